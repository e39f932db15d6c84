// list_build.cuh — two-phase ("mask") neighbour-list build for LINEAR_CELL keys: an opt-in alternative to
// k_build_neighbours (pbf_kernels.cuh), selected with akua_pbf_options::list_build / AKUA_LIST_BUILD.
//
// Replaces kernel_find_neighbours (src/CUDA/NeighbourSearchCUDA.cu:72-130) exactly like k_build_neighbours does: same
// candidates in the same order (rows dx -> dy, sorted order inside a row), strict d2 < h*h with the reference's contraction
// (dist2), self skipped, capped at maxNeighbours — so the lists are bit-identical to k_build_neighbours' lists.
//
// Why a second variant. ncu on k_build_neighbours at 1 M particles (profiles/r01_ncu_final_summary.txt): 178 M warp
// instructions (~26 per candidate), 62 % long-scoreboard stalls on the single candidate load in flight, and the L1 data
// stage at 80-86 % — about half of those wavefronts come from the ~200 divergent 4-byte list stores a warp executes (one
// per candidate iteration in which ANY lane has a hit). Here a row of candidates is processed in chunks of up to 32:
//   phase 1  tests the chunk with independent loads (UNROLL in flight per thread) and only records a hit bitmask —
//            no store, no dependence on the running neighbour count, ~10 instructions per candidate;
//   phase 2  walks the set bits (about three per row) and appends them — the divergent stores shrink from one per
//            candidate iteration to one per hit rank (~45 per warp instead of ~200).
// STATUS: written without GPU access (round 1, session 3). The per-particle function is __host__ __device__ and is checked
// on the CPU against an independent scan (tests/test_list_build_host.py); on the GPU it is covered by
// tests/test_parity_gpu.py::test_list_build_variants_identical. Not the default until it has been timed on a B200.
// Reachability culling (skipping neighbour cells whose nearest point is >= h away: ~24 % of the candidates) was measured on a
// B200 in round 2 and REJECTED: 242 us vs 193 us at 1 M particles, 918 vs 777 us at 4 M (profiles/r02_c3_list_build_culling_rejected.txt)
// — the per-row bookkeeping and the shorter, more ragged rows cost more than the saved distance tests.
#pragma once
#include "pbf_kernels.cuh"

namespace akua {

#ifdef __CUDA_ARCH__
#define AKUA_LD_RO(p) __ldg(p)
#define AKUA_FFS(m) __ffs((int)(m))
#else
#define AKUA_LD_RO(p) (*(p))
#define AKUA_FFS(m) __builtin_ffs((int)(m))
#endif
#define AKUA_OPAQUE(v) asm volatile("" : "+r"(v))

// Appends the neighbours of sorted particle i to its ELL list (list_slot layout) and returns their number.
// xs: predicted positions in sorted order (owned range, then ghost planes in slab mode); cellRange: (start, end) per cell.
template <int UNROLL>
__host__ __device__ __forceinline__ uint32_t build_list_mask(uint32_t i, const float4* __restrict__ xs,
                                                             const uint2* __restrict__ cellRange, uint32_t stride,
                                                             uint32_t maxN, uint32_t* __restrict__ list,
                                                             const GridParams& G, float h) {
    static_assert(UNROLL == 4 || UNROLL == 8, "chunk length 32 must be a multiple of UNROLL");
    const float4 xi = xs[i];
#ifdef __CUDA_ARCH__
    const float h2 = __fmul_rn(h, h);
#else
    const float h2 = h * h;
#endif
    const int3 c = cell_of(xi.x, xi.y, xi.z, G.lookupCellSize);
    const int cx = clampi(c.x - G.gridMin.x, 0, G.gridDim.x - 1);
    const int cy = clampi(c.y - G.gridMin.y, 0, G.gridDim.y - 1);
    const int cz = clampi(c.z - G.gridMin.z, 0, G.gridDim.z - 1);
    const int z0 = cz - 1 < 0 ? 0 : cz - 1, z1 = cz + 1 > G.gridDim.z - 1 ? G.gridDim.z - 1 : cz + 1;
    uint32_t count = 0;
    // write cursor: entry k of particle i lives at list_slot(i, k, stride); consecutive entries are adjacent inside a group
    // of four and a group step (4 * stride words) apart across groups
    uint32_t* wp = list + (size_t)i * 4;
    const size_t groupWrap = (size_t)stride * 4 - 3;
    for (int dx = -1; dx <= 1; dx++) {
        const int X = cx + dx;
        if (X < 0 || X >= G.gridDim.x) continue;
        for (int dy = -1; dy <= 1; dy++) {
            const int Y = cy + dy;
            if (Y < 0 || Y >= G.gridDim.y) continue;
            // the (up to) three z-adjacent cells are contiguous in sorted order: one row range
            const uint2* row = cellRange + ((size_t)X * G.gridDim.y + Y) * G.gridDim.z;
            uint32_t s = 0xffffffffu, e = 0;
            for (int z = z0; z <= z1; z++) {
                const uint2 r = AKUA_LD_RO(&row[z]);
                if (r.y > r.x) { s = s < r.x ? s : r.x; e = e > r.y ? e : r.y; }
            }
            for (uint32_t base = s; base < e && count < maxN; base += 32u) {
                const uint32_t len = e - base < 32u ? e - base : 32u;
                const float4* __restrict__ p = xs + base;
                // ---- phase 1: hit mask of the chunk; bit t <=> d2(i, base + t) < h2
                uint32_t mask = 0;
                uint32_t t = 0;
                for (; t + UNROLL <= len; t += UNROLL, p += UNROLL) {
                    float4 xj[UNROLL];
#pragma unroll
                    for (int u = 0; u < UNROLL; u++) xj[u] = AKUA_LD_RO(p + u);
                    uint32_t bits = 0;
#pragma unroll
                    for (int u = 0; u < UNROLL; u++) {
                        const float d2 = dist2(xi.x - xj[u].x, xi.y - xj[u].y, xi.z - xj[u].z);
                        bits |= d2 < h2 ? (1u << u) : 0u;
                    }
                    AKUA_OPAQUE(bits);      // keep the UNROLL constant-position bits together: ONE variable shift per group
                    mask |= bits << t;
                }
                for (; t < len; t++, p++) {
                    const float4 xj = AKUA_LD_RO(p);
                    const float d2 = dist2(xi.x - xj.x, xi.y - xj.y, xi.z - xj.z);
                    mask |= (d2 < h2 ? 1u : 0u) << t;
                }
                const uint32_t self = i - base;            // wraps to a huge value when i < base
                if (self < 32u) mask &= ~(1u << self);
                // ---- phase 2: append the hits in candidate order
                while (mask != 0u && count < maxN) {
                    const uint32_t b = (uint32_t)AKUA_FFS(mask) - 1u;
                    mask &= mask - 1u;
                    *wp = base + b;
                    wp += (count & 3u) == 3u ? groupWrap : (size_t)1;
                    count++;
                }
            }
        }
    }
    return count;
}

// Count word and the padding of the last group of four (the sweeps read whole groups; tail slots hold the particle's own
// index and are masked out by the count) — identical to the end of k_build_neighbours.
__host__ __device__ __forceinline__ void finish_list(uint32_t i, uint32_t count, uint32_t stride, uint32_t* __restrict__ list,
                                                     uint32_t* __restrict__ cnt) {
    cnt[i] = count;
    for (uint32_t k = count; k < ((count + 3u) & ~3u); k++) list[list_slot(i, k, stride)] = i;
}

#if defined(__CUDACC__) || defined(AKUA_HOST_EMU)
// Same parameter list as k_build_neighbours<KEY_LINEAR> so that the host code can launch either. MINB (resident CTAs per SM)
// fixes the register budget explicitly: <4, 5> compiles to 48 registers, <8, 4> to 64, both without spills (ptxas -v); left to
// its own heuristics ptxas squeezes the 8-deep variant into 48 registers and spills.
template <int UNROLL, int MINB, bool STRIDED>
__global__ void __launch_bounds__(256, MINB) k_build_neighbours_mask(const float4* __restrict__ xs,
                                                               const uint32_t* __restrict__ /*keysSorted*/,
                                                               const uint32_t* __restrict__ /*bucketStart*/,
                                                               const uint2* __restrict__ cellRange, uint32_t n,
                                                               uint32_t stride, uint32_t maxN, uint32_t* __restrict__ list,
                                                               uint32_t* __restrict__ cnt, GridParams G, float h,
                                                               const uint32_t* __restrict__ nPtr) {
    pdl_wait();
    if (STRIDED) n = live_count(n, nPtr);
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const uint32_t count = build_list_mask<UNROLL>(i, xs, cellRange, stride, maxN, list, G, h);
        if (!STRIDED) pdl_trigger();
        finish_list(i, count, stride, list, cnt);
        if (!STRIDED) break;
    }
    if (STRIDED) pdl_trigger();
}
#endif

}  // namespace akua
