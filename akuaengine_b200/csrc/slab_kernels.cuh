// slab_kernels.cuh — device side of the x-slab multi-GPU decomposition (no reference counterpart: the reference is
// single-GPU). Keys are x-major, so after the sort a rank's particles are ordered by x-plane: the
// boundary planes it must send as ghosts are contiguous ranges of every SoA array (no pack kernels on the per-iteration
// path), and particles that left the slab carry a sentinel key that sorts them past the end of the owned range.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "pbf_kernels.cuh"
#include "radix_sort.cuh"

namespace akua {
namespace slab {

struct MigRecord { float4 pos, vel, xs; uint4 meta; };  // meta.x = particle id

// Destination of a particle from its (grid-relative, clamped) x cell: 0 = left neighbour, 1 = stays, 2 = right neighbour.
__device__ __forceinline__ int dest_of(uint32_t key, uint32_t planeCells, int xLo, int xHi) {
    int cx = (int)(key / planeCells);
    return cx < xLo ? 0 : (cx >= xHi ? 2 : 1);
}

// Pass 1: per-CTA counts of leavers in each direction (deterministic compaction, no atomics).
__global__ void __launch_bounds__(256) k_mig_count(const uint32_t* __restrict__ keys, uint32_t n, uint32_t planeCells,
                                                   int xLo, int xHi, uint32_t* __restrict__ blockCnt /*[2][blocks]*/) {
    __shared__ uint32_t sL[8], sR[8];
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    int d = i < n ? dest_of(keys[i], planeCells, xLo, xHi) : 1;
    uint32_t bl = __ballot_sync(0xffffffffu, d == 0), br = __ballot_sync(0xffffffffu, d == 2);
    int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (lane == 0) { sL[warp] = __popc(bl); sR[warp] = __popc(br); }
    __syncthreads();
    if (threadIdx.x == 0) {
        uint32_t a = 0, b = 0;
        for (int w = 0; w < 8; w++) { a += sL[w]; b += sR[w]; }
        blockCnt[blockIdx.x] = a;
        blockCnt[gridDim.x + blockIdx.x] = b;
    }
}
// Pass 2 (one CTA): exclusive scan of the per-CTA counts in place; totals -> counts[0] (left), counts[1] (right).
__global__ void __launch_bounds__(256) k_mig_scan(uint32_t* __restrict__ blockCnt, uint32_t blocks, uint32_t* __restrict__ counts) {
    __shared__ uint32_t s8[8];
    for (int dir = 0; dir < 2; dir++) {
        uint32_t* row = blockCnt + (size_t)dir * blocks;
        uint32_t carry = 0;
        for (uint32_t base = 0; base < blocks; base += 256) {
            uint32_t idx = base + threadIdx.x;
            uint32_t v = idx < blocks ? row[idx] : 0u, tot;
            uint32_t ex = rsort::block_excl_scan_256(v, s8, &tot);
            if (idx < blocks) row[idx] = carry + ex;
            carry += tot;
        }
        if (threadIdx.x == 0) counts[dir] = carry;
        __syncthreads();
    }
}
// Pass 3: leavers are copied (in index order) into the send buffers and get the sentinel key, which sorts them past the
// owned range so the reorder drops them.
__global__ void __launch_bounds__(256) k_mig_pack(uint32_t* __restrict__ keys, uint32_t n, uint32_t planeCells, int xLo,
                                                  int xHi, const uint32_t* __restrict__ blockOff, uint32_t sentinel,
                                                  const float4* __restrict__ pos, const float4* __restrict__ vel,
                                                  const float4* __restrict__ xs, const uint32_t* __restrict__ id,
                                                  MigRecord* __restrict__ sendL, MigRecord* __restrict__ sendR,
                                                  uint32_t cap) {
    __shared__ uint32_t sL[8], sR[8];
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    int d = i < n ? dest_of(keys[i], planeCells, xLo, xHi) : 1;
    uint32_t bl = __ballot_sync(0xffffffffu, d == 0), br = __ballot_sync(0xffffffffu, d == 2);
    int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (lane == 0) { sL[warp] = __popc(bl); sR[warp] = __popc(br); }
    __syncthreads();
    if (d == 1) return;
    uint32_t lt = (1u << lane) - 1;
    uint32_t off = 0;
    for (int w = 0; w < warp; w++) off += (d == 0 ? sL[w] : sR[w]);
    off += __popc((d == 0 ? bl : br) & lt);
    uint32_t slot = (d == 0 ? blockOff[blockIdx.x] : blockOff[gridDim.x + blockIdx.x]) + off;
    if (slot < cap) {
        MigRecord r;
        r.pos = pos[i]; r.vel = vel[i]; r.xs = xs[i]; r.meta = make_uint4(id[i], 0, 0, 0);
        (d == 0 ? sendL : sendR)[slot] = r;
    }
    keys[i] = sentinel;
}
// Arrivals are appended after the resident particles (before the sort) and keyed like everyone else.
__global__ void __launch_bounds__(256) k_mig_unpack(const MigRecord* __restrict__ recv, uint32_t count, uint32_t base,
                                                    float4* __restrict__ pos, float4* __restrict__ vel,
                                                    float4* __restrict__ xs, uint32_t* __restrict__ id,
                                                    uint32_t* __restrict__ keys, GridParams G) {
    uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= count) return;
    MigRecord r = recv[t];
    uint32_t i = base + t;
    pos[i] = r.pos; vel[i] = r.vel; xs[i] = r.xs; id[i] = r.meta.x;
    keys[i] = linear_key(cell_of(r.xs.x, r.xs.y, r.xs.z, G.cellSize), G);
}
// counts[2] = number of owned particles in the first owned x-plane, counts[3] = in the last one (binary searches in the
// sorted keys; two threads).
__global__ void k_plane_counts(const uint32_t* __restrict__ keysSorted, uint32_t nOwn, uint32_t planeCells, int xLo, int xHi,
                               uint32_t* __restrict__ counts) {
    int t = threadIdx.x;
    if (t > 1) return;
    // first key of plane xLo+1 (t = 0) or of plane xHi-1 (t = 1)
    uint64_t bound = (uint64_t)(t == 0 ? (xLo + 1) : (xHi - 1)) * planeCells;
    uint32_t lo = 0, hi = nOwn;
    while (lo < hi) {
        uint32_t mid = (lo + hi) >> 1;
        if ((uint64_t)keysSorted[mid] < bound) lo = mid + 1; else hi = mid;
    }
    counts[2 + t] = t == 0 ? lo : nOwn - lo;
}
// Per-x-plane population of the owned (key-sorted) particles: hist[x] = #{ i : key_i / planeCells == x }, by two binary
// searches per plane (one thread per plane). Used to re-balance the slab boundaries.
__global__ void __launch_bounds__(256) k_plane_hist(const uint32_t* __restrict__ keysSorted, uint32_t nOwn, uint32_t planeCells,
                                                    int gx, unsigned long long* __restrict__ hist) {
    int x = blockIdx.x * blockDim.x + threadIdx.x;
    if (x >= gx) return;
    auto lower = [&](uint64_t bound) {
        uint32_t lo = 0, hi = nOwn;
        while (lo < hi) {
            uint32_t mid = (lo + hi) >> 1;
            if ((uint64_t)keysSorted[mid] < bound) lo = mid + 1; else hi = mid;
        }
        return lo;
    };
    hist[x] = (unsigned long long)(lower((uint64_t)(x + 1) * planeCells) - lower((uint64_t)x * planeCells));
}

// Cell ranges of a contiguous, already key-sorted block [begin, end) (ghost planes received from a neighbour).
__global__ void __launch_bounds__(256) k_ranges(const uint32_t* __restrict__ keysSorted, uint32_t begin, uint32_t end,
                                                uint2* __restrict__ cellRange) {
    uint32_t i = begin + blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= end) return;
    uint32_t k = keysSorted[i];
    if (i == begin || keysSorted[i - 1] != k) cellRange[k].x = i;
    if (i == end - 1 || keysSorted[i + 1] != k) cellRange[k].y = i + 1;
}

}  // namespace slab
}  // namespace akua
