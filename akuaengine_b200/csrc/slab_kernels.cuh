// slab_kernels.cuh — device side of the x-slab multi-GPU decomposition (no reference counterpart: the reference is
// single-GPU). Keys are x-major, so after the sort a rank's particles are ordered by x-plane: the
// boundary planes it must send as ghosts are contiguous ranges of every SoA array (no pack kernels on the per-iteration
// path), and particles that left the slab carry a sentinel key that sorts them past the end of the owned range.
//
// Every size of a step (owned count, arrivals, leavers, boundary / ghost plane sizes) lives in the device-resident `dims`
// block (pbf_kernels.cuh: D_*): the kernels below compute and consume them on the device, grids come from a host estimate
// and every kernel loops, so the step needs no host synchronisation and is replayed as a CUDA graph.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "pbf_kernels.cuh"
#include "radix_sort.cuh"

namespace akua {
namespace slab {

// A migrating particle: solver state + the render payload (Particle::color / ::size follow the particle like they do in the
// reference, which sorts whole structs: src/CUDA/NeighbourSearchCUDA.cu:167-170). meta = (global id, size bits, 0, 0).
struct MigRecord { float4 pos, vel, xs, color; uint4 meta; };

// A migration tile is 2048 consecutive particles: thread t of the CTA handles the 8 particles [8 t, 8 t + 8) of its tile (two
// 16-byte key loads), so leavers are compacted in index order — deterministic, no atomics on the compaction path — while the
// per-tile bookkeeping (one count per tile and direction, scanned by a single CTA) is 8 x smaller than with one particle per thread.
constexpr int kMigItems = 8;
constexpr int kMigTile = 256 * kMigItems;
enum : int { D_FREE_TOP = 49, D_FREE_POP = 50 };   // payload-slot free stack: entries in use; pop base of this step's arrivals

// Destination of a particle from its (slab-local, clamped) x cell: 0 = left neighbour, 1 = stays, 2 = right neighbour.
__device__ __forceinline__ int dest_of(uint32_t key, uint32_t planeCells, int xLo, int xHi) {
    int cx = (int)(key / planeCells);
    return cx < xLo ? 0 : (cx >= xHi ? 2 : 1);
}
// The (up to) 8 keys of thread `threadIdx.x` in tile `tile`; slots past n read as a stayer of an interior plane.
__device__ __forceinline__ void mig_load_keys(const uint32_t* __restrict__ keys, uint32_t n, uint32_t base, uint32_t stayKey,
                                              uint32_t (&k)[kMigItems]) {
    if (base + kMigItems <= n) {
        const uint4 a = *reinterpret_cast<const uint4*>(keys + base), b = *reinterpret_cast<const uint4*>(keys + base + 4);
        k[0] = a.x; k[1] = a.y; k[2] = a.z; k[3] = a.w; k[4] = b.x; k[5] = b.y; k[6] = b.z; k[7] = b.w;
    } else {
#pragma unroll
        for (int r = 0; r < kMigItems; r++) k[r] = base + r < n ? keys[base + r] : stayKey;
    }
}

// Pass 1: per-tile counts of leavers in each direction, plus the four plane populations that let every rank PREDICT its
// post-migration boundary-plane and ghost-plane sizes from one count exchange (extra[0] stayers in my first plane, [1]
// stayers in my last plane, [2] leavers to the left that land in the left rank's last plane, [3] leavers to the right that
// land in the right rank's first plane).
__global__ void __launch_bounds__(256) k_mig_count(const uint32_t* __restrict__ keys, const uint32_t* __restrict__ dims,
                                                   uint32_t planeCells,
                                                   uint32_t* __restrict__ blockCnt /*[2][tileStride]*/, uint32_t tileStride,
                                                   uint32_t* __restrict__ extra /*[4], zeroed*/) {
    __shared__ uint32_t sAcc[6];
    const uint32_t n = dims[D_N];
    const int xLo = (int)dims[D_XLO], xHi = (int)dims[D_XHI];
    const uint32_t numTiles = (n + kMigTile - 1) / kMigTile;
    const uint32_t stayKey = (uint32_t)(xLo + 1) * planeCells;   // an interior plane when the slab has one; never a leaver
    uint32_t tot[4] = {0, 0, 0, 0};                              // this thread's share of the four plane populations
    for (uint32_t tile = blockIdx.x; tile < numTiles; tile += gridDim.x) {
        if (threadIdx.x < 2) sAcc[threadIdx.x] = 0;
        __syncthreads();
        uint32_t k[kMigItems];
        const uint32_t base = tile * kMigTile + threadIdx.x * kMigItems;
        mig_load_keys(keys, n, base, stayKey, k);
        uint32_t cl = 0, cr = 0;
#pragma unroll
        for (int r = 0; r < kMigItems; r++) {
            const bool live = base + r < n;
            const int cx = (int)(k[r] / planeCells);
            const int d = cx < xLo ? 0 : (cx >= xHi ? 2 : 1);
            cl += live && d == 0; cr += live && d == 2;
            tot[0] += live && d == 1 && cx == xLo; tot[1] += live && d == 1 && cx == xHi - 1;
            tot[2] += live && d == 0 && cx == xLo - 1; tot[3] += live && d == 2 && cx == xHi;
        }
        cl = __reduce_add_sync(0xffffffffu, cl); cr = __reduce_add_sync(0xffffffffu, cr);
        if ((threadIdx.x & 31) == 0) { if (cl) atomicAdd(&sAcc[0], cl); if (cr) atomicAdd(&sAcc[1], cr); }
        __syncthreads();
        if (threadIdx.x == 0) { blockCnt[tile] = sAcc[0]; blockCnt[tileStride + tile] = sAcc[1]; }
        __syncthreads();
    }
    if (threadIdx.x < 4) sAcc[2 + threadIdx.x] = 0;
    __syncthreads();
#pragma unroll
    for (int q = 0; q < 4; q++) {
        const uint32_t v = __reduce_add_sync(0xffffffffu, tot[q]);
        if ((threadIdx.x & 31) == 0 && v) atomicAdd(&sAcc[2 + q], v);
    }
    __syncthreads();
    if (threadIdx.x < 4 && sAcc[2 + threadIdx.x]) atomicAdd(&extra[threadIdx.x], sAcc[2 + threadIdx.x]);
}
// Pass 2 (one CTA of 1024 threads): exclusive scan of the per-tile counts in place, both directions at once; totals ->
// dims[D_OUT_L], dims[D_OUT_R]; also assembles the two count messages.
__global__ void __launch_bounds__(1024) k_mig_scan(uint32_t* __restrict__ blockCnt, const uint32_t* __restrict__ nPtr,
                                                   uint32_t tileStride, uint32_t* __restrict__ counts) {
    __shared__ uint32_t wsum[2][32];
    __shared__ uint32_t tot[2];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t blocks = (*nPtr + kMigTile - 1) / kMigTile;
    uint32_t carry0 = 0, carry1 = 0;
    for (uint32_t base = 0; base < blocks; base += 1024) {
        const uint32_t idx = base + tid;
        const uint32_t v0 = idx < blocks ? blockCnt[idx] : 0u, v1 = idx < blocks ? blockCnt[tileStride + idx] : 0u;
        uint32_t i0 = v0, i1 = v1;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            uint32_t t0 = __shfl_up_sync(0xffffffffu, i0, o), t1 = __shfl_up_sync(0xffffffffu, i1, o);
            if (lane >= o) { i0 += t0; i1 += t1; }
        }
        if (lane == 31) { wsum[0][warp] = i0; wsum[1][warp] = i1; }
        __syncthreads();
        if (warp == 0) {
            uint32_t a = wsum[0][lane], b = wsum[1][lane], ia = a, ib = b;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                uint32_t t0 = __shfl_up_sync(0xffffffffu, ia, o), t1 = __shfl_up_sync(0xffffffffu, ib, o);
                if (lane >= o) { ia += t0; ib += t1; }
            }
            wsum[0][lane] = ia - a; wsum[1][lane] = ib - b;   // exclusive warp offsets
        }
        __syncthreads();
        const uint32_t off0 = wsum[0][warp], off1 = wsum[1][warp];
        if (idx < blocks) { blockCnt[idx] = carry0 + off0 + i0 - v0; blockCnt[tileStride + idx] = carry1 + off1 + i1 - v1; }
        // chunk totals = exclusive offset of the last warp + its inclusive sum
        if (tid == 1023) { tot[0] = off0 + i0; tot[1] = off1 + i1; }
        __syncthreads();
        carry0 += tot[0]; carry1 += tot[1];
        __syncthreads();
    }
    // messages for the single count exchange: to the left rank {leavers, leavers landing in its last plane, my
    // first-plane stayers}; to the right rank the mirror image
    if (tid == 0) {
        counts[D_OUT_L] = carry0; counts[D_OUT_R] = carry1;
        counts[D_MSG_TO_L] = carry0; counts[D_MSG_TO_L + 1] = counts[D_LAND_L]; counts[D_MSG_TO_L + 2] = counts[D_STAY_FIRST];
        counts[D_MSG_TO_R] = carry1; counts[D_MSG_TO_R + 1] = counts[D_LAND_R]; counts[D_MSG_TO_R + 2] = counts[D_STAY_LAST];
    }
}
// Pass 3: leavers are copied (in index order) into the send buffers — with the CUDA-IPC transport straight into the
// neighbour's inbox over NVLink — and get the sentinel key, which sorts them past the owned range so the reorder drops
// them. Their payload slots go back on the free stack (in send-buffer order: deterministic).
__global__ void __launch_bounds__(256) k_mig_pack(uint32_t* __restrict__ keys, const uint32_t* __restrict__ dims,
                                                  uint32_t planeCells, const uint32_t* __restrict__ blockOff,
                                                  uint32_t tileStride, uint32_t sentinel,
                                                  const float4* __restrict__ pos, const float4* __restrict__ vel,
                                                  const float4* __restrict__ xs, const uint32_t* __restrict__ id,
                                                  const uint32_t* __restrict__ slot, const float4* __restrict__ color,
                                                  const float* __restrict__ size, uint32_t* __restrict__ freeSlots,
                                                  MigRecord* __restrict__ sendL, MigRecord* __restrict__ sendR,
                                                  uint32_t cap) {
    __shared__ uint32_t sL[8], sR[8];
    const uint32_t n = dims[D_N];
    const int xLo = (int)dims[D_XLO], xHi = (int)dims[D_XHI];
    const uint32_t numTiles = (n + kMigTile - 1) / kMigTile;
    const uint32_t outL = min(dims[D_OUT_L], cap), freeTop = dims[D_FREE_TOP];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t stayKey = (uint32_t)(xLo + 1) * planeCells;
    for (uint32_t tile = blockIdx.x; tile < numTiles; tile += gridDim.x) {
        uint32_t k[kMigItems];
        const uint32_t base = tile * kMigTile + threadIdx.x * kMigItems;
        mig_load_keys(keys, n, base, stayKey, k);
        uint32_t cl = 0, cr = 0;
#pragma unroll
        for (int r = 0; r < kMigItems; r++) {
            const int d = base + r < n ? dest_of(k[r], planeCells, xLo, xHi) : 1;
            cl += d == 0; cr += d == 2;
        }
        // exclusive prefix of (cl, cr) over the threads of the tile: index order = (thread, item) order
        uint32_t il = cl, ir = cr;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t tl = __shfl_up_sync(0xffffffffu, il, o), tr = __shfl_up_sync(0xffffffffu, ir, o);
            if (lane >= o) { il += tl; ir += tr; }
        }
        if (lane == 31) { sL[warp] = il; sR[warp] = ir; }
        __syncthreads();
        if (sL[0] + sL[1] + sL[2] + sL[3] + sL[4] + sL[5] + sL[6] + sL[7] + sR[0] + sR[1] + sR[2] + sR[3] + sR[4] + sR[5] + sR[6] + sR[7]) {
            uint32_t offL = il - cl, offR = ir - cr;
            for (int w = 0; w < warp; w++) { offL += sL[w]; offR += sR[w]; }
            offL += blockOff[tile]; offR += blockOff[tileStride + tile];
            if (cl | cr) {
#pragma unroll 1
                for (int r = 0; r < kMigItems; r++) {
                    const uint32_t i = base + r;
                    const int d = i < n ? dest_of(k[r], planeCells, xLo, xHi) : 1;
                    if (d == 1) continue;
                    const uint32_t dst = d == 0 ? offL++ : offR++;
                    if (dst < cap) {
                        const uint32_t ps = slot[i];
                        MigRecord rec;
                        rec.pos = pos[i]; rec.vel = vel[i]; rec.xs = xs[i]; rec.color = color[ps];
                        rec.meta = make_uint4(id[i], __float_as_uint(size[ps]), 0, 0);
                        (d == 0 ? sendL : sendR)[dst] = rec;
                        freeSlots[freeTop + (d == 0 ? dst : outL + dst)] = ps;
                    }
                    keys[i] = sentinel;
                }
            }
        }
        __syncthreads();
    }
}

// The step's plan, computed on the device by one thread once both neighbours' count messages are in (CUDA-IPC transport:
// bounded in-kernel wait on the message epochs; NCCL transport: the messages were received before the launch). Every later
// kernel of the step reads its sizes from here. All sizes are clamped to the buffers they index (memory safety) and any
// clamp raises a sticky error bit the host reports at its next look.
struct PlanCaps {
    uint32_t packCap;        // leavers per direction that fit the outboxes / the neighbours' inboxes
    uint32_t migCap;         // arrivals per direction that fit this rank's inboxes
    uint32_t ownedCap;       // owned + arriving particles that fit below the ghost regions
    uint32_t planeCapL, planeCapR;   // boundary-plane particles that fit the left / right neighbour's ghost region
    uint32_t ghostCap;       // ghost particles per side that fit this rank's ghost regions
    uint32_t slotCap;        // payload slots
};
// `pub` (CUDA-IPC transport): the kernel first publishes this rank's own count message — the three counters for each
// neighbour (assembled by k_mig_scan) are stored straight into the neighbour's dims block (D_MSG_FROM_R of the left rank,
// D_MSG_FROM_L of the right rank), then the message epoch goes into the neighbour's flag word; the migration records were
// stored into the neighbour's inbox by k_mig_pack earlier in the same stream, so one flag covers both — and then waits for
// the neighbours' messages.
struct PlanPublish { uint32_t *peerDimsL = nullptr, *peerDimsR = nullptr, *peerFlagL = nullptr, *peerFlagR = nullptr; };
__global__ void k_slab_plan(uint32_t* __restrict__ dims, const uint32_t* __restrict__ countFlags /* [0] left, [1] right */,
                            int hasL, int hasR, int waitFlags, PlanCaps caps, long long timeoutCycles, PlanPublish pub) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    uint32_t err = 0;
    if (pub.peerFlagL || pub.peerFlagR) {
        const uint32_t epoch = dims[D_EPOCH] + 1u;
        __threadfence_system();   // the migration records k_mig_pack stored into the neighbours' inboxes come first
        if (pub.peerDimsL) for (int k = 0; k < 3; k++) ((volatile uint32_t*)pub.peerDimsL)[D_MSG_FROM_R + k] = dims[D_MSG_TO_L + k];
        if (pub.peerDimsR) for (int k = 0; k < 3; k++) ((volatile uint32_t*)pub.peerDimsR)[D_MSG_FROM_L + k] = dims[D_MSG_TO_R + k];
        __threadfence_system();
        if (pub.peerFlagL) *(volatile uint32_t*)pub.peerFlagL = epoch;
        if (pub.peerFlagR) *(volatile uint32_t*)pub.peerFlagR = epoch;
        __threadfence_system();
    }
    dims[D_PLAN_WAIT_NS] = 0u;
    if (waitFlags) {
        const uint32_t epoch = dims[D_EPOCH] + 1u;
        const unsigned long long t0 = global_timer_ns();
        if (hasL && !spin_until(countFlags + 0, epoch, timeoutCycles, dims + D_ERROR)) err |= SLAB_ERR_TIMEOUT;
        if (hasR && !(err & SLAB_ERR_TIMEOUT) && !spin_until(countFlags + 1, epoch, timeoutCycles, dims + D_ERROR)) err |= SLAB_ERR_TIMEOUT;
        __threadfence_system();
        // a rank that is ahead of its neighbours idles HERE once per step: the measured idle time is what the re-balancing
        // feeds back on (akua_pbf_rebalance)
        const unsigned long long waited = global_timer_ns() - t0;
        dims[D_PLAN_WAIT_NS] = waited > 0xffffffffull ? 0xffffffffu : (uint32_t)waited;
        *reinterpret_cast<unsigned long long*>(dims + D_STAT_PLAN_WAIT_NS) += waited;
    }
    const volatile uint32_t* d = dims;
    uint32_t outL = d[D_OUT_L], outR = d[D_OUT_R];
    if (outL > caps.packCap) { outL = caps.packCap; err |= SLAB_ERR_MIG_OVERFLOW; }
    if (outR > caps.packCap) { outR = caps.packCap; err |= SLAB_ERR_MIG_OVERFLOW; }
    uint32_t inL = hasL ? d[D_MSG_FROM_L] : 0u, inR = hasR ? d[D_MSG_FROM_R] : 0u;
    if (err & SLAB_ERR_TIMEOUT) inL = inR = 0;    // nothing trustworthy arrived
    if (inL > caps.migCap) { inL = caps.migCap; err |= SLAB_ERR_MIG_OVERFLOW; }
    if (inR > caps.migCap) { inR = caps.migCap; err |= SLAB_ERR_MIG_OVERFLOW; }
    const uint32_t n = d[D_N];
    uint32_t room = caps.ownedCap > n ? caps.ownedCap - n : 0u;
    const uint32_t freeTop = min(d[D_FREE_TOP] + outL + outR, caps.slotCap);
    room = min(room, freeTop);
    if (inL > room) { inL = room; err |= SLAB_ERR_CAPACITY; }
    if (inR > room - inL) { inR = room - inL; err |= SLAB_ERR_CAPACITY; }
    const uint32_t nPre = n + inL + inR;
    const uint32_t nOwn = nPre - outL - outR;
    // Post-migration plane sizes, known before the sort: my boundary planes = stayers + arrivals that land in them; a
    // neighbour's facing plane (= my ghosts) = its stayers there + my leavers that land there. (Arrivals from the far side
    // cannot reach the near plane: slabs are >= 2 planes wide and the stepper moves particles by << one slab.)
    uint32_t planeL = hasL ? d[D_STAY_FIRST] + d[D_MSG_FROM_L + 1] : 0u;
    uint32_t planeR = hasR ? d[D_STAY_LAST] + d[D_MSG_FROM_R + 1] : 0u;
    uint32_t ghostL = hasL ? d[D_MSG_FROM_L + 2] + d[D_LAND_L] : 0u;
    uint32_t ghostR = hasR ? d[D_MSG_FROM_R + 2] + d[D_LAND_R] : 0u;
    if (err & SLAB_ERR_TIMEOUT) ghostL = ghostR = 0;
    if (planeL > caps.planeCapL) { planeL = caps.planeCapL; err |= SLAB_ERR_GHOST_OVERFLOW; }
    if (planeR > caps.planeCapR) { planeR = caps.planeCapR; err |= SLAB_ERR_GHOST_OVERFLOW; }
    if (ghostL > caps.ghostCap) { ghostL = caps.ghostCap; err |= SLAB_ERR_GHOST_OVERFLOW; }
    if (ghostR > caps.ghostCap) { ghostR = caps.ghostCap; err |= SLAB_ERR_GHOST_OVERFLOW; }
    if (planeL > nOwn) planeL = nOwn;
    if (planeR > nOwn) planeR = nOwn;
    dims[D_OUT_L] = outL; dims[D_OUT_R] = outR;
    dims[D_IN_L] = inL; dims[D_IN_R] = inR; dims[D_NPRE] = nPre; dims[D_NOWN] = nOwn;
    dims[D_PLANE_L] = planeL; dims[D_PLANE_R] = planeR; dims[D_GHOST_L] = ghostL; dims[D_GHOST_R] = ghostR;
    dims[D_FREE_POP] = freeTop; dims[D_FREE_TOP] = freeTop - inL - inR;
    if (err) atomicOr(dims + D_ERROR, err);
}

// Arrivals (left inbox first, then right) are appended after the resident particles (before the sort), keyed like everyone
// else, and take a payload slot from the free stack. A record whose x* does not lie in this rank's planes (it crossed more
// than a slab in one step) is filed under the nearest owned plane; the next step's migration moves it on.
__global__ void __launch_bounds__(256) k_mig_unpack(const MigRecord* __restrict__ recvL, const MigRecord* __restrict__ recvR,
                                                    const uint32_t* __restrict__ dims,
                                                    float4* __restrict__ pos, float4* __restrict__ vel,
                                                    float4* __restrict__ xs, uint32_t* __restrict__ id,
                                                    uint32_t* __restrict__ slot, float4* __restrict__ color,
                                                    float* __restrict__ size, const uint32_t* __restrict__ freeSlots,
                                                    uint32_t* __restrict__ keys, GridParams G) {
    const int xLo = (int)dims[D_XLO], xHi = (int)dims[D_XHI];
    const uint32_t inL = dims[D_IN_L], count = inL + dims[D_IN_R], base = dims[D_N], popBase = dims[D_FREE_POP];
    for (uint32_t t = blockIdx.x * blockDim.x + threadIdx.x; t < count; t += gridDim.x * blockDim.x) {
        const MigRecord r = t < inL ? recvL[t] : recvR[t - inL];
        const uint32_t i = base + t;
        const uint32_t ps = freeSlots[popBase - 1u - t];
        pos[i] = r.pos; vel[i] = r.vel; xs[i] = r.xs; id[i] = r.meta.x;
        slot[i] = ps; color[ps] = r.color; size[ps] = __uint_as_float(r.meta.y);
        const int3 c = cell_of(r.xs.x, r.xs.y, r.xs.z, G.cellSize);
        const int x = clampi(c.x - G.gridMin.x, xLo, xHi - 1);
        const int y = clampi(c.y - G.gridMin.y, 0, G.gridDim.y - 1);
        const int z = clampi(c.z - G.gridMin.z, 0, G.gridDim.z - 1);
        keys[i] = (uint32_t)((x * G.gridDim.y + y) * G.gridDim.z + z);
    }
}
// Verifies the predicted boundary-plane sizes against the sorted keys (binary searches; two threads): a mismatch raises the
// sticky error word.
struct PlaneVerify { const uint32_t* keysSorted = nullptr; uint32_t planeCells = 0; int hasL = 0, hasR = 0; uint32_t* dims = nullptr; };
__device__ __forceinline__ void plane_verify(const uint32_t* __restrict__ keysSorted, uint32_t planeCells,
                                             int hasL, int hasR, uint32_t* __restrict__ dims) {
    int t = threadIdx.x;
    if (t > 1) return;
    const int xLo = (int)dims[D_XLO], xHi = (int)dims[D_XHI];
    const uint32_t nOwn = dims[D_NOWN];
    uint64_t bound = (uint64_t)(t == 0 ? (xLo + 1) : (xHi - 1)) * planeCells;
    uint32_t lo = 0, hi = nOwn;
    while (lo < hi) {
        uint32_t mid = (lo + hi) >> 1;
        if ((uint64_t)keysSorted[mid] < bound) lo = mid + 1; else hi = mid;
    }
    uint32_t actual = t == 0 ? lo : nOwn - lo;
    if (t == 0 && hasL && actual != dims[D_PLANE_L]) atomicOr(dims + D_ERROR, (uint32_t)SLAB_ERR_PLANE_PREDICTION);
    if (t == 1 && hasR && actual != dims[D_PLANE_R]) atomicOr(dims + D_ERROR, (uint32_t)SLAB_ERR_PLANE_PREDICTION);
}
__global__ void k_plane_verify(const uint32_t* __restrict__ keysSorted, uint32_t planeCells,
                               int hasL, int hasR, uint32_t* __restrict__ dims) {
    plane_verify(keysSorted, planeCells, hasL, hasR, dims);
}
// Per-x-plane population and WORK of the owned (key-sorted) particles, one CTA per plane: two binary searches give the plane's
// index range, count[planeOffset + x] = its size, work[planeOffset + x] = sum over it of (kWorkBase + the LARGEST neighbour count
// among the 32 consecutive particles it shares a warp with) — what a particle costs the sweeps, whose warps run as long as their
// busiest lane (a free surface or spray has few neighbours on average but ragged counts: it is dearer than its mean suggests).
// Used to re-balance the slab boundaries (a sloshing tank is denser, hence dearer, on one side).
constexpr uint32_t kWorkBase = 12;
__global__ void __launch_bounds__(256) k_plane_hist(const uint32_t* __restrict__ keysSorted, const uint32_t* __restrict__ nbrCount,
                                                    const uint32_t* __restrict__ nOwnPtr, uint32_t planeCells, int gx, int planeOffset,
                                                    unsigned long long* __restrict__ count, unsigned long long* __restrict__ work) {
    __shared__ uint32_t range[2];
    __shared__ unsigned long long part[8];
    const uint32_t nOwn = *nOwnPtr;
    const int x = blockIdx.x;
    if (x >= gx) return;
    if (threadIdx.x < 2) {
        const uint64_t bound = (uint64_t)(x + threadIdx.x) * planeCells;
        uint32_t lo = 0, hi = nOwn;
        while (lo < hi) {
            uint32_t mid = (lo + hi) >> 1;
            if ((uint64_t)keysSorted[mid] < bound) lo = mid + 1; else hi = mid;
        }
        range[threadIdx.x] = lo;
    }
    __syncthreads();
    const uint32_t a = range[0], b = range[1];
    unsigned long long w = 0;
    for (uint32_t i0 = a; i0 < b; i0 += blockDim.x) {        // uniform trip count: the warp-wide maximum needs every lane
        const uint32_t i = i0 + threadIdx.x;
        const uint32_t c = i < b ? nbrCount[i] : 0u;
        const uint32_t m = __reduce_max_sync(0xffffffffu, c);
        if (i < b) w += kWorkBase + m;
    }
    for (int o = 16; o > 0; o >>= 1) w += __shfl_down_sync(0xffffffffu, w, o);
    if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = w;
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int k = 1; k < 8; k++) w += part[k];
        count[planeOffset + x] = (unsigned long long)(b - a);
        work[planeOffset + x] = w;
    }
}

// ---- fused halo push (CUDA IPC path) ----
// Copies this rank's first / last boundary plane of `src` straight into the neighbours' ghost regions with P2P stores over
// NVLink; the last CTA publishes the exchange's epoch (halo_signal). Used where no sweep produces the planes (x* after the
// reorder; v when solverIterations == 0) — the sweeps push their own results (PeerPush in pbf_kernels.cuh).
// `pv` (optional): CTA 0 also checks the predicted plane sizes against the sorted keys (plane_verify) — the stand-alone
// k_plane_verify launch of the NCCL path folded into this one.
template <typename T>
__global__ void __launch_bounds__(256) k_push_planes(const T* __restrict__ src, PeerPush pp, HaloSync hs, PlaneVerify pv) {
    pdl_wait();
    if (pv.keysSorted && blockIdx.x == 0) plane_verify(pv.keysSorted, pv.planeCells, pv.hasL, pv.hasR, pv.dims);
    resolve_push(pp);
    const uint32_t n = pp.dims[D_NOWN];
    const uint32_t nL = pp.dstL ? pp.nL : 0u, nR = (pp.dstR && pp.startR <= n) ? n - pp.startR : 0u;
    for (uint32_t t = blockIdx.x * blockDim.x + threadIdx.x; t < nL + nR; t += gridDim.x * blockDim.x) {
        if (t < nL) static_cast<T*>(pp.dstL)[t] = src[t];
        else static_cast<T*>(pp.dstR)[t - nL] = src[pp.startR + (t - nL)];
    }
    halo_signal(hs, gridDim.x);
}
// Ghost planes received from the neighbours (already key-sorted: one x plane each, ordered like this rank's cells): keys from
// the received x*, and the (start, end) range of every ghost cell. Waits in-kernel for the x* exchange (CUDA IPC path).
__global__ void __launch_bounds__(256) k_ghost_ranges(const float4* __restrict__ xs, uint32_t* __restrict__ keysSorted,
                                                      const uint32_t* __restrict__ dims, uint32_t ghostBaseL,
                                                      uint32_t ghostBaseR, uint2* __restrict__ cellRange, GridParams G,
                                                      HaloSync hs) {
    pdl_wait();
    halo_wait(hs);
    const uint32_t gL = dims[D_GHOST_L], gR = dims[D_GHOST_R];
    for (uint32_t t = blockIdx.x * blockDim.x + threadIdx.x; t < gL + gR; t += gridDim.x * blockDim.x) {
        const bool left = t < gL;
        const uint32_t local = left ? t : t - gL, cnt = left ? gL : gR;
        const uint32_t j = (left ? ghostBaseL : ghostBaseR) + local;
        const float4 x = xs[j];
        const uint32_t k = linear_key(cell_of(x.x, x.y, x.z, G.cellSize), G);
        keysSorted[j] = k;
        bool first = local == 0, last = local + 1 == cnt;
        if (!first) { const float4 a = xs[j - 1]; first = linear_key(cell_of(a.x, a.y, a.z, G.cellSize), G) != k; }
        if (!last) { const float4 b = xs[j + 1]; last = linear_key(cell_of(b.x, b.y, b.z, G.cellSize), G) != k; }
        if (first) cellRange[k].x = j;
        if (last) cellRange[k].y = j + 1;
    }
}
// Closes a step on the device: the owned count carries over, the epoch base advances past this step's exchanges, the
// statistics accumulate. `planeBytes` = bytes pushed per boundary-plane particle over the whole step.
__global__ void k_step_end(uint32_t* __restrict__ dims, uint32_t exchanges, uint32_t planeBytes) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    dims[D_N] = dims[D_NOWN];
    dims[D_EPOCH] += exchanges;
    unsigned long long* st = reinterpret_cast<unsigned long long*>(dims + D_STAT_MIG_IN);
    st[0] += dims[D_IN_L] + dims[D_IN_R];
    st[1] += dims[D_OUT_L] + dims[D_OUT_R];
    st[2] += (unsigned long long)(dims[D_PLANE_L] + dims[D_PLANE_R]) * planeBytes
           + (unsigned long long)(dims[D_OUT_L] + dims[D_OUT_R]) * sizeof(MigRecord);
    dims[D_STEPS] += 1;
    // busy time of this rank: step to step on the GPU's clock, minus the idle wait for the neighbours' count messages
    unsigned long long* tl = reinterpret_cast<unsigned long long*>(dims + D_T_LAST);
    unsigned long long* busy = reinterpret_cast<unsigned long long*>(dims + D_BUSY_NS);
    const unsigned long long now = global_timer_ns();
    if (*tl != 0ull && now > *tl) {
        const unsigned long long dt = now - *tl, w = dims[D_PLAN_WAIT_NS];
        if (dt < 2000000000ull) { *busy += dt > w ? dt - w : 0ull; dims[D_BUSY_STEPS] += 1; }   // (a gap of seconds is the host, not a step)
    }
    *tl = now;
}
// Re-balancing: exports this rank's measured busy time per step (ns), the sum of its per-plane work and its current lower
// bound (global plane; 0 for the first rank), and restarts the busy-time accumulation. One CTA.
__global__ void __launch_bounds__(256) k_rank_busy(const unsigned long long* __restrict__ work, int gx, uint32_t* __restrict__ dims,
                                                   unsigned long long* __restrict__ busyOut, unsigned long long* __restrict__ workOut,
                                                   int lowerBound, int firstRank, unsigned long long* __restrict__ boundOut) {
    __shared__ unsigned long long part[8];
    unsigned long long w = 0;
    for (int x = threadIdx.x; x < gx; x += blockDim.x) w += work[x];
    for (int o = 16; o > 0; o >>= 1) w += __shfl_down_sync(0xffffffffu, w, o);
    if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = w;
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int k = 1; k < 8; k++) w += part[k];
        const unsigned long long busy = *reinterpret_cast<const unsigned long long*>(dims + D_BUSY_NS);
        const uint32_t steps = dims[D_BUSY_STEPS];
        *busyOut = steps ? busy / steps : 0ull;
        *workOut = w;
        *boundOut = firstRank ? 0ull : (unsigned long long)lowerBound;
        *reinterpret_cast<unsigned long long*>(dims + D_BUSY_NS) = 0ull;
        *reinterpret_cast<unsigned long long*>(dims + D_T_LAST) = 0ull;   // the host-side part of this call is not a step
        dims[D_BUSY_STEPS] = 0u;
    }
}
// (Re)initialises the device-side bookkeeping after an upload: owned count, identity payload slots, full free stack.
__global__ void __launch_bounds__(256) k_slab_reset(uint32_t* __restrict__ dims, uint32_t n, uint32_t cap,
                                                    uint32_t* __restrict__ slot, uint32_t* __restrict__ freeSlots) {
    const uint32_t t0 = blockIdx.x * blockDim.x + threadIdx.x;
    for (uint32_t t = t0; t < cap; t += gridDim.x * blockDim.x) {
        if (t < n) slot[t] = t;
        if (t < cap - n) freeSlots[t] = cap - 1u - t;   // popping yields n, n + 1, ...
    }
    if (t0 == 0) {
        dims[D_N] = n; dims[D_NOWN] = n; dims[D_NPRE] = n; dims[D_FREE_TOP] = cap - n; dims[D_FREE_POP] = cap - n;
        // an upload interrupts the step-to-step clock of the busy-time measurement
        dims[D_T_LAST] = 0u; dims[D_T_LAST + 1] = 0u; dims[D_BUSY_NS] = 0u; dims[D_BUSY_NS + 1] = 0u; dims[D_BUSY_STEPS] = 0u;
    }
}

}  // namespace slab
}  // namespace akua
