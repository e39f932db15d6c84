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

// Pass 1: per-CTA counts of leavers in each direction (deterministic compaction, no atomics on the compaction path),
// plus the four plane populations that let every rank PREDICT its post-migration boundary-plane and ghost-plane sizes
// from one count exchange (extra[0] stayers in my first plane, [1] stayers in my last plane, [2] leavers to the left that
// land in the left rank's last plane, [3] leavers to the right that land in the right rank's first plane).
__global__ void __launch_bounds__(256) k_mig_count(const uint32_t* __restrict__ keys, uint32_t n, uint32_t planeCells,
                                                   int xLo, int xHi, uint32_t* __restrict__ blockCnt /*[2][blocks]*/,
                                                   uint32_t* __restrict__ extra /*[4], zeroed*/) {
    __shared__ uint32_t sL[8], sR[8], sE[4];
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (threadIdx.x < 4) sE[threadIdx.x] = 0;
    __syncthreads();
    int cx = i < n ? (int)(keys[i] / planeCells) : xLo + 1;
    int d = i < n ? (cx < xLo ? 0 : (cx >= xHi ? 2 : 1)) : 1;
    uint32_t bl = __ballot_sync(0xffffffffu, d == 0), br = __ballot_sync(0xffffffffu, d == 2);
    uint32_t b0 = __ballot_sync(0xffffffffu, i < n && d == 1 && cx == xLo);
    uint32_t b1 = __ballot_sync(0xffffffffu, i < n && d == 1 && cx == xHi - 1);
    uint32_t b2 = __ballot_sync(0xffffffffu, d == 0 && cx == xLo - 1);
    uint32_t b3 = __ballot_sync(0xffffffffu, d == 2 && cx == xHi);
    int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (lane == 0) {
        sL[warp] = __popc(bl); sR[warp] = __popc(br);
        if (b0) atomicAdd(&sE[0], (uint32_t)__popc(b0));
        if (b1) atomicAdd(&sE[1], (uint32_t)__popc(b1));
        if (b2) atomicAdd(&sE[2], (uint32_t)__popc(b2));
        if (b3) atomicAdd(&sE[3], (uint32_t)__popc(b3));
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        uint32_t a = 0, b = 0;
        for (int w = 0; w < 8; w++) { a += sL[w]; b += sR[w]; }
        blockCnt[blockIdx.x] = a;
        blockCnt[gridDim.x + blockIdx.x] = b;
    }
    if (threadIdx.x < 4 && sE[threadIdx.x]) atomicAdd(&extra[threadIdx.x], sE[threadIdx.x]);
}
// Pass 2 (one CTA of 1024 threads): exclusive scan of the per-CTA counts in place, both directions at once; totals ->
// counts[0] (left), counts[1] (right).
__global__ void __launch_bounds__(1024) k_mig_scan(uint32_t* __restrict__ blockCnt, uint32_t blocks, uint32_t* __restrict__ counts) {
    __shared__ uint32_t wsum[2][32];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    uint32_t carry0 = 0, carry1 = 0;
    for (uint32_t base = 0; base < blocks; base += 1024) {
        const uint32_t idx = base + tid;
        const uint32_t v0 = idx < blocks ? blockCnt[idx] : 0u, v1 = idx < blocks ? blockCnt[blocks + idx] : 0u;
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
        if (idx < blocks) { blockCnt[idx] = carry0 + off0 + i0 - v0; blockCnt[blocks + idx] = carry1 + off1 + i1 - v1; }
        // chunk totals = exclusive offset of the last warp + its inclusive sum
        __shared__ uint32_t tot[2];
        if (tid == 1023) { tot[0] = off0 + i0; tot[1] = off1 + i1; }
        __syncthreads();
        carry0 += tot[0]; carry1 += tot[1];
        __syncthreads();
    }
    if (tid == 0) { counts[0] = carry0; counts[1] = carry1; }
    // messages for the single count exchange: to the left rank {leavers, leavers landing in its last plane, my
    // first-plane stayers} at counts[16..18]; to the right rank the mirror image at counts[20..22]
    if (threadIdx.x == 0) {
        counts[16] = counts[0]; counts[17] = counts[4]; counts[18] = counts[2];
        counts[20] = counts[1]; counts[21] = counts[5]; counts[22] = counts[3];
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
// Verifies the predicted boundary-plane sizes against the sorted keys (binary searches; two threads): a mismatch sets
// the sticky error word counts[31], which the host sees at the next step's count exchange.
__global__ void k_plane_verify(const uint32_t* __restrict__ keysSorted, uint32_t nOwn, uint32_t planeCells, int xLo, int xHi,
                               uint32_t predictFirst, uint32_t predictLast, int hasL, int hasR, uint32_t* __restrict__ counts) {
    int t = threadIdx.x;
    if (t > 1) return;
    uint64_t bound = (uint64_t)(t == 0 ? (xLo + 1) : (xHi - 1)) * planeCells;
    uint32_t lo = 0, hi = nOwn;
    while (lo < hi) {
        uint32_t mid = (lo + hi) >> 1;
        if ((uint64_t)keysSorted[mid] < bound) lo = mid + 1; else hi = mid;
    }
    uint32_t actual = t == 0 ? lo : nOwn - lo;
    if (t == 0 && hasL && actual != predictFirst) counts[31] = 1;
    if (t == 1 && hasR && actual != predictLast) counts[31] = 1;
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

// ---- peer-to-peer signalling (CUDA IPC path) ----
// After a rank has copied its boundary plane straight into the neighbour's ghost region (peer-mapped memory over NVLink),
// it publishes the exchange's epoch in the neighbour's flag word; the neighbour's stream runs k_wait_flags before the
// kernel that reads the ghosts. Epochs only grow, so "flag >= epoch" (wrap-safe) is the wait condition. The wait is
// bounded: on timeout it raises the sticky error word instead of hanging the GPU.
__global__ void k_signal_flag(uint32_t* __restrict__ peerFlag, uint32_t epoch) {
    __threadfence_system();
    *(volatile uint32_t*)peerFlag = epoch;
    __threadfence_system();
}
// CUDA-IPC transport of the per-step count message: the three counters for each neighbour (assembled by k_mig_scan at
// counts[16..18] / [20..22]) are stored straight into the neighbour's counter block (the slots it reads them from:
// [28..30] of the left rank = "from my right", [24..26] of the right rank = "from my left"), then the message epoch is
// published in the neighbour's flag word. The migration records were stored into the neighbour's inbox by k_mig_pack
// earlier in the same stream, so one flag covers both. One thread.
__global__ void k_publish_counts(const uint32_t* __restrict__ counts, uint32_t* __restrict__ peerCountsL,
                                 uint32_t* __restrict__ peerCountsR, uint32_t* __restrict__ peerFlagL,
                                 uint32_t* __restrict__ peerFlagR, uint32_t epoch) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    __threadfence_system();
    if (peerCountsL) for (int k = 0; k < 3; k++) ((volatile uint32_t*)peerCountsL)[28 + k] = counts[16 + k];
    if (peerCountsR) for (int k = 0; k < 3; k++) ((volatile uint32_t*)peerCountsR)[24 + k] = counts[20 + k];
    __threadfence_system();
    if (peerFlagL) *(volatile uint32_t*)peerFlagL = epoch;
    if (peerFlagR) *(volatile uint32_t*)peerFlagR = epoch;
    __threadfence_system();
}
__global__ void k_wait_flags(const uint32_t* __restrict__ flags, int waitL, int waitR, uint32_t epoch,
                             uint32_t* __restrict__ errWord, long long timeoutCycles) {
    const long long t0 = clock64();
    for (int side = 0; side < 2; side++) {
        if (!(side == 0 ? waitL : waitR)) continue;
        const volatile uint32_t* f = flags + side;
        while ((int32_t)(*f - epoch) < 0) {
            if (clock64() - t0 > timeoutCycles) { *errWord = 2; return; }
            __nanosleep(200);
        }
    }
    __threadfence_system();
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
