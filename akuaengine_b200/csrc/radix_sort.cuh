// radix_sort.cuh — hand-written stable LSD radix sort of (key32, index32) pairs for sm_100a.
//
// Replaces thrust::sort of whole 108-byte Particle structs by hash (reference: src/CUDA/NeighbourSearchCUDA.cu:167-170,
// a CUB merge sort moving 216 B per particle per merge level) with a sort of 8-byte pairs followed by one gather.
// Stable, so equal keys keep their input order — the same order the reference's (de-facto stable) merge sort leaves.
//
// The number of keys may live on the device (`nPtr`, x-slab mode: the host only knows an estimate): grids are sized from the
// host's value and every kernel loops over the tiles of the actual count; tileHist rows have a fixed stride.
// One pass per 8-bit digit, three launches per pass, no spin-waits between CTAs (nothing here can hang):
//   count   : one CTA per 4096-key tile -> 256-bin histogram, written bin-major  tileHist[bin][tile]
//   scan    : one CTA per bin           -> exclusive scan of that bin's row over tiles, row total -> binTotal[bin]
//   scatter : one CTA per tile          -> stable in-tile ranks (warp match + per-warp counters), keys/indices staged in
//                                          shared memory in digit order, then written out as contiguous runs
// HBM traffic per pass: count reads 4 B, scatter reads 8 B (4 B on the first pass: indices are implicit) and writes 8 B
// per key; the histogram rows are ~1/16 of a key each.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <algorithm>

namespace akua {
namespace rsort {

constexpr int kThreads = 256;
constexpr int kWarps = kThreads / 32;
// Keys per thread is a template parameter: 16 (4096-key tiles) for large inputs, 4 (1024-key tiles) below ~8 M keys so
// that the count / scatter grids still fill 148 SMs several times over (at 1 M keys a 4096-key tiling is only 245 CTAs).
constexpr int kItemsLarge = 16, kItemsSmall = 4;

#ifdef AKUA_HOST_EMU   // tests/emu (CPU emulation of the kernels, test infrastructure only)
__device__ __forceinline__ void pdl_wait() {}
#else
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }   // see pbf_kernels.cuh
#endif

__device__ __forceinline__ uint32_t lanemask_lt() {
    uint32_t m;
#ifdef AKUA_HOST_EMU
    m = (1u << (threadIdx.x & 31u)) - 1u;
#else
    asm("mov.u32 %0, %%lanemask_lt;" : "=r"(m));
#endif
    return m;
}

template <int kItems>
__global__ void __launch_bounds__(kThreads) k_count(const uint32_t* __restrict__ keys, uint32_t n,
                                                    const uint32_t* __restrict__ nPtr, int shift,
                                                    uint32_t* __restrict__ tileHist, uint32_t tileStride) {
    pdl_wait();
    constexpr int kTile = kThreads * kItems;
    __shared__ uint32_t hist[256];
    const int tid = threadIdx.x, lane = tid & 31;
    if (nPtr) n = *nPtr;
    const uint32_t numTiles = (n + kTile - 1) / kTile;
    for (uint32_t tile = blockIdx.x; tile < numTiles; tile += gridDim.x) {
        hist[tid] = 0;
        __syncthreads();
        const uint32_t base = tile * (uint32_t)kTile;
#pragma unroll 4
        for (int r = 0; r < kItems; r++) {
            uint32_t idx = base + r * kThreads + tid;
            bool valid = idx < n;
            uint32_t d = valid ? ((keys[idx] >> shift) & 255u) : 256u;
            uint32_t peers = __match_any_sync(0xffffffffu, d);
            if (valid && lane == __ffs(peers) - 1) atomicAdd(&hist[d], (uint32_t)__popc(peers));
        }
        __syncthreads();
        tileHist[(size_t)tid * tileStride + tile] = hist[tid];
    }
}

// Block-wide exclusive scan helper for 256 threads. Returns the exclusive prefix; *total gets the block sum.
__device__ __forceinline__ uint32_t block_excl_scan_256(uint32_t v, uint32_t* smem8, uint32_t* total) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    uint32_t incl = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        uint32_t t = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += t;
    }
    if (lane == 31) smem8[warp] = incl;
    __syncthreads();
    uint32_t warpOff = 0, tot = 0;
#pragma unroll
    for (int w = 0; w < kWarps; w++) {
        uint32_t s = smem8[w];
        if (w < warp) warpOff += s;
        tot += s;
    }
    __syncthreads();  // smem8 may be reused by the caller
    *total = tot;
    return incl - v + warpOff;
}

__global__ void __launch_bounds__(kThreads) k_scan(uint32_t* __restrict__ tileHist, uint32_t n,
                                                   const uint32_t* __restrict__ nPtr, uint32_t tileSize,
                                                   uint32_t tileStride, uint32_t* __restrict__ binTotal) {
    pdl_wait();
    __shared__ uint32_t s8[kWarps];
    if (nPtr) n = *nPtr;
    const uint32_t numTiles = (n + tileSize - 1) / tileSize;
    uint32_t* row = tileHist + (size_t)blockIdx.x * tileStride;
    uint32_t carry = 0;
    for (uint32_t base = 0; base < numTiles; base += kThreads) {
        uint32_t idx = base + threadIdx.x;
        uint32_t v = idx < numTiles ? row[idx] : 0u;
        uint32_t tot;
        uint32_t ex = block_excl_scan_256(v, s8, &tot);
        if (idx < numTiles) row[idx] = carry + ex;
        carry += tot;
    }
    if (threadIdx.x == 0) binTotal[blockIdx.x] = carry;
}

template <bool FIRST, int kItems>  // FIRST: input indices are implicit (idx itself)
__global__ void __launch_bounds__(kThreads) k_scatter(const uint32_t* __restrict__ keysIn,
                                                      const uint32_t* __restrict__ valsIn,
                                                      uint32_t* __restrict__ keysOut, uint32_t* __restrict__ valsOut,
                                                      uint32_t n, const uint32_t* __restrict__ nPtr, int shift,
                                                      const uint32_t* __restrict__ tileHist,
                                                      uint32_t tileStride, const uint32_t* __restrict__ binTotal) {
    pdl_wait();
    constexpr int kTile = kThreads * kItems;
    constexpr int kWarpSpan = 32 * kItems;  // keys handled by one warp (contiguous -> stability)
    __shared__ uint32_t warpCnt[kWarps][256];  // per-warp running digit counts, then exclusive warp prefixes
    __shared__ uint32_t tileBase[256];         // exclusive prefix of this tile's digit counts (slot of a digit's run in smem)
    __shared__ uint32_t globalBase[256];       // where this tile's run of each digit starts in the output
    __shared__ uint32_t s8[kWarps];
    __shared__ uint32_t stageKey[kTile];
    __shared__ uint32_t stageVal[kTile];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (nPtr) n = *nPtr;
    const uint32_t numTiles = (n + kTile - 1) / kTile;
    uint32_t binBase;
    {
        uint32_t tot;
        binBase = block_excl_scan_256(binTotal[tid], s8, &tot);  // syncs inside
    }
    for (uint32_t tile = blockIdx.x; tile < numTiles; tile += gridDim.x) {
    for (int i = tid; i < kWarps * 256; i += kThreads) (&warpCnt[0][0])[i] = 0;
    globalBase[tid] = binBase + tileHist[(size_t)tid * tileStride + tile];
    __syncthreads();

    uint32_t key[kItems], rank[kItems];
    const uint32_t wbase = tile * (uint32_t)kTile + warp * (uint32_t)kWarpSpan;
    const uint32_t lt = lanemask_lt();
#pragma unroll
    for (int r = 0; r < kItems; r++) {
        uint32_t idx = wbase + r * 32 + lane;
        bool valid = idx < n;
        key[r] = valid ? keysIn[idx] : 0xffffffffu;
        uint32_t d = valid ? ((key[r] >> shift) & 255u) : 256u;
        uint32_t peers = __match_any_sync(0xffffffffu, d);
        uint32_t pre = valid ? warpCnt[warp][d] : 0u;
        __syncwarp();
        if (valid && lane == __ffs(peers) - 1) warpCnt[warp][d] = pre + __popc(peers);
        __syncwarp();
        rank[r] = pre + __popc(peers & lt);
    }
    __syncthreads();
    // thread `tid` owns digit `tid`: exclusive prefix over warps, digit total for the tile
    uint32_t digitTotal = 0;
#pragma unroll
    for (int w = 0; w < kWarps; w++) {
        uint32_t c = warpCnt[w][tid];
        warpCnt[w][tid] = digitTotal;
        digitTotal += c;
    }
    {
        uint32_t tot;
        uint32_t ex = block_excl_scan_256(digitTotal, s8, &tot);  // syncs inside
        tileBase[tid] = ex;
    }
    __syncthreads();
    // stage in digit order (stable)
#pragma unroll
    for (int r = 0; r < kItems; r++) {
        uint32_t idx = wbase + r * 32 + lane;
        if (idx < n) {
            uint32_t d = (key[r] >> shift) & 255u;
            uint32_t slot = tileBase[d] + warpCnt[warp][d] + rank[r];
            stageKey[slot] = key[r];
            stageVal[slot] = FIRST ? idx : valsIn[idx];
        }
    }
    __syncthreads();
    // contiguous runs out: slot s holds digit d(s); its destination is globalBase[d] + (s - tileBase[d])
    const uint32_t tileBeg = tile * (uint32_t)kTile;
    const uint32_t tileCount = min((uint32_t)kTile, n - tileBeg);
#pragma unroll 4
    for (int r = 0; r < kItems; r++) {
        uint32_t s = r * kThreads + tid;
        if (s < tileCount) {
            uint32_t k = stageKey[s];
            uint32_t d = (k >> shift) & 255u;
            uint32_t dst = globalBase[d] + (s - tileBase[d]);
            keysOut[dst] = k;
            valsOut[dst] = stageVal[s];
        }
    }
    __syncthreads();   // the staging arrays are reused by the next tile of this CTA
    }
}

// ------------------------------------------------------------------------------------------------ one-sweep variant
// One read of the keys builds the digit histograms of ALL passes (k_hist); each pass is then ONE kernel (k_onesweep) that
// ranks a tile, publishes the tile's digit counts, obtains its global offsets by decoupled look-back over the preceding tiles
// (Merrill & Garland's single-pass scan; tiles are handed out by an atomic ticket, so every tile a CTA waits for belongs to a
// CTA that is already running) and scatters. HBM traffic 4 + 16 p bytes per key instead of 4 p + 16 p, and 1 + p launches
// instead of 3 p. Stability is by construction: ranks inside a tile follow (warp, item, lane) order = input order, tiles are
// ordered by the look-back.
constexpr uint32_t kLbAggregate = 1u << 30, kLbInclusive = 2u << 30, kLbCountMask = (1u << 30) - 1u, kLbFlagMask = 3u << 30;
constexpr int kMaxPasses = 4;

template <int kItems>
__global__ void __launch_bounds__(kThreads) k_hist(const uint32_t* __restrict__ keys, uint32_t n, const uint32_t* __restrict__ nPtr,
                                                   int passes, uint32_t* __restrict__ gHist /*[kMaxPasses][256]*/) {
    pdl_wait();
    constexpr int kTile = kThreads * kItems;
    __shared__ uint32_t hist[kMaxPasses][256];
    const int tid = threadIdx.x, lane = tid & 31;
    for (int q = 0; q < kMaxPasses; q++) hist[q][tid] = 0;
    __syncthreads();
    if (nPtr) n = *nPtr;
    const uint32_t numTiles = (n + kTile - 1) / kTile;
    for (uint32_t tile = blockIdx.x; tile < numTiles; tile += gridDim.x) {
        const uint32_t base = tile * (uint32_t)kTile;
#pragma unroll 2
        for (int r = 0; r < kItems; r++) {
            const uint32_t idx = base + r * kThreads + tid;
            const bool valid = idx < n;
            const uint32_t key = valid ? keys[idx] : 0u;
            for (int q = 0; q < passes; q++) {
                // warp-aggregated: neighbouring keys share their upper digits (the input is nearly sorted from the last step)
                const uint32_t d = valid ? ((key >> (8 * q)) & 255u) : 256u;
                const uint32_t peers = __match_any_sync(0xffffffffu, d);
                if (valid && lane == __ffs(peers) - 1) atomicAdd(&hist[q][d], (uint32_t)__popc(peers));
            }
        }
    }
    __syncthreads();
    for (int q = 0; q < passes; q++)
        if (hist[q][tid]) atomicAdd(&gHist[q * 256 + tid], hist[q][tid]);
}

__device__ __forceinline__ uint32_t ld_volatile_u32(const uint32_t* p) { return *reinterpret_cast<const volatile uint32_t*>(p); }
__device__ __forceinline__ void st_volatile_u32(uint32_t* p, uint32_t v) { *reinterpret_cast<volatile uint32_t*>(p) = v; }

template <bool FIRST, int kItems>  // FIRST: input indices are implicit (idx itself)
__global__ void __launch_bounds__(kThreads) k_onesweep(const uint32_t* __restrict__ keysIn, const uint32_t* __restrict__ valsIn,
                                                       uint32_t* __restrict__ keysOut, uint32_t* __restrict__ valsOut,
                                                       uint32_t n, const uint32_t* __restrict__ nPtr, int shift,
                                                       const uint32_t* __restrict__ gHistPass /*[256]*/,
                                                       uint32_t* __restrict__ ticket, uint32_t* __restrict__ lookback /*[tiles][256]*/) {
    pdl_wait();
    constexpr int kTile = kThreads * kItems;
    constexpr int kWarpSpan = 32 * kItems;  // keys handled by one warp (contiguous -> stability)
    __shared__ uint32_t warpCnt[kWarps][256];  // per-warp running digit counts, then exclusive warp prefixes
    __shared__ uint32_t tileBase[256];         // exclusive prefix of this tile's digit counts (slot of a digit's run in smem)
    __shared__ uint32_t globalBase[256];       // where this tile's run of each digit starts in the output
    __shared__ uint32_t s8[kWarps];
    __shared__ uint32_t stageKey[kTile];
    __shared__ uint32_t stageVal[kTile];
    __shared__ uint32_t sTile;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (nPtr) n = *nPtr;
    const uint32_t numTiles = (n + kTile - 1) / kTile;
    uint32_t binBase;
    {
        uint32_t tot;
        binBase = block_excl_scan_256(gHistPass[tid], s8, &tot);  // first output slot of digit `tid`; syncs inside
    }
    const uint32_t lt = lanemask_lt();
    for (;;) {
        if (tid == 0) sTile = atomicAdd(ticket, 1u);
        for (int i = tid; i < kWarps * 256; i += kThreads) (&warpCnt[0][0])[i] = 0;
        __syncthreads();
        const uint32_t tile = sTile;
        if (tile >= numTiles) break;

        uint32_t key[kItems], rank[kItems];
        const uint32_t wbase = tile * (uint32_t)kTile + warp * (uint32_t)kWarpSpan;
#pragma unroll
        for (int r = 0; r < kItems; r++) {
            const uint32_t idx = wbase + r * 32 + lane;
            const bool valid = idx < n;
            key[r] = valid ? keysIn[idx] : 0xffffffffu;
            const uint32_t d = valid ? ((key[r] >> shift) & 255u) : 256u;
            const uint32_t peers = __match_any_sync(0xffffffffu, d);
            const uint32_t pre = valid ? warpCnt[warp][d] : 0u;
            __syncwarp();
            if (valid && lane == __ffs(peers) - 1) warpCnt[warp][d] = pre + __popc(peers);
            __syncwarp();
            rank[r] = pre + __popc(peers & lt);
        }
        __syncthreads();
        // thread `tid` owns digit `tid`: exclusive prefix over warps, digit total for the tile
        uint32_t digitTotal = 0;
#pragma unroll
        for (int w = 0; w < kWarps; w++) {
            const uint32_t c = warpCnt[w][tid];
            warpCnt[w][tid] = digitTotal;
            digitTotal += c;
        }
        // decoupled look-back: publish this tile's count of digit `tid`, sum the preceding tiles' counts back to the nearest
        // inclusive prefix, publish this tile's inclusive prefix
        uint32_t* mine = lookback + (size_t)tile * 256 + tid;
        uint32_t excl = 0;
        if (tile == 0) {
            st_volatile_u32(mine, digitTotal | kLbInclusive);
        } else {
            st_volatile_u32(mine, digitTotal | kLbAggregate);
            for (uint32_t t = tile - 1;; t--) {
                const uint32_t* q = lookback + (size_t)t * 256 + tid;
                uint32_t v;
                do { v = ld_volatile_u32(q); } while ((v & kLbFlagMask) == 0u);
                excl += v & kLbCountMask;
                if (v & kLbInclusive) break;     // tile 0 always publishes an inclusive prefix: t never runs below 0
            }
            st_volatile_u32(mine, (excl + digitTotal) | kLbInclusive);
        }
        globalBase[tid] = binBase + excl;
        {
            uint32_t tot;
            const uint32_t ex = block_excl_scan_256(digitTotal, s8, &tot);  // syncs inside
            tileBase[tid] = ex;
        }
        __syncthreads();
        // stage in digit order (stable)
#pragma unroll
        for (int r = 0; r < kItems; r++) {
            const uint32_t idx = wbase + r * 32 + lane;
            if (idx < n) {
                const uint32_t d = (key[r] >> shift) & 255u;
                const uint32_t slot = tileBase[d] + warpCnt[warp][d] + rank[r];
                stageKey[slot] = key[r];
                stageVal[slot] = FIRST ? idx : valsIn[idx];
            }
        }
        __syncthreads();
        // contiguous runs out: slot s holds digit d(s); its destination is globalBase[d] + (s - tileBase[d])
        const uint32_t tileBeg = tile * (uint32_t)kTile;
        const uint32_t tileCount = min((uint32_t)kTile, n - tileBeg);
#pragma unroll 4
        for (int r = 0; r < kItems; r++) {
            const uint32_t sl = r * kThreads + tid;
            if (sl < tileCount) {
                const uint32_t k = stageKey[sl];
                const uint32_t d = (k >> shift) & 255u;
                const uint32_t dst = globalBase[d] + (sl - tileBase[d]);
                keysOut[dst] = k;
                valsOut[dst] = stageVal[sl];
            }
        }
        __syncthreads();   // staging arrays, warpCnt and sTile are reused by the next tile of this CTA
    }
}

struct Workspace {
    uint32_t* tileHist = nullptr;  // three-kernel variant: [256][maxTiles]; one-sweep variant: look-back words [passes][maxTiles][256]
    uint32_t* binTotal = nullptr;  // three-kernel variant: [256]; one-sweep variant: [kMaxPasses][256] histograms + kMaxPasses tickets
    uint32_t maxTiles = 0;
    int mode = 2;                  // 0 = three kernels per pass (count / scan / scatter), 1 = one-sweep, 2 = by size (default)
    int items = 0;                 // keys per thread of the one-sweep tiles (0 = by size)
};
constexpr size_t kCtrlWords = (size_t)kMaxPasses * 256 + 8;   // binTotal allocation
inline size_t tile_hist_words(uint32_t maxTiles) { return (size_t)256 * maxTiles * kMaxPasses; }

constexpr uint64_t kSmallLimit = 8u << 20;
inline int items_for(uint64_t n) { return n < kSmallLimit ? kItemsSmall : kItemsLarge; }
inline uint32_t tiles_for(uint64_t n) { uint64_t t = (uint64_t)kThreads * items_for(n); return (uint32_t)((n + t - 1) / t); }
// workspace must hold either tiling of any n <= capacity (with a device-side count the host picks the tiling from an
// estimate, so the small tiling can meet a count up to the capacity)
inline uint32_t max_tiles_for_capacity(uint64_t cap) {
    uint64_t a = (cap + (uint64_t)kThreads * kItemsSmall - 1) / ((uint64_t)kThreads * kItemsSmall);
    return (uint32_t)std::max<uint64_t>(a, 1);
}
inline int passes_for_bits(int bits) { return bits <= 0 ? 1 : (bits + 7) / 8; }

// Sorts n (key, index) pairs. keysIn is preserved. Results land in (*keysOut, *valsOut), which point into bufA or bufB.
// Returns the number of kernel launches issued.
// `pdl`: launch with programmatic stream serialization (every kernel here starts with pdl_wait()).
template <typename... KArgs, typename... Args>
inline void launch(bool pdl, void (*kernel)(KArgs...), uint32_t grid, cudaStream_t st, Args... args) {
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(grid); cfg.blockDim = dim3(kThreads); cfg.stream = st;
    cudaLaunchAttribute at{};
    if (pdl) {
        at.id = cudaLaunchAttributeProgrammaticStreamSerialization;
        at.val.programmaticStreamSerializationAllowed = 1;
        cfg.attrs = &at; cfg.numAttrs = 1;
    }
    cudaLaunchKernelEx(&cfg, kernel, KArgs(args)...);
}

// `nPtr` (optional): the actual number of pairs lives on the device; `n` is then the host's estimate (grid sizing only).
// `valsIn0` (optional): the values that travel with keysIn (default: the indices 0..n-1). Must not alias bufA / bufB.
inline int sort_pairs_three_kernel(const uint32_t* keysIn, uint32_t* keyA, uint32_t* valA, uint32_t* keyB, uint32_t* valB,
                      uint32_t n, int keyBits, const Workspace& ws, cudaStream_t st, uint32_t** keysOut,
                      uint32_t** valsOut, bool pdl = false, const uint32_t* nPtr = nullptr, const uint32_t* valsIn0 = nullptr) {
    const uint32_t numTiles = std::max(1u, tiles_for(n));
    const bool small = items_for(n) == kItemsSmall;
    const uint32_t tileSize = (uint32_t)kThreads * (small ? kItemsSmall : kItemsLarge);
    const uint32_t tileStride = ws.maxTiles;
    const int passes = passes_for_bits(keyBits);
    const uint32_t* kin = keysIn;
    const uint32_t* vin = valsIn0;
    uint32_t* kout = keyA;
    uint32_t* vout = valA;
    int launches = 0;
    for (int p = 0; p < passes; p++) {
        int shift = 8 * p;
        if (small) launch(pdl, k_count<kItemsSmall>, numTiles, st, kin, n, nPtr, shift, ws.tileHist, tileStride);
        else       launch(pdl, k_count<kItemsLarge>, numTiles, st, kin, n, nPtr, shift, ws.tileHist, tileStride);
        launch(pdl, k_scan, 256u, st, ws.tileHist, n, nPtr, tileSize, tileStride, ws.binTotal);
#define AK_SCATTER(F, I) launch(pdl, k_scatter<F, I>, numTiles, st, kin, vin, kout, vout, n, nPtr, shift, ws.tileHist, tileStride, ws.binTotal)
        if (p == 0 && !valsIn0) { if (small) AK_SCATTER(true, kItemsSmall); else AK_SCATTER(true, kItemsLarge); }
        else        { if (small) AK_SCATTER(false, kItemsSmall); else AK_SCATTER(false, kItemsLarge); }
#undef AK_SCATTER
        launches += 3;
        kin = kout;
        vin = vout;
        if (kout == keyA) { kout = keyB; vout = valB; } else { kout = keyA; vout = valA; }
    }
    *keysOut = const_cast<uint32_t*>(kin);
    *valsOut = const_cast<uint32_t*>(vin);
    return launches;
}

inline int onesweep_items_for(uint64_t n, int forced) {
    if (forced == 4 || forced == 8 || forced == 16) return forced;
    return n < (2u << 20) ? 4 : 16;   // measured on a B200 at 8 M keys: 16 keys per thread 249 us, 8: 275 us, 4: 380 us
}
// Below this many keys the three-kernel passes win (1 M keys: 56 us against 80 us — the one-sweep variant pays a histogram
// kernel, its look-back memsets and tile-serial look-back latency that small inputs cannot amortise); above it one-sweep does
// (8 M keys: 249 us against 283 us). profiles/r02_c5_sort_variants.txt
constexpr uint64_t kOnesweepMin = 4u << 20;
inline int sort_pairs_onesweep(const uint32_t* keysIn, uint32_t* keyA, uint32_t* valA, uint32_t* keyB, uint32_t* valB,
                      uint32_t n, int keyBits, const Workspace& ws, cudaStream_t st, uint32_t** keysOut,
                      uint32_t** valsOut, bool pdl, const uint32_t* nPtr, const uint32_t* valsIn0 = nullptr) {
    const int passes = std::min(passes_for_bits(keyBits), kMaxPasses);
    const int items = onesweep_items_for(n, ws.items);
    const uint32_t tileSize = (uint32_t)kThreads * items;
    const uint32_t numTiles = std::max<uint32_t>(1u, (uint32_t)(((uint64_t)n + tileSize - 1) / tileSize));
    // look-back words of the tiles that can exist: exact on one GPU, every tile of the capacity with a device-side count
    const uint32_t zeroTiles = nPtr ? (uint32_t)std::min<uint64_t>(ws.maxTiles, ((uint64_t)ws.maxTiles * kThreads * kItemsSmall + tileSize - 1) / tileSize) : numTiles;
    uint32_t* gHist = ws.binTotal;
    uint32_t* tickets = ws.binTotal + (size_t)kMaxPasses * 256;
    cudaMemsetAsync(ws.binTotal, 0, kCtrlWords * sizeof(uint32_t), st);
    for (int p = 0; p < passes; p++)
        cudaMemsetAsync(ws.tileHist + (size_t)p * ws.maxTiles * 256, 0, (size_t)zeroTiles * 256 * sizeof(uint32_t), st);
    const uint32_t histGrid = std::min<uint32_t>(numTiles, 148u * 8u);
    if (items == 4) launch(pdl, k_hist<4>, histGrid, st, keysIn, n, nPtr, passes, gHist);
    else if (items == 8) launch(pdl, k_hist<8>, histGrid, st, keysIn, n, nPtr, passes, gHist);
    else launch(pdl, k_hist<16>, histGrid, st, keysIn, n, nPtr, passes, gHist);
    const uint32_t* kin = keysIn;
    const uint32_t* vin = valsIn0;
    uint32_t* kout = keyA;
    uint32_t* vout = valA;
    int launches = 1;
    for (int p = 0; p < passes; p++) {
        const int shift = 8 * p;
        uint32_t* lb = ws.tileHist + (size_t)p * ws.maxTiles * 256;
#define AK_SWEEP(F, I) launch(pdl, k_onesweep<F, I>, numTiles, st, kin, vin, kout, vout, n, nPtr, shift, (const uint32_t*)(gHist + p * 256), tickets + p, lb)
#define AK_SWEEP_I(F) do { if (items == 4) AK_SWEEP(F, 4); else if (items == 8) AK_SWEEP(F, 8); else AK_SWEEP(F, 16); } while (0)
        if (p == 0 && !valsIn0) AK_SWEEP_I(true); else AK_SWEEP_I(false);
#undef AK_SWEEP_I
#undef AK_SWEEP
        launches++;
        kin = kout;
        vin = vout;
        if (kout == keyA) { kout = keyB; vout = valB; } else { kout = keyA; vout = valA; }
    }
    *keysOut = const_cast<uint32_t*>(kin);
    *valsOut = const_cast<uint32_t*>(vin);
    return launches;
}

inline int sort_pairs(const uint32_t* keysIn, uint32_t* keyA, uint32_t* valA, uint32_t* keyB, uint32_t* valB,
                      uint32_t n, int keyBits, const Workspace& ws, cudaStream_t st, uint32_t** keysOut,
                      uint32_t** valsOut, bool pdl = false, const uint32_t* nPtr = nullptr, const uint32_t* valsIn0 = nullptr) {
    if ((ws.mode == 1 || (ws.mode == 2 && n >= kOnesweepMin)) && passes_for_bits(keyBits) <= kMaxPasses)
        return sort_pairs_onesweep(keysIn, keyA, valA, keyB, valB, n, keyBits, ws, st, keysOut, valsOut, pdl, nPtr, valsIn0);
    return sort_pairs_three_kernel(keysIn, keyA, valA, keyB, valB, n, keyBits, ws, st, keysOut, valsOut, pdl, nPtr, valsIn0);
}

}  // namespace rsort
}  // namespace akua
