// pbf_kernels.cuh — hand-written sm_100a kernels for the PBF step (device-resident SoA float4 state).
//
// Each kernel cites the reference kernel(s) it replaces (paths relative to the reference repository root). The
// reference's 13 kernels run one thread per particle over a 108-byte AoS with a per-particle neighbour list that
// round-trips through host memory; here the state is SoA float4 (x,y,z,mass) / (vx,vy,vz,density), the neighbour list is a
// column-major (ELL) device array in groups of four (entry k of particle i lives in uint4 number (k/4)*stride + i, lane
// k%4), so that one coalesced 16-byte load per thread fetches four neighbour indices, the four gathers they feed are
// issued together, and the next group is prefetched while the current one is evaluated. The per-iteration kernels
// are fused so that each neighbour's position is gathered once per sweep:
//   pass A = K5+K6  (density, both lambda loops)            pass B = K7+K8 (+K9+K10 on the last iteration)
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#ifndef AKUA_SWEEP_BLOCK
#define AKUA_SWEEP_BLOCK 128
#endif
#ifndef AKUA_SWEEP_MINBLOCKS
#define AKUA_SWEEP_MINBLOCKS 8
#endif
// The SLAB instantiations (x-slab mode: device-resolved span, halo wait / signal, peer pushes) get the same residency as the
// single-GPU ones: the sweeps are latency-bound gathers, one resident CTA fewer per SM costs ~10 % (measured, profiles/r02_c5_*);
// the default instantiations (fast math, packed gathers, n = 4) fit 64 registers without spills either way.
#define AKUA_SWEEP_BOUNDS __launch_bounds__(AKUA_SWEEP_BLOCK, AKUA_SWEEP_MINBLOCKS)

namespace akua {

enum : int { KEY_HASH = 0, KEY_LINEAR = 1 };

struct SphParams {
    float h, h2;
    float poly6Coef;    // 315/(64*3.14*h^9)   (SmoothingKernelsCUDA.h:20; pi is 3.14f in the reference)
    float spikyCoef;    // -45/(3.14*h^6)      (SmoothingKernelsCUDA.h:27)
    float selfW;        // poly6(0)
    float invRestDensity;
    float relaxation;
    float corrK, corrN, invPoly6Dq;  // artificial pressure: -k * (W(d2)/W(dq^2))^n
    int corrNIsFour;
    float uniformMass;  // packed-gather sweeps only: the one mass every particle has (checked at upload)
};

struct GridParams {
    float cellSize;     // K2 cell size (the reference passes smoothRadius, NeighbourSearchCUDA.cu:163)
    float lookupCellSize;  // K4 cell size (spatialHashCellSize, :177); equal to cellSize by contract
    uint32_t tableSize; // REFERENCE_HASH
    int3 gridMin;       // LINEAR_CELL: cell coordinate of grid corner
    int3 gridDim;
};

struct BoxParams {
    float3 bmin, bmax;
    float collisionMinDist, collisionStiffness;   // 0.025, 0.5 (ConstraintSolverCUDA.cu:137-138)
    float dampingMinDist, restitution, oneMinusFriction;  // 0.025 (IntegrationCUDA.cu:88), 0, 1-0.95 (PBFSolver.cpp:64)
};

// ------------------------------------------------------------------------------------------------ programmatic dependent launch
// The step is ~23 short kernels in a row; with programmatic dependent launch (options.use_pdl) kernel N+1 is scheduled while
// kernel N drains instead of after it, which hides the launch latency between them. Every kernel of the step calls
// pdl_wait() before it touches global memory (it returns once ALL grids it depends on have completed and flushed — so the
// data dependencies are exactly those of plain stream order) and pdl_trigger() once its main loop is done (the next grid may
// start occupying SM slots from then on). Both are no-ops for a kernel launched without the attribute.
// (AKUA_HOST_EMU: the kernels compiled for the CPU by tests/emu — test infrastructure only, see tests/emu/cuda_runtime.h; the
// inline PTX of this file has a plain C++ equivalent under that macro and nowhere else.)
#ifdef AKUA_HOST_EMU
__device__ __forceinline__ void pdl_wait() {}
__device__ __forceinline__ void pdl_trigger() {}
#else
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
#endif

// Nanosecond wall clock of the GPU (x-slab mode: per-rank busy time for the re-balancing). 0 under the host emulation.
__device__ __forceinline__ unsigned long long global_timer_ns() {
#ifdef AKUA_HOST_EMU
    return 0ull;
#else
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
#endif
}

// ------------------------------------------------------------------------------------------------ device math
// dist2 / cell_of / clampi / linear_key are __host__ __device__ so that tests/cpp/list_build_host.cu can run the per-particle
// list-build functions of list_build.cuh on the CPU (same source, IEEE-identical float expressions); the device code is unchanged.
__host__ __device__ __forceinline__ float dist2(float dx, float dy, float dz) {
    // exactly the reference's contraction (kernel_find_neighbours SASS): fma(dz,dz, fma(dx,dx, dy*dy))
#ifdef __CUDA_ARCH__
    return __fmaf_rn(dz, dz, __fmaf_rn(dx, dx, __fmul_rn(dy, dy)));
#else
    const volatile float yy = dy * dy;   // volatile: the product is rounded before it enters the fma, whatever -ffp-contract says
    return fmaf(dz, dz, fmaf(dx, dx, yy));
#endif
}
__device__ __forceinline__ float poly6(float d2, const SphParams& P) {  // SmoothingKernelsCUDA.h:16-21
    // d2 > h2 ? 0 : coef * (h2 - d2)^3, with the cut-off as a clamp (one FMNMX instead of FSETP + FSEL): for d2 <= h2 the
    // clamp is the identity, beyond it the cube is exactly 0.
    float t = fmaxf(P.h2 - d2, 0.0f);
    return P.poly6Coef * (t * t * t);
}
// Spiky gradient, SmoothingKernelsCUDA.h:23-28. Returns the scalar s such that grad = s * r_vector
// (s = coef*(h-r)^2 / r, 0 outside (1e-5, h]; the r > h cut-off is a clamp of h - r, see poly6).
template <bool FAST>
__device__ __forceinline__ float spiky_scale(float d2, const SphParams& P) {
    float r, invr;
    if (FAST) {
        // one MUFU.RSQ: the operand is clamped to a normal float, so the denormal pre/post-scaling of rsqrtf() (three more
        // instructions per neighbour) can never trigger. d2 == 0 (coincident / masked self slot) -> r = 0 -> s = 0.
        float d = fmaxf(d2, 1e-30f);
#ifdef AKUA_HOST_EMU
        invr = 1.0f / sqrtf(d);
#else
        asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(invr) : "f"(d));
#endif
        r = d2 * invr;
    } else { r = sqrtf(d2); invr = 1.0f / r; }
    float t = fmaxf(P.h - r, 0.0f);
    float s = P.spikyCoef * (t * t) * invr;
    return r < 1e-5f ? 0.0f : s;
}
__host__ __device__ __forceinline__ int3 cell_of(float x, float y, float z, float cellSize) {
    // NeighbourSearchCUDA.cu:15-21: floorf of a true IEEE division
#ifdef __CUDA_ARCH__
    return make_int3((int)floorf(__fdiv_rn(x, cellSize)), (int)floorf(__fdiv_rn(y, cellSize)),
                     (int)floorf(__fdiv_rn(z, cellSize)));
#else
    return make_int3((int)floorf(x / cellSize), (int)floorf(y / cellSize), (int)floorf(z / cellSize));
#endif
}
__device__ __forceinline__ uint32_t ref_hash(int cx, int cy, int cz, uint32_t tableSize) {
    // NeighbourSearchCUDA.cu:23-27
    return (((uint32_t)cx * 73856093u) ^ ((uint32_t)cy * 19349663u) ^ ((uint32_t)cz * 83492791u)) % tableSize;
}
__host__ __device__ __forceinline__ int clampi(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }
__host__ __device__ __forceinline__ uint32_t linear_key(int3 c, const GridParams& G) {
    int x = clampi(c.x - G.gridMin.x, 0, G.gridDim.x - 1);
    int y = clampi(c.y - G.gridMin.y, 0, G.gridDim.y - 1);
    int z = clampi(c.z - G.gridMin.z, 0, G.gridDim.z - 1);
    return (uint32_t)((x * G.gridDim.y + y) * G.gridDim.z + z);
}

// ------------------------------------------------------------------------------------------------ x-slab step dimensions
// In x-slab (multi-GPU) mode every per-step size lives in a device-resident block of 64 u32 words ("dims"), written by the
// migration kernels and by slab::k_slab_plan and read by every kernel of the step: the host never learns this step's sizes,
// so the step needs no host synchronisation and can be replayed as a CUDA graph. Launch grids come from a host ESTIMATE
// (the sizes of an earlier step, read back asynchronously); kernels loop (grid-stride) so any estimate is correct.
enum : int {
    D_OUT_L = 0, D_OUT_R = 1,                                       // leavers of this step (k_mig_scan)
    D_STAY_FIRST = 2, D_STAY_LAST = 3, D_LAND_L = 4, D_LAND_R = 5,  // plane populations (k_mig_count)
    D_MASS = 8,                                                     // 8..11: scratch of the uniform-mass agreement
    D_MSG_TO_L = 16, D_MSG_TO_R = 20, D_MSG_FROM_L = 24, D_MSG_FROM_R = 28,   // the per-step count messages (3 words each)
    D_ERROR = 31,                                                   // sticky error word (akua_slab_error bits)
    D_N = 32,                                                       // owned particles between steps
    D_IN_L = 33, D_IN_R = 34, D_NPRE = 35, D_NOWN = 36,             // arrivals; resident before the sort; owned after it
    D_PLANE_L = 37, D_PLANE_R = 38, D_GHOST_L = 39, D_GHOST_R = 40, // boundary planes sent / ghost planes received
    D_EPOCH = 41,                                                   // epoch base of this step: exchange e carries D_EPOCH + e + 1
    D_STAT_MIG_IN = 42, D_STAT_MIG_OUT = 44, D_STAT_BYTES = 46,     // u64 accumulators (two words each)
    D_STEPS = 48,                                                   // steps completed on the device
    // 49, 50: payload-slot free stack (slab_kernels.cuh)
    D_PLAN_WAIT_NS = 51,                                            // this step: time k_slab_plan spent waiting for the neighbours' count messages
    D_T_LAST = 52, D_BUSY_NS = 54,                                  // u64: globaltimer at the end of the last step; busy time (step time minus
                                                                    // that wait) accumulated since the last re-balancing — what it balances
    D_BUSY_STEPS = 56,                                              // steps accumulated in D_BUSY_NS
    D_XLO = 57, D_XHI = 58,                                         // owned x planes [lo, hi) in slab-local grid coordinates: written by the
                                                                    // host when the slab interval changes (akua_pbf_set_slab / _rebalance), read
                                                                    // by the migration kernels — moving a boundary does not touch the step's graph
    D_STAT_PLAN_WAIT_NS = 60, D_STAT_HALO_WAIT_NS = 62,            // u64 accumulators: idle time waiting for the neighbours' count message;
                                                                    // time CTA 0 of the sweeps spent waiting for ghost planes (exposed halo latency)
    D_WORDS = 64
};
enum : uint32_t {   // bits of dims[D_ERROR]
    SLAB_ERR_PLANE_PREDICTION = 1u, SLAB_ERR_TIMEOUT = 2u, SLAB_ERR_MIG_OVERFLOW = 4u, SLAB_ERR_CAPACITY = 8u,
    SLAB_ERR_GHOST_OVERFLOW = 16u
};

// Index span of a sweep launch: thread t handles particle base + t (+ skip once t >= split). One launch can thus cover
// the whole owned range (single GPU), the interior of a slab, or its two boundary planes (multi-GPU overlap).
// mode 0: the host filled count/base/split/skip. Slab mode: resolved on the device from `dims` —
// mode 1 = all owned particles, 2 = slab interior (needs no ghost data), 3 = the two boundary planes.
// mode 4 (SPAN_FUSED, CUDA-IPC transport) = interior AND boundary planes in ONE launch: the first CTAs of the grid take the
// boundary planes (they wait for the ghosts, push their results to the neighbours and publish the epoch), the others the
// interior — see sweep_cta().
enum : int { SPAN_FIXED = 0, SPAN_OWNED = 1, SPAN_INTERIOR = 2, SPAN_BOUNDARY = 3, SPAN_FUSED = 4 };
struct Span { uint32_t count, base, split, skip; const uint32_t* dims; int mode; };
__device__ __forceinline__ Span resolve_span(Span sp) {
    if (sp.mode == SPAN_FIXED) return sp;
    const uint32_t n = sp.dims[D_NOWN], pl = sp.dims[D_PLANE_L], pr = sp.dims[D_PLANE_R];
    const bool allBoundary = (uint64_t)pl + pr >= n;   // slab only one or two planes wide: everything is boundary
    sp.split = 0xffffffffu; sp.skip = 0u; sp.base = 0u;
    if (sp.mode == SPAN_OWNED) sp.count = n;
    else if (sp.mode == SPAN_INTERIOR) { sp.count = allBoundary ? 0u : n - pl - pr; sp.base = pl; }
    else if (allBoundary) sp.count = n;
    else { sp.count = pl + pr; sp.split = pl; sp.skip = n - pr - pl; }
    return sp;
}
__device__ __forceinline__ uint32_t span_particle(const Span& sp, uint32_t t) { return sp.base + t + (t >= sp.split ? sp.skip : 0u); }
// What one CTA of a slab-mode sweep does: its first thread index, stride and count within the (resolved) span, whether it runs
// the halo protocol, and how many CTAs of the launch do (the last of them to finish publishes the epoch).
// SPAN_FUSED: boundary CTAs come FIRST in the grid (CTAs are dispatched in index order, so the boundary planes are computed and
// on their way over NVLink while the interior is still being swept: compute, halo push and synchronisation in one launch, no
// second stream, no launch priorities). At least one CTA takes the boundary role even when the planes are empty (the neighbours
// still wait for the epoch), and the interior keeps at least half of the grid whatever the host's size estimate was.
struct SweepCta { uint32_t t0, tstep, haloCtas; bool halo; };
__device__ __forceinline__ SweepCta sweep_cta(Span& sp) {
    SweepCta c;
    if (sp.mode != SPAN_FUSED) {
        sp = resolve_span(sp);
        c.t0 = blockIdx.x * blockDim.x + threadIdx.x; c.tstep = gridDim.x * blockDim.x; c.haloCtas = gridDim.x; c.halo = true;
        return c;
    }
    const uint32_t n = sp.dims[D_NOWN], pl = sp.dims[D_PLANE_L], pr = sp.dims[D_PLANE_R];
    const bool allBoundary = (uint64_t)pl + pr >= n;
    const uint32_t bcount = allBoundary ? n : pl + pr;
    const uint32_t want = (bcount + blockDim.x - 1) / blockDim.x;
    const uint32_t nb = max(1u, min(want, max(1u, gridDim.x >> 1)));
    c.haloCtas = nb;
    c.halo = blockIdx.x < nb;
    sp.split = 0xffffffffu; sp.skip = 0u; sp.base = 0u;
    if (c.halo) {
        sp.count = bcount;
        if (!allBoundary) { sp.split = pl; sp.skip = n - pr - pl; }
        c.t0 = blockIdx.x * blockDim.x + threadIdx.x; c.tstep = nb * blockDim.x;
    } else {
        sp.count = n - bcount; sp.base = pl;
        c.t0 = (blockIdx.x - nb) * blockDim.x + threadIdx.x; c.tstep = (gridDim.x - nb) * blockDim.x;
    }
    return c;
}

// Fused compute + halo push (multi-GPU, CUDA-IPC transport): a boundary particle's result is also stored straight into the
// neighbouring rank's ghost region through the peer-mapped pointer (NVLink P2P store), so no separate copy or
// collective follows the sweep. dstL / dstR already point at the ghost region's first element; null = nothing to push.
// The plane sizes come from `dims` (resolve_push).
struct PeerPush {
    void* dstL = nullptr;   // left rank: particles [0, nL) of this rank's owned range
    void* dstR = nullptr;   // right rank: particles [startR, nOwn)
    uint32_t nL = 0, startR = 0xffffffffu;
    const uint32_t* dims = nullptr;
};
__device__ __forceinline__ void resolve_push(PeerPush& pp) {
    if (!pp.dims) return;
    pp.nL = pp.dims[D_PLANE_L];
    pp.startR = pp.dims[D_NOWN] - pp.dims[D_PLANE_R];
}
template <typename T>
__device__ __forceinline__ void peer_push(const PeerPush& pp, uint32_t i, const T& v) {
    if (pp.dstL && i < pp.nL) static_cast<T*>(pp.dstL)[i] = v;
    if (pp.dstR && i >= pp.startR) static_cast<T*>(pp.dstR)[i - pp.startR] = v;
}

// In-kernel halo synchronisation: a boundary sweep first waits (one thread per CTA, bounded spin) until both neighbours
// have published the epoch of the ghost data it is about to read, and, when it has pushed its own results, the LAST CTA
// to finish publishes this exchange's epoch in the neighbours' flag words. No extra kernels, no copy engine, no
// collective: compute, communication and synchronisation are one launch. Epochs are dims[D_EPOCH] + index + 1, so the
// kernel arguments are the same every step (CUDA-graph replay).
struct HaloSync {
    const uint32_t* waitFlags = nullptr;   // this rank's flag words: [0] written by the left rank, [1] by the right rank
    int waitL = 0, waitR = 0;
    int waitIdx = -1;                      // exchange index to wait for (-1: none)
    uint32_t* signalL = nullptr;           // neighbours' flag words to publish into (null: no neighbour / nothing pushed)
    uint32_t* signalR = nullptr;
    int signalIdx = -1;
    uint32_t* doneCounter = nullptr;       // CTA completion counter for this launch (zero before and after)
    uint32_t* dims = nullptr;              // D_EPOCH, D_ERROR
    long long timeoutCycles = 0;
};
// Bounded spin of ONE thread on a flag word until it reaches `epoch` (wrap-safe); false on time-out. `err` (the sticky error
// word of dims): once ANY wait of this rank has timed out, the neighbour is gone and every later wait gives up at once — a rank
// whose neighbour died pays the time-out once, not once per queued kernel (its queue drains in milliseconds and the host's
// next look at the error word reports AKUA_ERR_COMM).
__device__ __forceinline__ bool spin_until(const volatile uint32_t* f, uint32_t epoch, long long timeoutCycles,
                                           const volatile uint32_t* err = nullptr) {
    const long long t0 = clock64();
    unsigned ns = 32;
    while ((int32_t)(*f - epoch) < 0) {
        if (err && (*err & SLAB_ERR_TIMEOUT)) return false;
        if (clock64() - t0 > timeoutCycles) return false;
        __nanosleep(ns);
        if (ns < 1024) ns <<= 1;
    }
    return true;
}
__device__ __forceinline__ void halo_wait(const HaloSync& hs) {
    if (!hs.waitFlags || hs.waitIdx < 0) return;
    if (threadIdx.x == 0) {
        const uint32_t epoch = hs.dims[D_EPOCH] + (uint32_t)hs.waitIdx + 1u;
        const unsigned long long t0 = blockIdx.x == 0 ? global_timer_ns() : 0ull;
        for (int side = 0; side < 2; side++) {
            if (!(side == 0 ? hs.waitL : hs.waitR)) continue;
            if (!spin_until(hs.waitFlags + side, epoch, hs.timeoutCycles, hs.dims + D_ERROR)) { atomicOr(hs.dims + D_ERROR, (uint32_t)SLAB_ERR_TIMEOUT); break; }
        }
        __threadfence_system();
        // the first boundary CTA of every sweep records how long the ghosts kept it waiting: the halo latency that was NOT hidden
        if (blockIdx.x == 0) atomicAdd(reinterpret_cast<unsigned long long*>(hs.dims + D_STAT_HALO_WAIT_NS), global_timer_ns() - t0);
    }
    __syncthreads();
}
// Call with ALL threads of the CTA (no early returns before it) once the CTA's pushes are issued.
__device__ __forceinline__ void halo_signal(const HaloSync& hs, uint32_t ctas) {
    if (!hs.doneCounter || hs.signalIdx < 0) return;
    __syncthreads();   // every thread's pushes happen-before thread 0's fence below (fences are cumulative)
    if (threadIdx.x == 0) {
        __threadfence_system();
        const uint32_t done = atomicAdd(hs.doneCounter, 1u);
        if (done == ctas - 1) {
            __threadfence_system();
            const uint32_t epoch = hs.dims[D_EPOCH] + (uint32_t)hs.signalIdx + 1u;
            if (hs.signalL) *(volatile uint32_t*)hs.signalL = epoch;
            if (hs.signalR) *(volatile uint32_t*)hs.signalR = epoch;
            *hs.doneCounter = 0;
            __threadfence_system();
        }
    }
}
// particle count of a kernel launch: the host's value, or (slab mode) a word of `dims`
__device__ __forceinline__ uint32_t live_count(uint32_t n, const uint32_t* nPtr) { return nPtr ? *nPtr : n; }

// ------------------------------------------------------------------------------------------------ neighbour list access
__host__ __device__ __forceinline__ size_t list_slot(uint32_t i, uint32_t k, uint32_t stride) {
    return ((size_t)(k >> 2) * stride + i) * 4 + (k & 3);
}
// Streaming 16-byte load of four neighbour indices: read once per sweep, so keep it out of L1 (the gathers want L1).
__device__ __forceinline__ uint4 ld_list4(const uint4* p) {
    uint4 v;
#ifdef AKUA_HOST_EMU
    v = *p;
#else
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
#endif
    return v;
}
// Post-solve gather record: committed position (w = mass) and velocity (w = density) of one particle side by side, so that
// K11 / K13 fetch a neighbour with ONE 32-byte gather (LDG.E.256, sm_100+) instead of two 16-byte gathers into two arrays.
// The gathers are bound by L1 data-stage wavefronts (a quarter-warp of scattered 16-byte reads collides on the L1 banks:
// ~9.8 wavefronts per request measured, profiles/r01_ncu_final_summary.txt), so one wider request beats two.
struct __align__(32) PosVel { float4 x, v; };
__device__ __forceinline__ PosVel ld_posvel(const PosVel* p) {
    PosVel r;
#ifdef AKUA_HOST_EMU
    r = *p;
#else
    asm volatile("ld.global.nc.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=f"(r.x.x), "=f"(r.x.y), "=f"(r.x.z), "=f"(r.x.w), "=f"(r.v.x), "=f"(r.v.y), "=f"(r.v.z), "=f"(r.v.w)
                 : "l"(p));
#endif
    return r;
}
// Drives one neighbour sweep for particle i: `load(j)` gathers whatever the sweep needs of neighbour j, `acc(payload,
// valid)` accumulates it. Indices come four at a time; the four gathers are independent and issued back to back, and
// the next index group is already in flight while the current one is evaluated. Tail slots of the last group hold the
// particle's own index (always cached) and are masked out through `valid`. Neighbours are visited in list order, so
// every accumulator sees its terms in the reference's order.
template <typename Payload, typename LoadF, typename AccF>
__device__ __forceinline__ void neighbour_sweep(const uint32_t* __restrict__ list, uint32_t i, uint32_t c,
                                                uint32_t stride, LoadF load, AccF acc) {
    if (c == 0) return;
    const uint4* lp = reinterpret_cast<const uint4*>(list) + i;
    const uint32_t groups = (c + 3) >> 2;
    uint4 cur = ld_list4(lp);
    for (uint32_t g = 0; g < groups; g++) {
        uint4 nxt = cur;
        if (g + 1 < groups) nxt = ld_list4(lp + (size_t)(g + 1) * stride);
        Payload p0 = load(cur.x), p1 = load(cur.y), p2 = load(cur.z), p3 = load(cur.w);
        const uint32_t k = g * 4;
        acc(p0, true);
        acc(p1, k + 1 < c);
        acc(p2, k + 2 < c);
        acc(p3, k + 3 < c);
        cur = nxt;
    }
}

// ------------------------------------------------------------------------------------------------ K1 + K2
// kernel_predict_position (IntegrationCUDA.cu:27-36) fused with kernel_compute_hashes (NeighbourSearchCUDA.cu:36-44).
template <int MODE>
__global__ void __launch_bounds__(256) k_predict_key(const float4* __restrict__ pos, const float4* __restrict__ vel,
                                                     float4* __restrict__ xs, uint32_t* __restrict__ keys, uint32_t n,
                                                     const uint32_t* __restrict__ nPtr, float dt, float3 g, GridParams G,
                                                     int doPredict) {
    pdl_wait();
    n = live_count(n, nPtr);
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        float4 x;
        if (doPredict) {
            float4 p = pos[i], v = vel[i];
            float vx = __fmaf_rn(g.x, dt, v.x), vy = __fmaf_rn(g.y, dt, v.y), vz = __fmaf_rn(g.z, dt, v.z);
            x = make_float4(__fmaf_rn(vx, dt, p.x), __fmaf_rn(vy, dt, p.y), __fmaf_rn(vz, dt, p.z), p.w);
            xs[i] = x;
        } else {
            x = xs[i];
        }
        if (keys) {
            int3 c = cell_of(x.x, x.y, x.z, G.cellSize);
            keys[i] = MODE == KEY_HASH ? ref_hash(c.x, c.y, c.z, G.tableSize) : linear_key(c, G);
        }
    }
}

// ------------------------------------------------------------------------------------------------ reorder + K3
// Gathers the state into key-sorted order (the reference physically sorts the structs, NeighbourSearchCUDA.cu:167-170)
// and records bucket / cell ranges (kernel_build_hash_table, :52-65).
template <int MODE>
__global__ void __launch_bounds__(256) k_reorder_ranges(const uint32_t* __restrict__ keysSorted,
                                                        const uint32_t* __restrict__ perm, uint32_t n,
                                                        const uint32_t* __restrict__ nPtr,
                                                        const float4* __restrict__ posIn, const float4* __restrict__ velIn,
                                                        const float4* __restrict__ xsIn, const uint32_t* __restrict__ idIn,
                                                        float4* __restrict__ posOut, float4* __restrict__ velOut,
                                                        float4* __restrict__ xsOut, uint32_t* __restrict__ idOut,
                                                        uint32_t* __restrict__ bucketStart, uint2* __restrict__ cellRange,
                                                        const uint32_t* __restrict__ slotIn, uint32_t* __restrict__ slotOut) {
    pdl_wait();
    n = live_count(n, nPtr);
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        uint32_t src = perm[i];
        posOut[i] = posIn[src];
        velOut[i] = velIn[src];
        xsOut[i] = xsIn[src];
        idOut[i] = idIn[src];
        if (slotIn) slotOut[i] = slotIn[src];   // x-slab mode: where the particle's render payload lives (see k_pack_aos)
        uint32_t k = keysSorted[i];
        bool first = (i == 0) || (keysSorted[i - 1] != k);
        if (MODE == KEY_HASH) {
            if (first) bucketStart[k] = i;
        } else {
            bool last = (i == n - 1) || (keysSorted[i + 1] != k);
            if (first) cellRange[k].x = i;
            if (last) cellRange[k].y = i + 1;
        }
    }
}
// options.canonical_order: after the particles were sorted by id (byId = that permutation), their cell keys are gathered in id
// order and travel with the permutation into the stable sort by key — so particles of one cell end up ordered by id.
__global__ void __launch_bounds__(256) k_gather_keys(const uint32_t* __restrict__ keys, const uint32_t* __restrict__ byId, uint32_t n,
                                                     const uint32_t* __restrict__ nPtr, uint32_t* __restrict__ keysOut,
                                                     uint32_t* __restrict__ valsOut) {
    pdl_wait();
    n = live_count(n, nPtr);
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const uint32_t src = byId[i];
        keysOut[i] = keys[src];
        valsOut[i] = src;
    }
}
// Undo last step's bucket-start writes instead of refilling the 128*N-entry table (the reference allocates and fills
// it with UINT32_MAX every step: NeighbourSearchCUDA.cu:157 — 512 B per particle per step).
__global__ void __launch_bounds__(256) k_clear_buckets(const uint32_t* __restrict__ keysSorted, uint32_t n,
                                                       uint32_t* __restrict__ bucketStart) {
    pdl_wait();
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint32_t k = keysSorted[i];
    if (i == 0 || keysSorted[i - 1] != k) bucketStart[k] = 0xffffffffu;
}
__global__ void __launch_bounds__(256) k_fill_u32(uint32_t* __restrict__ p, uint64_t count, uint32_t v) {
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (; i < count; i += stride) p[i] = v;
}

__global__ void __launch_bounds__(256) k_fill_payload(float4* __restrict__ color, float* __restrict__ size, uint32_t n, float4 c,
                                                      float sz) {
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) { color[i] = c; size[i] = sz; }
}

// ------------------------------------------------------------------------------------------------ K4
// kernel_find_neighbours (NeighbourSearchCUDA.cu:72-130): 27-cell scan in the reference's order (dx outer, dz inner,
// bucket order = sorted order), strict d2 < h*h, self skipped, capped at maxNeighbours. The list is written
// column-major in groups of four (list_slot); the unused tail of the last group is padded with the particle's own index.
// STRIDED (x-slab mode): the particle count lives on the device and the kernel loops; otherwise one particle per thread.
template <int MODE, bool STRIDED>
__global__ void __launch_bounds__(256) k_build_neighbours(const float4* __restrict__ xs,
                                                          const uint32_t* __restrict__ keysSorted,
                                                          const uint32_t* __restrict__ bucketStart,
                                                          const uint2* __restrict__ cellRange, uint32_t n,
                                                          uint32_t stride, uint32_t maxN, uint32_t* __restrict__ list,
                                                          uint32_t* __restrict__ cnt, GridParams G, float h,
                                                          const uint32_t* __restrict__ nPtr) {
    pdl_wait();
    if (STRIDED) n = live_count(n, nPtr);
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const float4 xi = xs[i];
    const float h2 = __fmul_rn(h, h);
    const int3 c = cell_of(xi.x, xi.y, xi.z, G.lookupCellSize);
    uint32_t count = 0;
    if (MODE == KEY_HASH) {
        for (int dx = -1; dx <= 1; dx++)
            for (int dy = -1; dy <= 1; dy++)
                for (int dz = -1; dz <= 1; dz++) {
                    uint32_t hash = ref_hash(c.x + dx, c.y + dy, c.z + dz, G.tableSize);
                    uint32_t cand = bucketStart[hash];
                    if (cand == 0xffffffffu) continue;
                    while (cand < n && count < maxN) {
                        if (cand == i) { cand++; continue; }
                        if (keysSorted[cand] != hash) break;
                        float4 xj = __ldg(&xs[cand]);
                        float d2 = dist2(xi.x - xj.x, xi.y - xj.y, xi.z - xj.z);
                        if (d2 < h2) { list[list_slot(i, count, stride)] = cand; count++; }
                        cand++;
                    }
                }
    } else {
        const int cx = clampi(c.x - G.gridMin.x, 0, G.gridDim.x - 1);
        const int cy = clampi(c.y - G.gridMin.y, 0, G.gridDim.y - 1);
        const int cz = clampi(c.z - G.gridMin.z, 0, G.gridDim.z - 1);
        const int z0 = max(cz - 1, 0), z1 = min(cz + 1, G.gridDim.z - 1);
        for (int dx = -1; dx <= 1; dx++) {
            int X = cx + dx;
            if (X < 0 || X >= G.gridDim.x) continue;
            for (int dy = -1; dy <= 1; dy++) {
                int Y = cy + dy;
                if (Y < 0 || Y >= G.gridDim.y) continue;
                // the (up to) three z-adjacent cells are contiguous in sorted order: one row range
                const uint2* row = cellRange + ((size_t)X * G.gridDim.y + Y) * G.gridDim.z;
                uint32_t s = 0xffffffffu, e = 0;
                for (int z = z0; z <= z1; z++) {
                    uint2 r = __ldg(&row[z]);
                    if (r.y > r.x) { s = min(s, r.x); e = max(e, r.y); }
                }
                for (uint32_t cand = s; cand < e && count < maxN; cand++) {
                    if (cand == i) continue;
                    float4 xj = __ldg(&xs[cand]);
                    float d2 = dist2(xi.x - xj.x, xi.y - xj.y, xi.z - xj.z);
                    if (d2 < h2) { list[list_slot(i, count, stride)] = cand; count++; }
                }
            }
        }
    }
    if (!STRIDED) pdl_trigger();
    cnt[i] = count;
    for (uint32_t k = count; k < ((count + 3u) & ~3u); k++) list[list_slot(i, k, stride)] = i;
    if (!STRIDED) break;
    }
    if (STRIDED) pdl_trigger();
}

// ------------------------------------------------------------------------------------------------ pass A = K5 + K6
// kernel_calculate_densities (ConstraintSolverCUDA.cu:16-42) + kernel_calculate_lambdas (:51-97), one neighbour loop.
// Each accumulator sees its terms in the reference's order, so fusing the loops does not change the sums.
template <bool FAST, bool SLAB>
__global__ void AKUA_SWEEP_BOUNDS k_density_lambda(const float4* __restrict__ xs, const uint32_t* __restrict__ list,
                                                        const uint32_t* __restrict__ cnt, uint32_t stride, Span sp,
                                                        float* __restrict__ density, float* __restrict__ lambda,
                                                        float4* __restrict__ xl, SphParams P, PeerPush pushLambda,
                                                        HaloSync hs) {
    pdl_wait();
    SweepCta cta{};
    if (SLAB) { cta = sweep_cta(sp); if (cta.halo) halo_wait(hs); resolve_push(pushLambda); }
    const uint32_t tstep = SLAB ? cta.tstep : 0u;
    for (uint32_t t = SLAB ? cta.t0 : blockIdx.x * blockDim.x + threadIdx.x; t < sp.count; t += tstep) {
    const uint32_t i = span_particle(sp, t);
    const float4 xi = xs[i];
    const uint32_t c = cnt[i];
    float rho = xi.w * P.selfW;
    float gx = 0.f, gy = 0.f, gz = 0.f, sum = 0.f;
    neighbour_sweep<float4>(list, i, c, stride,
        [&](uint32_t j) { return __ldg(&xs[j]); },
        [&](const float4& xj, bool valid) {
            float dx = xi.x - xj.x, dy = xi.y - xj.y, dz = xi.z - xj.z;
            float d2 = dist2(dx, dy, dz);
            float m = valid ? xj.w : 0.0f;                       // masked tail slots contribute exactly nothing
            rho = fmaf(m, poly6(d2, P), rho);
            float s = spiky_scale<FAST>(d2, P);
            float ax = s * dx, ay = s * dy, az = s * dz;        // grad W_spiky
            gx = fmaf(m, ax, gx); gy = fmaf(m, ay, gy); gz = fmaf(m, az, gz);
            float q = -P.invRestDensity * m;                     // grad_pj C_i = q * gradW
            if (FAST) {                                          // |q * s * r_vec|^2 = (q s)^2 d2: 3 instructions, not 8
                float qs = q * s;
                sum = fmaf(qs * qs, d2, sum);
            } else {
                float bx = q * ax, by = q * ay, bz = q * az;
                sum += fmaf(bz, bz, fmaf(bx, bx, by * by));
            }
        });
    if (!SLAB) pdl_trigger();
    gx *= P.invRestDensity; gy *= P.invRestDensity; gz *= P.invRestDensity;
    float C = rho * P.invRestDensity - 1.0f;
    float lam = -C / (sum + fmaf(gz, gz, fmaf(gx, gx, gy * gy)) + P.relaxation);
    density[i] = rho;
    lambda[i] = lam;
    if (xl) {   // packed-gather layout: pass B fetches x* and lambda in one gather (the halo push then carries the pair too)
        const float4 v = make_float4(xi.x, xi.y, xi.z, lam);
        xl[i] = v;
        if (SLAB) peer_push(pushLambda, i, v);
    } else {
        if (SLAB) peer_push(pushLambda, i, lam);
    }
    if (!SLAB) break;
    }
    if (SLAB) { pdl_trigger(); if (cta.halo) halo_signal(hs, cta.haloCtas); }
}

// ------------------------------------------------------------------------------------------------ K8 / K9 / K10 pieces
__device__ __forceinline__ float collide_axis(float x, float lo, float hi, const BoxParams& B) {
    // handle_particle_collision, ConstraintSolverCUDA.cu:136-157 (both tests see the uncorrected coordinate)
    float c = 0.0f;
    float a = lo + B.collisionMinDist, b = hi - B.collisionMinDist;
    if (x < a) c += B.collisionStiffness * (a - x);
    if (x > b) c += B.collisionStiffness * (b - x);
    return x + c;
}
// resolve_collision (IntegrationCUDA.cu:51-73) specialised to an axis-aligned plane with normal sign*e_axis:
// va is the velocity component along the axis, vb/vc the tangential ones.
__device__ __forceinline__ void damp_plane(float dist, float sign, float& va, float& vb, float& vc, const BoxParams& B) {
    float approaching = sign * va;
    if (dist < B.dampingMinDist) {
        if (approaching < 0.0f) {
            va = -B.restitution * va;
            vb = B.oneMinusFriction * vb;
            vc = B.oneMinusFriction * vc;
        } else if (fabsf(approaching) < 1e-5f) {
            va = 0.0f;
            vb = B.oneMinusFriction * vb;
            vc = B.oneMinusFriction * vc;
        }
    }
}
// kernel_apply_boundary_velocity_damping, IntegrationCUDA.cu:75-102: planes in order xmin,xmax,ymin,ymax,zmin,zmax
__device__ __forceinline__ void damp_velocity(float px, float py, float pz, float& vx, float& vy, float& vz,
                                              const BoxParams& B) {
    damp_plane(px - B.bmin.x, 1.0f, vx, vy, vz, B);
    damp_plane(B.bmax.x - px, -1.0f, vx, vy, vz, B);
    damp_plane(py - B.bmin.y, 1.0f, vy, vx, vz, B);
    damp_plane(B.bmax.y - py, -1.0f, vy, vx, vz, B);
    damp_plane(pz - B.bmin.z, 1.0f, vz, vx, vy, B);
    damp_plane(B.bmax.z - pz, -1.0f, vz, vx, vy, B);
}

// ------------------------------------------------------------------------------------------------ pass B = K7 + K8 (+K9+K10)
// kernel_calculate_position_delta (ConstraintSolverCUDA.cu:99-130) + kernel_correct_position (:159-169); Jacobi, so the
// corrected x* goes to the other half of a double buffer. FINAL additionally commits: kernel_update_position_and_velocity
// (IntegrationCUDA.cu:38-49) and kernel_apply_boundary_velocity_damping (:75-102), both per-particle.
// PACK (every particle has the mass P.uniformMass): x* and lambda of a neighbour come from ONE gather of pass A's packed
// (x*, lambda) array `xl` instead of a 16-byte and a 4-byte gather; the arithmetic is the same expression on the same values.
// CORR4: the artificial-pressure exponent n is 4 (the reference's default, PBFConfig.h:14): two multiplies instead of powf,
// and no per-neighbour branch on it.
template <bool FAST, bool FINAL, bool PACK, bool CORR4, bool SLAB>
__global__ void AKUA_SWEEP_BOUNDS k_delta_apply(const float4* __restrict__ xsIn, float4* __restrict__ xsOut,
                                                     const float* __restrict__ lambda, const float4* __restrict__ xl,
                                                     const uint32_t* __restrict__ list,
                                                     const uint32_t* __restrict__ cnt, uint32_t stride, Span sp,
                                                     SphParams P, BoxParams B, float4* __restrict__ dposOut,
                                                     float4* __restrict__ pos, float4* __restrict__ vel,
                                                     const float* __restrict__ density, PosVel* __restrict__ pvOut,
                                                     float dt, PeerPush pushX, PeerPush pushV, HaloSync hs) {
    pdl_wait();
    SweepCta cta{};
    if (SLAB) { cta = sweep_cta(sp); if (cta.halo) halo_wait(hs); resolve_push(pushX); resolve_push(pushV); }
    const uint32_t tstep = SLAB ? cta.tstep : 0u;
    for (uint32_t t = SLAB ? cta.t0 : blockIdx.x * blockDim.x + threadIdx.x; t < sp.count; t += tstep) {
    const uint32_t i = span_particle(sp, t);
    float4 xi;
    float li;
    if (PACK) { const float4 t = xl[i]; li = t.w; xi = make_float4(t.x, t.y, t.z, P.uniformMass); }
    else      { xi = xsIn[i]; li = lambda[i]; }
    const uint32_t c = cnt[i];
    float px = 0.f, py = 0.f, pz = 0.f;
    struct NB { float4 x; float l; };
    neighbour_sweep<NB>(list, i, c, stride,
        [&](uint32_t j) {
            NB r;
            if (PACK) { r.x = __ldg(&xl[j]); r.l = r.x.w; r.x.w = P.uniformMass; }
            else      { r.x = __ldg(&xsIn[j]); r.l = __ldg(&lambda[j]); }
            return r;
        },
        [&](const NB& nb, bool valid) {
            float dx = xi.x - nb.x.x, dy = xi.y - nb.x.y, dz = xi.z - nb.x.z;
            float d2 = dist2(dx, dy, dz);
            float ratio = poly6(d2, P) * P.invPoly6Dq;
            float pw;
            if (CORR4) { float r2 = ratio * ratio; pw = r2 * r2; }
            else pw = powf(ratio, P.corrN);
            float corr = -P.corrK * pw;
            float coef = (li + nb.l + corr) * nb.x.w * spiky_scale<FAST>(d2, P);
            coef = valid ? coef : 0.0f;
            px = fmaf(coef, dx, px); py = fmaf(coef, dy, py); pz = fmaf(coef, dz, pz);
        });
    if (!SLAB) pdl_trigger();
    px *= P.invRestDensity; py *= P.invRestDensity; pz *= P.invRestDensity;
    if (dposOut) dposOut[i] = make_float4(px, py, pz, 0.f);
    float x = collide_axis(xi.x + px, B.bmin.x, B.bmax.x, B);
    float y = collide_axis(xi.y + py, B.bmin.y, B.bmax.y, B);
    float z = collide_axis(xi.z + pz, B.bmin.z, B.bmax.z, B);
    xsOut[i] = make_float4(x, y, z, xi.w);
    if (SLAB) peer_push(pushX, i, make_float4(x, y, z, xi.w));
    if (FINAL) {
        float4 p = pos[i];
        float vx = (x - p.x) / dt, vy = (y - p.y) / dt, vz = (z - p.z) / dt;
        damp_velocity(x, y, z, vx, vy, vz, B);
        pos[i] = make_float4(x, y, z, xi.w);
        const float4 vout = make_float4(vx, vy, vz, density[i]);
        vel[i] = vout;
        if (pvOut) { PosVel r; r.x = make_float4(x, y, z, xi.w); r.v = vout; pvOut[i] = r; }   // post-solve gather records
        if (SLAB) peer_push(pushV, i, vout);
    }
    if (!SLAB) break;
    }
    if (SLAB) { pdl_trigger(); if (cta.halo) halo_signal(hs, cta.haloCtas); }
}

// (position, velocity) -> post-solve gather records, for callers that commit outside the fused final pass B (phase-level
// API, solverIterations == 0).
__global__ void __launch_bounds__(256) k_build_posvel(const float4* __restrict__ pos, const float4* __restrict__ vel,
                                                      uint32_t n, PosVel* __restrict__ pv) {
    pdl_wait();
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    PosVel r; r.x = pos[i]; r.v = vel[i];
    pv[i] = r;
}

// Stand-alone K9 / K10 for the phase-level API (and solverIterations == 0).
__global__ void __launch_bounds__(256) k_update(const float4* __restrict__ xs, float4* __restrict__ pos,
                                                float4* __restrict__ vel, const float* __restrict__ density, uint32_t n,
                                                float dt, const uint32_t* __restrict__ nPtr) {
    pdl_wait();
    n = live_count(n, nPtr);
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        float4 x = xs[i], p = pos[i];
        pos[i] = x;
        vel[i] = make_float4((x.x - p.x) / dt, (x.y - p.y) / dt, (x.z - p.z) / dt, density[i]);
    }
}
__global__ void __launch_bounds__(256) k_damping(const float4* __restrict__ pos, float4* __restrict__ vel, uint32_t n,
                                                 BoxParams B, const uint32_t* __restrict__ nPtr) {
    pdl_wait();
    n = live_count(n, nPtr);
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        float4 p = pos[i], v = vel[i];
        damp_velocity(p.x, p.y, p.z, v.x, v.y, v.z, B);
        vel[i] = v;
    }
}

// ------------------------------------------------------------------------------------------------ K11
// kernel_compute_vorticities, IntegrationCUDA.cu:104-128. Also stores |omega| so K12 gathers 4 B per neighbour, not 12.
// REC: neighbour position + velocity come from one 32-byte record gather (PosVel). `xw` (optional) receives the packed
// (x, |omega|) array K12's PACK variant gathers from.
template <bool FAST, bool REC, bool SLAB>
__global__ void AKUA_SWEEP_BOUNDS k_vorticity(const float4* __restrict__ xs, const float4* __restrict__ vel,
                                                   const PosVel* __restrict__ pv,
                                                   const uint32_t* __restrict__ list, const uint32_t* __restrict__ cnt,
                                                   uint32_t stride, Span sp, float4* __restrict__ omega,
                                                   float* __restrict__ omegaLen, float4* __restrict__ xw, SphParams P,
                                                   PeerPush pushLen, HaloSync hs) {
    pdl_wait();
    SweepCta cta{};
    if (SLAB) { cta = sweep_cta(sp); if (cta.halo) halo_wait(hs); resolve_push(pushLen); }
    const uint32_t tstep = SLAB ? cta.tstep : 0u;
    for (uint32_t t = SLAB ? cta.t0 : blockIdx.x * blockDim.x + threadIdx.x; t < sp.count; t += tstep) {
    const uint32_t i = span_particle(sp, t);
    float4 xi, vi;
    if (REC) { const PosVel r = pv[i]; xi = r.x; vi = r.v; }
    else     { xi = xs[i]; vi = vel[i]; }
    const uint32_t c = cnt[i];
    float wx = 0.f, wy = 0.f, wz = 0.f;
    neighbour_sweep<PosVel>(list, i, c, stride,
        [&](uint32_t j) {
            if (REC) return ld_posvel(pv + j);
            PosVel r; r.x = __ldg(&xs[j]); r.v = __ldg(&vel[j]); return r;
        },
        [&](const PosVel& nb, bool valid) {
            float dx = xi.x - nb.x.x, dy = xi.y - nb.x.y, dz = xi.z - nb.x.z;
            float s = spiky_scale<FAST>(dist2(dx, dy, dz), P);
            float gx = s * dx, gy = s * dy, gz = s * dz;
            float ux = nb.v.x - vi.x, uy = nb.v.y - vi.y, uz = nb.v.z - vi.z;
            float cx = uy * gz - uz * gy, cy = uz * gx - ux * gz, cz = ux * gy - uy * gx;  // MathUtilsCUDA.h:12-18
            float m = valid ? -nb.x.w : 0.0f;
            wx = fmaf(m, cx, wx); wy = fmaf(m, cy, wy); wz = fmaf(m, cz, wz);
        });
    if (!SLAB) pdl_trigger();
    float len = sqrtf(fmaf(wz, wz, fmaf(wx, wx, wy * wy)));
    omega[i] = make_float4(wx, wy, wz, len);
    omegaLen[i] = len;
    if (xw) {
        const float4 v = make_float4(xi.x, xi.y, xi.z, len);
        xw[i] = v;
        if (SLAB) peer_push(pushLen, i, v);
    } else {
        if (SLAB) peer_push(pushLen, i, len);
    }
    if (!SLAB) break;
    }
    if (SLAB) { pdl_trigger(); if (cta.halo) halo_signal(hs, cta.haloCtas); }
}

// ------------------------------------------------------------------------------------------------ K12
// kernel_apply_vorticity_confinement, IntegrationCUDA.cu:130-165. Reads neighbours' |omega|, writes only its own
// velocity: race-free in place. PACK (uniform mass): position and |omega| of a neighbour come from one gather of `xw`.
// `pvOut` (optional): the updated velocity is mirrored into the post-solve gather records K13 reads.
template <bool FAST, bool PACK, bool SLAB>
__global__ void AKUA_SWEEP_BOUNDS k_confinement(const float4* __restrict__ xs, const float4* __restrict__ omega,
                                                     const float* __restrict__ omegaLen, const float4* __restrict__ xw,
                                                     const float* __restrict__ density,
                                                     const uint32_t* __restrict__ list, const uint32_t* __restrict__ cnt,
                                                     uint32_t stride, Span sp, float4* __restrict__ vel,
                                                     PosVel* __restrict__ pvOut, SphParams P,
                                                     float dt, float eps, PeerPush pushV, HaloSync hs) {
    pdl_wait();
    SweepCta cta{};
    if (SLAB) { cta = sweep_cta(sp); if (cta.halo) halo_wait(hs); resolve_push(pushV); }
    const uint32_t tstep = SLAB ? cta.tstep : 0u;
    for (uint32_t t = SLAB ? cta.t0 : blockIdx.x * blockDim.x + threadIdx.x; t < sp.count; t += tstep) {
    const uint32_t i = span_particle(sp, t);
    const float4 xi = PACK ? xw[i] : xs[i];
    const float4 oi = omega[i];
    const uint32_t c = cnt[i];
    const float invDensity = 1.0f / density[i];
    float ex = 0.f, ey = 0.f, ez = 0.f;
    struct NB { float4 x; float l; };
    neighbour_sweep<NB>(list, i, c, stride,
        [&](uint32_t j) {
            NB r;
            if (PACK) { r.x = __ldg(&xw[j]); r.l = r.x.w; r.x.w = P.uniformMass; }
            else      { r.x = __ldg(&xs[j]); r.l = __ldg(&omegaLen[j]); }
            return r;
        },
        [&](const NB& nb, bool valid) {
            float dx = xi.x - nb.x.x, dy = xi.y - nb.x.y, dz = xi.z - nb.x.z;
            float coef = nb.x.w * (oi.w - nb.l) * spiky_scale<FAST>(dist2(dx, dy, dz), P);
            coef = valid ? coef : 0.0f;
            ex = fmaf(coef, dx, ex); ey = fmaf(coef, dy, ey); ez = fmaf(coef, dz, ez);
        });
    if (!SLAB) pdl_trigger();
    ex *= invDensity; ey *= invDensity; ez *= invDensity;
    float len = sqrtf(fmaf(ez, ez, fmaf(ex, ex, ey * ey)));
    if (len >= 1e-5f) {
    float nx = ex / len, ny = ey / len, nz = ez / len;
    float fx = eps * (ny * oi.z - nz * oi.y), fy = eps * (nz * oi.x - nx * oi.z), fz = eps * (nx * oi.y - ny * oi.x);
    float4 v = vel[i];
    v.x = fmaf(dt, fx, v.x); v.y = fmaf(dt, fy, v.y); v.z = fmaf(dt, fz, v.z);
    vel[i] = v;
    if (pvOut) pvOut[i].v = v;
    if (SLAB) peer_push(pushV, i, v);   // unchanged velocities were already pushed by the committing pass B
    }
    if (!SLAB) break;
    }
    if (SLAB) { pdl_trigger(); if (cta.halo) halo_signal(hs, cta.haloCtas); }
}

// ------------------------------------------------------------------------------------------------ K13
// kernel_apply_xsph_viscosity, IntegrationCUDA.cu:167-195 — as a Jacobi sweep (velIn -> velOut). The reference updates
// velocity in place while neighbours read it (a data race, :187,:194); Jacobi is one of its legal outcomes and is
// deterministic. REC: one 32-byte record gather per neighbour instead of two 16-byte gathers.
template <bool REC, bool SLAB>
__global__ void AKUA_SWEEP_BOUNDS k_xsph(const float4* __restrict__ xs, const float4* __restrict__ velIn,
                                              const PosVel* __restrict__ pv,
                                              const uint32_t* __restrict__ list, const uint32_t* __restrict__ cnt,
                                              uint32_t stride, Span sp, float4* __restrict__ velOut, SphParams P,
                                              float cvisc, HaloSync hs) {
    pdl_wait();
    SweepCta cta{};
    if (SLAB) { cta = sweep_cta(sp); if (cta.halo) halo_wait(hs); }
    const uint32_t tstep = SLAB ? cta.tstep : 0u;
    for (uint32_t t = SLAB ? cta.t0 : blockIdx.x * blockDim.x + threadIdx.x; t < sp.count; t += tstep) {
    const uint32_t i = span_particle(sp, t);
    float4 xi, vi;
    if (REC) { const PosVel r = pv[i]; xi = r.x; vi = r.v; }
    else     { xi = xs[i]; vi = velIn[i]; }
    const uint32_t c = cnt[i];
    float ax = 0.f, ay = 0.f, az = 0.f;
    neighbour_sweep<PosVel>(list, i, c, stride,
        [&](uint32_t j) {
            if (REC) return ld_posvel(pv + j);
            PosVel r; r.x = __ldg(&xs[j]); r.v = __ldg(&velIn[j]); return r;
        },
        [&](const PosVel& nb, bool valid) {
            float dx = xi.x - nb.x.x, dy = xi.y - nb.x.y, dz = xi.z - nb.x.z;
            float w = poly6(dist2(dx, dy, dz), P);
            float mr = nb.x.w / nb.v.w;  // m_j / rho_j
            w = valid ? w : 0.0f;
            ax = fmaf(mr * (nb.v.x - vi.x), w, ax); ay = fmaf(mr * (nb.v.y - vi.y), w, ay); az = fmaf(mr * (nb.v.z - vi.z), w, az);
        });
    if (!SLAB) pdl_trigger();
    velOut[i] = make_float4(fmaf(cvisc, ax, vi.x), fmaf(cvisc, ay, vi.y), fmaf(cvisc, az, vi.z), vi.w);
    if (!SLAB) break;
    }
    if (SLAB) pdl_trigger();
}

// ------------------------------------------------------------------------------------------------ AoS-108 interchange
// Particle layout: include/AkuaEngine/Simulation/Particle.h:8-31 (27 words). A CTA stages 256 structs through shared
// memory so that global accesses are coalesced; the 27-word stride is odd, so the per-thread reads are conflict-free.
constexpr int kAosWords = 27;
__global__ void __launch_bounds__(256) k_unpack_aos(const uint32_t* __restrict__ aos, uint32_t n, float4* __restrict__ pos,
                                                    float4* __restrict__ vel, float4* __restrict__ xs,
                                                    float4* __restrict__ omega, float* __restrict__ omegaLen,
                                                    float4* __restrict__ dpos, float* __restrict__ density,
                                                    float* __restrict__ lambda, uint32_t* __restrict__ keys,
                                                    float4* __restrict__ color, float* __restrict__ size,
                                                    uint32_t* __restrict__ id, uint32_t idBase) {
    __shared__ uint32_t sm[256 * kAosWords];
    const uint32_t base = blockIdx.x * 256u;
    const uint32_t count = min(256u, n - base);
    const uint32_t* src = aos + (size_t)base * kAosWords;
    for (uint32_t w = threadIdx.x; w < count * kAosWords; w += 256) sm[w] = src[w];
    __syncthreads();
    if (threadIdx.x >= count) return;
    const uint32_t i = base + threadIdx.x;
    const float* f = reinterpret_cast<const float*>(sm + threadIdx.x * kAosWords);
    float mass = f[18], rho = f[19];
    pos[i] = make_float4(f[0], f[1], f[2], mass);
    vel[i] = make_float4(f[3], f[4], f[5], rho);
    xs[i] = make_float4(f[6], f[7], f[8], mass);
    dpos[i] = make_float4(f[12], f[13], f[14], 0.f);
    float len = sqrtf(fmaf(f[17], f[17], fmaf(f[15], f[15], f[16] * f[16])));
    omega[i] = make_float4(f[15], f[16], f[17], len);
    omegaLen[i] = len;
    density[i] = rho;
    lambda[i] = f[20];
    keys[i] = sm[threadIdx.x * kAosWords + 21];
    color[i] = make_float4(f[22], f[23], f[24], f[25]);
    size[i] = f[26];
    id[i] = idBase + i;   // (a chunked upload passes arrays offset to the chunk: the id is the index in the whole upload)
}
__global__ void __launch_bounds__(256) k_pack_aos(uint32_t* __restrict__ aos, uint32_t n, const float4* __restrict__ pos,
                                                  const float4* __restrict__ vel, const float4* __restrict__ xs,
                                                  const float4* __restrict__ omega, const float4* __restrict__ dpos,
                                                  const float* __restrict__ density, const float* __restrict__ lambda,
                                                  const uint32_t* __restrict__ keys, const float4* __restrict__ color,
                                                  const float* __restrict__ size, const uint32_t* __restrict__ id,
                                                  const uint32_t* __restrict__ slot) {
    __shared__ uint32_t sm[256 * kAosWords];
    const uint32_t base = blockIdx.x * 256u;
    const uint32_t count = min(256u, n - base);
    if (threadIdx.x < count) {
        const uint32_t i = base + threadIdx.x;
        float* f = reinterpret_cast<float*>(sm + threadIdx.x * kAosWords);
        float4 p = pos[i], v = vel[i], x = xs[i], o = omega[i], d = dpos[i];
        // The render payload (Particle::color / ::size) is never read by the solver, so it stays where the upload put it and
        // only an index travels through the sorts: the upload index `id` on one GPU; in x-slab mode (ids are global there) a
        // per-rank payload slot that migration re-assigns on the receiving rank (slab::k_mig_unpack).
        const uint32_t pid = slot ? slot[i] : id[i];
        float4 c = color[pid];
        f[0] = p.x; f[1] = p.y; f[2] = p.z;
        f[3] = v.x; f[4] = v.y; f[5] = v.z;
        f[6] = x.x; f[7] = x.y; f[8] = x.z;
        f[9] = v.x; f[10] = v.y; f[11] = v.z;   // new_velocity := velocity (see akua_pbf.h)
        f[12] = d.x; f[13] = d.y; f[14] = d.z;
        f[15] = o.x; f[16] = o.y; f[17] = o.z;
        f[18] = p.w; f[19] = density[i]; f[20] = lambda[i];
        sm[threadIdx.x * kAosWords + 21] = keys[i];
        f[22] = c.x; f[23] = c.y; f[24] = c.z; f[25] = c.w;
        f[26] = size[pid];
    }
    __syncthreads();
    uint32_t* dst = aos + (size_t)base * kAosWords;
    for (uint32_t w = threadIdx.x; w < count * kAosWords; w += 256) dst[w] = sm[w];
}

// neighbour list: column-major device layout -> the reference's row-major n x maxNeighbours (debug tap only)
__global__ void __launch_bounds__(256) k_list_to_rowmajor(const uint32_t* __restrict__ list, const uint32_t* __restrict__ cnt,
                                                          uint32_t stride, uint32_t n, uint32_t maxN,
                                                          uint32_t* __restrict__ out) {
    uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (uint64_t)n * maxN) return;
    uint32_t i = (uint32_t)(t / maxN), k = (uint32_t)(t % maxN);
    out[t] = k < cnt[i] ? list[list_slot(i, k, stride)] : 0u;
}

// min / max of the particle masses (pos.w) as order-preserving u32 codes: out[0] = min code, out[1] = max code
__device__ __forceinline__ uint32_t float_order_code(float f) {
    uint32_t b = __float_as_uint(f);
    return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}
__global__ void __launch_bounds__(256) k_mass_range(const float4* __restrict__ pos, uint32_t n, uint32_t* __restrict__ out) {
    __shared__ uint32_t slo[8], shi[8];
    uint32_t lo = 0xffffffffu, hi = 0u;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        uint32_t c = float_order_code(pos[i].w);
        lo = min(lo, c); hi = max(hi, c);
    }
    lo = __reduce_min_sync(0xffffffffu, lo);
    hi = __reduce_max_sync(0xffffffffu, hi);
    if ((threadIdx.x & 31) == 0) { slo[threadIdx.x >> 5] = lo; shi[threadIdx.x >> 5] = hi; }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < 8; w++) { lo = min(lo, slo[w]); hi = max(hi, shi[w]); }
        atomicMin(&out[0], lo); atomicMax(&out[1], hi);
    }
}

// |rho/rho0 - 1| partial sums / maxima per CTA
__global__ void __launch_bounds__(256) k_density_error(const float* __restrict__ density, uint32_t n, float invRho0,
                                                       float* __restrict__ partSum, float* __restrict__ partMax) {
    __shared__ float ssum[8], smax[8];
    float s = 0.f, m = 0.f;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        float e = fabsf(density[i] * invRho0 - 1.0f);
        s += e; m = fmaxf(m, e);
    }
    for (int o = 16; o > 0; o >>= 1) {
        s += __shfl_xor_sync(0xffffffffu, s, o);
        m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    }
    int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (lane == 0) { ssum[warp] = s; smax[warp] = m; }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < 8; w++) { s += ssum[w]; m = fmaxf(m, smax[w]); }
        partSum[blockIdx.x] = s; partMax[blockIdx.x] = m;
    }
}

}  // namespace akua
