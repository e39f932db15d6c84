// pbf_solver.cu — host side of the B200-native PBF step: owns the device-resident SoA state, sequences the kernels of
// pbf_kernels.cuh / radix_sort.cuh on one stream, and exports the C ABI declared in include/akua_pbf.h.
//
// Mirrors AkuaEngine::PBFSolver (include/AkuaEngine/Simulation/PBFSolver.h, src/Simulation/PBFSolver.cpp). Where the
// reference maps a GL buffer, allocates thrust::device_vectors, copies the neighbour list to the host and back and calls
// cudaDeviceSynchronize after every kernel (25 per step), this solver allocates everything once in create(), keeps all
// state on the device, and issues the whole step asynchronously on one stream.
#include "../../include/akua_pbf.h"
#include "pbf_kernels.cuh"
#include "pbf_params.h"
#include "list_build.cuh"
#include "radix_sort.cuh"
#include "slab_kernels.cuh"
#include "wall_model.cuh"

#include <dlfcn.h>
#include <nccl.h>  // types only: the library itself is dlopen'ed by akua_pbf_comm_init (pbf_slab.inl)

#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>
#include <string>
#include <vector>

using namespace akua;

namespace {

#ifndef AKUA_SWEEP_BLOCK
#define AKUA_SWEEP_BLOCK 128
#endif
constexpr int kBlock = 256;
constexpr int kSweepBlock = AKUA_SWEEP_BLOCK;
inline Span fullSpan(uint32_t n) { return Span{n, 0u, 0xffffffffu, 0u, nullptr, SPAN_FIXED}; }
inline uint32_t sweepGrid(uint64_t n) { return (uint32_t)((n + kSweepBlock - 1) / kSweepBlock); }
inline uint32_t gridFor(uint64_t n) { return (uint32_t)((n + kBlock - 1) / kBlock); }

enum Phase { PH_PREDICT = 0, PH_SORT, PH_REORDER, PH_LISTS, PH_SOLVE, PH_POST, PH_END, PH_COUNT };

}  // namespace

// x-slab multi-GPU state (pbf_slab.inl)
struct SlabPeer {            // a neighbour's allocations, opened through CUDA IPC
    bool open = false;
    float4* xsBuf[2] = {nullptr, nullptr};
    float4* velBuf[2] = {nullptr, nullptr};
    float* lambda = nullptr;
    float* omegaLen = nullptr;
    float4 *xl = nullptr, *xw = nullptr;                        // packed gather arrays (null when a rank has none)
    uint32_t* flags = nullptr;
    uint32_t* dims = nullptr;                                  // the neighbour's dims block (count exchange by P2P stores)
    slab::MigRecord *recvL = nullptr, *recvR = nullptr;       // the neighbour's migration inboxes
    uint32_t migCap = 0;
    uint32_t ghostBaseL = 0, ghostBaseR = 0, ghostCap = 0;
};
struct SlabState {
    bool enabled = false;
    int rank = 0, nranks = 1;
    void* comm = nullptr;                  // ncclComm_t
    int xLoAbs = 0, xHiAbs = 0;            // owned absolute x cells [lo, hi)
    uint32_t* dims = nullptr;              // device u32[D_WORDS]: every size of the current step (pbf_kernels.cuh D_*)
    volatile uint32_t* hDims = nullptr;    // pinned mirror, refreshed asynchronously at the end of every step
    uint32_t* blockCnt = nullptr;          // [2][migBlocksCap]
    uint32_t migBlocksCap = 0;
    slab::MigRecord *sendL = nullptr, *sendR = nullptr, *recvL = nullptr, *recvR = nullptr;
    uint32_t migCap = 0;
    uint32_t *slot = nullptr, *slotAlt = nullptr, *freeSlots = nullptr;   // render-payload slots (travel with the particles)
    // host-side view of the sizes: exact after slabRefresh() (stream synchronised), otherwise an earlier step's
    uint32_t nPlaneL = 0, nPlaneR = 0, nGhostL = 0, nGhostR = 0;
    // launch-size estimates (bucketed; kernels loop, so any estimate is correct) and the step's layout in the local grid
    uint32_t estN = 1, estBnd = 1, estGhost = 1, estIn = 1;
    int xLoL = -1, xHiL = -1;              // owned planes in slab-local (window) grid coordinates; mirrored in dims[D_XLO / D_XHI]
    bool intervalOnDevice = false;
    uint32_t hInterval[2] = {0, 0};
    // the window of global grid planes the local grid covers (slabLayout): fixed while the slab stays inside it
    bool winValid = false;
    int winX0 = 0, winX1 = 0;
    int3 winGmin = {0, 0, 0}, winGdim = {0, 0, 0};
    int64_t windowChanges = 0;
    int planeOffset = 0;                   // local plane 0 in global grid planes
    int gxGlobal = 0;
    uint32_t sentinel = 0;
    int sortBits = 1;
    int64_t exchanges = 0;
    cudaStream_t commStream = nullptr;     // set-up / re-balancing collectives and the NCCL fallback's traffic
    static constexpr int kEvents = 256;
    cudaEvent_t evPool[kEvents] = {};
    int evNext = 0;
    // fixed ghost regions at the top of every per-particle array
    uint32_t ghostBaseL = 0, ghostBaseR = 0, ghostCap = 0;
    // CUDA IPC transport (default): pushes are P2P stores issued by the kernels that produce the data, waits are in-kernel;
    // otherwise NCCL send/recv with one host synchronisation per step
    bool p2p = false;
    bool graphBroken = false;              // a capture of the slab step failed once: stay eager
    SlabPeer peerL, peerR;
    void* xsBuf[2] = {nullptr, nullptr};   // this rank's two x* / velocity allocations (pointer identity across swaps)
    void* velBuf[2] = {nullptr, nullptr};
    uint32_t* flags = nullptr;             // [0]/[1] ghost-exchange epoch published by the left / right rank, [3] CTA
                                           // completion counter, [4]/[5] epoch of the left / right rank's count message
    unsigned long long *dHist = nullptr, *hHist = nullptr;  // re-balancing histogram (+ current bounds)
    size_t histCap = 0;
    int64_t rebalances = 0;            // calls that moved a boundary
    // a measurement in flight (akua_pbf_rebalance_async): enqueued, not yet applied
    cudaEvent_t evRebalance = nullptr;
    bool pendValid = false;
    int pendGx = 0, pendGminGlobalX = 0;
    int64_t pendStep = 0;
    bool lastRebalanceMeasured = false;   // the last akua_pbf_rebalance balanced measured busy time (else the raw work estimate)
    double keepBelow = 1.02;           // akua_pbf_rebalance leaves a partition alone whose heaviest slab is within this of the mean
};

struct akua_pbf_solver {
    int64_t n = 0;         // live (owned) particles
    int64_t capacity = 0;  // array capacity
    akua_pbf_config cfg{};
    akua_corr_params corr{};
    akua_pbf_options opt{};
    int device = 0;
    cudaStream_t stream = nullptr;
    std::string err;

    // SoA state in the solver's current (key-sorted) order
    float4 *pos = nullptr, *posAlt = nullptr;  // x,y,z,mass
    float4 *vel = nullptr, *velAlt = nullptr;  // vx,vy,vz,density
    float4 *xs = nullptr, *xsAlt = nullptr;    // predicted position x*, w = mass (double buffer)
    uint32_t *id = nullptr, *idAlt = nullptr;  // upload index of each particle
    float *density = nullptr, *lambda = nullptr, *omegaLen = nullptr;
    float4 *omega = nullptr, *dpos = nullptr;
    // gather layouts of the sweeps (options.gather_layout; single-GPU path): packed (x*, lambda) written by pass A for
    // pass B, packed (x, |omega|) written by K11 for K12 — both need every particle to have the same mass — and the 32-byte
    // (position, velocity) records K11 / K13 gather from
    float4 *xl = nullptr, *xw = nullptr;
    PosVel* pv = nullptr;
    bool massUniform = false;     // established by the uploads
    float uniformMass = 0.0f;
    bool pvFresh = false;         // pv holds the committed (pos, vel) of this step (written by the final pass B)
    uint32_t *dMassRange = nullptr, *hMassRange = nullptr;   // order-preserving u32 encodings of min / max mass
    // payload in upload order (the solver never reads it: Particle::color / ::size)
    float4* color = nullptr;
    float* size = nullptr;
    // neighbour search
    uint32_t *keysUnsorted = nullptr, *keyA = nullptr, *keyB = nullptr, *valA = nullptr, *valB = nullptr;
    uint32_t *keysSorted = nullptr, *perm = nullptr;  // alias keyA/B, valA/B after a sort
    uint32_t* bucketStart = nullptr;                  // REFERENCE_HASH: tableSize entries
    bool bucketsDirty = false;
    uint32_t bucketsN = 0;                            // number of sorted keys the bucket table currently holds entries for
    bool hashFromUpload = false;                      // Particle::hash of a download comes from the last upload until the next sort
    uint2* cellRange = nullptr;                       // LINEAR_CELL: numCells entries
    int64_t cellCapacity = 0;
    GridParams grid{};
    int keyBits = 1;
    uint32_t *nbrList = nullptr, *nbrCount = nullptr;
    uint32_t nbrStride = 0;
    rsort::Workspace sortWs;
    uint32_t *canonKeys = nullptr, *canonVals = nullptr;   // options.canonical_order: (cell key, index) pairs in id order
    // interchange staging; the AoS transfers are chunked: copy engine on copyStream, (un)pack kernels on the solver's stream
    void* aosStage = nullptr;
    cudaStream_t copyStream = nullptr;
    static constexpr int kXferChunks = 8;
    cudaEvent_t xferEv[kXferChunks + 1] = {};
    float *partSum = nullptr, *partMax = nullptr;
    // bookkeeping
    akua_pbf_counters ctr{};
    bool haveBox = false;
    float lastBoxMin[3] = {0, 0, 0}, lastBoxMax[3] = {0, 0, 0};
    bool timing = false;
    cudaEvent_t ev[PH_COUNT] = {};
    static constexpr int kMaxTimedIters = 16;
    cudaEvent_t evPass[kMaxTimedIters][3] = {};  // before A, between A and B, after B
    int timedIters = 0;
    bool timingValid = false;
    SlabState slab;
    // CUDA-graph replay of the whole step (single-GPU path). The step's kernel arguments depend on which half of each
    // double buffer is current, so up to kGraphSlots graphs are cached, keyed by the pre-step state + parameters.
    struct GraphEntry {
        bool used = false;
        uint64_t key[20] = {};
        cudaGraphExec_t exec = nullptr;
        // host-side state after the step (the pointer swaps and flags the captured calls performed)
        float4 *pos, *posAlt, *vel, *velAlt, *xs, *xsAlt;
        uint32_t *id, *idAlt, *keysSorted, *perm, *slot, *slotAlt;
        bool bucketsDirty;
        int64_t launches, sortPasses, exchanges;
    };
    static constexpr int kGraphSlots = 4;
    GraphEntry graphs[kGraphSlots];
    int graphNext = 0;
    int graphMissStreak = 0;     // consecutive steps whose parameters matched no cached graph
    int graphCooldown = 0;       // steps to run eagerly after a burst of misses (callers that change dt / box every step)
    float accumulator = 0.0f;  // fixed-timestep driver (akua_pbf_advance)
    // launch timeline of one step (akua_pbf_trace_next_step): an event after every launch, on the stream it went to
    struct TraceRec { const char* name; int lane; cudaEvent_t ev; };
    std::vector<TraceRec> trace;
    bool traceArmed = false, tracing = false;
    std::string tracePath;
};

namespace {

#define AK_CUDA(s, call)                                                                               \
    do {                                                                                               \
        cudaError_t e_ = (call);                                                                       \
        if (e_ != cudaSuccess) {                                                                       \
            (s)->err = std::string(#call) + ": " + cudaGetErrorString(e_);                             \
            return AKUA_ERR_CUDA;                                                                      \
        }                                                                                              \
    } while (0)

#define AK_LAUNCH_CHECK(s, name)                                                                       \
    do {                                                                                               \
        cudaError_t e_ = cudaGetLastError();                                                           \
        if (e_ != cudaSuccess) {                                                                       \
            (s)->err = std::string("launch ") + name + ": " + cudaGetErrorString(e_);                  \
            return AKUA_ERR_CUDA;                                                                      \
        }                                                                                              \
        (s)->ctr.kernel_launches++;                                                                    \
        if ((s)->tracing) traceMark((s), name);                                                        \
    } while (0)

// One event after a launch, on the solver's stream (every launch of a step goes there; `lane` is kept for tools that read it).
inline void traceMark(akua_pbf_solver* s, const char* name) {
    cudaEvent_t e = nullptr;
    if (cudaEventCreate(&e) != cudaSuccess) return;
    cudaEventRecord(e, s->stream);
    s->trace.push_back({name, 0, e});
}

// Kernel launch on the solver's current stream. With options.use_pdl the launch carries the programmatic
// stream serialization attribute: the kernel may be scheduled while its predecessor drains (pdl_wait() in every kernel keeps
// the data dependencies those of plain stream order); captured into the step's CUDA graph as programmatic edges.
inline bool usePdl(const akua_pbf_solver* s) { return s->opt.use_pdl != 0; }
template <typename... KArgs, typename... Args>
inline void launchK(const akua_pbf_solver* s, void (*kernel)(KArgs...), uint32_t grid, uint32_t block, Args... args) {
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(grid); cfg.blockDim = dim3(block); cfg.stream = s->stream;
    cudaLaunchAttribute at{};
    if (usePdl(s)) {
        at.id = cudaLaunchAttributeProgrammaticStreamSerialization;
        at.val.programmaticStreamSerializationAllowed = 1;
        cfg.attrs = &at; cfg.numAttrs = 1;
    }
    cudaLaunchKernelEx(&cfg, kernel, KArgs(args)...);
}

// Plain launch (no programmatic serialization) of a kernel that does not begin with pdl_wait(). Every launch of this file goes
// through launchK / launchPlain / rsort::launch, i.e. cudaLaunchKernelEx: no <<<>>> syntax, so that the host code can also be
// compiled by g++ against the SIMT emulator of tests/emu (test infrastructure: the whole C ABI, multi-rank step included, on
// the CPU).
template <typename... KArgs, typename... Args>
inline void launchPlain(cudaStream_t st, void (*kernel)(KArgs...), uint32_t grid, uint32_t block, Args... args) {
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(grid); cfg.blockDim = dim3(block); cfg.stream = st;
    cudaLaunchKernelEx(&cfg, kernel, KArgs(args)...);
}

template <typename T>
cudaError_t dalloc(T** p, size_t count) {
    *p = nullptr;
    if (count == 0) count = 1;
    return cudaMalloc(reinterpret_cast<void**>(p), count * sizeof(T));
}

SphParams makeSph(const akua_pbf_solver* s) { return make_sph_params(s->cfg, s->corr, s->uniformMass); }
inline bool wallOn(const akua_pbf_solver* s) { return s->opt.wall_model == AKUA_WALL_VIRTUAL_FLUID; }
WallParams makeWall(const akua_pbf_solver* s, const SphParams& P, const float* bmin, const float* bmax) {
    WallParams W{};
    const float pi = 3.14159265358979f;
    W.h = P.h; W.rhoW = s->cfg.restDensity;
    W.fCoef = 0.25f * pi * P.poly6Coef;
    W.pH = wall_P(P.h, P.h);
    W.sCoef = std::fabs(P.spikyCoef) * (2.0f * pi / 3.0f);
    W.bmin = make_float3(bmin[0], bmin[1], bmin[2]); W.bmax = make_float3(bmax[0], bmax[1], bmax[2]);
    return W;
}
BoxParams makeBox(const float* bmin, const float* bmax) { return make_box_params(bmin, bmax); }

int bitsFor(uint64_t maxKey) { return bits_for_key(maxKey); }

// AKUA_SLAB_VERBOSE=1: host-side cost of the rare events of a slab run (re-balancing, graph re-capture, cell-table growth) on stderr
inline bool slabVerbose() { static const bool v = [] { const char* e = std::getenv("AKUA_SLAB_VERBOSE"); return e && e[0] == '1'; }(); return v; }
inline double hostMs() { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count(); }

void rememberBox(akua_pbf_solver* s, const float* bmin, const float* bmax) {
    for (int a = 0; a < 3; a++) { s->lastBoxMin[a] = bmin[a]; s->lastBoxMax[a] = bmax[a]; }
    s->haveBox = true;
}

// LINEAR_CELL grid: covers the box plus a two-cell margin (the collision response is a soft clamp, so particles can sit
// slightly outside the box). Particles beyond the margin are clamped into the border cells — still correct (the distance
// test decides), only slower.
int layoutGrid(akua_pbf_solver* s, const float* bmin, const float* bmax) {
    if (s->opt.key_mode != AKUA_KEY_LINEAR_CELL) return AKUA_OK;
    int3 gmin, gdim;
    const int64_t cells = layout_linear_grid(s->cfg.smoothRadius, bmin, bmax, &gmin, &gdim);
    if (cells < 0) { s->err = "box max must exceed box min"; return AKUA_ERR_INVALID; }
    if (cells >= (int64_t)1 << 31) { s->err = "LINEAR_CELL grid too large (>= 2^31 cells)"; return AKUA_ERR_INVALID; }
    if (cells > s->cellCapacity) {  // only when the box grows beyond anything seen so far
        if (s->cellRange) { AK_CUDA(s, cudaStreamSynchronize(s->stream)); AK_CUDA(s, cudaFree(s->cellRange)); }
        s->cellRange = nullptr;
        AK_CUDA(s, dalloc(&s->cellRange, (size_t)cells));
        s->cellCapacity = cells;
    }
    s->grid.gridMin = gmin;
    s->grid.gridDim = gdim;
    s->keyBits = bitsFor((uint64_t)cells - 1);
    s->ctr.num_cells = cells;
    return AKUA_OK;
}

int ensureStage(akua_pbf_solver* s) {
    if (!s->aosStage) AK_CUDA(s, cudaMalloc(&s->aosStage, (size_t)s->capacity * 108));
    if (!s->copyStream) {
        AK_CUDA(s, cudaStreamCreateWithFlags(&s->copyStream, cudaStreamNonBlocking));
        for (auto& e : s->xferEv) AK_CUDA(s, cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    }
    return AKUA_OK;
}
// Chunk c of an n-particle AoS transfer: [first, first + count), chunk boundaries on multiples of the 256-particle tiles of
// k_pack_aos / k_unpack_aos; small transfers are one chunk.
inline int xferChunks(int64_t n) { return (int)std::max<int64_t>(1, std::min<int64_t>(akua_pbf_solver::kXferChunks, n / 131072)); }
inline void xferChunk(int64_t n, int chunks, int c, int64_t* first, int64_t* count) {
    const int64_t per = ((n + chunks - 1) / chunks + 255) / 256 * 256;
    *first = std::min<int64_t>(n, per * c);
    *count = std::min<int64_t>(n, per * (c + 1)) - *first;
}

void mark(akua_pbf_solver* s, Phase p) {
    if (s->timing) cudaEventRecord(s->ev[p], s->stream);
}

// ---------------------------------------------------------------------------------------------------- phases
// x-slab mode: the particle counts live on the device (dims); the host passes an ESTIMATE for the grid and a pointer to the
// exact word. Single GPU: the host's count is exact and the pointer is null.
inline bool slabOn(const akua_pbf_solver* s) { return s->slab.enabled; }
inline uint32_t gridCount(const akua_pbf_solver* s) { return slabOn(s) ? s->slab.estN : (uint32_t)s->n; }
inline const uint32_t* dimWord(const akua_pbf_solver* s, int w) { return slabOn(s) ? s->slab.dims + w : nullptr; }

int phasePredictKey(akua_pbf_solver* s, float dt, bool doPredict, bool doKeys) {
    const uint32_t n = gridCount(s);
    if (n == 0) return AKUA_OK;
    float3 g = make_float3(s->cfg.gravity[0], s->cfg.gravity[1], s->cfg.gravity[2]);
    uint32_t* keys = doKeys ? s->keysUnsorted : nullptr;
    if (s->opt.key_mode == AKUA_KEY_REFERENCE_HASH)
        launchK(s, k_predict_key<KEY_HASH>, gridFor(n), kBlock, s->pos, s->vel, s->xs, keys, n, dimWord(s, D_N), dt, g, s->grid, doPredict ? 1 : 0);
    else
        launchK(s, k_predict_key<KEY_LINEAR>, gridFor(n), kBlock, s->pos, s->vel, s->xs, keys, n, dimWord(s, D_N), dt, g, s->grid, doPredict ? 1 : 0);
    AK_LAUNCH_CHECK(s, "k_predict_key");
    return AKUA_OK;
}

// K4 launch over the owned particles: REFERENCE_HASH scans its buckets; LINEAR_CELL scans row ranges one candidate
// at a time or with the two-phase mask variants of list_build.cuh (options.list_build). Same lists either way.
int launchBuildNeighbours(akua_pbf_solver* s) {
    const uint32_t n = gridCount(s);
    if (n == 0) return AKUA_OK;
#define AK_BUILD(K) launchK(s, K, gridFor(n), kBlock, s->xs, s->keysSorted, s->bucketStart, s->cellRange, n, s->nbrStride, \
            (uint32_t)s->cfg.maxNeighbours, s->nbrList, s->nbrCount, s->grid, s->cfg.smoothRadius, dimWord(s, D_NOWN))
    const bool st = slabOn(s);
    if (s->opt.key_mode == AKUA_KEY_REFERENCE_HASH) AK_BUILD((k_build_neighbours<KEY_HASH, false>));
    // (the mask kernels are always the looping instantiation: it allocates registers better than the single-trip one)
    else if (s->opt.list_build == AKUA_LIST_BUILD_MASK4) AK_BUILD((k_build_neighbours_mask<4, 5, true>));
    else if (s->opt.list_build == AKUA_LIST_BUILD_MASK8) AK_BUILD((k_build_neighbours_mask<8, 4, true>));
    else { if (st) AK_BUILD((k_build_neighbours<KEY_LINEAR, true>)); else AK_BUILD((k_build_neighbours<KEY_LINEAR, false>)); }
#undef AK_BUILD
    AK_LAUNCH_CHECK(s, "k_build_neighbours");
    return AKUA_OK;
}

// The step's sort: (cell key, index) pairs, stable. options.canonical_order first sorts the particles by id, so that ties inside
// a cell are broken by id instead of by last step's order: the result no longer depends on how the particles were distributed
// over GPUs or in which order migrants arrived (bit-identical trajectories at any GPU count; costs four more digit passes).
// `n` sizes the grids; `nPtr` (x-slab mode) is the device-side count.
int sortParticles(akua_pbf_solver* s, uint32_t n, const uint32_t* nPtr, int keyBits) {
    int launches;
    if (!s->opt.canonical_order) {
        launches = rsort::sort_pairs(s->keysUnsorted, s->keyA, s->valA, s->keyB, s->valB, n, keyBits, s->sortWs, s->stream,
                                     &s->keysSorted, &s->perm, usePdl(s), nPtr);
    } else {
        uint32_t *idSorted = nullptr, *byId = nullptr;
        launches = rsort::sort_pairs(s->id, s->keyA, s->valA, s->keyB, s->valB, n, 32, s->sortWs, s->stream, &idSorted, &byId,
                                     usePdl(s), nPtr);
        launchK(s, k_gather_keys, std::max(1u, gridFor(n)), kBlock, s->keysUnsorted, byId, n, nPtr, s->canonKeys, s->canonVals);
        launches += 1 + rsort::sort_pairs(s->canonKeys, s->keyA, s->valA, s->keyB, s->valB, n, keyBits, s->sortWs, s->stream,
                                          &s->keysSorted, &s->perm, usePdl(s), nPtr, s->canonVals);
    }
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) { s->err = std::string("radix sort: ") + cudaGetErrorString(e); return AKUA_ERR_CUDA; }
    s->ctr.kernel_launches += launches;
    s->ctr.sort_passes_last = rsort::passes_for_bits(keyBits) + (s->opt.canonical_order ? 4 : 0);
    if (s->tracing) traceMark(s, "radix sort");
    return AKUA_OK;
}

int phaseSortReorderLists(akua_pbf_solver* s) {
    const uint32_t n = (uint32_t)s->n;
    if (n == 0) return AKUA_OK;
    const bool hash = s->opt.key_mode == AKUA_KEY_REFERENCE_HASH;
    if (hash) {
        if (s->bucketsDirty && s->bucketsN) {   // un-write the entries of the LAST sort (its count, not the current one)
            launchK(s, k_clear_buckets, gridFor(s->bucketsN), kBlock, s->keysSorted, s->bucketsN, s->bucketStart);
            AK_LAUNCH_CHECK(s, "k_clear_buckets");
        }
    } else {
        AK_CUDA(s, cudaMemsetAsync(s->cellRange, 0, (size_t)s->ctr.num_cells * sizeof(uint2), s->stream));
    }
    if (int rcs = sortParticles(s, n, nullptr, s->keyBits)) return rcs;
    mark(s, PH_REORDER);
    if (hash)
        launchK(s, k_reorder_ranges<KEY_HASH>, gridFor(n), kBlock, s->keysSorted, s->perm, n, nullptr, s->pos, s->vel, s->xs, s->id,
            s->posAlt, s->velAlt, s->xsAlt, s->idAlt, s->bucketStart, s->cellRange, nullptr, nullptr);
    else
        launchK(s, k_reorder_ranges<KEY_LINEAR>, gridFor(n), kBlock, s->keysSorted, s->perm, n, nullptr, s->pos, s->vel, s->xs, s->id,
            s->posAlt, s->velAlt, s->xsAlt, s->idAlt, s->bucketStart, s->cellRange, nullptr, nullptr);
    AK_LAUNCH_CHECK(s, "k_reorder_ranges");
    std::swap(s->pos, s->posAlt);
    std::swap(s->vel, s->velAlt);
    std::swap(s->xs, s->xsAlt);
    std::swap(s->id, s->idAlt);
    s->bucketsDirty = hash;
    s->bucketsN = hash ? n : 0;
    s->hashFromUpload = false;
    mark(s, PH_LISTS);
    return launchBuildNeighbours(s);
}

// ---- slab-mode plumbing used by the sweeps below (definitions in pbf_slab.inl) ----
int slabAgreeMass(akua_pbf_solver* s);
template <typename T> PeerPush slabPush(const akua_pbf_solver* s, T* arr);        // null pushes unless the p2p transport is on
// In-kernel wait (exchange index waitIdx, -1 none) / signal (signalIdx, -1 none) block of a boundary launch (p2p transport).
HaloSync slabHalo(const akua_pbf_solver* s, int waitIdx, int signalIdx);
template <typename T> int slabNcclPlanes(akua_pbf_solver* s, T* arr);             // NCCL fallback: blocking plane exchange
struct SweepSpans { Span interior, boundary, fused; uint32_t gridInterior, gridBoundary, gridFused; };

// ---- sweep launchers over an index span ----
// Gather layouts (akua_pbf_options::gather_layout). The 32-byte records are single-GPU only.
inline bool usePack(const akua_pbf_solver* s) {
    const int g = s->opt.gather_layout;
    // x-slab mode: massUniform is then the verdict ALL ranks agreed on (slabAgreeMass), because the packed arrays are
    // what the halo exchanges carry
    return s->massUniform && s->xl && (g == AKUA_GATHER_AUTO || g == AKUA_GATHER_PACKED || g == AKUA_GATHER_PACKED_RECORDS);
}
inline bool useRec(const akua_pbf_solver* s) {
    const int g = s->opt.gather_layout;
    return !s->slab.enabled && (g == AKUA_GATHER_RECORDS || g == AKUA_GATHER_PACKED_RECORDS);
}
SweepSpans sweepSpans(const akua_pbf_solver* s) {
    SweepSpans sp;
    if (!slabOn(s)) {
        sp.interior = fullSpan((uint32_t)s->n);
        sp.boundary = Span{0, 0, 0, 0, nullptr, SPAN_FIXED};
        sp.gridInterior = sweepGrid((uint64_t)s->n); sp.gridBoundary = 0;
        sp.fused = sp.boundary; sp.gridFused = 0;
    } else {
        sp.interior = Span{0, 0, 0, 0, s->slab.dims, SPAN_INTERIOR};
        sp.boundary = Span{0, 0, 0, 0, s->slab.dims, SPAN_BOUNDARY};
        sp.gridInterior = sweepGrid(s->slab.estN); sp.gridBoundary = sweepGrid(s->slab.estBnd);
        // CUDA-IPC transport: ONE launch per sweep, boundary CTAs first (sweep_cta in pbf_kernels.cuh)
        sp.fused = Span{0, 0, 0, 0, s->slab.dims, SPAN_FUSED};
        sp.gridFused = std::max(2u, sp.gridInterior + sp.gridBoundary);
    }
    return sp;
}

// Every sweep exists in two instantiations: SLAB = false is the single-GPU kernel (one particle per thread, no halo code at
// all), SLAB = true loops over a device-resolved span, waits for / publishes halo epochs and pushes boundary results.
int launchPassA(akua_pbf_solver* s, Span sp, uint32_t grid, const SphParams& P, bool push = false, const HaloSync& hs = HaloSync{}) {
    if (!grid) return AKUA_OK;
    float4* xl = usePack(s) ? s->xl : nullptr;
    const PeerPush pl = !push ? PeerPush{} : (xl ? slabPush(s, xl) : slabPush(s, s->lambda));
#define AK_A(F, S) launchK(s, k_density_lambda<F, S>, grid, kSweepBlock, s->xs, s->nbrList, s->nbrCount, s->nbrStride, sp, s->density, s->lambda, xl, P, pl, hs)
    if (slabOn(s)) { if (s->opt.fast_math) AK_A(true, true); else AK_A(false, true); }
    else           { if (s->opt.fast_math) AK_A(true, false); else AK_A(false, false); }
#undef AK_A
    AK_LAUNCH_CHECK(s, "k_density_lambda");
    return AKUA_OK;
}
int launchPassB(akua_pbf_solver* s, Span sp, uint32_t grid, const SphParams& P, const BoxParams& B, bool fin, float dt, bool push = false,
                const HaloSync& hs = HaloSync{}) {
    if (!grid) return AKUA_OK;
    const PeerPush px = push ? slabPush(s, s->xsAlt) : PeerPush{};
    const PeerPush pv = (push && fin) ? slabPush(s, s->vel) : PeerPush{};
    PosVel* rec = (fin && useRec(s)) ? s->pv : nullptr;
#define AK_DELTA(F, L, K, C, S) launchK(s, k_delta_apply<F, L, K, C, S>, grid, kSweepBlock, s->xs, s->xsAlt, s->lambda, s->xl, \
            s->nbrList, s->nbrCount, s->nbrStride, sp, P, B, s->dpos, s->pos, s->vel, s->density, rec, dt, px, pv, hs)
#define AK_DELTA_S(F, L, K, C) do { if (slabOn(s)) AK_DELTA(F, L, K, C, true); else AK_DELTA(F, L, K, C, false); } while (0)
#define AK_DELTA_C(F, L, K) do { if (P.corrNIsFour) AK_DELTA_S(F, L, K, true); else AK_DELTA_S(F, L, K, false); } while (0)
#define AK_DELTA_L(F, K) do { if (fin) AK_DELTA_C(F, true, K); else AK_DELTA_C(F, false, K); } while (0)
    if (usePack(s)) { if (s->opt.fast_math) AK_DELTA_L(true, true); else AK_DELTA_L(false, true); }
    else            { if (s->opt.fast_math) AK_DELTA_L(true, false); else AK_DELTA_L(false, false); }
#undef AK_DELTA_L
#undef AK_DELTA_C
#undef AK_DELTA_S
#undef AK_DELTA
    AK_LAUNCH_CHECK(s, "k_delta_apply");
    if (rec) s->pvFresh = true;
    return AKUA_OK;
}
int launchVorticity(akua_pbf_solver* s, Span sp, uint32_t grid, const SphParams& P, bool push = false, const HaloSync& hs = HaloSync{}) {
    if (!grid) return AKUA_OK;
    float4* xw = usePack(s) ? s->xw : nullptr;
    const PeerPush pw = !push ? PeerPush{} : (xw ? slabPush(s, xw) : slabPush(s, s->omegaLen));
#define AK_VORT(F, R, S) launchK(s, k_vorticity<F, R, S>, grid, kSweepBlock, s->xs, s->vel, s->pv, s->nbrList, s->nbrCount, \
            s->nbrStride, sp, s->omega, s->omegaLen, xw, P, pw, hs)
    if (slabOn(s)) { if (s->opt.fast_math) AK_VORT(true, false, true); else AK_VORT(false, false, true); }
    else if (useRec(s)) { if (s->opt.fast_math) AK_VORT(true, true, false); else AK_VORT(false, true, false); }
    else           { if (s->opt.fast_math) AK_VORT(true, false, false); else AK_VORT(false, false, false); }
#undef AK_VORT
    AK_LAUNCH_CHECK(s, "k_vorticity");
    return AKUA_OK;
}
int launchConfinement(akua_pbf_solver* s, Span sp, uint32_t grid, const SphParams& P, float dt, bool push = false, const HaloSync& hs = HaloSync{}) {
    if (!grid) return AKUA_OK;
    const PeerPush pv = push ? slabPush(s, s->vel) : PeerPush{};
    PosVel* rec = useRec(s) ? s->pv : nullptr;
#define AK_CONF(F, K, S) launchK(s, k_confinement<F, K, S>, grid, kSweepBlock, s->xs, s->omega, s->omegaLen, s->xw, s->density, \
            s->nbrList, s->nbrCount, s->nbrStride, sp, s->vel, rec, P, dt, s->cfg.vorticityEpsilon, pv, hs)
#define AK_CONF_S(F, K) do { if (slabOn(s)) AK_CONF(F, K, true); else AK_CONF(F, K, false); } while (0)
    if (usePack(s)) { if (s->opt.fast_math) AK_CONF_S(true, true); else AK_CONF_S(false, true); }
    else            { if (s->opt.fast_math) AK_CONF_S(true, false); else AK_CONF_S(false, false); }
#undef AK_CONF_S
#undef AK_CONF
    AK_LAUNCH_CHECK(s, "k_confinement");
    return AKUA_OK;
}
int launchXsph(akua_pbf_solver* s, Span sp, uint32_t grid, const SphParams& P, const HaloSync& hs = HaloSync{}) {
    if (!grid) return AKUA_OK;
#define AK_X(R, S) launchK(s, k_xsph<R, S>, grid, kSweepBlock, s->xs, s->vel, s->pv, s->nbrList, s->nbrCount, s->nbrStride, sp, s->velAlt, P, s->cfg.viscosity, hs)
    if (slabOn(s)) AK_X(false, true);
    else if (useRec(s)) AK_X(true, false);
    else AK_X(false, false);
#undef AK_X
    AK_LAUNCH_CHECK(s, "k_xsph");
    return AKUA_OK;
}

// Exchange indices of a slab step (the epoch of exchange e is dims[D_EPOCH] + e + 1; all ranks run the same sequence):
//   0 count message + migration records      1 x* after the reorder (for the list build)
//   2 + 2 it : (x*, lambda) after pass A of iteration it         3 + 2 it : x* (last iteration: + v, rho) after pass B
//   solverIterations == 0: 2 = v after the stand-alone commit
//   post-solve: pb = (x, |omega|) after K11, pb + 1 = v after K12, with pb = 2 + 2 I (3 when I == 0)
inline int slabLastXIdx(int iterations) { return iterations > 0 ? 1 + 2 * iterations : 2; }
inline int slabPostBaseIdx(int iterations) { return iterations > 0 ? 2 + 2 * iterations : 3; }
inline int slabExchangesPerStep(int iterations) { return slabPostBaseIdx(iterations) + 2; }

// `commit`: fold K9+K10 into the last iteration's pass B (whole-step path). dt is only read when commit is set.
// Slab mode: a sweep covers the slab interior (needs no ghost data) and its two boundary planes. With the CUDA-IPC transport
// both are ONE launch whose first CTAs take the boundary planes: they wait IN-KERNEL for the ghosts they read, store their
// results straight into the neighbours' ghost regions (P2P stores over NVLink) and the last of them publishes the epoch,
// while the other CTAs sweep the interior — no exchange sits on the critical path, and the step is a linear chain of
// launches on one stream (CUDA-graph replay, programmatic dependent launch between all of them).
// (NCCL fallback: interior launch, boundary launch, blocking send/recv.)
int phaseSolve(akua_pbf_solver* s, int iterations, const float* bmin, const float* bmax, bool commit, float dt,
               bool* committed) {
    *committed = false;
    const bool slabMode = s->slab.enabled;
    if (s->n == 0 && !slabMode) return AKUA_OK;  // a slab rank without particles still takes part in the exchanges
    int rc;
    const SphParams P = makeSph(s);
    const BoxParams B = makeBox(bmin, bmax);
    const SweepSpans sp = sweepSpans(s);
    const bool p2p = s->slab.p2p;
    s->timedIters = 0;
    for (int it = 0; it < iterations; it++) {
        const bool timeIt = s->timing && it < akua_pbf_solver::kMaxTimedIters;
        const bool fin = commit && !wallOn(s) && it == iterations - 1;   // (the wall model inserts a kernel between pass B and the commit)
        if (timeIt) cudaEventRecord(s->evPass[it][0], s->stream);
        if (slabMode && p2p) {
            // ---- CUDA-IPC transport: one launch per sweep. Its first CTAs take the boundary planes: they wait in-kernel for the
            // ghosts this sweep reads, store their results straight into the neighbours' ghost regions (P2P stores over NVLink)
            // and the last of them publishes the epoch; the remaining CTAs sweep the interior meanwhile. Pass A needs the
            // ghosts' x* and sends (x*, lambda); pass B needs those and sends the corrected x* (after the commit also v, rho).
            if ((rc = launchPassA(s, sp.fused, sp.gridFused, P, true, slabHalo(s, 1 + 2 * it, 2 + 2 * it)))) return rc;
            if (timeIt) cudaEventRecord(s->evPass[it][1], s->stream);
            if ((rc = launchPassB(s, sp.fused, sp.gridFused, P, B, fin, dt, true, slabHalo(s, 2 + 2 * it, 3 + 2 * it)))) return rc;
        } else if (slabMode) {
            // ---- NCCL fallback: interior launch, boundary launch, then a blocking send/recv of the planes, all in stream order
            if ((rc = launchPassA(s, sp.interior, sp.gridInterior, P))) return rc;
            if ((rc = launchPassA(s, sp.boundary, sp.gridBoundary, P))) return rc;
            if ((rc = usePack(s) ? slabNcclPlanes(s, s->xl) : slabNcclPlanes(s, s->lambda))) return rc;
            if (timeIt) cudaEventRecord(s->evPass[it][1], s->stream);
            if ((rc = launchPassB(s, sp.interior, sp.gridInterior, P, B, fin, dt))) return rc;
            if ((rc = launchPassB(s, sp.boundary, sp.gridBoundary, P, B, fin, dt))) return rc;
            if ((rc = slabNcclPlanes(s, s->xsAlt))) return rc;
            if (fin && (rc = slabNcclPlanes(s, s->vel))) return rc;
        } else {
            if ((rc = launchPassA(s, sp.interior, sp.gridInterior, P))) return rc;
            if (wallOn(s)) {   // opt-in wall model: near-wall particles get rho, grad C and lambda with the virtual half-spaces
                const WallParams W = makeWall(s, P, bmin, bmax);
                launchK(s, k_wall_lambda, sweepGrid((uint64_t)s->n), kSweepBlock, s->xs, s->nbrList, s->nbrCount, s->nbrStride, (uint32_t)s->n,
                        s->density, s->lambda, usePack(s) ? s->xl : nullptr, P, W);
                AK_LAUNCH_CHECK(s, "k_wall_lambda");
            }
            if (timeIt) cudaEventRecord(s->evPass[it][1], s->stream);
            if ((rc = launchPassB(s, sp.interior, sp.gridInterior, P, B, fin, dt))) return rc;
            if (wallOn(s)) {   // ... and the wall's share of delta-p (pass B never commits in this mode: see stepEager)
                const WallParams W = makeWall(s, P, bmin, bmax);
                launchK(s, k_wall_dp, sweepGrid((uint64_t)s->n), kSweepBlock, s->xs, s->xsAlt, s->lambda, s->dpos, (uint32_t)s->n, P, B, W);
                AK_LAUNCH_CHECK(s, "k_wall_dp");
            }
        }
        if (timeIt) { cudaEventRecord(s->evPass[it][2], s->stream); s->timedIters = it + 1; }
        std::swap(s->xs, s->xsAlt);
        if (fin) *committed = true;
    }
    return AKUA_OK;
}

int phaseUpdate(akua_pbf_solver* s, float dt) {
    const uint32_t n = gridCount(s);
    if (n == 0) return AKUA_OK;
    launchK(s, k_update, gridFor(n), kBlock, s->xs, s->pos, s->vel, s->density, n, dt, dimWord(s, D_NOWN));
    AK_LAUNCH_CHECK(s, "k_update");
    return AKUA_OK;
}
int phaseDamping(akua_pbf_solver* s, const float* bmin, const float* bmax) {
    const uint32_t n = gridCount(s);
    if (n == 0) return AKUA_OK;
    launchK(s, k_damping, gridFor(n), kBlock, s->pos, s->vel, n, makeBox(bmin, bmax), dimWord(s, D_NOWN));
    AK_LAUNCH_CHECK(s, "k_damping");
    return AKUA_OK;
}
// `iterations`: only read in slab mode (it fixes the exchange indices of the post-solve sweeps)
int phasePost(akua_pbf_solver* s, float dt, int iterations) {
    const bool slabMode = s->slab.enabled;
    if (s->n == 0 && !slabMode) return AKUA_OK;
    int rc;
    const SphParams P = makeSph(s);
    const SweepSpans sp = sweepSpans(s);
    if (!slabMode) {
        const uint32_t n = (uint32_t)s->n;
        if (useRec(s) && !s->pvFresh) {   // committed outside the fused final pass B (phase-level API, 0 iterations)
            launchK(s, k_build_posvel, gridFor(n), kBlock, s->xs, s->vel, n, s->pv);
            AK_LAUNCH_CHECK(s, "k_build_posvel");
        }
        s->pvFresh = false;               // K12 / K13 leave the records behind the committed state
        if ((rc = launchVorticity(s, sp.interior, sp.gridInterior, P))) return rc;
        if ((rc = launchConfinement(s, sp.interior, sp.gridInterior, P, dt))) return rc;
        if ((rc = launchXsph(s, sp.interior, sp.gridInterior, P))) return rc;
        std::swap(s->vel, s->velAlt);
        return AKUA_OK;
    }
    const bool p2p = s->slab.p2p;
    const int lastX = slabLastXIdx(iterations), pb = slabPostBaseIdx(iterations);
    if (p2p) {
        // one fused launch per sweep (boundary CTAs first): K11 needs the ghosts' final x*, v and sends (x, |omega|); K12 needs
        // those and sends the post-confinement v; K13 needs that
        if ((rc = launchVorticity(s, sp.fused, sp.gridFused, P, true, slabHalo(s, lastX, pb)))) return rc;
        if ((rc = launchConfinement(s, sp.fused, sp.gridFused, P, dt, true, slabHalo(s, pb, pb + 1)))) return rc;
        if ((rc = launchXsph(s, sp.fused, sp.gridFused, P, slabHalo(s, pb + 1, -1)))) return rc;
    } else {
        if ((rc = launchVorticity(s, sp.interior, sp.gridInterior, P))) return rc;
        if ((rc = launchVorticity(s, sp.boundary, sp.gridBoundary, P))) return rc;
        if ((rc = usePack(s) ? slabNcclPlanes(s, s->xw) : slabNcclPlanes(s, s->omegaLen))) return rc;
        if ((rc = launchConfinement(s, sp.interior, sp.gridInterior, P, dt))) return rc;
        if ((rc = launchConfinement(s, sp.boundary, sp.gridBoundary, P, dt))) return rc;
        if ((rc = slabNcclPlanes(s, s->vel))) return rc;
        if ((rc = launchXsph(s, sp.interior, sp.gridInterior, P))) return rc;
        if ((rc = launchXsph(s, sp.boundary, sp.gridBoundary, P))) return rc;
    }
    std::swap(s->vel, s->velAlt);
    return AKUA_OK;
}

#include "pbf_slab.inl"

int stepEager(akua_pbf_solver* s, float dt, int iterations, const float* bmin, const float* bmax);

void graphKey(const akua_pbf_solver* s, float dt, int iterations, const float* bmin, const float* bmax, uint64_t key[20]) {
    auto f2 = [](float a, float b) { uint32_t x, y; std::memcpy(&x, &a, 4); std::memcpy(&y, &b, 4); return ((uint64_t)x << 32) | y; };
    key[0] = (uint64_t)s->pos; key[1] = (uint64_t)s->vel; key[2] = (uint64_t)s->xs; key[3] = (uint64_t)s->id;
    key[4] = (uint64_t)s->keysSorted; key[5] = slabOn(s) ? 0 : (uint64_t)s->n; key[6] = ((uint64_t)iterations << 1) | (s->bucketsDirty ? 1 : 0);
    key[7] = f2(dt, s->cfg.gravity[0]); key[8] = f2(s->cfg.gravity[1], s->cfg.gravity[2]);
    key[9] = f2(bmin[0], bmin[1]); key[10] = f2(bmin[2], bmax[0]); key[11] = f2(bmax[1], bmax[2]);
    key[12] = (uint64_t)s->cellRange; key[13] = (uint64_t)s->perm; key[14] = (uint64_t)s->opt.fast_math;
    uint32_t mbits; std::memcpy(&mbits, &s->uniformMass, 4);
    key[15] = ((uint64_t)mbits << 8) | (usePack(s) ? 1u : 0u) | (useRec(s) ? 2u : 0u) | (usePdl(s) ? 4u : 0u) | (slabOn(s) ? 8u : 0u) |
              (wallOn(s) ? 16u : 0u);
    // x-slab mode: the slab interval fixes the local grid, the (bucketed) size estimates fix the launch grids
    const SlabState& sl = s->slab;
    key[16] = slabOn(s) ? (((uint64_t)(uint32_t)sl.winX0 << 32) | (uint32_t)sl.winX1) : 0;   // the window, not the interval inside it
    key[17] = slabOn(s) ? (((uint64_t)sl.estN << 32) | sl.estBnd) : 0;
    key[18] = slabOn(s) ? (((uint64_t)sl.estGhost << 32) | sl.estIn) : 0;
    key[19] = slabOn(s) ? (uint64_t)sl.slot : (uint64_t)s->bucketsN;
}

// akua_pbf_trace_next_step: runs one step eagerly with an event after every launch and appends the timeline (completion time
// of each launch relative to the start of the step, per stream) to the armed file, one JSON object per line.
template <typename Body>
int tracedStep(akua_pbf_solver* s, Body& body) {
    s->traceArmed = false;
    s->trace.clear();
    s->tracing = true;
    traceMark(s, "step begin");
    const int rc = body();
    s->tracing = false;
    cudaStreamSynchronize(s->stream);
    FILE* f = std::fopen(s->tracePath.c_str(), "a");
    if (f && !s->trace.empty()) {
        float prev[2] = {0.0f, 0.0f};
        for (size_t k = 1; k < s->trace.size(); k++) {
            float t = 0.0f;
            cudaEventElapsedTime(&t, s->trace[0].ev, s->trace[k].ev);
            const int lane = s->trace[k].lane;
            std::fprintf(f, "{\"rank\": %d, \"step\": %lld, \"seq\": %zu, \"lane\": %d, \"name\": \"%s\", \"end_ms\": %.4f, \"since_prev_on_lane_ms\": %.4f}\n",
                         s->slab.rank, (long long)s->ctr.steps, k, lane, s->trace[k].name, t, t - prev[lane]);
            prev[lane] = t;
        }
    }
    if (f) std::fclose(f);
    for (auto& r : s->trace) cudaEventDestroy(r.ev);
    s->trace.clear();
    cudaGetLastError();
    return rc;
}

int stepImpl(akua_pbf_solver* s, float dt, int iterations, const float* bmin, const float* bmax) {
    if (!s || !bmin || !bmax) return AKUA_ERR_INVALID;
    if (iterations < 0) { s->err = "solverIterations must be >= 0"; return AKUA_ERR_INVALID; }
    AK_CUDA(s, cudaSetDevice(s->device));
    rememberBox(s, bmin, bmax);
    const bool slabMode = s->slab.enabled;
    int rc;
    if (slabMode) {
        if ((rc = slabCheckError(s))) return rc;     // a device-side error of an earlier step (pinned mirror, no sync)
        if ((rc = slabLayout(s, bmin, bmax))) return rc;
        slabUpdateEstimates(s);
    } else {
        rc = layoutGrid(s, bmin, bmax);  // may (re)allocate the cell table: stays outside any capture
        if (rc) return rc;
    }
    auto body = [&]() {
        return slabMode ? stepSlabBody(s, dt, iterations, bmin, bmax) : stepEager(s, dt, iterations, bmin, bmax);
    };
    if (s->traceArmed) return tracedStep(s, body);
    const bool graphable = s->opt.use_graph && !s->timing && (slabMode ? (s->slab.p2p && !s->slab.graphBroken) : s->n != 0);
    if (!graphable) return body();
    uint64_t key[20];
    graphKey(s, dt, iterations, bmin, bmax, key);
    akua_pbf_solver::GraphEntry* hit = nullptr;
    for (auto& g : s->graphs)
        if (g.used && std::memcmp(g.key, key, sizeof(key)) == 0) { hit = &g; break; }
    if (hit) s->graphMissStreak = 0;
    else if (s->graphCooldown > 0) { s->graphCooldown--; return body(); }
    else if (++s->graphMissStreak > 6) {
        // parameters keep changing (adaptive dt, moving box): capturing a graph per step costs more than it saves
        s->graphMissStreak = 0; s->graphCooldown = 64;
        return body();
    }
    if (!hit) {
        // capture this step (the launches below are recorded, not executed), instantiate, then replay it
        akua_pbf_solver::GraphEntry& g = s->graphs[s->graphNext];
        s->graphNext = (s->graphNext + 1) % akua_pbf_solver::kGraphSlots;
        if (g.used && g.exec) { cudaGraphExecDestroy(g.exec); g.exec = nullptr; }
        g.used = false;
        // pre-state, restored before the replay so that capture + replay advance the state exactly once
        float4 *pos = s->pos, *posAlt = s->posAlt, *vel = s->vel, *velAlt = s->velAlt, *xs = s->xs, *xsAlt = s->xsAlt;
        uint32_t *id = s->id, *idAlt = s->idAlt, *ks = s->keysSorted, *pm = s->perm, *sl0 = s->slab.slot, *sl1 = s->slab.slotAlt;
        const bool dirty = s->bucketsDirty, hfu0 = s->hashFromUpload;
        const uint32_t bucketsN0 = s->bucketsN;
        const int64_t launches0 = s->ctr.kernel_launches, steps0 = s->ctr.steps, exch0 = s->slab.exchanges;
        auto restore = [&]() {
            s->pos = pos; s->posAlt = posAlt; s->vel = vel; s->velAlt = velAlt; s->xs = xs; s->xsAlt = xsAlt;
            s->id = id; s->idAlt = idAlt; s->keysSorted = ks; s->perm = pm; s->bucketsDirty = dirty;
            s->slab.slot = sl0; s->slab.slotAlt = sl1; s->bucketsN = bucketsN0; s->hashFromUpload = hfu0;
            s->ctr.kernel_launches = launches0; s->ctr.steps = steps0; s->slab.exchanges = exch0;
        };
        const double tCap0 = hostMs();
        if (cudaStreamBeginCapture(s->stream, cudaStreamCaptureModeThreadLocal) != cudaSuccess) {
            cudaGetLastError();
            s->opt.use_graph = 0;     // a stream that cannot be captured (or a runtime without graphs): step eagerly from now on
            return body();
        }
        rc = body();
        cudaGraph_t graph = nullptr;
        cudaError_t ce = cudaStreamEndCapture(s->stream, &graph);
        if (rc == AKUA_OK && ce == cudaSuccess && graph) {
            ce = cudaGraphInstantiate(&g.exec, graph, 0);
            if (ce != cudaSuccess) g.exec = nullptr;
        }
        if (graph) cudaGraphDestroy(graph);
        if (rc != AKUA_OK || ce != cudaSuccess || !g.exec) {
            // nothing was executed: put the host-side state back where it was before the capture
            const std::string why = rc != AKUA_OK ? s->err : std::string("graph capture: ") + cudaGetErrorString(ce);
            cudaGetLastError();
            restore();
            if (slabMode && rc == AKUA_OK) {   // e.g. a driver that cannot capture this multi-stream pattern: run eagerly from now on
                s->slab.graphBroken = true;
                return body();
            }
            s->err = why;
            return rc != AKUA_OK ? rc : AKUA_ERR_CUDA;
        }
        std::memcpy(g.key, key, sizeof(key));
        g.pos = s->pos; g.posAlt = s->posAlt; g.vel = s->vel; g.velAlt = s->velAlt; g.xs = s->xs; g.xsAlt = s->xsAlt;
        g.id = s->id; g.idAlt = s->idAlt; g.keysSorted = s->keysSorted; g.perm = s->perm; g.bucketsDirty = s->bucketsDirty;
        g.slot = s->slab.slot; g.slotAlt = s->slab.slotAlt;
        g.launches = s->ctr.kernel_launches - launches0; g.sortPasses = s->ctr.sort_passes_last;
        g.exchanges = s->slab.exchanges - exch0;
        g.used = true;
        restore();
        hit = &g;
        if (slabVerbose()) std::fprintf(stderr, "[akua rank %d] step %lld: graph captured + instantiated in %.2f ms (host)\n", s->slab.rank,
                                        (long long)s->ctr.steps, hostMs() - tCap0);
    }
    AK_CUDA(s, cudaGraphLaunch(hit->exec, s->stream));
    s->pos = hit->pos; s->posAlt = hit->posAlt; s->vel = hit->vel; s->velAlt = hit->velAlt; s->xs = hit->xs; s->xsAlt = hit->xsAlt;
    s->id = hit->id; s->idAlt = hit->idAlt; s->keysSorted = hit->keysSorted; s->perm = hit->perm; s->bucketsDirty = hit->bucketsDirty;
    s->slab.slot = hit->slot; s->slab.slotAlt = hit->slotAlt;
    s->bucketsN = hit->bucketsDirty ? (uint32_t)s->n : 0; s->hashFromUpload = false;
    s->ctr.kernel_launches += hit->launches; s->ctr.sort_passes_last = hit->sortPasses; s->ctr.steps++;
    s->slab.exchanges += hit->exchanges;
    s->ctr.graph_replays++;
    return AKUA_OK;
}

int stepEager(akua_pbf_solver* s, float dt, int iterations, const float* bmin, const float* bmax) {
    int rc;
    mark(s, PH_PREDICT);
    if ((rc = phasePredictKey(s, dt, true, true))) return rc;            // PBFSolver.cpp:30 (+ K2)
    mark(s, PH_SORT);
    if ((rc = phaseSortReorderLists(s))) return rc;                       // PBFSolver.cpp:33
    mark(s, PH_SOLVE);
    bool committed = false;
    if ((rc = phaseSolve(s, iterations, bmin, bmax, true, dt, &committed))) return rc;  // PBFSolver.cpp:45
    if (!committed) {                                                     // solverIterations == 0
        if ((rc = phaseUpdate(s, dt))) return rc;                         // PBFSolver.cpp:61
        if ((rc = phaseDamping(s, bmin, bmax))) return rc;                // PBFSolver.cpp:64
    }
    mark(s, PH_POST);
    if ((rc = phasePost(s, dt, iterations))) return rc;                   // PBFSolver.cpp:67
    mark(s, PH_END);
    s->timingValid = s->timing;
    s->ctr.steps++;
    return AKUA_OK;
}

// Uniform-mass detection for the packed gather layouts: min / max of pos.w through an order-preserving u32 encoding,
// fetched with the copy the upload already waits for. A NaN mass encodes above +inf on one side only -> "not uniform".
int massRangeAsync(akua_pbf_solver* s) {
    const uint32_t n = (uint32_t)s->n;
    AK_CUDA(s, cudaMemsetAsync(s->dMassRange, 0xff, sizeof(uint32_t), s->stream));
    AK_CUDA(s, cudaMemsetAsync(s->dMassRange + 1, 0, sizeof(uint32_t), s->stream));
    launchPlain(s->stream, k_mass_range, std::min<uint32_t>(gridFor(n), 148 * 8), kBlock, s->pos, n, s->dMassRange);
    AK_LAUNCH_CHECK(s, "k_mass_range");
    AK_CUDA(s, cudaMemcpyAsync(s->hMassRange, s->dMassRange, 2 * sizeof(uint32_t), cudaMemcpyDeviceToHost, s->stream));
    return AKUA_OK;
}
void massRangeFinish(akua_pbf_solver* s) {   // after the stream synchronisation
    const uint32_t lo = s->hMassRange[0], hi = s->hMassRange[1];
    s->massUniform = s->n > 0 && lo == hi;
    const uint32_t bits = (lo & 0x80000000u) ? (lo ^ 0x80000000u) : ~lo;   // decode
    std::memcpy(&s->uniformMass, &bits, 4);
}

// Every upload replaces the particle set: the live count follows it and, in x-slab mode, the device-side bookkeeping (owned
// count, payload slots) is reset to match.
int beforeUpload(akua_pbf_solver* s, int64_t n) {
    s->n = n;
    s->hashFromUpload = true;
    return slabResetCounts(s);
}
// Particle::hash of an AoS export: the sorted keys of the current order (what the reference's struct holds after a step), or
// the uploaded hash field while no sort has happened since the upload.
const uint32_t* hashField(const akua_pbf_solver* s) { return s->hashFromUpload ? s->keysUnsorted : s->keysSorted; }

}  // namespace

// ======================================================================================================== C ABI
extern "C" {

int akua_pbf_abi_version(void) { return AKUA_PBF_ABI_VERSION; }

void akua_pbf_default_config(akua_pbf_config* c) {  // PBFConfig.h:18-29
    if (!c) return;
    c->restDensity = 7600.0f; c->particle_spacing = 0.05f; c->smoothRadius = 0.1f; c->spatialHashCellSize = 0.1f;
    c->relaxation = 600.0f; c->vorticityEpsilon = 0.00001f; c->viscosity = 0.01f; c->maxNeighbours = 128;
    c->solverIterations = 4; c->gravity[0] = 0.0f; c->gravity[1] = -9.8f; c->gravity[2] = 0.0f;
}
void akua_pbf_default_corr(akua_corr_params* c) {  // PBFConfig.h:10-15
    if (!c) return;
    c->enabled = 1; c->k = 0.0001f; c->n = 4.0f; c->delta_q = 0.03f;
}
void akua_pbf_default_options(akua_pbf_options* o) {
    if (!o) return;
    std::memset(o, 0, sizeof(*o));
    o->key_mode = AKUA_KEY_LINEAR_CELL; o->device = 0; o->use_graph = 1; o->fast_math = 1; o->capacity_factor = 1.0f;
    o->use_pdl = 1;
    o->list_build = AKUA_LIST_BUILD_MASK4;
}

int akua_pbf_create(akua_pbf_solver** out, int64_t numParticles, const akua_pbf_config* cfg, const akua_corr_params* corr,
                    const akua_pbf_options* opt) {
    if (!out) return AKUA_ERR_INVALID;
    *out = nullptr;
    if (!cfg || !corr || numParticles < 0 || numParticles >= ((int64_t)1 << 31)) return AKUA_ERR_INVALID;
    if (!(cfg->smoothRadius > 0.0f) || cfg->maxNeighbours <= 0 || cfg->solverIterations < 0) return AKUA_ERR_INVALID;
    if (cfg->spatialHashCellSize != cfg->smoothRadius) return AKUA_ERR_INVALID;  // see akua_pbf.h
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0) return AKUA_ERR_NO_DEVICE;
    akua_pbf_solver* s = new (std::nothrow) akua_pbf_solver();
    if (!s) return AKUA_ERR_ALLOC;
    s->cfg = *cfg;
    s->corr = *corr;
    if (opt) s->opt = *opt; else akua_pbf_default_options(&s->opt);
    if (s->opt.capacity_factor < 1.0f) s->opt.capacity_factor = 1.0f;
    s->device = s->opt.device;
    s->n = numParticles;
    s->capacity = (int64_t)std::ceil((double)numParticles * s->opt.capacity_factor);
    if (s->capacity < 1) s->capacity = 1;
    *out = s;  // from here on the caller can read last_error and must destroy
    if (s->device < 0 || s->device >= ndev) { s->err = "invalid device ordinal"; return AKUA_ERR_INVALID; }
    AK_CUDA(s, cudaSetDevice(s->device));
    {
        cudaDeviceProp prop;
        AK_CUDA(s, cudaGetDeviceProperties(&prop, s->device));
        if (prop.major < 10) { s->err = "this library is built for sm_100a (Blackwell) only"; return AKUA_ERR_NO_DEVICE; }
    }
    AK_CUDA(s, cudaStreamCreateWithFlags(&s->stream, cudaStreamNonBlocking));
    const size_t cap = (size_t)s->capacity;
    AK_CUDA(s, dalloc(&s->pos, cap)); AK_CUDA(s, dalloc(&s->posAlt, cap));
    AK_CUDA(s, dalloc(&s->vel, cap)); AK_CUDA(s, dalloc(&s->velAlt, cap));
    AK_CUDA(s, dalloc(&s->xs, cap));  AK_CUDA(s, dalloc(&s->xsAlt, cap));
    AK_CUDA(s, dalloc(&s->id, cap));  AK_CUDA(s, dalloc(&s->idAlt, cap));
    AK_CUDA(s, dalloc(&s->density, cap)); AK_CUDA(s, dalloc(&s->lambda, cap)); AK_CUDA(s, dalloc(&s->omegaLen, cap));
    AK_CUDA(s, dalloc(&s->omega, cap)); AK_CUDA(s, dalloc(&s->dpos, cap));
    AK_CUDA(s, dalloc(&s->color, cap)); AK_CUDA(s, dalloc(&s->size, cap));
    if (const char* e = std::getenv("AKUA_PDL")) s->opt.use_pdl = std::atoi(e) != 0;   // tuning experiments
    if (const char* e = std::getenv("AKUA_GATHER_LAYOUT")) s->opt.gather_layout = std::atoi(e);   // tuning experiments
    if (s->opt.gather_layout < AKUA_GATHER_AUTO || s->opt.gather_layout > AKUA_GATHER_PACKED_RECORDS) { s->err = "unknown gather_layout"; return AKUA_ERR_INVALID; }
    if (const char* e = std::getenv("AKUA_LIST_BUILD")) s->opt.list_build = std::atoi(e);   // tuning experiments
    if (s->opt.list_build < AKUA_LIST_BUILD_SCAN || s->opt.list_build > AKUA_LIST_BUILD_MASK8) { s->err = "unknown list_build"; return AKUA_ERR_INVALID; }
    if (const char* e = std::getenv("AKUA_WALL_MODEL")) s->opt.wall_model = std::atoi(e);   // experiments
    if (s->opt.wall_model != AKUA_WALL_REFERENCE && s->opt.wall_model != AKUA_WALL_VIRTUAL_FLUID) { s->err = "unknown wall_model"; return AKUA_ERR_INVALID; }
    {
        const int g = s->opt.gather_layout;
        if (g != AKUA_GATHER_PLAIN && g != AKUA_GATHER_RECORDS) {
            AK_CUDA(s, dalloc(&s->xl, cap)); AK_CUDA(s, dalloc(&s->xw, cap));
            AK_CUDA(s, cudaMemsetAsync(s->xl, 0, cap * sizeof(float4), s->stream));
            AK_CUDA(s, cudaMemsetAsync(s->xw, 0, cap * sizeof(float4), s->stream));
        }
        if (g == AKUA_GATHER_RECORDS || g == AKUA_GATHER_PACKED_RECORDS) {
            AK_CUDA(s, dalloc(&s->pv, cap));
            AK_CUDA(s, cudaMemsetAsync(s->pv, 0, cap * sizeof(PosVel), s->stream));
        }
    }
    AK_CUDA(s, dalloc(&s->dMassRange, 2));
    AK_CUDA(s, cudaMallocHost(reinterpret_cast<void**>(&s->hMassRange), 2 * sizeof(uint32_t)));
    AK_CUDA(s, dalloc(&s->keysUnsorted, cap));
    AK_CUDA(s, dalloc(&s->keyA, cap)); AK_CUDA(s, dalloc(&s->keyB, cap));
    AK_CUDA(s, dalloc(&s->valA, cap)); AK_CUDA(s, dalloc(&s->valB, cap));
    s->keysSorted = s->keyA; s->perm = s->valA;
    s->nbrStride = (uint32_t)((cap + 31) / 32 * 32);
    AK_CUDA(s, dalloc(&s->nbrList, (size_t)s->nbrStride * (size_t)((cfg->maxNeighbours + 3) / 4 * 4)));
    AK_CUDA(s, dalloc(&s->nbrCount, cap));
    s->sortWs.maxTiles = rsort::max_tiles_for_capacity(cap);
    AK_CUDA(s, dalloc(&s->sortWs.tileHist, rsort::tile_hist_words(s->sortWs.maxTiles)));
    AK_CUDA(s, dalloc(&s->sortWs.binTotal, rsort::kCtrlWords));
    if (const char* e = std::getenv("AKUA_SORT_MODE")) s->sortWs.mode = std::atoi(e);     // tuning experiments: 0 = three kernels per pass
    if (const char* e = std::getenv("AKUA_SORT_ITEMS")) s->sortWs.items = std::atoi(e);   // tuning experiments: keys per thread (4 / 8 / 16)
    if (const char* e = std::getenv("AKUA_CANONICAL_ORDER")) s->opt.canonical_order = std::atoi(e) != 0;   // tests / experiments
    if (s->opt.canonical_order) { AK_CUDA(s, dalloc(&s->canonKeys, cap)); AK_CUDA(s, dalloc(&s->canonVals, cap)); }
    AK_CUDA(s, dalloc(&s->partSum, 1024)); AK_CUDA(s, dalloc(&s->partMax, 1024));
    for (float4* p : {s->pos, s->posAlt, s->vel, s->velAlt, s->xs, s->xsAlt, s->omega, s->dpos, s->color})
        AK_CUDA(s, cudaMemsetAsync(p, 0, cap * sizeof(float4), s->stream));
    for (float* p : {s->density, s->lambda, s->omegaLen, s->size}) AK_CUDA(s, cudaMemsetAsync(p, 0, cap * sizeof(float), s->stream));
    for (uint32_t* p : {s->id, s->idAlt, s->keysUnsorted, s->keyA, s->keyB, s->valA, s->valB, s->nbrCount})
        AK_CUDA(s, cudaMemsetAsync(p, 0, cap * sizeof(uint32_t), s->stream));
    s->grid.cellSize = cfg->smoothRadius;             // NeighbourSearchCUDA.cu:163
    s->grid.lookupCellSize = cfg->spatialHashCellSize;  // NeighbourSearchCUDA.cu:177
    if (s->opt.key_mode == AKUA_KEY_REFERENCE_HASH) {
        // tableSize = maxNeighbours * numParticles as an int (PBFSolver.cpp:15); undefined in the reference beyond INT_MAX.
        int64_t ts = (int64_t)cfg->maxNeighbours * numParticles;
        if (numParticles == 0) ts = 1;  // the reference would compute `% 0`; nothing is ever hashed with n = 0
        if (ts <= 0 || ts > 0x7fffffffLL) { s->err = "REFERENCE_HASH: maxNeighbours*numParticles must be in [1, 2^31) (the reference's int tableSize); use LINEAR_CELL"; return AKUA_ERR_INVALID; }
        s->grid.tableSize = (uint32_t)ts;
        s->keyBits = bitsFor((uint64_t)ts - 1);
        s->ctr.num_cells = ts;
        AK_CUDA(s, dalloc(&s->bucketStart, (size_t)ts));
        launchPlain(s->stream, k_fill_u32, 148 * 8, kBlock, s->bucketStart, (uint64_t)ts, 0xffffffffu);
        AK_LAUNCH_CHECK(s, "k_fill_u32");
    } else if (s->opt.key_mode != AKUA_KEY_LINEAR_CELL) {
        s->err = "unknown key_mode"; return AKUA_ERR_INVALID;
    }
    for (int p = 0; p < PH_COUNT; p++) AK_CUDA(s, cudaEventCreate(&s->ev[p]));
    for (int i = 0; i < akua_pbf_solver::kMaxTimedIters; i++)
        for (int k = 0; k < 3; k++) AK_CUDA(s, cudaEventCreate(&s->evPass[i][k]));
    AK_CUDA(s, cudaStreamSynchronize(s->stream));
    return AKUA_OK;
}

void akua_pbf_destroy(akua_pbf_solver* s) {
    if (!s) return;
    cudaSetDevice(s->device);
    if (s->stream) cudaStreamSynchronize(s->stream);
    void* ptrs[] = {s->pos, s->posAlt, s->vel, s->velAlt, s->xs, s->xsAlt, s->id, s->idAlt, s->density, s->lambda,
                    s->omegaLen, s->omega, s->dpos, s->color, s->size, s->keysUnsorted, s->keyA, s->keyB, s->valA, s->valB,
                    s->bucketStart, s->cellRange, s->nbrList, s->nbrCount, s->sortWs.tileHist, s->sortWs.binTotal,
                    s->aosStage, s->partSum, s->partMax, s->xl, s->xw, s->pv, s->dMassRange, s->canonKeys, s->canonVals};
    if (s->hMassRange) cudaFreeHost(s->hMassRange);
    for (auto& g : s->graphs) if (g.used && g.exec) cudaGraphExecDestroy(g.exec);
    for (void* p : ptrs) if (p) cudaFree(p);
    {
        SlabState& sl = s->slab;
        void* sp[] = {sl.dims, sl.blockCnt, sl.sendL, sl.sendR, sl.recvL, sl.recvR, sl.slot, sl.slotAlt, sl.freeSlots};
        for (void* p : sp) if (p) cudaFree(p);
        for (SlabPeer* p : {&sl.peerL, &sl.peerR})
            if (p->open) for (void* q : {(void*)p->xsBuf[0], (void*)p->xsBuf[1], (void*)p->velBuf[0], (void*)p->velBuf[1],
                                         (void*)p->lambda, (void*)p->omegaLen, (void*)p->flags, (void*)p->dims,
                                         (void*)p->recvL, (void*)p->recvR, (void*)p->xl, (void*)p->xw}) if (q) cudaIpcCloseMemHandle(q);
        if (sl.flags) cudaFree(sl.flags);
        if (sl.hDims) cudaFreeHost((void*)sl.hDims);
        if (sl.dHist) cudaFree(sl.dHist);
        if (sl.hHist) cudaFreeHost(sl.hHist);
        if (sl.evRebalance) cudaEventDestroy(sl.evRebalance);
        if (sl.commStream) { cudaStreamSynchronize(sl.commStream); cudaStreamDestroy(sl.commStream); }
        for (int e = 0; e < SlabState::kEvents; e++) if (sl.evPool[e]) cudaEventDestroy(sl.evPool[e]);
        if (sl.comm && g_nccl.CommDestroy) g_nccl.CommDestroy((ncclComm_t)sl.comm);
    }
    for (int p = 0; p < PH_COUNT; p++) if (s->ev[p]) cudaEventDestroy(s->ev[p]);
    for (int i = 0; i < akua_pbf_solver::kMaxTimedIters; i++)
        for (int k = 0; k < 3; k++) if (s->evPass[i][k]) cudaEventDestroy(s->evPass[i][k]);
    if (s->copyStream) { cudaStreamSynchronize(s->copyStream); cudaStreamDestroy(s->copyStream); }
    for (auto& e : s->xferEv) if (e) cudaEventDestroy(e);
    if (s->stream) cudaStreamDestroy(s->stream);
    delete s;
}

int akua_pbf_step(akua_pbf_solver* s, float dt, const float boxMin[3], const float boxMax[3]) {
    if (!s) return AKUA_ERR_INVALID;
    return stepImpl(s, dt, s->cfg.solverIterations, boxMin, boxMax);
}
int akua_pbf_step_iters(akua_pbf_solver* s, float dt, int32_t iters, const float boxMin[3], const float boxMax[3]) {
    if (!s) return AKUA_ERR_INVALID;
    return stepImpl(s, dt, iters, boxMin, boxMax);
}
// Fixed-timestep driver: the accumulator loop of Application::run (src/Application/Application.cpp:37-70), headless.
int akua_pbf_advance(akua_pbf_solver* s, float frameTime, float deltaTime, int32_t maxStepsPerFrame, const float boxMin[3],
                     const float boxMax[3], int32_t* stepsDone) {
    if (!s || !(deltaTime > 0.0f) || frameTime < 0.0f) return AKUA_ERR_INVALID;
    s->accumulator += frameTime;
    int n = 0;
    while (s->accumulator >= deltaTime && n < maxStepsPerFrame) {   // Application.cpp:63-70
        int rc = stepImpl(s, deltaTime, s->cfg.solverIterations, boxMin, boxMax);
        if (rc) { if (stepsDone) *stepsDone = n; return rc; }
        s->accumulator -= deltaTime;
        n++;
    }
    if (stepsDone) *stepsDone = n;
    return AKUA_OK;
}
int akua_pbf_run_steps(akua_pbf_solver* s, int32_t steps, float deltaTime, const float boxMin[3], const float boxMax[3]) {
    if (!s || steps < 0) return AKUA_ERR_INVALID;
    for (int i = 0; i < steps; i++) {
        int rc = stepImpl(s, deltaTime, s->cfg.solverIterations, boxMin, boxMax);
        if (rc) return rc;
    }
    return AKUA_OK;
}
int akua_pbf_set_gravity(akua_pbf_solver* s, const float g[3]) {
    if (!s || !g) return AKUA_ERR_INVALID;
    s->cfg.gravity[0] = g[0]; s->cfg.gravity[1] = g[1]; s->cfg.gravity[2] = g[2];
    return AKUA_OK;
}
int akua_pbf_sync(akua_pbf_solver* s) {
    if (!s) return AKUA_ERR_INVALID;
    AK_CUDA(s, cudaSetDevice(s->device));
    if (s->slab.enabled) return slabRefresh(s);   // synchronises, refreshes the live count, reports device-side errors
    AK_CUDA(s, cudaStreamSynchronize(s->stream));
    return AKUA_OK;
}
const char* akua_pbf_last_error(const akua_pbf_solver* s) { return s ? s->err.c_str() : "null solver"; }
int64_t akua_pbf_num_particles(const akua_pbf_solver* s) {
    if (!s) return -1;
    if (s->slab.enabled) {   // the owned count changes on the device (migration): make the host's view exact
        akua_pbf_solver* m = const_cast<akua_pbf_solver*>(s);
        cudaSetDevice(m->device);
        slabRefresh(m);
    }
    return s->n;
}

int akua_pbf_upload_aos108(akua_pbf_solver* s, const void* src, int64_t n) {
    if (!s || !src || n < 0 || n > s->capacity) { if (s) s->err = "upload_aos108: n must be in [0, capacity]"; return AKUA_ERR_INVALID; }
    AK_CUDA(s, cudaSetDevice(s->device));
    int rc = beforeUpload(s, n);   // the live particle count follows the upload (slab ranks upload their own share)
    if (rc) return rc;
    if (n == 0) { s->massUniform = false; return slabAgreeMass(s); }   // still takes part in the slab-mode verdict
    if ((rc = ensureStage(s))) return rc;
    // Chunked: the copy engine moves chunk c + 1 from the host while k_unpack_aos scatters chunk c into the SoA arrays.
    // Particle::hash goes to the debug copy of the unsorted keys, NOT to keysSorted: in REFERENCE_HASH mode keysSorted must keep
    // describing the bucket table's current entries (they are un-written from it at the next step)
    {
        const int chunks = xferChunks(n);
        AK_CUDA(s, cudaEventRecord(s->xferEv[akua_pbf_solver::kXferChunks], s->stream));       // the staging buffer's last user
        AK_CUDA(s, cudaStreamWaitEvent(s->copyStream, s->xferEv[akua_pbf_solver::kXferChunks], 0));
        for (int c = 0; c < chunks; c++) {
            int64_t a, m;
            xferChunk(n, chunks, c, &a, &m);
            if (m <= 0) continue;
            char* stage = (char*)s->aosStage + (size_t)a * 108;
            AK_CUDA(s, cudaMemcpyAsync(stage, (const char*)src + (size_t)a * 108, (size_t)m * 108, cudaMemcpyHostToDevice, s->copyStream));
            AK_CUDA(s, cudaEventRecord(s->xferEv[c], s->copyStream));
            AK_CUDA(s, cudaStreamWaitEvent(s->stream, s->xferEv[c], 0));
            launchPlain(s->stream, k_unpack_aos, gridFor(m), kBlock, (const uint32_t*)stage, (uint32_t)m, s->pos + a, s->vel + a, s->xs + a,
                s->omega + a, s->omegaLen + a, s->dpos + a, s->density + a, s->lambda + a, s->keysUnsorted + a, s->color + a, s->size + a,
                s->id + a, (uint32_t)a);
            AK_LAUNCH_CHECK(s, "k_unpack_aos");
        }
        s->ctr.h2d_bytes += n * 108;
    }
    int rcm = massRangeAsync(s);
    if (rcm) return rcm;
    AK_CUDA(s, cudaStreamSynchronize(s->stream));  // `src` may be reused by the caller as soon as we return
    massRangeFinish(s);
    return slabAgreeMass(s);
}
int akua_pbf_download_aos108(akua_pbf_solver* s, void* dst, int64_t n) {
    if (!s || !dst) return AKUA_ERR_INVALID;
    AK_CUDA(s, cudaSetDevice(s->device));
    int rc = slabRefresh(s);
    if (rc) return rc;
    if (n != s->n) { s->err = "download_aos108: n must equal numParticles"; return AKUA_ERR_INVALID; }
    if (n == 0) return AKUA_OK;
    if ((rc = ensureStage(s))) return rc;
    // Chunked: k_pack_aos gathers chunk c + 1 into the staging buffer while the copy engine moves chunk c to the host.
    {
        const int chunks = xferChunks(n);
        const uint32_t* slot = s->slab.enabled ? s->slab.slot : nullptr;
        for (int c = 0; c < chunks; c++) {
            int64_t a, m;
            xferChunk(n, chunks, c, &a, &m);
            if (m <= 0) continue;
            char* stage = (char*)s->aosStage + (size_t)a * 108;
            launchPlain(s->stream, k_pack_aos, gridFor(m), kBlock, (uint32_t*)stage, (uint32_t)m, s->pos + a, s->vel + a, s->xs + a, s->omega + a,
                s->dpos + a, s->density + a, s->lambda + a, hashField(s) + a, s->color, s->size, s->id + a, slot ? slot + a : nullptr);
            AK_LAUNCH_CHECK(s, "k_pack_aos");
            AK_CUDA(s, cudaEventRecord(s->xferEv[c], s->stream));
            AK_CUDA(s, cudaStreamWaitEvent(s->copyStream, s->xferEv[c], 0));
            AK_CUDA(s, cudaMemcpyAsync((char*)dst + (size_t)a * 108, stage, (size_t)m * 108, cudaMemcpyDeviceToHost, s->copyStream));
        }
        AK_CUDA(s, cudaStreamSynchronize(s->copyStream));
        AK_CUDA(s, cudaStreamSynchronize(s->stream));
        s->ctr.d2h_bytes += n * 108;
    }
    return AKUA_OK;
}

int akua_pbf_export_aos108_device(akua_pbf_solver* s, void* device_dst, int64_t n) {
    if (!s || !device_dst) return AKUA_ERR_INVALID;
    AK_CUDA(s, cudaSetDevice(s->device));
    int rc = slabRefresh(s);
    if (rc) return rc;
    if (n != s->n) { s->err = "export_aos108_device: n must equal numParticles"; return AKUA_ERR_INVALID; }
    if (n == 0) return AKUA_OK;
    launchPlain(s->stream, k_pack_aos, gridFor(n), kBlock, (uint32_t*)device_dst, (uint32_t)n, s->pos, s->vel, s->xs, s->omega, s->dpos,
        s->density, s->lambda, hashField(s), s->color, s->size, s->id, s->slab.enabled ? s->slab.slot : nullptr);
    AK_LAUNCH_CHECK(s, "k_pack_aos");
    return AKUA_OK;
}

int akua_pbf_export_to_graphics_resource(akua_pbf_solver* s, void* graphicsResource) {
    if (!s || !graphicsResource) { if (s) s->err = "export_to_graphics_resource: null resource"; return AKUA_ERR_INVALID; }
    AK_CUDA(s, cudaSetDevice(s->device));
    cudaGraphicsResource_t res = reinterpret_cast<cudaGraphicsResource_t>(graphicsResource);
    AK_CUDA(s, cudaGraphicsMapResources(1, &res, s->stream));
    void* dst = nullptr;
    size_t bytes = 0;
    int rc = AKUA_OK;
    cudaError_t e = cudaGraphicsResourceGetMappedPointer(&dst, &bytes, res);
    if (e != cudaSuccess) { s->err = std::string("cudaGraphicsResourceGetMappedPointer: ") + cudaGetErrorString(e); rc = AKUA_ERR_CUDA; }
    else if ((rc = slabRefresh(s)) != AKUA_OK) {}
    else if (bytes < (size_t)s->n * 108) { s->err = "export_to_graphics_resource: the buffer is smaller than 108 * numParticles bytes"; rc = AKUA_ERR_INVALID; }
    else rc = akua_pbf_export_aos108_device(s, dst, s->n);
    e = cudaGraphicsUnmapResources(1, &res, s->stream);   // stream-ordered behind the pack kernel
    if (rc == AKUA_OK && e != cudaSuccess) { s->err = std::string("cudaGraphicsUnmapResources: ") + cudaGetErrorString(e); rc = AKUA_ERR_CUDA; }
    return rc;
}

int akua_pbf_upload_soa(akua_pbf_solver* s, const float* pos_xyz, const float* vel_xyz, const float* mass, int64_t n) {
    if (!s || !pos_xyz || n < 0 || n > s->capacity) { if (s) s->err = "upload_soa: n must be in [0, capacity]"; return AKUA_ERR_INVALID; }
    AK_CUDA(s, cudaSetDevice(s->device));
    {
        int rc = beforeUpload(s, n);
        if (rc) return rc;
    }
    if (n == 0) { s->massUniform = false; return slabAgreeMass(s); }
    // Host-side widening to float4, then two async copies. (Setup path; the per-step e2e path is AoS-108.)
    std::vector<float4> p4((size_t)n), v4((size_t)n);
    std::vector<uint32_t> ids((size_t)n);
    bool uniform = true;
    const float m0 = mass ? mass[0] : 1.0f;
    for (int64_t i = 0; i < n; i++) {
        float m = mass ? mass[i] : 1.0f;
        uniform = uniform && std::memcmp(&m, &m0, 4) == 0;
        p4[i] = make_float4(pos_xyz[3 * i], pos_xyz[3 * i + 1], pos_xyz[3 * i + 2], m);
        v4[i] = vel_xyz ? make_float4(vel_xyz[3 * i], vel_xyz[3 * i + 1], vel_xyz[3 * i + 2], 0.f) : make_float4(0, 0, 0, 0);
        ids[i] = (uint32_t)i;
    }
    AK_CUDA(s, cudaMemcpyAsync(s->pos, p4.data(), (size_t)n * 16, cudaMemcpyHostToDevice, s->stream));
    AK_CUDA(s, cudaMemcpyAsync(s->xs, p4.data(), (size_t)n * 16, cudaMemcpyHostToDevice, s->stream));
    AK_CUDA(s, cudaMemcpyAsync(s->vel, v4.data(), (size_t)n * 16, cudaMemcpyHostToDevice, s->stream));
    AK_CUDA(s, cudaMemcpyAsync(s->id, ids.data(), (size_t)n * 4, cudaMemcpyHostToDevice, s->stream));
    // the fields a lean upload does not carry get the scene defaults of Application.cpp:186-187 (blue, size 50) / zero
    launchPlain(s->stream, k_fill_payload, std::min<uint32_t>(gridFor(n), 148 * 8), kBlock, s->color, s->size, (uint32_t)n,
        make_float4(0.f, 0.f, 1.f, 1.f), 50.0f);
    AK_LAUNCH_CHECK(s, "k_fill_payload");
    for (float4* q : {s->omega, s->dpos}) AK_CUDA(s, cudaMemsetAsync(q, 0, (size_t)n * sizeof(float4), s->stream));
    for (float* q : {s->density, s->lambda, s->omegaLen}) AK_CUDA(s, cudaMemsetAsync(q, 0, (size_t)n * sizeof(float), s->stream));
    AK_CUDA(s, cudaMemsetAsync(s->keysUnsorted, 0, (size_t)n * sizeof(uint32_t), s->stream));
    AK_CUDA(s, cudaStreamSynchronize(s->stream));
    s->ctr.h2d_bytes += n * 52;
    s->massUniform = uniform; s->uniformMass = m0;
    return slabAgreeMass(s);
}
int akua_pbf_download_soa(akua_pbf_solver* s, float* pos4, float* vel4, uint32_t* id, int64_t n) {
    if (!s) return AKUA_ERR_INVALID;
    AK_CUDA(s, cudaSetDevice(s->device));
    {
        int rc = slabRefresh(s);
        if (rc) return rc;
    }
    if (n != s->n) { s->err = "download_soa: n must equal numParticles"; return AKUA_ERR_INVALID; }
    if (pos4) { AK_CUDA(s, cudaMemcpyAsync(pos4, s->pos, (size_t)n * 16, cudaMemcpyDeviceToHost, s->stream)); s->ctr.d2h_bytes += n * 16; }
    if (vel4) { AK_CUDA(s, cudaMemcpyAsync(vel4, s->vel, (size_t)n * 16, cudaMemcpyDeviceToHost, s->stream)); s->ctr.d2h_bytes += n * 16; }
    if (id)   { AK_CUDA(s, cudaMemcpyAsync(id, s->id, (size_t)n * 4, cudaMemcpyDeviceToHost, s->stream)); s->ctr.d2h_bytes += n * 4; }
    AK_CUDA(s, cudaStreamSynchronize(s->stream));
    return AKUA_OK;
}
const float* akua_pbf_positions_device(akua_pbf_solver* s) { return s ? reinterpret_cast<const float*>(s->pos) : nullptr; }
const float* akua_pbf_velocities_device(akua_pbf_solver* s) { return s ? reinterpret_cast<const float*>(s->vel) : nullptr; }

void* akua_pbf_host_alloc(int64_t bytes) {
    void* p = nullptr;
    if (bytes <= 0 || cudaMallocHost(&p, (size_t)bytes) != cudaSuccess) return nullptr;
    return p;
}
void akua_pbf_host_free(void* p) { if (p) cudaFreeHost(p); }

// ---- checkpoint / resume (the reference keeps its state only in the GL VBO) ----
namespace {
struct CkptHeader {
    char magic[8];          // "AKUAPBF2"
    uint32_t headerBytes;   // sizeof(CkptHeader): a file written by a different layout is rejected, not misread
    uint32_t version;
    int64_t n;
    akua_pbf_config cfg;
    akua_corr_params corr;
    int64_t steps;
    float accumulator;
    int32_t key_mode;
};
constexpr uint32_t kCkptVersion = 2;
bool sameBytes(const void* a, const void* b, size_t n) { return std::memcmp(a, b, n) == 0; }
}
// File: header, then per particle in the solver's CURRENT order: position (x,y,z,mass), velocity (vx,vy,vz,density), id,
// color, size. The payload is stored per particle, so the file does not depend on where the solver keeps it.
int akua_pbf_checkpoint_save(akua_pbf_solver* s, const char* path) {
    if (!s || !path) return AKUA_ERR_INVALID;
    AK_CUDA(s, cudaSetDevice(s->device));
    int rc = slabRefresh(s);
    if (rc) return rc;
    AK_CUDA(s, cudaStreamSynchronize(s->stream));
    const size_t n = (size_t)s->n, cap = (size_t)s->capacity;
    std::vector<float4> pos(n), vel(n), colorAll(cap), color(n);
    std::vector<float> sizeAll(cap), size(n);
    std::vector<uint32_t> id(n), pidx(n);
    if (n) {
        AK_CUDA(s, cudaMemcpy(pos.data(), s->pos, n * 16, cudaMemcpyDeviceToHost));
        AK_CUDA(s, cudaMemcpy(vel.data(), s->vel, n * 16, cudaMemcpyDeviceToHost));
        AK_CUDA(s, cudaMemcpy(id.data(), s->id, n * 4, cudaMemcpyDeviceToHost));
        AK_CUDA(s, cudaMemcpy(pidx.data(), s->slab.enabled ? s->slab.slot : s->id, n * 4, cudaMemcpyDeviceToHost));
        AK_CUDA(s, cudaMemcpy(colorAll.data(), s->color, cap * 16, cudaMemcpyDeviceToHost));
        AK_CUDA(s, cudaMemcpy(sizeAll.data(), s->size, cap * 4, cudaMemcpyDeviceToHost));
        for (size_t i = 0; i < n; i++) { const uint32_t q = pidx[i] < cap ? pidx[i] : 0u; color[i] = colorAll[q]; size[i] = sizeAll[q]; }
    }
    FILE* f = std::fopen(path, "wb");
    if (!f) { s->err = std::string("checkpoint_save: cannot open ") + path; return AKUA_ERR_INVALID; }
    CkptHeader h{};
    std::memcpy(h.magic, "AKUAPBF2", 8);
    h.headerBytes = (uint32_t)sizeof(CkptHeader); h.version = kCkptVersion;
    h.n = (int64_t)n; h.cfg = s->cfg; h.corr = s->corr; h.steps = s->ctr.steps; h.accumulator = s->accumulator;
    h.key_mode = s->opt.key_mode;
    bool ok = std::fwrite(&h, sizeof(h), 1, f) == 1;
    ok = ok && (n == 0 || (std::fwrite(pos.data(), 16, n, f) == n && std::fwrite(vel.data(), 16, n, f) == n &&
                           std::fwrite(id.data(), 4, n, f) == n && std::fwrite(color.data(), 16, n, f) == n &&
                           std::fwrite(size.data(), 4, n, f) == n));
    ok = (std::fclose(f) == 0) && ok;
    if (!ok) { s->err = "checkpoint_save: short write"; return AKUA_ERR_INVALID; }
    return AKUA_OK;
}
// A checkpoint continues bit-identically only under the parameters it was written with: the solver's config, correction
// parameters and key mode must equal the file's (gravity excepted: it is runtime state and is restored from the file).
int akua_pbf_checkpoint_load(akua_pbf_solver* s, const char* path) {
    if (!s || !path) return AKUA_ERR_INVALID;
    AK_CUDA(s, cudaSetDevice(s->device));
    FILE* f = std::fopen(path, "rb");
    if (!f) { s->err = std::string("checkpoint_load: cannot open ") + path; return AKUA_ERR_INVALID; }
    CkptHeader h{};
    if (std::fread(&h, sizeof(h), 1, f) != 1 || std::memcmp(h.magic, "AKUAPBF2", 8) != 0 || h.headerBytes != sizeof(CkptHeader) ||
        h.version != kCkptVersion) {
        std::fclose(f); s->err = "checkpoint_load: not an AKUAPBF2 file of this library version"; return AKUA_ERR_INVALID;
    }
    if (h.n < 0 || h.n > s->capacity) { std::fclose(f); s->err = "checkpoint_load: particle count exceeds solver capacity"; return AKUA_ERR_INVALID; }
    {
        akua_pbf_config a = h.cfg, b = s->cfg;
        for (int k = 0; k < 3; k++) a.gravity[k] = b.gravity[k] = 0.0f;
        const char* why = nullptr;
        if (!sameBytes(&a, &b, sizeof(a))) why = "checkpoint_load: the file was written with a different PBFConfig";
        else if (!sameBytes(&h.corr, &s->corr, sizeof(h.corr))) why = "checkpoint_load: the file was written with different LambdaCorrParams";
        else if (h.key_mode != s->opt.key_mode) why = "checkpoint_load: the file was written in a different key mode";
        else if (s->opt.key_mode == AKUA_KEY_REFERENCE_HASH && (int64_t)s->cfg.maxNeighbours * h.n != (int64_t)s->grid.tableSize)
            why = "checkpoint_load: REFERENCE_HASH needs the particle count the solver was created with (tableSize = 128 * n)";
        if (why) { std::fclose(f); s->err = why; return AKUA_ERR_INVALID; }
    }
    const size_t n = (size_t)h.n;
    std::vector<float4> pos(n), vel(n), color(n);
    std::vector<float> size(n);
    std::vector<uint32_t> id(n);
    bool ok = n == 0 || (std::fread(pos.data(), 16, n, f) == n && std::fread(vel.data(), 16, n, f) == n &&
                         std::fread(id.data(), 4, n, f) == n && std::fread(color.data(), 16, n, f) == n &&
                         std::fread(size.data(), 4, n, f) == n);
    std::fclose(f);
    if (!ok) { s->err = "checkpoint_load: truncated file"; return AKUA_ERR_INVALID; }
    AK_CUDA(s, cudaStreamSynchronize(s->stream));
    int rc = beforeUpload(s, h.n);
    if (rc) return rc;
    if (n) {
        // payload placement: slot i in x-slab mode (identity slots after the reset), index id[i] on one GPU
        std::vector<float4> colorAt(n);
        std::vector<float> sizeAt(n);
        const bool byId = !s->slab.enabled;
        for (size_t i = 0; i < n; i++) {
            if (byId && id[i] >= (uint32_t)s->capacity) { s->err = "checkpoint_load: particle id outside the solver's capacity"; return AKUA_ERR_INVALID; }
        }
        if (byId) {
            std::vector<float4> cfull((size_t)s->capacity, make_float4(0.f, 0.f, 1.f, 1.f));
            std::vector<float> sfull((size_t)s->capacity, 50.0f);
            for (size_t i = 0; i < n; i++) { cfull[id[i]] = color[i]; sfull[id[i]] = size[i]; }
            AK_CUDA(s, cudaMemcpy(s->color, cfull.data(), cfull.size() * 16, cudaMemcpyHostToDevice));
            AK_CUDA(s, cudaMemcpy(s->size, sfull.data(), sfull.size() * 4, cudaMemcpyHostToDevice));
        } else {
            AK_CUDA(s, cudaMemcpy(s->color, color.data(), n * 16, cudaMemcpyHostToDevice));
            AK_CUDA(s, cudaMemcpy(s->size, size.data(), n * 4, cudaMemcpyHostToDevice));
        }
        AK_CUDA(s, cudaMemcpy(s->pos, pos.data(), n * 16, cudaMemcpyHostToDevice));
        AK_CUDA(s, cudaMemcpy(s->xs, pos.data(), n * 16, cudaMemcpyHostToDevice));
        AK_CUDA(s, cudaMemcpy(s->vel, vel.data(), n * 16, cudaMemcpyHostToDevice));
        AK_CUDA(s, cudaMemcpy(s->id, id.data(), n * 4, cudaMemcpyHostToDevice));
        AK_CUDA(s, cudaMemsetAsync(s->keysUnsorted, 0, n * sizeof(uint32_t), s->stream));
        AK_CUDA(s, cudaStreamSynchronize(s->stream));
    }
    s->massUniform = n > 0;
    s->uniformMass = n ? pos[0].w : 0.0f;
    for (size_t i = 1; i < n && s->massUniform; i++) s->massUniform = std::memcmp(&pos[i].w, &pos[0].w, 4) == 0;
    s->cfg.gravity[0] = h.cfg.gravity[0]; s->cfg.gravity[1] = h.cfg.gravity[1]; s->cfg.gravity[2] = h.cfg.gravity[2];
    s->accumulator = h.accumulator;
    s->ctr.steps = h.steps;
    return slabAgreeMass(s);
}

// ---- multi-GPU (x-slab) ----
int akua_pbf_upload_ids(akua_pbf_solver* s, const uint32_t* ids, int64_t n) {
    if (!s || !ids) return AKUA_ERR_INVALID;
    AK_CUDA(s, cudaSetDevice(s->device));
    {
        int rc = slabRefresh(s);
        if (rc) return rc;
    }
    if (n != s->n) { s->err = "upload_ids: n must equal the live particle count"; return AKUA_ERR_INVALID; }
    if (n == 0) return AKUA_OK;
    if (!s->slab.enabled) {
        // on one GPU the id is also the index of the particle's render payload (k_pack_aos): it must stay inside the arrays
        for (int64_t i = 0; i < n; i++)
            if ((int64_t)ids[i] >= s->capacity) { s->err = "upload_ids: outside x-slab mode every id must be < capacity"; return AKUA_ERR_INVALID; }
    }
    AK_CUDA(s, cudaMemcpyAsync(s->id, ids, (size_t)n * 4, cudaMemcpyHostToDevice, s->stream));
    AK_CUDA(s, cudaStreamSynchronize(s->stream));
    return AKUA_OK;
}
int akua_pbf_comm_unique_id(void* out, int64_t out_bytes) {
    if (!out || out_bytes < (int64_t)sizeof(ncclUniqueId)) return AKUA_ERR_INVALID;
    if (loadNccl()) return AKUA_ERR_COMM;
    ncclUniqueId id;
    if (g_nccl.GetUniqueId(&id) != ncclSuccess) return AKUA_ERR_COMM;
    std::memcpy(out, &id, sizeof(id));
    return AKUA_OK;
}
int akua_pbf_comm_init(akua_pbf_solver* s, int32_t rank, int32_t nranks, const void* unique_id) {
    if (!s || !unique_id || nranks < 1 || rank < 0 || rank >= nranks) return AKUA_ERR_INVALID;
    if (s->opt.key_mode != AKUA_KEY_LINEAR_CELL) { s->err = "comm_init: slab mode needs LINEAR_CELL keys"; return AKUA_ERR_INVALID; }
    if (wallOn(s)) { s->err = "comm_init: the opt-in wall model is single-GPU only"; return AKUA_ERR_INVALID; }
    if (const char* e = loadNccl()) { s->err = e; return AKUA_ERR_COMM; }
    AK_CUDA(s, cudaSetDevice(s->device));
    SlabState& sl = s->slab;
    ncclUniqueId id;
    std::memcpy(&id, unique_id, sizeof(id));
    ncclComm_t comm = nullptr;
    AK_NCCL(s, g_nccl.CommInitRank(&comm, nranks, id, rank));
    sl.comm = comm; sl.rank = rank; sl.nranks = nranks;
    sl.migCap = (uint32_t)std::max<int64_t>(4096, s->capacity / 8);
    sl.migBlocksCap = (uint32_t)((s->capacity + slab::kMigTile - 1) / slab::kMigTile) + 1;
    {
        int lo = 0, hi = 0;
        AK_CUDA(s, cudaDeviceGetStreamPriorityRange(&lo, &hi));
        AK_CUDA(s, cudaStreamCreateWithPriority(&sl.commStream, cudaStreamNonBlocking, hi));
        for (int e = 0; e < SlabState::kEvents; e++) AK_CUDA(s, cudaEventCreateWithFlags(&sl.evPool[e], cudaEventDisableTiming));
    }
    AK_CUDA(s, dalloc(&sl.dims, D_WORDS));
    AK_CUDA(s, cudaMemsetAsync(sl.dims, 0, D_WORDS * sizeof(uint32_t), s->stream));
    AK_CUDA(s, cudaMallocHost((void**)&sl.hDims, D_WORDS * sizeof(uint32_t)));
    std::memset((void*)sl.hDims, 0, D_WORDS * sizeof(uint32_t));
    AK_CUDA(s, dalloc(&sl.slot, (size_t)s->capacity)); AK_CUDA(s, dalloc(&sl.slotAlt, (size_t)s->capacity));
    AK_CUDA(s, dalloc(&sl.freeSlots, (size_t)s->capacity));
    AK_CUDA(s, cudaMemsetAsync(sl.slotAlt, 0, (size_t)s->capacity * sizeof(uint32_t), s->stream));
    AK_CUDA(s, dalloc(&sl.blockCnt, (size_t)2 * sl.migBlocksCap));
    AK_CUDA(s, dalloc(&sl.sendL, sl.migCap)); AK_CUDA(s, dalloc(&sl.sendR, sl.migCap));
    AK_CUDA(s, dalloc(&sl.recvL, sl.migCap)); AK_CUDA(s, dalloc(&sl.recvR, sl.migCap));
    // fixed ghost regions at the top of every per-particle array: an eighth of the capacity each for small solvers (a plane of
    // a test scene can be a tenth of the slab), a sixteenth each from a million particles on (a plane of an 8 M-particle slab is
    // 1 - 2 % of it; the space goes to the owned region instead)
    sl.ghostCap = (uint32_t)(s->capacity < (1 << 20) ? s->capacity / 8 : s->capacity / 16);
    sl.ghostBaseL = (uint32_t)s->capacity - 2 * sl.ghostCap;
    sl.ghostBaseR = (uint32_t)s->capacity - sl.ghostCap;
    if (s->n > (int64_t)sl.ghostBaseL) { s->err = "comm_init: capacity_factor too small for the ghost regions (use >= 1.4)"; return AKUA_ERR_INVALID; }
    sl.xsBuf[0] = s->xs; sl.xsBuf[1] = s->xsAlt; sl.velBuf[0] = s->vel; sl.velBuf[1] = s->velAlt;
    // flag words: [0] / [1] ghost-exchange epoch published by the left / right rank, [3] CTA completion counter of the fused
    // pushes, [4] / [5] epoch of the left / right rank's per-step count + migration message
    AK_CUDA(s, dalloc(&sl.flags, 8));
    AK_CUDA(s, cudaMemsetAsync(sl.flags, 0, 8 * sizeof(uint32_t), s->stream));
    AK_CUDA(s, cudaStreamSynchronize(s->stream));
    // ---- CUDA IPC: open the neighbours' arrays so that ghost planes can be copied straight into them ----
    const char* env = std::getenv("AKUA_SLAB_P2P");
    const bool wantP2p = !(env && env[0] == '0') && nranks > 1;
    constexpr int kIpc = 12;   // the last two (packed gather arrays) may be absent
    struct PeerMsg { cudaIpcMemHandle_t h[kIpc]; uint32_t ghostBaseL, ghostBaseR, ok, migCap, hasPack, ghostCap, pad[2]; };
    PeerMsg mine{};
    mine.ghostBaseL = sl.ghostBaseL; mine.ghostBaseR = sl.ghostBaseR; mine.ok = wantP2p ? 1u : 0u; mine.migCap = sl.migCap;
    mine.ghostCap = sl.ghostCap;
    if (wantP2p) {
        void* arrs[kIpc] = {sl.xsBuf[0], sl.xsBuf[1], sl.velBuf[0], sl.velBuf[1], s->lambda, s->omegaLen, sl.flags,
                            sl.dims, sl.recvL, sl.recvR, s->xl, s->xw};
        mine.hasPack = (s->xl && s->xw) ? 1u : 0u;
        for (int k = 0; k < (mine.hasPack ? kIpc : kIpc - 2); k++)
            if (cudaIpcGetMemHandle(&mine.h[k], arrs[k]) != cudaSuccess) { mine.ok = 0; cudaGetLastError(); break; }
    }
    if (nranks > 1) {
        PeerMsg* dmsg = nullptr;   // [0] mine, [1] from left, [2] from right
        AK_CUDA(s, dalloc(&dmsg, 3));
        AK_CUDA(s, cudaMemcpy(dmsg, &mine, sizeof(mine), cudaMemcpyHostToDevice));
        const bool hasL = rank > 0, hasR = rank + 1 < nranks;
        AK_NCCL(s, g_nccl.GroupStart());
        if (hasL) AK_NCCL(s, g_nccl.Send(dmsg, sizeof(PeerMsg), ncclUint8, rank - 1, comm, sl.commStream));
        if (hasR) AK_NCCL(s, g_nccl.Send(dmsg, sizeof(PeerMsg), ncclUint8, rank + 1, comm, sl.commStream));
        if (hasL) AK_NCCL(s, g_nccl.Recv(dmsg + 1, sizeof(PeerMsg), ncclUint8, rank - 1, comm, sl.commStream));
        if (hasR) AK_NCCL(s, g_nccl.Recv(dmsg + 2, sizeof(PeerMsg), ncclUint8, rank + 1, comm, sl.commStream));
        AK_NCCL(s, g_nccl.GroupEnd());
        AK_CUDA(s, cudaStreamSynchronize(sl.commStream));
        PeerMsg got[2]{};
        AK_CUDA(s, cudaMemcpy(got, dmsg + 1, 2 * sizeof(PeerMsg), cudaMemcpyDeviceToHost));
        cudaFree(dmsg);
        bool ok = mine.ok != 0 && (!hasL || got[0].ok) && (!hasR || got[1].ok);
        auto openPeer = [&](const PeerMsg& m, SlabPeer& p) {
            void* ptr[kIpc] = {};
            for (int k = 0; k < (m.hasPack ? kIpc : kIpc - 2); k++)
                if (cudaIpcOpenMemHandle(&ptr[k], m.h[k], cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) { cudaGetLastError(); return false; }
            p.xsBuf[0] = (float4*)ptr[0]; p.xsBuf[1] = (float4*)ptr[1]; p.velBuf[0] = (float4*)ptr[2]; p.velBuf[1] = (float4*)ptr[3];
            p.lambda = (float*)ptr[4]; p.omegaLen = (float*)ptr[5]; p.flags = (uint32_t*)ptr[6];
            p.dims = (uint32_t*)ptr[7]; p.recvL = (slab::MigRecord*)ptr[8]; p.recvR = (slab::MigRecord*)ptr[9];
            p.migCap = m.migCap; p.xl = (float4*)ptr[10]; p.xw = (float4*)ptr[11];
            p.ghostBaseL = m.ghostBaseL; p.ghostBaseR = m.ghostBaseR; p.ghostCap = m.ghostCap; p.open = true;
            return true;
        };
        if (ok && hasL) ok = openPeer(got[0], sl.peerL);
        if (ok && hasR) ok = openPeer(got[1], sl.peerR);
        // all ranks must agree on the transport: a tiny all-reduce (min) over the local verdicts
        uint32_t* dflag = nullptr;
        AK_CUDA(s, dalloc(&dflag, 1));
        uint32_t v = ok ? 1u : 0u;
        AK_CUDA(s, cudaMemcpy(dflag, &v, 4, cudaMemcpyHostToDevice));
        AK_NCCL(s, g_nccl.AllReduce(dflag, dflag, 1, ncclUint32, ncclMin, comm, sl.commStream));
        AK_CUDA(s, cudaStreamSynchronize(sl.commStream));
        AK_CUDA(s, cudaMemcpy(&v, dflag, 4, cudaMemcpyDeviceToHost));
        cudaFree(dflag);
        sl.p2p = v != 0;
    }
    if (const char* eg = std::getenv("AKUA_SLAB_GRAPH")) { if (eg[0] == '0') sl.graphBroken = true; }   // tuning experiments
    if (const char* ek = std::getenv("AKUA_SLAB_KEEP_BELOW")) { const double v = std::atof(ek); if (v >= 1.0) sl.keepBelow = v; }
    AK_CUDA(s, cudaStreamSynchronize(s->stream));
    return AKUA_OK;
}
int akua_pbf_set_slab(akua_pbf_solver* s, int32_t xCellLo, int32_t xCellHi) {
    if (!s || !s->slab.comm) { if (s) s->err = "set_slab: call akua_pbf_comm_init first"; return AKUA_ERR_INVALID; }
    if (xCellHi <= xCellLo) { s->err = "set_slab: empty interval"; return AKUA_ERR_INVALID; }
    s->slab.xLoAbs = xCellLo; s->slab.xHiAbs = xCellHi;
    const bool first = !s->slab.enabled;
    s->slab.enabled = true;
    if (!first) return AKUA_OK;
    int rc = slabResetCounts(s);                 // the particles uploaded so far are this rank's owned set
    if (rc) return rc;
    return slabAgreeMass(s);                     // collective, like the call itself
}
int akua_pbf_rebalance(akua_pbf_solver* s) {
    if (!s) return AKUA_ERR_INVALID;
    AK_CUDA(s, cudaSetDevice(s->device));
    int rc = slabRebalanceApply(s);          // a measurement left in flight by akua_pbf_rebalance_async comes first
    if (rc) return rc;
    if ((rc = slabRebalanceMeasure(s))) return rc;
    return slabRebalanceApply(s);
}
int akua_pbf_rebalance_async(akua_pbf_solver* s) {
    if (!s) return AKUA_ERR_INVALID;
    AK_CUDA(s, cudaSetDevice(s->device));
    int rc = slabRebalanceApply(s);
    if (rc) return rc;
    return slabRebalanceMeasure(s);
}
int akua_pbf_slab_stats(const akua_pbf_solver* s, int64_t out[8]) {
    if (!s || !out) return AKUA_ERR_INVALID;
    akua_pbf_solver* m = const_cast<akua_pbf_solver*>(s);
    cudaSetDevice(m->device);
    const int rc = slabRefresh(m);
    const SlabState& sl = s->slab;
    unsigned long long st[3] = {0, 0, 0};
    if (sl.hDims) std::memcpy(st, (const void*)(sl.hDims + D_STAT_MIG_IN), sizeof(st));
    out[0] = s->n; out[1] = sl.nGhostL; out[2] = sl.nGhostR; out[3] = sl.nPlaneL; out[4] = sl.nPlaneR;
    out[5] = sl.exchanges; out[6] = sl.p2p ? -(int64_t)st[2] : (int64_t)st[2]; out[7] = (int64_t)st[0];  // negative bytes: p2p transport
    return rc;
}
int akua_pbf_slab_wait_stats(const akua_pbf_solver* s, int64_t out[4]) {
    if (!s || !out) return AKUA_ERR_INVALID;
    akua_pbf_solver* m = const_cast<akua_pbf_solver*>(s);
    cudaSetDevice(m->device);
    const int rc = slabRefresh(m);
    const SlabState& sl = s->slab;
    unsigned long long w[2] = {0, 0};
    if (sl.hDims) std::memcpy(w, (const void*)(sl.hDims + D_STAT_PLAN_WAIT_NS), sizeof(w));
    out[0] = (int64_t)w[0]; out[1] = (int64_t)w[1]; out[2] = sl.hDims ? (int64_t)sl.hDims[D_STEPS] : 0; out[3] = sl.rebalances;
    return rc;
}
// Balanced x-slab boundaries from a histogram of particles per absolute x cell column (pure host code, no CUDA):
// bounds[r] .. bounds[r+1] is rank r's interval of columns (indices into hist); bounds[0] = 0, bounds[nranks] = ncols.
int akua_slab_partition(const int64_t* hist, int32_t ncols, int32_t nranks, int32_t* bounds) {
    if (!hist || !bounds || ncols < nranks || nranks < 1) return AKUA_ERR_INVALID;
    int64_t total = 0;
    for (int c = 0; c < ncols; c++) total += hist[c];
    bounds[0] = 0;
    int64_t cum = 0;
    int c = 0;
    for (int r = 1; r < nranks; r++) {
        const int64_t target = (total * r + nranks / 2) / nranks;
        while (c < ncols && cum + hist[c] <= target) { cum += hist[c]; c++; }
        // keep every slab at least two columns wide (one when there are too few columns) and leave room for the rest
        const int minW = ncols >= 2 * nranks ? 2 : 1;
        int lo = bounds[r - 1] + minW, hi = ncols - minW * (nranks - r);
        int b = std::min(std::max(c, lo), hi);
        while (c < b) { cum += hist[c]; c++; }
        bounds[r] = b;
    }
    bounds[nranks] = ncols;
    return AKUA_OK;
}

// New slab boundaries for a re-balancing step (pure host code, no CUDA; used by akua_pbf_rebalance on every rank with the
// all-reduced histogram, so all ranks compute the same result). Starts from the balanced partition of `hist` and clamps
// every boundary r (between ranks r-1 and r) so that the ORDINARY per-step migration of the next step can carry the
// transfer: it stays strictly inside the two old slabs it separates (particles only ever move to an adjacent rank, and
// arrivals never reach a slab's far boundary plane), every slab stays at least two planes wide, and at most maxMove
// particles cross it.
int akua_slab_rebalance_bounds_weighted(const int64_t* work, const int64_t* count, int32_t ncols, int32_t nranks,
                                        const int32_t* oldBounds, int64_t maxMove, double keepBelow, int64_t maxCount, int32_t* bounds) {
    if (!work || !count || !oldBounds || !bounds || nranks < 1) return AKUA_ERR_INVALID;
    const int R = nranks;
    const int32_t* old = oldBounds;
    std::vector<int64_t> pre((size_t)ncols + 1, 0);
    for (int x = 0; x < ncols; x++) pre[x + 1] = pre[x] + count[x];
    auto cnt = [&](int a, int b) { a = std::min(std::max(a, 0), ncols); b = std::min(std::max(b, 0), ncols); return b > a ? pre[b] - pre[a] : (int64_t)0; };
    bool oldFits = true;
    if (maxCount > 0) for (int r = 0; r < R; r++) oldFits = oldFits && cnt(old[r], old[r + 1]) <= maxCount;
    if (keepBelow > 1.0 && oldFits) {
        // hysteresis: a partition whose heaviest slab is within keepBelow of the mean is left alone (moving a boundary costs a
        // migration burst and a re-capture of the step's CUDA graph)
        int64_t total = 0, heaviest = 0;
        for (int r = 0; r < R; r++) {
            int64_t w = 0;
            for (int x = std::max(old[r], 0); x < std::min(old[r + 1], ncols); x++) w += work[x];
            total += w; heaviest = std::max(heaviest, w);
        }
        if (total > 0 && (double)heaviest * R <= keepBelow * (double)total) {
            for (int r = 0; r <= R; r++) bounds[r] = old[r];
            return AKUA_OK;
        }
    }
    if (akua_slab_partition(work, ncols, nranks, bounds) != AKUA_OK) return AKUA_ERR_INVALID;
    if (maxCount > 0) {
        // no slab may hold more particles than its arrays have room for, whatever the work says: pull the upper boundary of an
        // over-full slab down (left to right), then push the lower boundary of one that is still over-full up (right to left)
        for (int r = 0; r + 1 < R; r++)
            while (cnt(bounds[r], bounds[r + 1]) > maxCount && bounds[r + 1] > bounds[r] + 2) bounds[r + 1]--;
        for (int r = R - 1; r >= 1; r--)
            while (cnt(bounds[r], bounds[r + 1]) > maxCount && bounds[r] < bounds[r + 1] - 2) bounds[r]++;
    }
    for (int r = 1; r < R; r++) {
        // stay inside the two old slabs and keep every slab at least two planes wide
        int b = std::min(std::max(bounds[r], std::max(old[r - 1] + 1, bounds[r - 1] + 2)), old[r + 1] - 2);
        int64_t moved = 0;
        if (b > old[r]) { int x = old[r]; while (x < b && moved + count[x] <= maxMove) { moved += count[x]; x++; } b = x; }
        else if (b < old[r]) { int x = old[r]; while (x > b && moved + count[x - 1] <= maxMove) { moved += count[x - 1]; x--; } b = x; }
        if (b < bounds[r - 1] + 2) b = std::min(bounds[r - 1] + 2, old[r + 1] - 2);
        bounds[r] = b;
    }
    return AKUA_OK;
}
int akua_slab_rebalance_bounds(const int64_t* hist, int32_t ncols, int32_t nranks, const int32_t* oldBounds, int64_t maxMove,
                               int32_t* bounds) {
    return akua_slab_rebalance_bounds_weighted(hist, hist, ncols, nranks, oldBounds, maxMove, 0.0, 0, bounds);
}

// ---- phase-level operators ----
static int phaseApiGuard(akua_pbf_solver* s) {
    if (!s) return AKUA_ERR_INVALID;
    if (s->slab.enabled) { s->err = "phase-level operators are single-GPU only; use akua_pbf_step in slab mode"; return AKUA_ERR_INVALID; }
    return AKUA_OK;
}
int akua_pbf_phase_predict(akua_pbf_solver* s, float dt) {
    if (int g = phaseApiGuard(s)) return g;
    AK_CUDA(s, cudaSetDevice(s->device));
    return phasePredictKey(s, dt, true, false);
}
int akua_pbf_phase_neighbours(akua_pbf_solver* s, const float boxMin[3], const float boxMax[3]) {
    if (int g = phaseApiGuard(s)) return g;
    if ((boxMin == nullptr) != (boxMax == nullptr)) { s->err = "phase_neighbours: pass both box corners or neither"; return AKUA_ERR_INVALID; }
    AK_CUDA(s, cudaSetDevice(s->device));
    const float* bmin = boxMin ? boxMin : s->lastBoxMin;
    const float* bmax = boxMax ? boxMax : s->lastBoxMax;
    if (s->opt.key_mode == AKUA_KEY_LINEAR_CELL && !boxMin && !s->haveBox) { s->err = "phase_neighbours: LINEAR_CELL needs a box"; return AKUA_ERR_INVALID; }
    if (boxMin && boxMax) rememberBox(s, boxMin, boxMax);
    int rc = layoutGrid(s, bmin, bmax);
    if (rc) return rc;
    mark(s, PH_PREDICT);
    if ((rc = phasePredictKey(s, 0.0f, false, true))) return rc;  // keys from the current x* (K2)
    mark(s, PH_SORT);
    if ((rc = phaseSortReorderLists(s))) return rc;
    mark(s, PH_SOLVE); mark(s, PH_POST); mark(s, PH_END);            // so akua_pbf_last_step_timing covers this call too
    s->timedIters = 0;
    s->timingValid = s->timing;
    return AKUA_OK;
}
int akua_pbf_phase_solve(akua_pbf_solver* s, int32_t iters, const float boxMin[3], const float boxMax[3]) {
    if (!s || !boxMin || !boxMax || iters < 0) return AKUA_ERR_INVALID;
    if (int g = phaseApiGuard(s)) return g;
    AK_CUDA(s, cudaSetDevice(s->device));
    rememberBox(s, boxMin, boxMax);
    bool committed;
    return phaseSolve(s, iters, boxMin, boxMax, false, 0.0f, &committed);
}
int akua_pbf_phase_update(akua_pbf_solver* s, float dt) {
    if (int g = phaseApiGuard(s)) return g;
    AK_CUDA(s, cudaSetDevice(s->device));
    return phaseUpdate(s, dt);
}
int akua_pbf_phase_damping(akua_pbf_solver* s, const float boxMin[3], const float boxMax[3]) {
    if (!s || !boxMin || !boxMax) return AKUA_ERR_INVALID;
    if (int g = phaseApiGuard(s)) return g;
    AK_CUDA(s, cudaSetDevice(s->device));
    return phaseDamping(s, boxMin, boxMax);
}
int akua_pbf_phase_vorticity_viscosity(akua_pbf_solver* s, float dt) {
    if (int g = phaseApiGuard(s)) return g;
    AK_CUDA(s, cudaSetDevice(s->device));
    return phasePost(s, dt, 0);
}

// ---- debug taps ----
int64_t akua_pbf_debug_size(akua_pbf_solver* s, int32_t which) {
    if (!s) return -1;
    if (s->slab.enabled) { cudaSetDevice(s->device); slabRefresh(s); }
    const int64_t n = s->n;
    switch (which) {
        case AKUA_DBG_KEYS_UNSORTED: case AKUA_DBG_KEYS_SORTED: case AKUA_DBG_PERM: case AKUA_DBG_ID:
        case AKUA_DBG_NBR_COUNT: case AKUA_DBG_DENSITY: case AKUA_DBG_LAMBDA: return n * 4;
        case AKUA_DBG_BUCKET_START: return s->opt.key_mode == AKUA_KEY_REFERENCE_HASH ? (int64_t)s->grid.tableSize * 4 : -1;
        case AKUA_DBG_CELL_RANGE: return s->opt.key_mode == AKUA_KEY_LINEAR_CELL ? s->ctr.num_cells * 8 : -1;
        case AKUA_DBG_NBR_LIST: return n * (int64_t)s->cfg.maxNeighbours * 4;
        case AKUA_DBG_XSTAR: case AKUA_DBG_POSITION: case AKUA_DBG_VELOCITY: case AKUA_DBG_VORTICITY:
        case AKUA_DBG_DELTA_P: return n * 16;
        default: return -1;
    }
}
int akua_pbf_debug_get(akua_pbf_solver* s, int32_t which, void* dst, int64_t dst_bytes) {
    if (!s || !dst) return AKUA_ERR_INVALID;
    const int64_t need = akua_pbf_debug_size(s, which);
    if (need < 0 || dst_bytes < need) { s->err = "debug_get: array not available or destination too small"; return AKUA_ERR_INVALID; }
    AK_CUDA(s, cudaSetDevice(s->device));
    if (need == 0) return AKUA_OK;
    const void* src = nullptr;
    switch (which) {
        case AKUA_DBG_KEYS_UNSORTED: src = s->keysUnsorted; break;
        case AKUA_DBG_KEYS_SORTED: src = s->keysSorted; break;
        case AKUA_DBG_PERM: src = s->perm; break;
        case AKUA_DBG_ID: src = s->id; break;
        case AKUA_DBG_BUCKET_START: src = s->bucketStart; break;
        case AKUA_DBG_CELL_RANGE: src = s->cellRange; break;
        case AKUA_DBG_NBR_COUNT: src = s->nbrCount; break;
        case AKUA_DBG_DENSITY: src = s->density; break;
        case AKUA_DBG_LAMBDA: src = s->lambda; break;
        case AKUA_DBG_XSTAR: src = s->xs; break;
        case AKUA_DBG_POSITION: src = s->pos; break;
        case AKUA_DBG_VELOCITY: src = s->vel; break;
        case AKUA_DBG_VORTICITY: src = s->omega; break;
        case AKUA_DBG_DELTA_P: src = s->dpos; break;
        case AKUA_DBG_NBR_LIST: {
            uint32_t* tmp = nullptr;
            AK_CUDA(s, cudaMalloc(&tmp, (size_t)need));
            uint64_t total = (uint64_t)s->n * s->cfg.maxNeighbours;
            launchPlain(s->stream, k_list_to_rowmajor, gridFor(total), kBlock, s->nbrList, s->nbrCount, s->nbrStride, (uint32_t)s->n,
                (uint32_t)s->cfg.maxNeighbours, tmp);
            AK_LAUNCH_CHECK(s, "k_list_to_rowmajor");
            cudaError_t e = cudaMemcpyAsync(dst, tmp, (size_t)need, cudaMemcpyDeviceToHost, s->stream);
            if (e == cudaSuccess) e = cudaStreamSynchronize(s->stream);
            cudaFree(tmp);
            if (e != cudaSuccess) { s->err = std::string("debug_get: ") + cudaGetErrorString(e); return AKUA_ERR_CUDA; }
            s->ctr.d2h_bytes += need;
            return AKUA_OK;
        }
        default: return AKUA_ERR_INVALID;
    }
    AK_CUDA(s, cudaMemcpyAsync(dst, src, (size_t)need, cudaMemcpyDeviceToHost, s->stream));
    AK_CUDA(s, cudaStreamSynchronize(s->stream));
    s->ctr.d2h_bytes += need;
    return AKUA_OK;
}

int akua_pbf_density_error(akua_pbf_solver* s, float* mean, float* maxv) {
    if (!s) return AKUA_ERR_INVALID;
    AK_CUDA(s, cudaSetDevice(s->device));
    {
        int rc = slabRefresh(s);
        if (rc) return rc;
    }
    const uint32_t n = (uint32_t)s->n;
    if (n == 0) { if (mean) *mean = 0; if (maxv) *maxv = 0; return AKUA_OK; }
    uint32_t blocks = gridFor(n);
    if (blocks > 1024) blocks = 1024;
    launchPlain(s->stream, k_density_error, blocks, kBlock, s->density, n, 1.0f / s->cfg.restDensity, s->partSum, s->partMax);
    AK_LAUNCH_CHECK(s, "k_density_error");
    std::vector<float> hs(blocks), hm(blocks);
    AK_CUDA(s, cudaMemcpyAsync(hs.data(), s->partSum, blocks * 4, cudaMemcpyDeviceToHost, s->stream));
    AK_CUDA(s, cudaMemcpyAsync(hm.data(), s->partMax, blocks * 4, cudaMemcpyDeviceToHost, s->stream));
    AK_CUDA(s, cudaStreamSynchronize(s->stream));
    double sum = 0; float mx = 0;
    for (uint32_t b = 0; b < blocks; b++) { sum += hs[b]; mx = std::fmax(mx, hm[b]); }
    if (mean) *mean = (float)(sum / n);
    if (maxv) *maxv = mx;
    return AKUA_OK;
}

int akua_pbf_get_counters(const akua_pbf_solver* s, akua_pbf_counters* out) {
    if (!s || !out) return AKUA_ERR_INVALID;
    *out = s->ctr;
    return AKUA_OK;
}
int akua_pbf_enable_timing(akua_pbf_solver* s, int32_t on) {
    if (!s) return AKUA_ERR_INVALID;
    s->timing = on != 0;
    s->timingValid = false;
    return AKUA_OK;
}
void* akua_pbf_stream(akua_pbf_solver* s) { return s ? (void*)s->stream : nullptr; }
int akua_pbf_trace_next_step(akua_pbf_solver* s, const char* path) {
    if (!s || !path || !*path) return AKUA_ERR_INVALID;
    s->tracePath = path;
    s->traceArmed = true;
    return AKUA_OK;
}
int akua_pbf_last_step_timing(akua_pbf_solver* s, float ms[10]) {
    if (!s || !ms) return AKUA_ERR_INVALID;
    if (!s->timingValid) { s->err = "no timed step recorded (call akua_pbf_enable_timing first)"; return AKUA_ERR_INVALID; }
    AK_CUDA(s, cudaSetDevice(s->device));
    AK_CUDA(s, cudaEventSynchronize(s->ev[PH_END]));
    for (int p = 0; p < 6; p++) AK_CUDA(s, cudaEventElapsedTime(&ms[p], s->ev[p], s->ev[p + 1]));
    AK_CUDA(s, cudaEventElapsedTime(&ms[6], s->ev[PH_PREDICT], s->ev[PH_END]));
    ms[7] = ms[8] = 0.0f;
    ms[9] = (float)s->timedIters;
    for (int it = 0; it < s->timedIters; it++) {
        float a = 0, b = 0;
        AK_CUDA(s, cudaEventElapsedTime(&a, s->evPass[it][0], s->evPass[it][1]));
        AK_CUDA(s, cudaEventElapsedTime(&b, s->evPass[it][1], s->evPass[it][2]));
        ms[7] += a; ms[8] += b;
    }
    return AKUA_OK;
}

}  // extern "C"
