// pbf_slab.inl — host side of the x-slab multi-GPU step (included by pbf_solver.cu inside its anonymous namespace).
//
// One process per GPU. Rank r owns the particles whose predicted position x* lies in grid x-planes [xLo, xHi) (absolute
// cell coordinates floor(x/h); the first / last rank own everything below / above). Layout of every per-particle array
// during a step:   [ owned, key-sorted | ghost plane from the left rank | ghost plane from the right rank ].
// Because keys are x-major, the planes a rank sends are the first and last contiguous stretch of its owned range, and
// what it receives is already sorted: no pack/unpack kernels and no re-sort on the per-iteration path. With the CUDA-IPC
// transport (default) the first CTAs of every sweep store those planes straight into the neighbours' ghost regions and
// publish an epoch the neighbours' CTAs wait for (pbf_kernels.cuh: sweep_cta, peer_push, halo_wait / halo_signal); every size
// of the step lives on the device (dims) and the step is one CUDA graph. The fallback is ncclSend/ncclRecv of array slices on
// the solver's stream with one host synchronisation per step. Exchanges per step (1-cell halo): x* once for the neighbour
// search; lambda after pass A and x* after pass B in every iteration; v after the commit; |omega| after K11; v after K12.
// Migration happens once per step, right after the prediction: leavers are compacted deterministically into the
// neighbours' inboxes and get a sentinel key that sorts them out of the owned range; arrivals are appended before the sort.
// NCCL (set-up, re-balancing, fallback transport) is loaded with dlopen so single-GPU users need no NCCL at all.

struct NcclApi {
    void* lib = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*Send)(const void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Recv)(void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
};
NcclApi g_nccl;

const char* loadNccl() {
    if (g_nccl.lib) return nullptr;
#ifdef AKUA_HOST_EMU   // tests/emu: the in-process stand-in of tests/emu/nccl.h (test infrastructure only)
    g_nccl.GetUniqueId = ncclGetUniqueId; g_nccl.CommInitRank = ncclCommInitRank; g_nccl.CommDestroy = ncclCommDestroy;
    g_nccl.Send = ncclSend; g_nccl.Recv = ncclRecv; g_nccl.GroupStart = ncclGroupStart; g_nccl.GroupEnd = ncclGroupEnd;
    g_nccl.AllReduce = ncclAllReduce; g_nccl.GetErrorString = ncclGetErrorString;
    g_nccl.lib = &g_nccl;
    return nullptr;
#endif
    void* lib = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!lib) lib = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
    if (!lib) return "cannot dlopen libnccl.so.2";
#define AK_SYM(field, name) \
    *(void**)(&g_nccl.field) = dlsym(lib, name); \
    if (!g_nccl.field) return "libnccl is missing " name;
    AK_SYM(GetUniqueId, "ncclGetUniqueId") AK_SYM(CommInitRank, "ncclCommInitRank") AK_SYM(CommDestroy, "ncclCommDestroy")
    AK_SYM(Send, "ncclSend") AK_SYM(Recv, "ncclRecv") AK_SYM(GroupStart, "ncclGroupStart") AK_SYM(GroupEnd, "ncclGroupEnd")
    AK_SYM(GetErrorString, "ncclGetErrorString") AK_SYM(AllReduce, "ncclAllReduce")
#undef AK_SYM
    g_nccl.lib = lib;
    return nullptr;
}

#define AK_NCCL(s, call)                                                                          \
    do {                                                                                          \
        ncclResult_t r_ = (call);                                                                 \
        if (r_ != ncclSuccess) {                                                                  \
            (s)->err = std::string(#call) + ": " + g_nccl.GetErrorString(r_);                     \
            return AKUA_ERR_COMM;                                                                 \
        }                                                                                         \
    } while (0)

// Bound of the in-kernel flag waits in SM clock cycles (~10 s at 1.97 GHz): long enough for any host-side skew between the
// ranks' launch loops (the count exchange at the top of a step is where a late rank is waited for), short enough that a
// lost neighbour ends in an error instead of a hung GPU. One bound for every wait (count message and ghost planes); after the
// first time-out of a rank every later wait returns at once (spin_until), so a dead neighbour costs this once.
constexpr long long kFlagWaitCyclesDefault = 20000000000LL;
long long flagWaitCycles() {   // AKUA_SLAB_WAIT_CYCLES: override for tests (a time-out must be reachable in a unit test)
    static const long long v = [] { const char* e = std::getenv("AKUA_SLAB_WAIT_CYCLES"); return e ? std::atoll(e) : kFlagWaitCyclesDefault; }();
    return v > 0 ? v : kFlagWaitCyclesDefault;
}
#define kFlagWaitCycles flagWaitCycles()

cudaEvent_t slabNextEvent(akua_pbf_solver* s) {
    SlabState& sl = s->slab;
    cudaEvent_t e = sl.evPool[sl.evNext];
    sl.evNext = (sl.evNext + 1) % SlabState::kEvents;
    return e;
}
// comm stream waits for everything issued on the main stream so far
int slabCommAfterMain(akua_pbf_solver* s) {
    cudaEvent_t e = slabNextEvent(s);
    AK_CUDA(s, cudaEventRecord(e, s->stream));
    AK_CUDA(s, cudaStreamWaitEvent(s->slab.commStream, e, 0));
    return AKUA_OK;
}
int slabMainAfterComm(akua_pbf_solver* s) {
    cudaEvent_t e = slabNextEvent(s);
    AK_CUDA(s, cudaEventRecord(e, s->slab.commStream));
    AK_CUDA(s, cudaStreamWaitEvent(s->stream, e, 0));
    return AKUA_OK;
}

const char* slabErrorText(uint32_t bits) {
    if (bits & SLAB_ERR_TIMEOUT) return "slab: timed out waiting for a neighbour rank (count message or ghost planes)";
    if (bits & SLAB_ERR_CAPACITY) return "slab: particle capacity exceeded by arrivals (raise capacity_factor)";
    if (bits & SLAB_ERR_MIG_OVERFLOW) return "slab: migration buffer overflow (raise capacity_factor)";
    if (bits & SLAB_ERR_GHOST_OVERFLOW) return "slab: boundary plane larger than the ghost region (raise capacity_factor)";
    if (bits & SLAB_ERR_PLANE_PREDICTION) return "slab: boundary-plane size prediction failed (a particle crossed more than one slab in one step?)";
    return "slab: unknown device-side error";
}
// Reports a sticky device-side error the pinned mirror already shows (no synchronisation: the mirror is refreshed at the end
// of every step, so an error surfaces at the latest one call after the step that raised it, and at every synchronising call).
int slabCheckError(akua_pbf_solver* s) {
    const uint32_t e = s->slab.hDims ? s->slab.hDims[D_ERROR] : 0u;
    if (!e) return AKUA_OK;
    s->err = slabErrorText(e);
    return (e & (SLAB_ERR_TIMEOUT | SLAB_ERR_PLANE_PREDICTION)) ? AKUA_ERR_COMM : AKUA_ERR_ALLOC;
}
// Synchronises the solver's stream and makes the host's view of the step sizes exact (s->n, plane / ghost sizes).
int slabRefresh(akua_pbf_solver* s) {
    SlabState& sl = s->slab;
    if (!sl.enabled) return AKUA_OK;
    AK_CUDA(s, cudaStreamSynchronize(s->stream));
    AK_CUDA(s, cudaMemcpy((void*)sl.hDims, sl.dims, D_WORDS * sizeof(uint32_t), cudaMemcpyDeviceToHost));
    s->n = sl.hDims[D_N];
    sl.nPlaneL = sl.hDims[D_PLANE_L]; sl.nPlaneR = sl.hDims[D_PLANE_R];
    sl.nGhostL = sl.hDims[D_GHOST_L]; sl.nGhostR = sl.hDims[D_GHOST_R];
    return slabCheckError(s);
}
// After an upload (or checkpoint load) of n particles: owned count, identity payload slots, full free-slot stack.
int slabResetCounts(akua_pbf_solver* s) {
    SlabState& sl = s->slab;
    if (!sl.dims) return AKUA_OK;
    const uint32_t n = (uint32_t)s->n, cap = (uint32_t)s->capacity;
    launchPlain(s->stream, slab::k_slab_reset, std::min<uint32_t>(gridFor(cap), 148 * 8), kBlock, sl.dims, n, cap, sl.slot, sl.freeSlots);
    AK_LAUNCH_CHECK(s, "k_slab_reset");
    AK_CUDA(s, cudaStreamSynchronize(s->stream));
    sl.hDims[D_N] = n; sl.hDims[D_NOWN] = n;
    sl.hDims[D_PLANE_L] = sl.hDims[D_PLANE_R] = sl.hDims[D_GHOST_L] = sl.hDims[D_GHOST_R] = 0;
    sl.hDims[D_IN_L] = sl.hDims[D_IN_R] = 0;
    // (the launch-size estimates stay: they only size grids, and keeping them keeps the step's CUDA graph across uploads)
    return AKUA_OK;
}

// ---- ghost-plane exchange --------------------------------------------------------------------------------------------
// Ghost planes live at FIXED offsets at the top of every per-particle array (ghostBaseL for the plane received from the
// left rank, ghostBaseR for the one from the right), so neither side needs the other's particle count.
// Two transports:
//   * CUDA IPC (default when the neighbours' allocations can be opened): the kernel that PRODUCES a boundary plane stores it
//     straight into the neighbour's ghost region (P2P stores over NVLink, PeerPush) and its last CTA publishes the exchange's
//     epoch in the neighbour's flag word; the kernel that CONSUMES ghosts waits in-kernel for the epoch (HaloSync). No copy
//     engine, no NCCL kernel, no rendezvous, no host-side size: the step is one CUDA graph.
//   * NCCL send/recv (fallback; AKUA_SLAB_P2P=0): one grouped call per exchange with host-known sizes, which costs one host
//     synchronisation per step and rules out graph capture.
template <typename T> T* peerOf(const akua_pbf_solver* s, const SlabPeer& peer, T* mine) {
    const SlabState& sl = s->slab;
    const void* m = mine;
    if (m == sl.xsBuf[0]) return reinterpret_cast<T*>(peer.xsBuf[0]);
    if (m == sl.xsBuf[1]) return reinterpret_cast<T*>(peer.xsBuf[1]);
    if (m == sl.velBuf[0]) return reinterpret_cast<T*>(peer.velBuf[0]);
    if (m == sl.velBuf[1]) return reinterpret_cast<T*>(peer.velBuf[1]);
    if (m == s->lambda) return reinterpret_cast<T*>(peer.lambda);
    if (m == s->omegaLen) return reinterpret_cast<T*>(peer.omegaLen);
    if (m == s->xl) return reinterpret_cast<T*>(peer.xl);
    if (m == s->xw) return reinterpret_cast<T*>(peer.xw);
    return nullptr;
}
template <typename T>
PeerPush slabPush(const akua_pbf_solver* s, T* arr) {
    const SlabState& sl = s->slab;
    PeerPush pp;
    if (!sl.p2p) return pp;
    const bool hasL = sl.rank > 0, hasR = sl.rank + 1 < sl.nranks;
    // my first plane -> the left rank's "from the right" ghost region; my last plane -> the right rank's "from the left"
    if (hasL) { T* d = peerOf(s, sl.peerL, arr); if (d) pp.dstL = d + sl.peerL.ghostBaseR; }
    if (hasR) { T* d = peerOf(s, sl.peerR, arr); if (d) pp.dstR = d + sl.peerR.ghostBaseL; }
    pp.dims = sl.dims;   // plane sizes are resolved on the device
    return pp;
}
HaloSync slabHalo(const akua_pbf_solver* s, int waitIdx, int signalIdx) {
    const SlabState& sl = s->slab;
    const bool hasL = sl.rank > 0, hasR = sl.rank + 1 < sl.nranks;
    HaloSync hs;
    hs.waitFlags = sl.flags; hs.waitL = hasL ? 1 : 0; hs.waitR = hasR ? 1 : 0; hs.waitIdx = waitIdx;
    hs.signalL = hasL ? sl.peerL.flags + 1 : nullptr;
    hs.signalR = hasR ? sl.peerR.flags + 0 : nullptr;
    hs.signalIdx = signalIdx;
    hs.doneCounter = sl.flags + 3;
    hs.dims = sl.dims;
    hs.timeoutCycles = kFlagWaitCycles;
    return hs;
}
// NCCL fallback: blocking (in stream order on the main stream) exchange of the boundary planes of `arr`, host-known sizes.
template <typename T>
int slabNcclPlanes(akua_pbf_solver* s, T* arr) {
    SlabState& sl = s->slab;
    const bool hasL = sl.rank > 0, hasR = sl.rank + 1 < sl.nranks;
    const size_t nOwn = (size_t)s->n;
    ncclComm_t comm = (ncclComm_t)sl.comm;
    cudaStream_t st = s->stream;
    AK_NCCL(s, g_nccl.GroupStart());
    if (hasL && sl.nPlaneL) AK_NCCL(s, g_nccl.Send(arr, (size_t)sl.nPlaneL * sizeof(T), ncclUint8, sl.rank - 1, comm, st));
    if (hasR && sl.nPlaneR) AK_NCCL(s, g_nccl.Send(arr + (nOwn - sl.nPlaneR), (size_t)sl.nPlaneR * sizeof(T), ncclUint8, sl.rank + 1, comm, st));
    if (hasL && sl.nGhostL) AK_NCCL(s, g_nccl.Recv(arr + sl.ghostBaseL, (size_t)sl.nGhostL * sizeof(T), ncclUint8, sl.rank - 1, comm, st));
    if (hasR && sl.nGhostR) AK_NCCL(s, g_nccl.Recv(arr + sl.ghostBaseR, (size_t)sl.nGhostR * sizeof(T), ncclUint8, sl.rank + 1, comm, st));
    AK_NCCL(s, g_nccl.GroupEnd());
    return AKUA_OK;
}
// p2p: pushes the boundary planes of `arr` with a small copy kernel (no sweep produces them) and publishes exchange `idx`.
template <typename T>
int slabPushPlanes(akua_pbf_solver* s, T* arr, int idx, const slab::PlaneVerify& pv = slab::PlaneVerify{}) {
    SlabState& sl = s->slab;
    if (!sl.p2p) return slabNcclPlanes(s, arr);
    const PeerPush pp = slabPush(s, arr);
    const HaloSync hs = slabHalo(s, -1, idx);
    launchK(s, slab::k_push_planes<T>, std::max(1u, gridFor(sl.estBnd)), kBlock, (const T*)arr, pp, hs, pv);
    AK_LAUNCH_CHECK(s, "k_push_planes");
    return AKUA_OK;
}

// Packed gather layout in x-slab mode: pass B / K12 read ghosts through the packed (x*, lambda) / (x, |omega|) arrays and the
// sweeps multiply by ONE mass, so the layout is only used when every particle of every rank has the same mass and every
// rank holds the packed arrays. Global uniformity cannot change through migration, only through uploads, so the verdict is
// taken where the masses enter: in akua_pbf_set_slab and, once slab mode is on, in every upload — which makes those calls
// COLLECTIVE in slab mode (every rank calls them, like akua_pbf_rebalance). One MIN all-reduce of (lo, ~hi, has-arrays).
int slabAgreeMass(akua_pbf_solver* s) {
    SlabState& sl = s->slab;
    if (!sl.enabled || !sl.comm || sl.nranks < 2) return AKUA_OK;
    uint32_t lo = 0xffffffffu, hi = 0u;                       // neutral: a rank without particles
    if (s->n > 0) {
        if (s->massUniform) { uint32_t b; std::memcpy(&b, &s->uniformMass, 4); lo = hi = (b & 0x80000000u) ? ~b : (b | 0x80000000u); }
        else { lo = 0u; hi = 0xffffffffu; }
    }
    uint32_t msg[4] = {lo, ~hi, (s->xl && s->xw) ? 1u : 0u, 0u};
    uint32_t* d = sl.dims + D_MASS;                           // words 8..11 are not used by the step
    AK_CUDA(s, cudaStreamSynchronize(s->stream));
    AK_CUDA(s, cudaMemcpy(d, msg, sizeof(msg), cudaMemcpyHostToDevice));
    AK_NCCL(s, g_nccl.AllReduce(d, d, 4, ncclUint32, ncclMin, (ncclComm_t)sl.comm, sl.commStream));
    AK_CUDA(s, cudaStreamSynchronize(sl.commStream));
    AK_CUDA(s, cudaMemcpy(msg, d, sizeof(msg), cudaMemcpyDeviceToHost));
    const uint32_t glo = msg[0], ghi = ~msg[1];
    s->massUniform = msg[2] != 0 && glo == ghi && glo != 0xffffffffu;
    const uint32_t bits = (glo & 0x80000000u) ? (glo ^ 0x80000000u) : ~glo;
    std::memcpy(&s->uniformMass, &bits, 4);
    return AKUA_OK;
}

// ---- the slab step ------------------------------------------------------------------------------------------------------
// Launch-size estimates from the (asynchronously refreshed, possibly a step or two old) pinned mirror of dims. They only size
// grids (every kernel loops over the device-side counts, so any estimate is correct) but they are arguments of the step's CUDA
// graph: an estimate that changes re-captures it. So: the owned count is bucketed with hysteresis; the plane, ghost and arrival
// sizes — small launches whose grids cost nothing when oversized, and which jump after every re-balancing — only ever GROW,
// with 50 % head-room (reset by an upload).
void slabEstimate(uint32_t actual, uint32_t bucket, uint32_t* est) {
    if (actual > *est || (uint64_t)actual + 2ull * bucket < *est || *est == 0)
        *est = (uint32_t)(((uint64_t)actual + bucket / 2 + bucket) / bucket * bucket);
}
void slabEstimateGrowOnly(uint32_t actual, uint32_t bucket, uint32_t* est) {
    if (actual > *est || *est == 0)
        *est = (uint32_t)(((uint64_t)actual + actual / 2 + bucket) / bucket * bucket);
}
void slabUpdateEstimates(akua_pbf_solver* s) {
    SlabState& sl = s->slab;
    const uint32_t cap = (uint32_t)s->capacity;
    const uint32_t bN = std::max(8192u, cap / 64), bS = std::max(2048u, cap / 1024);
    slabEstimate(sl.hDims[D_N], bN, &sl.estN);
    slabEstimateGrowOnly(sl.hDims[D_PLANE_L] + sl.hDims[D_PLANE_R], bS, &sl.estBnd);
    slabEstimateGrowOnly(sl.hDims[D_GHOST_L] + sl.hDims[D_GHOST_R], bS, &sl.estGhost);
    slabEstimateGrowOnly(sl.hDims[D_IN_L] + sl.hDims[D_IN_R], bS, &sl.estIn);
    sl.estN = std::min(sl.estN, cap);
}

// Slab-local cell grid: the global grid of the box (shared by all ranks: same y / z extent and origin) restricted in x to a
// WINDOW of planes around this rank's slab: its owned planes, towards each neighbour the ghost plane and one "far" plane into
// which everything beyond is clamped (a leaver that lands deeper than the neighbour's boundary plane must not be counted into
// that plane), plus windowMargin() spare planes on each inner side. Keys are window-relative: a 64 M-particle scene on 8 GPUs
// sorts 23-bit keys (3 digit passes) and clears an 8th of the cell table. The window — and with it every kernel argument of
// the step: grid, sentinel key, sort passes, cell-table size — only changes when a re-balancing pushes a boundary out of the
// margin; the owned interval itself lives in dims[D_XLO / D_XHI], so an ordinary re-balancing costs two words written to the
// device and NO re-capture of the step's CUDA graph.
// Spare planes per inner side: a quarter of the slab's width, between 6 and 32 (a sloshing tank on 8 GPUs moves a boundary by up
// to eight planes between two re-balancing calls ten steps apart: profiles/r02_c14_*).
inline int windowMargin(int width) { return std::max(6, std::min(32, width / 4)); }
int slabLayout(akua_pbf_solver* s, const float* bmin, const float* bmax) {
    SlabState& sl = s->slab;
    if (s->opt.key_mode != AKUA_KEY_LINEAR_CELL) { s->err = "slab mode needs LINEAR_CELL keys"; return AKUA_ERR_INVALID; }
    int3 gmin, gdim;
    const int64_t cellsGlobal = layout_linear_grid(s->cfg.smoothRadius, bmin, bmax, &gmin, &gdim);
    if (cellsGlobal < 0) { s->err = "box max must exceed box min"; return AKUA_ERR_INVALID; }
    const bool hasL = sl.rank > 0, hasR = sl.rank + 1 < sl.nranks;
    // ownership in global grid planes; the end ranks own the clamped border planes too
    const int xLoG = !hasL ? 0 : std::min(std::max(sl.xLoAbs - gmin.x, 0), gdim.x);
    const int xHiG = !hasR ? gdim.x : std::min(std::max(sl.xHiAbs - gmin.x, 0), gdim.x);
    if (xHiG <= xLoG) { s->err = "slab is empty in the current grid (box does not cover this rank's x range)"; return AKUA_ERR_INVALID; }
    if (xHiG - xLoG < 2 && sl.nranks > 1) { s->err = "slab must be at least two x planes wide"; return AKUA_ERR_INVALID; }
    const int needLo = hasL ? std::max(xLoG - 2, 0) : 0, needHi = hasR ? std::min(xHiG + 2, gdim.x) : gdim.x;
    const bool sameGrid = sl.winValid && sl.winGmin.x == gmin.x && sl.winGmin.y == gmin.y && sl.winGmin.z == gmin.z &&
                          sl.winGdim.x == gdim.x && sl.winGdim.y == gdim.y && sl.winGdim.z == gdim.z;
    if (!sameGrid || needLo < sl.winX0 || needHi > sl.winX1) {
        const int margin = windowMargin(xHiG - xLoG);
        sl.winX0 = hasL ? std::max(needLo - margin, 0) : 0;
        sl.winX1 = hasR ? std::min(needHi + margin, gdim.x) : gdim.x;
        sl.winGmin = gmin; sl.winGdim = gdim; sl.winValid = true;
        sl.windowChanges++;
    }
    const int x0 = sl.winX0, x1 = sl.winX1;
    const int64_t cells = (int64_t)(x1 - x0) * gdim.y * gdim.z;
    if (cells + 1 >= (int64_t)1 << 31) { s->err = "LINEAR_CELL grid too large (>= 2^31 cells)"; return AKUA_ERR_INVALID; }
    if (cells > s->cellCapacity) {
        const double t0 = hostMs();
        if (s->cellRange) { AK_CUDA(s, cudaStreamSynchronize(s->stream)); AK_CUDA(s, cudaFree(s->cellRange)); }
        s->cellRange = nullptr;
        const int64_t want = cells + cells / 4;   // head-room: the next, slightly wider window does not reallocate
        AK_CUDA(s, dalloc(&s->cellRange, (size_t)want));
        s->cellCapacity = want;
        if (slabVerbose()) std::fprintf(stderr, "[akua rank %d] cell table grown to %lld cells in %.2f ms (host)\n", sl.rank, (long long)want, hostMs() - t0);
    }
    s->grid.gridMin = make_int3(gmin.x + x0, gmin.y, gmin.z);
    s->grid.gridDim = make_int3(x1 - x0, gdim.y, gdim.z);
    s->ctr.num_cells = cells;
    const int xLoL = xLoG - x0, xHiL = xHiG - x0;
    if (xLoL != sl.xLoL || xHiL != sl.xHiL || !sl.intervalOnDevice) {
        // stream-ordered ahead of the step that uses it (and outside any capture: slabLayout runs before the step is replayed)
        sl.hInterval[0] = (uint32_t)xLoL; sl.hInterval[1] = (uint32_t)xHiL;
        AK_CUDA(s, cudaMemcpyAsync(sl.dims + D_XLO, (const void*)sl.hInterval, 2 * sizeof(uint32_t), cudaMemcpyHostToDevice, s->stream));
        sl.intervalOnDevice = true;
    }
    sl.xLoL = xLoL; sl.xHiL = xHiL; sl.planeOffset = x0; sl.gxGlobal = gdim.x;
    sl.sentinel = (uint32_t)cells;                 // one past the last valid key: leavers sort behind the owned range
    sl.sortBits = bitsFor((uint64_t)cells);
    s->keyBits = sl.sortBits;
    return AKUA_OK;
}

int stepSlabBody(akua_pbf_solver* s, float dt, int iterations, const float* bmin, const float* bmax) {
    SlabState& sl = s->slab;
    int rc;
    const GridParams& G = s->grid;
    const uint32_t planeCells = (uint32_t)G.gridDim.y * (uint32_t)G.gridDim.z;
    const bool hasL = sl.rank > 0, hasR = sl.rank + 1 < sl.nranks;
    const bool p2p = sl.p2p;
    uint32_t* dims = sl.dims;

    // ---- 1. predict + key (PBFSolver.cpp:30 + K2) ----
    mark(s, PH_PREDICT);
    if ((rc = phasePredictKey(s, dt, true, true))) return rc;

    // ---- 2. migration: leavers out (sentinel key), arrivals appended ----
    mark(s, PH_SORT);
    const uint32_t migGrid = std::max(1u, std::min((sl.estN + slab::kMigTile - 1) / slab::kMigTile, sl.migBlocksCap));
    AK_CUDA(s, cudaMemsetAsync(dims + D_STAY_FIRST, 0, 4 * sizeof(uint32_t), s->stream));   // the four plane populations
    launchPlain(s->stream, slab::k_mig_count, migGrid, kBlock, s->keysUnsorted, (const uint32_t*)dims, planeCells, sl.blockCnt,
                                                         sl.migBlocksCap, dims + D_STAY_FIRST);
    AK_LAUNCH_CHECK(s, "k_mig_count");
    launchPlain(s->stream, slab::k_mig_scan, 1, 1024, sl.blockCnt, dims + D_N, sl.migBlocksCap, dims);
    AK_LAUNCH_CHECK(s, "k_mig_scan");
    // p2p transport: leavers are packed straight into the neighbours' inboxes (my left neighbour receives them "from its
    // right"); otherwise into local send buffers that NCCL ships after the count exchange
    slab::MigRecord* outBufL = (p2p && hasL) ? sl.peerL.recvR : sl.sendL;
    slab::MigRecord* outBufR = (p2p && hasR) ? sl.peerR.recvL : sl.sendR;
    uint32_t packCap = sl.migCap;
    if (p2p && hasL) packCap = std::min(packCap, sl.peerL.migCap);
    if (p2p && hasR) packCap = std::min(packCap, sl.peerR.migCap);
    launchPlain(s->stream, slab::k_mig_pack, migGrid, kBlock, s->keysUnsorted, dims, planeCells, sl.blockCnt, sl.migBlocksCap,
                                                        sl.sentinel, s->pos, s->vel, s->xs, s->id, sl.slot, s->color, s->size,
                                                        sl.freeSlots, outBufL, outBufR, packCap);
    AK_LAUNCH_CHECK(s, "k_mig_pack");
    slab::PlanCaps caps{};
    caps.packCap = packCap; caps.migCap = sl.migCap; caps.ownedCap = sl.ghostBaseL; caps.ghostCap = sl.ghostCap;
    caps.planeCapL = hasL ? (p2p ? std::min(sl.ghostCap, sl.peerL.ghostCap) : sl.ghostCap) : 0;
    caps.planeCapR = hasR ? (p2p ? std::min(sl.ghostCap, sl.peerR.ghostCap) : sl.ghostCap) : 0;
    caps.slotCap = (uint32_t)s->capacity;
    if (p2p) {
        // the count message and the records travel by P2P stores: the plan kernel publishes this rank's message, then waits
        // in-kernel for the neighbours'
        slab::PlanPublish pub;
        if (hasL) { pub.peerDimsL = sl.peerL.dims; pub.peerFlagL = sl.peerL.flags + 5; }
        if (hasR) { pub.peerDimsR = sl.peerR.dims; pub.peerFlagR = sl.peerR.flags + 4; }
        launchPlain(s->stream, slab::k_slab_plan, 1, 32, dims, sl.flags + 4, hasL ? 1 : 0, hasR ? 1 : 0, 1, caps, kFlagWaitCycles, pub);
        AK_LAUNCH_CHECK(s, "k_slab_plan");
    } else {
        // NCCL fallback: count messages, then ONE host synchronisation (the record sizes are needed on the host), then records
        ncclComm_t comm = (ncclComm_t)sl.comm;
        cudaStream_t st = s->stream;
        AK_NCCL(s, g_nccl.GroupStart());
        if (hasL) AK_NCCL(s, g_nccl.Send(dims + D_MSG_TO_L, 12, ncclUint8, sl.rank - 1, comm, st));
        if (hasR) AK_NCCL(s, g_nccl.Send(dims + D_MSG_TO_R, 12, ncclUint8, sl.rank + 1, comm, st));
        if (hasL) AK_NCCL(s, g_nccl.Recv(dims + D_MSG_FROM_L, 12, ncclUint8, sl.rank - 1, comm, st));
        if (hasR) AK_NCCL(s, g_nccl.Recv(dims + D_MSG_FROM_R, 12, ncclUint8, sl.rank + 1, comm, st));
        AK_NCCL(s, g_nccl.GroupEnd());
        launchPlain(st, slab::k_slab_plan, 1, 32, dims, sl.flags + 4, hasL ? 1 : 0, hasR ? 1 : 0, 0, caps, kFlagWaitCycles, slab::PlanPublish{});
        AK_LAUNCH_CHECK(s, "k_slab_plan");
        AK_CUDA(s, cudaMemcpyAsync((void*)sl.hDims, dims, D_WORDS * sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
        AK_CUDA(s, cudaStreamSynchronize(st));
        if ((rc = slabCheckError(s))) return rc;
        const uint32_t outL = sl.hDims[D_OUT_L], outR = sl.hDims[D_OUT_R], inL = sl.hDims[D_IN_L], inR = sl.hDims[D_IN_R];
        AK_NCCL(s, g_nccl.GroupStart());
        if (hasL && outL) AK_NCCL(s, g_nccl.Send(sl.sendL, (size_t)outL * sizeof(slab::MigRecord), ncclUint8, sl.rank - 1, comm, st));
        if (hasR && outR) AK_NCCL(s, g_nccl.Send(sl.sendR, (size_t)outR * sizeof(slab::MigRecord), ncclUint8, sl.rank + 1, comm, st));
        if (hasL && inL) AK_NCCL(s, g_nccl.Recv(sl.recvL, (size_t)inL * sizeof(slab::MigRecord), ncclUint8, sl.rank - 1, comm, st));
        if (hasR && inR) AK_NCCL(s, g_nccl.Recv(sl.recvR, (size_t)inR * sizeof(slab::MigRecord), ncclUint8, sl.rank + 1, comm, st));
        AK_NCCL(s, g_nccl.GroupEnd());
        // the host now knows this step's sizes exactly
        s->n = sl.hDims[D_NOWN];
        sl.nPlaneL = sl.hDims[D_PLANE_L]; sl.nPlaneR = sl.hDims[D_PLANE_R];
        sl.nGhostL = sl.hDims[D_GHOST_L]; sl.nGhostR = sl.hDims[D_GHOST_R];
        sl.estN = std::max<uint32_t>(sl.estN, (uint32_t)sl.hDims[D_NPRE]);
    }
    launchPlain(s->stream, slab::k_mig_unpack, std::max(1u, gridFor(sl.estIn)), kBlock, sl.recvL, sl.recvR, dims, s->pos, s->vel, s->xs, s->id,
        sl.slot, s->color, s->size, sl.freeSlots, s->keysUnsorted, G);
    AK_LAUNCH_CHECK(s, "k_mig_unpack");

    // ---- 3. sort everything resident (leavers end up past nOwn), reorder the owned range, owned cell ranges ----
    AK_CUDA(s, cudaMemsetAsync(s->cellRange, 0, (size_t)s->ctr.num_cells * sizeof(uint2), s->stream));
    {
        const uint32_t nSort = std::min<uint32_t>(sl.estN + sl.estIn, (uint32_t)s->capacity);
        if ((rc = sortParticles(s, nSort, dims + D_NPRE, sl.sortBits))) return rc;
    }
    mark(s, PH_REORDER);
    launchK(s, k_reorder_ranges<KEY_LINEAR>, gridFor(sl.estN), kBlock, s->keysSorted, s->perm, sl.estN, dims + D_NOWN, s->pos, s->vel, s->xs,
        s->id, s->posAlt, s->velAlt, s->xsAlt, s->idAlt, s->bucketStart, s->cellRange, sl.slot, sl.slotAlt);
    AK_LAUNCH_CHECK(s, "k_reorder_ranges");
    std::swap(s->pos, s->posAlt); std::swap(s->vel, s->velAlt); std::swap(s->xs, s->xsAlt); std::swap(s->id, s->idAlt);
    std::swap(sl.slot, sl.slotAlt);

    // ---- 4. ghost planes: x* of the neighbours' boundary planes (exchange 1), keyed and ranged in place ----
    if (p2p) {   // the plane-size check rides on the push kernel
        slab::PlaneVerify pv;
        pv.keysSorted = s->keysSorted; pv.planeCells = planeCells; pv.hasL = hasL; pv.hasR = hasR; pv.dims = dims;
        if ((rc = slabPushPlanes(s, s->xs, 1, pv))) return rc;
    } else {
        launchPlain(s->stream, slab::k_plane_verify, 1, 32, s->keysSorted, planeCells, hasL ? 1 : 0, hasR ? 1 : 0, dims);
        AK_LAUNCH_CHECK(s, "k_plane_verify");
        if ((rc = slabPushPlanes(s, s->xs, 1))) return rc;
    }
    {
        const HaloSync hs = p2p ? slabHalo(s, 1, -1) : HaloSync{};
        launchK(s, slab::k_ghost_ranges, std::max(1u, gridFor(sl.estGhost)), kBlock, (const float4*)s->xs, s->keysSorted, (const uint32_t*)dims,
                sl.ghostBaseL, sl.ghostBaseR, s->cellRange, G, hs);
        AK_LAUNCH_CHECK(s, "k_ghost_ranges");
    }

    // ---- 5. neighbour lists of the owned particles (candidates include the ghost planes) ----
    mark(s, PH_LISTS);
    if ((rc = launchBuildNeighbours(s))) return rc;

    // ---- 6. constraint solve and post-solve on the owned range, with the per-pass ghost exchanges inside ----
    mark(s, PH_SOLVE);
    bool committed = false;
    if ((rc = phaseSolve(s, iterations, bmin, bmax, true, dt, &committed))) return rc;
    if (!committed) {  // solverIterations == 0
        if ((rc = phaseUpdate(s, dt))) return rc;
        if ((rc = phaseDamping(s, bmin, bmax))) return rc;
        if ((rc = slabPushPlanes(s, s->vel, 2))) return rc;
    }
    mark(s, PH_POST);
    if ((rc = phasePost(s, dt, iterations))) return rc;
    // ---- 7. close the step on the device and refresh the host's (asynchronous) view of its sizes ----
    {
        // bytes pushed per boundary-plane particle over the step: x* (16) once, per iteration (x*, lambda) 16 + x* 16,
        // v 16 after the commit, (x, |omega|) 16, v 16
        const uint32_t planeBytes = 16u + 32u * (uint32_t)iterations + 16u + 32u;
        launchPlain(s->stream, slab::k_step_end, 1, 32, dims, (uint32_t)slabExchangesPerStep(iterations), planeBytes);
        AK_LAUNCH_CHECK(s, "k_step_end");
        AK_CUDA(s, cudaMemcpyAsync((void*)sl.hDims, dims, D_WORDS * sizeof(uint32_t), cudaMemcpyDeviceToHost, s->stream));
    }
    sl.exchanges += slabExchangesPerStep(iterations);
    mark(s, PH_END);
    s->timingValid = s->timing;
    s->ctr.steps++;
    return AKUA_OK;
}

// Re-balances the slab boundaries from the current per-x-plane particle counts and work of all ranks (collective: every rank
// calls it at the same step). The owned particles are sorted by x plane after a step, so each rank histograms its own planes
// (binary searches + a sum of neighbour counts), the histograms are summed with one ncclAllReduce, and every rank computes the
// same new boundaries. A boundary may only move inside the two slabs it separates, and by no more particles than the
// migration buffers hold, so the ordinary per-step migration of the NEXT step performs the transfer.
// Two halves: slabRebalanceMeasure enqueues the histogram kernels, the all-reduce and the copy to pinned memory WITHOUT any
// host synchronisation; slabRebalanceApply waits for that copy and moves the boundaries. akua_pbf_rebalance = measure + apply
// (one pipeline drain per call); akua_pbf_rebalance_async = apply the PREVIOUS call's measurement, then measure again: the
// host never waits and the GPU never idles, at the price of boundaries that lag one call behind the fluid.
int slabRebalanceMeasure(akua_pbf_solver* s) {
    SlabState& sl = s->slab;
    if (!sl.enabled || !s->haveBox) { s->err = "rebalance: slab mode with at least one completed step required"; return AKUA_ERR_INVALID; }
    int rc;
    if ((rc = slabCheckError(s))) return rc;
    const GridParams& G = s->grid;   // slab-local grid of the last step; plane p of it is global plane planeOffset + p
    const int gx = sl.gxGlobal, R = sl.nranks;
    const uint32_t planeCells = (uint32_t)G.gridDim.y * (uint32_t)G.gridDim.z;
    // [0, gx): plane counts; [gx, 2 gx): plane work (particles weighted by their neighbour count); then per rank: its current
    // lower bound, its measured busy time per step (ns) and the sum of its work
    const size_t words = (size_t)2 * gx + 3 * (size_t)R;
    if (words > sl.histCap) {
        AK_CUDA(s, cudaStreamSynchronize(sl.commStream));
        if (sl.dHist) cudaFree(sl.dHist);
        if (sl.hHist) cudaFreeHost(sl.hHist);
        sl.dHist = nullptr; sl.hHist = nullptr;
        AK_CUDA(s, dalloc(&sl.dHist, words));
        AK_CUDA(s, cudaMallocHost((void**)&sl.hHist, words * sizeof(unsigned long long)));
        sl.histCap = words;
    }
    if (!sl.evRebalance) AK_CUDA(s, cudaEventCreateWithFlags(&sl.evRebalance, cudaEventDisableTiming));
    AK_CUDA(s, cudaMemsetAsync(sl.dHist, 0, words * sizeof(unsigned long long), s->stream));
    launchPlain(s->stream, slab::k_plane_hist, (uint32_t)G.gridDim.x, 256, s->keysSorted, s->nbrCount, (const uint32_t*)(sl.dims + D_N), planeCells,
                G.gridDim.x, sl.planeOffset, sl.dHist, sl.dHist + gx);
    AK_LAUNCH_CHECK(s, "k_plane_hist");
    launchPlain(s->stream, slab::k_rank_busy, 1, 256, (const unsigned long long*)(sl.dHist + gx), gx, sl.dims,
                sl.dHist + 2 * (size_t)gx + R + sl.rank, sl.dHist + 2 * (size_t)gx + 2 * (size_t)R + sl.rank, sl.planeOffset + sl.xLoL,
                sl.rank == 0 ? 1 : 0, sl.dHist + 2 * (size_t)gx + sl.rank);
    AK_LAUNCH_CHECK(s, "k_rank_busy");
    if ((rc = slabCommAfterMain(s))) return rc;
    AK_NCCL(s, g_nccl.AllReduce(sl.dHist, sl.dHist, words, ncclUint64, ncclSum, (ncclComm_t)sl.comm, sl.commStream));
    AK_CUDA(s, cudaMemcpyAsync(sl.hHist, sl.dHist, words * sizeof(unsigned long long), cudaMemcpyDeviceToHost, sl.commStream));
    AK_CUDA(s, cudaEventRecord(sl.evRebalance, sl.commStream));
    // the next step's memset of dHist must not overtake the all-reduce: the main stream waits for the comm stream's copy only
    // when the next measurement is enqueued (slabCommAfterMain orders comm after main; the reverse edge is the event below)
    sl.pendValid = true; sl.pendGx = gx; sl.pendGminGlobalX = G.gridMin.x - sl.planeOffset; sl.pendStep = s->ctr.steps;
    return AKUA_OK;
}
int slabRebalanceApply(akua_pbf_solver* s) {
    SlabState& sl = s->slab;
    if (!sl.pendValid) return AKUA_OK;
    const double tReb0 = hostMs();
    AK_CUDA(s, cudaEventSynchronize(sl.evRebalance));
    AK_CUDA(s, cudaStreamWaitEvent(s->stream, sl.evRebalance, 0));   // dHist is free for the next measurement's memset
    sl.pendValid = false;
    const int gx = sl.pendGx, R = sl.nranks;
    std::vector<int64_t> hist(gx), work(gx);
    for (int x = 0; x < gx; x++) { hist[x] = (int64_t)sl.hHist[x]; work[x] = (int64_t)sl.hHist[gx + x]; }
    std::vector<int32_t> bounds(R + 1), old(R + 1);
    for (int r = 0; r < R; r++) old[r] = (int32_t)sl.hHist[2 * (size_t)gx + r];
    old[0] = 0; old[R] = gx;
    // Feedback from the clock: a rank whose measured busy time per unit of work is above (below) the mean has its planes made
    // dearer (cheaper) by that ratio, clamped to [0.8, 1.25] — the work estimate decides the bulk, the measurement corrects what
    // it cannot know (how well a rank's gathers cache, a free surface, a wall). Applied only when every rank with particles had
    // a measurement since the last call.
    bool measured = true;
    double busyAll = 0.0, workAll = 0.0;
    for (int r = 0; r < R; r++) {
        const double b = (double)sl.hHist[2 * (size_t)gx + R + r], w = (double)sl.hHist[2 * (size_t)gx + 2 * (size_t)R + r];
        if (w > 0.0 && b <= 0.0) measured = false;
        busyAll += b; workAll += w;
    }
    // (the clock only means something when a step is long enough to be bound by the particles rather than by launch latency:
    // every rank's busy time must reach 2 ms per step)
    double busyMin = 1e30;
    for (int r = 0; r < R; r++) if ((double)sl.hHist[2 * (size_t)gx + 2 * (size_t)R + r] > 0.0) busyMin = std::min(busyMin, (double)sl.hHist[2 * (size_t)gx + R + r]);
    if (busyMin < 2.0e6) measured = false;
    if (measured && busyAll > 0.0 && workAll > 0.0) {
        for (int r = 0; r < R; r++) {
            const double b = (double)sl.hHist[2 * (size_t)gx + R + r], w = (double)sl.hHist[2 * (size_t)gx + 2 * (size_t)R + r];
            if (w <= 0.0) continue;
            const double c = std::min(1.25, std::max(0.8, (b / w) / (busyAll / workAll)));
            for (int x = std::max(old[r], 0); x < std::min(old[r + 1], gx); x++) work[x] = (int64_t)((double)work[x] * c);
        }
    } else measured = false;
    sl.lastRebalanceMeasured = measured;
    // no slab gets more particles than 80 % of the room below its ghost regions (the fluid keeps flowing between two calls)
    const int64_t maxCount = (int64_t)sl.ghostBaseL * 4 / 5;
    if (akua_slab_rebalance_bounds_weighted(work.data(), hist.data(), gx, R, old.data(), (int64_t)sl.migCap / 2, sl.keepBelow, maxCount, bounds.data()) != AKUA_OK) {
        s->err = "rebalance: grid has fewer x planes than ranks"; return AKUA_ERR_INVALID;
    }
    bool movedAny = false;
    for (int r = 0; r <= R; r++) movedAny = movedAny || bounds[r] != old[r];
    if (slabVerbose() && sl.rank == 0) {
        std::fprintf(stderr, "[akua] rebalance measured after step %lld, applied after step %lld: %s, %s, %.2f ms (host); bounds", (long long)sl.pendStep,
                     (long long)s->ctr.steps,
                     measured ? "work estimate x measured busy-time correction" : "work estimate", movedAny ? "moved" : "kept", hostMs() - tReb0);
        for (int r = 0; r <= R; r++) std::fprintf(stderr, " %d", bounds[r]);
        std::fprintf(stderr, "\n");
    }
    if (!movedAny) return AKUA_OK;
    // monotonic by construction (each stays within its old neighbours' interval); take this rank's new interval
    const int gminGlobalX = sl.pendGminGlobalX;
    sl.xLoAbs = gminGlobalX + bounds[sl.rank];
    sl.xHiAbs = gminGlobalX + bounds[sl.rank + 1];
    sl.rebalances++;
    return AKUA_OK;
}
