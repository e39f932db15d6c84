// pbf_slab.inl — host side of the x-slab multi-GPU step (included by pbf_solver.cu inside its anonymous namespace).
//
// One process per GPU. Rank r owns the particles whose predicted position x* lies in grid x-planes [xLo, xHi) (absolute
// cell coordinates floor(x/h); the first / last rank own everything below / above). Layout of every per-particle array
// during a step:   [ owned, key-sorted | ghost plane from the left rank | ghost plane from the right rank ].
// Because keys are x-major, the planes a rank sends are the first and last contiguous stretch of its owned range, and
// what it receives is already sorted: no pack/unpack kernels and no re-sort on the per-iteration path, only
// ncclSend/ncclRecv of array slices on the solver's stream. Exchanges per step (1-cell halo): x* once for the neighbour
// search; lambda after pass A and x* after pass B in every iteration; v after the commit; |omega| after K11; v after K12.
// Migration happens once per step, right after the prediction: leavers are compacted deterministically into send
// buffers and get a sentinel key that sorts them out of the owned range; arrivals are appended before the sort.
// NCCL is loaded with dlopen so single-GPU users need no NCCL at all.

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

// Bound of the k_wait_flags spins in SM clock cycles (~20 s at 1.97 GHz): long enough for any host-side skew between the
// ranks' launch loops (the count exchange at the top of a step is where a late rank is waited for), short enough that a
// lost neighbour ends in an error instead of a hung GPU.
constexpr long long kFlagWaitCycles = 40000000000LL;

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

// ---- ghost-plane exchange --------------------------------------------------------------------------------------------
// Ghost planes live at FIXED offsets at the top of every per-particle array (ghostBaseL for the plane received from the
// left rank, ghostBaseR for the one from the right), so neither side needs the other's particle count.
// Two transports:
//   * CUDA IPC (default when the neighbours' allocations can be opened): the plane is copied straight from this rank's
//     array into the neighbour's ghost region over NVLink by the copy engine (cudaMemcpyAsync to the peer-mapped pointer)
//     and the exchange's epoch is published in the neighbour's flag word; the neighbour's main stream runs a bounded
//     spin-wait kernel (k_wait_flags) before the kernel that reads the ghosts. No NCCL kernel, no SM time, no rendezvous.
//   * NCCL send/recv (fallback; AKUA_SLAB_P2P=0): one grouped call per exchange.
// A SlabTicket is what the main stream has to wait for before it touches the ghosts of an exchange.
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
int slabPlanesOnComm(akua_pbf_solver* s, T* arr) {
    SlabState& sl = s->slab;
    const bool hasL = sl.rank > 0, hasR = sl.rank + 1 < sl.nranks;
    const size_t nOwn = (size_t)s->n;
    cudaStream_t st = sl.commStream;
    if (sl.p2p) {
        // my first plane -> the left rank's "from the right" ghost region; my last plane -> the right rank's "from the left"
        if (hasL && sl.nPlaneL) {
            T* dst = peerOf(s, sl.peerL, arr);
            if (!dst) { s->err = "slab p2p: array has no peer mapping"; return AKUA_ERR_COMM; }
            AK_CUDA(s, cudaMemcpyAsync(dst + sl.peerL.ghostBaseR, arr, (size_t)sl.nPlaneL * sizeof(T), cudaMemcpyDeviceToDevice, st));
        }
        if (hasR && sl.nPlaneR) {
            T* dst = peerOf(s, sl.peerR, arr);
            if (!dst) { s->err = "slab p2p: array has no peer mapping"; return AKUA_ERR_COMM; }
            AK_CUDA(s, cudaMemcpyAsync(dst + sl.peerR.ghostBaseL, arr + (nOwn - sl.nPlaneR), (size_t)sl.nPlaneR * sizeof(T),
                                       cudaMemcpyDeviceToDevice, st));
        }
    } else {
        ncclComm_t comm = (ncclComm_t)sl.comm;
        AK_NCCL(s, g_nccl.GroupStart());
        if (hasL && sl.nPlaneL) AK_NCCL(s, g_nccl.Send(arr, (size_t)sl.nPlaneL * sizeof(T), ncclUint8, sl.rank - 1, comm, st));
        if (hasR && sl.nPlaneR) AK_NCCL(s, g_nccl.Send(arr + (nOwn - sl.nPlaneR), (size_t)sl.nPlaneR * sizeof(T), ncclUint8, sl.rank + 1, comm, st));
        if (hasL && sl.nGhostL) AK_NCCL(s, g_nccl.Recv(arr + sl.ghostBaseL, (size_t)sl.nGhostL * sizeof(T), ncclUint8, sl.rank - 1, comm, st));
        if (hasR && sl.nGhostR) AK_NCCL(s, g_nccl.Recv(arr + sl.ghostBaseR, (size_t)sl.nGhostR * sizeof(T), ncclUint8, sl.rank + 1, comm, st));
        AK_NCCL(s, g_nccl.GroupEnd());
    }
    sl.exchanges++;
    sl.bytesSent += ((hasL ? sl.nPlaneL : 0) + (hasR ? sl.nPlaneR : 0)) * sizeof(T);
    return AKUA_OK;
}
// Closes an exchange on the comm stream: publishes the epoch to the neighbours (p2p) and records the local event.
int slabFinishExchange(akua_pbf_solver* s, SlabTicket* t) {
    SlabState& sl = s->slab;
    const bool hasL = sl.rank > 0, hasR = sl.rank + 1 < sl.nranks;
    t->epoch = 0;
    t->valid = true;
    if (sl.p2p) {
        t->epoch = ++sl.epoch;
        if (hasL) { slab::k_signal_flag<<<1, 1, 0, sl.commStream>>>(sl.peerL.flags + 1, t->epoch); AK_LAUNCH_CHECK(s, "k_signal_flag"); }
        if (hasR) { slab::k_signal_flag<<<1, 1, 0, sl.commStream>>>(sl.peerR.flags + 0, t->epoch); AK_LAUNCH_CHECK(s, "k_signal_flag"); }
    }
    t->ev = slabNextEvent(s);
    AK_CUDA(s, cudaEventRecord(t->ev, sl.commStream));
    return AKUA_OK;
}
// Asynchronous ghost-plane exchange: ordered after the main stream's work so far, runs on the comm stream.
template <typename T>
int slabExchangeAsync(akua_pbf_solver* s, T* arr, SlabTicket* t) {
    int rc;
    if ((rc = slabCommAfterMain(s))) return rc;
    if ((rc = slabPlanesOnComm(s, arr))) return rc;
    return slabFinishExchange(s, t);
}
template <typename T, typename U>
int slabExchangeAsync2(akua_pbf_solver* s, T* a, U* b, SlabTicket* t) {
    int rc;
    if ((rc = slabCommAfterMain(s))) return rc;
    if ((rc = slabPlanesOnComm(s, a))) return rc;
    if ((rc = slabPlanesOnComm(s, b))) return rc;
    return slabFinishExchange(s, t);
}
// Main stream: do not run past this point before the exchange behind `t` has delivered this rank's ghosts (and has
// finished reading this rank's planes).
int slabWait(akua_pbf_solver* s, const SlabTicket& t) {
    SlabState& sl = s->slab;
    if (!t.valid) return AKUA_OK;
    if (t.ev) AK_CUDA(s, cudaStreamWaitEvent(s->stream, t.ev, 0));
    if (sl.p2p) {
        const int hasL = sl.rank > 0, hasR = sl.rank + 1 < sl.nranks;
        if (hasL || hasR) {
            slab::k_wait_flags<<<1, 1, 0, s->stream>>>(sl.flags, hasL, hasR, t.epoch, sl.dCounts + 31, kFlagWaitCycles);
            AK_LAUNCH_CHECK(s, "k_wait_flags");
        }
    }
    return AKUA_OK;
}
// Fused path: the boundary kernel that just ran on the main stream already stored its planes into the neighbours' ghost
// regions (PeerPush); all that is left is to publish the epoch, in stream order right behind that kernel.
int slabSignalAfterKernel(akua_pbf_solver* s, SlabTicket* t) {
    SlabState& sl = s->slab;
    const bool hasL = sl.rank > 0, hasR = sl.rank + 1 < sl.nranks;
    t->valid = true; t->ev = nullptr; t->epoch = ++sl.epoch;
    if (hasL) { slab::k_signal_flag<<<1, 1, 0, s->stream>>>(sl.peerL.flags + 1, t->epoch); AK_LAUNCH_CHECK(s, "k_signal_flag"); }
    if (hasR) { slab::k_signal_flag<<<1, 1, 0, s->stream>>>(sl.peerR.flags + 0, t->epoch); AK_LAUNCH_CHECK(s, "k_signal_flag"); }
    sl.exchanges++;
    return AKUA_OK;
}
int slabHalo(akua_pbf_solver* s, const SlabTicket& waitFor, SlabTicket* out, bool launchHappens, HaloSync* hs) {
    SlabState& sl = s->slab;
    const bool hasL = sl.rank > 0, hasR = sl.rank + 1 < sl.nranks;
    *hs = HaloSync{};
    hs->errWord = sl.dCounts + 31;
    if (!launchHappens) {
        // no boundary particles on this rank: nothing reads ghosts, but the neighbours still expect the epoch
        if (out) return slabSignalAfterKernel(s, out);
        return AKUA_OK;
    }
    if (waitFor.valid) {
        if (waitFor.ev) AK_CUDA(s, cudaStreamWaitEvent(s->stream, waitFor.ev, 0));
        hs->waitFlags = sl.flags; hs->waitL = hasL ? 1 : 0; hs->waitR = hasR ? 1 : 0; hs->waitEpoch = waitFor.epoch;
    }
    if (out) {
        out->valid = true; out->ev = nullptr; out->epoch = ++sl.epoch;
        hs->signalL = hasL ? sl.peerL.flags + 1 : nullptr;
        hs->signalR = hasR ? sl.peerR.flags + 0 : nullptr;
        hs->signalEpoch = out->epoch;
        hs->doneCounter = sl.flags + 3;
        sl.exchanges++;
    }
    return AKUA_OK;
}
template <typename T>
PeerPush slabPush(const akua_pbf_solver* s, T* arr) {
    const SlabState& sl = s->slab;
    PeerPush pp;
    if (!sl.p2p) return pp;
    const bool hasL = sl.rank > 0, hasR = sl.rank + 1 < sl.nranks;
    const uint32_t nOwn = (uint32_t)s->n;
    if (hasL && sl.nPlaneL) { T* d = peerOf(s, sl.peerL, arr); if (d) { pp.dstL = d + sl.peerL.ghostBaseR; pp.nL = sl.nPlaneL; } }
    if (hasR && sl.nPlaneR) { T* d = peerOf(s, sl.peerR, arr); if (d) { pp.dstR = d + sl.peerR.ghostBaseL; pp.startR = nOwn - sl.nPlaneR; } }
    return pp;
}
// Blocking flavour: the main stream waits for the exchange.
template <typename T>
int slabExchangePlanes(akua_pbf_solver* s, T* arr) {
    if (!s->slab.enabled) return AKUA_OK;
    SlabTicket t;
    int rc = slabExchangeAsync(s, arr, &t);
    if (rc) return rc;
    return slabWait(s, t);
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
    uint32_t* d = sl.dCounts + 8;                             // words 8..11 are not used by the step
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

// Interior / boundary index spans of the owned range for the current step's plane sizes.
SweepSpans sweepSpans(const akua_pbf_solver* s) {
    const SlabState& sl = s->slab;
    const uint32_t n = (uint32_t)s->n, pl = sl.nPlaneL, pr = sl.nPlaneR;
    SweepSpans sp;
    if ((uint64_t)pl + pr >= n) {  // slab only one or two planes wide: everything is boundary
        sp.boundary = Span{n, 0u, 0xffffffffu, 0u};
        sp.interior = Span{0u, 0u, 0xffffffffu, 0u};
    } else {
        sp.boundary = Span{pl + pr, 0u, pl, n - pr - pl};
        sp.interior = Span{n - pl - pr, pl, 0xffffffffu, 0u};
    }
    return sp;
}

// The one count exchange of a step: three u32 to each neighbour (assembled by k_mig_scan at dCounts[16..18] for the left
// rank, [20..22] for the right rank), three from each (landing at [24..26] from the left, [28..30] from the right), then
// all 32 counters go to the host — the only host synchronisation of the step.
int slabSwapCounts(akua_pbf_solver* s) {
    SlabState& sl = s->slab;
    const bool hasL = sl.rank > 0, hasR = sl.rank + 1 < sl.nranks;
    if (sl.p2p) {
        // CUDA-IPC transport: counters (and, through k_mig_pack, the migration records) are stored straight into the
        // neighbours' memory; everything stays on the main stream — no NCCL kernel, no stream hop.
        const uint32_t epoch = ++sl.countEpoch;
        slab::k_publish_counts<<<1, 32, 0, s->stream>>>(sl.dCounts, hasL ? sl.peerL.dCounts : nullptr, hasR ? sl.peerR.dCounts : nullptr,
                                                        hasL ? sl.peerL.flags + 5 : nullptr, hasR ? sl.peerR.flags + 4 : nullptr, epoch);
        AK_LAUNCH_CHECK(s, "k_publish_counts");
        slab::k_wait_flags<<<1, 1, 0, s->stream>>>(sl.flags + 4, hasL ? 1 : 0, hasR ? 1 : 0, epoch, sl.dCounts + 31, kFlagWaitCycles);
        AK_LAUNCH_CHECK(s, "k_wait_flags");
        AK_CUDA(s, cudaMemcpyAsync(sl.hCounts, sl.dCounts, 32 * sizeof(uint32_t), cudaMemcpyDeviceToHost, s->stream));
        AK_CUDA(s, cudaStreamSynchronize(s->stream));
        if (!hasL) sl.hCounts[24] = sl.hCounts[25] = sl.hCounts[26] = 0;
        if (!hasR) sl.hCounts[28] = sl.hCounts[29] = sl.hCounts[30] = 0;
        return AKUA_OK;
    }
    ncclComm_t comm = (ncclComm_t)sl.comm;
    cudaStream_t st = sl.commStream;
    int rc;
    if ((rc = slabCommAfterMain(s))) return rc;
    AK_NCCL(s, g_nccl.GroupStart());
    if (hasL) AK_NCCL(s, g_nccl.Send(sl.dCounts + 16, 12, ncclUint8, sl.rank - 1, comm, st));
    if (hasR) AK_NCCL(s, g_nccl.Send(sl.dCounts + 20, 12, ncclUint8, sl.rank + 1, comm, st));
    if (hasL) AK_NCCL(s, g_nccl.Recv(sl.dCounts + 24, 12, ncclUint8, sl.rank - 1, comm, st));
    if (hasR) AK_NCCL(s, g_nccl.Recv(sl.dCounts + 28, 12, ncclUint8, sl.rank + 1, comm, st));
    AK_NCCL(s, g_nccl.GroupEnd());
    AK_CUDA(s, cudaMemcpyAsync(sl.hCounts, sl.dCounts, 32 * sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
    AK_CUDA(s, cudaStreamSynchronize(st));
    if (!hasL) sl.hCounts[24] = sl.hCounts[25] = sl.hCounts[26] = 0;
    if (!hasR) sl.hCounts[28] = sl.hCounts[29] = sl.hCounts[30] = 0;
    return AKUA_OK;
}

int stepSlab(akua_pbf_solver* s, float dt, int iterations, const float* bmin, const float* bmax) {
    SlabState& sl = s->slab;
    if (s->opt.key_mode != AKUA_KEY_LINEAR_CELL) { s->err = "slab mode needs LINEAR_CELL keys"; return AKUA_ERR_INVALID; }
    int rc = layoutGrid(s, bmin, bmax);
    if (rc) return rc;
    const GridParams& G = s->grid;
    const uint32_t planeCells = (uint32_t)G.gridDim.y * (uint32_t)G.gridDim.z;
    // ownership in grid-relative x planes; the end ranks own the clamped border planes too
    int xLo = sl.rank == 0 ? 0 : std::min(std::max(sl.xLoAbs - G.gridMin.x, 0), G.gridDim.x);
    int xHi = sl.rank + 1 == sl.nranks ? G.gridDim.x : std::min(std::max(sl.xHiAbs - G.gridMin.x, 0), G.gridDim.x);
    if (xHi <= xLo) { s->err = "slab is empty in the current grid (box does not cover this rank's x range)"; return AKUA_ERR_INVALID; }
    const uint32_t sentinel = (uint32_t)s->ctr.num_cells;  // one past the last valid key
    const int sortBits = bitsFor((uint64_t)sentinel);
    uint32_t n = (uint32_t)s->n;

    // ---- 1. predict + key (PBFSolver.cpp:30 + K2) ----
    mark(s, PH_PREDICT);
    if ((rc = phasePredictKey(s, dt, true, true))) return rc;

    // ---- 2. migration: leavers out (sentinel key), arrivals appended ----
    mark(s, PH_SORT);
    const uint32_t blocks = std::max(1u, gridFor(n));
    if (blocks > sl.migBlocksCap) { s->err = "slab: migration scratch too small"; return AKUA_ERR_INVALID; }
    if (xHi - xLo < 2 && sl.nranks > 1) { s->err = "slab must be at least two x planes wide"; return AKUA_ERR_INVALID; }
    AK_CUDA(s, cudaMemsetAsync(sl.dCounts + 2, 0, 4 * sizeof(uint32_t), s->stream));   // the four plane populations
    slab::k_mig_count<<<blocks, kBlock, 0, s->stream>>>(s->keysUnsorted, n, planeCells, xLo, xHi, sl.blockCnt, sl.dCounts + 2);
    AK_LAUNCH_CHECK(s, "k_mig_count");
    slab::k_mig_scan<<<1, 1024, 0, s->stream>>>(sl.blockCnt, blocks, sl.dCounts);
    AK_LAUNCH_CHECK(s, "k_mig_scan");
    // p2p transport: leavers are packed straight into the neighbours' inboxes (my left neighbour receives them "from its
    // right"); otherwise into local send buffers that NCCL ships after the count exchange
    const bool hasLn = sl.rank > 0, hasRn = sl.rank + 1 < sl.nranks;
    slab::MigRecord* outBufL = (sl.p2p && hasLn) ? sl.peerL.recvR : sl.sendL;
    slab::MigRecord* outBufR = (sl.p2p && hasRn) ? sl.peerR.recvL : sl.sendR;
    uint32_t packCap = sl.migCap;
    if (sl.p2p && hasLn) packCap = std::min(packCap, sl.peerL.migCap);
    if (sl.p2p && hasRn) packCap = std::min(packCap, sl.peerR.migCap);
    slab::k_mig_pack<<<blocks, kBlock, 0, s->stream>>>(s->keysUnsorted, n, planeCells, xLo, xHi, sl.blockCnt, sentinel, s->pos,
                                                       s->vel, s->xs, s->id, outBufL, outBufR, packCap);
    AK_LAUNCH_CHECK(s, "k_mig_pack");
    if ((rc = slabSwapCounts(s))) return rc;
    const uint32_t* hc = sl.hCounts;
    if (hc[31] == 2) { s->err = "slab p2p: timed out waiting for a neighbour's ghost planes in the previous step"; return AKUA_ERR_COMM; }
    if (hc[31]) { s->err = "slab: boundary-plane size prediction failed in the previous step (a particle crossed more than one slab?)"; return AKUA_ERR_INVALID; }
    const uint32_t outL = hc[0], outR = hc[1], inL = hc[24], inR = hc[28];
    // Post-migration plane sizes, known before the sort: my boundary planes = stayers + arrivals that land in them;
    // a neighbour's facing plane (= my ghosts) = its stayers there + my leavers that land there. (Arrivals from the far
    // side cannot reach the near plane: slabs are >= 2 planes wide and the stepper moves particles by << one slab.)
    sl.nPlaneL = hasLn ? hc[2] + hc[25] : 0;
    sl.nPlaneR = hasRn ? hc[3] + hc[29] : 0;
    sl.nGhostL = hasLn ? hc[26] + hc[4] : 0;
    sl.nGhostR = hasRn ? hc[30] + hc[5] : 0;
    if (sl.rank == 0 && outL) { s->err = "slab: internal error (leavers beyond the first rank)"; return AKUA_ERR_INVALID; }
    if (outL > packCap || outR > packCap || inL > sl.migCap || inR > sl.migCap) { s->err = "slab: migration buffer overflow (raise capacity_factor)"; return AKUA_ERR_ALLOC; }
    if ((uint64_t)n + inL + inR > (uint64_t)sl.ghostBaseL) { s->err = "slab: particle capacity exceeded by arrivals (raise capacity_factor)"; return AKUA_ERR_ALLOC; }
    if (sl.p2p) {   // the records arrived with the count message
        sl.exchanges++;
        sl.bytesSent += ((size_t)outL + outR) * sizeof(slab::MigRecord);
    } else {
        const bool hasL = sl.rank > 0, hasR = sl.rank + 1 < sl.nranks;
        ncclComm_t comm = (ncclComm_t)sl.comm;
        cudaStream_t st = sl.commStream;   // already ordered after the pack kernel by the count swap's host sync
        AK_NCCL(s, g_nccl.GroupStart());
        if (hasL && outL) AK_NCCL(s, g_nccl.Send(sl.sendL, (size_t)outL * sizeof(slab::MigRecord), ncclUint8, sl.rank - 1, comm, st));
        if (hasR && outR) AK_NCCL(s, g_nccl.Send(sl.sendR, (size_t)outR * sizeof(slab::MigRecord), ncclUint8, sl.rank + 1, comm, st));
        if (hasL && inL) AK_NCCL(s, g_nccl.Recv(sl.recvL, (size_t)inL * sizeof(slab::MigRecord), ncclUint8, sl.rank - 1, comm, st));
        if (hasR && inR) AK_NCCL(s, g_nccl.Recv(sl.recvR, (size_t)inR * sizeof(slab::MigRecord), ncclUint8, sl.rank + 1, comm, st));
        AK_NCCL(s, g_nccl.GroupEnd());
        {
            cudaEvent_t e = slabNextEvent(s);
            AK_CUDA(s, cudaEventRecord(e, st));
            AK_CUDA(s, cudaStreamWaitEvent(s->stream, e, 0));
        }
        sl.exchanges++;
        sl.bytesSent += ((size_t)outL + outR) * sizeof(slab::MigRecord);
    }
    if (inL) { slab::k_mig_unpack<<<gridFor(inL), kBlock, 0, s->stream>>>(sl.recvL, inL, n, s->pos, s->vel, s->xs, s->id, s->keysUnsorted, G); AK_LAUNCH_CHECK(s, "k_mig_unpack"); }
    if (inR) { slab::k_mig_unpack<<<gridFor(inR), kBlock, 0, s->stream>>>(sl.recvR, inR, n + inL, s->pos, s->vel, s->xs, s->id, s->keysUnsorted, G); AK_LAUNCH_CHECK(s, "k_mig_unpack"); }
    const uint32_t nPre = n + inL + inR;
    const uint32_t nOwn = nPre - outL - outR;
    sl.migratedIn += inL + inR; sl.migratedOut += outL + outR;

    // ---- 3. sort everything resident (leavers end up past nOwn), reorder the owned range, owned cell ranges ----
    AK_CUDA(s, cudaMemsetAsync(s->cellRange, 0, (size_t)s->ctr.num_cells * sizeof(uint2), s->stream));
    if (nPre) {
        int launches = rsort::sort_pairs(s->keysUnsorted, s->keyA, s->valA, s->keyB, s->valB, nPre, sortBits, s->sortWs, s->stream,
                                         &s->keysSorted, &s->perm, usePdl(s));
        cudaError_t e = cudaGetLastError();
        if (e != cudaSuccess) { s->err = std::string("radix sort: ") + cudaGetErrorString(e); return AKUA_ERR_CUDA; }
        s->ctr.kernel_launches += launches;
        s->ctr.sort_passes_last = launches / 3;
    }
    mark(s, PH_REORDER);
    s->n = nOwn;
    if (nOwn) {
        launchK(s, k_reorder_ranges<KEY_LINEAR>, gridFor(nOwn), kBlock, s->keysSorted, s->perm, nOwn, s->pos, s->vel, s->xs, s->id,
            s->posAlt, s->velAlt, s->xsAlt, s->idAlt, s->bucketStart, s->cellRange);
        AK_LAUNCH_CHECK(s, "k_reorder_ranges");
    }
    std::swap(s->pos, s->posAlt); std::swap(s->vel, s->velAlt); std::swap(s->xs, s->xsAlt); std::swap(s->id, s->idAlt);

    // ---- 4. ghost planes: sizes, then x* of the neighbours' boundary planes, keyed and ranged in place ----
    slab::k_plane_verify<<<1, 32, 0, s->stream>>>(s->keysSorted, nOwn, planeCells, xLo, xHi, sl.nPlaneL, sl.nPlaneR,
                                                 hasLn ? 1 : 0, hasRn ? 1 : 0, sl.dCounts);
    AK_LAUNCH_CHECK(s, "k_plane_verify");
    if (sl.nGhostL > sl.ghostCap || sl.nGhostR > sl.ghostCap || sl.nPlaneL > sl.ghostCap || sl.nPlaneR > sl.ghostCap) {
        s->err = "slab: boundary plane larger than the ghost region (raise capacity_factor)"; return AKUA_ERR_ALLOC;
    }
    if ((rc = slabExchangePlanes(s, s->xs))) return rc;
    for (int side = 0; side < 2; side++) {   // ghost planes: keys from the received x*, then their cell ranges
        const uint32_t cntG = side == 0 ? sl.nGhostL : sl.nGhostR, base = side == 0 ? sl.ghostBaseL : sl.ghostBaseR;
        if (!cntG) continue;
        float3 g0 = make_float3(0, 0, 0);
        launchK(s, k_predict_key<KEY_LINEAR>, gridFor(cntG), kBlock, nullptr, nullptr, s->xs + base, s->keysSorted + base, cntG, 0.0f, g0, G, 0);
        AK_LAUNCH_CHECK(s, "k_predict_key(ghosts)");
        slab::k_ranges<<<gridFor(cntG), kBlock, 0, s->stream>>>(s->keysSorted, base, base + cntG, s->cellRange);
        AK_LAUNCH_CHECK(s, "k_ranges(ghosts)");
    }

    // ---- 5. neighbour lists of the owned particles (candidates include the ghost planes) ----
    mark(s, PH_LISTS);
    if ((rc = launchBuildNeighbours(s, nOwn))) return rc;

    sl.pending = SlabTicket{};
    // ---- 6. constraint solve and post-solve on the owned range, with the per-pass ghost exchanges inside ----
    mark(s, PH_SOLVE);
    bool committed = false;
    if ((rc = phaseSolve(s, iterations, bmin, bmax, true, dt, &committed))) return rc;
    if (!committed) {  // solverIterations == 0
        if ((rc = phaseUpdate(s, dt))) return rc;
        if ((rc = phaseDamping(s, bmin, bmax))) return rc;
        if ((rc = slabExchangeAsync(s, s->vel, &sl.pending))) return rc;
    }
    mark(s, PH_POST);
    if ((rc = phasePost(s, dt))) return rc;
    mark(s, PH_END);
    s->timingValid = s->timing;
    s->ctr.steps++;
    return AKUA_OK;
}

// Re-balances the slab boundaries from the current per-x-plane particle counts of all ranks (collective: every rank
// calls it at the same step). The owned particles are sorted by x plane after a step, so each rank histograms its own
// planes with binary searches, the histograms are summed with one ncclAllReduce, and every rank computes the same new
// boundaries (akua_slab_partition). A boundary may only move inside the two slabs it separates, and by no more
// particles than the migration buffers hold, so the ordinary per-step migration of the NEXT step performs the transfer.
int slabRebalance(akua_pbf_solver* s) {
    SlabState& sl = s->slab;
    if (!sl.enabled || !s->haveBox) { s->err = "rebalance: slab mode with at least one completed step required"; return AKUA_ERR_INVALID; }
    const GridParams& G = s->grid;
    const int gx = G.gridDim.x, R = sl.nranks;
    const uint32_t planeCells = (uint32_t)G.gridDim.y * (uint32_t)G.gridDim.z;
    const size_t words = (size_t)gx + R;  // [0,gx): plane counts; [gx, gx+R): current lower bounds (grid-relative)
    if (words > sl.histCap) {
        if (sl.dHist) cudaFree(sl.dHist);
        if (sl.hHist) cudaFreeHost(sl.hHist);
        sl.dHist = nullptr; sl.hHist = nullptr;
        AK_CUDA(s, dalloc(&sl.dHist, words));
        AK_CUDA(s, cudaMallocHost((void**)&sl.hHist, words * sizeof(unsigned long long)));
        sl.histCap = words;
    }
    AK_CUDA(s, cudaMemsetAsync(sl.dHist, 0, words * sizeof(unsigned long long), s->stream));
    const uint32_t n = (uint32_t)s->n;
    slab::k_plane_hist<<<(gx + 255) / 256, 256, 0, s->stream>>>(s->keysSorted, n, planeCells, gx, sl.dHist);
    AK_LAUNCH_CHECK(s, "k_plane_hist");
    const int curLo = sl.rank == 0 ? 0 : std::min(std::max(sl.xLoAbs - G.gridMin.x, 0), gx);
    unsigned long long lo64 = (unsigned long long)curLo;
    AK_CUDA(s, cudaMemcpyAsync(sl.dHist + gx + sl.rank, &lo64, sizeof(lo64), cudaMemcpyHostToDevice, s->stream));
    int rc;
    if ((rc = slabCommAfterMain(s))) return rc;
    AK_NCCL(s, g_nccl.AllReduce(sl.dHist, sl.dHist, words, ncclUint64, ncclSum, (ncclComm_t)sl.comm, sl.commStream));
    AK_CUDA(s, cudaMemcpyAsync(sl.hHist, sl.dHist, words * sizeof(unsigned long long), cudaMemcpyDeviceToHost, sl.commStream));
    AK_CUDA(s, cudaStreamSynchronize(sl.commStream));
    std::vector<int64_t> hist(gx);
    for (int x = 0; x < gx; x++) hist[x] = (int64_t)sl.hHist[x];
    std::vector<int32_t> bounds(R + 1), old(R + 1);
    for (int r = 0; r < R; r++) old[r] = (int32_t)sl.hHist[gx + r];
    old[0] = 0; old[R] = gx;
    if (akua_slab_rebalance_bounds(hist.data(), gx, R, old.data(), (int64_t)sl.migCap / 2, bounds.data()) != AKUA_OK) {
        s->err = "rebalance: grid has fewer x planes than ranks"; return AKUA_ERR_INVALID;
    }
    // monotonic by construction (each stays within its old neighbours' interval); take this rank's new interval
    sl.xLoAbs = G.gridMin.x + bounds[sl.rank];
    sl.xHiAbs = G.gridMin.x + bounds[sl.rank + 1];
    sl.rebalances++;
    return AKUA_OK;
}
