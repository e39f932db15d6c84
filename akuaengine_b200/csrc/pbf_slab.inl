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

// Sends [sendLoff, +sendLcnt) to the left rank and [sendRoff, +sendRcnt) to the right rank, receives recvLcnt elements
// from the left at recvLoff and recvRcnt from the right at recvRoff. One grouped NCCL call on the comm stream.
int slabExchange(akua_pbf_solver* s, void* base, size_t elemBytes, size_t sendLoff, size_t sendLcnt, size_t sendRoff,
                 size_t sendRcnt, size_t recvLoff, size_t recvLcnt, size_t recvRoff, size_t recvRcnt) {
    SlabState& sl = s->slab;
    const bool hasL = sl.rank > 0, hasR = sl.rank + 1 < sl.nranks;
    char* b = static_cast<char*>(base);
    ncclComm_t comm = (ncclComm_t)sl.comm;
    cudaStream_t st = sl.commStream;
    AK_NCCL(s, g_nccl.GroupStart());
    if (hasL && sendLcnt) AK_NCCL(s, g_nccl.Send(b + sendLoff * elemBytes, sendLcnt * elemBytes, ncclUint8, sl.rank - 1, comm, st));
    if (hasR && sendRcnt) AK_NCCL(s, g_nccl.Send(b + sendRoff * elemBytes, sendRcnt * elemBytes, ncclUint8, sl.rank + 1, comm, st));
    if (hasL && recvLcnt) AK_NCCL(s, g_nccl.Recv(b + recvLoff * elemBytes, recvLcnt * elemBytes, ncclUint8, sl.rank - 1, comm, st));
    if (hasR && recvRcnt) AK_NCCL(s, g_nccl.Recv(b + recvRoff * elemBytes, recvRcnt * elemBytes, ncclUint8, sl.rank + 1, comm, st));
    AK_NCCL(s, g_nccl.GroupEnd());
    sl.exchanges++;
    sl.bytesSent += (hasL ? sendLcnt : 0) * elemBytes + (hasR ? sendRcnt : 0) * elemBytes;
    return AKUA_OK;
}
template <typename T>
int slabPlanesOnComm(akua_pbf_solver* s, T* arr) {
    const SlabState& sl = s->slab;
    const size_t nOwn = (size_t)s->n;
    return slabExchange(s, arr, sizeof(T), 0, sl.nPlaneL, nOwn - sl.nPlaneR, sl.nPlaneR, nOwn, sl.nGhostL, nOwn + sl.nGhostL,
                        sl.nGhostR);
}
// Asynchronous ghost-plane exchange: ordered after the main stream's work so far, runs on the comm stream; `*done`
// must be waited on (cudaStreamWaitEvent) before the main stream touches the received ghosts.
template <typename T>
int slabExchangeAsync(akua_pbf_solver* s, T* arr, cudaEvent_t* done) {
    int rc;
    if ((rc = slabCommAfterMain(s))) return rc;
    if ((rc = slabPlanesOnComm(s, arr))) return rc;
    *done = slabNextEvent(s);
    AK_CUDA(s, cudaEventRecord(*done, s->slab.commStream));
    return AKUA_OK;
}
template <typename T, typename U>
int slabExchangeAsync2(akua_pbf_solver* s, T* a, U* b, cudaEvent_t* done) {
    int rc;
    if ((rc = slabCommAfterMain(s))) return rc;
    if ((rc = slabPlanesOnComm(s, a))) return rc;
    if ((rc = slabPlanesOnComm(s, b))) return rc;
    *done = slabNextEvent(s);
    AK_CUDA(s, cudaEventRecord(*done, s->slab.commStream));
    return AKUA_OK;
}
// Blocking flavour: the main stream waits for the exchange.
template <typename T>
int slabExchangePlanes(akua_pbf_solver* s, T* arr) {
    if (!s->slab.enabled) return AKUA_OK;
    cudaEvent_t done;
    int rc = slabExchangeAsync(s, arr, &done);
    if (rc) return rc;
    AK_CUDA(s, cudaStreamWaitEvent(s->stream, done, 0));
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
    slab::k_mig_pack<<<blocks, kBlock, 0, s->stream>>>(s->keysUnsorted, n, planeCells, xLo, xHi, sl.blockCnt, sentinel, s->pos,
                                                       s->vel, s->xs, s->id, sl.sendL, sl.sendR, sl.migCap);
    AK_LAUNCH_CHECK(s, "k_mig_pack");
    if ((rc = slabSwapCounts(s))) return rc;
    const uint32_t* hc = sl.hCounts;
    if (hc[31]) { s->err = "slab: boundary-plane size prediction failed in the previous step (a particle crossed more than one slab?)"; return AKUA_ERR_INVALID; }
    const uint32_t outL = hc[0], outR = hc[1], inL = hc[24], inR = hc[28];
    // Post-migration plane sizes, known before the sort: my boundary planes = stayers + arrivals that land in them;
    // a neighbour's facing plane (= my ghosts) = its stayers there + my leavers that land there. (Arrivals from the far
    // side cannot reach the near plane: slabs are >= 2 planes wide and the stepper moves particles by << one slab.)
    const bool hasLn = sl.rank > 0, hasRn = sl.rank + 1 < sl.nranks;
    sl.nPlaneL = hasLn ? hc[2] + hc[25] : 0;
    sl.nPlaneR = hasRn ? hc[3] + hc[29] : 0;
    sl.nGhostL = hasLn ? hc[26] + hc[4] : 0;
    sl.nGhostR = hasRn ? hc[30] + hc[5] : 0;
    if (sl.rank == 0 && outL) { s->err = "slab: internal error (leavers beyond the first rank)"; return AKUA_ERR_INVALID; }
    if (outL > sl.migCap || outR > sl.migCap || inL > sl.migCap || inR > sl.migCap) { s->err = "slab: migration buffer overflow (raise capacity_factor)"; return AKUA_ERR_ALLOC; }
    if ((uint64_t)n + inL + inR > (uint64_t)s->capacity) { s->err = "slab: particle capacity exceeded by arrivals (raise capacity_factor)"; return AKUA_ERR_ALLOC; }
    {
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
                                         &s->keysSorted, &s->perm);
        cudaError_t e = cudaGetLastError();
        if (e != cudaSuccess) { s->err = std::string("radix sort: ") + cudaGetErrorString(e); return AKUA_ERR_CUDA; }
        s->ctr.kernel_launches += launches;
        s->ctr.sort_passes_last = launches / 3;
    }
    mark(s, PH_REORDER);
    s->n = nOwn;
    if (nOwn) {
        k_reorder_ranges<KEY_LINEAR><<<gridFor(nOwn), kBlock, 0, s->stream>>>(s->keysSorted, s->perm, nOwn, s->pos, s->vel, s->xs, s->id,
            s->posAlt, s->velAlt, s->xsAlt, s->idAlt, s->bucketStart, s->cellRange);
        AK_LAUNCH_CHECK(s, "k_reorder_ranges");
    }
    std::swap(s->pos, s->posAlt); std::swap(s->vel, s->velAlt); std::swap(s->xs, s->xsAlt); std::swap(s->id, s->idAlt);

    // ---- 4. ghost planes: sizes, then x* of the neighbours' boundary planes, keyed and ranged in place ----
    slab::k_plane_verify<<<1, 32, 0, s->stream>>>(s->keysSorted, nOwn, planeCells, xLo, xHi, sl.nPlaneL, sl.nPlaneR,
                                                 hasLn ? 1 : 0, hasRn ? 1 : 0, sl.dCounts);
    AK_LAUNCH_CHECK(s, "k_plane_verify");
    const uint64_t nTot = (uint64_t)nOwn + sl.nGhostL + sl.nGhostR;
    if (nTot > (uint64_t)s->capacity) { s->err = "slab: particle capacity exceeded by ghosts (raise capacity_factor)"; return AKUA_ERR_ALLOC; }
    if ((rc = slabExchangePlanes(s, s->xs))) return rc;
    const uint32_t nGhost = sl.nGhostL + sl.nGhostR;
    if (nGhost) {
        float3 g0 = make_float3(0, 0, 0);
        k_predict_key<KEY_LINEAR><<<gridFor(nGhost), kBlock, 0, s->stream>>>(nullptr, nullptr, s->xs + nOwn, s->keysSorted + nOwn, nGhost, 0.0f, g0, G, 0);
        AK_LAUNCH_CHECK(s, "k_predict_key(ghosts)");
        slab::k_ranges<<<gridFor(nGhost), kBlock, 0, s->stream>>>(s->keysSorted, nOwn, (uint32_t)nTot, s->cellRange);
        AK_LAUNCH_CHECK(s, "k_ranges(ghosts)");
    }

    // ---- 5. neighbour lists of the owned particles (candidates include the ghost planes) ----
    mark(s, PH_LISTS);
    if (nOwn) {
        k_build_neighbours<KEY_LINEAR><<<gridFor(nOwn), kBlock, 0, s->stream>>>(s->xs, s->keysSorted, s->bucketStart, s->cellRange, nOwn,
            s->nbrStride, (uint32_t)s->cfg.maxNeighbours, s->nbrList, s->nbrCount, s->grid, s->cfg.smoothRadius);
        AK_LAUNCH_CHECK(s, "k_build_neighbours");
    }

    sl.pending = nullptr;
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
    if (akua_slab_partition(hist.data(), gx, R, bounds.data()) != AKUA_OK) { s->err = "rebalance: grid has fewer x planes than ranks"; return AKUA_ERR_INVALID; }
    for (int r = 0; r < R; r++) old[r] = (int32_t)sl.hHist[gx + r];
    old[0] = 0; old[R] = gx;
    // clamp: boundary r (between ranks r-1 and r) stays strictly inside (old[r-1], old[r+1]) and moves at most
    // migCap/2 particles
    const int64_t maxMove = (int64_t)sl.migCap / 2;
    for (int r = 1; r < R; r++) {
        // stay inside the two old slabs and keep every slab at least two planes wide
        int b = std::min(std::max(bounds[r], std::max(old[r - 1] + 1, bounds[r - 1] + 2)), old[r + 1] - 2);
        int64_t moved = 0;
        if (b > old[r]) { int x = old[r]; while (x < b && moved + hist[x] <= maxMove) { moved += hist[x]; x++; } b = x; }
        else if (b < old[r]) { int x = old[r]; while (x > b && moved + hist[x - 1] <= maxMove) { moved += hist[x - 1]; x--; } b = x; }
        if (b < bounds[r - 1] + 2) b = std::min(bounds[r - 1] + 2, old[r + 1] - 2);
        bounds[r] = b;
    }
    // monotonic by construction (each stays within its old neighbours' interval); take this rank's new interval
    sl.xLoAbs = G.gridMin.x + bounds[sl.rank];
    sl.xHiAbs = G.gridMin.x + bounds[sl.rank + 1];
    sl.rebalances++;
    return AKUA_OK;
}
