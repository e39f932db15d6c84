// tests/emu/emu_harness.cpp — TEST INFRASTRUCTURE: C entry points (ctypes, tests/test_emu_kernels.py) that run the kernels
// of akuaengine_b200/csrc/ on the CPU through the SIMT emulator of tests/emu/cuda_runtime.h. Built by g++ with
// -DAKUA_HOST_EMU -I tests/emu; never part of libakua_pbf.so and never loaded by the product package.
//
// emu_step() sequences the kernels exactly as stepEager() in akuaengine_b200/csrc/pbf_solver.cu does (same kernels, same
// arguments, same buffer swaps) on plain host arrays, so a whole PBF step of the CUDA source can be compared with the CPU
// port of the reference and with the golden fixtures without a GPU.
#include "../../akuaengine_b200/csrc/pbf_kernels.cuh"
#include "../../akuaengine_b200/csrc/list_build.cuh"
#include "../../akuaengine_b200/csrc/pbf_params.h"
#include "../../akuaengine_b200/csrc/radix_sort.cuh"
#include "../../akuaengine_b200/csrc/slab_kernels.cuh"

#include <algorithm>
#include <cmath>
#include <vector>

using namespace akua;

namespace {

template <typename... KArgs, typename... Args>
void run(void (*kernel)(KArgs...), uint32_t grid, uint32_t block, Args... args) {
    if (grid == 0) return;
    emu::launch(dim3(grid), dim3(block), [&]() { kernel(KArgs(args)...); });
}
inline uint32_t gridFor(uint64_t n, uint32_t block = 256) { return (uint32_t)((n + block - 1) / block); }
int bitsFor(uint64_t maxKey) { int b = 1; while (b < 32 && (maxKey >> b) != 0) b++; return b; }

}  // namespace

extern "C" {

// Stable LSD radix sort of (key, index) pairs: the three kernels of radix_sort.cuh driven by rsort::sort_pairs itself.
// mode: 0 = three kernels per pass, 1 = one-sweep (decoupled look-back); items: keys per thread of the one-sweep tiles (0 = auto)
int emu_sort_pairs(const uint32_t* keysIn, uint32_t n, int keyBits, uint32_t* keysOut, uint32_t* valsOut, int mode, int items) {
    std::vector<uint32_t> keyA(n), keyB(n), valA(n), valB(n);
    rsort::Workspace ws;
    ws.mode = mode; ws.items = items;
    ws.maxTiles = rsort::max_tiles_for_capacity(n);
    std::vector<uint32_t> tileHist(rsort::tile_hist_words(ws.maxTiles), 0xdeadbeefu), binTotal(rsort::kCtrlWords, 0xdeadbeefu);
    ws.tileHist = tileHist.data();
    ws.binTotal = binTotal.data();
    uint32_t *ko = nullptr, *vo = nullptr;
    const int launches = rsort::sort_pairs(keysIn, keyA.data(), valA.data(), keyB.data(), valB.data(), n, keyBits, ws, nullptr, &ko, &vo, true);
    std::copy(ko, ko + n, keysOut);
    std::copy(vo, vo + n, valsOut);
    return launches;
}


// ---- a whole solver on host arrays: the sequence of stepEager() / phaseSolve() / phasePost() in pbf_solver.cu ----
struct EmuSolver {
    uint32_t n = 0;
    akua_pbf_config cfg{};
    akua_corr_params corr{};
    int keyMode = AKUA_KEY_LINEAR_CELL, fastMath = 1, pack = 1, listBuild = 0;   // pack: bit 0 packed arrays, bit 1 32-byte records
    std::vector<PosVel> pv;
    std::vector<float4> pos, posAlt, vel, velAlt, xs, xsAlt, omega, dpos, color, xl, xw;
    std::vector<float> density, lambda, omegaLen, size;
    std::vector<uint32_t> id, idAlt, keysUnsorted, keyA, keyB, valA, valB, bucketStart, nbrList, nbrCount, tileHist, binTotal;
    std::vector<uint2> cellRange;
    uint32_t *keysSorted = nullptr, *perm = nullptr;
    float4 *pPos, *pPosAlt, *pVel, *pVelAlt, *pXs, *pXsAlt;
    uint32_t *pId, *pIdAlt;
    GridParams grid{};
    int keyBits = 1;
    uint32_t nbrStride = 0;
    bool bucketsDirty = false, massUniform = false;
    float uniformMass = 0.f;
    rsort::Workspace ws;
    long launches = 0;
};

void* emu_create(uint32_t n, const akua_pbf_config* cfg, const akua_corr_params* corr, int keyMode, int fastMath, int pack,
                 int listBuild) {
    EmuSolver* s = new EmuSolver();
    s->n = n; s->cfg = *cfg; s->corr = *corr; s->keyMode = keyMode; s->fastMath = fastMath; s->pack = pack; s->listBuild = listBuild;
    const size_t cap = n ? n : 1;
    for (auto* v : {&s->pos, &s->posAlt, &s->vel, &s->velAlt, &s->xs, &s->xsAlt, &s->omega, &s->dpos, &s->color, &s->xl, &s->xw})
        v->assign(cap, make_float4(0, 0, 0, 0));
    for (auto* v : {&s->density, &s->lambda, &s->omegaLen, &s->size}) v->assign(cap, 0.f);
    for (auto* v : {&s->id, &s->idAlt, &s->keysUnsorted, &s->keyA, &s->keyB, &s->valA, &s->valB, &s->nbrCount}) v->assign(cap, 0u);
    s->pPos = s->pos.data(); s->pPosAlt = s->posAlt.data(); s->pVel = s->vel.data(); s->pVelAlt = s->velAlt.data();
    s->pXs = s->xs.data(); s->pXsAlt = s->xsAlt.data(); s->pId = s->id.data(); s->pIdAlt = s->idAlt.data();
    s->keysSorted = s->keyA.data(); s->perm = s->valA.data();
    s->nbrStride = (uint32_t)((cap + 31) / 32 * 32);
    s->nbrList.assign((size_t)s->nbrStride * (size_t)((cfg->maxNeighbours + 3) / 4 * 4), 0u);
    s->ws.maxTiles = rsort::max_tiles_for_capacity(cap);
    s->tileHist.assign(rsort::tile_hist_words(s->ws.maxTiles), 0u); s->binTotal.assign(rsort::kCtrlWords, 0u);
    s->ws.tileHist = s->tileHist.data(); s->ws.binTotal = s->binTotal.data();
    s->grid.cellSize = cfg->smoothRadius;
    s->grid.lookupCellSize = cfg->spatialHashCellSize;
    if (keyMode == AKUA_KEY_REFERENCE_HASH) {
        const int64_t ts = n ? (int64_t)cfg->maxNeighbours * n : 1;   // PBFSolver.cpp:15
        s->grid.tableSize = (uint32_t)ts;
        s->keyBits = bits_for_key((uint64_t)ts - 1);
        s->bucketStart.assign((size_t)ts, 0xffffffffu);
    }
    return s;
}
void emu_destroy(void* h) { delete static_cast<EmuSolver*>(h); }

int emu_upload_aos108(void* h, const void* src) {
    EmuSolver* s = static_cast<EmuSolver*>(h);
    const uint32_t n = s->n;
    if (!n) return 0;
    run(k_unpack_aos, gridFor(n), 256u, (const uint32_t*)src, n, s->pPos, s->pVel, s->pXs, s->omega.data(), s->omegaLen.data(),
        s->dpos.data(), s->density.data(), s->lambda.data(), s->keysSorted, s->color.data(), s->size.data(), s->pId, 0u);
    uint32_t range[2] = {0xffffffffu, 0u};   // massRangeAsync / massRangeFinish
    run(k_mass_range, std::min<uint32_t>(gridFor(n), 148 * 8), 256u, (const float4*)s->pPos, n, range);
    s->massUniform = range[0] == range[1];
    const uint32_t bits = (range[0] & 0x80000000u) ? (range[0] ^ 0x80000000u) : ~range[0];
    std::memcpy(&s->uniformMass, &bits, 4);
    return 0;
}
int emu_download_aos108(void* h, void* dst) {
    EmuSolver* s = static_cast<EmuSolver*>(h);
    const uint32_t n = s->n;
    if (!n) return 0;
    run(k_pack_aos, gridFor(n), 256u, (uint32_t*)dst, n, (const float4*)s->pPos, (const float4*)s->pVel, (const float4*)s->pXs,
        (const float4*)s->omega.data(), (const float4*)s->dpos.data(), (const float*)s->density.data(), (const float*)s->lambda.data(),
        (const uint32_t*)s->keysSorted, (const float4*)s->color.data(), (const float*)s->size.data(), (const uint32_t*)s->pId,
        (const uint32_t*)nullptr);
    return 0;
}

int emu_step(void* h, float dt, int iterations, const float* bmin, const float* bmax) {
    EmuSolver* s = static_cast<EmuSolver*>(h);
    const uint32_t n = s->n;
    if (!n) return 0;
    const bool hash = s->keyMode == AKUA_KEY_REFERENCE_HASH;
    const bool pack = (s->pack & 1) && s->massUniform;
    const bool rec = (s->pack & 2) != 0;
    if (rec && s->pv.size() < n) s->pv.assign(n, PosVel{});
    PosVel* pvp = rec ? s->pv.data() : nullptr;
    if (!hash) {   // layoutGrid
        int3 gmin, gdim;
        const int64_t cells = layout_linear_grid(s->cfg.smoothRadius, bmin, bmax, &gmin, &gdim);
        if (cells < 0 || cells >= (int64_t)1 << 31) return 1;
        s->cellRange.assign((size_t)cells, make_uint2(0, 0));   // = the per-step memset
        s->grid.gridMin = gmin; s->grid.gridDim = gdim;
        s->keyBits = bits_for_key((uint64_t)cells - 1);
    }
    const float3 g = make_float3(s->cfg.gravity[0], s->cfg.gravity[1], s->cfg.gravity[2]);
    // phasePredictKey
    if (hash) run(k_predict_key<KEY_HASH>, gridFor(n), 256u, (const float4*)s->pPos, (const float4*)s->pVel, s->pXs, s->keysUnsorted.data(), n, (const uint32_t*)nullptr, dt, g, s->grid, 1);
    else      run(k_predict_key<KEY_LINEAR>, gridFor(n), 256u, (const float4*)s->pPos, (const float4*)s->pVel, s->pXs, s->keysUnsorted.data(), n, (const uint32_t*)nullptr, dt, g, s->grid, 1);
    // phaseSortReorderLists
    if (hash && s->bucketsDirty) run(k_clear_buckets, gridFor(n), 256u, (const uint32_t*)s->keysSorted, n, s->bucketStart.data());
    s->launches += rsort::sort_pairs(s->keysUnsorted.data(), s->keyA.data(), s->valA.data(), s->keyB.data(), s->valB.data(), n, s->keyBits,
                                     s->ws, nullptr, &s->keysSorted, &s->perm, true);
    if (hash) run(k_reorder_ranges<KEY_HASH>, gridFor(n), 256u, (const uint32_t*)s->keysSorted, (const uint32_t*)s->perm, n, (const uint32_t*)nullptr, (const float4*)s->pPos,
                  (const float4*)s->pVel, (const float4*)s->pXs, (const uint32_t*)s->pId, s->pPosAlt, s->pVelAlt, s->pXsAlt, s->pIdAlt,
                  s->bucketStart.data(), s->cellRange.data(), (const uint32_t*)nullptr, (uint32_t*)nullptr);
    else      run(k_reorder_ranges<KEY_LINEAR>, gridFor(n), 256u, (const uint32_t*)s->keysSorted, (const uint32_t*)s->perm, n, (const uint32_t*)nullptr, (const float4*)s->pPos,
                  (const float4*)s->pVel, (const float4*)s->pXs, (const uint32_t*)s->pId, s->pPosAlt, s->pVelAlt, s->pXsAlt, s->pIdAlt,
                  s->bucketStart.data(), s->cellRange.data(), (const uint32_t*)nullptr, (uint32_t*)nullptr);
    std::swap(s->pPos, s->pPosAlt); std::swap(s->pVel, s->pVelAlt); std::swap(s->pXs, s->pXsAlt); std::swap(s->pId, s->pIdAlt);
    s->bucketsDirty = hash;
#define EMU_BUILD(K) run(K, gridFor(n), 256u, (const float4*)s->pXs, (const uint32_t*)s->keysSorted, (const uint32_t*)s->bucketStart.data(), \
                         (const uint2*)s->cellRange.data(), n, s->nbrStride, (uint32_t)s->cfg.maxNeighbours, s->nbrList.data(), s->nbrCount.data(), \
                         s->grid, s->cfg.smoothRadius, (const uint32_t*)nullptr)
    if (hash) EMU_BUILD((k_build_neighbours<KEY_HASH, false>));
    else if (s->listBuild == 1) EMU_BUILD((k_build_neighbours_mask<4, 5, true>));
    else if (s->listBuild == 2) EMU_BUILD((k_build_neighbours_mask<8, 4, true>));
    else EMU_BUILD((k_build_neighbours<KEY_LINEAR, false>));
#undef EMU_BUILD
    // phaseSolve (single GPU: one span over all particles; the last pass B commits)
    const SphParams P = make_sph_params(s->cfg, s->corr, s->uniformMass);
    const BoxParams B = make_box_params(bmin, bmax);
    const Span all{n, 0u, 0xffffffffu, 0u, nullptr, SPAN_FIXED};
    const uint32_t sg = gridFor(n, AKUA_SWEEP_BLOCK), sb = AKUA_SWEEP_BLOCK;
    float4* xl = pack ? s->xl.data() : nullptr;
    float4* xw = pack ? s->xw.data() : nullptr;
    const PeerPush nop{};
    const HaloSync nohs{};
    bool committed = false;
    for (int it = 0; it < iterations; it++) {
        const bool fin = it == iterations - 1;
#define EMU_A(F) run(k_density_lambda<F, false>, sg, sb, (const float4*)s->pXs, (const uint32_t*)s->nbrList.data(), (const uint32_t*)s->nbrCount.data(), \
                     s->nbrStride, all, s->density.data(), s->lambda.data(), xl, P, nop, nohs)
        if (s->fastMath) EMU_A(true); else EMU_A(false);
#undef EMU_A
#define EMU_B(F, L, K, C) run(k_delta_apply<F, L, K, C, false>, sg, sb, (const float4*)s->pXs, s->pXsAlt, (const float*)s->lambda.data(), (const float4*)s->xl.data(), \
                     (const uint32_t*)s->nbrList.data(), (const uint32_t*)s->nbrCount.data(), s->nbrStride, all, P, B, s->dpos.data(), s->pPos, s->pVel, \
                     (const float*)s->density.data(), (fin ? pvp : (PosVel*)nullptr), dt, nop, nop, nohs)
#define EMU_B_C(F, L, K) do { if (P.corrNIsFour) EMU_B(F, L, K, true); else EMU_B(F, L, K, false); } while (0)
#define EMU_B_L(F, K) do { if (fin) EMU_B_C(F, true, K); else EMU_B_C(F, false, K); } while (0)
        if (pack) { if (s->fastMath) EMU_B_L(true, true); else EMU_B_L(false, true); }
        else      { if (s->fastMath) EMU_B_L(true, false); else EMU_B_L(false, false); }
#undef EMU_B_L
#undef EMU_B_C
#undef EMU_B
        std::swap(s->pXs, s->pXsAlt);
        if (fin) committed = true;
    }
    if (!committed) {   // solverIterations == 0
        run(k_update, gridFor(n), 256u, (const float4*)s->pXs, s->pPos, s->pVel, (const float*)s->density.data(), n, dt, (const uint32_t*)nullptr);
        run(k_damping, gridFor(n), 256u, (const float4*)s->pPos, s->pVel, n, B, (const uint32_t*)nullptr);
    }
    // phasePost
    if (rec && !committed) run(k_build_posvel, gridFor(n), 256u, (const float4*)s->pXs, (const float4*)s->pVel, n, pvp);   // as pbf_solver.cu does
#define EMU_V(F, R) run(k_vorticity<F, R, false>, sg, sb, (const float4*)s->pXs, (const float4*)s->pVel, (const PosVel*)pvp, (const uint32_t*)s->nbrList.data(), \
                     (const uint32_t*)s->nbrCount.data(), s->nbrStride, all, s->omega.data(), s->omegaLen.data(), xw, P, nop, nohs)
    if (rec) { if (s->fastMath) EMU_V(true, true); else EMU_V(false, true); }
    else     { if (s->fastMath) EMU_V(true, false); else EMU_V(false, false); }
#undef EMU_V
#define EMU_C(F, K) run(k_confinement<F, K, false>, sg, sb, (const float4*)s->pXs, (const float4*)s->omega.data(), (const float*)s->omegaLen.data(), \
                     (const float4*)s->xw.data(), (const float*)s->density.data(), (const uint32_t*)s->nbrList.data(), (const uint32_t*)s->nbrCount.data(), \
                     s->nbrStride, all, s->pVel, pvp, P, dt, s->cfg.vorticityEpsilon, nop, nohs)
    if (pack) { if (s->fastMath) EMU_C(true, true); else EMU_C(false, true); }
    else      { if (s->fastMath) EMU_C(true, false); else EMU_C(false, false); }
#undef EMU_C
    if (rec) run(k_xsph<true, false>, sg, sb, (const float4*)s->pXs, (const float4*)s->pVel, (const PosVel*)pvp, (const uint32_t*)s->nbrList.data(),
                 (const uint32_t*)s->nbrCount.data(), s->nbrStride, all, s->pVelAlt, P, s->cfg.viscosity, nohs);
    else     run(k_xsph<false, false>, sg, sb, (const float4*)s->pXs, (const float4*)s->pVel, (const PosVel*)nullptr, (const uint32_t*)s->nbrList.data(),
                 (const uint32_t*)s->nbrCount.data(), s->nbrStride, all, s->pVelAlt, P, s->cfg.viscosity, nohs);
    std::swap(s->pVel, s->pVelAlt);
    return 0;
}

// debug taps (which: 0 unsorted keys, 1 sorted keys, 2 permutation, 3 ids, 6 neighbour counts, 7 neighbour lists row-major)
int emu_debug_get(void* h, int which, uint32_t* dst) {
    EmuSolver* s = static_cast<EmuSolver*>(h);
    const uint32_t n = s->n;
    switch (which) {
        case 0: std::copy(s->keysUnsorted.begin(), s->keysUnsorted.begin() + n, dst); return 0;
        case 1: std::copy(s->keysSorted, s->keysSorted + n, dst); return 0;
        case 2: std::copy(s->perm, s->perm + n, dst); return 0;
        case 3: std::copy(s->pId, s->pId + n, dst); return 0;
        case 6: std::copy(s->nbrCount.begin(), s->nbrCount.begin() + n, dst); return 0;
        case 7: {
            const uint32_t maxN = (uint32_t)s->cfg.maxNeighbours;
            run(k_list_to_rowmajor, gridFor((uint64_t)n * maxN), 256u, (const uint32_t*)s->nbrList.data(), (const uint32_t*)s->nbrCount.data(),
                s->nbrStride, n, maxN, dst);
            return 0;
        }
    }
    return 1;
}


// ---- interior / boundary split of a sweep and the fused halo push (multi-GPU overlap, pbf_solver.cu: sweepSpans, slabPush) ----
// After at least one emu_step (lists exist): runs pass A once over the whole range (single-GPU kernel) and once as the x-slab
// interior + boundary launches: SLAB kernels on a deliberately SMALL grid (so the grid-stride loop runs several trips), spans
// and plane sizes resolved on the device from a dims block, the boundary launch pushing (x*, lambda) / lambda of the first
// planeL and last planeR particles into "peer" buffers and publishing the epoch. Returns 0 when density, lambda and the packed
// array are bit-identical, the pushed planes equal the corresponding slices and the epoch words were published.
int emu_span_push_check(void* h, uint32_t planeL, uint32_t planeR) {
    EmuSolver* s = static_cast<EmuSolver*>(h);
    const uint32_t n = s->n;
    if ((uint64_t)planeL + planeR >= n) return -1;
    const SphParams P = make_sph_params(s->cfg, s->corr, s->uniformMass);
    const bool pack = s->pack && s->massUniform;
    const uint32_t sb = AKUA_SWEEP_BLOCK;
    const HaloSync nohs{};
    std::vector<float> d0(n, -1.f), l0(n, -1.f), d1(n, -2.f), l1(n, -2.f);
    std::vector<float4> x0(n, make_float4(0, 0, 0, 0)), x1(n, make_float4(9, 9, 9, 9));
    run(k_density_lambda<true, false>, gridFor(n, sb), sb, (const float4*)s->pXs, (const uint32_t*)s->nbrList.data(),
        (const uint32_t*)s->nbrCount.data(), s->nbrStride, Span{n, 0u, 0xffffffffu, 0u, nullptr, SPAN_FIXED}, d0.data(), l0.data(),
        pack ? x0.data() : nullptr, P, PeerPush{}, nohs);
    std::vector<float4> peerL4(planeL + 1), peerR4(planeR + 1);
    std::vector<float> peerL1(planeL + 1), peerR1(planeR + 1);
    uint32_t dims[D_WORDS] = {};
    dims[D_NOWN] = n; dims[D_PLANE_L] = planeL; dims[D_PLANE_R] = planeR; dims[D_EPOCH] = 40;
    uint32_t flags[8] = {41, 41, 0, 0, 0, 0, 0, 0}, peerFlagL = 0, peerFlagR = 0;   // ghosts of exchange 0 "arrived"
    PeerPush pp;
    pp.dstL = pack ? (void*)peerL4.data() : (void*)peerL1.data();
    pp.dstR = pack ? (void*)peerR4.data() : (void*)peerR1.data();
    pp.dims = dims;
    HaloSync hs;
    hs.waitFlags = flags; hs.waitL = hs.waitR = 1; hs.waitIdx = 0;
    hs.signalL = &peerFlagL; hs.signalR = &peerFlagR; hs.signalIdx = 1; hs.doneCounter = flags + 3; hs.dims = dims; hs.timeoutCycles = 1000;
    auto passA = [&](int mode, const PeerPush& push, const HaloSync& sync) {
        run(k_density_lambda<true, true>, 3u, sb, (const float4*)s->pXs, (const uint32_t*)s->nbrList.data(),
            (const uint32_t*)s->nbrCount.data(), s->nbrStride, Span{0u, 0u, 0u, 0u, dims, mode}, d1.data(), l1.data(),
            pack ? x1.data() : nullptr, P, push, sync);
    };
    passA(SPAN_INTERIOR, PeerPush{}, nohs);
    passA(SPAN_BOUNDARY, pp, hs);
    if (std::memcmp(d0.data(), d1.data(), n * 4) || std::memcmp(l0.data(), l1.data(), n * 4)) return 1;
    if (pack && std::memcmp(x0.data(), x1.data(), (size_t)n * 16)) return 2;
    for (uint32_t k = 0; k < planeL; k++)
        if (pack ? std::memcmp(&peerL4[k], &x0[k], 16) != 0 : std::memcmp(&peerL1[k], &l0[k], 4) != 0) return 3;
    for (uint32_t k = 0; k < planeR; k++)
        if (pack ? std::memcmp(&peerR4[k], &x0[n - planeR + k], 16) != 0 : std::memcmp(&peerR1[k], &l0[n - planeR + k], 4) != 0) return 4;
    if (peerFlagL != 42 || peerFlagR != 42 || flags[3] != 0 || dims[D_ERROR] != 0) return 5;
    // a wait for an epoch that never comes times out into the sticky error word instead of hanging
    hs.waitIdx = 7; hs.signalIdx = -1;
    passA(SPAN_BOUNDARY, PeerPush{}, hs);
    if (!(dims[D_ERROR] & SLAB_ERR_TIMEOUT)) return 6;
    return 0;
}

// ---- x-slab migration compaction: k_mig_count -> k_mig_scan -> k_mig_pack on host arrays (slab_kernels.cuh) ----
// keys: n unsorted LINEAR_CELL keys (modified: leavers get `sentinel`). ids: particle ids carried in MigRecord::meta.x.
// counts[64] = the device dims block; idsL / idsR receive the ids of the records packed for the left / right rank in order.
int emu_migration(uint32_t* keys, uint32_t n, uint32_t planeCells, int xLo, int xHi, uint32_t sentinel, const uint32_t* ids,
                  uint32_t cap, uint32_t* counts, uint32_t* idsL, uint32_t* idsR) {
    const uint32_t blocks = std::max(1u, (n + slab::kMigTile - 1) / slab::kMigTile);
    const uint32_t tileStride = blocks + 1;
    std::vector<uint32_t> blockCnt((size_t)2 * tileStride, 0u);
    std::vector<float4> pos(n ? n : 1, make_float4(1, 2, 3, 4)), vel(pos), xs(pos), color(pos);
    std::vector<float> size(n ? n : 1, 7.0f);
    std::vector<uint32_t> slot(n ? n : 1), freeSlots(2 * (size_t)n + 2, 0u);
    for (uint32_t i = 0; i < n; i++) slot[i] = i;
    std::vector<slab::MigRecord> sendL(cap), sendR(cap);
    std::fill(counts, counts + D_WORDS, 0u);
    counts[D_N] = n; counts[D_XLO] = (uint32_t)xLo; counts[D_XHI] = (uint32_t)xHi;
    const uint32_t grid = std::max(1u, blocks / 3);   // fewer CTAs than tiles: the tile loops run
    run(slab::k_mig_count, grid, 256u, (const uint32_t*)keys, (const uint32_t*)counts, planeCells, blockCnt.data(), tileStride,
        counts + D_STAY_FIRST);
    run(slab::k_mig_scan, 1u, 1024u, blockCnt.data(), (const uint32_t*)(counts + D_N), tileStride, counts);
    run(slab::k_mig_pack, grid, 256u, keys, (const uint32_t*)counts, planeCells, (const uint32_t*)blockCnt.data(), tileStride, sentinel,
        (const float4*)pos.data(), (const float4*)vel.data(), (const float4*)xs.data(), ids, (const uint32_t*)slot.data(),
        (const float4*)color.data(), (const float*)size.data(), freeSlots.data(), sendL.data(), sendR.data(), cap);
    for (uint32_t k = 0; k < std::min(counts[0], cap); k++) idsL[k] = sendL[k].meta.x;
    for (uint32_t k = 0; k < std::min(counts[1], cap); k++) idsR[k] = sendR[k].meta.x;
    return 0;
}

// Per-x-plane histogram of sorted keys (k_plane_hist, used by akua_pbf_rebalance) and the plane-size check (k_plane_verify).
int emu_plane_hist(const uint32_t* keysSorted, const uint32_t* nbrCount, uint32_t n, uint32_t planeCells, int gx, unsigned long long* hist,
                   unsigned long long* work) {
    const uint32_t nOwn = n;
    run(slab::k_plane_hist, (uint32_t)gx, 256u, keysSorted, nbrCount, (const uint32_t*)&nOwn, planeCells, gx, 0, hist, work);
    return 0;
}
int emu_plane_verify(const uint32_t* keysSorted, uint32_t n, uint32_t planeCells, int xLo, int xHi, uint32_t predictFirst,
                     uint32_t predictLast, uint32_t* counts64) {
    counts64[D_NOWN] = n; counts64[D_PLANE_L] = predictFirst; counts64[D_PLANE_R] = predictLast;
    counts64[D_XLO] = (uint32_t)xLo; counts64[D_XHI] = (uint32_t)xHi;
    run(slab::k_plane_verify, 1u, 32u, keysSorted, planeCells, 1, 1, counts64);
    return 0;
}

}  // extern "C"
