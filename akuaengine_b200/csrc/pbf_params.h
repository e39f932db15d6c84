// pbf_params.h — host-side derivation of the kernel parameter blocks from the reference's configuration structs. Pure
// functions (no CUDA calls), shared by the host solver (pbf_solver.cu) and by the CPU emulation harness of the tests
// (tests/emu/emu_harness.cpp), so both feed the kernels exactly the same constants.
#pragma once
#include <cmath>
#include <cstdint>

#include "../../include/akua_pbf.h"
#include "pbf_kernels.cuh"

namespace akua {

inline SphParams make_sph_params(const akua_pbf_config& cfg, const akua_corr_params& corr, float uniformMass) {
    SphParams P{};
    const float h = cfg.smoothRadius;
    P.h = h;
    P.h2 = h * h;
    // Same float expressions as include/AkuaEngine/CUDA/SmoothingKernelsCUDA.h:20,27, evaluated once on the host.
    P.poly6Coef = 315.0f / (64.0f * 3.14f * powf(h, 9.0f));
    P.spikyCoef = -45.0f / (3.14f * powf(h, 6.0f));
    float t0 = P.h2 - 0.0f;
    P.selfW = P.poly6Coef * (t0 * t0 * t0);
    P.invRestDensity = 1.0f / cfg.restDensity;  // ConstraintSolverCUDA.cu:201
    P.relaxation = cfg.relaxation;
    P.corrK = corr.k;
    P.corrN = corr.n;
    float dq2 = corr.delta_q * corr.delta_q;
    float tq = P.h2 - dq2;
    float wdq = dq2 > P.h2 ? 0.0f : P.poly6Coef * (tq * tq * tq);
    P.invPoly6Dq = 1.0f / wdq;
    P.corrNIsFour = (corr.n == 4.0f) ? 1 : 0;
    P.uniformMass = uniformMass;
    return P;
}

inline BoxParams make_box_params(const float* bmin, const float* bmax) {
    BoxParams B{};
    B.bmin = make_float3(bmin[0], bmin[1], bmin[2]);
    B.bmax = make_float3(bmax[0], bmax[1], bmax[2]);
    B.collisionMinDist = 0.025f;   // ConstraintSolverCUDA.cu:137
    B.collisionStiffness = 0.5f;   // ConstraintSolverCUDA.cu:138
    B.dampingMinDist = 0.025f;     // IntegrationCUDA.cu:88
    B.restitution = 0.0f;          // PBFSolver.cpp:64
    B.oneMinusFriction = 1.0f - 0.95f;
    return B;
}

// LINEAR_CELL grid: covers the box plus a two-cell margin (the collision response is a soft clamp, so particles can sit
// slightly outside the box). Returns the number of cells, or -1 when the box is empty along an axis.
inline int64_t layout_linear_grid(float cellSize, const float* bmin, const float* bmax, int3* gridMin, int3* gridDim) {
    int lo[3], dim[3];
    int64_t cells = 1;
    for (int a = 0; a < 3; a++) {
        if (!(bmax[a] > bmin[a])) return -1;
        lo[a] = (int)std::floor(bmin[a] / cellSize) - 2;
        int hi = (int)std::floor(bmax[a] / cellSize) + 2;
        dim[a] = hi - lo[a] + 1;
        cells *= dim[a];
    }
    *gridMin = make_int3(lo[0], lo[1], lo[2]);
    *gridDim = make_int3(dim[0], dim[1], dim[2]);
    return cells;
}

inline int bits_for_key(uint64_t maxKey) {
    int b = 1;
    while (b < 32 && (maxKey >> b) != 0) b++;
    return b;
}

}  // namespace akua
