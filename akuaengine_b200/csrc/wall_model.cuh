// wall_model.cuh — OPT-IN boundary-handling upgrade (akua_pbf_options::wall_model = 1). Outside the parity
// contract: it changes results, the default (0) is the reference's behaviour bit for bit.
//
// What it replaces. The reference has no boundary model: a particle next to a wall simply misses the neighbours a wall would
// stand for, its density reads low, the constraint pulls fluid INTO the wall, and only the soft clamp of
// handle_particle_collision (src/CUDA/ConstraintSolverCUDA.cu:132-157) — marked temporary by its author: "Later we will use
// virtual particles" — pushes it back. This is that later step in its continuum limit: every box wall is backed by a half-space
// of virtual fluid at rest density, sampled infinitely finely, so that its contribution has a closed form.
//   density      rho_i += rho_w * F(d),  F(d) = integral of W_poly6 over the half-space at distance d
//                                              = (pi c / 4) [P(h) - P(d)],  P(t) = h^8 t - 4/3 h^6 t^3 + 6/5 h^4 t^5 - 4/7 h^2 t^7 + t^9 / 9
//   gradient     sum_j m_j grad W_spiky(x_i - x_j) over the half-space = -rho_w S(d) n,  S(d) = |c_s| (2 pi / 3) u^4 (h / 4 - u / 5), u = h - d
//                (n = wall normal into the fluid; c, c_s = the poly6 / spiky coefficients of SmoothingKernelsCUDA.h:20,27)
// The virtual fluid does not move: it enters the density, the particle's own constraint gradient (hence lambda) and its
// position correction (with lambda_i only), not the sum over neighbours' gradients.
//
// How it is wired WITHOUT touching the measured kernels (pbf_kernels.cuh is unchanged; the sweeps' SASS is what was profiled):
//   pass A (k_density_lambda)                        as always
//   k_wall_lambda   particles within h of a wall     re-evaluate rho, grad C, lambda with the wall terms (re-sweeps their lists)
//   pass B (k_delta_apply, never the committing one) as always; it also leaves delta-p in `dpos`
//   k_wall_dp       particles within h of a wall     x* = collide(x*_in + delta-p + wall term)
//   after the last iteration                         stand-alone commit + damping (k_update, k_damping)
// Only a few per cent of the particles are near a wall, so the two extra launches cost a few per cent of a sweep; the
// un-fused commit costs one more pass over the state. Single-GPU only (akua_pbf_create rejects the option in x-slab mode... it
// is checked in akua_pbf_comm_init).
#pragma once
#include "pbf_kernels.cuh"

namespace akua {

struct WallParams {
    float h;
    float rhoW;      // density of the virtual fluid behind the walls (rest density)
    float fCoef;     // pi c / 4
    float pH;        // P(h)
    float sCoef;     // |c_s| 2 pi / 3
    float3 bmin, bmax;
};

__host__ __device__ __forceinline__ float wall_P(float t, float h) {
    const float h2 = h * h, t2 = t * t;
    // h^8 t - 4/3 h^6 t^3 + 6/5 h^4 t^5 - 4/7 h^2 t^7 + t^9 / 9, Horner in t^2
    return t * (h2 * h2 * h2 * h2 - t2 * ((4.0f / 3.0f) * h2 * h2 * h2 - t2 * ((6.0f / 5.0f) * h2 * h2 - t2 * ((4.0f / 7.0f) * h2 - t2 * (1.0f / 9.0f)))));
}
// Density factor F(d) and gradient magnitude S(d) of one wall at distance d (clamped to [0, h]: a particle that the soft clamp
// lets sit slightly outside the box sees the wall at distance 0).
__host__ __device__ __forceinline__ void wall_terms(float d, const WallParams& W, float* F, float* S) {
    if (d >= W.h) { *F = 0.0f; *S = 0.0f; return; }
    const float dd = d > 0.0f ? d : 0.0f;
    *F = W.fCoef * (W.pH - wall_P(dd, W.h));
    const float u = W.h - dd, u2 = u * u;
    *S = W.sCoef * (u2 * u2) * (0.25f * W.h - 0.2f * u);
}
__device__ __forceinline__ bool near_wall(const float4& x, const WallParams& W) {
    return x.x - W.bmin.x < W.h || W.bmax.x - x.x < W.h || x.y - W.bmin.y < W.h || W.bmax.y - x.y < W.h ||
           x.z - W.bmin.z < W.h || W.bmax.z - x.z < W.h;
}
// Sum over the six walls: density rhoW * sum F, and the vector sum_j m_j grad W = -rhoW * sum S n.
__device__ __forceinline__ void wall_sum(const float4& x, const WallParams& W, float* rho, float* gx, float* gy, float* gz) {
    float F, S, r = 0.f, ax = 0.f, ay = 0.f, az = 0.f;
    wall_terms(x.x - W.bmin.x, W, &F, &S); r += F; ax -= S;      // lower wall: n = +e, contribution -S n
    wall_terms(W.bmax.x - x.x, W, &F, &S); r += F; ax += S;      // upper wall: n = -e
    wall_terms(x.y - W.bmin.y, W, &F, &S); r += F; ay -= S;
    wall_terms(W.bmax.y - x.y, W, &F, &S); r += F; ay += S;
    wall_terms(x.z - W.bmin.z, W, &F, &S); r += F; az -= S;
    wall_terms(W.bmax.z - x.z, W, &F, &S); r += F; az += S;
    *rho = W.rhoW * r; *gx = W.rhoW * ax; *gy = W.rhoW * ay; *gz = W.rhoW * az;
}

// After pass A: density, constraint gradient and lambda of the near-wall particles with the wall terms (K5 + K6 of the
// reference, ConstraintSolverCUDA.cu:16-97, plus the virtual half-spaces). IEEE sqrt / division throughout.
__global__ void __launch_bounds__(128) k_wall_lambda(const float4* __restrict__ xs, const uint32_t* __restrict__ list,
                                                     const uint32_t* __restrict__ cnt, uint32_t stride, uint32_t n,
                                                     float* __restrict__ density, float* __restrict__ lambda,
                                                     float4* __restrict__ xl, SphParams P, WallParams W) {
    pdl_wait();
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float4 xi = xs[i];
    if (!near_wall(xi, W)) return;
    float rho = xi.w * P.selfW;
    float gx = 0.f, gy = 0.f, gz = 0.f, sum = 0.f;
    neighbour_sweep<float4>(list, i, cnt[i], stride,
        [&](uint32_t j) { return __ldg(&xs[j]); },
        [&](const float4& xj, bool valid) {
            const float dx = xi.x - xj.x, dy = xi.y - xj.y, dz = xi.z - xj.z;
            const float d2 = dist2(dx, dy, dz);
            const float m = valid ? xj.w : 0.0f;
            rho = fmaf(m, poly6(d2, P), rho);
            const float s = spiky_scale<false>(d2, P);
            const float ax = s * dx, ay = s * dy, az = s * dz;
            gx = fmaf(m, ax, gx); gy = fmaf(m, ay, gy); gz = fmaf(m, az, gz);
            const float q = -P.invRestDensity * m;
            const float bx = q * ax, by = q * ay, bz = q * az;
            sum += fmaf(bz, bz, fmaf(bx, bx, by * by));
        });
    float wr, wx, wy, wz;
    wall_sum(xi, W, &wr, &wx, &wy, &wz);
    rho += wr; gx += wx; gy += wy; gz += wz;
    gx *= P.invRestDensity; gy *= P.invRestDensity; gz *= P.invRestDensity;
    const float C = rho * P.invRestDensity - 1.0f;
    const float lam = -C / (sum + fmaf(gz, gz, fmaf(gx, gx, gy * gy)) + P.relaxation);
    density[i] = rho;
    lambda[i] = lam;
    if (xl) xl[i] = make_float4(xi.x, xi.y, xi.z, lam);
}

// After pass B (non-committing): the near-wall particles' corrected position with the wall's share of delta-p,
// x* = collide(x*_in + delta-p + (1 / rho0) lambda_i sum_walls(-rhoW S n)) — K7 + K8 (ConstraintSolverCUDA.cu:99-169).
__global__ void __launch_bounds__(128) k_wall_dp(const float4* __restrict__ xsIn, float4* __restrict__ xsOut,
                                                 const float* __restrict__ lambda, float4* __restrict__ dpos, uint32_t n,
                                                 SphParams P, BoxParams B, WallParams W) {
    pdl_wait();
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float4 xi = xsIn[i];
    if (!near_wall(xi, W)) return;
    float wr, wx, wy, wz;
    wall_sum(xi, W, &wr, &wx, &wy, &wz);
    const float li = lambda[i] * P.invRestDensity;
    const float4 d = dpos[i];
    const float px = d.x + li * wx, py = d.y + li * wy, pz = d.z + li * wz;
    dpos[i] = make_float4(px, py, pz, 0.f);
    xsOut[i] = make_float4(collide_axis(xi.x + px, B.bmin.x, B.bmax.x, B), collide_axis(xi.y + py, B.bmin.y, B.bmax.y, B),
                           collide_axis(xi.z + pz, B.bmin.z, B.bmax.z, B), xi.w);
}

}  // namespace akua
