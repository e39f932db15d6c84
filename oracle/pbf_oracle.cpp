// oracle/pbf_oracle.cpp — TEST INFRASTRUCTURE, not product code.
//
// Host-C++ restatement ("port") of AkuaEngine's PBF simulation step. Only tests/, __graft_entry__.smoke() and
// bench.py's cpu_baseline / --impl reference legs may load this library; the product path (akuaengine_b200/) never does.
//
// The reference has no CPU path: its solver is 13 CUDA kernels + one thrust::sort working in place on an array of
// 108-byte `Particle` structs. This file restates each of them as a loop over particles (all kernels are Jacobi, so a
// loop per kernel is an exact restatement of the data flow), in the reference's order of operations, citing the
// reference file:line each function follows. All citations are relative to /root/reference.
//
// Pinning: the reference holds NO tests / golden vectors for this path. This port is pinned
// instead against outputs of the reference's own kernels rebuilt headless (oracle/_ref/libakua_ref.so, run on a B200
// via gpurun): tests/golden/*.npz were produced by tests/golden/make_golden.py from that library, and
// tests/test_oracle_golden.py checks this port against them on CPU (integer structures bit-exact, floats to 1e-5 rel).
//
// Bit-faithfulness notes (verified against the reference's sm_100a SASS):
//   * nvcc contracts K1 to v* = fma(g, dt, v), x* = fma(v*, dt, x)            -> fmaf here.
//   * cell coordinate = floorf(x* / cellSize) with a true IEEE division        -> plain '/' here.
//   * K4 distance     = fma(dz,dz, fma(dx,dx, dy*dy)), compared  d2 < h*h      -> dot3() here.
//   This file is compiled with -ffp-contract=off so only the explicit fmaf calls fuse.
//   libdevice powf / sqrtf differ from glibc's in the last ulp, so float fields agree to ~1e-6 relative, not bitwise.
// Documented deviation: XSPH (K13) is evaluated as a Jacobi sweep (reads the pre-sweep velocities). The reference
// updates `velocity` in place while neighbouring threads read it (src/CUDA/IntegrationCUDA.cu:187,194) — a data race
// whose outcome depends on scheduling; Jacobi is one of its legal outcomes.
#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <vector>
#ifdef _OPENMP
#include <omp.h>
#endif

namespace {

struct V3 { float x, y, z; };
inline V3 operator+(V3 a, V3 b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
inline V3 operator-(V3 a, V3 b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
inline V3 operator*(V3 a, float s) { return {a.x * s, a.y * s, a.z * s}; }
inline V3 operator*(float s, V3 a) { return {s * a.x, s * a.y, s * a.z}; }
inline V3 operator/(V3 a, float s) { return {a.x / s, a.y / s, a.z / s}; }  // glm component-wise IEEE divide
inline V3 operator-(V3 a) { return {-a.x, -a.y, -a.z}; }
// glm::dot for vec3 is tmp = a*b; tmp.x + tmp.y + tmp.z (include/glm/detail/func_geometric.inl:48-55);
// nvcc contracts it to fma(z,z', fma(x,x', y*y')) (seen in kernel_find_neighbours SASS).
inline float dot3(V3 a, V3 b) { return fmaf(a.z, b.z, fmaf(a.x, b.x, a.y * b.y)); }
inline float length3(V3 a) { return sqrtf(dot3(a, a)); }
// include/AkuaEngine/CUDA/MathUtilsCUDA.h:12-18
inline V3 cross3(V3 a, V3 b) { return {a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x}; }

// include/AkuaEngine/Simulation/Particle.h:8-31 — 108 bytes, alignment 4.
struct Particle {
    V3 position, velocity, new_position, new_velocity, position_delta, vorticity;
    float mass, density, lambda;
    uint32_t hash;
    float color[4];
    float size;
};
static_assert(sizeof(Particle) == 108, "Particle layout must match the reference");

// include/AkuaEngine/CUDA/SmoothingKernelsCUDA.h:16-21
inline float poly6(float d2, float h) {
    float h2 = h * h;
    if (d2 > h2) return 0.0f;
    return 315.0f / (64.0f * 3.14f * powf(h, 9.0f)) * powf(h2 - d2, 3.0f);
}
// include/AkuaEngine/CUDA/SmoothingKernelsCUDA.h:23-28
inline V3 gradSpiky(V3 r, float h) {
    float rm = length3(r);
    if (rm > h || rm < 1e-5f) return {0.0f, 0.0f, 0.0f};
    return (-45.0f / (3.14f * powf(h, 6.0f)) * powf(h - rm, 2.0f) * r) / rm;
}

struct IV3 { int x, y, z; };
// src/CUDA/NeighbourSearchCUDA.cu:15-21
inline IV3 discretize(const Particle& p, float cellSize) {
    return {(int)floorf(p.new_position.x / cellSize), (int)floorf(p.new_position.y / cellSize),
            (int)floorf(p.new_position.z / cellSize)};
}
// src/CUDA/NeighbourSearchCUDA.cu:23-27 — signed wrapping multiply, reinterpret as u32, xor, unsigned modulo.
inline uint32_t cellHash(IV3 c, int tableSize) {
    uint32_t a = (uint32_t)c.x * 73856093u, b = (uint32_t)c.y * 19349663u, d = (uint32_t)c.z * 83492791u;
    return (a ^ b ^ d) % (uint32_t)tableSize;
}

struct Oracle {
    int n;
    // include/AkuaEngine/Simulation/PBFConfig.h:10-29
    float restDensity, particleSpacing, smoothRadius, cellSize, relaxation, vorticityEpsilon, viscosity;
    int maxNeighbours, solverIterations;
    V3 gravity;
    float corrK, corrN, corrDeltaQ;
    // src/Simulation/PBFSolver.cpp:13-20
    int tableSize;
    std::vector<uint32_t> nbrArray, nbrCount;
    std::vector<uint32_t> table;       // hashToFirstParticleIndex (src/CUDA/NeighbourSearchCUDA.cu:157)
    std::vector<Particle> scratch;     // merge buffer for the stable sort
    std::vector<V3> vtmp;              // Jacobi buffer for XSPH
};

// ---- K1: src/CUDA/IntegrationCUDA.cu:27-36 ----
void predict(Oracle& o, Particle* P, float dt) {
    const V3 g = o.gravity;
#pragma omp parallel for schedule(static)
    for (int i = 0; i < o.n; i++) {
        Particle& p = P[i];
        p.new_velocity = {fmaf(g.x, dt, p.velocity.x), fmaf(g.y, dt, p.velocity.y), fmaf(g.z, dt, p.velocity.z)};
        p.new_position = {fmaf(p.new_velocity.x, dt, p.position.x), fmaf(p.new_velocity.y, dt, p.position.y),
                          fmaf(p.new_velocity.z, dt, p.position.z)};
    }
}

// parallel stable merge sort by hash — stands in for thrust::sort (src/CUDA/NeighbourSearchCUDA.cu:167-170), which is
// a stable merge sort in CCCL 2.8.2.
void stableSortByHash(Oracle& o, Particle* P) {
    const int n = o.n;
    auto cmp = [](const Particle& a, const Particle& b) { return a.hash < b.hash; };
    int nt = 1;
#ifdef _OPENMP
    nt = omp_get_max_threads();
#endif
    if (nt <= 1 || n < 4096) { std::stable_sort(P, P + n, cmp); return; }
    int chunks = 1; while (chunks < nt) chunks <<= 1;
    std::vector<int> b(chunks + 1);
    for (int c = 0; c <= chunks; c++) b[c] = (int)((int64_t)n * c / chunks);
#pragma omp parallel for schedule(dynamic, 1)
    for (int c = 0; c < chunks; c++) std::stable_sort(P + b[c], P + b[c + 1], cmp);
    o.scratch.resize(n);
    Particle* src = P; Particle* dst = o.scratch.data();
    for (int w = 1; w < chunks; w <<= 1) {
#pragma omp parallel for schedule(dynamic, 1)
        for (int c = 0; c < chunks; c += 2 * w)
            std::merge(src + b[c], src + b[c + w], src + b[c + w], src + b[c + 2 * w], dst + b[c], cmp);  // stable
        std::swap(src, dst);
    }
    if (src != P) std::memcpy(P, src, (size_t)n * sizeof(Particle));
}

// ---- findParticleNeighboursCUDA: src/CUDA/NeighbourSearchCUDA.cu:134-187 ----
void neighbours(Oracle& o, Particle* P) {
    const int n = o.n, tableSize = o.tableSize, maxN = o.maxNeighbours;
    // :157 table filled with UINT32_MAX every call
    o.table.resize((size_t)tableSize);
#pragma omp parallel for schedule(static)
    for (int64_t t = 0; t < (int64_t)tableSize; t++) o.table[t] = UINT32_MAX;
    // K2 :36-44 — note the wrapper passes smoothRadius as the cell size here (:163)
#pragma omp parallel for schedule(static)
    for (int i = 0; i < n; i++) P[i].hash = cellHash(discretize(P[i], o.smoothRadius), tableSize);
    // :167-170
    stableSortByHash(o, P);
    // K3 :52-65
#pragma omp parallel for schedule(static)
    for (int i = 0; i < n; i++) {
        uint32_t h = P[i].hash;
        if (i == 0 || h != P[i - 1].hash) o.table[h] = (uint32_t)i;
    }
    // K4 :72-130 — passed cellSize (:177-178)
    const float h2 = o.smoothRadius * o.smoothRadius;
#pragma omp parallel for schedule(dynamic, 256)
    for (int i = 0; i < n; i++) {
        const Particle& pi = P[i];
        IV3 cell = discretize(pi, o.cellSize);
        int count = 0;
        size_t start = (size_t)i * maxN;
        for (int dx = -1; dx <= 1; dx++)
            for (int dy = -1; dy <= 1; dy++)
                for (int dz = -1; dz <= 1; dz++) {
                    uint32_t hash = cellHash({cell.x + dx, cell.y + dy, cell.z + dz}, tableSize);
                    int cand = (int)o.table[hash];
                    if ((uint32_t)cand == UINT32_MAX) continue;
                    while (cand < n && count < maxN) {
                        if (cand == i) { cand++; continue; }
                        const Particle& pj = P[cand];
                        if (pj.hash != hash) break;
                        V3 sep = pi.new_position - pj.new_position;
                        if (dot3(sep, sep) < h2) { o.nbrArray[start + count] = (uint32_t)cand; count++; }
                        cand++;
                    }
                }
        o.nbrCount[i] = (uint32_t)count;
    }
}

// ---- runConstraintSolverCUDA: src/CUDA/ConstraintSolverCUDA.cu:173-222 ----
void solve(Oracle& o, Particle* P, int iterations, V3 bmin, V3 bmax) {
    const int n = o.n, maxN = o.maxNeighbours;
    const float h = o.smoothRadius;
    const float invRho0 = 1.0f / o.restDensity;  // :201
    while (iterations-- > 0) {
        // K5 :16-42
#pragma omp parallel for schedule(dynamic, 256)
        for (int i = 0; i < n; i++) {
            Particle& pi = P[i];
            float density = pi.mass * poly6(0.0f, h);
            size_t start = (size_t)i * maxN;
            for (uint32_t k = 0; k < o.nbrCount[i]; k++) {
                const Particle& pj = P[o.nbrArray[start + k]];
                V3 sep = pi.new_position - pj.new_position;
                density = fmaf(pj.mass, poly6(dot3(sep, sep), h), density);
            }
            pi.density = density;
        }
        // K6 :51-97
#pragma omp parallel for schedule(dynamic, 256)
        for (int i = 0; i < n; i++) {
            Particle& pi = P[i];
            V3 gradI = {0, 0, 0};
            size_t start = (size_t)i * maxN;
            for (uint32_t k = 0; k < o.nbrCount[i]; k++) {
                const Particle& pj = P[o.nbrArray[start + k]];
                gradI = gradI + pj.mass * gradSpiky(pi.new_position - pj.new_position, h);
            }
            gradI = gradI * invRho0;
            float sum = 0.0f;
            for (uint32_t k = 0; k < o.nbrCount[i]; k++) {
                const Particle& pj = P[o.nbrArray[start + k]];
                V3 gradJ = (-invRho0 * pj.mass) * gradSpiky(pi.new_position - pj.new_position, h);
                sum += dot3(gradJ, gradJ);
            }
            float C = pi.density * invRho0 - 1.0f;
            pi.lambda = -C / (sum + dot3(gradI, gradI) + o.relaxation);
        }
        // K7 :99-130 — the `enabled` flag is never read; artificial pressure is always applied (:123-126)
#pragma omp parallel for schedule(dynamic, 256)
        for (int i = 0; i < n; i++) {
            Particle& pi = P[i];
            V3 dp = {0, 0, 0};
            size_t start = (size_t)i * maxN;
            for (uint32_t k = 0; k < o.nbrCount[i]; k++) {
                const Particle& pj = P[o.nbrArray[start + k]];
                V3 sep = pi.new_position - pj.new_position;
                float d2 = dot3(sep, sep);
                float dq2 = o.corrDeltaQ * o.corrDeltaQ;
                float corr = -o.corrK * powf(poly6(d2, h) / poly6(dq2, h), o.corrN);
                dp = dp + ((pi.lambda + pj.lambda + corr) * pj.mass) * gradSpiky(sep, h);
            }
            pi.position_delta = dp * invRho0;
        }
        // K8 :159-169 with handle_particle_collision :136-157
#pragma omp parallel for schedule(static)
        for (int i = 0; i < n; i++) {
            Particle& p = P[i];
            p.new_position = p.new_position + p.position_delta;
            const float minDist = 0.025f, stiffness = 0.5f;
            V3 c = {0, 0, 0};
            if (p.new_position.x < bmin.x + minDist) c.x += stiffness * (bmin.x + minDist - p.new_position.x);
            if (p.new_position.x > bmax.x - minDist) c.x += stiffness * (bmax.x - minDist - p.new_position.x);
            if (p.new_position.y < bmin.y + minDist) c.y += stiffness * (bmin.y + minDist - p.new_position.y);
            if (p.new_position.y > bmax.y - minDist) c.y += stiffness * (bmax.y - minDist - p.new_position.y);
            if (p.new_position.z < bmin.z + minDist) c.z += stiffness * (bmin.z + minDist - p.new_position.z);
            if (p.new_position.z > bmax.z - minDist) c.z += stiffness * (bmax.z - minDist - p.new_position.z);
            p.new_position = p.new_position + c;
        }
    }
}

// ---- K9: src/CUDA/IntegrationCUDA.cu:38-49 ----
void update(Oracle& o, Particle* P, float dt) {
#pragma omp parallel for schedule(static)
    for (int i = 0; i < o.n; i++) {
        Particle& p = P[i];
        p.new_velocity = (p.new_position - p.position) / dt;
        p.position = p.new_position;
        p.velocity = p.new_velocity;
    }
}

// ---- K10: src/CUDA/IntegrationCUDA.cu:75-102 with resolve_collision :51-73 ----
inline void resolveCollision(V3& position, V3& velocity, V3 planePoint, V3 normal, float minDist, float restitution,
                             float friction) {
    float distance = dot3(position - planePoint, normal);
    float approaching = dot3(velocity, normal);
    if (distance < minDist) {
        V3 vn = approaching * normal;
        V3 vt = velocity - vn;
        if (approaching < 0.0f) velocity = -restitution * vn + (1.0f - friction) * vt;
        else if (fabsf(approaching) < 1e-5f) velocity = (1.0f - friction) * vt;
    }
}
void damping(Oracle& o, Particle* P, V3 bmin, V3 bmax, float restitution, float friction) {
#pragma omp parallel for schedule(static)
    for (int i = 0; i < o.n; i++) {
        Particle& p = P[i];
        const float minDist = 0.025f;
        const V3 pts[6] = {{bmin.x, 0, 0}, {bmax.x, 0, 0}, {0, bmin.y, 0}, {0, bmax.y, 0}, {0, 0, bmin.z}, {0, 0, bmax.z}};
        const V3 nrm[6] = {{1, 0, 0}, {-1, 0, 0}, {0, 1, 0}, {0, -1, 0}, {0, 0, 1}, {0, 0, -1}};
        for (int k = 0; k < 6; k++) resolveCollision(p.position, p.velocity, pts[k], nrm[k], minDist, restitution, friction);
    }
}

// ---- applyVorticityAndViscosityCUDA: src/CUDA/IntegrationCUDA.cu:253-286 ----
void vorticityViscosity(Oracle& o, Particle* P, float dt) {
    const int n = o.n, maxN = o.maxNeighbours;
    const float h = o.smoothRadius;
    // K11 :104-128
#pragma omp parallel for schedule(dynamic, 256)
    for (int i = 0; i < n; i++) {
        Particle& pi = P[i];
        V3 w = {0, 0, 0};
        size_t start = (size_t)i * maxN;
        for (uint32_t k = 0; k < o.nbrCount[i]; k++) {
            const Particle& pj = P[o.nbrArray[start + k]];
            V3 vij = pj.velocity - pi.velocity;
            V3 sep = pi.new_position - pj.new_position;
            w = w + (-pj.mass) * cross3(vij, gradSpiky(sep, h));
        }
        pi.vorticity = w;
    }
    // K12 :130-165 (reads vorticity, writes own velocity: race-free)
#pragma omp parallel for schedule(dynamic, 256)
    for (int i = 0; i < n; i++) {
        Particle& pi = P[i];
        float invDensity = 1 / pi.density;
        V3 eta = {0, 0, 0};
        size_t start = (size_t)i * maxN;
        for (uint32_t k = 0; k < o.nbrCount[i]; k++) {
            const Particle& pj = P[o.nbrArray[start + k]];
            float diff = length3(pi.vorticity) - length3(pj.vorticity);
            V3 sep = pi.new_position - pj.new_position;
            eta = eta + (pj.mass * diff) * gradSpiky(sep, h);
        }
        eta = eta * invDensity;
        float etaLen = length3(eta);
        if (etaLen < 1e-5f) continue;
        V3 N = eta / etaLen;
        V3 force = o.vorticityEpsilon * cross3(N, pi.vorticity);
        pi.velocity = pi.velocity + dt * force;
    }
    // K13 :167-195, as a Jacobi sweep (see header)
    o.vtmp.resize(n);
#pragma omp parallel for schedule(dynamic, 256)
    for (int i = 0; i < n; i++) {
        const Particle& pi = P[i];
        V3 dv = {0, 0, 0};
        size_t start = (size_t)i * maxN;
        for (uint32_t k = 0; k < o.nbrCount[i]; k++) {
            const Particle& pj = P[o.nbrArray[start + k]];
            V3 vij = pj.velocity - pi.velocity;
            V3 sep = pi.new_position - pj.new_position;
            float w = poly6(dot3(sep, sep), h);
            dv = dv + ((pj.mass / pj.density) * vij) * w;
        }
        o.vtmp[i] = pi.velocity + o.viscosity * dv;
    }
#pragma omp parallel for schedule(static)
    for (int i = 0; i < n; i++) P[i].velocity = o.vtmp[i];
}

inline V3 v3(const float* p) { return {p[0], p[1], p[2]}; }

}  // namespace

// ------------------------------------------------------------------ C ABI (mirrors oracle/ref_harness.cu's akref_*)
// Parameter block layout: see ref_harness.cu (15 floats).
extern "C" {

void* pbfo_create(int n, const float* p) {
    auto* o = new Oracle();
    o->n = n;
    o->restDensity = p[0]; o->particleSpacing = p[1]; o->smoothRadius = p[2]; o->cellSize = p[3];
    o->relaxation = p[4]; o->vorticityEpsilon = p[5]; o->viscosity = p[6];
    o->maxNeighbours = (int)p[7]; o->solverIterations = (int)p[8];
    o->gravity = {p[9], p[10], p[11]};
    o->corrK = p[12]; o->corrN = p[13]; o->corrDeltaQ = p[14];
    o->tableSize = o->maxNeighbours * n;  // src/Simulation/PBFSolver.cpp:15 (int; overflows at n >= 2^24, as in the reference)
    o->nbrArray.assign((size_t)n * o->maxNeighbours, 0u);
    o->nbrCount.assign((size_t)n, 0u);
    return o;
}
void pbfo_destroy(void* h) { delete static_cast<Oracle*>(h); }
int pbfo_set_gravity(void* h, const float* g) { static_cast<Oracle*>(h)->gravity = v3(g); return 0; }
int pbfo_threads() {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

int pbfo_phase_predict(void* h, void* particles, float dt) {
    predict(*static_cast<Oracle*>(h), static_cast<Particle*>(particles), dt); return 0;
}
int pbfo_phase_neighbours(void* h, void* particles) {
    neighbours(*static_cast<Oracle*>(h), static_cast<Particle*>(particles)); return 0;
}
int pbfo_phase_solve(void* h, void* particles, int iterations, const float* bmin, const float* bmax) {
    solve(*static_cast<Oracle*>(h), static_cast<Particle*>(particles), iterations, v3(bmin), v3(bmax)); return 0;
}
int pbfo_phase_update(void* h, void* particles, float dt) {
    update(*static_cast<Oracle*>(h), static_cast<Particle*>(particles), dt); return 0;
}
int pbfo_phase_damping(void* h, void* particles, const float* bmin, const float* bmax) {
    damping(*static_cast<Oracle*>(h), static_cast<Particle*>(particles), v3(bmin), v3(bmax), 0.0f, 0.95f); return 0;
}
int pbfo_phase_vorticity_viscosity(void* h, void* particles, float dt) {
    vorticityViscosity(*static_cast<Oracle*>(h), static_cast<Particle*>(particles), dt); return 0;
}
// PBFSolver::step — src/Simulation/PBFSolver.cpp:22-78
int pbfo_step(void* h, void* particles, float dt, const float* bmin, const float* bmax) {
    Oracle& o = *static_cast<Oracle*>(h);
    Particle* P = static_cast<Particle*>(particles);
    predict(o, P, dt);                                  // :30
    neighbours(o, P);                                   // :33
    solve(o, P, o.solverIterations, v3(bmin), v3(bmax)); // :45
    update(o, P, dt);                                   // :61
    damping(o, P, v3(bmin), v3(bmax), 0.0f, 0.95f);     // :64
    vorticityViscosity(o, P, dt);                       // :67
    return 0;
}
int pbfo_get_neighbours(void* h, uint32_t* arr, uint32_t* cnt) {
    Oracle& o = *static_cast<Oracle*>(h);
    if (arr) std::memcpy(arr, o.nbrArray.data(), o.nbrArray.size() * sizeof(uint32_t));
    if (cnt) std::memcpy(cnt, o.nbrCount.data(), o.nbrCount.size() * sizeof(uint32_t));
    return 0;
}
int pbfo_max_neighbours(void* h) { return static_cast<Oracle*>(h)->maxNeighbours; }

}  // extern "C"
