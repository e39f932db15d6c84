// oracle/ref_harness.cu — TEST INFRASTRUCTURE, not product code.
//
// Headless harness around the UNMODIFIED reference solver. The reference sources are compiled where they lie
// under /root/reference (see oracle/Makefile); nothing is copied. The reference only needs OpenGL interop to
// obtain a raw device pointer to its Particle[] VBO (src/CUDA/NeighbourSearchCUDA.cu:145-148,186 and the same
// pattern in every wrapper), so the Makefile renames the five cudaGraphics* entry points it calls to the
// akref_gl_* stand-ins defined here, which hand out a cudaMalloc'd buffer instead of a mapped VBO.
//
// Two ways to drive it, both exported with a C ABI for ctypes:
//   * akref_step            -> AkuaEngine::PBFSolver::step (src/Simulation/PBFSolver.cpp:22-78), untouched.
//   * akref_phase_*         -> the six free wrapper functions, called in the order of PBFSolver.cpp:30-77, so the
//                              parity tests can observe / teacher-force the state between phases.
#include <AkuaEngine/Simulation/PBFSolver.h>
#include <AkuaEngine/Simulation/PBFConfig.h>
#include <AkuaEngine/Simulation/Particle.h>
#include <AkuaEngine/Interop/InteropResource.h>
#include <AkuaEngine/CUDA/IntegrationCUDA.h>
#include <AkuaEngine/CUDA/NeighbourSearchCUDA.h>
#include <AkuaEngine/CUDA/ConstraintSolverCUDA.h>
#include <cuda_runtime.h>
#include <cstdint>
#include <cstring>
#include <map>
#include <memory>
#include <vector>

using namespace AkuaEngine;

static_assert(sizeof(Particle) == 108, "reference Particle must be 108 bytes");

namespace {
struct FakeVbo { void* dptr; size_t bytes; };
std::map<unsigned, FakeVbo*>& vboTable() { static std::map<unsigned, FakeVbo*> t; return t; }
unsigned g_nextVbo = 1;
}

// ---- stand-ins for the GL-interop calls (signatures identical to the CUDA runtime's) ----
extern "C" cudaError_t akref_gl_register(cudaGraphicsResource** resource, GLuint buffer, unsigned int) {
    auto it = vboTable().find(buffer);
    if (it == vboTable().end()) return cudaErrorInvalidValue;
    *resource = reinterpret_cast<cudaGraphicsResource*>(it->second);
    return cudaSuccess;
}
extern "C" cudaError_t akref_gl_unregister(cudaGraphicsResource_t) { return cudaSuccess; }
extern "C" cudaError_t akref_gl_map(int, cudaGraphicsResource_t*, cudaStream_t) { return cudaSuccess; }
extern "C" cudaError_t akref_gl_unmap(int, cudaGraphicsResource_t*, cudaStream_t) { return cudaDeviceSynchronize(); }
extern "C" cudaError_t akref_gl_get_ptr(void** devPtr, size_t* size, cudaGraphicsResource_t resource) {
    FakeVbo* v = reinterpret_cast<FakeVbo*>(resource);
    *devPtr = v->dptr;
    if (size) *size = v->bytes;
    return cudaSuccess;
}

struct AkRef {
    int n = 0;
    unsigned vboId = 0;
    FakeVbo vbo{nullptr, 0};
    PBFConfig cfg;
    LambdaCorrParams corr;
    std::unique_ptr<InteropResource> interop;
    std::unique_ptr<PBFSolver> solver;
    // phase-mode mirrors of PBFSolver's private members (src/Simulation/PBFSolver.cpp:13-20)
    int tableSize = 0;
    std::vector<uint32_t> nbrArray, nbrCount;
};

// Flat parameter block so ctypes does not depend on glm/bool layout:
// [0] restDensity [1] particle_spacing [2] smoothRadius [3] spatialHashCellSize [4] relaxation
// [5] vorticityEpsilon [6] viscosity [7] maxNeighbours [8] solverIterations [9..11] gravity
// [12] corr.k [13] corr.n [14] corr.delta_q
extern "C" void* akref_create(int n, const float* p) {
    auto* r = new AkRef();
    r->n = n;
    r->cfg.restDensity = p[0]; r->cfg.particle_spacing = p[1]; r->cfg.smoothRadius = p[2];
    r->cfg.spatialHashCellSize = p[3]; r->cfg.relaxation = p[4]; r->cfg.vorticityEpsilon = p[5];
    r->cfg.viscosity = p[6]; r->cfg.maxNeighbours = (int)p[7]; r->cfg.solverIterations = (int)p[8];
    r->cfg.gravity = glm::vec3(p[9], p[10], p[11]);
    r->corr.enabled = true; r->corr.k = p[12]; r->corr.n = p[13]; r->corr.delta_q = p[14];
    r->vbo.bytes = (size_t)n * sizeof(Particle);
    if (cudaMalloc(&r->vbo.dptr, r->vbo.bytes) != cudaSuccess) { delete r; return nullptr; }
    cudaMemset(r->vbo.dptr, 0, r->vbo.bytes);
    r->vboId = g_nextVbo++;
    vboTable()[r->vboId] = &r->vbo;
    r->interop.reset(new InteropResource(r->vboId));
    r->solver.reset(new PBFSolver(n, r->cfg, r->corr));
    r->tableSize = r->cfg.maxNeighbours * n;
    r->nbrArray.assign((size_t)n * r->cfg.maxNeighbours, 0u);
    r->nbrCount.assign((size_t)n, 0u);
    return r;
}
extern "C" void akref_destroy(void* h) {
    auto* r = static_cast<AkRef*>(h);
    if (!r) return;
    r->solver.reset(); r->interop.reset();
    vboTable().erase(r->vboId);
    cudaFree(r->vbo.dptr);
    delete r;
}
extern "C" int akref_upload(void* h, const void* aos108) {
    auto* r = static_cast<AkRef*>(h);
    return (int)cudaMemcpy(r->vbo.dptr, aos108, r->vbo.bytes, cudaMemcpyHostToDevice);
}
extern "C" int akref_download(void* h, void* aos108) {
    auto* r = static_cast<AkRef*>(h);
    return (int)cudaMemcpy(aos108, r->vbo.dptr, r->vbo.bytes, cudaMemcpyDeviceToHost);
}
extern "C" int akref_set_gravity(void* h, const float* g) {
    auto* r = static_cast<AkRef*>(h);
    r->cfg.gravity = glm::vec3(g[0], g[1], g[2]);
    r->solver->setGravity(r->cfg.gravity);
    return 0;
}
static int lastErr() { cudaDeviceSynchronize(); return (int)cudaGetLastError(); }

extern "C" int akref_step(void* h, float dt, const float* bmin, const float* bmax) {
    auto* r = static_cast<AkRef*>(h);
    r->solver->step(*r->interop, dt, glm::vec3(bmin[0], bmin[1], bmin[2]), glm::vec3(bmax[0], bmax[1], bmax[2]));
    return lastErr();
}
// ---- phase mode: same calls, same order, same arguments as PBFSolver::step ----
extern "C" int akref_phase_predict(void* h, float dt) {
    auto* r = static_cast<AkRef*>(h);
    IntegrationCUDA::predictNewPositionCUDA(r->interop->getGraphicsResource(), r->n, r->cfg.gravity, dt);
    return lastErr();
}
extern "C" int akref_phase_neighbours(void* h) {
    auto* r = static_cast<AkRef*>(h);
    NeighbourSearchCUDA::findParticleNeighboursCUDA(r->interop->getGraphicsResource(), r->n, r->nbrArray.data(),
        r->nbrCount.data(), r->cfg.smoothRadius, r->cfg.spatialHashCellSize, r->tableSize, r->cfg.maxNeighbours);
    return lastErr();
}
extern "C" int akref_phase_solve(void* h, int iterations, const float* bmin, const float* bmax) {
    auto* r = static_cast<AkRef*>(h);
    ConstraintSolverCUDA::runConstraintSolverCUDA(r->interop->getGraphicsResource(), r->n, iterations,
        r->nbrArray.data(), r->nbrCount.data(), r->cfg.smoothRadius, r->cfg.maxNeighbours, r->cfg.restDensity,
        r->cfg.relaxation, glm::vec3(bmin[0], bmin[1], bmin[2]), glm::vec3(bmax[0], bmax[1], bmax[2]), r->corr);
    return lastErr();
}
extern "C" int akref_phase_update(void* h, float dt) {
    auto* r = static_cast<AkRef*>(h);
    IntegrationCUDA::updatePositionAndVelocityCUDA(r->interop->getGraphicsResource(), r->n, dt);
    return lastErr();
}
extern "C" int akref_phase_damping(void* h, const float* bmin, const float* bmax) {
    auto* r = static_cast<AkRef*>(h);
    IntegrationCUDA::applyBoundaryVelocityDampingCUDA(r->interop->getGraphicsResource(), r->n,
        glm::vec3(bmin[0], bmin[1], bmin[2]), glm::vec3(bmax[0], bmax[1], bmax[2]), 0.0f, 0.95f);
    return lastErr();
}
extern "C" int akref_phase_vorticity_viscosity(void* h, float dt) {
    auto* r = static_cast<AkRef*>(h);
    IntegrationCUDA::applyVorticityAndViscosityCUDA(r->interop->getGraphicsResource(), r->n, r->nbrArray.data(),
        r->nbrCount.data(), r->cfg.smoothRadius, dt, r->cfg.vorticityEpsilon, r->cfg.viscosity, r->cfg.maxNeighbours);
    return lastErr();
}
extern "C" int akref_get_neighbours(void* h, uint32_t* arr, uint32_t* cnt) {
    auto* r = static_cast<AkRef*>(h);
    if (arr) std::memcpy(arr, r->nbrArray.data(), r->nbrArray.size() * sizeof(uint32_t));
    if (cnt) std::memcpy(cnt, r->nbrCount.data(), r->nbrCount.size() * sizeof(uint32_t));
    return 0;
}
extern "C" int akref_max_neighbours(void* h) { return static_cast<AkRef*>(h)->cfg.maxNeighbours; }
