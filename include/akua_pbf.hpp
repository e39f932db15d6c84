// akua_pbf.hpp — header-only C++ facade over the C ABI (akua_pbf.h), shaped like AkuaEngine::PBFSolver
// (include/AkuaEngine/Simulation/PBFSolver.h:12-27) so it can be dropped in behind src/Simulation:
//
//     akua::PBFSolver solver(numParticles, config, corrParams);   // PBFSolver.h:14
//     solver.uploadParticles(hostParticles, n);                   // replaces the VBO upload (Renderer.cpp:185-219)
//     solver.step(dt, boxMin, boxMax);                            // PBFSolver.h:17
//     solver.setGravity(g);                                       // PBFSolver.h:18
//
// Vectors are anything indexable with [0..2] (glm::vec3, float[3], std::array<float,3>), so the reference's call sites
// compile unchanged apart from the particle-buffer argument. Errors surface as akua::Error (the reference ignores them).
#ifndef AKUA_PBF_HPP
#define AKUA_PBF_HPP

#include <cstdint>
#include <stdexcept>
#include <string>

#include "akua_pbf.h"

namespace akua {

struct Error : std::runtime_error {
    int status;
    Error(int s, const std::string& what) : std::runtime_error(what), status(s) {}
};

// Field-for-field twins of the reference's structs (PBFConfig.h:10-29) with the reference's defaults.
struct LambdaCorrParams : akua_corr_params {
    LambdaCorrParams() { akua_pbf_default_corr(this); }
};
struct PBFConfig : akua_pbf_config {
    PBFConfig() { akua_pbf_default_config(this); }
};
struct Options : akua_pbf_options {
    Options() { akua_pbf_default_options(this); }
};

class PBFSolver {
public:
    PBFSolver(int64_t numParticles, const PBFConfig& config, const LambdaCorrParams& corrParams,
              const Options& options = Options()) {
        int rc = akua_pbf_create(&h_, numParticles, &config, &corrParams, &options);
        if (rc != AKUA_OK) {
            std::string msg = h_ ? akua_pbf_last_error(h_) : "invalid arguments or no CUDA device";
            if (h_) akua_pbf_destroy(h_);
            h_ = nullptr;
            throw Error(rc, "akua_pbf_create: " + msg);
        }
    }
    ~PBFSolver() { if (h_) akua_pbf_destroy(h_); }
    PBFSolver(const PBFSolver&) = delete;
    PBFSolver& operator=(const PBFSolver&) = delete;
    PBFSolver(PBFSolver&& o) noexcept : h_(o.h_) { o.h_ = nullptr; }
    PBFSolver& operator=(PBFSolver&& o) noexcept {
        if (this != &o) { if (h_) akua_pbf_destroy(h_); h_ = o.h_; o.h_ = nullptr; }
        return *this;
    }

    // void step(const InteropResource&, float deltaTime, glm::vec3 boxMin, glm::vec3 boxMax) — PBFSolver.h:17
    template <typename V3>
    void step(float deltaTime, const V3& boxMin, const V3& boxMax) {
        const float a[3] = {boxMin[0], boxMin[1], boxMin[2]}, b[3] = {boxMax[0], boxMax[1], boxMax[2]};
        check(akua_pbf_step(h_, deltaTime, a, b), "step");
    }
    // step(dt, solverIterations) shape
    template <typename V3>
    void step(float deltaTime, int solverIterations, const V3& boxMin, const V3& boxMax) {
        const float a[3] = {boxMin[0], boxMin[1], boxMin[2]}, b[3] = {boxMax[0], boxMax[1], boxMax[2]};
        check(akua_pbf_step_iters(h_, deltaTime, solverIterations, a, b), "step");
    }
    // void setGravity(glm::vec3) — PBFSolver.h:18
    template <typename V3>
    void setGravity(const V3& g) {
        const float v[3] = {g[0], g[1], g[2]};
        check(akua_pbf_set_gravity(h_, v), "setGravity");
    }
    void sync() { check(akua_pbf_sync(h_), "sync"); }

    // `particles` points to n reference-layout structs (sizeof == 108), e.g. std::vector<AkuaEngine::Particle>::data()
    void uploadParticles(const void* particles, int64_t n) { check(akua_pbf_upload_aos108(h_, particles, n), "upload"); }
    void downloadParticles(void* particles, int64_t n) { check(akua_pbf_download_aos108(h_, particles, n), "download"); }
    // GL consumer: `deviceDst` is device memory, e.g. the mapped particle VBO (stride 108, Renderer.cpp:201-213)
    void downloadParticlesDevice(void* deviceDst) { check(akua_pbf_export_aos108_device(h_, deviceDst, numParticles()), "export"); }
    // the same through the registered VBO handle the reference keeps in InteropResource (map, write, unmap on the solver's stream)
    void exportToGraphicsResource(void* cudaGraphicsResource) { check(akua_pbf_export_to_graphics_resource(h_, cudaGraphicsResource), "export"); }
    const float* positionsDevice() { return akua_pbf_positions_device(h_); }   // float4 per particle, device memory
    const float* velocitiesDevice() { return akua_pbf_velocities_device(h_); }
    int64_t numParticles() const { return akua_pbf_num_particles(h_); }
    void densityError(float& mean, float& max) { check(akua_pbf_density_error(h_, &mean, &max), "densityError"); }
    akua_pbf_solver* handle() { return h_; }

private:
    void check(int rc, const char* what) {
        if (rc != AKUA_OK) throw Error(rc, std::string(what) + ": " + akua_pbf_last_error(h_));
    }
    akua_pbf_solver* h_ = nullptr;
};

}  // namespace akua
#endif  // AKUA_PBF_HPP
