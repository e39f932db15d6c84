// C++ host-side smoke test of the drop-in facade (include/akua_pbf.hpp): builds the README dam-break lattice the way
// Application::prepareDamBreak does (src/Application/Application.cpp:162-192), steps it, and checks conservation and
// that the fluid falls. Compiled by tests/test_cpp_facade.py with plain g++ against libakua_pbf.so (no CUDA headers).
#include <array>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <vector>

#include "akua_pbf.hpp"

struct Particle {  // include/AkuaEngine/Simulation/Particle.h:8-31 with glm::vec3 -> float[3]
    float position[3], velocity[3], new_position[3], new_velocity[3], position_delta[3], vorticity[3];
    float mass, density, lambda;
    uint32_t hash;
    float color[4];
    float size;
};
static_assert(sizeof(Particle) == 108, "layout");

int main(int argc, char** argv) {
    const bool run = argc > 1 && std::strcmp(argv[1], "--run") == 0;
    akua::PBFConfig config;
    akua::LambdaCorrParams corr;
    if (config.restDensity != 7600.0f || config.maxNeighbours != 128 || corr.n != 4.0f) return 2;
    if (!run) { std::puts("link-ok"); return 0; }
    const int side = 20;
    std::vector<Particle> particles;
    for (int x = 0; x < side; ++x)
        for (int y = 0; y < side; ++y)
            for (int z = 0; z < side; ++z) {
                Particle p{};
                p.position[0] = 2.0f + x * config.particle_spacing;
                p.position[1] = 1.0f + y * config.particle_spacing;
                p.position[2] = 2.0f + z * config.particle_spacing;
                p.mass = 1.0f;
                p.color[2] = p.color[3] = 1.0f;
                p.size = 50.0f;
                particles.push_back(p);
            }
    try {
        akua::PBFSolver solver((int64_t)particles.size(), config, corr);
        solver.uploadParticles(particles.data(), (int64_t)particles.size());
        const std::array<float, 3> boxMin{1.5f, 0.0f, 1.5f}, boxMax{4.5f, 4.0f, 4.5f};
        double y0 = 0, y1 = 0;
        for (auto& p : particles) y0 += p.position[1];
        for (int i = 0; i < 30; ++i) solver.step(0.0083f, boxMin, boxMax);
        solver.downloadParticles(particles.data(), (int64_t)particles.size());
        for (auto& p : particles) {
            if (!std::isfinite(p.position[0] + p.position[1] + p.position[2])) return 3;
            y1 += p.position[1];
        }
        float mean, mx;
        solver.densityError(mean, mx);
        std::printf("mean height %.4f -> %.4f, density error mean %.4f max %.4f\n", y0 / particles.size(),
                    y1 / particles.size(), mean, mx);
        if (!(y1 < y0)) return 4;  // gravity acts
    } catch (const akua::Error& e) {
        std::fprintf(stderr, "akua::Error %d: %s\n", e.status, e.what());
        return 5;
    }
    std::puts("facade-ok");
    return 0;
}
