/* akua_pbf.h — C ABI of the B200-native PBF simulation step (drop-in for AkuaEngine's src/Simulation solver surface).
 *
 * Every entry point cites the reference interface it replaces (paths relative to the reference repository root).
 * Plain pointers and sizes only; no C++ / torch / glm types cross this boundary. All functions return 0 on success and
 * a non-zero akua_status otherwise (the reference's wrappers return void and ignore CUDA errors; see
 * akua_pbf_last_error). Nothing throws across the ABI. A solver handle is single-caller and not re-entrant, like the
 * reference's PBFSolver (one caller: Application::run, src/Application/Application.cpp:63-70).
 *
 * The library is CUDA-only (sm_100a). There is no CPU fallback: akua_pbf_create fails with AKUA_ERR_NO_DEVICE when
 * no CUDA device is usable.
 */
#ifndef AKUA_PBF_H
#define AKUA_PBF_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define AKUA_PBF_ABI_VERSION 1

typedef struct akua_pbf_solver akua_pbf_solver; /* opaque; replaces AkuaEngine::PBFSolver (PBFSolver.h:12-27) */

typedef enum akua_status {
    AKUA_OK = 0,
    AKUA_ERR_INVALID = 1,   /* bad argument / unsupported configuration */
    AKUA_ERR_NO_DEVICE = 2, /* no usable CUDA device (no CPU fallback exists) */
    AKUA_ERR_CUDA = 3,      /* a CUDA call or kernel failed; see akua_pbf_last_error */
    AKUA_ERR_ALLOC = 4,
    AKUA_ERR_COMM = 5
} akua_status;

/* Mirrors AkuaEngine::LambdaCorrParams field-for-field (include/AkuaEngine/Simulation/PBFConfig.h:10-15).
 * `enabled` is carried for layout parity only: the reference never reads it (src/CUDA/ConstraintSolverCUDA.cu:123-126
 * always applies the term), and neither does this library. Use k = 0 to switch artificial pressure off. */
typedef struct akua_corr_params {
    int32_t enabled; /* bool in the reference */
    float k;         /* 0.0001f */
    float n;         /* 4.0f */
    float delta_q;   /* 0.03f */
} akua_corr_params;

/* Mirrors AkuaEngine::PBFConfig field-for-field and in order (PBFConfig.h:18-29); glm::vec3 gravity -> float[3]. */
typedef struct akua_pbf_config {
    float restDensity;         /* 7600 */
    float particle_spacing;    /* 0.05 (scene construction only; the solver does not read it) */
    float smoothRadius;        /* 0.1  */
    float spatialHashCellSize; /* 0.1; must equal smoothRadius (the reference silently breaks otherwise:
                                  NeighbourSearchCUDA.cu:163 hashes with smoothRadius, :177 looks up with cellSize) */
    float relaxation;          /* 600  */
    float vorticityEpsilon;    /* 1e-5 */
    float viscosity;           /* 0.01 */
    int32_t maxNeighbours;     /* 128  */
    int32_t solverIterations;  /* 4    */
    float gravity[3];          /* 0,-9.8,0 */
} akua_pbf_config;

/* How particles are keyed for the neighbour search. Neighbour SETS are identical in both modes (when the neighbour cap
 * does not bite and no two of a particle's 27 cells collide in the reference's hash); order inside a set, and therefore
 * float summation order, differs. */
typedef enum akua_key_mode {
    /* key = reference hash (NeighbourSearchCUDA.cu:15-27) mod tableSize = maxNeighbours*n (PBFSolver.cpp:15).
     * Keys, sorted permutation, bucket-start table and neighbour lists are bit-identical to the reference. */
    AKUA_KEY_REFERENCE_HASH = 0,
    /* key = dense linear cell index (x-major, z-fastest) over a grid covering the box; spatially coherent order. */
    AKUA_KEY_LINEAR_CELL = 1
} akua_key_mode;

/* Knobs that exist only on the B200 side (no reference counterpart). Zero-initialise, then akua_pbf_default_options. */
/* How the neighbour sweeps fetch a neighbour's data (single-GPU path; x-slab mode always uses the plain layout).
 * The sweeps are bound by L1 gather wavefronts (profiles/), so each sweep wants ONE gather per neighbour:
 *   packed  : pass A also writes (x*, lambda) as one float4 and K11 writes (x, |omega|), so that pass B and K12 gather
 *             16 bytes from one array instead of 16 + 4 bytes from two (-11 % on both kernels, measured). Needs every
 *             particle to have the same mass (checked at each upload; the reference's scenes use mass 1,
 *             Application.cpp:186) — otherwise those two sweeps silently keep the plain layout.
 *   records : the committing pass B also writes 32-byte (position, velocity) records; K11 and K13 fetch a neighbour with
 *             one 256-bit load (LDG.E.256) instead of two 128-bit loads. Measured -3 % on those kernels, cancelled by the
 *             record writes: kept as an opt-in, not part of the default.
 * Results are bit-identical in every layout (tests/test_parity_gpu.py::test_gather_layouts_are_bit_identical). */
typedef enum akua_gather_layout {
    AKUA_GATHER_AUTO = 0,            /* = packed */
    AKUA_GATHER_PLAIN = 1,
    AKUA_GATHER_PACKED = 2,
    AKUA_GATHER_RECORDS = 3,
    AKUA_GATHER_PACKED_RECORDS = 4
} akua_gather_layout;

/* How the neighbour lists are built in LINEAR_CELL mode (REFERENCE_HASH always scans its buckets). Lists are bit-identical
 * in every variant (tests/test_list_build_host.py on the CPU, test_list_build_variants_identical on the GPU).
 *   scan  : one candidate at a time, test and append in the same loop (k_build_neighbours): the plain statement of the
 *           reference's traversal, kept as the cross-check of the other two.
 *   mask4 / mask8 : two phases per row chunk of <= 32 candidates — a hit bitmask from 4 / 8 independent loads in flight, then
 *           the set bits are appended (akuaengine_b200/csrc/list_build.cuh). mask4 is the default since round 2 (measured on a
 *           B200: -14 % at 1 M, -7 % at 4 M particles against scan; reachability culling on top of it was measured and
 *           rejected, profiles/r02_c3_list_build_culling_rejected.txt). A zero-initialised options struct still selects scan. */
typedef enum akua_list_build {
    AKUA_LIST_BUILD_SCAN = 0,
    AKUA_LIST_BUILD_MASK4 = 1,
    AKUA_LIST_BUILD_MASK8 = 2
} akua_list_build;

/* Boundary handling. 0 = the reference's: no boundary model, only the soft clamp of handle_particle_collision
 * (src/CUDA/ConstraintSolverCUDA.cu:132-157, "Later we will use virtual particles") — the default, and the only mode inside the
 * parity contract. 1 = OPT-IN upgrade: every box wall is backed by a half-space of virtual fluid at rest density in closed form
 * (its poly6 density and spiky gradient enter a near-wall particle's density, lambda and delta-p; akuaengine_b200/csrc/wall_model.cuh),
 * on top of the clamp. Changes results by design; single-GPU only; the measured kernels are not touched by it. */
typedef enum akua_wall_model {
    AKUA_WALL_REFERENCE = 0,
    AKUA_WALL_VIRTUAL_FLUID = 1
} akua_wall_model;

typedef struct akua_pbf_options {
    int32_t key_mode;        /* akua_key_mode; default AKUA_KEY_LINEAR_CELL */
    int32_t device;          /* CUDA device ordinal; default 0 */
    int32_t use_graph;       /* capture the step into a CUDA graph and replay it (default 1) */
    int32_t fast_math;       /* 1 (default) = r and 1/r of the spiky gradient from one MUFU rsqrt (max 2 ulp, the same error
                                class as the reference's own powf calls); 0 = IEEE sqrtf and division */
    float capacity_factor;   /* device arrays are sized for capacity_factor * n particles (ghosts, migration); default 1 */
    int32_t gather_layout;   /* akua_gather_layout; default AKUA_GATHER_AUTO. Results are bit-identical in every layout. */
    int32_t use_pdl;         /* 1 (default) = the step's kernels are launched with programmatic dependent launch: each is
                                scheduled while its predecessor drains (also between the eagerly launched kernels of the x-slab path); 0 = plain stream order */
    int32_t list_build;      /* akua_list_build; akua_pbf_default_options selects AKUA_LIST_BUILD_MASK4 (the field took the first of
                                the formerly reserved words: a zero-initialised struct from an older caller selects scan) */
    int32_t canonical_order; /* 0 (default) = particles of one cell keep last step's relative order (stable sort by cell key, like the
                                reference's sort); 1 = they are ordered by particle id (an id sort ahead of the key sort: four more
                                digit passes). With 1 the neighbour order, hence every float sum, no longer depends on history: an
                                x-slab run on any number of GPUs, with migration and re-balancing, is then BIT-IDENTICAL to the
                                single-GPU run (tests/mgpu_worker.py --canonical). Ids must be unique. */
    int32_t wall_model;      /* akua_wall_model; default AKUA_WALL_REFERENCE */
    int32_t reserved[3];
} akua_pbf_options;

void akua_pbf_default_config(akua_pbf_config* cfg);   /* PBFConfig{} defaults */
void akua_pbf_default_corr(akua_corr_params* corr);   /* LambdaCorrParams{} defaults */
void akua_pbf_default_options(akua_pbf_options* opt);
int akua_pbf_abi_version(void);

/* PBFSolver::PBFSolver(int numParticles, const PBFConfig&, const LambdaCorrParams&) — PBFSolver.h:14,
 * src/Simulation/PBFSolver.cpp:13-20. All device memory is allocated here; step() allocates nothing. `opt` may be NULL. */
int akua_pbf_create(akua_pbf_solver** out, int64_t numParticles, const akua_pbf_config* cfg,
                    const akua_corr_params* corr, const akua_pbf_options* opt);
void akua_pbf_destroy(akua_pbf_solver* s);

/* void PBFSolver::step(const InteropResource&, float deltaTime, glm::vec3 boxMin, glm::vec3 boxMax) — PBFSolver.h:17,
 * src/Simulation/PBFSolver.cpp:22-78. The particle buffer is owned by the solver (see upload/download) instead of
 * being passed as a GL-interop handle. Iteration count comes from config.solverIterations (PBFSolver.cpp:48).
 * Asynchronous on the solver's stream; akua_pbf_sync waits. */
int akua_pbf_step(akua_pbf_solver* s, float deltaTime, const float boxMin[3], const float boxMax[3]);
/* step(dt, solverIterations) shape named by the north-star: same as akua_pbf_step with an explicit iteration count. */
int akua_pbf_step_iters(akua_pbf_solver* s, float deltaTime, int32_t solverIterations, const float boxMin[3],
                        const float boxMax[3]);
/* Headless fixed-timestep driver: the accumulator loop of Application::run (src/Application/Application.cpp:37-70).
 * accumulator += frameTime; while (accumulator >= deltaTime && steps < maxStepsPerFrame) step. MAX_STEPS_PER_FRAME is 3
 * in the reference (Application.cpp:19). Steps replay a captured CUDA graph when options.use_graph is set. */
int akua_pbf_advance(akua_pbf_solver* s, float frameTime, float deltaTime, int32_t maxStepsPerFrame, const float boxMin[3],
                     const float boxMax[3], int32_t* stepsDone);
/* `steps` back-to-back steps of deltaTime (asynchronous). */
int akua_pbf_run_steps(akua_pbf_solver* s, int32_t steps, float deltaTime, const float boxMin[3], const float boxMax[3]);
/* void PBFSolver::setGravity(glm::vec3) — PBFSolver.h:18,29-31 */
int akua_pbf_set_gravity(akua_pbf_solver* s, const float gravity[3]);
int akua_pbf_sync(akua_pbf_solver* s);
const char* akua_pbf_last_error(const akua_pbf_solver* s);
int64_t akua_pbf_num_particles(const akua_pbf_solver* s);

/* ---- particle buffer interchange (replaces the caller-owned VBO of Particle[N], Renderer.cpp:185-219) ----
 * AoS-108 is the reference's `Particle` (include/AkuaEngine/Simulation/Particle.h:8-31): position@0 velocity@12
 * new_position@24 new_velocity@36 position_delta@48 vorticity@60 mass@72 density@76 lambda@80 hash@84 color@88 size@104.
 * upload: `src` is a HOST pointer to n <= capacity structs. All solver-visible fields are imported; upload order defines
 * particle ids and n becomes the live particle count.
 * download: `dst` is a HOST pointer; particles come back in the solver's current (key-sorted) order, exactly as the
 * reference leaves its VBO after the in-place sort (NeighbourSearchCUDA.cu:167-170). Fields the reference overwrites
 * before reading on the next step and this solver does not keep (new_velocity) are filled with their commit-time
 * equivalents (new_velocity := velocity). */
int akua_pbf_upload_aos108(akua_pbf_solver* s, const void* src, int64_t n);
int akua_pbf_download_aos108(akua_pbf_solver* s, void* dst, int64_t n);
/* Optional OpenGL consumer (replaces the in-place VBO update): writes the AoS-108 buffer straight into DEVICE memory,
 * e.g. the pointer obtained from cudaGraphicsResourceGetMappedPointer on the renderer's VBO (layout the reference's
 * renderer binds: position@0, color@88, size@104, stride 108 — src/Rendering/Renderer.cpp:201-213). Asynchronous. */
int akua_pbf_export_aos108_device(akua_pbf_solver* s, void* device_dst, int64_t n);
/* The same, through a registered graphics resource: `graphicsResource` is the cudaGraphicsResource* of the particle VBO
 * (the handle the reference keeps in InteropResource, src/Interop/InteropResource.cpp:9-35, and passes to every wrapper,
 * e.g. src/CUDA/IntegrationCUDA.cu:199-214). Maps it on the solver's stream, checks its size (>= 108 * numParticles bytes),
 * writes the AoS-108 buffer and unmaps — the map / get-pointer / unmap triple each reference wrapper performs, once per
 * frame instead of six times per step. Asynchronous (stream-ordered); no GL headers are needed on either side. */
int akua_pbf_export_to_graphics_resource(akua_pbf_solver* s, void* graphicsResource);
/* Lean interchange: xyz triples + mass (mass may be NULL => 1.0). Host pointers. */
int akua_pbf_upload_soa(akua_pbf_solver* s, const float* pos_xyz, const float* vel_xyz, const float* mass, int64_t n);
/* pos4/vel4 are n float4 (xyz + mass / xyz + density); id is the upload index of each returned particle. Any may be NULL. */
int akua_pbf_download_soa(akua_pbf_solver* s, float* pos4, float* vel4, uint32_t* id, int64_t n);
/* Zero-copy consumers (the "optional OpenGL consumer": a renderer can cudaMemcpy / map from these). DEVICE pointers to
 * n float4, valid until the next step. */
const float* akua_pbf_positions_device(akua_pbf_solver* s);
const float* akua_pbf_velocities_device(akua_pbf_solver* s);

/* ---- checkpoint / resume ----
 * Saves / restores the simulation state a step depends on: positions (+mass), velocities (+density), particle ids,
 * payload (color, size), gravity, step count and the fixed-timestep accumulator. Particle order is preserved, so a run
 * resumed from a checkpoint continues bit-identically. Synchronises. */
int akua_pbf_checkpoint_save(akua_pbf_solver* s, const char* path);
int akua_pbf_checkpoint_load(akua_pbf_solver* s, const char* path);

/* ---- multi-GPU: x-slab domain decomposition, one process per GPU (no reference counterpart) ----
 * Each rank creates its own solver (capacity_factor > 1 leaves room for ghosts and arrivals), then:
 *   rank 0: akua_pbf_comm_unique_id(buf, 128) and broadcasts buf by any means (torch.distributed, MPI, a file);
 *   all:    akua_pbf_comm_init(s, rank, nranks, buf)  — loads NCCL with dlopen and builds the communicator;
 *           akua_pbf_set_slab(s, xCellLo, xCellHi)    — owned interval of absolute x cells floor(x / smoothRadius);
 *                                                        contiguous across ranks; the end ranks own everything beyond;
 *           upload the owned particles (upload_* sets the live count; akua_pbf_upload_ids gives them global ids);
 *           akua_pbf_step(...) with the SAME box on every rank. It migrates particles whose predicted position left the
 *           slab, exchanges ghost planes over NVLink (x*, lambda per iteration; v, |omega| post-solve) and steps the owned
 *           particles. Nothing of a step's sizes is known to the host: migration counts, plane sizes and the exchange
 *           epochs live on the device, every sweep is ONE launch whose first CTAs compute the boundary planes, store them
 *           into the neighbours' arrays (CUDA-IPC peer pointers) and publish an epoch the neighbours' CTAs wait for — the
 *           step is a single CUDA graph (NCCL send/recv with one host synchronisation per step is the fallback transport).
 *           Downloads return the owned particles; akua_pbf_num_particles is the owned count. The render payload (color,
 *           size) migrates with its particle, like the reference's struct follows its sort (NeighbourSearchCUDA.cu:167-170). */
int akua_pbf_comm_unique_id(void* out, int64_t out_bytes);
/* COLLECTIVE CALLS in slab mode: akua_pbf_set_slab, akua_pbf_rebalance and — once akua_pbf_set_slab has been called — every
 * upload (akua_pbf_upload_aos108 / _soa, akua_pbf_checkpoint_load): each contains one small all-reduce in which the ranks
 * agree whether all particles of all ranks share one mass (the packed gather layout, akua_gather_layout, is what the halo
 * exchanges then carry). Every rank must make these calls the same number of times, a rank without particles with n = 0. */
int akua_pbf_comm_init(akua_pbf_solver* s, int32_t rank, int32_t nranks, const void* unique_id);
int akua_pbf_set_slab(akua_pbf_solver* s, int32_t xCellLo, int32_t xCellHi);
int akua_pbf_upload_ids(akua_pbf_solver* s, const uint32_t* ids, int64_t n);
/* Collective (every rank, same step): moves the slab boundaries towards equal WORK (a particle weighs 12 + the largest
 * neighbour count in its warp, corrected per rank by its measured busy time when steps are long enough for the clock to
 * mean something) using the current per-x-plane sums (one small ncclAllReduce); no slab is given more particles than 80 % of
 * what its arrays hold; the following step's migration transfers the particles.
 * A partition whose heaviest slab is within 2 % of the mean is left alone (environment AKUA_SLAB_KEEP_BELOW, default 1.02).
 * Call every few dozen steps for scenes whose fluid moves along x (dam break, sloshing tank). */
int akua_pbf_rebalance(akua_pbf_solver* s);
/* The same without a host synchronisation: applies the measurement the PREVIOUS call enqueued (histogram kernels, all-reduce,
 * copy to pinned memory — long finished by then), then enqueues the next one behind the last step. The host never waits and
 * the GPU never idles; the boundaries lag one call behind the fluid. Collective like akua_pbf_rebalance; the two may be mixed. */
int akua_pbf_rebalance_async(akua_pbf_solver* s);
/* out: 0 owned, 1 ghosts from left, 2 ghosts from right, 3 first-plane size, 4 last-plane size, 5 exchanges so far,
 * 6 bytes sent so far (NEGATIVE when the CUDA-IPC peer-to-peer transport is in use, positive for NCCL send/recv),
 * 7 particles migrated in so far. Transport: ghost planes are copied straight into the neighbour's arrays through
 * CUDA IPC when every rank could open its neighbours' allocations (environment AKUA_SLAB_P2P=0 forces NCCL). */
int akua_pbf_slab_stats(const akua_pbf_solver* s, int64_t out[8]);
/* Where a rank waited for its neighbours, cumulative, measured on the device clock: out[0] nanoseconds idle in the per-step
 * count exchange (a rank that is ahead of its neighbours idles there: load imbalance), out[1] nanoseconds the first boundary
 * CTA of the sweeps waited for ghost planes (halo latency that was not hidden behind the interior), out[2] steps completed
 * on the device, out[3] re-balancing calls that moved a boundary. Synchronises. */
int akua_pbf_slab_wait_stats(const akua_pbf_solver* s, int64_t out[4]);
/* Balanced slab boundaries from a per-x-column particle histogram. Pure host code (callable without a GPU). */
int akua_slab_partition(const int64_t* hist, int32_t ncols, int32_t nranks, int32_t* bounds /* nranks + 1 */);
/* Boundaries for a re-balancing step from the global per-column histogram and the current boundaries (pure host code; what
 * akua_pbf_rebalance evaluates on every rank): the balanced partition, clamped so that the next step's ordinary migration
 * can carry the transfer — boundary r stays strictly inside the two old slabs it separates, slabs stay >= 2 columns wide,
 * and at most maxMove particles cross a boundary. oldBounds / bounds hold nranks + 1 entries (first 0, last ncols). */
int akua_slab_rebalance_bounds(const int64_t* hist, int32_t ncols, int32_t nranks, const int32_t* oldBounds, int64_t maxMove,
                               int32_t* bounds);
/* The same with a WORK histogram deciding where the boundaries go (akua_pbf_rebalance weighs a particle by 12 + its neighbour
 * count: the sweeps' cost follows the neighbour count, and a sloshing tank is denser on one side) while `count` (particles per
 * column) still bounds what may cross a boundary. keepBelow > 1: if the heaviest slab of the CURRENT partition carries at most
 * keepBelow x the mean work, the boundaries are left where they are. maxCount > 0: no slab of the target partition holds more
 * particles than that (the room of a rank's arrays), whatever the work says. */
int akua_slab_rebalance_bounds_weighted(const int64_t* work, const int64_t* count, int32_t ncols, int32_t nranks,
                                        const int32_t* oldBounds, int64_t maxMove, double keepBelow, int64_t maxCount, int32_t* bounds);

/* Page-locked host memory for the interchange buffers (so uploads/downloads run at full PCIe rate). */
void* akua_pbf_host_alloc(int64_t bytes);
void akua_pbf_host_free(void* p);

/* ---- phase-level operators: the six free functions PBFSolver::step calls, same decomposition, same order ----
 * They double as the teacher-forced parity hooks (upload the oracle's pre-phase state, run one phase, compare). */
/* predictNewPositionCUDA — include/AkuaEngine/CUDA/IntegrationCUDA.h:12, src/CUDA/IntegrationCUDA.cu:199-214 */
int akua_pbf_phase_predict(akua_pbf_solver* s, float deltaTime);
/* findParticleNeighboursCUDA — NeighbourSearchCUDA.h:12-21, src/CUDA/NeighbourSearchCUDA.cu:134-187.
 * Uses the box of the last step/solve call (or the grid given at creation) to lay out the LINEAR_CELL grid. */
int akua_pbf_phase_neighbours(akua_pbf_solver* s, const float boxMin[3], const float boxMax[3]);
/* runConstraintSolverCUDA — ConstraintSolverCUDA.h:14-27, src/CUDA/ConstraintSolverCUDA.cu:173-222 */
int akua_pbf_phase_solve(akua_pbf_solver* s, int32_t solverIterations, const float boxMin[3], const float boxMax[3]);
/* updatePositionAndVelocityCUDA — IntegrationCUDA.h:13, src/CUDA/IntegrationCUDA.cu:216-230 */
int akua_pbf_phase_update(akua_pbf_solver* s, float deltaTime);
/* applyBoundaryVelocityDampingCUDA(…, restitution 0, friction 0.95) — IntegrationCUDA.h:14-21, PBFSolver.cpp:64 */
int akua_pbf_phase_damping(akua_pbf_solver* s, const float boxMin[3], const float boxMax[3]);
/* applyVorticityAndViscosityCUDA — IntegrationCUDA.h:22-32, src/CUDA/IntegrationCUDA.cu:253-286 */
int akua_pbf_phase_vorticity_viscosity(akua_pbf_solver* s, float deltaTime);

/* ---- debug taps (no reference counterpart; required to check integer parity) ---- */
typedef enum akua_debug_which {
    AKUA_DBG_KEYS_UNSORTED = 0, /* u32[n]  key of each particle in pre-sort order (= Particle::hash after K2) */
    AKUA_DBG_KEYS_SORTED = 1,   /* u32[n]  (= Particle::hash after the sort) */
    AKUA_DBG_PERM = 2,          /* u32[n]  sorted slot -> pre-sort slot */
    AKUA_DBG_ID = 3,            /* u32[n]  sorted slot -> upload index */
    AKUA_DBG_BUCKET_START = 4,  /* u32[tableSize] hashToFirstParticleIndex (REFERENCE_HASH mode; UINT32_MAX = empty) */
    AKUA_DBG_CELL_RANGE = 5,    /* u32[2*numCells] (start,end) per linear cell (LINEAR_CELL mode) */
    AKUA_DBG_NBR_COUNT = 6,     /* u32[n] */
    AKUA_DBG_NBR_LIST = 7,      /* u32[n*maxNeighbours], row-major like the reference's neighbourArray */
    AKUA_DBG_DENSITY = 8,       /* f32[n] */
    AKUA_DBG_LAMBDA = 9,        /* f32[n] */
    AKUA_DBG_XSTAR = 10,        /* f32[4n] predicted position, w = mass */
    AKUA_DBG_POSITION = 11,     /* f32[4n] */
    AKUA_DBG_VELOCITY = 12,     /* f32[4n] w = density */
    AKUA_DBG_VORTICITY = 13,    /* f32[4n] w = |omega| */
    AKUA_DBG_DELTA_P = 14       /* f32[4n] last position_delta */
} akua_debug_which;
/* Copies the selected array to host memory `dst` (dst_bytes must be large enough). Synchronises. */
int akua_pbf_debug_get(akua_pbf_solver* s, int32_t which, void* dst, int64_t dst_bytes);
int64_t akua_pbf_debug_size(akua_pbf_solver* s, int32_t which); /* bytes needed for akua_pbf_debug_get, <0 if n/a */

/* Per-step density-constraint error |rho_i/rho0 - 1| over all particles, from the last density pass (north-star
 * acceptance criterion for long runs). Synchronises. */
int akua_pbf_density_error(akua_pbf_solver* s, float* mean, float* max);

/* ---- instrumentation ---- */
typedef struct akua_pbf_counters {
    int64_t kernel_launches;  /* kernels of this library launched (or replayed inside a graph) since creation */
    int64_t steps;
    int64_t sort_passes_last; /* radix passes used by the last neighbour phase */
    int64_t num_cells;        /* linear grid cells (LINEAR_CELL) or tableSize (REFERENCE_HASH) */
    int64_t h2d_bytes, d2h_bytes; /* bytes moved by upload_* / download_* / debug_get since creation */
    int64_t graph_replays;    /* steps executed by replaying a captured CUDA graph */
} akua_pbf_counters;
int akua_pbf_get_counters(const akua_pbf_solver* s, akua_pbf_counters* out);
/* Timing of the phases of the LAST akua_pbf_step call, from CUDA events recorded on the solver's stream when
 * enabled with akua_pbf_enable_timing(s, 1) (disables graph replay). ms[]: 0 predict+key, 1 sort, 2 reorder+ranges,
 * 3 neighbour lists, 4 constraint solve (all iterations), 5 post-solve (vorticity, confinement, XSPH), 6 whole step,
 * 7 sum of the density+lambda launches (pass A), 8 sum of the delta-p+apply launches (pass B), 9 launches per pass. */
int akua_pbf_enable_timing(akua_pbf_solver* s, int32_t on);
int akua_pbf_last_step_timing(akua_pbf_solver* s, float ms[10]);
/* Launch timeline of ONE step (a diagnostic, no reference counterpart): the NEXT akua_pbf_step runs eagerly (no graph replay)
 * with a CUDA event after every launch on the stream it went to, synchronises, and appends one JSON object per launch to
 * `path` (rank, step, seq, lane (always 0: every launch goes to the solver's stream), name, end_ms since the start of the step,
 * since_prev_on_lane_ms). In x-slab mode the time a rank spends waiting for a neighbour shows up in the launch that waits
 * (k_slab_plan: the count message; a sweep: the ghost planes its boundary CTAs read). */
int akua_pbf_trace_next_step(akua_pbf_solver* s, const char* path);
/* The CUDA stream (cudaStream_t) all of this solver's work is issued on, so callers can record their own events on it
 * or order other work after it. */
void* akua_pbf_stream(akua_pbf_solver* s);

#ifdef __cplusplus
}
#endif
#endif /* AKUA_PBF_H */
