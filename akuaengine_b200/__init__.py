"""akuaengine_b200 — B200-native Position-Based-Fluids step behind AkuaEngine's solver surface.

The product is the CUDA library `libakua_pbf.so` (hand-written sm_100a kernels + C++ host solver, C ABI in
include/akua_pbf.h). This package is the thin Python binding over that C ABI, mirroring the reference's
`AkuaEngine::PBFSolver` / `PBFConfig` / `LambdaCorrParams` (include/AkuaEngine/Simulation/*.h) so tests and benchmarks read
like code written against the reference. There is NO CPU fallback: if the library is missing or no CUDA device is
usable, construction raises.
"""
from __future__ import annotations

import ctypes as C
from pathlib import Path

import numpy as np

from .build import LIB_PATH, build_library  # noqa: F401

__all__ = [
    "PBFConfig", "LambdaCorrParams", "PBFOptions", "PBFSolver", "PARTICLE_DTYPE", "AkuaError", "load_library",
    "KEY_REFERENCE_HASH", "KEY_LINEAR_CELL", "DBG", "PinnedBuffer",
    "GATHER_AUTO", "GATHER_PLAIN", "GATHER_PACKED", "GATHER_RECORDS", "GATHER_PACKED_RECORDS",
    "LIST_BUILD_SCAN", "LIST_BUILD_MASK4", "LIST_BUILD_MASK8",
]

KEY_REFERENCE_HASH = 0
KEY_LINEAR_CELL = 1
# akua_gather_layout (include/akua_pbf.h): how the sweeps fetch a neighbour; results are bit-identical in every layout
GATHER_AUTO, GATHER_PLAIN, GATHER_PACKED, GATHER_RECORDS, GATHER_PACKED_RECORDS = 0, 1, 2, 3, 4
# akua_list_build (include/akua_pbf.h): how LINEAR_CELL neighbour lists are built; lists are bit-identical in every variant
LIST_BUILD_SCAN, LIST_BUILD_MASK4, LIST_BUILD_MASK8 = 0, 1, 2

# include/AkuaEngine/Simulation/Particle.h:8-31 — packed, 108 bytes
PARTICLE_DTYPE = np.dtype([
    ("position", "<f4", 3), ("velocity", "<f4", 3), ("new_position", "<f4", 3), ("new_velocity", "<f4", 3),
    ("position_delta", "<f4", 3), ("vorticity", "<f4", 3), ("mass", "<f4"), ("density", "<f4"), ("lambda", "<f4"),
    ("hash", "<u4"), ("color", "<f4", 4), ("size", "<f4"),
])
assert PARTICLE_DTYPE.itemsize == 108


class AkuaError(RuntimeError):
    pass


class LambdaCorrParams(C.Structure):
    """AkuaEngine::LambdaCorrParams (PBFConfig.h:10-15)."""
    _fields_ = [("enabled", C.c_int32), ("k", C.c_float), ("n", C.c_float), ("delta_q", C.c_float)]

    def __init__(self, **kw):
        super().__init__()
        self.enabled, self.k, self.n, self.delta_q = 1, 0.0001, 4.0, 0.03
        for k, v in kw.items():
            setattr(self, k, v)


class PBFConfig(C.Structure):
    """AkuaEngine::PBFConfig (PBFConfig.h:18-29), field names as in the reference."""
    _fields_ = [
        ("restDensity", C.c_float), ("particle_spacing", C.c_float), ("smoothRadius", C.c_float),
        ("spatialHashCellSize", C.c_float), ("relaxation", C.c_float), ("vorticityEpsilon", C.c_float),
        ("viscosity", C.c_float), ("maxNeighbours", C.c_int32), ("solverIterations", C.c_int32),
        ("gravity", C.c_float * 3),
    ]

    def __init__(self, **kw):
        super().__init__()
        self.restDensity, self.particle_spacing, self.smoothRadius, self.spatialHashCellSize = 7600.0, 0.05, 0.1, 0.1
        self.relaxation, self.vorticityEpsilon, self.viscosity = 600.0, 0.00001, 0.01
        self.maxNeighbours, self.solverIterations = 128, 4
        self.gravity[:] = [0.0, -9.8, 0.0]
        for k, v in kw.items():
            if k == "gravity":
                self.gravity[:] = list(v)
            else:
                setattr(self, k, v)

    def as_param_block(self, corr: "LambdaCorrParams") -> np.ndarray:
        """The flat 15-float block the oracle libraries take (oracle/ref_harness.cu)."""
        return np.array([self.restDensity, self.particle_spacing, self.smoothRadius, self.spatialHashCellSize,
                         self.relaxation, self.vorticityEpsilon, self.viscosity, self.maxNeighbours,
                         self.solverIterations, *self.gravity, corr.k, corr.n, corr.delta_q], dtype=np.float32)


class PBFOptions(C.Structure):
    _fields_ = [("key_mode", C.c_int32), ("device", C.c_int32), ("use_graph", C.c_int32), ("fast_math", C.c_int32),
                ("capacity_factor", C.c_float), ("gather_layout", C.c_int32), ("use_pdl", C.c_int32), ("list_build", C.c_int32), ("canonical_order", C.c_int32),
                ("wall_model", C.c_int32), ("reserved", C.c_int32 * 3)]


class Counters(C.Structure):
    _fields_ = [("kernel_launches", C.c_int64), ("steps", C.c_int64), ("sort_passes_last", C.c_int64),
                ("num_cells", C.c_int64), ("h2d_bytes", C.c_int64), ("d2h_bytes", C.c_int64), ("graph_replays", C.c_int64)]


class DBG:
    KEYS_UNSORTED, KEYS_SORTED, PERM, ID, BUCKET_START, CELL_RANGE, NBR_COUNT, NBR_LIST = range(8)
    DENSITY, LAMBDA, XSTAR, POSITION, VELOCITY, VORTICITY, DELTA_P = range(8, 15)
    _DTYPES = {0: "u4", 1: "u4", 2: "u4", 3: "u4", 4: "u4", 5: "u4", 6: "u4", 7: "u4",
               8: "f4", 9: "f4", 10: "f4", 11: "f4", 12: "f4", 13: "f4", 14: "f4"}


# every symbol include/akua_pbf.h declares (tests check that the library exports all of them)
ABI_SYMBOLS = [
    "akua_pbf_abi_version", "akua_pbf_default_config", "akua_pbf_default_corr", "akua_pbf_default_options",
    "akua_pbf_create", "akua_pbf_destroy", "akua_pbf_step", "akua_pbf_step_iters", "akua_pbf_set_gravity",
    "akua_pbf_advance", "akua_pbf_run_steps", "akua_pbf_checkpoint_save", "akua_pbf_checkpoint_load",
    "akua_pbf_sync", "akua_pbf_last_error", "akua_pbf_num_particles", "akua_pbf_upload_aos108",
    "akua_pbf_download_aos108", "akua_pbf_export_aos108_device", "akua_pbf_export_to_graphics_resource", "akua_pbf_upload_soa", "akua_pbf_download_soa", "akua_pbf_positions_device",
    "akua_pbf_velocities_device", "akua_pbf_host_alloc", "akua_pbf_host_free", "akua_pbf_phase_predict",
    "akua_pbf_phase_neighbours", "akua_pbf_phase_solve", "akua_pbf_phase_update", "akua_pbf_phase_damping",
    "akua_pbf_phase_vorticity_viscosity", "akua_pbf_debug_get", "akua_pbf_debug_size", "akua_pbf_density_error",
    "akua_pbf_get_counters", "akua_pbf_enable_timing", "akua_pbf_last_step_timing", "akua_pbf_trace_next_step", "akua_pbf_stream",
    "akua_pbf_comm_unique_id", "akua_pbf_comm_init", "akua_pbf_set_slab", "akua_pbf_upload_ids", "akua_pbf_slab_stats", "akua_pbf_slab_wait_stats", "akua_pbf_rebalance", "akua_pbf_rebalance_async", "akua_slab_partition", "akua_slab_rebalance_bounds", "akua_slab_rebalance_bounds_weighted",
]

_lib = None


def load_library(path: str | Path | None = None) -> C.CDLL:
    """Loads libakua_pbf.so (built in-tree by akuaengine_b200.build / __graft_entry__.build). Fails loudly if absent."""
    global _lib
    if _lib is not None and path is None:
        return _lib
    import os
    if path is None and os.environ.get("AKUA_PBF_LIB"):
        path = os.environ["AKUA_PBF_LIB"]  # alternative build of the same library (tuning experiments)
    p = Path(path) if path else LIB_PATH
    if not p.exists():
        raise AkuaError(f"{p} not found: the CUDA library is not built. Run `python -m akuaengine_b200.build` "
                        "(or __graft_entry__.build()). There is no CPU fallback.")
    lib = C.CDLL(str(p))
    f3 = C.POINTER(C.c_float)
    vp = C.c_void_p
    lib.akua_pbf_abi_version.restype = C.c_int
    lib.akua_pbf_create.argtypes = [C.POINTER(vp), C.c_int64, C.POINTER(PBFConfig), C.POINTER(LambdaCorrParams),
                                    C.POINTER(PBFOptions)]
    lib.akua_pbf_destroy.argtypes = [vp]
    lib.akua_pbf_destroy.restype = None
    lib.akua_pbf_default_options.argtypes = [C.POINTER(PBFOptions)]
    lib.akua_pbf_default_options.restype = None
    lib.akua_pbf_default_config.argtypes = [C.POINTER(PBFConfig)]
    lib.akua_pbf_default_config.restype = None
    lib.akua_pbf_default_corr.argtypes = [C.POINTER(LambdaCorrParams)]
    lib.akua_pbf_default_corr.restype = None
    lib.akua_pbf_step.argtypes = [vp, C.c_float, f3, f3]
    lib.akua_pbf_step_iters.argtypes = [vp, C.c_float, C.c_int32, f3, f3]
    lib.akua_pbf_set_gravity.argtypes = [vp, f3]
    lib.akua_pbf_advance.argtypes = [vp, C.c_float, C.c_float, C.c_int32, f3, f3, C.POINTER(C.c_int32)]
    lib.akua_pbf_run_steps.argtypes = [vp, C.c_int32, C.c_float, f3, f3]
    lib.akua_pbf_checkpoint_save.argtypes = [vp, C.c_char_p]
    lib.akua_pbf_checkpoint_load.argtypes = [vp, C.c_char_p]
    lib.akua_pbf_sync.argtypes = [vp]
    lib.akua_pbf_last_error.argtypes = [vp]
    lib.akua_pbf_last_error.restype = C.c_char_p
    lib.akua_pbf_num_particles.argtypes = [vp]
    lib.akua_pbf_num_particles.restype = C.c_int64
    lib.akua_pbf_upload_aos108.argtypes = [vp, vp, C.c_int64]
    lib.akua_pbf_download_aos108.argtypes = [vp, vp, C.c_int64]
    lib.akua_pbf_export_aos108_device.argtypes = [vp, vp, C.c_int64]
    lib.akua_pbf_export_to_graphics_resource.argtypes = [vp, vp]
    lib.akua_pbf_upload_soa.argtypes = [vp, vp, vp, vp, C.c_int64]
    lib.akua_pbf_download_soa.argtypes = [vp, vp, vp, vp, C.c_int64]
    lib.akua_pbf_positions_device.argtypes = [vp]
    lib.akua_pbf_positions_device.restype = vp
    lib.akua_pbf_velocities_device.argtypes = [vp]
    lib.akua_pbf_velocities_device.restype = vp
    lib.akua_pbf_host_alloc.argtypes = [C.c_int64]
    lib.akua_pbf_host_alloc.restype = vp
    lib.akua_pbf_host_free.argtypes = [vp]
    lib.akua_pbf_host_free.restype = None
    lib.akua_pbf_phase_predict.argtypes = [vp, C.c_float]
    lib.akua_pbf_phase_neighbours.argtypes = [vp, f3, f3]
    lib.akua_pbf_phase_solve.argtypes = [vp, C.c_int32, f3, f3]
    lib.akua_pbf_phase_update.argtypes = [vp, C.c_float]
    lib.akua_pbf_phase_damping.argtypes = [vp, f3, f3]
    lib.akua_pbf_phase_vorticity_viscosity.argtypes = [vp, C.c_float]
    lib.akua_pbf_debug_get.argtypes = [vp, C.c_int32, vp, C.c_int64]
    lib.akua_pbf_debug_size.argtypes = [vp, C.c_int32]
    lib.akua_pbf_debug_size.restype = C.c_int64
    lib.akua_pbf_density_error.argtypes = [vp, f3, f3]
    lib.akua_pbf_get_counters.argtypes = [vp, C.POINTER(Counters)]
    lib.akua_pbf_enable_timing.argtypes = [vp, C.c_int32]
    lib.akua_pbf_last_step_timing.argtypes = [vp, f3]
    lib.akua_pbf_trace_next_step.argtypes = [vp, C.c_char_p]
    lib.akua_pbf_stream.argtypes = [vp]
    lib.akua_pbf_stream.restype = vp
    lib.akua_pbf_comm_unique_id.argtypes = [vp, C.c_int64]
    lib.akua_pbf_comm_init.argtypes = [vp, C.c_int32, C.c_int32, vp]
    lib.akua_pbf_set_slab.argtypes = [vp, C.c_int32, C.c_int32]
    lib.akua_pbf_upload_ids.argtypes = [vp, vp, C.c_int64]
    lib.akua_pbf_slab_stats.argtypes = [vp, C.POINTER(C.c_int64)]
    lib.akua_pbf_slab_wait_stats.argtypes = [vp, C.POINTER(C.c_int64)]
    lib.akua_pbf_rebalance.argtypes = [vp]
    lib.akua_pbf_rebalance_async.argtypes = [vp]
    lib.akua_slab_partition.argtypes = [C.POINTER(C.c_int64), C.c_int32, C.c_int32, C.POINTER(C.c_int32)]
    lib.akua_slab_rebalance_bounds.argtypes = [C.POINTER(C.c_int64), C.c_int32, C.c_int32, C.POINTER(C.c_int32), C.c_int64,
                                               C.POINTER(C.c_int32)]
    lib.akua_slab_rebalance_bounds_weighted.argtypes = [C.POINTER(C.c_int64), C.POINTER(C.c_int64), C.c_int32, C.c_int32, C.POINTER(C.c_int32),
                                                        C.c_int64, C.c_double, C.c_int64, C.POINTER(C.c_int32)]
    if path is None:
        _lib = lib
    return lib


def _vec3(v):
    return (C.c_float * 3)(*[float(x) for x in v])


class PinnedBuffer:
    """Page-locked host array (cudaMallocHost through the C ABI), exposed as a numpy array."""

    def __init__(self, shape, dtype):
        self._lib = load_library()
        dt = np.dtype(dtype)
        count = int(np.prod(shape))
        self._ptr = self._lib.akua_pbf_host_alloc(max(1, count * dt.itemsize))
        if not self._ptr:
            raise AkuaError("cudaMallocHost failed")
        buf = (C.c_char * (count * dt.itemsize)).from_address(self._ptr)
        self.array = np.frombuffer(buf, dtype=dt, count=count).reshape(shape)

    def free(self):
        if self._ptr:
            self.array = None
            self._lib.akua_pbf_host_free(self._ptr)
            self._ptr = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


class PBFSolver:
    """Mirror of AkuaEngine::PBFSolver (PBFSolver.h:12-27): ctor(numParticles, config, corrParams), step, setGravity.

    The particle buffer the reference receives as a GL-interop handle is owned by the solver here; use
    upload_particles / download_particles (AoS-108, the reference's Particle layout) or upload / download (lean SoA).
    """

    def __init__(self, numParticles: int, config: PBFConfig | None = None, corrParams: LambdaCorrParams | None = None,
                 key_mode: int = KEY_LINEAR_CELL, device: int = 0, fast_math: bool = True, use_graph: bool = True,
                 capacity_factor: float = 1.0, gather_layout: int = GATHER_AUTO, use_pdl: bool | None = None,
                 list_build: int | None = None, canonical_order: bool = False, wall_model: int = 0, lib=None):
        # `lib`: an alternative build of the same C ABI (load_library(path)), e.g. the host-emulated build of tests/emu
        self._lib = lib if lib is not None else load_library()
        self.config = config or PBFConfig()
        self.corrParams = corrParams or LambdaCorrParams()
        self.numParticles = int(numParticles)
        opt = PBFOptions()
        self._lib.akua_pbf_default_options(C.byref(opt))
        opt.key_mode, opt.device, opt.fast_math, opt.use_graph = int(key_mode), int(device), int(fast_math), int(use_graph)
        opt.capacity_factor = float(capacity_factor)
        opt.gather_layout = int(gather_layout)
        if use_pdl is not None:
            opt.use_pdl = int(use_pdl)
        if list_build is not None:
            opt.list_build = int(list_build)
        opt.canonical_order = int(bool(canonical_order))
        opt.wall_model = int(wall_model)   # 0 = reference (soft clamp only), 1 = opt-in virtual-fluid walls (akua_wall_model)
        self.options = opt
        self._h = C.c_void_p()
        rc = self._lib.akua_pbf_create(C.byref(self._h), self.numParticles, C.byref(self.config),
                                       C.byref(self.corrParams), C.byref(opt))
        if rc != 0:
            msg = self._lib.akua_pbf_last_error(self._h).decode() if self._h else "invalid arguments or no CUDA device"
            if self._h:
                self._lib.akua_pbf_destroy(self._h)
                self._h = None
            raise AkuaError(f"akua_pbf_create failed (status {rc}): {msg}")

    # -- lifetime
    def close(self):
        if getattr(self, "_h", None):
            self._lib.akua_pbf_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def _ck(self, rc, what):
        if rc != 0:
            raise AkuaError(f"{what} failed (status {rc}): {self._lib.akua_pbf_last_error(self._h).decode()}")

    # -- the reference surface
    def step(self, deltaTime: float, boxMin, boxMax, solverIterations: int | None = None):
        if solverIterations is None:
            self._ck(self._lib.akua_pbf_step(self._h, deltaTime, _vec3(boxMin), _vec3(boxMax)), "step")
        else:
            self._ck(self._lib.akua_pbf_step_iters(self._h, deltaTime, int(solverIterations), _vec3(boxMin),
                                                   _vec3(boxMax)), "step")

    def advance(self, frameTime: float, deltaTime: float, boxMin, boxMax, maxStepsPerFrame: int = 3) -> int:
        """Fixed-timestep accumulator loop of Application::run (Application.cpp:63-70); returns the steps taken."""
        n = C.c_int32(0)
        self._ck(self._lib.akua_pbf_advance(self._h, frameTime, deltaTime, maxStepsPerFrame, _vec3(boxMin), _vec3(boxMax),
                                            C.byref(n)), "advance")
        return int(n.value)

    def run_steps(self, steps: int, deltaTime: float, boxMin, boxMax):
        self._ck(self._lib.akua_pbf_run_steps(self._h, int(steps), deltaTime, _vec3(boxMin), _vec3(boxMax)), "run_steps")

    def save_checkpoint(self, path):
        self._ck(self._lib.akua_pbf_checkpoint_save(self._h, str(path).encode()), "checkpoint_save")

    def load_checkpoint(self, path):
        self._ck(self._lib.akua_pbf_checkpoint_load(self._h, str(path).encode()), "checkpoint_load")

    def setGravity(self, gravity):
        self.config.gravity[:] = [float(x) for x in gravity]
        self._ck(self._lib.akua_pbf_set_gravity(self._h, _vec3(gravity)), "setGravity")

    def sync(self):
        self._ck(self._lib.akua_pbf_sync(self._h), "sync")

    # -- particle buffer
    @property
    def n(self) -> int:
        """Live particle count (== numParticles on a single GPU; the owned count of this rank in slab mode)."""
        return int(self._lib.akua_pbf_num_particles(self._h))

    def upload_particles(self, particles: np.ndarray):
        assert particles.dtype == PARTICLE_DTYPE and particles.flags.c_contiguous
        self._ck(self._lib.akua_pbf_upload_aos108(self._h, particles.ctypes.data, len(particles)), "upload_aos108")

    def download_particles(self, out: np.ndarray | None = None) -> np.ndarray:
        if out is None:
            out = np.empty(self.n, dtype=PARTICLE_DTYPE)
        assert out.dtype == PARTICLE_DTYPE and out.flags.c_contiguous and len(out) == self.n
        self._ck(self._lib.akua_pbf_download_aos108(self._h, out.ctypes.data, len(out)), "download_aos108")
        return out

    def upload(self, pos, vel=None, mass=None):
        pos = np.ascontiguousarray(pos, dtype=np.float32).reshape(-1, 3)
        vel_p = None
        if vel is not None:
            vel = np.ascontiguousarray(vel, dtype=np.float32).reshape(-1, 3)
            vel_p = vel.ctypes.data
        mass_p = None
        if mass is not None:
            mass = np.ascontiguousarray(mass, dtype=np.float32)
            mass_p = mass.ctypes.data
        self._ck(self._lib.akua_pbf_upload_soa(self._h, pos.ctypes.data, vel_p, mass_p, len(pos)), "upload_soa")

    def download(self):
        n = self.n
        pos4 = np.empty((n, 4), np.float32)
        vel4 = np.empty((n, 4), np.float32)
        pid = np.empty(n, np.uint32)
        self._ck(self._lib.akua_pbf_download_soa(self._h, pos4.ctypes.data, vel4.ctypes.data, pid.ctypes.data, n),
                 "download_soa")
        return pos4, vel4, pid

    def positions_device_ptr(self) -> int:
        return int(self._lib.akua_pbf_positions_device(self._h) or 0)

    def velocities_device_ptr(self) -> int:
        return int(self._lib.akua_pbf_velocities_device(self._h) or 0)

    # -- phase-level operators (names follow the reference's free functions)
    def predictNewPosition(self, deltaTime):
        self._ck(self._lib.akua_pbf_phase_predict(self._h, deltaTime), "phase_predict")

    def findParticleNeighbours(self, boxMin=None, boxMax=None):
        a = _vec3(boxMin) if boxMin is not None else None
        b = _vec3(boxMax) if boxMax is not None else None
        self._ck(self._lib.akua_pbf_phase_neighbours(self._h, a, b), "phase_neighbours")

    def runConstraintSolver(self, solverIterations, boxMin, boxMax):
        self._ck(self._lib.akua_pbf_phase_solve(self._h, int(solverIterations), _vec3(boxMin), _vec3(boxMax)),
                 "phase_solve")

    def updatePositionAndVelocity(self, deltaTime):
        self._ck(self._lib.akua_pbf_phase_update(self._h, deltaTime), "phase_update")

    def applyBoundaryVelocityDamping(self, boxMin, boxMax):
        self._ck(self._lib.akua_pbf_phase_damping(self._h, _vec3(boxMin), _vec3(boxMax)), "phase_damping")

    def applyVorticityAndViscosity(self, deltaTime):
        self._ck(self._lib.akua_pbf_phase_vorticity_viscosity(self._h, deltaTime), "phase_vorticity_viscosity")

    # -- debug taps / metrics
    def debug(self, which: int) -> np.ndarray:
        nbytes = self._lib.akua_pbf_debug_size(self._h, which)
        if nbytes < 0:
            raise AkuaError(f"debug array {which} not available in this key mode")
        out = np.empty(nbytes // 4, dtype=DBG._DTYPES[which])
        self._ck(self._lib.akua_pbf_debug_get(self._h, which, out.ctypes.data, nbytes), "debug_get")
        if which in (DBG.XSTAR, DBG.POSITION, DBG.VELOCITY, DBG.VORTICITY, DBG.DELTA_P):
            return out.reshape(-1, 4)
        if which == DBG.NBR_LIST:
            return out.reshape(self.n, self.config.maxNeighbours)
        if which == DBG.CELL_RANGE:
            return out.reshape(-1, 2)
        return out

    def density_error(self):
        m, x = C.c_float(), C.c_float()
        self._ck(self._lib.akua_pbf_density_error(self._h, C.byref(m), C.byref(x)), "density_error")
        return float(m.value), float(x.value)

    def counters(self) -> dict:
        c = Counters()
        self._ck(self._lib.akua_pbf_get_counters(self._h, C.byref(c)), "get_counters")
        return {k: int(getattr(c, k)) for k, _ in Counters._fields_}

    # -- multi-GPU (x-slab) plumbing; see akuaengine_b200/slab.py for the torch.distributed driver
    @staticmethod
    def comm_unique_id(lib=None) -> np.ndarray:
        buf = np.zeros(128, np.uint8)
        rc = (lib if lib is not None else load_library()).akua_pbf_comm_unique_id(buf.ctypes.data, 128)
        if rc != 0:
            raise AkuaError(f"akua_pbf_comm_unique_id failed (status {rc}): is libnccl.so.2 loadable?")
        return buf

    def comm_init(self, rank: int, nranks: int, unique_id: np.ndarray):
        uid = np.ascontiguousarray(unique_id, dtype=np.uint8)
        assert uid.size == 128
        self._ck(self._lib.akua_pbf_comm_init(self._h, rank, nranks, uid.ctypes.data), "comm_init")

    def set_slab(self, xCellLo: int, xCellHi: int):
        self._ck(self._lib.akua_pbf_set_slab(self._h, int(xCellLo), int(xCellHi)), "set_slab")

    def upload_ids(self, ids: np.ndarray):
        ids = np.ascontiguousarray(ids, dtype=np.uint32)
        self._ck(self._lib.akua_pbf_upload_ids(self._h, ids.ctypes.data, len(ids)), "upload_ids")

    def rebalance_async(self):
        """akua_pbf_rebalance without the host synchronisation: applies the previous call's measurement, enqueues the next."""
        self._ck(self._lib.akua_pbf_rebalance_async(self._h), "rebalance_async")

    def rebalance(self):
        """Collective: re-balance the slab boundaries from the current particle distribution (all ranks, same step)."""
        self._ck(self._lib.akua_pbf_rebalance(self._h), "rebalance")

    def slab_stats(self) -> dict:
        out = (C.c_int64 * 8)()
        self._ck(self._lib.akua_pbf_slab_stats(self._h, out), "slab_stats")
        names = ["owned", "ghost_left", "ghost_right", "plane_left", "plane_right", "exchanges", "bytes_sent", "migrated_in"]
        d = dict(zip(names, [int(x) for x in out]))
        d["transport"] = "cuda-ipc p2p" if d["bytes_sent"] < 0 else "nccl"
        d["bytes_sent"] = abs(d["bytes_sent"])
        return d

    def slab_wait_stats(self) -> dict:
        out = (C.c_int64 * 4)()
        self._ck(self._lib.akua_pbf_slab_wait_stats(self._h, out), "slab_wait_stats")
        return {"plan_wait_ns": int(out[0]), "halo_wait_ns": int(out[1]), "device_steps": int(out[2]), "rebalances_moved": int(out[3])}

    def stream_ptr(self) -> int:
        """cudaStream_t of the solver, e.g. for torch.cuda.ExternalStream."""
        return int(self._lib.akua_pbf_stream(self._h) or 0)

    def enable_timing(self, on=True):
        self._ck(self._lib.akua_pbf_enable_timing(self._h, int(on)), "enable_timing")

    def trace_next_step(self, path: str):
        """The next step runs eagerly with an event after every launch; the timeline is appended to `path` (JSON lines)."""
        self._ck(self._lib.akua_pbf_trace_next_step(self._h, str(path).encode()), "trace_next_step")

    def last_step_timing(self) -> dict:
        ms = (C.c_float * 10)()
        self._ck(self._lib.akua_pbf_last_step_timing(self._h, ms), "last_step_timing")
        names = ["predict_key", "sort", "reorder_ranges", "neighbour_lists", "solve", "post", "step", "pass_a_sum",
                 "pass_b_sum", "timed_iterations"]
        return dict(zip(names, [float(x) for x in ms]))
