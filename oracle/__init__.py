"""oracle/ — TEST INFRASTRUCTURE (checkers), never imported by the product package akuaengine_b200.

  RefOracle  : the UNMODIFIED reference kernels rebuilt headless (oracle/_ref/libakua_ref.so, needs a GPU).
  PortOracle : the host-C++ restatement (oracle/libpbf_oracle.so, CPU, OpenMP).
Both expose the same methods, named after the reference's free functions, and operate on AoS-108 particle arrays.
"""
from __future__ import annotations

import ctypes as C
import subprocess
from pathlib import Path

import numpy as np

ORACLE_DIR = Path(__file__).resolve().parent
PORT_LIB = ORACLE_DIR / "libpbf_oracle.so"
REF_LIB = ORACLE_DIR / "_ref" / "libakua_ref.so"

PARTICLE_DTYPE = np.dtype([
    ("position", "<f4", 3), ("velocity", "<f4", 3), ("new_position", "<f4", 3), ("new_velocity", "<f4", 3),
    ("position_delta", "<f4", 3), ("vorticity", "<f4", 3), ("mass", "<f4"), ("density", "<f4"), ("lambda", "<f4"),
    ("hash", "<u4"), ("color", "<f4", 4), ("size", "<f4"),
])


def build_port(force: bool = False) -> Path:
    src = ORACLE_DIR / "pbf_oracle.cpp"
    if force or not PORT_LIB.exists() or PORT_LIB.stat().st_mtime < src.stat().st_mtime:
        subprocess.run(["make", "-C", str(ORACLE_DIR), "-B", "port"], check=True, capture_output=True)
    return PORT_LIB


def build_ref(force: bool = False) -> Path | None:
    """Only possible where /root/reference exists (the dev container); the built .so travels to the GPU box."""
    if not Path("/root/reference/src/CUDA").exists():
        return REF_LIB if REF_LIB.exists() else None
    if force or not REF_LIB.exists() or REF_LIB.stat().st_mtime < (ORACLE_DIR / "ref_harness.cu").stat().st_mtime:
        subprocess.run(["make", "-C", str(ORACLE_DIR), "-B", "ref"], check=True, capture_output=True)
    return REF_LIB


def _f3(v):
    return (C.c_float * 3)(*[float(x) for x in v])


def param_block(cfg=None, corr=None, **kw) -> np.ndarray:
    """15-float parameter block (layout in ref_harness.cu). cfg/corr: akuaengine_b200.PBFConfig / LambdaCorrParams."""
    d = dict(restDensity=7600.0, particle_spacing=0.05, smoothRadius=0.1, spatialHashCellSize=0.1, relaxation=600.0,
             vorticityEpsilon=1e-5, viscosity=0.01, maxNeighbours=128, solverIterations=4, gravity=(0.0, -9.8, 0.0),
             k=1e-4, n=4.0, delta_q=0.03)
    if cfg is not None:
        for f in ("restDensity", "particle_spacing", "smoothRadius", "spatialHashCellSize", "relaxation",
                  "vorticityEpsilon", "viscosity", "maxNeighbours", "solverIterations"):
            d[f] = getattr(cfg, f)
        d["gravity"] = tuple(cfg.gravity)
    if corr is not None:
        d["k"], d["n"], d["delta_q"] = corr.k, corr.n, corr.delta_q
    d.update(kw)
    return np.array([d["restDensity"], d["particle_spacing"], d["smoothRadius"], d["spatialHashCellSize"],
                     d["relaxation"], d["vorticityEpsilon"], d["viscosity"], d["maxNeighbours"], d["solverIterations"],
                     *d["gravity"], d["k"], d["n"], d["delta_q"]], dtype=np.float32)


class PortOracle:
    """CPU restatement. Particles live in a caller-visible numpy AoS array (self.particles), mutated in place."""

    kind = "port"

    def __init__(self, particles: np.ndarray, params: np.ndarray):
        build_port()
        lib = C.CDLL(str(PORT_LIB))
        vp, f3 = C.c_void_p, C.POINTER(C.c_float)
        lib.pbfo_create.restype = vp
        lib.pbfo_create.argtypes = [C.c_int, f3]
        lib.pbfo_destroy.argtypes = [vp]
        lib.pbfo_destroy.restype = None
        lib.pbfo_set_gravity.argtypes = [vp, f3]
        lib.pbfo_step.argtypes = [vp, vp, C.c_float, f3, f3]
        lib.pbfo_phase_predict.argtypes = [vp, vp, C.c_float]
        lib.pbfo_phase_neighbours.argtypes = [vp, vp]
        lib.pbfo_phase_solve.argtypes = [vp, vp, C.c_int, f3, f3]
        lib.pbfo_phase_update.argtypes = [vp, vp, C.c_float]
        lib.pbfo_phase_damping.argtypes = [vp, vp, f3, f3]
        lib.pbfo_phase_vorticity_viscosity.argtypes = [vp, vp, C.c_float]
        lib.pbfo_get_neighbours.argtypes = [vp, vp, vp]
        self._lib = lib
        self.params = np.ascontiguousarray(params, dtype=np.float32)
        self.particles = np.ascontiguousarray(particles.copy())
        assert self.particles.dtype.itemsize == 108
        self.n = len(self.particles)
        self.maxNeighbours = int(self.params[7])
        self._h = lib.pbfo_create(self.n, self.params.ctypes.data_as(f3))
        self.threads = lib.pbfo_threads()

    def close(self):
        if self._h:
            self._lib.pbfo_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def _p(self):
        return self.particles.ctypes.data

    def upload(self, particles):
        self.particles[...] = particles

    def download(self):
        return self.particles.copy()

    def setGravity(self, g):
        self._lib.pbfo_set_gravity(self._h, _f3(g))

    def step(self, dt, bmin, bmax):
        self._lib.pbfo_step(self._h, self._p, dt, _f3(bmin), _f3(bmax))

    def predictNewPosition(self, dt):
        self._lib.pbfo_phase_predict(self._h, self._p, dt)

    def findParticleNeighbours(self):
        self._lib.pbfo_phase_neighbours(self._h, self._p)

    def runConstraintSolver(self, iters, bmin, bmax):
        self._lib.pbfo_phase_solve(self._h, self._p, int(iters), _f3(bmin), _f3(bmax))

    def updatePositionAndVelocity(self, dt):
        self._lib.pbfo_phase_update(self._h, self._p, dt)

    def applyBoundaryVelocityDamping(self, bmin, bmax):
        self._lib.pbfo_phase_damping(self._h, self._p, _f3(bmin), _f3(bmax))

    def applyVorticityAndViscosity(self, dt):
        self._lib.pbfo_phase_vorticity_viscosity(self._h, self._p, dt)

    def neighbours(self):
        arr = np.empty((self.n, self.maxNeighbours), np.uint32)
        cnt = np.empty(self.n, np.uint32)
        self._lib.pbfo_get_neighbours(self._h, arr.ctypes.data, cnt.ctypes.data)
        return arr, cnt


class RefOracle:
    """The unmodified reference (GPU). Same surface as PortOracle."""

    kind = "reference"

    def __init__(self, particles: np.ndarray, params: np.ndarray):
        if not REF_LIB.exists():
            raise FileNotFoundError(f"{REF_LIB} not built (make -C oracle ref, needs /root/reference)")
        lib = C.CDLL(str(REF_LIB))
        vp, f3 = C.c_void_p, C.POINTER(C.c_float)
        lib.akref_create.restype = vp
        lib.akref_create.argtypes = [C.c_int, f3]
        lib.akref_destroy.argtypes = [vp]
        lib.akref_destroy.restype = None
        lib.akref_upload.argtypes = [vp, vp]
        lib.akref_download.argtypes = [vp, vp]
        lib.akref_set_gravity.argtypes = [vp, f3]
        lib.akref_step.argtypes = [vp, C.c_float, f3, f3]
        lib.akref_phase_predict.argtypes = [vp, C.c_float]
        lib.akref_phase_neighbours.argtypes = [vp]
        lib.akref_phase_solve.argtypes = [vp, C.c_int, f3, f3]
        lib.akref_phase_update.argtypes = [vp, C.c_float]
        lib.akref_phase_damping.argtypes = [vp, f3, f3]
        lib.akref_phase_vorticity_viscosity.argtypes = [vp, C.c_float]
        lib.akref_get_neighbours.argtypes = [vp, vp, vp]
        self._lib = lib
        self.params = np.ascontiguousarray(params, dtype=np.float32)
        self.n = len(particles)
        self.maxNeighbours = int(self.params[7])
        self._h = lib.akref_create(self.n, self.params.ctypes.data_as(f3))
        if not self._h:
            raise RuntimeError("akref_create failed (no CUDA device?)")
        self.threads = 1
        self.upload(particles)

    def close(self):
        if self._h:
            self._lib.akref_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _ck(self, rc, what):
        if rc != 0:
            raise RuntimeError(f"reference {what} failed: CUDA error {rc}")

    def upload(self, particles):
        p = np.ascontiguousarray(particles)
        assert p.dtype.itemsize == 108 and len(p) == self.n
        self._ck(self._lib.akref_upload(self._h, p.ctypes.data), "upload")

    def download(self):
        out = np.empty(self.n, dtype=PARTICLE_DTYPE)
        self._ck(self._lib.akref_download(self._h, out.ctypes.data), "download")
        return out

    @property
    def particles(self):
        return self.download()

    def setGravity(self, g):
        self._lib.akref_set_gravity(self._h, _f3(g))

    def step(self, dt, bmin, bmax):
        self._ck(self._lib.akref_step(self._h, dt, _f3(bmin), _f3(bmax)), "step")

    def predictNewPosition(self, dt):
        self._ck(self._lib.akref_phase_predict(self._h, dt), "predict")

    def findParticleNeighbours(self):
        self._ck(self._lib.akref_phase_neighbours(self._h), "neighbours")

    def runConstraintSolver(self, iters, bmin, bmax):
        self._ck(self._lib.akref_phase_solve(self._h, int(iters), _f3(bmin), _f3(bmax)), "solve")

    def updatePositionAndVelocity(self, dt):
        self._ck(self._lib.akref_phase_update(self._h, dt), "update")

    def applyBoundaryVelocityDamping(self, bmin, bmax):
        self._ck(self._lib.akref_phase_damping(self._h, _f3(bmin), _f3(bmax)), "damping")

    def applyVorticityAndViscosity(self, dt):
        self._ck(self._lib.akref_phase_vorticity_viscosity(self._h, dt), "vorticity_viscosity")

    def neighbours(self):
        arr = np.empty((self.n, self.maxNeighbours), np.uint32)
        cnt = np.empty(self.n, np.uint32)
        self._lib.akref_get_neighbours(self._h, arr.ctypes.data, cnt.ctypes.data)
        return arr, cnt
