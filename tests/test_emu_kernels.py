"""The CUDA kernel SOURCE, run on the CPU: akuaengine_b200/csrc/*.cuh compiled by g++ against the SIMT emulator of
tests/emu/ (every CUDA thread a fiber; barriers, warp shuffles / ballots / match-any with CUDA semantics) and driven in the
order the host solver drives it (tests/emu/emu_harness.cpp mirrors stepEager() of pbf_solver.cu).

This is test infrastructure for the CPU-only container: it checks kernel LOGIC (the hand-written radix sort, the reorder /
range detection, both list builds, the fused constraint passes, the post-solve sweeps, the x-slab migration compaction)
against the golden fixtures of the unmodified reference kernels and against numpy. It says nothing about the GPU build —
`-m gpu` tests do that through the C ABI — and the product never loads the emulated code.

Tolerances are those of tests/test_parity_gpu.py: integer structures bit-exact (REFERENCE_HASH), 2e-5 after one step,
1e-3 max after ten (positions / h, velocities / (h/dt)).
"""
import ctypes as C
import subprocess
from pathlib import Path

import numpy as np
import pytest

import parity_lib as pl
from akuaengine_b200 import KEY_LINEAR_CELL, KEY_REFERENCE_HASH, PARTICLE_DTYPE, LambdaCorrParams, PBFConfig

REPO = Path(__file__).resolve().parents[1]
EMU = REPO / "tests" / "emu"
OUT = EMU / "_build" / "libakua_emu.so"
GOLDEN = REPO / "tests" / "golden"
CSRC = REPO / "akuaengine_b200" / "csrc"
DEPS = [EMU / "emu_harness.cpp", EMU / "emu_core.cpp", EMU / "cuda_runtime.h", *CSRC.glob("*.cuh"), CSRC / "pbf_params.h"]


@pytest.fixture(scope="module")
def emu():
    if not OUT.exists() or any(p.stat().st_mtime > OUT.stat().st_mtime for p in DEPS):
        OUT.parent.mkdir(parents=True, exist_ok=True)
        cmd = ["g++", "-O2", "-std=c++17", "-DAKUA_HOST_EMU", "-U_FORTIFY_SOURCE", "-ffp-contract=off", "-fPIC", "-shared", "-I", str(EMU),
               "-o", str(OUT), str(EMU / "emu_harness.cpp"), str(EMU / "emu_core.cpp")]
        res = subprocess.run(cmd, capture_output=True, text=True)
        assert res.returncode == 0, res.stderr[-4000:]
    lib = C.CDLL(str(OUT))
    vp = C.c_void_p
    lib.emu_sort_pairs.argtypes = [vp, C.c_uint32, C.c_int, vp, vp, C.c_int, C.c_int]
    lib.emu_create.restype = vp
    lib.emu_create.argtypes = [C.c_uint32, vp, vp, C.c_int, C.c_int, C.c_int, C.c_int]
    lib.emu_destroy.argtypes = [vp]
    lib.emu_upload_aos108.argtypes = [vp, vp]
    lib.emu_download_aos108.argtypes = [vp, vp]
    lib.emu_step.argtypes = [vp, C.c_float, C.c_int, vp, vp]
    lib.emu_debug_get.argtypes = [vp, C.c_int, vp]
    lib.emu_migration.argtypes = [vp, C.c_uint32, C.c_uint32, C.c_int, C.c_int, C.c_uint32, vp, C.c_uint32, vp, vp, vp]
    lib.emu_span_push_check.argtypes = [vp, C.c_uint32, C.c_uint32]
    lib.emu_plane_hist.argtypes = [vp, vp, C.c_uint32, C.c_uint32, C.c_int, vp, vp]
    lib.emu_plane_verify.argtypes = [vp, C.c_uint32, C.c_uint32, C.c_int, C.c_int, C.c_uint32, C.c_uint32, vp]
    return lib


class EmuSolver:
    """The kernels of the product, sequenced like the host solver, on host arrays."""

    def __init__(self, lib, n, params, key_mode, fast_math=True, pack=True, list_build=0):
        self.lib, self.n = lib, n
        self.cfg = PBFConfig(restDensity=float(params[0]), particle_spacing=float(params[1]), smoothRadius=float(params[2]),
                             spatialHashCellSize=float(params[3]), relaxation=float(params[4]), vorticityEpsilon=float(params[5]),
                             viscosity=float(params[6]), maxNeighbours=int(params[7]), solverIterations=int(params[8]),
                             gravity=[float(x) for x in params[9:12]])
        self.corr = LambdaCorrParams(k=float(params[12]), n=float(params[13]), delta_q=float(params[14]))
        self.h = lib.emu_create(n, C.byref(self.cfg), C.byref(self.corr), key_mode, int(fast_math), int(pack), list_build)

    def upload(self, particles):
        buf = np.ascontiguousarray(particles)
        assert buf.dtype == PARTICLE_DTYPE and len(buf) == self.n
        assert self.lib.emu_upload_aos108(self.h, buf.ctypes.data) == 0

    def step(self, dt, bmin, bmax, iters=None):
        bmin = np.ascontiguousarray(bmin, np.float32); bmax = np.ascontiguousarray(bmax, np.float32)
        it = self.cfg.solverIterations if iters is None else iters
        assert self.lib.emu_step(self.h, dt, it, bmin.ctypes.data, bmax.ctypes.data) == 0

    def download(self):
        out = np.empty(self.n, PARTICLE_DTYPE)
        assert self.lib.emu_download_aos108(self.h, out.ctypes.data) == 0
        return out

    def debug(self, which, shape=None):
        out = np.empty(shape or self.n, np.uint32)
        assert self.lib.emu_debug_get(self.h, which, out.ctypes.data) == 0
        return out

    def close(self):
        if self.h:
            self.lib.emu_destroy(self.h)
            self.h = None


@pytest.mark.parametrize("mode,items", [(0, 0), (1, 0), (1, 4), (1, 8), (1, 16)], ids=["three-kernel", "onesweep", "onesweep4", "onesweep8", "onesweep16"])
@pytest.mark.parametrize("n,bits", [(1, 8), (31, 5), (1024, 21), (1025, 21), (4097, 13), (20000, 32), (70001, 22)])
def test_radix_sort_kernels_are_a_stable_sort(emu, n, bits, mode, items):
    """Both sort variants driven by rsort::sort_pairs itself (three kernels per pass: k_count / k_scan / k_scatter; one-sweep:
    k_hist + one k_onesweep per pass with decoupled look-back): keys sorted, equal keys in input order."""
    rng = np.random.default_rng(n)
    keys = rng.integers(0, 2 ** bits, n, dtype=np.uint64).astype(np.uint32)
    if n == 70001:   # nearly sorted input with long runs of equal keys, like the cell keys of consecutive steps
        keys = np.sort(keys) // 64 * 64
        swap = rng.integers(0, n - 1, n // 20)
        keys[swap], keys[swap + 1] = keys[swap + 1].copy(), keys[swap].copy()
    ko, vo = np.empty(n, np.uint32), np.empty(n, np.uint32)
    launches = emu.emu_sort_pairs(keys.ctypes.data, n, bits, ko.ctypes.data, vo.ctypes.data, mode, items)
    passes = (bits + 7) // 8
    assert launches == (3 * passes if mode == 0 else 1 + passes)
    want = np.argsort(keys, kind="stable").astype(np.uint32)
    assert np.array_equal(vo, want) and np.array_equal(ko, keys[want])


@pytest.mark.parametrize("mode", [KEY_REFERENCE_HASH, KEY_LINEAR_CELL], ids=["hash", "linear"])
@pytest.mark.parametrize("name", ["lattice12", "jitter"])
def test_kernel_source_reproduces_the_reference_golden(emu, name, mode):
    g = dict(np.load(GOLDEN / f"{name}.npz"))
    init, dt, bmin, bmax = g["init"], float(g["dt"]), g["box_min"], g["box_max"]
    s = EmuSolver(emu, len(init), g["params"], mode)
    s.upload(init)
    for k in range(1, 11):
        s.step(dt, bmin, bmax)
        if k == 1 and mode == KEY_REFERENCE_HASH:
            # integer structures of the first neighbour phase, bit for bit (fixture: unmodified reference kernels on a B200)
            want = g["after_neighbours"]
            got = s.download()
            assert np.array_equal(got["hash"], want["hash"])                      # cell keys in sorted order
            assert np.array_equal(pl.ids_of(got), pl.ids_of(want))                # sorted permutation
            cnt = s.debug(6)
            assert np.array_equal(cnt, g["nbr_count"])
            lst = s.debug(7, (len(init), s.cfg.maxNeighbours))
            want_lst = pl.ragged_to_padded(g["nbr_flat"], g["nbr_count"], lst.shape[1])
            mask = np.arange(lst.shape[1])[None, :] < cnt[:, None]
            assert np.array_equal(lst[mask], want_lst[mask])
        if k in (1, 10):
            p = s.download()
            a = np.argsort(pl.ids_of(p)); b = np.argsort(g[f"step{k}_id"])
            dp = np.abs(p["position"][a].astype(np.float64) - g[f"step{k}_position"][b]).max() / pl.H
            dv = np.abs(p["velocity"][a].astype(np.float64) - g[f"step{k}_velocity"][b]).max() / (pl.H / dt)
            tol = 2e-5 if k == 1 else 1e-3
            assert dp < tol and dv < tol, (name, mode, k, dp, dv)
    s.close()


def test_variants_leave_identical_state(emu):
    """Gather layout (plain / packed / records / both) and list build (scan / mask4 / mask8) must not change a single bit of the state — the
    emulated counterpart of test_gather_layouts_are_bit_identical / test_list_build_variants_identical."""
    g = dict(np.load(GOLDEN / "lattice12.npz"))     # uniform masses: the packed layout is really taken
    init, dt, bmin, bmax = g["init"], float(g["dt"]), g["box_min"], g["box_max"]
    outs = {}
    for pack in (0, 1, 2, 3):           # bit 0: packed (x*, lambda) / (x, |omega|) arrays; bit 1: 32-byte (position, velocity) records
        for lb in (0, 1, 2):
            if pack >= 2 and lb:
                continue
            s = EmuSolver(emu, len(init), g["params"], KEY_LINEAR_CELL, pack=pack, list_build=lb)
            s.upload(init)
            for _ in range(3):
                s.step(dt, bmin, bmax)
            s.step(dt, bmin, bmax, iters=0)     # commit outside the fused pass B: records rebuilt by k_build_posvel
            s.step(dt, bmin, bmax)
            outs[(pack, lb)] = (s.download().tobytes(), s.debug(6).tobytes(), s.debug(7, (len(init), s.cfg.maxNeighbours)).tobytes())
            s.close()
    ref = outs[(0, 0)]
    for k, v in outs.items():
        assert v == ref, k


def test_ieee_math_path_and_zero_iterations(emu):
    """fast_math = 0 (sqrtf + division) stays within the one-step tolerance; solverIterations = 0 commits through the
    stand-alone K9 / K10 kernels."""
    g = dict(np.load(GOLDEN / "jitter.npz"))
    init, dt, bmin, bmax = g["init"], float(g["dt"]), g["box_min"], g["box_max"]
    s = EmuSolver(emu, len(init), g["params"], KEY_LINEAR_CELL, fast_math=False)
    s.upload(init)
    s.step(dt, bmin, bmax)
    p = s.download()
    a = np.argsort(pl.ids_of(p)); b = np.argsort(g["step1_id"])
    assert np.abs(p["position"][a] - g["step1_position"][b]).max() / pl.H < 2e-5
    s.step(dt, bmin, bmax, iters=0)
    q = s.download()
    assert np.isfinite(q["position"]).all() and np.isfinite(q["velocity"]).all()
    # with no constraint iterations the committed position is the (collision-free) prediction of the previous state
    a2 = np.argsort(pl.ids_of(q))
    pred = p["position"][a] + dt * (p["velocity"][a] + dt * np.array(list(s.cfg.gravity), np.float32))
    inside = np.all((pred > bmin + 0.03) & (pred < bmax - 0.03), axis=1)
    assert inside.sum() > len(p) // 2
    assert np.abs(q["position"][a2][inside] - pred[inside]).max() < 1e-5
    s.close()


@pytest.mark.parametrize("n", [3000, 2200013])     # the larger case makes k_mig_scan carry across 1024-tile chunks (2048 particles per tile)
def test_migration_compaction_is_deterministic_and_ordered(emu, n):
    """k_mig_count / k_mig_scan / k_mig_pack: leavers are packed in index order per direction, get the sentinel key, and the
    count message carries the plane populations the single count exchange of a slab step relies on."""
    rng = np.random.default_rng(3)
    gy, gz, gx = 7, 5, 12
    plane = gy * gz
    cx = rng.integers(0, gx, n)
    keys = (cx * plane + rng.integers(0, plane, n)).astype(np.uint32)
    ids = rng.permutation(n).astype(np.uint32)
    x_lo, x_hi = 4, 9
    sentinel = gx * plane
    k2 = keys.copy()
    counts = np.zeros(64, np.uint32)
    ids_l, ids_r = np.zeros(n, np.uint32), np.zeros(n, np.uint32)
    emu.emu_migration(k2.ctypes.data, n, plane, x_lo, x_hi, sentinel, ids.ctypes.data, n, counts.ctypes.data,
                      ids_l.ctypes.data, ids_r.ctypes.data)
    left, right = cx < x_lo, cx >= x_hi
    assert counts[0] == left.sum() and counts[1] == right.sum()
    assert np.array_equal(ids_l[:counts[0]], ids[left]) and np.array_equal(ids_r[:counts[1]], ids[right])   # index order
    assert np.all(k2[left | right] == sentinel) and np.array_equal(k2[~(left | right)], keys[~(left | right)])
    stay = ~(left | right)
    assert counts[2] == (stay & (cx == x_lo)).sum() and counts[3] == (stay & (cx == x_hi - 1)).sum()
    assert counts[4] == (cx == x_lo - 1).sum() and counts[5] == (cx == x_hi).sum()
    # the outgoing count messages: to the left {leavers, landing in its last plane, my first-plane stayers}, mirrored right
    assert list(counts[16:19]) == [counts[0], counts[4], counts[2]] and list(counts[20:23]) == [counts[1], counts[5], counts[3]]


def test_plane_histogram_and_plane_verify(emu):
    rng = np.random.default_rng(4)
    plane, gx, n = 35, 20, 5000
    keys = np.sort((rng.integers(2, 17, n) * plane + rng.integers(0, plane, n)).astype(np.uint32))
    hist, work = np.zeros(gx, np.uint64), np.zeros(gx, np.uint64)
    nbr = rng.integers(0, 60, n).astype(np.uint32)
    emu.emu_plane_hist(keys.ctypes.data, nbr.ctypes.data, n, plane, gx, hist.ctypes.data, work.ctypes.data)
    assert np.array_equal(hist, np.bincount(keys // plane, minlength=gx).astype(np.uint64))
    # the work histogram weighs a particle by 12 + the largest neighbour count among the 32 consecutive particles of its warp
    # (groups of 32 counted from the start of its plane): what akua_pbf_rebalance balances
    want = np.zeros(gx, np.uint64)
    for x in range(gx):
        idx = np.nonzero(keys // plane == x)[0]
        for g0 in range(0, len(idx), 32):
            grp = nbr[idx[g0:g0 + 32]]
            want[x] += len(grp) * (12 + int(grp.max()))
    assert np.array_equal(work, want)
    x_lo, x_hi = 2, 17
    first, last = int((keys // plane == x_lo).sum()), int((keys // plane == x_hi - 1).sum())
    c = np.zeros(64, np.uint32)
    emu.emu_plane_verify(keys.ctypes.data, n, plane, x_lo, x_hi, first, last, c.ctypes.data)
    assert c[31] == 0
    emu.emu_plane_verify(keys.ctypes.data, n, plane, x_lo, x_hi, first + 1, last, c.ctypes.data)
    assert c[31] == 1      # a wrong prediction raises the sticky error word (bit 0)


def test_fuzz_against_the_cpu_port(emu):
    """Differential fuzz of the emulated kernel source (REFERENCE_HASH mode) against the CPU port of the reference on small
    random scenes chosen to hit the edges: 1-3 particles, mixed masses, neighbour caps of 5 / 16 that bite, dense blobs,
    particles outside the box and around the origin (where the reference's hash collides). Integer
    structures must agree bit for bit after every step (each step teacher-forced from the port's state); floats within 2e-4
    except in the dense blobs, whose violent steps amplify last-ulp differences."""
    from oracle import PortOracle, param_block
    rng = np.random.default_rng(0)
    bmin, bmax = np.array([1.5, 0, 1.5], np.float32), np.array([4.5, 4, 4.5], np.float32)
    dt = 0.0083
    for trial in range(14):
        n = int(rng.choice([1, 2, 3, 7, 33, 100, 257, 600]))
        kind = int(rng.integers(0, 4))
        if kind == 0:
            pos = rng.uniform([1.6, 0.1, 1.6], [2.2, 0.6, 2.2], (n, 3))
        elif kind == 1:
            pos = np.array([3, 2, 3]) + rng.normal(0, 0.05, (n, 3))
        elif kind == 2:
            pos = rng.uniform([1.0, -0.5, 1.0], [5.0, 4.5, 5.0], (n, 3))
        else:
            pos = np.array([-0.02, 0.03, 0.01]) + rng.uniform(-0.15, 0.15, (n, 3))
        p = np.zeros(n, PARTICLE_DTYPE)
        p["position"] = pos.astype(np.float32)
        p["velocity"] = rng.normal(0, 0.5, (n, 3)).astype(np.float32)
        p["mass"] = 1.0 if rng.random() < 0.5 else rng.uniform(0.5, 1.5, n).astype(np.float32)
        p["color"][:, 0] = np.arange(n)
        max_n = int(rng.choice([128, 128, 16, 5]))
        params = param_block(maxNeighbours=max_n)
        o = PortOracle(p.copy(), params)
        s = EmuSolver(emu, n, params, KEY_REFERENCE_HASH)
        s.upload(p)
        for step in range(2):
            o.step(dt, bmin, bmax)
            s.step(dt, bmin, bmax)
            got, want = s.download(), o.particles
            arr, cnt = o.neighbours()
            where = (trial, n, kind, max_n, step)
            assert np.array_equal(got["hash"], want["hash"]), where
            assert np.array_equal(pl.ids_of(got), pl.ids_of(want)), where
            assert np.array_equal(s.debug(6), cnt), where
            lst = s.debug(7, (n, max_n))
            mask = np.arange(max_n)[None, :] < cnt[:, None]
            assert np.array_equal(lst[mask], arr[:, :max_n][mask]), where
            if kind != 1 and np.isfinite(want["position"]).all():
                assert np.abs(got["position"] - want["position"]).max() / pl.H < 2e-4, where
                assert np.abs(got["velocity"] - want["velocity"]).max() / (pl.H / dt) < 2e-4, where
            # teacher forcing: the next step starts from the port's state, so that integer structures stay comparable (free
            # running, last-ulp differences flip a neighbour at d == h or move a particle across a cell face)
            s.upload(want.copy())
        s.close(); o.close()


def test_linear_cell_mode_finds_the_same_neighbour_sets(emu):
    """LINEAR_CELL keys change the order of particles and of list entries, never the neighbour SETS (away from the
    reference's origin hash collisions and while the cap does not bite): compare as sets of particle ids with the CPU port."""
    from oracle import PortOracle, param_block
    rng = np.random.default_rng(21)
    bmin, bmax = np.array([1.5, 0, 1.5], np.float32), np.array([4.5, 4, 4.5], np.float32)
    for n, lo, hi in [(400, [1.6, 0.1, 1.6], [2.3, 0.7, 2.3]), (700, [1.0, -0.5, 1.0], [3.0, 1.5, 3.0])]:
        p = np.zeros(n, PARTICLE_DTYPE)
        p["position"] = rng.uniform(lo, hi, (n, 3)).astype(np.float32)
        p["mass"] = 1.0
        p["color"][:, 0] = np.arange(n)
        params = param_block()
        o = PortOracle(p.copy(), params)
        o.step(0.0083, bmin, bmax)
        arr, cnt = o.neighbours()
        assert cnt.max() < 128
        want = pl.neighbour_sets_by_id(arr, cnt, pl.ids_of(o.particles))
        for lb in (0, 1):
            s = EmuSolver(emu, n, params, KEY_LINEAR_CELL, list_build=lb)
            s.upload(p)
            s.step(0.0083, bmin, bmax)
            got_p = s.download()
            got = pl.neighbour_sets_by_id(s.debug(7, (n, 128)), s.debug(6), pl.ids_of(got_p))
            assert got == want, (n, lb)
            a, b = np.argsort(pl.ids_of(got_p)), np.argsort(pl.ids_of(o.particles))
            assert np.abs(got_p["position"][a] - o.particles["position"][b]).max() / pl.H < 2e-5
            s.close()
        o.close()


def test_readme_dam_break_27k_ten_steps(emu):
    """BASELINE config 1 (the README scene: 30^3 particles, dt 0.0083, 4 iterations) through the emulated kernel source,
    against the fixture recorded from the unmodified reference kernels: same tolerances as the GPU test."""
    from akuaengine_b200 import scenes
    g = dict(np.load(GOLDEN / "dambreak27k.npz"))
    init, _, _ = scenes.dam_break(30)
    init["color"][:, 0] = np.arange(len(init), dtype=np.float32)
    dt = float(g["dt"])
    s = EmuSolver(emu, len(init), g["params"], KEY_LINEAR_CELL)
    s.upload(init)
    for k in range(1, 11):
        s.step(dt, g["box_min"], g["box_max"])
        if k in (1, 10):
            p = s.download()
            a = np.argsort(pl.ids_of(p)); b = np.argsort(g[f"step{k}_id"])
            dp = np.abs(p["position"][a].astype(np.float64) - g[f"step{k}_position"][b]).max() / pl.H
            dv = np.abs(p["velocity"][a].astype(np.float64) - g[f"step{k}_velocity"][b]).max() / (pl.H / dt)
            tol = 2e-5 if k == 1 else 1e-3
            assert dp < tol and dv < tol, (k, dp, dv)
    s.close()


@pytest.mark.parametrize("pack", [1, 0], ids=["packed", "plain"])
def test_interior_boundary_split_and_fused_push(emu, pack):
    """Multi-GPU overlap splits every sweep into the slab interior and its two boundary planes (Span) and lets the boundary
    launch store its results straight into the neighbours' ghost regions (PeerPush): the split must compute exactly what one
    launch over the whole range computes, and the pushed planes must be the first / last stretch of the result."""
    g = dict(np.load(GOLDEN / "lattice12.npz"))
    s = EmuSolver(emu, len(g["init"]), g["params"], KEY_LINEAR_CELL, pack=pack)
    s.upload(g["init"])
    s.step(float(g["dt"]), g["box_min"], g["box_max"])
    for plane_l, plane_r in [(0, 0), (100, 250), (1, 1726), (300, 0), (0, 17), (127, 129)]:
        assert emu.emu_span_push_check(s.h, plane_l, plane_r) == 0, (plane_l, plane_r)
    s.close()
