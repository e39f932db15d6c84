"""CPU check of the two-phase ("mask") neighbour-list build (akuaengine_b200/csrc/list_build.cuh).

The per-particle function is __host__ __device__; tests/cpp/list_build_host.cu runs it on the CPU next to an independent
one-candidate-at-a-time scan (the order kernel_find_neighbours visits candidates in, src/CUDA/NeighbourSearchCUDA.cu:72-130).
Lists must agree bit for bit, including where the maxNeighbours cap bites, rows longer than one 32-candidate chunk,
particles clamped into border cells and the padding of the last group of four. This is test infrastructure: the product
runs the function only inside k_build_neighbours_mask on the GPU.
"""
import ctypes as C
import shutil
import subprocess
from pathlib import Path

import numpy as np
import pytest

REPO = Path(__file__).resolve().parents[1]
SRC = REPO / "tests" / "cpp" / "list_build_host.cu"
OUT = REPO / "tests" / "cpp" / "_build" / "liblist_build_host.so"
DEPS = [SRC, REPO / "akuaengine_b200" / "csrc" / "list_build.cuh", REPO / "akuaengine_b200" / "csrc" / "pbf_kernels.cuh"]


@pytest.fixture(scope="module")
def host_lib():
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not Path(nvcc).exists():
        pytest.skip("nvcc not available")
    if not OUT.exists() or any(p.stat().st_mtime > OUT.stat().st_mtime for p in DEPS):
        OUT.parent.mkdir(parents=True, exist_ok=True)
        cmd = [nvcc, "-O2", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-Xcompiler", "-fPIC,-ffp-contract=off",
               "-shared", "-o", str(OUT), str(SRC)]
        res = subprocess.run(cmd, capture_output=True, text=True)
        assert res.returncode == 0, res.stdout + res.stderr
    lib = C.CDLL(str(OUT))
    fn = lib.akua_test_list_build_host
    fn.restype = C.c_int
    fn.argtypes = [C.c_void_p, C.c_uint32, C.c_void_p, C.c_void_p, C.c_float, C.c_uint32, C.c_int, C.c_void_p, C.c_void_p,
                   C.c_void_p]
    return fn


def _run(fn, xyz, bmin, bmax, h, max_n, variant):
    xyz = np.ascontiguousarray(xyz, np.float32)
    n = len(xyz)
    bmin = np.asarray(bmin, np.float32); bmax = np.asarray(bmax, np.float32)
    order = np.empty(n, np.uint32); lst = np.empty((n, max_n), np.uint32); cnt = np.empty(n, np.uint32)
    rc = fn(xyz.ctypes.data, n, bmin.ctypes.data, bmax.ctypes.data, h, max_n, variant, order.ctypes.data, lst.ctypes.data,
            cnt.ctypes.data)
    assert rc == 0, f"host harness returned {rc} (variant {variant})"
    return order, lst, cnt


def _lattice(nx, ny, nz, origin, spacing=0.05):
    # Application.cpp:173-181: fl(origin + fl(i * spacing)), x outer / z inner
    i, j, k = np.meshgrid(np.arange(nx), np.arange(ny), np.arange(nz), indexing="ij")
    p = np.stack([i, j, k], -1).reshape(-1, 3).astype(np.float32) * np.float32(spacing)
    return (p + np.asarray(origin, np.float32)).astype(np.float32)


def _scenes():
    rng = np.random.default_rng(7)
    box = ([1.5, 0.0, 1.5], [4.5, 4.0, 4.5])
    out = {}
    out["lattice 14^3 (ties at d == h)"] = (_lattice(14, 14, 14, [2.0, 1.0, 2.0]), *box, 128)
    jit = _lattice(12, 10, 11, [1.52, 0.02, 1.52]) + rng.uniform(-0.02, 0.02, (12 * 10 * 11, 3)).astype(np.float32)
    out["jittered block in a box corner"] = (jit, *box, 128)
    out["jittered block, cap 16 bites"] = (jit, *box, 16)
    out["jittered block, cap 5 (not a multiple of four)"] = (jit, *box, 5)
    blob = (np.array([3.0, 2.0, 3.0]) + rng.normal(0, 0.06, (3000, 3))).astype(np.float32)
    out["dense blob: rows longer than 32 candidates, cap 128 bites"] = (blob, *box, 128)
    out["dense blob, cap 40"] = (blob, *box, 40)
    one_cell = (np.array([3.01, 2.01, 3.01]) + rng.uniform(0, 0.08, (150, 3))).astype(np.float32)
    out["150 particles in one cell"] = (one_cell, *box, 128)
    outside = rng.uniform([0.9, -0.6, 0.9], [5.1, 4.6, 5.1], (4000, 3)).astype(np.float32)
    out["particles beyond the grid margin (clamped into border cells)"] = (outside, *box, 128)
    far = (np.array([0.8, -0.5, 4.9]) + rng.normal(0, 0.08, (800, 3))).astype(np.float32)
    out["blob wholly outside the grid (everything clamped into border cells)"] = (far, *box, 128)
    out["single particle"] = (np.array([[2.0, 1.0, 2.0]], np.float32), *box, 128)
    out["two coincident particles"] = (np.array([[2.0, 1.0, 2.0], [2.0, 1.0, 2.0]], np.float32), *box, 128)
    sparse = rng.uniform([1.5, 0, 1.5], [4.5, 4, 4.5], (500, 3)).astype(np.float32)
    out["sparse cloud (most rows empty)"] = (sparse, *box, 128)
    return out


SCENES = _scenes()


@pytest.mark.parametrize("name", list(SCENES))
@pytest.mark.parametrize("variant", [1, 2])
def test_mask_variants_match_the_scan(host_lib, name, variant):
    xyz, bmin, bmax, max_n = SCENES[name]
    o0, l0, c0 = _run(host_lib, xyz, bmin, bmax, 0.1, max_n, 0)
    o1, l1, c1 = _run(host_lib, xyz, bmin, bmax, 0.1, max_n, variant)
    assert np.array_equal(o0, o1)
    assert np.array_equal(c0, c1), f"neighbour counts differ: {np.flatnonzero(c0 != c1)[:8]}"
    assert np.array_equal(l0, l1), "neighbour lists differ"
    assert c0.max() <= max_n


def test_scan_agrees_with_brute_force(host_lib):
    """Pins the harness's own scan: away from the d2 == h2 boundary the neighbour SET is the brute-force set."""
    xyz, bmin, bmax, max_n = SCENES["jittered block in a box corner"]
    order, lst, cnt = _run(host_lib, xyz, bmin, bmax, 0.1, max_n, 0)
    p = xyz[order].astype(np.float64)
    d2 = ((p[:, None, :] - p[None, :, :]) ** 2).sum(-1)
    h2 = np.float64(np.float32(0.1) * np.float32(0.1))
    sure_in = d2 < h2 * (1 - 1e-5)
    sure_out = d2 > h2 * (1 + 1e-5)
    np.fill_diagonal(sure_in, False)
    n = len(p)
    member = np.zeros((n, n), bool)
    rows = np.repeat(np.arange(n), cnt)
    cols = np.concatenate([lst[i, :cnt[i]] for i in range(n)])
    member[rows, cols] = True
    assert not member.diagonal().any(), "a particle lists itself"
    assert (member | ~sure_in).all(), "a certain neighbour is missing"
    assert not (member & sure_out).any(), "a certain non-neighbour is listed"
    assert cnt.max() < max_n      # the cap did not bite, so the sets are complete


def test_cap_truncates_in_scan_order(host_lib):
    """With a smaller cap the list is a PREFIX of the uncapped list (the reference stops scanning when the cap is hit)."""
    xyz, bmin, bmax, _ = SCENES["dense blob, cap 40"]
    for variant in (0, 1, 2):
        _, full, cf = _run(host_lib, xyz, bmin, bmax, 0.1, 512, variant)
        _, cut, cc = _run(host_lib, xyz, bmin, bmax, 0.1, 40, variant)
        assert np.array_equal(cc, np.minimum(cf, 40))
        assert np.array_equal(cut, full[:, :40])
