"""The x-slab multi-GPU step on the CPU: the product's HOST solver (akuaengine_b200/csrc/pbf_solver.cu + pbf_slab.inl, the very
source nvcc builds) compiled by g++ against the SIMT emulator and the host-runtime shim of tests/emu/, driven through the real C
ABI and the real Python binding, with ONE OS THREAD PER RANK: peer-to-peer stores are plain stores into the other thread's
arrays, the in-kernel flag waits are real waits, NCCL is an in-process rendezvous (tests/emu/emu_nccl.cpp).

What this covers without a GPU: the device-driven step plan (no host-side sizes), slab-local keys, migration with payload
slots, ghost-plane ranges, the device-resolved interior / boundary spans with fused halo pushes and epoch waits, re-balancing,
the NCCL fallback transport, and the host bookkeeping around them. What it cannot cover: CUDA-graph capture (refused by the
shim; the solver then steps eagerly), real concurrency between kernels of one device, timing. `-m gpu` tests and
tests/mgpu_worker.py do that on real GPUs. Test infrastructure only: nothing under akuaengine_b200/ can load this build.
"""
import os
import subprocess
import threading
from pathlib import Path

import numpy as np
import pytest

from akuaengine_b200 import PARTICLE_DTYPE, PBFSolver, load_library, scenes
from akuaengine_b200.slab import partition_columns, x_columns

REPO = Path(__file__).resolve().parents[1]
EMU = REPO / "tests" / "emu"
OUT = EMU / "_build" / "libakua_pbf_emu.so"
CSRC = REPO / "akuaengine_b200" / "csrc"
DEPS = [EMU / "emu_core.cpp", EMU / "emu_nccl.cpp", EMU / "cuda_runtime.h", EMU / "cuda_host_shim.h", EMU / "nccl.h",
        *CSRC.glob("*.cuh"), *CSRC.glob("*.h"), *CSRC.glob("*.inl"), CSRC / "pbf_solver.cu", REPO / "include" / "akua_pbf.h"]
H, DT = 0.1, 0.0083
# a rank that stops (error) must not hang its neighbours for the emulator's lifetime: bounded waits, short for the tests. Set by
# the fixture (not at import: collecting this module on a GPU box must not shorten the real library's time-outs, which the
# multi-GPU tests' worker processes would inherit) and read once by the emulated library at its first wait.
EMU_WAIT_CYCLES = "30000000"


@pytest.fixture(scope="module")
def emulib():
    os.environ.setdefault("AKUA_SLAB_WAIT_CYCLES", EMU_WAIT_CYCLES)
    if not OUT.exists() or any(p.stat().st_mtime > OUT.stat().st_mtime for p in DEPS):
        OUT.parent.mkdir(parents=True, exist_ok=True)
        cmd = ["g++", "-O1", "-std=c++17", "-DAKUA_HOST_EMU", "-U_FORTIFY_SOURCE", "-ffp-contract=off", "-fPIC", "-shared", "-pthread",
               "-I", str(EMU), "-x", "c++", str(CSRC / "pbf_solver.cu"), "-x", "c++", str(EMU / "emu_core.cpp"), str(EMU / "emu_nccl.cpp"),
               "-o", str(OUT), "-ldl"]
        res = subprocess.run(cmd, capture_output=True, text=True)
        assert res.returncode == 0, res.stderr[-4000:]
    return load_library(OUT)


def _scene(nx=24, ny=8, nz=10, vx=0.0):
    p, bmin, bmax = scenes.tank(nx, ny, nz)
    n = len(p)
    p["velocity"][:, 0] = np.float32(vx)
    p["color"][:, 0] = (np.arange(n) % 251).astype(np.float32)
    p["size"] = (np.arange(n) % 17 + 1).astype(np.float32)
    return p, bmin, bmax


def _run_single(lib, p, bmin, bmax, steps, g, canonical=False):
    s = PBFSolver(len(p), lib=lib, use_graph=False, canonical_order=canonical)
    s.upload_particles(p)
    s.setGravity(g)
    for _ in range(steps):
        s.step(DT, bmin, bmax)
    pos, vel, pid = s.download()
    err = s.density_error()
    s.close()
    return pos, vel, pid, err


def _run_slab(lib, p, bmin, bmax, steps, g, world, skew=0.0, rebalance_every=0, capacity_factor=4.0, canonical=False,
              rebalance_async=False):
    """One thread per rank. Returns per-rank (pos4, vel4, ids, aos, stats) and raises the first exception of any rank."""
    n = len(p)
    ids = np.arange(n, dtype=np.uint32)
    cols = x_columns(p["position"][:, 0], H)
    col_min = int(cols.min())
    hist = np.bincount(cols - col_min).astype(np.int64)
    bounds = partition_columns(hist, world)
    if skew:
        for r in range(1, world):
            bounds[r] = max(2 * r, int(bounds[r] * (1.0 - skew)))
    uid = PBFSolver.comm_unique_id(lib)
    out, errs = [None] * world, []

    def work(rank):
        try:
            lo, hi = col_min + int(bounds[rank]), col_min + int(bounds[rank + 1])
            mine = (cols >= lo) & (cols < hi)
            own = np.ascontiguousarray(p[mine])
            s = PBFSolver(n if skew else max(n // world, len(own)), lib=lib, use_graph=False, capacity_factor=capacity_factor, device=rank,
                          canonical_order=canonical)
            s.comm_init(rank, world, uid)
            s.set_slab(lo, hi)
            s.upload_particles(own)
            s.upload_ids(ids[mine])
            s.setGravity(g)
            for k in range(steps):
                s.step(DT, bmin, bmax)
                if rebalance_every and (k + 1) % rebalance_every == 0:
                    s.rebalance_async() if rebalance_async else s.rebalance()
            pos, vel, pid = s.download()
            aos = s.download_particles()
            st = s.slab_stats()
            st["waits"] = s.slab_wait_stats()
            err = s.density_error()
            out[rank] = (pos, vel, pid, aos, st, err)
            s.close()
        except Exception as e:  # noqa: BLE001 - re-raised in the main thread
            errs.append((rank, e))

    threads = [threading.Thread(target=work, args=(r,)) for r in range(world)]
    for t in threads:
        t.start()
    for t in threads:
        t.join(timeout=900)
    assert not any(t.is_alive() for t in threads), "a rank is stuck (deadlock between emulated ranks)"
    if errs:
        raise AssertionError(f"rank {errs[0][0]} failed: {errs[0][1]}")
    return out


def _compare(p, single, slab, tol):
    n = len(p)
    sp, sv, sid, _ = single
    pos = np.concatenate([o[0] for o in slab]); vel = np.concatenate([o[1] for o in slab]); pid = np.concatenate([o[2] for o in slab])
    aos = np.concatenate([o[3] for o in slab])
    assert len(pid) == n and np.array_equal(np.sort(pid), np.arange(n)), "particles not conserved across the ranks"
    a, b = np.argsort(pid), np.argsort(sid)
    dp = np.abs(pos[a, :3] - sp[b, :3]).max() / H
    dv = np.abs(vel[a, :3] - sv[b, :3]).max() / (H / DT)
    assert dp < tol and dv < tol, (dp, dv)
    # the render payload follows its particle through sorts and migrations (reference: the struct is sorted as a whole)
    assert np.array_equal(aos["color"][a, 0], p["color"][:, 0]) and np.array_equal(aos["size"][a], p["size"])
    assert np.array_equal(aos["position"][a], pos[a, :3])
    return dp, dv


@pytest.mark.parametrize("world", [2, 3])
def test_slab_without_migration_is_bit_identical_to_one_rank(emulib, world):
    p, bmin, bmax = _scene()
    g = np.array([0.0, -9.8, 0.0], np.float32)
    single = _run_single(emulib, p, bmin, bmax, 3, g)
    slab = _run_slab(emulib, p, bmin, bmax, 3, g, world)
    migrated = sum(o[4]["migrated_in"] for o in slab)
    dp, dv = _compare(p, single, slab, 1e-6 if migrated == 0 else 1e-3)
    if migrated == 0:
        assert dp == 0.0 and dv == 0.0      # same sort order, same list order: no float may differ
    assert all(o[4]["transport"] == "cuda-ipc p2p" and o[4]["exchanges"] > 0 for o in slab)
    # the device-side wait statistics come through the ABI (the emulator's device clock stands still: zero nanoseconds, 3 steps)
    assert all(o[4]["waits"]["device_steps"] == 3 and o[4]["waits"]["plan_wait_ns"] == 0 for o in slab)
    assert sum(o[4]["bytes_sent"] for o in slab) > 0


def test_slab_with_migration_matches_one_rank(emulib):
    p, bmin, bmax = _scene(vx=2.0)
    g = scenes.tank_gravity(15.0)
    steps = 10
    single = _run_single(emulib, p, bmin, bmax, steps, g)
    slab = _run_slab(emulib, p, bmin, bmax, steps, g, 3)
    assert sum(o[4]["migrated_in"] for o in slab) > 0, "the scene was meant to migrate particles"
    _compare(p, single, slab, 1e-3)
    # global density-constraint error agrees with the single-rank run
    tot = sum(o[5][0] * len(o[2]) for o in slab) / len(p)
    assert abs(tot - single[3][0]) < 1e-3 * max(single[3][0], 1e-3) + 1e-5


@pytest.mark.parametrize("world,skew,rebalance_every", [(3, 0.0, 0), (2, 0.5, 2)])
def test_canonical_order_makes_the_slab_run_bit_identical_with_migration(emulib, world, skew, rebalance_every):
    """options.canonical_order: a cell's particles are ordered by id on every rank, so neighbour order — and every float sum —
    is that of the single-rank run, whatever migrated and however the slabs were re-balanced."""
    p, bmin, bmax = _scene(nx=32 if skew else 24, vx=2.0)
    g = scenes.tank_gravity(15.0)
    steps = 10
    single = _run_single(emulib, p, bmin, bmax, steps, g, canonical=True)
    slab = _run_slab(emulib, p, bmin, bmax, steps, g, world, skew=skew, rebalance_every=rebalance_every,
                     capacity_factor=2.5 if skew else 4.0, canonical=True)
    assert sum(o[4]["migrated_in"] for o in slab) > 0, "the scene was meant to migrate particles"
    dp, dv = _compare(p, single, slab, 1e-6)
    assert dp == 0.0 and dv == 0.0
    # and it is a different order from the default one (otherwise the option would be vacuous)
    plain = _run_single(emulib, p, bmin, bmax, steps, g, canonical=False)
    assert np.array_equal(np.sort(plain[2]), np.sort(single[2]))


def test_slab_rebalances_from_a_skewed_start(emulib):
    p, bmin, bmax = _scene(nx=32)
    g = np.array([0.0, -9.8, 0.0], np.float32)
    steps = 12
    single = _run_single(emulib, p, bmin, bmax, steps, g)
    slab = _run_slab(emulib, p, bmin, bmax, steps, g, 2, skew=0.5, rebalance_every=2, capacity_factor=2.5)
    _compare(p, single, slab, 1e-3)
    owned = [len(o[2]) for o in slab]
    assert max(owned) / (sum(owned) / 2) < 1.2, owned


def test_async_rebalance_lags_one_call_and_stays_bit_identical(emulib):
    """akua_pbf_rebalance_async applies the previous call's measurement (no host synchronisation): from a skewed start the slabs
    still balance, one call later, and in canonical order the run stays bit-identical to one rank."""
    p, bmin, bmax = _scene(nx=32)
    g = np.array([0.0, -9.8, 0.0], np.float32)
    steps = 14
    single = _run_single(emulib, p, bmin, bmax, steps, g, canonical=True)
    slab = _run_slab(emulib, p, bmin, bmax, steps, g, 2, skew=0.5, rebalance_every=2, capacity_factor=2.5, canonical=True,
                     rebalance_async=True)
    dp, dv = _compare(p, single, slab, 1e-6)
    assert dp == 0.0 and dv == 0.0
    owned = [len(o[2]) for o in slab]
    assert max(owned) / (sum(owned) / 2) < 1.2, owned


def test_a_dead_neighbour_costs_one_timeout_not_one_per_kernel(emulib):
    """Rank 1 stops stepping. Rank 0's next step waits for its count message, times out ONCE (bounded in-kernel spin), raises the
    sticky error word, and every later wait of that rank gives up at once: the queue drains, the next host call reports
    AKUA_ERR_COMM, and closing the solver does not hang. (On a GPU the same logic turns a lost peer into an error within ~10 s
    instead of one time-out per queued kernel.)"""
    import time
    from akuaengine_b200 import AkuaError
    p, bmin, bmax = _scene()
    g = np.array([0.0, -9.8, 0.0], np.float32)
    n, world = len(p), 2
    ids = np.arange(n, dtype=np.uint32)
    cols = x_columns(p["position"][:, 0], H)
    col_min = int(cols.min())
    bounds = partition_columns(np.bincount(cols - col_min).astype(np.int64), world)
    uid = PBFSolver.comm_unique_id(emulib)
    res, errs = {}, []
    gate = threading.Event()

    def work(rank):
        try:
            lo, hi = col_min + int(bounds[rank]), col_min + int(bounds[rank + 1])
            mine = (cols >= lo) & (cols < hi)
            s = PBFSolver(n // world, lib=emulib, use_graph=False, capacity_factor=4.0)
            s.comm_init(rank, world, uid)
            s.set_slab(lo, hi)
            s.upload_particles(np.ascontiguousarray(p[mine]))
            s.upload_ids(ids[mine])
            s.setGravity(g)
            for _ in range(2):
                s.step(DT, bmin, bmax)
            s.sync()
            if rank == 1:
                gate.wait(timeout=600)       # the "dead" rank: alive (its memory stays mapped) but silent
                s.close()
                return
            t0 = time.perf_counter()
            failed = None
            try:
                for _ in range(3):
                    s.step(DT, bmin, bmax)
                s.sync()
            except AkuaError as e:
                failed = str(e)
            res["elapsed"] = time.perf_counter() - t0
            res["error"] = failed
            t1 = time.perf_counter()
            s.close()
            res["close"] = time.perf_counter() - t1
        except Exception as e:  # noqa: BLE001
            errs.append((rank, e))
        finally:
            if rank == 0:
                gate.set()

    ts = [threading.Thread(target=work, args=(r,)) for r in range(world)]
    [t.start() for t in ts]
    [t.join(timeout=600) for t in ts]
    assert not any(t.is_alive() for t in ts) and not errs, errs
    assert res["error"] and "timed out" in res["error"], res
    # three steps = dozens of waits: with one time-out per wait this would take tens of times longer than the first one did
    assert res["close"] < max(2.0, res["elapsed"]), res


def test_slab_nccl_fallback_transport(emulib, monkeypatch):
    monkeypatch.setenv("AKUA_SLAB_P2P", "0")
    p, bmin, bmax = _scene(vx=1.5)
    g = np.array([0.0, -9.8, 0.0], np.float32)
    single = _run_single(emulib, p, bmin, bmax, 6, g)
    slab = _run_slab(emulib, p, bmin, bmax, 6, g, 2)
    assert all(o[4]["transport"] == "nccl" for o in slab)
    _compare(p, single, slab, 1e-3)


def test_zero_iterations_and_upload_roundtrip_in_slab_mode(emulib):
    """solverIterations == 0 takes the stand-alone commit + its own v exchange; an AoS download / upload round trip between
    steps (what bench.py's e2e loop does) must not disturb the run."""
    p, bmin, bmax = _scene()
    g = np.array([0.0, -9.8, 0.0], np.float32)
    n, world = len(p), 2
    ids = np.arange(n, dtype=np.uint32)
    cols = x_columns(p["position"][:, 0], H)
    col_min = int(cols.min())
    bounds = partition_columns(np.bincount(cols - col_min).astype(np.int64), world)
    uid = PBFSolver.comm_unique_id(emulib)
    res, errs = [None] * world, []

    def work(rank):
        try:
            lo, hi = col_min + int(bounds[rank]), col_min + int(bounds[rank + 1])
            mine = (cols >= lo) & (cols < hi)
            s = PBFSolver(n // world, lib=emulib, use_graph=False, capacity_factor=4.0)
            s.comm_init(rank, world, uid)
            s.set_slab(lo, hi)
            s.upload_particles(np.ascontiguousarray(p[mine]))
            s.upload_ids(ids[mine])
            s.step(DT, bmin, bmax)                       # densities exist from here on (XSPH divides by them)
            s.step(DT, bmin, bmax, solverIterations=0)
            for _ in range(2):
                aos = s.download_particles()
                pid = s.debug(3)
                s.upload_particles(aos)
                s.upload_ids(pid)
                s.step(DT, bmin, bmax)
            res[rank] = (s.download(), s.download_particles())
            s.close()
        except Exception as e:  # noqa: BLE001
            errs.append((rank, e))

    ts = [threading.Thread(target=work, args=(r,)) for r in range(world)]
    [t.start() for t in ts]
    [t.join(timeout=600) for t in ts]
    assert not errs, errs
    ref = PBFSolver(n, lib=emulib, use_graph=False)
    ref.upload_particles(p)
    ref.step(DT, bmin, bmax)
    ref.step(DT, bmin, bmax, solverIterations=0)
    for _ in range(2):
        ref.step(DT, bmin, bmax)
    rp, rv, rid = ref.download()
    ref.close()
    pos = np.concatenate([r[0][0] for r in res]); pid = np.concatenate([r[0][2] for r in res])
    assert np.array_equal(np.sort(pid), np.arange(n))
    a, b = np.argsort(pid), np.argsort(rid)
    assert np.abs(pos[a, :3] - rp[b, :3]).max() / H < 1e-3
