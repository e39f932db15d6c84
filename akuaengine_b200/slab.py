"""x-slab multi-GPU driver glue (one process per GPU, torch.distributed for the rendezvous only).

The data path — migration, ghost-plane exchange over NCCL/NVLink — lives in the CUDA library (csrc/pbf_slab.inl). This
module only (1) chooses balanced slab boundaries from an all-reduced x-column histogram (host code in the library:
akua_slab_partition), (2) distributes the NCCL unique id, and (3) uploads each rank's share.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import PARTICLE_DTYPE, PBFSolver, load_library

INT_MIN, INT_MAX = -(2 ** 31), 2 ** 31 - 1


def x_columns(pos_x: np.ndarray, h: float) -> np.ndarray:
    """Absolute x cell of each particle, exactly as the kernels compute it: floor of an IEEE float32 division."""
    return np.floor(pos_x.astype(np.float32) / np.float32(h)).astype(np.int64)


def partition_columns(hist: np.ndarray, nranks: int) -> np.ndarray:
    """Balanced contiguous column intervals; returns nranks + 1 indices into `hist` (library host code, no GPU needed)."""
    hist = np.ascontiguousarray(hist, dtype=np.int64)
    bounds = np.zeros(nranks + 1, np.int32)
    rc = load_library().akua_slab_partition(hist.ctypes.data_as(C.POINTER(C.c_int64)), len(hist), nranks,
                                            bounds.ctypes.data_as(C.POINTER(C.c_int32)))
    if rc != 0:
        raise ValueError(f"akua_slab_partition failed (status {rc}): need at least one column per rank")
    return bounds


def rebalance_bounds(hist: np.ndarray, old_bounds: np.ndarray, max_move: int) -> np.ndarray:
    """Boundaries akua_pbf_rebalance would adopt for the global histogram `hist` and the current `old_bounds` when at most
    `max_move` particles may cross a boundary in one step (library host code, no GPU needed)."""
    hist = np.ascontiguousarray(hist, dtype=np.int64)
    old = np.ascontiguousarray(old_bounds, dtype=np.int32)
    bounds = np.zeros(len(old), np.int32)
    rc = load_library().akua_slab_rebalance_bounds(hist.ctypes.data_as(C.POINTER(C.c_int64)), len(hist), len(old) - 1,
                                                   old.ctypes.data_as(C.POINTER(C.c_int32)), int(max_move),
                                                   bounds.ctypes.data_as(C.POINTER(C.c_int32)))
    if rc != 0:
        raise ValueError(f"akua_slab_rebalance_bounds failed (status {rc})")
    return bounds


def global_histogram(cols_local: np.ndarray, dist=None):
    """(col_min, histogram over [col_min, col_max]) of all ranks' particles. `dist` = torch.distributed or None."""
    lo = int(cols_local.min()) if len(cols_local) else INT_MAX
    hi = int(cols_local.max()) if len(cols_local) else INT_MIN
    if dist is not None and dist.is_initialized() and dist.get_world_size() > 1:
        import torch
        dev = "cuda" if dist.get_backend() == "nccl" else "cpu"
        t = torch.tensor([-lo, hi], dtype=torch.int64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        lo, hi = -int(t[0]), int(t[1])
        hist = torch.from_numpy(np.bincount(cols_local - lo, minlength=hi - lo + 1).astype(np.int64)).to(dev)
        dist.all_reduce(hist)
        return lo, hist.cpu().numpy()
    return lo, np.bincount(cols_local - lo, minlength=hi - lo + 1).astype(np.int64)


def slab_interval(rank: int, nranks: int, col_min: int, bounds: np.ndarray):
    """Absolute x-cell interval [lo, hi) of a rank."""
    return col_min + int(bounds[rank]), col_min + int(bounds[rank + 1])


def broadcast_unique_id(dist, rank: int) -> np.ndarray:
    import torch
    dev = "cuda" if dist.get_backend() == "nccl" else "cpu"
    uid = PBFSolver.comm_unique_id() if rank == 0 else np.zeros(128, np.uint8)
    t = torch.from_numpy(uid).to(dev)
    dist.broadcast(t, src=0)
    return t.cpu().numpy()


def setup_slab_solver(particles: np.ndarray, ids: np.ndarray, dist, rank: int, nranks: int, device: int, h: float = 0.1,
                      capacity_factor: float = 1.5, skew: float = 0.0, **solver_kw) -> PBFSolver:
    """`particles` / `ids`: any subset of the scene held by this rank (e.g. a 1/nranks share, or everything on every rank
    with `ids` selecting a share); the routine keeps what falls into this rank's slab. Every particle of the scene must be
    held by exactly one rank whose slab it falls into — the simplest way is for every rank to pass the whole scene."""
    assert particles.dtype == PARTICLE_DTYPE
    cols = x_columns(particles["position"][:, 0], h)
    col_min, hist = global_histogram(cols, None)  # callers pass the whole scene: the histogram is already global
    bounds = partition_columns(hist, nranks)
    if skew:  # deliberately unbalanced start (tests of akua_pbf_rebalance): interior boundaries pulled towards column 0
        for r in range(1, nranks):
            bounds[r] = max(2 * r, int(bounds[r] * (1.0 - skew)))   # slabs stay at least two planes wide
    lo, hi = slab_interval(rank, nranks, col_min, bounds)
    mine = (cols >= lo) & (cols < hi)
    if rank == 0:
        mine |= cols < lo
    if rank == nranks - 1:
        mine |= cols >= hi
    own = np.ascontiguousarray(particles[mine])
    cap_n = max(int(len(particles) / nranks), len(own), 1) if not skew else len(particles)
    solver = PBFSolver(cap_n, device=device, capacity_factor=capacity_factor, **solver_kw)
    if nranks > 1:
        solver.comm_init(rank, nranks, broadcast_unique_id(dist, rank))
        solver.set_slab(lo, hi)
    solver.upload_particles(own)
    solver.upload_ids(np.ascontiguousarray(ids[mine], dtype=np.uint32))
    return solver


def slab_selfcheck(dist, rank: int, nranks: int, device: int, steps: int = 24, dt: float = 0.0083, h: float = 0.1,
                   vx: float = 1.0, rebalance_every: int = 8) -> dict:
    """Runs a small tank scene (32*nranks x 24 x 32 lattice, tilted gravity, initial x velocity so that particles migrate,
    re-balanced every `rebalance_every` steps) on all ranks through the x-slab path and on rank 0 through the single-GPU path,
    and compares them particle by particle (matched by id). Collective; the same verdict dict on every rank.

    Two passes. (1) options.canonical_order = 1 on both sides: particles of a cell are ordered by id, so neighbour order and
    with it every float sum is independent of how the particles are distributed — the N-GPU result must be BIT-IDENTICAL to
    the single-GPU result, migration and re-balancing included. (2) the default order (a cell keeps last step's order, like
    the reference's stable sort; migrants are appended): neighbour ORDER then differs between 1 and N GPUs, float sums differ in
    the last bit, and the scene amplifies that — a statistical bound only (99 % of the particles within 1e-3 h, rms within
    5e-4 h, nobody further than half a cell, equal density-constraint error)."""
    import torch
    from . import scenes
    particles, bmin, bmax = scenes.tank(32 * nranks, 24, 32)
    particles["velocity"][:, 0] = np.float32(vx)
    # payload that must follow its particle through sorts and migrations (reference: the struct is sorted as a whole)
    n = len(particles)
    particles["color"][:, 0] = (np.arange(n) % 251).astype(np.float32)
    particles["size"] = (np.arange(n) % 17 + 1).astype(np.float32)
    ids = np.arange(n, dtype=np.uint32)
    g = scenes.tank_gravity(15.0)

    def one_pass(canonical: bool) -> dict:
        solver = setup_slab_solver(particles, ids, dist, rank, nranks, device, h, capacity_factor=2.0, canonical_order=canonical)
        solver.setGravity(g)
        for k in range(steps):
            solver.step(dt, bmin, bmax)
            if rebalance_every and nranks > 1 and (k + 1) % rebalance_every == 0:
                solver.rebalance()
        pos4, vel4, pid = solver.download()
        aos = solver.download_particles()
        st = solver.slab_stats()
        merr, _ = solver.density_error()
        m = solver.n
        solver.close()
        owned = torch.tensor([m, st["migrated_in"]], device="cuda", dtype=torch.int64)
        all_owned = [torch.zeros_like(owned) for _ in range(nranks)]
        dist.all_gather(all_owned, owned)
        counts = [int(t[0]) for t in all_owned]
        migrated = sum(int(t[1]) for t in all_owned)
        mx = max(counts)
        pack = torch.zeros((mx, 12), dtype=torch.float64, device="cuda")
        pack[:m, 0:3] = torch.from_numpy(pos4[:, :3].astype(np.float64)).cuda()
        pack[:m, 3:7] = torch.from_numpy(vel4.astype(np.float64)).cuda()
        pack[:m, 7] = torch.from_numpy(pid.astype(np.float64)).cuda()
        pack[:m, 8] = torch.from_numpy(aos["color"][:, 0].astype(np.float64)).cuda()
        pack[:m, 9] = torch.from_numpy(aos["size"].astype(np.float64)).cuda()
        gathered = [torch.zeros_like(pack) for _ in range(nranks)]
        dist.all_gather(gathered, pack)
        verdict = torch.zeros(9, dtype=torch.float64, device="cuda")
        if rank == 0:
            allp = np.concatenate([t[:c].cpu().numpy() for t, c in zip(gathered, counts)])
            got_ids = allp[:, 7].astype(np.int64)
            conserved = len(got_ids) == n and np.array_equal(np.sort(got_ids), np.arange(n))
            ref = PBFSolver(n, device=device, canonical_order=canonical)
            ref.upload_particles(particles)
            ref.setGravity(g)
            for _ in range(steps):
                ref.step(dt, bmin, bmax)
            rp, rv, rid = ref.download()
            rerr, _ = ref.density_error()
            ref.close()
            dp = dv = payload_ok = p99 = rms = drho = float("nan")
            if conserved:
                o1, o2 = np.argsort(got_ids), np.argsort(rid)
                d = np.abs(allp[o1, 0:3] - rp[o2, :3]).max(axis=1) / h
                dp, p99, rms = float(d.max()), float(np.quantile(d, 0.99)), float(np.sqrt((d * d).mean()))
                dv = float(np.abs(allp[o1, 3:6] - rv[o2, :3]).max() / (h / dt))
                drho = float(np.abs(allp[o1, 6] - rv[o2, 3]).max())
                payload_ok = float(np.array_equal(allp[o1, 8], particles["color"][:, 0].astype(np.float64))
                                   and np.array_equal(allp[o1, 9], particles["size"].astype(np.float64)))
            verdict = torch.tensor([float(conserved), dp, dv, rerr, float(migrated), payload_ok, p99, rms, drho],
                                   dtype=torch.float64, device="cuda")
        dist.broadcast(verdict, src=0)
        errs = torch.tensor([merr * m, float(m)], device="cuda", dtype=torch.float64)
        dist.all_reduce(errs)
        v = verdict.cpu().numpy()
        slab_err = float(errs[0] / max(float(errs[1]), 1.0))
        out = {"ids_conserved": bool(v[0]), "payload_follows_particles": bool(v[5] == 1.0),
               "max_dpos_over_h": float(v[1]), "p99_dpos_over_h": float(v[6]), "rms_dpos_over_h": float(v[7]),
               "max_dvel_over_h_dt": float(v[2]), "max_ddensity": float(v[8]),
               "density_error_mean_slab": slab_err, "density_error_mean_single_gpu": float(v[3]),
               "migrated": int(v[4]), "owned_per_rank": counts, "transport": st["transport"], "rebalances": st.get("rebalances")}
        base = out["ids_conserved"] and out["payload_follows_particles"]
        if canonical:
            out["ok"] = base and out["max_dpos_over_h"] == 0.0 and out["max_dvel_over_h_dt"] == 0.0 and out["max_ddensity"] == 0.0
        else:
            err_ok = abs(slab_err - float(v[3])) <= 1e-3 * max(float(v[3]), 1e-6)
            out["ok"] = base and err_ok and out["p99_dpos_over_h"] < 5e-3 and out["rms_dpos_over_h"] < 2e-3 and out["max_dpos_over_h"] < 0.5
        return out

    canon = one_pass(True)
    default = one_pass(False)
    return {"particles": n, "steps": steps, "ranks": nranks, "rebalance_every": rebalance_every,
            "canonical_order_bit_identical_to_single_gpu": canon, "default_order_statistical": default,
            "ok": bool(canon["ok"] and default["ok"])}
