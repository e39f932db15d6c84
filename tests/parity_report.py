"""TEST INFRASTRUCTURE (imports oracle/). Diagnostic run for a GPU box: golden generation, port-vs-reference check, CUDA-vs-reference parity, quick timings.
Usage (under gpurun): python tests/parity_report.py [--golden] [--timing]"""
import json
import sys
import time
from pathlib import Path

import numpy as np

REPO = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(REPO))
sys.path.insert(0, str(REPO / "tests"))

from akuaengine_b200 import KEY_LINEAR_CELL, KEY_REFERENCE_HASH, PBFSolver, scenes  # noqa: E402
import parity_lib as pl  # noqa: E402

OUT = REPO / "gpurun_out"
OUT.mkdir(exist_ok=True)


def port_vs_trace(trace):
    """Teacher-forced check of the CPU port against a reference trace."""
    from oracle import PortOracle
    rep = {}
    dt = float(trace["dt"]); bmin, bmax = trace["box_min"], trace["box_max"]; iters = int(trace["iters"])
    o = PortOracle(trace["init"], trace["params"])
    o.predictNewPosition(dt)
    rep["predict_bitexact"] = bool(np.array_equal(o.particles["new_position"].view(np.uint32), trace["after_predict"]["new_position"].view(np.uint32)))
    o.upload(trace["after_predict"]); o.findParticleNeighbours()
    w = trace["after_neighbours"]
    rep["hash_bitexact"] = bool(np.array_equal(o.particles["hash"], w["hash"]))
    rep["perm_bitexact"] = bool(np.array_equal(pl.ids_of(o.particles), pl.ids_of(w)))
    arr, cnt = o.neighbours()
    rep["nbr_count_bitexact"] = bool(np.array_equal(cnt, trace["nbr_count"]))
    want = pl.ragged_to_padded(trace["nbr_flat"], trace["nbr_count"], arr.shape[1])
    mask = np.arange(arr.shape[1])[None, :] < trace["nbr_count"][:, None]
    rep["nbr_list_bitexact"] = bool(rep["nbr_count_bitexact"] and np.array_equal(arr[mask], want[mask]))
    o.runConstraintSolver(iters, bmin, bmax)
    w = trace["after_solve"]
    rep["solve_xstar_rel_h"] = pl.rel_err(o.particles["new_position"], w["new_position"], pl.H)
    rep["solve_density_rel"] = pl.rel_err(o.particles["density"], w["density"])
    rep["solve_lambda_rel"] = pl.rel_err(o.particles["lambda"], w["lambda"])
    o.upload(trace["after_solve"]); o.updatePositionAndVelocity(dt)
    rep["update_vel_rel"] = pl.rel_err(o.particles["velocity"], trace["after_update"]["velocity"], pl.H / dt)
    o.upload(trace["after_update"]); o.applyBoundaryVelocityDamping(bmin, bmax)
    rep["damping_vel_rel"] = pl.rel_err(o.particles["velocity"], trace["after_damping"]["velocity"], pl.H / dt)
    o.upload(trace["after_damping"]); o.applyVorticityAndViscosity(dt)
    rep["vv_vel_rel"] = pl.rel_err(o.particles["velocity"], trace["after_vv"]["velocity"], pl.H / dt)
    rep["vv_vorticity_rel"] = pl.rel_err(o.particles["vorticity"], trace["after_vv"]["vorticity"])
    return rep


def timing(n_side, key_mode, steps=20, warm=5, fast=False):
    p, bmin, bmax = scenes.dam_break(n_side)
    s = PBFSolver(len(p), key_mode=key_mode, fast_math=fast)
    s.upload_particles(p)
    for _ in range(warm):
        s.step(0.0083, bmin, bmax)
    s.sync()
    t = time.perf_counter()
    for _ in range(steps):
        s.step(0.0083, bmin, bmax)
    s.sync()
    wall = (time.perf_counter() - t) / steps
    s.enable_timing(True)
    s.step(0.0083, bmin, bmax)
    ph = s.last_step_timing()
    err = s.density_error()
    cnt = s.debug(6)
    s.close()
    return {"n": len(p), "mode": key_mode, "fast": fast, "ms_per_step": wall * 1e3, "pi_per_s": len(p) * 4 / wall,
            "phases_ms": ph, "density_err": err, "nbr_mean": float(cnt.mean()), "nbr_max": int(cnt.max())}


def main():
    log = []

    def emit(name, obj):
        print(name, json.dumps(obj, default=float), flush=True)
        log.append({name: obj})

    gold_dir = OUT / "golden"
    if "--golden" in sys.argv:
        sys.path.insert(0, str(REPO / "tests" / "golden"))
        import make_golden
        make_golden.main(gold_dir)
    else:
        gold_dir = REPO / "tests" / "golden"
    for name in ("lattice12", "jitter", "jitter_k0"):
        f = gold_dir / f"{name}.npz"
        if not f.exists():
            continue
        tr = dict(np.load(f))
        emit(f"port_vs_ref[{name}]", port_vs_trace(tr))
        for mode in (KEY_REFERENCE_HASH, KEY_LINEAR_CELL):
            emit(f"cuda_vs_ref[{name},mode={mode}]", pl.phase_report(tr, mode))
        emit(f"cuda_vs_ref[{name},mode=1,fast]", pl.phase_report(tr, KEY_LINEAR_CELL, fast_math=True))
        if "step1_id" in tr:
            for mode in (KEY_REFERENCE_HASH, KEY_LINEAR_CELL):
                emit(f"traj[{name},mode={mode}]", pl.trajectory_report(tr["init"], tr["box_min"], tr["box_max"], tr["params"], float(tr["dt"]), tr, key_mode=mode))
    f = gold_dir / "dambreak27k.npz"
    if f.exists():
        g = dict(np.load(f))
        init, bmin, bmax = scenes.dam_break(30)
        init["color"][:, 0] = np.arange(len(init), dtype=np.float32)
        emit("ref_rerun_noise", {"dpos": float(g["rerun_step10_max_abs_dpos"]), "dvel": float(g["rerun_step10_max_abs_dvel"])})
        for mode in (KEY_REFERENCE_HASH, KEY_LINEAR_CELL):
            emit(f"traj[dambreak27k,mode={mode}]", pl.trajectory_report(init, bmin, bmax, g["params"], float(g["dt"]), g, key_mode=mode))
        emit("traj[dambreak27k,mode=1,fast]", pl.trajectory_report(init, bmin, bmax, g["params"], float(g["dt"]), g, key_mode=KEY_LINEAR_CELL, fast_math=True))
    if "--timing" in sys.argv:
        for n_side in (30, 100):
            for mode in (KEY_LINEAR_CELL, KEY_REFERENCE_HASH):
                emit("timing", timing(n_side, mode))
        emit("timing", timing(100, KEY_LINEAR_CELL, fast=True))
        emit("timing", timing(160, KEY_LINEAR_CELL))
    (OUT / "parity_report.json").write_text(json.dumps(log, default=float, indent=1))


if __name__ == "__main__":
    main()
