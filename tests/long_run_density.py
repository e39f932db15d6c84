"""TEST INFRASTRUCTURE (imports oracle/). Long-run statistics (north-star acceptance criterion): mean / max density-constraint error |rho/rho0 - 1| per step of
the README dam break, this library vs the unmodified reference kernels (oracle/_ref), both on the GPU.
    python tests/long_run_density.py [steps] > profiles/r01_long_run_density_error.json"""
import json
import sys
from pathlib import Path

import numpy as np

REPO = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(REPO))
from akuaengine_b200 import KEY_LINEAR_CELL, KEY_REFERENCE_HASH, PBFSolver, scenes  # noqa: E402
from oracle import RefOracle, param_block  # noqa: E402


def curves(steps, n_side=30, dt=0.0083):
    init, bmin, bmax = scenes.dam_break(n_side)
    rho0 = 7600.0
    out = {}
    ref = RefOracle(init, param_block())
    m, x = [], []
    for _ in range(steps):
        ref.step(dt, bmin, bmax)
        e = np.abs(ref.download()["density"].astype(np.float64) / rho0 - 1.0)
        m.append(float(e.mean())); x.append(float(e.max()))
    ref.close()
    out["reference"] = {"mean": m, "max": x}
    # the reference against itself: its XSPH kernel is a data race, so two runs of the same binary diverge chaotically —
    # this is the noise floor for any comparison of long-run statistics
    ref = RefOracle(init, param_block())
    m, x = [], []
    for _ in range(steps):
        ref.step(dt, bmin, bmax)
        e = np.abs(ref.download()["density"].astype(np.float64) / rho0 - 1.0)
        m.append(float(e.mean())); x.append(float(e.max()))
    ref.close()
    out["reference_rerun"] = {"mean": m, "max": x}
    for name, mode in (("linear", KEY_LINEAR_CELL), ("hash", KEY_REFERENCE_HASH)):
        s = PBFSolver(len(init), key_mode=mode)
        s.upload_particles(init)
        m, x = [], []
        for _ in range(steps):
            s.step(dt, bmin, bmax)
            a, b = s.density_error()
            m.append(a); x.append(b)
        s.close()
        out[name] = {"mean": m, "max": x}
    return out


if __name__ == "__main__":
    steps = int(sys.argv[1]) if len(sys.argv) > 1 else 300
    c = curves(steps)
    r = c["reference"]
    summary = {"steps": steps, "scene": "README dam break, 27 000 particles, dt = 0.0083"}
    for name in ("linear", "hash", "reference_rerun"):
        dm = np.abs(np.array(c[name]["mean"]) - np.array(r["mean"])) / np.array(r["mean"])
        dx = np.abs(np.array(c[name]["max"]) - np.array(r["max"])) / np.array(r["max"])
        summary[name] = {"mean_err_rel_diff_max": float(dm.max()), "mean_err_rel_diff_avg": float(dm.mean()),
                         "max_err_rel_diff_max": float(dx.max()), "max_err_rel_diff_avg": float(dx.mean()),
                         "time_avg_mean_err": float(np.mean(c[name]["mean"])), "time_avg_max_err": float(np.mean(c[name]["max"]))}
    summary["reference"] = {"time_avg_mean_err": float(np.mean(r["mean"])), "time_avg_max_err": float(np.mean(r["max"]))}
    print(json.dumps({"summary": summary, "curves": {k: {kk: [round(v, 6) for v in vv[::10]] for kk, vv in c[k].items()} for k in c}}))
