"""Generates the golden fixtures in tests/golden/ from the UNMODIFIED reference kernels (oracle/_ref/libakua_ref.so).

Needs a GPU: run on a B200 box with
    gpurun -- 'python tests/golden/make_golden.py gpurun_out/golden'
then copy gpurun_out/golden/*.npz into tests/golden/. The reference holds no tests or golden vectors of its own
(SURVEY.md §4), so these outputs of its own kernels are what pins the CPU port (tests/test_oracle_golden.py) and,
through it and directly, the CUDA path (tests/test_parity_gpu.py).

Particle identity: the reference has no id field and physically permutes the structs, so the id of each particle is
stashed as a float in color.x (the solver never reads color; exact for ids < 2^24).
"""
from __future__ import annotations

import sys
from pathlib import Path

import numpy as np

REPO = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(REPO))

from akuaengine_b200 import scenes  # noqa: E402
from oracle import RefOracle, param_block  # noqa: E402

DT = np.float32(0.0083)


def with_ids(p):
    p = p.copy()
    p["color"][:, 0] = np.arange(len(p), dtype=np.float32)
    return p


def scene_lattice12():
    """12^3 lattice block resting in the corner of the README box: exercises the d == h boundary cases of the lattice,
    the floor / wall collision clamp and the velocity damping."""
    pos = scenes._lattice(12, 12, 12, np.array([1.52, 0.02, 1.52], np.float32))
    p = scenes.particles_from_positions(pos)
    return with_ids(p), np.array([1.5, 0.0, 1.5], np.float32), np.array([4.5, 4.0, 4.5], np.float32)


def scene_jitter():
    """2000 particles at about rest density with pseudo-random positions, velocities and masses, touching two walls."""
    n = 2000
    u = scenes._uniform01(7 * n, 7).reshape(n, 7)
    side = (n / 7.6) ** (1 / 3) * 0.1
    pos = (np.array([1.5, 0.0, 1.5], np.float32) + u[:, :3] * np.float32(side)).astype(np.float32)
    p = scenes.particles_from_positions(pos)
    p["velocity"] = ((u[:, 3:6] - 0.5) * 2.0).astype(np.float32)
    p["mass"] = (0.9 + 0.2 * u[:, 6]).astype(np.float32)
    return with_ids(p), np.array([1.5, 0.0, 1.5], np.float32), np.array([4.5, 4.0, 4.5], np.float32)


def ragged(arr, cnt):
    flat = np.concatenate([arr[i, :cnt[i]] for i in range(len(cnt))]) if len(cnt) else np.zeros(0, np.uint32)
    return flat.astype(np.uint32)


def phase_trace(init, bmin, bmax, params, iters=4):
    """One step through the six wrappers, snapshotting the full AoS state after each (teacher-forcing fixtures)."""
    ref = RefOracle(init, params)
    out = {"init": init, "dt": DT, "box_min": bmin, "box_max": bmax, "params": params, "iters": np.int32(iters)}
    ref.predictNewPosition(DT)
    out["after_predict"] = ref.download()
    ref.findParticleNeighbours()
    out["after_neighbours"] = ref.download()
    arr, cnt = ref.neighbours()
    out["nbr_count"] = cnt
    out["nbr_flat"] = ragged(arr, cnt)
    ref.runConstraintSolver(iters, bmin, bmax)
    out["after_solve"] = ref.download()
    ref.updatePositionAndVelocity(DT)
    out["after_update"] = ref.download()
    ref.applyBoundaryVelocityDamping(bmin, bmax)
    out["after_damping"] = ref.download()
    ref.applyVorticityAndViscosity(DT)
    out["after_vv"] = ref.download()
    ref.close()
    return out


def trajectory(init, bmin, bmax, params, steps):
    """Free-running PBFSolver::step; returns id / position / velocity (+density) snapshots at the requested steps."""
    ref = RefOracle(init, params)
    out = {}
    for s in range(1, max(steps) + 1):
        ref.step(DT, bmin, bmax)
        if s in steps:
            p = ref.download()
            out[f"step{s}_id"] = p["color"][:, 0].astype(np.uint32)
            out[f"step{s}_position"] = p["position"].copy()
            out[f"step{s}_velocity"] = p["velocity"].copy()
            out[f"step{s}_density"] = p["density"].copy()
    ref.close()
    return out


def main(outdir):
    outdir = Path(outdir)
    outdir.mkdir(parents=True, exist_ok=True)
    params = param_block()
    for name, fn in (("lattice12", scene_lattice12), ("jitter", scene_jitter)):
        init, bmin, bmax = fn()
        d = phase_trace(init, bmin, bmax, params)
        d.update(trajectory(init, bmin, bmax, params, (1, 10)))
        np.savez_compressed(outdir / f"{name}.npz", **d)
        print(name, "n =", len(init), "mean nbrs", d["nbr_count"].mean(), "max", d["nbr_count"].max())
    # artificial pressure off (k = 0), README scene constants otherwise
    init, bmin, bmax = scene_jitter()
    p0 = param_block(k=0.0)
    d = phase_trace(init, bmin, bmax, p0)
    np.savez_compressed(outdir / "jitter_k0.npz", **d)
    # config 1: README dam break, 27 000 particles; trajectory only (positions/velocities after 1 and 10 steps)
    init, bmin, bmax = scenes.dam_break(30)
    init = with_ids(init)
    d = {"dt": DT, "box_min": bmin, "box_max": bmax, "params": params}
    d.update(trajectory(init, bmin, bmax, params, (1, 10)))
    # run-to-run noise floor of the reference itself (XSPH race): a second identical run
    d2 = trajectory(init, bmin, bmax, params, (10,))
    o1 = np.argsort(d["step10_id"]); o2 = np.argsort(d2["step10_id"])
    d["rerun_step10_max_abs_dpos"] = np.abs(d["step10_position"][o1] - d2["step10_position"][o2]).max()
    d["rerun_step10_max_abs_dvel"] = np.abs(d["step10_velocity"][o1] - d2["step10_velocity"][o2]).max()
    np.savez_compressed(outdir / "dambreak27k.npz", **d)
    print("dambreak27k rerun noise", d["rerun_step10_max_abs_dpos"], d["rerun_step10_max_abs_dvel"])


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else str(Path(__file__).resolve().parent))
