"""How much does neighbour ORDER alone move a trajectory? One GPU, the multi-GPU self-check scene (tilted tank, initial x
velocity, 24 steps), run in the three orders the library knows: LINEAR_CELL (cells keep last step's order), REFERENCE_HASH
(the reference's bucket order) and LINEAR_CELL + canonical_order (cells ordered by id). Neighbour SETS are identical; only the
order of the float sums differs. The pairwise differences are the floor any N-GPU-vs-1-GPU comparison in the default order
sits on (akuaengine_b200/slab.py: slab_selfcheck, pass 2).   python tools/order_sensitivity.py > profiles/r02_order_sensitivity.json"""
import json
import sys
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from akuaengine_b200 import KEY_LINEAR_CELL, KEY_REFERENCE_HASH, PBFSolver, scenes  # noqa: E402

H, DT, STEPS = 0.1, 0.0083, 24
p, bmin, bmax = scenes.tank(64, 24, 32)
p["velocity"][:, 0] = np.float32(1.0)
g = scenes.tank_gravity(15.0)


def run(**kw):
    s = PBFSolver(len(p), **kw)
    s.upload_particles(p)
    s.setGravity(g)
    for _ in range(STEPS):
        s.step(DT, bmin, bmax)
    pos, vel, pid = s.download()
    s.close()
    o = np.argsort(pid)
    return pos[o, :3]


runs = {"linear": run(key_mode=KEY_LINEAR_CELL), "hash": run(key_mode=KEY_REFERENCE_HASH),
        "linear_canonical": run(key_mode=KEY_LINEAR_CELL, canonical_order=True), "linear_again": run(key_mode=KEY_LINEAR_CELL)}
out = {"scene": f"tank 64x24x32 ({len(p)} particles), vx = 1, gravity tilted 15 deg, {STEPS} steps, one GPU", "pairs": {}}
names = list(runs)
for i, a in enumerate(names):
    for b in names[i + 1:]:
        d = np.abs(runs[a] - runs[b]).max(axis=1) / H
        out["pairs"][f"{a} vs {b}"] = {"max_dpos_over_h": float(d.max()), "p99": float(np.quantile(d, 0.99)), "rms": float(np.sqrt((d * d).mean()))}
print(json.dumps(out, indent=1))
