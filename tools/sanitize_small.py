"""Tiny workload for compute-sanitizer (memcheck / racecheck): two steps + interchange in both key modes."""
import sys
from pathlib import Path
import numpy as np
REPO = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(REPO))
from akuaengine_b200 import KEY_LINEAR_CELL, KEY_REFERENCE_HASH, PBFSolver, scenes

pos = scenes._lattice(13, 11, 9, np.array([1.52, 0.02, 1.52], np.float32))
p = scenes.particles_from_positions(pos)
bmin, bmax = np.array([1.5, 0, 1.5], np.float32), np.array([4.5, 4, 4.5], np.float32)
for mode in (KEY_REFERENCE_HASH, KEY_LINEAR_CELL):
    s = PBFSolver(len(p), key_mode=mode)
    s.upload_particles(p)
    for _ in range(2):
        s.step(0.0083, bmin, bmax)
    q = s.download_particles()
    s.debug(7)
    print(mode, "ok", s.density_error(), np.isfinite(q["position"]).all(), s.counters())
    s.close()
