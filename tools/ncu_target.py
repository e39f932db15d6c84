"""Target process of the ncu captures: steps a scene past its settle phase so that the kernels ncu profiles see a disordered
fluid. Usage (under ncu, see tools/r02_call7.sh): python tools/ncu_target.py tank200|dam252|dam100 SETTLE STEPS"""
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from akuaengine_b200 import PBFSolver, scenes  # noqa: E402

scene, settle, steps = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
g = None
if scene.startswith("tank"):
    side = int(scene[4:])
    (nx, ny, nz), origin, bmin, bmax = scenes.tank_layout(side, side, side)
    pos, _ = scenes.lattice_slab(nx, ny, nz, origin, 0, nx)
    p = scenes.particles_from_positions(pos)
    g = scenes.tank_gravity(15.0)
else:
    p, bmin, bmax = scenes.dam_break(int(scene[3:]))
s = PBFSolver(len(p), use_graph=False)   # eager: every launch is a plain kernel launch ncu can count with -s / -c
s.upload_particles(p)
if g is not None:
    s.setGravity(g)
for _ in range(settle + steps):
    s.step(0.0083, bmin, bmax)
s.sync()
print("particles", len(p), "steps", settle + steps, "launches", s.counters()["kernel_launches"])
s.close()
