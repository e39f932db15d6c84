"""Times alternative builds of libakua_pbf.so and the gather layouts (tuning experiments):
python tools/time_variants.py [build/*.so]   — every library x fast_math {0,1} x AKUA_GATHER_LAYOUT {plain, packed, records, packed+records}
Environment: AKUA_TV_LAYOUTS=2 (layouts to run), AKUA_TV_NSIDE=100,160, AKUA_TV_PDL=0,1, AKUA_TV_LIST_BUILD=0,1,2 (akua_list_build:
scan / mask4 / mask8)."""
import json, os, subprocess, sys
from pathlib import Path
REPO = Path(__file__).resolve().parents[1]
CODE = r'''
import sys, json, time
sys.path.insert(0, %r)
from akuaengine_b200 import PBFSolver, scenes, KEY_LINEAR_CELL
n_side = int(sys.argv[1]); fast = int(sys.argv[2])
p, bmin, bmax = scenes.dam_break(n_side)
s = PBFSolver(len(p), key_mode=KEY_LINEAR_CELL, fast_math=bool(fast))
s.upload_particles(p)
for _ in range(20): s.step(0.0083, bmin, bmax)
s.sync(); t = time.perf_counter()
for _ in range(30): s.step(0.0083, bmin, bmax)
s.sync(); wall = (time.perf_counter() - t) / 30
s.enable_timing(True); s.step(0.0083, bmin, bmax); ph = s.last_step_timing()
print(json.dumps({"ms": round(wall*1e3, 4), "A": round(ph["pass_a_sum"]/4*1e3), "B": round(ph["pass_b_sum"]/4*1e3), "lists": round(ph["neighbour_lists"]*1e3), "post": round(ph["post"]*1e3), "sort": round(ph["sort"]*1e3)}))
''' % str(REPO)
libs = sys.argv[1:] or [""]
LAYOUTS = {1: "plain", 2: "packed", 3: "records", 4: "packed+records"}
if os.environ.get("AKUA_TV_LAYOUTS"):
    LAYOUTS = {int(k): LAYOUTS[int(k)] for k in os.environ["AKUA_TV_LAYOUTS"].split(",")}
for n_side in [int(v) for v in os.environ.get("AKUA_TV_NSIDE", "100").split(",")]:
    for fast in (1,):
        for lib in libs:
            for layout, lname in LAYOUTS.items():
                env = dict(os.environ)
                if lib: env["AKUA_PBF_LIB"] = str(Path(lib).resolve())
                env["AKUA_GATHER_LAYOUT"] = str(layout)
                for pdl in os.environ.get("AKUA_TV_PDL", "1").split(","):
                    env["AKUA_PDL"] = pdl
                    for lb in os.environ.get("AKUA_TV_LIST_BUILD", "0").split(","):
                        env["AKUA_LIST_BUILD"] = lb
                        r = subprocess.run([sys.executable, "-c", CODE, str(n_side), str(fast)], env=env, capture_output=True, text=True)
                        print(f"n_side={n_side} fast={fast} {Path(lib).name or 'default':20s} {lname:8s} pdl={pdl} list_build={lb}", r.stdout.strip() or r.stderr[-300:], flush=True)
