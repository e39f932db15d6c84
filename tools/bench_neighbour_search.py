"""Neighbour-search microbench (BASELINE config 5): cell keys + radix sort + cell ranges + reorder, and separately the
27-cell list build, at 1 M - 128 M particles, uniform vs clustered density. CUDA-event phase timings from the library.
    python tools/bench_neighbour_search.py [--sizes 1,4,16,64] > profiles/r01_neighbour_search_microbench.jsonl"""
import argparse
import json
import sys
from pathlib import Path

import numpy as np

REPO = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(REPO))
from akuaengine_b200 import DBG, KEY_LINEAR_CELL, KEY_REFERENCE_HASH, PBFSolver, scenes  # noqa: E402


def positions(kind, n):
    if kind == "uniform":
        p, bmin, bmax = scenes.uniform_cloud(n, seed=42)
    else:
        p, bmin, bmax = scenes.clustered_cloud(n)
    return p["position"].copy(), bmin, bmax


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--sizes", default="1,4,16,64", help="millions of particles")
    ap.add_argument("--reps", type=int, default=5)
    ap.add_argument("--linear-only", action="store_true")
    a = ap.parse_args()
    for m in [int(x) for x in a.sizes.split(",")]:
        n = m * 1_000_000
        for kind in ("uniform", "clustered"):
            pos, bmin, bmax = positions(kind, n)
            for mode, mname in ((KEY_LINEAR_CELL, "linear"), (KEY_REFERENCE_HASH, "hash")):
                if mode == KEY_REFERENCE_HASH and (n * 128 >= 2 ** 31 or a.linear_only):
                    continue  # the reference's int tableSize overflows (PBFSolver.cpp:15)
                s = PBFSolver(n, key_mode=mode)
                s.upload(pos)
                s.enable_timing(True)
                rows = []
                for _ in range(a.reps + 1):
                    s.findParticleNeighbours(bmin, bmax)
                    rows.append(s.last_step_timing())
                rows = rows[1:]
                med = {k: float(np.median([r[k] for r in rows])) for k in ("predict_key", "sort", "reorder_ranges", "neighbour_lists")}
                search = med["predict_key"] + med["sort"] + med["reorder_ranges"]
                cnt = s.debug(DBG.NBR_COUNT)
                out = {"n": n, "density": kind, "key_mode": mname, "ms": med, "search_ms": search,
                       "search_particles_per_s": n / (search * 1e-3), "search_alg_GBps": (52 - 32 + 4 + 16 * s.counters()["sort_passes_last"] + 104) * n / (search * 1e-3) / 1e9,
                       "list_build_particles_per_s": n / (med["neighbour_lists"] * 1e-3), "sort_passes": s.counters()["sort_passes_last"],
                       "nbr_mean": float(cnt.mean()), "nbr_max": int(cnt.max()), "capped_particles": int((cnt >= 128).sum())}
                print(json.dumps(out), flush=True)
                s.close()


if __name__ == "__main__":
    main()
