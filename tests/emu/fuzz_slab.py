"""Diagnostic (not a test): differential fuzz of the x-slab step on the CPU. Random small tanks, 2 - 5 emulated ranks, random initial
velocities and gravity tilt, random skewed starts, random re-balancing cadence (synchronous or asynchronous), canonical order:
every run must conserve particles, keep the payload with its particle and be BIT-IDENTICAL to the single-rank run.
    python tests/emu/fuzz_slab.py [CASES] [SEED]"""
import sys
import time
from pathlib import Path

import os

import numpy as np

os.environ.setdefault("AKUA_SLAB_WAIT_CYCLES", "300000000")   # bounded waits in the emulated library (a failed rank must not hang the others)

REPO = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(REPO / "tests")); sys.path.insert(0, str(REPO))
import test_emu_slab as T  # noqa: E402
from akuaengine_b200 import load_library, scenes  # noqa: E402


def main():
    cases = int(sys.argv[1]) if len(sys.argv) > 1 else 12
    seed = int(sys.argv[2]) if len(sys.argv) > 2 else 1
    rng = np.random.default_rng(seed)
    import subprocess
    if not T.OUT.exists():
        subprocess.run([sys.executable, "-m", "pytest", str(REPO / "tests" / "test_emu_slab.py"), "-q", "-k", "without_migration"], check=True)
    lib = load_library(T.OUT)
    bad = 0
    for c in range(cases):
        world = int(rng.integers(2, 6))
        nx = int(rng.integers(6 * world, 10 * world))
        ny, nz = int(rng.integers(5, 9)), int(rng.integers(5, 11))
        p, bmin, bmax = scenes.tank(nx, ny, nz)
        n = len(p)
        p["velocity"][:, 0] = np.float32(rng.uniform(-2.5, 2.5))
        p["velocity"][:, 1] = np.float32(rng.uniform(-1.0, 1.0))
        p["color"][:, 0] = (np.arange(n) % 251).astype(np.float32)
        p["size"] = (np.arange(n) % 17 + 1).astype(np.float32)
        g = scenes.tank_gravity(float(rng.uniform(-25, 25)))
        steps = int(rng.integers(6, 16))
        skew = float(rng.choice([0.0, 0.0, 0.3, 0.5]))
        reb = int(rng.choice([0, 2, 3, 5]))
        asyn = bool(rng.integers(0, 2))
        t0 = time.time()
        try:
            single = T._run_single(lib, p, bmin, bmax, steps, g, canonical=True)
            slab = T._run_slab(lib, p, bmin, bmax, steps, g, world, skew=skew, rebalance_every=reb, capacity_factor=4.0,
                               canonical=True, rebalance_async=asyn)
            dp, dv = T._compare(p, single, slab, 1e-6)
            mig = sum(o[4]["migrated_in"] for o in slab)
            ok = dp == 0.0 and dv == 0.0
            verdict = "bit-identical" if ok else f"DIFFERS dp={dp} dv={dv}"
        except Exception as e:  # noqa: BLE001
            ok, mig, verdict = False, -1, f"ERROR {str(e)[:160]}"
        bad += not ok
        print(f"case {c}: {world} ranks, tank {nx}x{ny}x{nz} ({n} particles), {steps} steps, skew {skew}, rebalance every {reb} "
              f"({'async' if asyn else 'sync'}), migrated {mig}: {verdict} [{time.time() - t0:.0f} s]", flush=True)
    print(f"{cases - bad} / {cases} cases bit-identical to one rank")
    sys.exit(1 if bad else 0)


if __name__ == "__main__":
    main()
