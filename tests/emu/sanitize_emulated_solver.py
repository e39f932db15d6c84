"""Diagnostic (not a test): the product's HOST solver + kernels (pbf_solver.cu, pbf_slab.inl, wall_model.cuh) compiled against the
SIMT emulator under AddressSanitizer + UndefinedBehaviorSanitizer, driven through the C ABI: single-rank steps (default and opt-in
wall model), and multi-rank x-slab runs with migration, re-balancing (synchronous and asynchronous) and canonical order. Device
allocations are exact-size heap blocks in the emulator, so an index past a ghost region, an inbox, the cell table or a payload
slot is a heap overflow ASan reports.      python tests/emu/sanitize_emulated_solver.py
"""
import os
import subprocess
import sys
from pathlib import Path

REPO = Path(__file__).resolve().parents[2]
EMU = REPO / "tests" / "emu"
CSRC = REPO / "akuaengine_b200" / "csrc"
LIB = Path("/tmp/libakua_pbf_emu_asan.so")


def build():
    cmd = ["g++", "-O1", "-g", "-std=c++17", "-DAKUA_HOST_EMU", "-DEMU_USE_SWAPCONTEXT", "-U_FORTIFY_SOURCE", "-ffp-contract=off",
           "-fsanitize=address,undefined", "-fno-omit-frame-pointer", "-fPIC", "-shared", "-pthread", "-I", str(EMU),
           "-x", "c++", str(CSRC / "pbf_solver.cu"), "-x", "c++", str(EMU / "emu_core.cpp"), str(EMU / "emu_nccl.cpp"), "-o", str(LIB), "-ldl"]
    subprocess.run(cmd, check=True)


def main():
    if os.environ.get("AKUA_EMU_ASAN_CHILD") != "1":
        build()
        asan = subprocess.run(["gcc", "-print-file-name=libasan.so"], capture_output=True, text=True, check=True).stdout.strip()
        env = dict(os.environ, LD_PRELOAD=asan, ASAN_OPTIONS="detect_leaks=0:detect_stack_use_after_return=0", AKUA_EMU_ASAN_CHILD="1",
                   AKUA_SLAB_WAIT_CYCLES="300000000")
        sys.exit(subprocess.run([sys.executable, __file__], env=env).returncode)
    import numpy as np
    sys.path.insert(0, str(REPO)); sys.path.insert(0, str(REPO / "tests"))
    import test_emu_slab as T
    import test_wall_model as W
    from akuaengine_b200 import load_library, scenes
    lib = load_library(LIB)
    p, bmin, bmax = T._scene(vx=2.0)
    g = scenes.tank_gravity(15.0)
    single = T._run_single(lib, p, bmin, bmax, 6, g, canonical=True)
    print("single rank, canonical order: clean", flush=True)
    for world, skew, reb, asyn in ((2, 0.0, 0, False), (3, 0.0, 0, False), (2, 0.5, 2, False), (2, 0.5, 2, True)):
        pp, b0, b1 = T._scene(nx=32 if skew else 24, vx=2.0)
        ref = T._run_single(lib, pp, b0, b1, 6, g, canonical=True)
        slab = T._run_slab(lib, pp, b0, b1, 6, g, world, skew=skew, rebalance_every=reb, capacity_factor=2.5 if skew else 4.0,
                           canonical=True, rebalance_async=asyn)
        dp, dv = T._compare(pp, ref, slab, 1e-6)
        assert dp == 0.0 and dv == 0.0
        print(f"{world} ranks, skew {skew}, rebalance every {reb} ({'async' if asyn else 'sync'}): clean, bit-identical to one rank", flush=True)
    W.one_particle(lib, 0.03, 1)
    W._resting_block(lib, 1, steps=6)
    print("opt-in wall model: clean", flush=True)


if __name__ == "__main__":
    main()
