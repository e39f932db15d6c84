"""Diagnostic (not a test): the emulated kernels under AddressSanitizer + UndefinedBehaviorSanitizer.

    python tests/emu/sanitize_emulated_kernels.py          # builds /tmp/libakua_emu_asan.so, re-executes itself under libasan

Every buffer of the emulated solver is a std::vector of exactly n elements, so any out-of-range index of a kernel (reads past
the last particle, list slots past the allocation, shared-memory overruns) is a heap / global buffer overflow ASan reports.
Covers: the radix sort, whole steps in both key modes with every list build and gather layout, the zero-iteration commit
path, the neighbour-list export, the AoS pack / unpack kernels and the migration compaction.
"""
import ctypes as C
import os
import subprocess
import sys
from pathlib import Path

REPO = Path(__file__).resolve().parents[2]
EMU = REPO / "tests" / "emu"
LIB = Path("/tmp/libakua_emu_asan.so")


def build():
    cmd = ["g++", "-O1", "-g", "-std=c++17", "-DAKUA_HOST_EMU", "-DEMU_USE_SWAPCONTEXT", "-U_FORTIFY_SOURCE", "-ffp-contract=off",
           "-fsanitize=address,undefined", "-fno-omit-frame-pointer", "-fPIC", "-shared", "-I", str(EMU), "-o", str(LIB),
           str(EMU / "emu_harness.cpp"), str(EMU / "emu_core.cpp")]
    subprocess.run(cmd, check=True)


def main():
    if os.environ.get("AKUA_EMU_ASAN_CHILD") != "1":
        build()
        asan = subprocess.run(["gcc", "-print-file-name=libasan.so"], capture_output=True, text=True, check=True).stdout.strip()
        env = dict(os.environ, LD_PRELOAD=asan, ASAN_OPTIONS="detect_leaks=0:detect_stack_use_after_return=0", AKUA_EMU_ASAN_CHILD="1")
        sys.exit(subprocess.run([sys.executable, __file__], env=env).returncode)
    import numpy as np
    sys.path.insert(0, str(REPO)); sys.path.insert(0, str(REPO / "tests"))
    import test_emu_kernels as T
    import test_zz_list_build_gpu as Z
    from oracle import param_block
    lib = C.CDLL(str(LIB))
    vp = C.c_void_p
    lib.emu_create.restype = vp
    lib.emu_create.argtypes = [C.c_uint32, vp, vp, C.c_int, C.c_int, C.c_int, C.c_int]
    lib.emu_destroy.argtypes = [vp]; lib.emu_upload_aos108.argtypes = [vp, vp]; lib.emu_download_aos108.argtypes = [vp, vp]
    lib.emu_step.argtypes = [vp, C.c_float, C.c_int, vp, vp]; lib.emu_debug_get.argtypes = [vp, C.c_int, vp]
    lib.emu_sort_pairs.argtypes = [vp, C.c_uint32, C.c_int, vp, vp, C.c_int, C.c_int]
    lib.emu_plane_hist.argtypes = [vp, vp, C.c_uint32, C.c_uint32, C.c_int, vp, vp]
    lib.emu_migration.argtypes = [vp, C.c_uint32, C.c_uint32, C.c_int, C.c_int, C.c_uint32, vp, C.c_uint32, vp, vp, vp]
    rng = np.random.default_rng(0)
    for n, bits in [(1, 8), (1025, 21), (5000, 32)]:
        keys = rng.integers(0, 2 ** bits, n, dtype=np.uint64).astype(np.uint32)
        ko, vo = np.empty(n, np.uint32), np.empty(n, np.uint32)
        for mode, items in ((0, 0), (1, 4), (1, 8), (1, 16)):
            lib.emu_sort_pairs(keys.ctypes.data, n, bits, ko.ctypes.data, vo.ctypes.data, mode, items)
        print("sort", n, "clean (three-kernel and one-sweep variants)", flush=True)
    for name in ["jittered block, cap 20", "dense blob (cap 128 bites, rows > 32 candidates)"]:
        init, bmin, bmax, cap = Z._scene(name)
        init = init[:1500]
        for mode, lb in ((1, 0), (1, 1), (1, 2), (0, 0)):
            for pack in (0, 1):
                s = T.EmuSolver(lib, len(init), param_block(maxNeighbours=cap), mode, list_build=lb, pack=pack)
                s.upload(init)
                for _ in range(2):
                    s.step(0.0083, bmin, bmax)
                s.step(0.0083, bmin, bmax, iters=0)
                s.debug(7, (len(init), cap)); s.download(); s.close()
                print(f"{name}: key_mode {mode} list_build {lb} pack {pack} clean", flush=True)
    for n in (1, 3000, 2048 * 3 + 5):     # migration tiles are 2048 particles: one partial tile, two tiles, a ragged fourth
        keys = (rng.integers(0, 12, n) * 35 + rng.integers(0, 35, n)).astype(np.uint32)
        ids = np.arange(n, dtype=np.uint32)
        counts, il, ir = np.zeros(64, np.uint32), np.zeros(n, np.uint32), np.zeros(n, np.uint32)
        lib.emu_migration(keys.ctypes.data, n, 35, 4, 9, 12 * 35, ids.ctypes.data, n, counts.ctypes.data, il.ctypes.data, ir.ctypes.data)
        ks = np.sort(keys)
        hist, work = np.zeros(12, np.uint64), np.zeros(12, np.uint64)
        nbr = rng.integers(0, 60, n).astype(np.uint32)
        lib.emu_plane_hist(ks.ctypes.data, nbr.ctypes.data, n, 35, 12, hist.ctypes.data, work.ctypes.data)
        print("migration + plane histogram", n, "clean", flush=True)


if __name__ == "__main__":
    main()
