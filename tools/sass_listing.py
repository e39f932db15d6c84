"""SASS evidence for the default kernels: python tools/sass_listing.py > profiles/r02_sass_hot_kernels.txt
Part 1: opcode histogram of every kernel in libakua_pbf.so (which load / store widths, which Blackwell-only opcodes appear).
Part 2: the full SASS of the kernels a default step spends its time in."""
import collections
import re
import subprocess
import sys
from pathlib import Path

REPO = Path(__file__).resolve().parents[1]
lib = REPO / "akuaengine_b200" / "libakua_pbf.so"
sass = subprocess.run(["cuobjdump", "-sass", str(lib)], capture_output=True, text=True).stdout
names = subprocess.run(["c++filt"], input="\n".join(re.findall(r"Function : (\S+)", sass)), capture_output=True, text=True).stdout.split("\n")
blocks = re.split(r"\n\s*Function : ", sass)[1:]
HOT = ["k_density_lambda<true, false>", "k_delta_apply<true, false, true, true, false>", "k_delta_apply<true, true, true, true, false>",
       "k_density_lambda<true, true>", "k_build_neighbours_mask<4, 5, true>", "k_onesweep<false, 16>", "k_vorticity<true, false, false>",
       "k_confinement<true, true, false>", "k_xsph<false, false>"]
INTEREST = ("LDG", "STG", "LDS", "STS", "LDGSTS", "UTMA", "UBLKCP", "SYNCS", "MUFU", "FFMA2", "FMUL2", "FADD2", "ACQBULK", "CCTL", "MATCH", "REDUX", "ATOM", "RED", "NANOSLEEP")
print("# Part 1: per-kernel instruction count and the memory / special opcodes in it\n")
keep = {}
for blk, name in zip(blocks, names):
    ops = collections.Counter()
    n = 0
    for line in blk.split("\n"):
        m = re.search(r"/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d\s+)?([A-Z0-9_.]+)", line)
        if m:
            n += 1
            op = m.group(1)
            if op.startswith(INTEREST):
                ops[op] += 1
    short = re.sub(r"\(.*", "", name).replace("void ", "").replace("akua::", "")
    print(f"{short:70s} {n:5d} instr  " + ", ".join(f"{k} x{v}" for k, v in sorted(ops.items())))
    if any(h in name for h in HOT):
        keep[short] = blk
print("\n# Part 2: full SASS of the hot kernels\n")
for short, blk in keep.items():
    print("=" * 30, short)
    for line in blk.split("\n"):
        if re.search(r"/\*[0-9a-f]{4}\*/", line):
            print(re.sub(r"\s*/\* 0x[0-9a-f]+ \*/\s*$", "", line).rstrip())
