"""Which kernels changed? Compares the SASS of two builds of libakua_pbf.so function by function.

    python tools/sass_diff.py OLD.so NEW.so        (or two files written by `cuobjdump -sass`)

Prints NEW / GONE / DIFF per kernel (instruction counts) and nothing for kernels whose instruction streams are identical —
the check used when a change is claimed not to touch the measured kernels (e.g. making helpers __host__ __device__ or adding
an opt-in variant next to them). Addresses and encodings are ignored; opcodes, operands and order are compared.
"""
import re
import subprocess
import sys


def sass_of(path):
    if path.endswith(".so") or path.endswith(".cubin") or path.endswith(".o"):
        return subprocess.run(["cuobjdump", "-sass", path], capture_output=True, text=True, check=True).stdout
    return open(path).read()


def parse(text):
    funcs, cur = {}, None
    for line in text.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = m.group(1)
            funcs[cur] = []
            continue
        m = re.match(r"\s+/\*[0-9a-f]{4}\*/\s+(.*?);", line)
        if m and cur:
            funcs[cur].append(m.group(1).strip())
    return funcs


def main():
    if len(sys.argv) != 3:
        sys.exit(__doc__)
    a, b = parse(sass_of(sys.argv[1])), parse(sass_of(sys.argv[2]))
    changed = 0
    for f in sorted(set(a) | set(b)):
        if f not in a:
            print(f"NEW   {len(b[f]):5d}        {f}"); changed += 1
        elif f not in b:
            print(f"GONE  {len(a[f]):5d}        {f}"); changed += 1
        elif a[f] != b[f]:
            print(f"DIFF  {len(a[f]):5d} -> {len(b[f]):5d} {f}"); changed += 1
    print(f"{len(a)} kernels before, {len(b)} after, {changed} differ")


if __name__ == "__main__":
    main()
