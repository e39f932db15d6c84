#!/bin/bash
# usage: tools/regs.sh [extra nvcc flags]  — registers / spills per kernel of libakua_pbf.so (ptxas -v), demangled
nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -Xptxas -v -Xcompiler -fPIC -shared -o /tmp/regs_probe.so "$@" akuaengine_b200/csrc/pbf_solver.cu -ldl 2>&1 | python3 -c "
import re,sys,subprocess
txt=sys.stdin.read()
names=re.findall(r\"Compiling entry function '(\S+)' for 'sm_100a'\",txt)
used=re.findall(r'Used (\d+) registers',txt)
spill=re.findall(r'(\d+) bytes spill stores',txt)
dem=subprocess.run(['c++filt']+names,capture_output=True,text=True).stdout.splitlines()
for d,u,sp in sorted(zip(dem,used,spill)):
    print(u.rjust(4), ('spill '+sp if sp!='0' else '').ljust(10), d.split('(')[0][:100])
errs=[l for l in txt.splitlines() if 'error' in l.lower()]
print('\n'.join(errs[:20]))
"
