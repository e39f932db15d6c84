"""Diagnostic: a long x-slab run of the product's host solver + kernels on the CPU (SIMT emulator build of tests/test_emu_slab.py),
8 ranks as 8 threads, sloshing tank, asynchronous re-balancing every CAD steps: no rank may raise (capacity, ghost-plane or
migration overflow, time-outs), particles must be conserved, and the owned counts are printed every 20 steps.
    python tools/emu_long_slab_run.py SIDE STEPS CAD        (e.g. 16 300 5 -> profiles/r02_emu_8rank_300steps_async_rebalance.txt)"""
import os, sys, numpy as np, time, threading
os.environ.setdefault("AKUA_SLAB_WAIT_CYCLES", "300000000")   # bounded waits in the emulated library
from pathlib import Path
REPO = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(REPO / 'tests')); sys.path.insert(0, str(REPO))
import test_emu_slab as t
from akuaengine_b200 import load_library, scenes, PBFSolver
from akuaengine_b200.slab import partition_columns, x_columns
lib = load_library(t.OUT)
world=8; side=int(sys.argv[1]); steps=int(sys.argv[2]); cad=int(sys.argv[3])
(nx,ny,nz),origin,bmin,bmax = scenes.tank_layout(side*world, side, side)
pos,_ = scenes.lattice_slab(nx,ny,nz,origin,0,nx)
p = scenes.particles_from_positions(pos)
n=len(p); ids=np.arange(n,dtype=np.uint32)
g=scenes.tank_gravity(15.0)
cols=x_columns(p["position"][:,0],0.1); col_min=int(cols.min())
bounds=partition_columns(np.bincount(cols-col_min).astype(np.int64), world)
uid=PBFSolver.comm_unique_id(lib)
log=[[] for _ in range(world)]; errs=[]
def work(rank):
    try:
        lo,hi=col_min+int(bounds[rank]),col_min+int(bounds[rank+1])
        mine=(cols>=lo)&(cols<hi)
        s=PBFSolver(n//world, lib=lib, use_graph=False, capacity_factor=4.0, device=rank)
        s.comm_init(rank,world,uid); s.set_slab(lo,hi)
        s.upload_particles(np.ascontiguousarray(p[mine])); s.upload_ids(ids[mine]); s.setGravity(g)
        for k in range(steps):
            s.step(0.0083,bmin,bmax)
            if (k+1)%cad==0:
                s.rebalance_async()
            if (k+1)%20==0:
                log[rank].append(s.n)
        s.close()
    except Exception as e:
        errs.append((rank,e))
ts=[threading.Thread(target=work,args=(r,)) for r in range(world)]
t0=time.time()
[x.start() for x in ts]; [x.join() for x in ts]
print('n',n,'errs',errs,'t',time.time()-t0)
for i in range(len(log[0])):
    row=[log[r][i] for r in range(world)]
    print((i+1)*20, row, 'imb %.2f'%(max(row)/(sum(row)/world)), 'sum', sum(row))
