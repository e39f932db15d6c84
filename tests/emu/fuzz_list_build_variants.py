"""Diagnostic (not a test): long differential fuzz of the list-build variants under emulation (LINEAR_CELL keys).

    python tests/emu/fuzz_list_build_variants.py SEED TRIALS

Random scenes (uniform boxes, Gaussian blobs, lattice-aligned points with ties at d == h, near-coincident clumps, mixtures;
random boxes that leave particles outside the grid; caps 128 / 64 / 17 / 4): k_build_neighbours_mask<4,5> and <8,4> must leave
exactly the counts, lists and post-step state of k_build_neighbours. tests/test_emu_kernels.py holds the short version."""
import sys, ctypes as C, numpy as np, time
from pathlib import Path; R=Path(__file__).resolve().parents[2]; sys.path.insert(0,str(R)); sys.path.insert(0,str(R/'tests'))
import test_emu_kernels as T
from akuaengine_b200 import PARTICLE_DTYPE
from oracle import param_block
lib=C.CDLL(str(T.OUT)); vp=C.c_void_p
lib.emu_create.restype=vp; lib.emu_create.argtypes=[C.c_uint32,vp,vp,C.c_int,C.c_int,C.c_int,C.c_int]
lib.emu_destroy.argtypes=[vp]; lib.emu_upload_aos108.argtypes=[vp,vp]; lib.emu_download_aos108.argtypes=[vp,vp]
lib.emu_step.argtypes=[vp,C.c_float,C.c_int,vp,vp]; lib.emu_debug_get.argtypes=[vp,C.c_int,vp]
rng=np.random.default_rng(int(sys.argv[1])); bad=0
for trial in range(int(sys.argv[2])):
    n=int(rng.choice([1,2,5,31,32,33,100,500,1500,4000]))
    kind=int(rng.integers(0,5))
    lo=rng.uniform([-1,-1,-1],[2,1,2]); ext=rng.uniform(0.05,1.5,3)
    if kind==0: pos=rng.uniform(lo,lo+ext,(n,3))
    elif kind==1: pos=lo+rng.normal(0,rng.uniform(0.02,0.2),(n,3))
    elif kind==2: pos=np.round(rng.uniform(lo,lo+ext,(n,3))/0.05)*0.05          # lattice-aligned: ties at d == h
    elif kind==3: pos=np.repeat(rng.uniform(lo,lo+ext,((n+7)//8,3)),8,axis=0)[:n]+rng.normal(0,1e-4,(n,3))  # near-coincident clumps
    else: pos=np.concatenate([rng.uniform(lo,lo+ext,(n-n//2,3)), lo+rng.normal(0,0.03,(n//2,3))])
    bmin=(lo+rng.uniform(-0.3,0.3,3)).astype(np.float32); bmax=(bmin+ext+rng.uniform(0.1,1.0,3)).astype(np.float32)
    p=np.zeros(n,PARTICLE_DTYPE); p["position"]=pos.astype(np.float32); p["mass"]=1.0; p["color"][:,0]=np.arange(n)
    cap=int(rng.choice([128,64,17,4]))
    outs=[]
    for lb in (0,1,2):
        s=T.EmuSolver(lib,n,param_block(maxNeighbours=cap),1,list_build=lb); s.upload(p)
        s.step(0.0083,bmin,bmax)
        cnt=s.debug(6); lst=s.debug(7,(n,cap)); m=np.arange(cap)[None,:]<cnt[:,None]
        outs.append((cnt.tobytes(),np.where(m,lst,0).tobytes(),s.download().tobytes())); s.close()
    ok=outs[0]==outs[1]==outs[2]
    cnt=np.frombuffer(outs[0][0],np.uint32)
    if not ok: bad+=1
    print(f"trial {trial} n={n} kind={kind} cap={cap} maxcnt={cnt.max()} capped={(cnt==cap).sum()} identical={ok}", flush=True)
print("BAD",bad)
