"""Random configurations of the host-compiled library (tests/hostsim, test infrastructure) against the oracle: random small sizes,
iteration counts 2..6, seeds, uniform values and garbage in motion / depth / roughness; three frames each, whole-frame tolerances of
tests/_harness.py.  python tools/hostsim_fuzz.py <rng seed> <seconds>.  Last run: profiles/r02_hostsim_memcheck.md."""
import os, sys, time
import numpy as np
ROOT=os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT,"tests"))
from raytracedgl1_b200 import denoiser
denoiser._LIB_PATH = os.path.join(ROOT, "tests/hostsim/libasvgf_hostsim.so")
denoiser._TEST_ONLY_ALLOW_HOST_BUILD = True
from _harness import Pair
rng=np.random.default_rng(int(sys.argv[1]))
t_end=time.time()+float(sys.argv[2])
n=0; fails=0
while time.time()<t_end:
    W=int(rng.choice([34,66,70,98,130])); H=int(rng.integers(7,60)); N=int(rng.choice([2,3,4,4,4,5,6]))
    seed=int(rng.integers(1,1<<30))
    def garbage(g, r=np.random.default_rng(seed)):
        if r.random()<0.3:
            m=g["Motion"]; k=r.integers(3,9)
            m[::k, ::k+1, 0]*=np.float16(r.uniform(-8,8)); m[1::k, ::k, 1]+=np.float16(r.uniform(-1,1))
        if r.random()<0.2:
            g["DepthWorld"][::r.integers(2,7), ::r.integers(2,7)] = np.float16(r.choice([0.0,-1.0,65504.0,1e-3]))
        if r.random()<0.2:
            g["MetallicRoughness"][::3, ::2, 1] = r.integers(0,256)
    def edit(u, r=np.random.default_rng(seed+1)):
        u.timeDelta=float(r.choice([0.001,0.016,0.033,0.5])); u.lightmapScreenCoverage=float(r.choice([0,0,0.3,0.5,1.0]))
        u.antiFireflyEnabled=int(r.random()<0.8)
        u.gradientMultDiffuse=float(r.uniform(0,2)); u.gradientMultSpecular=float(r.uniform(0,2)); u.gradientMultIndirect=float(r.uniform(0,2))
    p=Pair(W,H,iterations=N,seed=seed)
    try:
        for frame in range(3):
            fi,u=p.run_frame(frame, uniform_edit=edit, mutate=garbage)
            res=p.compare(fi)
            bad={k:(v["n_bad"],v["n_outlier"],round(v["worst"],2),round(v["exact_frac"],4)) for k,v in res.items() if not v["ok"]}
            if bad:
                fails+=1; print("FAIL",W,H,N,seed,frame,bad,flush=True)
            p.sync_state_to_cuda()
    finally:
        p.close()
    n+=1
print("fuzz done:",n,"configs,",fails,"failing frames",flush=True)
