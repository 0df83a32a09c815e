"""Stand-alone KM (ghicp_km_solve, the auction) on the CPU emulator against the optimum: random rectangular instances — float costs,
integer costs with masses of ties, sparse graphs, all-penalty graphs, constant costs — must give a valid partial matching on
candidate edges whose energy lies within max(n, m) * eps of scipy's linear_sum_assignment on the reference's padded graph
(src/ghicp_reg.cpp:348-365).  Developer tool, no GPU:   python tools/emu_fuzz_km.py [seconds] [seed]
(Dense near-square integer instances can take minutes HERE: tens of thousands of single-chain reverse rounds, DESIGN §9 limit 2.)"""
import os, sys, tempfile, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import conftest, ghicp_b200 as g, oracle as orc
from scipy.optimize import linear_sum_assignment
orc.build(); g.build_library()
d=tempfile.mkdtemp(); conftest.swap_in_library(g, conftest.build_emulated_library(d)); os.chdir(d)
rng=np.random.default_rng(int(sys.argv[2]) if len(sys.argv)>2 else 5)
t0=time.time(); c=0
while time.time()-t0<(float(sys.argv[1]) if len(sys.argv)>1 else 300.0):
    n=int(rng.integers(1,140)); m=int(rng.integers(1,140))
    kind=rng.choice(['float','int','sparse','allpen','const'])
    eps=float(rng.choice([0.01,0.01,0.1,0.001]))
    if kind=='float': CD=rng.random((n,m))*50; pen=float(rng.uniform(1,60))
    elif kind=='int': CD=rng.integers(150,230,size=(n,m)).astype(float); pen=float(rng.choice([171.3,200.0,150.0,231.0]))
    elif kind=='sparse': CD=rng.random((n,m))*50; pen=float(rng.uniform(0.2,2.0))
    elif kind=='allpen': CD=rng.random((n,m))*50+10; pen=5.0
    else: CD=np.full((n,m),7.0); pen=float(rng.choice([7.0,7.5,3.0]))
    print(c,kind,n,m,pen,eps,flush=True)
    G=orc.km_graph(CD,pen); size=max(n,m)
    match,energy,rounds=g.km_solve(G,sp=n,tp=m,eps=eps,penalty=pen)
    used=[x for x in match if x>=0]
    assert len(used)==len(set(used))
    for y,x in enumerate(match):
        if x>=0: assert x<n and y<m and CD[x,y]<pen,(x,y)
    r,cc=linear_sum_assignment(-G); e_opt=-G[r,cc].sum()
    assert e_opt-1e-9<=energy<=e_opt+size*eps+1e-9,(energy,e_opt,size*eps)
    c+=1
print(c,'cases ok')
