"""Experiment (not product code): do the issue-bound P2G and the latency-bound G2P overlap when run on two
streams?  Two independent solvers on one GPU; prints ms per iteration alone and concurrently."""
import sys; sys.path.insert(0,".")
import torch, os
from bench import make_scene
from femflow_b200.mpm import MpmSolver
sc=make_scene("3d16m"); n=sc.n
def mk():
    s=MpmSolver(3,sc.res,sc.dt,sc.volume,sc.gravity,sc.hardening,capacity=n)
    s.set_particles(sc.x,sc.v,sc.F,sc.C,None,sc.mass,sc.mu_0,sc.lambda_0)
    s.substep(4); torch.cuda.synchronize(); return s
A=mk(); B=mk()
sa=torch.cuda.Stream(); sb=torch.cuda.Stream()
def run_p2g(k):
    with torch.cuda.stream(sa):
        for _ in range(k): A.clear_grid(); A.p2g()
def run_g2p(k):
    with torch.cuda.stream(sb):
        for _ in range(k): B.bin(); B.g2p()
def timeit(fn, reps=3):
    best=1e9
    for _ in range(reps):
        torch.cuda.synchronize(); e0=torch.cuda.Event(enable_timing=True); e1=torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); torch.cuda.synchronize(); e1.record(); torch.cuda.synchronize(); best=min(best,e0.elapsed_time(e1))
    return best
# prepare B grid with velocities
B.clear_grid(); B.bin(); B.p2g(); B.grid_op(); A.bin(); torch.cuda.synchronize()
K=10
tp=timeit(lambda: run_p2g(K)); tg=timeit(lambda: run_g2p(K))
def both(): run_p2g(K); run_g2p(K)
tb=timeit(both)
print("P2G alone %.3f ms/iter, bin+G2P alone %.3f ms/iter, concurrent %.3f ms/iter (sum %.3f)"%(tp/K,tg/K,tb/K,(tp+tg)/K))
