"""CUDA-graph replay of the substep against eager launches: parity (3D, reordering pipeline) and
timing (2D 1M, where the four launches of a 71 us substep are latency-bound).  One GPU, ~20 s."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from femflow_b200 import scenes  # noqa: E402
from femflow_b200.mpm import MpmSolver  # noqa: E402


def solver_for(sc):
    s = MpmSolver(sc.dim, sc.res, sc.dt, sc.volume, sc.gravity, sc.hardening, capacity=sc.n, device="cuda:0",
                  mass=sc.mass, mu_0=sc.mu_0, lambda_0=sc.lambda_0)
    s.set_particles(sc.x, sc.v, sc.F, sc.C, None, *((sc.mass, sc.mu_0, sc.lambda_0) if sc.dim == 3 else (None,) * 3))
    return s


def timed(fn, reps):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


out = {}
# ---- parity: 3D block, 2 warm substeps, then 6 more eagerly or as 3 replays of a 2-substep graph ----
sc = scenes.elastic_block(3, 64, 24, 2, seed=1)
a, b = solver_for(sc), solver_for(sc)
a.substep(8)
b.substep(2)
side = torch.cuda.Stream()
with torch.cuda.stream(side):
    g = b.make_graph(2)          # capture does not execute
    for _ in range(3):
        g.replay()
torch.cuda.synchronize()
pa, pb = a.get_particles(), b.get_particles()
out["parity_max_abs"] = {k: float((pa[k] - pb[k]).abs().max()) for k in pa}
out["oob"] = [a.poll_error(), b.poll_error()]
out["graph_launches_per_2_substeps"] = b.graph_launches

# ---- timing: 2D 1M ----
sc2 = scenes.config_2d_1m(0)
s2 = solver_for(sc2)
s2.substep(10)
eager = timed(lambda: s2.substep(100), 3) / 100
with torch.cuda.stream(side):
    g2 = s2.make_graph(100)
    graph = timed(g2.replay, 3) / 100
out["2d1m_ms_per_substep"] = {"eager": eager, "graph": graph}

# ---- timing: 3D 2M (reordering pipeline with the internal stream inside the capture) ----
sc3 = scenes.elastic_block(3, 256, 64, 2, seed=0)
s3 = solver_for(sc3)
s3.substep(10)
eager3 = timed(lambda: s3.substep(50), 3) / 50
with torch.cuda.stream(side):
    g3 = s3.make_graph(50)
    graph3 = timed(g3.replay, 3) / 50
out["3d2m_ms_per_substep"] = {"eager": eager3, "graph": graph3}
out["oob_after"] = [s2.poll_error(), s3.poll_error()]
os.makedirs("gpurun_out", exist_ok=True)
with open("gpurun_out/graph_experiment.json", "w") as f:
    json.dump(out, f, indent=1)
print(json.dumps(out))
