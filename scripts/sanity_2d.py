"""Small 2D run for compute-sanitizer (memcheck / racecheck) over the warp-window kernels of csrc/mpm_2d_window.cuh:
lattice order (node tile), cell-sorted order (lanes taking turns) and shuffled order (direct reductions), 3 substeps each,
checked against the NumPy oracle."""
import sys

import numpy as np

sys.path.insert(0, ".")
from femflow_b200 import scenes          # noqa: E402
from femflow_b200.mpm import MpmSolver   # noqa: E402
from oracle import mpm_oracle as O       # noqa: E402

sc = scenes.elastic_block(2, 64, 40, 3, seed=4)
n = sc.n - 21
rng = np.random.default_rng(4)
for order in ("lattice", "cell_sorted", "shuffled"):
    x, v, F, C = (a[:n].astype(np.float64) for a in (sc.x, sc.v, sc.F, sc.C))
    if order == "shuffled":
        perm = rng.permutation(n)
    elif order == "cell_sorted":
        base = np.floor(np.float32(x) * np.float32(sc.res) - np.float32(0.5)).astype(np.int64)
        perm = np.argsort(base[:, 0] * (sc.res + 1) + base[:, 1], kind="stable")
    else:
        perm = np.arange(n)
    x, v, F, C = x[perm], v[perm], F[perm], C[perm]
    Jp = np.ones((n, 1))
    s = MpmSolver(2, sc.res, sc.dt, sc.volume, sc.gravity, 1.0, capacity=n, mass=sc.mass, mu_0=sc.mu_0, lambda_0=sc.lambda_0)
    s.set_particles(x, v, F, C, Jp)
    s.substep(3)
    s.check_errors()
    for _ in range(3):
        O.solve_mls_mpm_2d(sc.res, float(sc.res), 1.0, sc.mu_0, sc.lambda_0, sc.mass, 1 / sc.res, sc.dt, sc.volume, sc.gravity,
                           x, v, F, C, Jp)
    out = {k: t.double().cpu().numpy() for k, t in s.get_particles().items()}
    V = max(np.abs(v).max(), sc.dt * 9.8)
    errs = {"x": np.abs(out["x"] - x).max(), "v": np.abs(out["v"] - v).max() / V, "F": np.abs(out["F"] - F).max(),
            "C": np.abs(out["C"] - C).max() / (4 * sc.res * V)}
    print(order, n, errs, "launches", s.launch_count(), flush=True)
    assert all(e < 3e-5 for e in errs.values()), errs
    s.close()
print("sanity_2d ok")
