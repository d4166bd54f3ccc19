"""Time the UNMODIFIED reference's numba substep (femflow/solvers/mpm/mls_mpm.py:40-79 solve_mls_mpm_3d) on this
box's host: the paper scene's particle count (35 321, paper_1.py:75-114) as a jittered block on the 64^3 grid.

BENCH INFRASTRUCTURE (bench.py's cpu_baseline leg runs it as a subprocess).  Needs ``baseline/_ref`` (staged by
oracle/vendor_reference.py in the build container) and numba; prints ONE JSON line, ``{"unavailable": why}`` when
either is missing.  The reference's kernels are serial ``@nb.njit`` loops: 1 core, whatever the box has.
"""
from __future__ import annotations

import json
import os
import sys
import time
import warnings

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.path.join(HERE, "..", "baseline", "_ref")


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 35321
    substeps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
    if not os.path.isdir(os.path.join(REF, "femflow")):
        print(json.dumps({"unavailable": "baseline/_ref not staged (oracle/vendor_reference.py runs in the build container)"}))
        return
    warnings.filterwarnings("ignore")
    try:
        sys.path.insert(0, REF)
        from femflow.solvers.mpm.mls_mpm import solve_mls_mpm_3d
        from femflow.solvers.mpm.particle import Particle
        from numba.typed import List as NbList
    except Exception as e:  # numba / scipy missing on this box
        print(json.dumps({"unavailable": f"reference import failed: {e!r}"}))
        return
    rng = np.random.default_rng(0)
    res = 64
    side = int(round((n / 8) ** (1 / 3))) + 1
    cells = np.stack(np.meshgrid(*[np.arange(side)] * 3, indexing="ij"), -1).reshape(-1, 3)
    x = (np.repeat(cells, 8, 0)[:n] + (res - side) // 2 + rng.uniform(0.05, 0.95, (n, 3))) / res
    x = x.astype(np.float32).astype(np.float64)
    # paper_1.py:75-86 gyroid material: mass 1, E 140, nu 0.2 ; volume 1, hardening 0.7, dt 1e-4
    E, nu = 140.0, 0.2
    mu, lam = E / (2 * (1 + nu)), E * nu / ((1 + nu) * (1 - 2 * nu))
    t0 = time.perf_counter()
    particles = NbList()
    for i in range(n):
        particles.append(Particle(x[i].copy(), 0.0, 1.0, lam, mu))     # particle.py:20-27: (pos, force, mass, lambda, mu)
    t_list = time.perf_counter() - t0
    v = np.zeros((n, 3)); F = np.tile(np.eye(3), (n, 1, 1)); C = np.zeros((n, 3, 3)); Jp = np.ones((n, 1))
    args = (res, float(res), 0.7, 1.0 / res, 1e-4, 1.0, -9.8, particles, v, F, C, Jp)
    t0 = time.perf_counter()
    solve_mls_mpm_3d(*args)                                            # JIT + first substep
    t_jit = time.perf_counter() - t0
    t0 = time.perf_counter()
    for _ in range(substeps):
        solve_mls_mpm_3d(*args)
    sec = (time.perf_counter() - t0) / substeps
    print(json.dumps({"value": n / sec, "unit": "particle-substeps/s", "cores": 1, "kind": "numba",
                      "sample": f"{n} particles (paper scene count) on the 64^3 grid, {substeps} substeps after the JIT call; "
                                f"unmodified femflow.solvers.mpm.mls_mpm.solve_mls_mpm_3d (serial njit: 1 core of "
                                f"{os.cpu_count()}), {sec:.2f} s per substep, JIT call {t_jit:.1f} s, typed-list build {t_list:.1f} s",
                      "seconds_per_substep": sec}))


if __name__ == "__main__":
    main()
