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


def main_2d(n: int, substeps: int):
    """two_d.p2g / grid_op / g2p chained as the reference's own 2D tests chain them (there is no 2D driver function;
    test_particle_to_grid.py:30-80, test_grid_to_particle.py:30-103): a jittered block of 4 particles per cell on a
    grid sized to hold it, mls-mpm88-style constants scaled so that the block stays stable."""
    warnings.filterwarnings("ignore")
    try:
        sys.path.insert(0, REF)
        from femflow.solvers.mpm import two_d
    except Exception as e:
        print(json.dumps({"unavailable": f"reference import failed: {e!r}"}))
        return
    rng = np.random.default_rng(0)
    side = int(np.ceil(np.sqrt(n / 4)))
    res = 1 << int(np.ceil(np.log2(2 * side)))
    cells = np.stack(np.meshgrid(*[np.arange(side)] * 2, indexing="ij"), -1).reshape(-1, 2)
    x = (np.repeat(cells, 4, 0)[:n] + (res - side) // 2 + rng.uniform(0.05, 0.95, (n, 2))) / res
    x = x.astype(np.float32).astype(np.float64)
    dx, inv_dx = 1.0 / res, float(res)
    E, nu, rho = 1e4, 0.2, 1.0
    mu, lam = E / (2 * (1 + nu)), E * nu / ((1 + nu) * (1 - 2 * nu))
    volume = (dx / 2) ** 2
    mass = rho * volume
    dt = 0.2 * dx / np.sqrt((lam + 2 * mu) / rho)
    v = np.zeros((n, 2)); F = np.tile(np.eye(2), (n, 1, 1)); C = np.zeros((n, 2, 2)); Jp = np.ones((n, 1))

    def substep():
        gv = np.zeros((res + 1, res + 1, 2)); gm = np.zeros((res + 1, res + 1, 1))
        two_d.p2g(inv_dx, 1.0, mu, lam, mass, dx, dt, volume, gv, gm, x, v, F, C, Jp)
        two_d.grid_op(res, dt, -9.8, gv, gm)
        two_d.g2p(inv_dx, dt, gv, x, v, F, C, Jp)
    t0 = time.perf_counter()
    substep()
    t_jit = time.perf_counter() - t0
    t0 = time.perf_counter()
    for _ in range(substeps):
        substep()
    sec = (time.perf_counter() - t0) / substeps
    print(json.dumps({"value": n / sec, "unit": "particle-substeps/s", "cores": 1, "kind": "numba", "dim": 2,
                      "sample": f"{n} particles, 4 per cell, on the {res}^2 grid, {substeps} substeps after the JIT call; unmodified "
                                f"femflow.solvers.mpm.two_d.p2g / grid_op / g2p (serial njit: 1 core of {os.cpu_count()}), "
                                f"{sec:.2f} s per substep, JIT call {t_jit:.1f} s",
                      "seconds_per_substep": sec}))


def main():
    if len(sys.argv) > 1 and sys.argv[1] == "2d":
        if not os.path.isdir(os.path.join(REF, "femflow")):
            print(json.dumps({"unavailable": "baseline/_ref not staged (oracle/vendor_reference.py runs in the build container)"}))
            return
        return main_2d(int(sys.argv[2]) if len(sys.argv) > 2 else 65536, int(sys.argv[3]) if len(sys.argv) > 3 else 3)
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
    side = int(round((n / 8) ** (1 / 3))) + 1
    res = 64 if side <= 60 else 1 << int(np.ceil(np.log2(side + 4)))
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
                      "sample": f"{n} particles{' (paper scene count)' if n == 35321 else ''} on the {res}^3 grid, {substeps} substeps after the JIT call; "
                                f"unmodified femflow.solvers.mpm.mls_mpm.solve_mls_mpm_3d (serial njit: 1 core of "
                                f"{os.cpu_count()}), {sec:.2f} s per substep, JIT call {t_jit:.1f} s, typed-list build {t_list:.1f} s",
                      "seconds_per_substep": sec}))


if __name__ == "__main__":
    main()
