"""``solve_mls_mpm_3d`` with the reference's exact signature
(femflow/solvers/mpm/mls_mpm.py:40-79) running the CUDA substep."""
from __future__ import annotations

from typing import List

import numpy as np

from . import _runtime as R
from .particle import Particle, particles_to_soa, write_back_positions


def make_mls_mpm_coefficients(lenx: int, dim: int):
    """mls_mpm.py:9-15"""
    v = np.zeros((lenx, dim), dtype=np.float64)
    F = np.tile(np.eye(dim, dtype=np.float64), (lenx, 1, 1))
    C = np.zeros((lenx, dim, dim), dtype=np.float64)
    Jp = np.ones((lenx, 1), dtype=np.float64)
    return v, F, C, Jp


def solve_mls_mpm_3d(res: int, inv_dx: float, hardening: float, dx: float, dt: float, volume: float,
                     gravity: float, particles: List[Particle], v: np.ndarray, F: np.ndarray, C: np.ndarray,
                     Jp: np.ndarray, *, p2g_mode: str = "auto"):
    """One substep (zeroed grid, P2G, grid update, G2P); mutates ``particles[i].pos``,
    ``v``, ``F``, ``C`` in place and returns None.  ``Jp`` is untouched: the driver
    hard-codes ``model = "neo_hookean"`` (mls_mpm.py:58).  Raises RuntimeError when a
    particle's stencil leaves the grid (three_d/p2g.py:51-52)."""
    soa = particles_to_soa(particles)
    s = R.solver_for(3, int(res), len(soa), inv_dx=float(inv_dx), dx=float(dx), dt=float(dt),
                     volume=float(volume), gravity=float(gravity), hardening=float(hardening),
                     model="neo_hookean", p2g_mode=p2g_mode)
    s.set_particles(soa.pos, v, F, C, None, soa.mass, soa.mu_0, soa.lambda_0)
    s.substep(1)
    s.check_errors()
    out = s.get_particles()
    write_back_positions(particles, out["x"].double().cpu().numpy())
    v[...] = out["v"].double().cpu().numpy()
    F[...] = out["F"].double().cpu().numpy()
    C[...] = out["C"].double().cpu().numpy()
