"""2D phase functions with the reference's signatures
(femflow/solvers/mpm/two_d/{p2g,grid_op,g2p}.py), executed on the GPU."""
from __future__ import annotations

import numpy as np

from . import _runtime as R


def _solver(res, n, dt, volume, gravity, hardening, mass, mu_0, lambda_0, model, inv_dx=None, dx=None):
    # inv_dx / dx as the caller passed them (two_d/p2g.py:12-20, g2p.py:6-8): the reference never derives one
    # from the other or from the grid shape, so neither do the wrappers
    if inv_dx is not None and dx is None:
        dx = 1.0 / float(inv_dx)
    extra = {} if inv_dx is None else {"inv_dx": float(inv_dx), "dx": float(dx)}
    return R.solver_for(2, res, n, dt=float(dt), volume=float(volume), gravity=float(gravity),
                        hardening=float(hardening), mass=float(mass), mu_0=float(mu_0),
                        lambda_0=float(lambda_0), model=model, **extra)


def p2g(inv_dx, hardening, mu_0, lambda_0, mass, dx, dt, volume, grid_velocity, grid_mass, x, v, F, C, Jp,
        model: str = "neo_hookean"):
    """two_d/p2g.py:11-76 (``+=`` into the caller's grids).  Unlike the reference
    (no bounds check: undefined behaviour) a stencil outside the grid raises
    RuntimeError."""
    G = grid_velocity.shape[0]
    s = _solver(G - 1, len(x), dt, volume, 0.0, hardening, mass, mu_0, lambda_0, model, inv_dx, dx)
    s.set_particles(x, v, F, C, Jp)
    s.clear_grid()
    s.p2g()
    s.check_errors()
    gv, gm = R.grid_from_device(s)
    grid_velocity += gv
    grid_mass += gm


def grid_op(grid_resolution, dt, gravity, grid_velocity, grid_mass):
    """two_d/grid_op.py:5-24, in place on ``grid_velocity``."""
    s = _solver(grid_resolution, 0, dt, 1.0, gravity, 1.0, 1.0, 1.0, 1.0, "neo_hookean")
    R.grid_to_device(s, grid_velocity, grid_mass)
    s.grid_op()
    gv, _ = R.grid_from_device(s)
    grid_velocity[...] = gv


def g2p(inv_dx, dt, grid_velocity, x, v, F, C, Jp, model: str = "neo_hookean"):
    """two_d/g2p.py:5-47: mutates ``x, v, F, C, Jp`` in place."""
    G = grid_velocity.shape[0]
    s = _solver(G - 1, len(x), dt, 1.0, 0.0, 1.0, 1.0, 1.0, 1.0, model, inv_dx)
    s.set_particles(x, v, F, C, Jp)
    R.grid_to_device(s, grid_velocity, np.zeros(grid_velocity.shape[:-1] + (1,)))
    s.g2p()
    s.check_errors()
    out = s.get_particles()
    x[...] = out["x"].double().cpu().numpy()
    v[...] = out["v"].double().cpu().numpy()
    F[...] = out["F"].double().cpu().numpy()
    C[...] = out["C"].double().cpu().numpy()
    Jp[...] = out["Jp"].double().cpu().numpy()
