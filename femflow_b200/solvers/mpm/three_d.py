"""3D phase functions with the reference's signatures
(femflow/solvers/mpm/three_d/{p2g,grid_op,g2p}.py), executed on the GPU.

These wrappers upload their NumPy arguments, run one CUDA phase and write the
results back in place -- the slow, exact drop-in used by the parity tests.  The
fast path keeps state on the device (femflow_b200.mpm.MpmSolver,
femflow_b200.simulation.mpm.MPMSimulation)."""
from __future__ import annotations

import numpy as np

from . import _runtime as R
from .particle import particles_to_soa, write_back_positions


# Which kernels the phase functions run: "direct" = one thread per particle on the caller's particle order
# (no binning); "production" = what ffmpm_substep runs -- cell binning, the warp-autonomous bulk P2G over
# cell runs, the tiled reordering G2P (results are returned in the caller's order through the id plane).
_KERNELS = "direct"


def set_kernels(kind: str) -> str:
    """Select "direct" or "production" kernels for the 3D phase functions; returns the previous setting."""
    global _KERNELS
    if kind not in ("direct", "production"):
        raise ValueError(f"unknown kernel path {kind!r}")
    prev, _KERNELS = _KERNELS, kind
    return prev


def _solver(res, n, inv_dx, dx, dt, volume, gravity, hardening, model):
    prod = _KERNELS == "production"
    return R.solver_for(3, res, n, inv_dx=float(inv_dx), dx=float(dx), dt=float(dt), volume=float(volume),
                        gravity=float(gravity), hardening=float(hardening), model=model,
                        p2g_mode="tiled" if prod else "scatter", reorder=prod)


def _upload(s, soa, v, F, C, Jp):
    s.set_particles(soa.pos, v, F, C, Jp if s.model == "snow" else None, soa.mass, soa.mu_0, soa.lambda_0)


def p2g(inv_dx, hardening, dx, dt, volume, grid_velocity, grid_mass, particles, v, F, C, Jp,
        model: str = "neo_hookean"):
    """three_d/p2g.py:14-80.  Accumulates momentum into ``grid_velocity`` and mass
    into ``grid_mass`` (both caller-owned, ``+=`` semantics)."""
    soa = particles_to_soa(particles)
    G = grid_velocity.shape[0]
    s = _solver(G - 1, len(soa), inv_dx, dx, dt, volume, 0.0, hardening, model)
    _upload(s, soa, v, F, C, Jp)
    s.clear_grid()
    if s.reorder:
        s.bin()
    s.p2g()
    s.check_errors()
    gv, gm = R.grid_from_device(s)
    grid_velocity += gv
    grid_mass += gm


def grid_op(grid_resolution, dx, dt, gravity, grid_velocity, grid_mass):
    """three_d/grid_op.py:5-47, in place on ``grid_velocity``."""
    s = _solver(grid_resolution, 0, 1.0 / dx, dx, dt, 1.0, gravity, 1.0, "neo_hookean")
    R.grid_to_device(s, grid_velocity, grid_mass)
    s.grid_op()
    gv, _ = R.grid_from_device(s)
    grid_velocity[...] = gv


def check_collision_points(points, normals, grid_resolution, dx, grid_velocity):
    """three_d/grid_op.py:50-67: zero ``grid_velocity`` behind the planes (point, normal), in place."""
    s = _solver(grid_resolution, 0, 1.0 / dx, dx, 1.0, 1.0, 0.0, 1.0, "neo_hookean")
    s.set_colliders(points, normals)
    try:
        R.grid_to_device(s, grid_velocity, np.zeros(grid_velocity.shape[:-1] + (1,)))
        s.collide()
        gv, _ = R.grid_from_device(s)
    finally:
        s.set_colliders(np.zeros((0, 3)), np.zeros((0, 3)))
    grid_velocity[...] = gv


def g2p(inv_dx, dt, grid_velocity, particles, v, F, C, Jp, model: str = "neo_hookean"):
    """three_d/g2p.py:9-59: mutates ``particles[i].pos``, ``v``, ``F``, ``C`` (and ``Jp`` for ``model="snow"``,
    g2p.py:48-58, with LAPACK's singular-vector signs: csrc/mpm_svd3.cuh) in place."""
    soa = particles_to_soa(particles)
    G = grid_velocity.shape[0]
    s = _solver(G - 1, len(soa), inv_dx, 1.0 / inv_dx, dt, 1.0, 0.0, 1.0, model)
    _upload(s, soa, v, F, C, Jp)
    R.grid_to_device(s, grid_velocity, np.zeros(grid_velocity.shape[:-1] + (1,)))
    if s.reorder:
        s.bin()
    s.g2p()
    s.check_errors()
    out = s.get_particles()
    write_back_positions(particles, out["x"].double().cpu().numpy())
    v[...] = out["v"].double().cpu().numpy()
    F[...] = out["F"].double().cpu().numpy()
    C[...] = out["C"].double().cpu().numpy()
    if model == "snow":
        Jp[...] = out["Jp"].double().cpu().numpy().reshape(np.shape(Jp))
