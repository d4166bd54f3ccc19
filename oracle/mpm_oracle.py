"""CPU oracle (fp64, vectorised NumPy) for the FEMFlow MLS-MPM substep.

TEST INFRASTRUCTURE ONLY.  This module restates, in batched NumPy, the algorithm
of the reference's numba loops so that the CUDA path can be checked at sizes the
serial numba code cannot reach.  Only ``tests/``, ``__graft_entry__.smoke()`` and
``bench.py``'s CPU-baseline / ``--impl reference`` legs may import it; the product
package ``femflow_b200`` never does.

Parity status: PINNED.  ``oracle/gen_golden.py`` runs the *real* reference
(``/root/reference/femflow``, numba) in the build container and stores its
outputs in ``tests/golden/*.npz``; ``tests/test_oracle_golden.py`` checks this
restatement against those files and against the reference's own 2D known-answer
values (solvers/mpm/tests/ground_truth_grid.txt, test_utils.py:7-15,
test_particle_to_grid.py:30-38, test_grid_to_particle.py:45-103).

State layout here is SoA (``x`` is an ``(N, d)`` array, per-particle ``mass``,
``mu0``, ``lam0`` are ``(N,)`` arrays) instead of the reference's typed list of
``Particle`` objects (solvers/mpm/particle.py:10-27); grids use the reference's
own layout: ``grid_velocity (G,)*d + (d,)`` and ``grid_mass (G,)*d + (1,)``,
C-order, with ``G = res + 1`` (solvers/mpm/mls_mpm.py:54-56).
"""
from __future__ import annotations

import numpy as np

__all__ = [
    "Ev_to_mu", "Ev_to_lambda", "bspline_weights", "base_and_fx", "cell_keys",
    "polar_decomp_2d", "polar_decomp_3d",
    "fixed_corotated_stress_2d", "fixed_corotated_stress_3d",
    "p2g_3d", "grid_op_3d", "check_collision_points", "g2p_3d", "solve_mls_mpm_3d",
    "p2g_2d", "grid_op_2d", "g2p_2d", "solve_mls_mpm_2d",
    "boundary_masks_2d",
]


# --------------------------------------------------------------------------- #
# numerics/fem.py:1-6
# --------------------------------------------------------------------------- #
def Ev_to_mu(E: float, v: float) -> float:
    return E / (2 * (1 + v))


def Ev_to_lambda(E: float, v: float) -> float:
    return E * v / ((1 + v) * (1 - 2 * v))


# --------------------------------------------------------------------------- #
# Indexing: three_d/p2g.py:50-55, two_d/p2g.py:50-53 (identical in g2p)
# --------------------------------------------------------------------------- #
def base_and_fx(x: np.ndarray, inv_dx: float):
    """``base = trunc(x*inv_dx - 0.5)`` (astype(int64) truncates toward zero),
    ``fx = x*inv_dx - base``."""
    xs = np.asarray(x, dtype=np.float64) * inv_dx
    base = (xs - 0.5).astype(np.int64)
    fx = xs - base.astype(np.float64)
    return base, fx


def bspline_weights(fx: np.ndarray) -> np.ndarray:
    """Quadratic B-spline weights, shape ``(3,) + fx.shape``
    (three_d/p2g.py:55)."""
    return np.stack([0.5 * (1.5 - fx) ** 2, 0.75 - (fx - 1) ** 2, 0.5 * (fx - 0.5) ** 2])


def cell_keys(base: np.ndarray, G: int) -> np.ndarray:
    """Linear C-order id of the base node, ``(i*G + j)*G + k`` (3D) or
    ``i*G + j`` (2D) -- the binning key of the CUDA path."""
    key = base[:, 0].astype(np.int64)
    for d in range(1, base.shape[1]):
        key = key * G + base[:, d]
    return key


def _check_oob(base: np.ndarray, G: int) -> None:
    """utils.py:138-150 called with res = grid.shape[0] = G for the base node
    and for every stencil offset 0..2 (three_d/p2g.py:51-52,70-71)."""
    if np.any(base < 0) or np.any(base + 2 >= G):
        raise RuntimeError("particle stencil leaves the grid")


# --------------------------------------------------------------------------- #
# numerics/linear_algebra.py:96-135
# --------------------------------------------------------------------------- #
def polar_decomp_2d(F: np.ndarray) -> np.ndarray:
    """Rotation factor only, with the reference's ``+1e-10`` in the norm
    (linear_algebra.py:108-113)."""
    x = F[:, 0, 0] + F[:, 1, 1]
    y = F[:, 1, 0] - F[:, 0, 1]
    scale = 1.0 / (np.sqrt(x * x + y * y) + 1e-10)
    c = x * scale
    s = y * scale
    R = np.empty_like(F)
    R[:, 0, 0] = c
    R[:, 0, 1] = -s
    R[:, 1, 0] = s
    R[:, 1, 1] = c
    return R


def polar_decomp_3d(F: np.ndarray) -> np.ndarray:
    """``R = U @ Vh`` from the SVD (linear_algebra.py:131-132)."""
    U, _, Vh = np.linalg.svd(F)
    return U @ Vh


# --------------------------------------------------------------------------- #
# solvers/mpm/utils.py:52-135
# --------------------------------------------------------------------------- #
def _fixed_corotated(F, R, inv_dx, mu, lam, dt, volume, mass, C):
    J = np.linalg.det(F)
    D_inv = 4 * inv_dx * inv_dx
    mu = np.asarray(mu, dtype=np.float64).reshape(-1, 1, 1)
    lam = np.asarray(lam, dtype=np.float64).reshape(-1, 1, 1)
    mass = np.asarray(mass, dtype=np.float64).reshape(-1, 1, 1)
    # utils.py:86 / :129 -- the scalar lambda*(J-1)*J is broadcast onto ALL entries.
    PF = (2 * mu * (F - R)) @ np.swapaxes(F, 1, 2) + lam * ((J - 1) * J).reshape(-1, 1, 1)
    stress = -(dt * volume) * (D_inv * PF)
    return stress + mass * C


def fixed_corotated_stress_2d(F, inv_dx, mu, lam, dt, volume, mass, C):
    return _fixed_corotated(F, polar_decomp_2d(F), inv_dx, mu, lam, dt, volume, mass, C)


def fixed_corotated_stress_3d(F, inv_dx, mu, lam, dt, volume, mass, C):
    return _fixed_corotated(F, polar_decomp_3d(F), inv_dx, mu, lam, dt, volume, mass, C)


def _hardening(mu0, lam0, hardening, Jp, model):
    """utils.py:7-49; three_d/p2g.py:57-61."""
    if model == "neo_hookean":
        e = hardening
    else:
        e = np.exp(hardening * (1.0 - np.asarray(Jp, dtype=np.float64).reshape(-1)))
    return np.asarray(mu0, dtype=np.float64) * e, np.asarray(lam0, dtype=np.float64) * e


def _scatter(grid_velocity, grid_mass, base, contrib_fn, d):
    """Sum all stencil contributions into the (caller-owned, accumulate-in-place)
    grids.  ``np.bincount`` on linear node ids replaces the reference's serial
    ``+=`` (three_d/p2g.py:75-80); summation order differs, values agree to
    round-off."""
    G = grid_velocity.shape[0]
    n_nodes = G ** d
    gv = grid_velocity.reshape(n_nodes, d)
    gm = grid_mass.reshape(n_nodes)
    offsets = np.stack(np.meshgrid(*([np.arange(3)] * d), indexing="ij"), -1).reshape(-1, d)
    for off in offsets:
        node = cell_keys(base + off, G)
        mom, m = contrib_fn(off)
        for c in range(d):
            gv[:, c] += np.bincount(node, weights=mom[:, c], minlength=n_nodes)
        gm += np.bincount(node, weights=m, minlength=n_nodes)


# --------------------------------------------------------------------------- #
# 3D phases
# --------------------------------------------------------------------------- #
def p2g_3d(inv_dx, hardening, dx, dt, volume, grid_velocity, grid_mass,
           x, mass, mu0, lam0, v, F, C, Jp, model="neo_hookean"):
    """three_d/p2g.py:14-80.  ``grid_velocity`` receives *momentum*."""
    n = len(x)
    if n == 0:
        return
    G = grid_velocity.shape[0]
    base, fx = base_and_fx(x, inv_dx)
    _check_oob(base, G)
    w = bspline_weights(fx)
    mass = np.broadcast_to(np.asarray(mass, dtype=np.float64), (n,))
    mu, lam = _hardening(np.broadcast_to(mu0, (n,)), np.broadcast_to(lam0, (n,)),
                         hardening, Jp, model)
    affine = fixed_corotated_stress_3d(F, inv_dx, mu, lam, dt, volume, mass, C)
    mv = v * mass[:, None]

    def contrib(off):
        dpos = (off - fx) * dx
        weight = w[off[0], :, 0] * w[off[1], :, 1] * w[off[2], :, 2]
        return weight[:, None] * (mv + np.einsum("nij,nj->ni", affine, dpos)), weight * mass

    _scatter(grid_velocity, grid_mass, base, contrib, 3)


def grid_op_3d(grid_resolution, dx, dt, gravity, grid_velocity, grid_mass):
    """three_d/grid_op.py:5-47.  Momentum -> velocity, gravity on axis 1,
    clamp to +-0.9*dx/dt, then per-axis sticky walls on *every* node."""
    R = grid_resolution
    v_allowed = dx * 0.9 / dt
    m = grid_mass[..., 0]
    act = m > 0
    gv = grid_velocity
    gv[act] /= m[act][:, None]
    gv[act, 1] += dt * gravity
    gv[act] = np.clip(gv[act], -v_allowed, v_allowed)
    boundary = 1
    idx = np.arange(grid_velocity.shape[0])
    wall = (idx < boundary) | (idx >= R - boundary)
    gv[wall, :, :, 0] = 0
    gv[:, wall, :, 1] = 0
    gv[:, :, wall, 2] = 0


def check_collision_points(points, normals, grid_resolution, dx, grid_velocity):
    """three_d/grid_op.py:50-67: zero the velocity of every node behind any of the planes
    (point, normal); the reference adds the scalar 1/|normal| to every component of the
    normal before the test (grid_op.py:59-60)."""
    G = grid_resolution + 1
    I = np.stack(np.meshgrid(*([np.arange(G)] * 3), indexing="ij"), -1).astype(np.float64)
    for point, normal in zip(np.asarray(points, dtype=np.float64), np.asarray(normals, dtype=np.float64)):
        normal = normal + (1.0 / np.sqrt(np.sum(np.square(normal))))
        offset = I * dx - point
        grid_velocity[offset @ normal < 0] = 0.0


def g2p_3d(inv_dx, dt, grid_velocity, x, v, F, C, Jp, model="neo_hookean"):
    """three_d/g2p.py:9-59.  Mutates x, v, F, C (and Jp for snow) in place."""
    n = len(x)
    if n == 0:
        return
    G = grid_velocity.shape[0]
    base, fx = base_and_fx(x, inv_dx)
    _check_oob(base, G)
    w = bspline_weights(fx)
    new_v = np.zeros((n, 3))
    new_C = np.zeros((n, 3, 3))
    for i in range(3):
        for j in range(3):
            for k in range(3):
                dpos = np.array((i, j, k)) - fx
                gv = grid_velocity[base[:, 0] + i, base[:, 1] + j, base[:, 2] + k]
                weight = w[i, :, 0] * w[j, :, 1] * w[k, :, 2]
                wgv = weight[:, None] * gv
                new_v += wgv
                new_C += 4 * inv_dx * wgv[:, :, None] * dpos[:, None, :]
    v[:] = new_v
    C[:] = new_C
    x += dt * new_v
    F_ = (np.eye(3) + dt * new_C) @ F
    if model == "snow":
        U, sig, Vh = np.linalg.svd(F_)
        sig = np.clip(sig, 1.0 - 2.5e-2, 1.0 + 7.5e-3)
        old_J = np.linalg.det(F_)
        # g2p.py:55 uses ``V.T`` where V is numpy's Vh: reproduce U @ S @ Vh^T.
        F_ = (U * sig[:, None, :]) @ np.swapaxes(Vh, 1, 2)
        det = np.linalg.det(F_) + 1e-10
        Jp[:, 0] = np.clip(Jp[:, 0] * old_J / det, 0.6, 20.0)
    F[:] = F_


def solve_mls_mpm_3d(res, inv_dx, hardening, dx, dt, volume, gravity,
                     x, mass, mu0, lam0, v, F, C, Jp, return_grids=False):
    """solvers/mpm/mls_mpm.py:40-79 on SoA state."""
    dres = res + 1
    grid_velocity = np.zeros((dres, dres, dres, 3))
    grid_mass = np.zeros((dres, dres, dres, 1))
    model = "neo_hookean"
    p2g_3d(inv_dx, hardening, dx, dt, volume, grid_velocity, grid_mass,
           x, mass, mu0, lam0, v, F, C, Jp, model)
    grid_momentum = grid_velocity.copy() if return_grids else None
    grid_op_3d(res, dx, dt, gravity, grid_velocity, grid_mass)
    g2p_3d(inv_dx, dt, grid_velocity, x, v, F, C, Jp, model)
    if return_grids:
        return grid_momentum, grid_mass, grid_velocity


# --------------------------------------------------------------------------- #
# 2D phases
# --------------------------------------------------------------------------- #
def p2g_2d(inv_dx, hardening, mu_0, lambda_0, mass, dx, dt, volume,
           grid_velocity, grid_mass, x, v, F, C, Jp, model="neo_hookean"):
    """two_d/p2g.py:11-76 (global material scalars, no bounds check)."""
    n = len(x)
    if n == 0:
        return
    base, fx = base_and_fx(x, inv_dx)
    w = bspline_weights(fx)
    mu, lam = _hardening(np.full(n, mu_0), np.full(n, lambda_0), hardening, Jp, model)
    massv = np.full(n, float(mass))
    affine = fixed_corotated_stress_2d(F, inv_dx, mu, lam, dt, volume, massv, C)
    mv = v * mass

    def contrib(off):
        dpos = (off - fx) * dx
        weight = w[off[0], :, 0] * w[off[1], :, 1]
        return weight[:, None] * (mv + np.einsum("nij,nj->ni", affine, dpos)), weight * massv

    _scatter(grid_velocity, grid_mass, base, contrib, 2)


def boundary_masks_2d(grid_resolution: int):
    """The f64 wall predicates of two_d/grid_op.py:18-23 as per-index masks:
    ``sticky[i]`` <=> ``i/R < 0.05 or i/R > 1-0.05`` (x axis), ``top[j]`` <=>
    ``j/R > 1-0.05``, ``floor[j]`` <=> ``j/R < 0.05``."""
    boundary = 0.05
    c = np.arange(grid_resolution + 1) / grid_resolution
    return (c < boundary) | (c > 1 - boundary), c > 1 - boundary, c < boundary


def grid_op_2d(grid_resolution, dt, gravity, grid_velocity, grid_mass):
    """two_d/grid_op.py:5-24.  Only nodes with mass > 0 are touched."""
    m = grid_mass[..., 0]
    act = m > 0
    gv = grid_velocity
    gv[act] /= m[act][:, None]
    gv[act, 1] += dt * gravity
    sticky_x, top_y, floor_y = boundary_masks_2d(grid_resolution)
    sticky = (sticky_x[:, None] | top_y[None, :]) & act
    gv[sticky] = 0.0
    fl = np.broadcast_to(floor_y[None, :], act.shape) & act
    gv[fl, 1] = np.maximum(0.0, gv[fl, 1])


def g2p_2d(inv_dx, dt, grid_velocity, x, v, F, C, Jp, model="neo_hookean"):
    """two_d/g2p.py:5-47.  The SVD round trip and the Jp update run for every
    model; ``F_ = U @ diag(sig) @ V.T`` with V = numpy's Vh (g2p.py:37-47)."""
    n = len(x)
    if n == 0:
        return
    base, fx = base_and_fx(x, inv_dx)
    w = bspline_weights(fx)
    new_v = np.zeros((n, 2))
    new_C = np.zeros((n, 2, 2))
    for i in range(3):
        for j in range(3):
            dpos = np.array((i, j)) - fx
            gv = grid_velocity[base[:, 0] + i, base[:, 1] + j]
            weight = w[i, :, 0] * w[j, :, 1]
            wgv = weight[:, None] * gv
            new_v += wgv
            new_C += 4 * inv_dx * wgv[:, :, None] * dpos[:, None, :]
    v[:] = new_v
    C[:] = new_C
    x += dt * new_v
    F_ = (np.eye(2) + dt * new_C) @ F
    U, sig, Vh = np.linalg.svd(F_)
    if model == "snow":
        sig = np.clip(sig, 1.0 - 2.5e-2, 1.0 + 7.5e-3)
    old_J = np.linalg.det(F_)
    F_ = (U * sig[:, None, :]) @ np.swapaxes(Vh, 1, 2)
    det = np.linalg.det(F_) + 1e-10
    Jp[:, 0] = np.clip(Jp[:, 0] * old_J / det, 0.6, 20.0)
    F[:] = F_


def solve_mls_mpm_2d(res, inv_dx, hardening, mu_0, lambda_0, mass, dx, dt, volume, gravity,
                     x, v, F, C, Jp, model="neo_hookean", return_grids=False):
    """The 2D driver the reference never wrote: zeroed grids, then
    two_d.p2g -> two_d.grid_op -> two_d.g2p."""
    G = res + 1
    grid_velocity = np.zeros((G, G, 2))
    grid_mass = np.zeros((G, G, 1))
    p2g_2d(inv_dx, hardening, mu_0, lambda_0, mass, dx, dt, volume,
           grid_velocity, grid_mass, x, v, F, C, Jp, model)
    grid_momentum = grid_velocity.copy() if return_grids else None
    grid_op_2d(res, dt, gravity, grid_velocity, grid_mass)
    g2p_2d(inv_dx, dt, grid_velocity, x, v, F, C, Jp, model)
    if return_grids:
        return grid_momentum, grid_mass, grid_velocity
