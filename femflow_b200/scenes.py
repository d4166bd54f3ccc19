"""Synthetic particle blocks of the BASELINE.json shapes (SURVEY 8d "Synthetic inputs").

Generated with ``numpy.random.default_rng(seed)`` on the host, float32-rounded, in
the reference's AoS layouts (``x (N, d)``, ``F (N, d, d)``).  Particles are emitted
cell by cell in C order, ``ppc_side**d`` jittered particles per cell.
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np


@dataclass
class Scene:
    dim: int
    res: int
    dt: float
    volume: float
    gravity: float
    hardening: float
    x: np.ndarray
    v: np.ndarray
    F: np.ndarray
    C: np.ndarray
    mass: float
    mu_0: float
    lambda_0: float
    active_nodes: int
    name: str

    @property
    def n(self) -> int:
        return len(self.x)


def _lame(E, nu):
    return E / (2 * (1 + nu)), E * nu / ((1 + nu) * (1 - 2 * nu))


def elastic_block(dim: int, res: int, cells: int, ppc_side: int = 2, seed: int = 0, E: float = 1e4,
                  nu: float = 0.2, rho: float = 1.0, perturb: bool = True, origin_cell=None,
                  cells_x=None, shape=None) -> Scene:
    """Block of ``cells**dim`` grid cells centred in the unit domain (``cells_x`` cells
    along axis 0 when given, or an explicit per-axis ``shape``), ``ppc_side**dim`` jittered particles per cell
    (jitter U(-0.25, 0.25) * dx / ppc_side).  ``perturb`` adds v0 ~ N(0, 0.05) and
    F = I + N(0, 0.01) so the constitutive path does real work."""
    rng = np.random.default_rng(seed)
    dx = 1.0 / res
    shape = list(shape) if shape is not None else [cells_x or cells] + [cells] * (dim - 1)
    if origin_cell is None:
        origin_cell = [(res - c) // 2 for c in shape]
    sub = (np.arange(ppc_side) + 0.5) / ppc_side
    cell_axes = [np.arange(c, dtype=np.float32) + o for c, o in zip(shape, origin_cell)]
    # cell-major ordering: all particles of a cell are consecutive
    grids = np.meshgrid(*cell_axes, *([sub] * dim), indexing="ij")
    cell = np.stack([g.reshape(-1) for g in grids[:dim]], -1)
    off = np.stack([g.reshape(-1) for g in grids[dim:]], -1)
    n = len(cell)
    jitter = rng.uniform(-0.25, 0.25, size=(n, dim)).astype(np.float32) / ppc_side
    x = ((cell + off + jitter) * np.float32(dx)).astype(np.float32)
    if perturb:
        v = rng.normal(0, 0.05, size=(n, dim)).astype(np.float32)
        F = (np.eye(dim, dtype=np.float32) + rng.normal(0, 0.01, size=(n, dim, dim)).astype(np.float32))
    else:
        v = np.zeros((n, dim), dtype=np.float32)
        F = np.tile(np.eye(dim, dtype=np.float32), (n, 1, 1))
    C = np.zeros((n, dim, dim), dtype=np.float32)
    volume = float(np.float32((dx / ppc_side) ** dim))
    mu, lam = _lame(E, nu)
    c_wave = np.sqrt((lam + 2 * mu) / rho)
    dt = float(np.float32(0.2 * dx / c_wave))
    active = int(np.prod([c + 3 for c in shape]))
    return Scene(dim, res, dt, volume, -9.8, 1.0, x, v, F, C, float(np.float32(rho * volume)),
                 float(np.float32(mu)), float(np.float32(lam)), active,
                 f"{dim}D elastic block {'x'.join(str(c) for c in shape)} cells x {ppc_side ** dim} ppc, res {res}")


def config_2d_1m(seed: int = 0) -> Scene:
    """BASELINE configs[1]: 2D, 1 048 576 particles, 1024^2 grid."""
    return elastic_block(2, 1024, 512, 2, seed)


def config_3d_16m(seed: int = 0) -> Scene:
    """BASELINE configs[2]: 3D, 16 777 216 particles, 256^3 grid."""
    return elastic_block(3, 256, 128, 2, seed)


def dam_break_slab(world: int, rank: int, res: int = 256, n_total: int = 33_554_432, seed: int = 0,
                   x_range=None, late: bool = False) -> Scene:
    """BASELINE configs[4]: a tall soft column at one x-end of a (res*world) x res x res
    domain, ~8.9 particles per cell, that collapses onto the floor (high atomic contention
    near the floor, strong load imbalance between slabs).  Returns only the particles whose
    x lies in ``x_range`` (this rank's owned interval); all ranks draw from the same
    deterministic per-cell streams, so the union over ranks is one scene."""
    dx = 1.0 / res
    L = float(world)
    x0, x1 = 0.05 * L, 0.20 * L
    y0, y1, z0, z1 = 0.02, 0.92, 0.40, 0.60
    if late:
        # the LATE window of the same scene, as an initial condition: the same particles spread over the floor
        # (~27 per cell: the dense, contended cells of a settled column; every slab loaded).  Running the column until it
        # has spread is not an option: under the reference's own stress formula (lambda (J-1) J added to ALL entries,
        # utils.py:129) the soft column goes unstable after a few thousand substeps -- on one GPU as on eight.
        x0, x1 = 0.05 * L, 0.95 * L
        y0, y1 = 0.02, 0.07
    cx0, cx1 = int(round(x0 * res)), int(round(x1 * res))
    cy0, cy1 = int(round(y0 * res)), int(round(y1 * res))
    cz0, cz1 = int(round(z0 * res)), int(round(z1 * res))
    n_cells = (cx1 - cx0) * (cy1 - cy0) * (cz1 - cz0)
    ppc = max(1, int(round(n_total / n_cells)))
    if x_range is not None:
        lo = max(cx0, int(np.floor(x_range[0] * res)))
        hi = min(cx1, int(np.ceil(x_range[1] * res)))
    else:
        lo, hi = cx0, cx1
    xs = np.arange(lo, max(lo, hi), dtype=np.float32)
    ys = np.arange(cy0, cy1, dtype=np.float32)
    zs = np.arange(cz0, cz1, dtype=np.float32)
    rng = np.random.default_rng(seed * 1000003 + rank)
    X, Y, Z = np.meshgrid(xs, ys, zs, indexing="ij")
    cell = np.stack([X.reshape(-1), Y.reshape(-1), Z.reshape(-1)], -1)
    cell = np.repeat(cell, ppc, axis=0)
    pos = ((cell + rng.uniform(0.02, 0.98, size=cell.shape).astype(np.float32)) * np.float32(dx)).astype(np.float32)
    if x_range is not None and len(pos):
        keep = (pos[:, 0] >= x_range[0]) & (pos[:, 0] < x_range[1])
        pos = pos[keep]
    n = len(pos)
    E, nu, rho = 1e3, 0.3, 1.0
    mu, lam = _lame(E, nu)
    volume = float(np.float32(dx ** 3 / ppc))
    c_wave = np.sqrt((lam + 2 * mu) / rho)
    dt = float(np.float32(0.2 * dx / c_wave))
    return Scene(3, res, dt, volume, -9.8, 1.0, pos, np.zeros((n, 3), np.float32),
                 np.tile(np.eye(3, dtype=np.float32), (n, 1, 1)), np.zeros((n, 3, 3), np.float32),
                 float(np.float32(rho * volume)), float(np.float32(mu)), float(np.float32(lam)), 0,
                 f"3D dam break{' (late: spread over the floor)' if late else ''}, {ppc} ppc "
                 f"{'layer' if late else 'column'} {cx1 - cx0}x{cy1 - cy0}x{cz1 - cz0} cells on {res * world}x{res}x{res}")
