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
                  cells_x=None) -> Scene:
    """Block of ``cells**dim`` grid cells centred in the unit domain (``cells_x`` cells
    along axis 0 when given), ``ppc_side**dim`` jittered particles per cell
    (jitter U(-0.25, 0.25) * dx / ppc_side).  ``perturb`` adds v0 ~ N(0, 0.05) and
    F = I + N(0, 0.01) so the constitutive path does real work."""
    rng = np.random.default_rng(seed)
    dx = 1.0 / res
    shape = [cells_x or cells] + [cells] * (dim - 1)
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
