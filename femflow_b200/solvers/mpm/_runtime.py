"""Solver cache and array marshalling shared by the reference-signature wrappers."""
from __future__ import annotations

import os
from typing import Dict, Tuple

import numpy as np
import torch

from ...mpm import MpmSolver

_DTYPES = {"float32": torch.float32, "float64": torch.float64}
_default_dtype = _DTYPES[os.environ.get("FEMFLOW_B200_DTYPE", "float32")]
_cache: Dict[Tuple, MpmSolver] = {}


def set_default_dtype(dtype) -> None:
    """Storage/compute type of the wrappers: torch.float32 (product path, fp64 only
    inside the constitutive evaluation) or torch.float64 (validation build)."""
    global _default_dtype
    _default_dtype = _DTYPES[dtype] if isinstance(dtype, str) else dtype


def get_default_dtype():
    return _default_dtype


def clear_cache() -> None:
    for s in _cache.values():
        s.close()
    _cache.clear()


def solver_for(dim, res, n, **scalars) -> MpmSolver:
    """A cached solver with capacity >= n for this exact scalar configuration."""
    dtype = scalars.pop("dtype", None) or _default_dtype
    key = (dim, res, dtype, tuple(sorted(scalars.items())))
    s = _cache.get(key)
    if s is None or s.capacity < n:
        if s is not None:
            s.close()
        cap = max(1024, int(n))
        if len(_cache) > 16:
            clear_cache()
        s = MpmSolver(dim, res, capacity=cap, dtype=dtype, **scalars)
        _cache[key] = s
    return s


def grid_to_device(solver: MpmSolver, grid_velocity: np.ndarray, grid_mass: np.ndarray) -> None:
    """Reference grids ``(G,..,d)`` + ``(G,..,1)`` -> interleaved device grid."""
    g = solver.grid()
    d = solver.dim
    gv = torch.as_tensor(np.ascontiguousarray(grid_velocity), device=g.device).to(g.dtype)
    gm = torch.as_tensor(np.ascontiguousarray(grid_mass), device=g.device).to(g.dtype)
    g.zero_()
    if d == 3:
        g[..., :3] = gv
        g[..., 3] = gm[..., 0]
    else:
        g[:, :, 0, :2] = gv
        g[:, :, 0, 2] = gm[..., 0]


def grid_from_device(solver: MpmSolver):
    g = solver.grid().double().cpu().numpy()
    if solver.dim == 3:
        return g[..., :3], g[..., 3:4]
    return g[:, :, 0, :2], g[:, :, 0, 2:3]
