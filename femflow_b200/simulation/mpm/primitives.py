"""Scene point generators (reference femflow/simulation/mpm/primitives.py:8-76 and
numerics/geometry.py:101-116), vectorised: the reference evaluates the implicit
function point by point in Python."""
from __future__ import annotations

from typing import Callable, Tuple, Union

import numpy as np


def grid(res: np.ndarray) -> np.ndarray:
    """numerics/geometry.py:101-116: lattice of prod(res) points in [0,1]^d, axis 0
    fastest; coordinate = index / (res-1) (index / res when res == 1)."""
    res = np.asarray(res, dtype=np.int64)
    axes = [np.arange(r) / (r - 1 if r != 1 else r) for r in res]
    mesh = np.meshgrid(*axes[::-1], indexing="ij")       # slowest axis first
    return np.stack([m.reshape(-1) for m in mesh[::-1]], axis=1).astype(np.float64)


def primitive(k: float, t: float, pos: np.ndarray):
    two_pi = (2.0 * np.pi) / k
    x, y, z = pos.T
    return np.cos(two_pi * x) + np.cos(two_pi * y) + np.cos(two_pi * z) - t


def gyroid(k: float, t: float, pos: np.ndarray):
    two_pi = (2.0 * np.pi) / k
    x, y, z = pos.T
    return (np.sin(two_pi * x) * np.cos(two_pi * y) + np.sin(two_pi * y) * np.cos(two_pi * z)
            + np.sin(two_pi * z) * np.cos(two_pi * x) - t)


def diamond(k: float, t: float, pos: np.ndarray):
    two_pi = (2.0 * np.pi) / k
    x, y, z = pos.T
    sx, sy, sz = np.sin(two_pi * x), np.sin(two_pi * y), np.sin(two_pi * z)
    cx, cy, cz = np.cos(two_pi * x), np.cos(two_pi * y), np.cos(two_pi * z)
    return sx * sy * sz + sx * cy * cz + cx * sy * cz + cx * cy * sz - t


_FNS = {"gyroid": gyroid, "diamond": diamond, "primitive": primitive}


_KINDS = {"gyroid": 0, "diamond": 1, "primitive": 2}


def generate_implicit_points_gpu(implicit_fn: str, k: float, t: float, res: int, device="cuda:0"):
    """``generate_implicit_points`` on the GPU (csrc/mpm_scene.cuh through ``ffmpm_gen_implicit_points``): the res^3
    lattice is never materialised, the selected points come back as an ``(N, 3)`` float64 device tensor in the
    reference's row order.  Two passes over the lattice: count, then write."""
    import ctypes as C

    import torch

    from ... import _native as N
    if implicit_fn not in _KINDS:
        raise ValueError("Invalid implicit function specified")
    lib, dev = N.lib(), torch.device(device)
    with torch.cuda.device(dev):
        nbytes = lib.ffmpm_scene_scratch_bytes(int(res))
        if nbytes < 0:
            N.check(int(nbytes))
        scratch = torch.zeros(nbytes, dtype=torch.uint8, device=dev)
        stream = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
        args = (_KINDS[implicit_fn], float(k), float(t), int(res), scratch.data_ptr())
        N.check(lib.ffmpm_gen_implicit_points(*args, None, 0, stream))
        count = int(scratch[:8].view(torch.int64).item())
        out = torch.empty((count, 3), dtype=torch.float64, device=dev)
        if count:
            N.check(lib.ffmpm_gen_implicit_points(*args, out.data_ptr(), count, stream))
    return out


def generate_cube_points_gpu(xb, yb, zb, res: int = 10, device="cuda:0"):
    """``generate_cube_points`` on the GPU: ``(res^3, 3)`` float64 device tensor, rows [z, y, x] with x fastest."""
    import ctypes as C

    import torch

    from ... import _native as N
    lib, dev = N.lib(), torch.device(device)
    bounds = (C.c_double * 6)(float(xb[0]), float(xb[1]), float(yb[0]), float(yb[1]), float(zb[0]), float(zb[1]))
    with torch.cuda.device(dev):
        out = torch.empty((int(res) ** 3, 3), dtype=torch.float64, device=dev)
        N.check(lib.ffmpm_gen_cube_points(bounds, int(res), out.data_ptr(), C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)))
    return out


def generate_implicit_points(implicit_fn: Union[Callable, str], k: float, t: float, res: int) -> np.ndarray:
    """primitives.py:46-61: lattice points with ``f(p) - t > t`` (the reference
    subtracts t inside the function and compares against t again)."""
    if isinstance(implicit_fn, str):
        if implicit_fn not in _FNS:
            raise ValueError("Invalid implicit function specified")
        implicit_fn = _FNS[implicit_fn]
    g = grid(np.array((res, res, res)))
    inside = implicit_fn(k, t, g)
    return g[inside > t]


def generate_cube_points(xb: Tuple[float, float], yb: Tuple[float, float], zb: Tuple[float, float],
                         res: int = 10) -> np.ndarray:
    """primitives.py:64-76: note the reference's axis order -- rows are [z, y, x]
    with x fastest."""
    x = np.linspace(*xb, num=res)
    y = np.linspace(*yb, num=res)
    z = np.linspace(*zb, num=res)
    L, R, Cc = np.meshgrid(z, y, x, indexing="ij")
    return np.stack([L.reshape(-1), R.reshape(-1), Cc.reshape(-1)], axis=1).astype(np.float64)
