"""3D snow G2P (three_d/g2p.py:48-58) needs the SIGNS of LAPACK's singular vectors: ``U @ S @ Vh.T`` is not invariant
under (u_i, v_i) -> (-u_i, -v_i).  oracle/lapack_svd3.py restates DGESDD's operation sequence for 3x3 input and
femflow_b200/csrc/mpm_svd3.cuh is the device version; both are pinned here against ``np.linalg.svd`` itself (the
routine the reference calls) and against the reference's own outputs in tests/golden/snow3d.npz."""
import ctypes as C
import os

import numpy as np
import pytest

from oracle import lapack_svd3 as L
from oracle import mpm_oracle as O
from test_kernel_math_host import km, ptr  # noqa: F401  (module-scoped fixture: the host build of csrc/)

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def matrices(rng, n):
    """Near-identity (the snow regime), generic, badly scaled, graded, negative determinant."""
    out = []
    for k in range(n):
        kind = k % 6
        if kind == 0:
            A = np.eye(3) + 0.05 * rng.standard_normal((3, 3))
        elif kind == 1:
            A = rng.standard_normal((3, 3))
        elif kind == 2:
            A = np.diag(10 ** rng.uniform(-2, 2, 3)) @ (np.eye(3) + 0.1 * rng.standard_normal((3, 3)))
        elif kind == 3:
            q1, _ = np.linalg.qr(rng.standard_normal((3, 3)))
            q2, _ = np.linalg.qr(rng.standard_normal((3, 3)))
            A = q1 @ np.diag(10 ** rng.uniform(-3, 1, 3)) @ q2
        elif kind == 4:
            A = np.eye(3) + 1e-4 * rng.standard_normal((3, 3))
        else:
            A = np.diag([1.0, 1.0, -1.0]) + 0.05 * rng.standard_normal((3, 3))
        out.append(A)
    return np.array(out)


def vector_bar(s):
    """Singular vectors are conditioned by the relative gap; 1e-10 at a gap of 1."""
    gap = np.minimum(np.abs(s[:, 0] - s[:, 1]), np.abs(s[:, 1] - s[:, 2])) / s[:, 0]
    return 1e-10 / np.maximum(gap, 1e-6)


def test_python_restatement_matches_numpy_svd():
    A = matrices(np.random.default_rng(0), 3000)
    U, s, Vh = np.linalg.svd(A)
    bar = vector_bar(s)
    for k in range(len(A)):
        U2, s2, Vh2 = L.svd3(A[k])
        assert np.abs(s2 - s[k]).max() <= 1e-13 * s[k, 0]
        assert max(np.abs(U2 - U[k]).max(), np.abs(Vh2 - Vh[k]).max()) <= bar[k], k


def test_device_source_matches_numpy_svd(km):
    n = 120_000
    A = np.ascontiguousarray(matrices(np.random.default_rng(1), n))
    U = np.zeros((n, 3, 3)); S = np.zeros((n, 3)); Vh = np.zeros((n, 3, 3)); ok = np.zeros(n, np.int32)
    km.km_svd3_lapack(C.c_longlong(n), ptr(A), ptr(U), ptr(S), ptr(Vh), ptr(ok))
    assert ok.all()
    Ur, sr, Vhr = np.linalg.svd(A)
    bar = vector_bar(sr)
    assert (np.abs(S - sr).max(axis=1) <= 1e-13 * sr[:, 0]).all()
    err = np.maximum(np.abs(U - Ur).max(axis=(1, 2)), np.abs(Vh - Vhr).max(axis=(1, 2)))
    assert (err <= bar).all(), (np.argmax(err / bar), (err / bar).max())
    # the quantity the snow branch forms, and the one a sign error would change
    Q = (U * S[:, None, :]) @ np.swapaxes(Vh, 1, 2)
    Qr = (Ur * sr[:, None, :]) @ np.swapaxes(Vhr, 1, 2)
    assert (np.abs(Q - Qr).max(axis=(1, 2)) <= 4 * bar * sr[:, 0]).all()


def _snow3d():
    g = np.load(os.path.join(GOLD, "snow3d.npz"))
    res = int(g["res"])
    grid = np.zeros(tuple(g["grid_momentum_shape"]))
    grid.reshape(-1, 3)[g["grid_momentum_idx"]] = g["grid_momentum_val"]
    mass = np.zeros(tuple(g["grid_mass_shape"]))
    mass.reshape(-1, 1)[g["grid_mass_idx"]] = g["grid_mass_val"]
    return g, res, grid, mass


def test_snow_return_map_matches_reference_outputs(km):
    """The reference's own F_out / Jp_out (numba + LAPACK, oracle/gen_golden.py) from F, the new C and Jp."""
    g, res, grid, mass = _snow3d()
    O.grid_op_3d(res, float(g["dx"]), float(g["dt"]), float(g["gravity"]), grid, mass)
    x, v, F, Cm, Jp = (g[k].copy() for k in ("x", "v", "F", "C", "Jp"))
    Fold = F.copy()
    O.g2p_3d(float(g["inv_dx"]), float(g["dt"]), grid, x, v, F, Cm, Jp, "snow")       # numpy's own SVD
    np.testing.assert_allclose(F, g["F_out"], rtol=0, atol=1e-11)
    n = len(x)
    Fout = np.zeros((n, 9)); jp_out = np.zeros(n)
    km.km_snow_return_map3(C.c_longlong(n), C.c_double(float(g["dt"])), ptr(np.ascontiguousarray(Fold.reshape(n, 9))),
                           ptr(np.ascontiguousarray(g["C_out"].reshape(n, 9))), ptr(np.ascontiguousarray(g["Jp"][:, 0])),
                           ptr(Fout), ptr(jp_out))
    np.testing.assert_allclose(Fout.reshape(n, 3, 3), g["F_out"], rtol=0, atol=1e-10)
    np.testing.assert_allclose(jp_out, g["Jp_out"][:, 0], rtol=0, atol=1e-10)
    # fp32 storage: F and C rounded to float32 before the fp64 return map (what the fp32 build's kernel sees)
    F32 = Fold.astype(np.float32).astype(np.float64).reshape(n, 9)
    C32 = g["C_out"].astype(np.float32).astype(np.float64).reshape(n, 9)
    km.km_snow_return_map3(C.c_longlong(n), C.c_double(float(g["dt"])), ptr(np.ascontiguousarray(F32)), ptr(np.ascontiguousarray(C32)),
                           ptr(np.ascontiguousarray(g["Jp"][:, 0])), ptr(Fout), ptr(jp_out))
    assert np.abs(Fout.reshape(n, 3, 3) - g["F_out"]).max() < 1e-7 and np.abs(jp_out - g["Jp_out"][:, 0]).max() < 1e-6
