"""The kernels' own per-particle math on the CPU.

Cell indexing, B-spline weights, both constitutive evaluations (fp32 perturbation series and fp64
Newton polar), the P2G payload and the separable G2P stencil sums are ``__host__ __device__``
functions of femflow_b200/csrc; tests/native/kernel_math_host.cu compiles those very sources for the
host and this file checks them against LAPACK / the NumPy oracle -- the same bars as the GPU parity
tests, without a GPU.  (The warp-level scatter, the binning and the atomics only exist on the device:
tests/test_gpu_parity.py.)"""
import ctypes as C
import os
import shutil
import subprocess

import numpy as np
import pytest

from oracle import mpm_oracle as O

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "native", "kernel_math_host.cu")
CSRC = os.path.join(os.path.dirname(HERE), "femflow_b200", "csrc")
LIB = os.path.join(HERE, "native", "_build", "libkernel_math_host.so")


def build_host_harness(force: bool = False) -> str:
    deps = [SRC] + [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".cuh")]
    if not force and os.path.exists(LIB) and all(os.path.getmtime(d) <= os.path.getmtime(LIB) for d in deps):
        return LIB
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        raise FileNotFoundError("nvcc not available")
    os.makedirs(os.path.dirname(LIB), exist_ok=True)
    cmd = [nvcc, "-O2", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-diag-suppress", "20013,20011,20015",
           "-Xcompiler", "-fPIC", "-shared", "-o", LIB, SRC]
    proc = subprocess.run(cmd, capture_output=True, text=True)
    assert proc.returncode == 0, proc.stderr
    return LIB


@pytest.fixture(scope="module")
def km():
    try:
        return C.CDLL(build_host_harness())
    except FileNotFoundError as e:
        pytest.skip(str(e))


def ptr(a):
    return a.ctypes.data_as(C.c_void_p)


def prepare3(km, dtype, res, x, v, Cm, F, mass, mu, lam, dt, volume, hardening=1.0, fp32_stress=1, index_fp32=None, jp=None,
             model=0):
    n = len(x)
    np_dt = np.float32 if dtype == "f32" else np.float64
    arrs = [np.ascontiguousarray(a, dtype=np_dt) for a in (x, v, Cm.reshape(n, 9), F.reshape(n, 9), mass, mu, lam)]
    base = np.zeros((n, 3), np.int32); fx = np.zeros((n, 3), np_dt); aff = np.zeros((n, 9), np_dt)
    mv = np.zeros((n, 3), np_dt); m = np.zeros(n, np_dt); ok = np.zeros(n, np.int32)
    if index_fp32 is None:
        index_fp32 = int(res & (res - 1) == 0)
    jp_arr = None if jp is None else np.ascontiguousarray(jp, dtype=np.float64)
    fn = getattr(km, f"km_prepare3_{dtype}")
    fn(C.c_int(res), C.c_int(res + 1), C.c_double(float(res)), C.c_double(1.0 / res), C.c_double(dt), C.c_double(volume),
       C.c_double(hardening), C.c_int(model), C.c_int(fp32_stress), C.c_int(index_fp32), C.c_longlong(n),
       *(ptr(a) for a in arrs), ptr(jp_arr) if jp_arr is not None else None, ptr(base), ptr(fx), ptr(aff), ptr(mv), ptr(m),
       ptr(ok))
    return base, fx, aff.reshape(n, 3, 3), mv, m, ok.astype(bool)


def f32(a):
    return np.asarray(a, dtype=np.float32).astype(np.float64)


# --------------------------------------------------------------------------- #
@pytest.mark.parametrize("res", [37, 64, 80, 256, 1024])
def test_cell_indexing_is_bit_exact(km, res):
    """base = trunc(x*inv_dx - 0.5) (quirks 1, 11): positions one fp32 ulp either side of every cell
    boundary, plus the (-1, 0) band that truncation sends to cell 0 -- fp64 evaluation for any R, and
    the fp32 evaluation the kernels switch to when inv_dx is a power of two."""
    k = np.arange(0, min(res, 300), dtype=np.float64)
    edge = ((k + 0.5) / res).astype(np.float32)
    xs = np.concatenate([edge, np.nextafter(edge, np.float32(-1)), np.nextafter(edge, np.float32(2)),
                         np.float32([0.0, 1e-9, -1e-9, 0.2 / res, 0.49999 / res]),
                         np.random.default_rng(res).uniform(0, 1, 4000).astype(np.float32)])
    want_base, want_fx = O.base_and_fx(xs.astype(np.float64)[:, None], float(res))
    pow2 = res & (res - 1) == 0
    for index_fp32 in ([0, 1] if pow2 else [0]):
        base = np.zeros(len(xs), np.int32); fx = np.zeros(len(xs), np.float32)
        km.km_base_fx_f32(C.c_double(float(res)), C.c_int(index_fp32), C.c_longlong(len(xs)), ptr(xs), ptr(base), ptr(fx))
        assert np.array_equal(base, want_base[:, 0]), (res, index_fp32)
        assert np.array_equal(fx, want_fx[:, 0].astype(np.float32)), (res, index_fp32)   # fx is exact, then rounded once
    xd = xs.astype(np.float64)
    base = np.zeros(len(xs), np.int32); fxd = np.zeros(len(xs), np.float64)
    km.km_base_fx_f64(C.c_double(float(res)), C.c_longlong(len(xs)), ptr(xd), ptr(base), ptr(fxd))
    assert np.array_equal(base, want_base[:, 0]) and np.array_equal(fxd, want_fx[:, 0])


@pytest.mark.parametrize("strain", [0.0, 1e-6, 1e-4, 1.5e-3, 1e-2, 4e-2, 0.3])
def test_fp32_stress_tiers_against_lapack(km, strain):
    """affine = stress + mass*C of three_d/p2g.py:57-65 (utils.py:120-135) from the fp32 kernel path: the
    left-form series tiers (degree 2 .. 5 by the Frobenius norm of G = F F^T - I) and the fp64 Newton fallback beyond, each
    to 1e-5 of the batch's largest STRESS entry (C = 0, so nothing hides the stress) -- the bar of
    test_stress_accuracy_across_series_tiers on the GPU."""
    rng = np.random.default_rng(11)
    res, n = 32, 4000
    dx = 1.0 / res
    vol = float(f32((dx / 2) ** 3))
    x = f32(rng.uniform(0.25, 0.75, size=(n, 3)))
    F = f32(np.eye(3) + strain * rng.uniform(-1, 1, size=(n, 3, 3)))
    Z = np.zeros((n, 3, 3))
    mass = np.full(n, vol); mu = np.full(n, f32(4166.67)); lam = np.full(n, f32(2777.78))
    want = O.fixed_corotated_stress_3d(F, float(res), mu, lam, 1e-4, vol, mass, Z)
    base, fx, aff, mv, m, ok = prepare3(km, "f32", res, x, np.zeros((n, 3)), Z, F, mass, mu, lam, 1e-4, vol)
    assert ok.all()
    scale = np.abs(want).max()
    if strain == 0.0:
        assert np.all(aff == 0.0) and scale == 0.0            # F = I: exactly no stress (every reference scene starts so)
    else:
        assert np.abs(aff - want).max() / scale < 1e-5, strain
    # the tier decision itself: the series declines beyond ||G||_F = 0.15 and the caller goes to fp64
    A = np.zeros(9, np.float32)
    Ff = np.ascontiguousarray(F[0].reshape(9), dtype=np.float32)
    Cf = np.zeros(9, np.float32)
    km.km_affine3_f32.restype = C.c_int
    took = km.km_affine3_f32(ptr(Ff), ptr(Cf), C.c_float(1.0), C.c_float(1.0), C.c_float(1.0), C.c_float(1.0), ptr(A))
    G = F[0] @ F[0].T - np.eye(3)
    assert bool(took) == bool(np.linalg.norm(G) < 0.15 * (1 - 1e-4)) or abs(np.linalg.norm(G) - 0.15) < 1e-4


def test_fp64_stress_and_polar_against_lapack(km):
    """fp64 build / large-strain fallback: Newton polar factor == U @ Vh of the SVD (also for det F < 0,
    where both give the det = -1 factor), stress to 1e-12; F = I returns R = I exactly."""
    rng = np.random.default_rng(5)
    n = 3000
    F = np.eye(3) + rng.uniform(-1, 1, size=(n, 3, 3)) * rng.choice([1e-6, 1e-3, 0.1, 0.5, 1.5], size=(n, 1, 1))
    F[0] = np.eye(3)
    F[1] = np.diag([1.0, 1.0, -1.0]) @ F[1]                                   # a reflection
    R = np.zeros((n, 9)); det = np.zeros(n)
    km.km_polar3(C.c_longlong(n), ptr(np.ascontiguousarray(F.reshape(n, 9))), ptr(R), ptr(det))
    R = R.reshape(n, 3, 3)
    want = O.polar_decomp_3d(F)
    cond = np.linalg.cond(F)
    sel = cond < 1e3                                                          # R is ill-defined near singular F
    assert sel.sum() > 0.9 * n and (np.linalg.det(F) < 0).sum() > 10
    assert np.abs(R[sel] - want[sel]).max() < 1e-11
    assert np.array_equal(R[0], np.eye(3))
    assert np.allclose(det, np.linalg.det(F), rtol=1e-12, atol=1e-14)
    res = 16
    x = np.full((n, 3), 0.5)
    Cm = rng.normal(0, 0.3, size=(n, 3, 3))
    mass = rng.uniform(0.5, 2, n); mu = rng.uniform(10, 50, n); lam = rng.uniform(10, 50, n)
    want_aff = O.fixed_corotated_stress_3d(F, float(res), 0.7 * mu, 0.7 * lam, 1e-3, 0.5, mass, Cm)
    base, fx, aff, mv, m, ok = prepare3(km, "f64", res, x, rng.normal(size=(n, 3)), Cm, F, mass, mu, lam, 1e-3, 0.5,
                                        hardening=0.7)
    assert ok.all()
    assert np.abs(aff[sel] - want_aff[sel]).max() / np.abs(want_aff[sel]).max() < 1e-12


def test_p2g_payload_and_oob_flag(km):
    """mass*v, mass, base, fx as the scatter consumes them; particles whose stencil leaves [0, R] are
    flagged instead of written (three_d/p2g.py:51-52, utils.py:138-150); snow hardening multiplier."""
    rng = np.random.default_rng(2)
    res, n = 16, 500
    x = f32(rng.uniform(0.1, 0.9, size=(n, 3)))
    x[0] = [0.99, 0.5, 0.5]          # base 15 -> stencil reaches node 17 > R
    x[1] = [0.5, -0.2, 0.5]          # base -3
    x[2] = [np.nan, 0.5, 0.5]
    x[3] = [0.0, 0.5, 0.5]           # x*inv_dx - 0.5 in (-1, 0): truncates to cell 0, legal (quirk 1)
    v = f32(rng.normal(size=(n, 3)))
    F = f32(np.eye(3) + 0.01 * rng.normal(size=(n, 3, 3)))
    Cm = f32(rng.normal(0, 0.2, size=(n, 3, 3)))
    mass = f32(rng.uniform(0.5, 2, n)); mu = f32(rng.uniform(10, 50, n)); lam = f32(rng.uniform(10, 50, n))
    base, fx, aff, mv, m, ok = prepare3(km, "f32", res, x, v, Cm, F, mass, mu, lam, 1e-3, 0.5)
    assert not ok[0] and not ok[1] and not ok[2] and ok[3] and ok[4:].all()
    wb, wfx = O.base_and_fx(x[3:], float(res))
    assert np.array_equal(base[3:], wb) and base[3, 0] == 0
    assert np.array_equal(fx[3:], wfx.astype(np.float32))
    assert np.array_equal(m[3:], mass[3:].astype(np.float32))
    assert np.array_equal(mv[3:], (mass[3:, None].astype(np.float32) * v[3:].astype(np.float32)))
    want = O.fixed_corotated_stress_3d(F[3:], float(res), mu[3:], lam[3:], 1e-3, 0.5, mass[3:], Cm[3:])
    assert np.abs(aff[3:] - want).max() / np.abs(want).max() < 1e-5
    # snow (three_d/p2g.py:59-61): mu, lam scaled by exp(h (1 - Jp))
    jp = rng.uniform(0.8, 1.2, n)
    h = 3.0
    _, _, aff_s, _, _, ok_s = prepare3(km, "f64", res, x, v, Cm, F, mass, mu, lam, 1e-3, 0.5, hardening=h, jp=jp, model=1)
    e = np.exp(h * (1 - jp[3:]))
    want_s = O.fixed_corotated_stress_3d(F[3:], float(res), mu[3:] * e, lam[3:] * e, 1e-3, 0.5, mass[3:], Cm[3:])
    assert np.abs(aff_s[3:] - want_s).max() / np.abs(want_s).max() < 1e-12


@pytest.mark.parametrize("dtype,tol", [("f32", 2e-6), ("f64", 1e-14)])
def test_separable_g2p_stencil_sums(km, dtype, tol):
    """v = sum w gv, C = sum (w gv) (x) dpos (three_d/g2p.py:31-43) folded z -> y -> x against the
    node-by-node sums of the reference."""
    rng = np.random.default_rng(8)
    n = 2000
    np_dt = np.float64 if dtype == "f64" else np.float32
    f = rng.uniform(0.5, 1.5, size=(n, 3)).astype(np_dt)
    gv = rng.normal(size=(n, 3, 3, 3, 3)).astype(np_dt)
    v = np.zeros((n, 3), np_dt); c = np.zeros((n, 9), np_dt)
    getattr(km, f"km_g2p_accumulate3_{dtype}")(C.c_longlong(n), ptr(f), ptr(np.ascontiguousarray(gv)), ptr(v), ptr(c))
    fd = f.astype(np.float64)
    w = O.bspline_weights(fd)                                     # (3, n, 3): [offset, particle, axis]
    want_v = np.zeros((n, 3)); want_c = np.zeros((n, 3, 3))
    for i in range(3):
        for j in range(3):
            for k in range(3):
                wt = w[i, :, 0] * w[j, :, 1] * w[k, :, 2]
                wgv = wt[:, None] * gv[:, i, j, k].astype(np.float64)
                dpos = np.array((i, j, k)) - fd
                want_v += wgv
                want_c += wgv[:, :, None] * dpos[:, None, :]
    assert np.abs(v - want_v).max() / np.abs(want_v).max() < tol
    assert np.abs(c.reshape(n, 3, 3) - want_c).max() / np.abs(want_c).max() < tol


def test_2d_stress_and_svd_roundtrip(km):
    """2D constitutive update with the reference's +1e-10 (quirks 3, 12) and the U diag(s) Vh^T round trip
    of two_d/g2p.py:37-43 including det F < 0 (where the wrong transpose matters) and the snow clamp."""
    rng = np.random.default_rng(4)
    n = 4000
    F = np.eye(2) + rng.uniform(-1, 1, size=(n, 2, 2)) * rng.choice([1e-8, 1e-3, 0.2, 1.2], size=(n, 1, 1))
    F[0] = np.eye(2)
    Cm = rng.normal(0, 0.3, size=(n, 2, 2))
    mu, lam, mass, dt, vol, inv_dx = 4166.67, 2777.78, 1.0, 1e-4, 1.0, 80.0
    k = dt * vol * 4 * inv_dx * inv_dx
    A = np.zeros((n, 4))
    km.km_affine2(C.c_longlong(n), ptr(np.ascontiguousarray(F.reshape(n, 4))), ptr(np.ascontiguousarray(Cm.reshape(n, 4))),
                  C.c_double(mu), C.c_double(lam), C.c_double(mass), C.c_double(k), ptr(A))
    want = O.fixed_corotated_stress_2d(F, inv_dx, np.full(n, mu), np.full(n, lam), dt, vol, np.full(n, mass), Cm)
    assert np.abs(A.reshape(n, 2, 2) - want).max() / np.abs(want).max() < 1e-13
    assert abs(A[0, 0] - want[0, 0, 0]) < 1e-12 * abs(want[0, 0, 0]) and want[0, 0, 0] != 0     # the 1e-10 stress at F = I
    sel = np.linalg.cond(F) < 1e4
    assert (np.linalg.det(F[sel]) < 0).sum() > 20
    for snow in (0, 1):
        G = np.zeros((n, 4)); det = np.zeros(n)
        km.km_svd_roundtrip2(C.c_longlong(n), ptr(np.ascontiguousarray(F.reshape(n, 4))), C.c_int(snow), ptr(G), ptr(det))
        U, sig, Vh = np.linalg.svd(F)
        if snow:
            sig = np.clip(sig, 1.0 - 2.5e-2, 1.0 + 7.5e-3)
        wantG = (U * sig[:, None, :]) @ np.swapaxes(Vh, 1, 2)
        assert np.abs(G.reshape(n, 2, 2)[sel] - wantG[sel]).max() < 1e-9, snow
        assert np.abs(det[sel] - np.linalg.det(wantG[sel])).max() < 1e-9, snow


@pytest.mark.parametrize("angle", [0.0, 1e-3, 0.1, 1.4, 3.1])
@pytest.mark.parametrize("strain", [0.0, 1e-7, 1e-5, 1e-3, 3e-2, 0.3, 1.2])
def test_2d_fp32_closed_form_stress(km, strain, angle):
    """The cancellation-free fp32 form of the 2D stress (fixed_corotated_affine2_f32: the default of the fp32 build)
    against the fp64 oracle of utils.py:53-92 on the same fp32 inputs (C = 0, so nothing hides the stress): to 1e-5 of
    the batch's largest stress entry, plus the fp32 resolution of the ROTATION, 4 eps angle^2 as a strain (the form
    has no first-order sensitivity to the angle: a block rotated by 0.1 rad resolves strains of 5e-9) -- and the
    reference's spurious 1e-10 stress at F = I (quirks 3, 12), which plain fp32 arithmetic would lose, to 1e-5 of
    itself."""
    rng = np.random.default_rng(21)
    n = 6000
    th = rng.uniform(-angle, angle, size=n)
    R = np.stack([np.stack([np.cos(th), -np.sin(th)], -1), np.stack([np.sin(th), np.cos(th)], -1)], -2)
    F = f32(R @ (np.eye(2) + strain * rng.uniform(-1, 1, size=(n, 2, 2))))
    F[0] = np.eye(2)
    Z = np.zeros((n, 2, 2))
    mu, lam, mass, dt, vol, inv_dx = float(f32(4166.67)), float(f32(2777.78)), 1.0, 1e-4, 1.0, 1024.0
    k = float(f32(dt * vol * 4 * inv_dx * inv_dx))
    A = np.zeros((n, 4), np.float32); took = np.zeros(n, np.int32)
    km.km_affine2_f32(C.c_longlong(n), ptr(np.ascontiguousarray(F.reshape(n, 4), dtype=np.float32)), ptr(np.zeros((n, 4), np.float32)),
                      C.c_float(mu), C.c_float(lam), C.c_float(mass), C.c_float(k), ptr(A), ptr(took))
    want = O.fixed_corotated_stress_2d(F, inv_dx, np.full(n, mu), np.full(n, lam), dt, vol, np.full(n, mass), Z)
    want = want * (k / (dt * vol * 4 * inv_dx * inv_dx))
    t = took.astype(bool)
    assert t.mean() > 0.999                             # declines only next to a reflection (r < 1e-3)
    scale = np.abs(want[t]).max()
    floor = 2 * mu * k * 4 * np.finfo(np.float32).eps * angle ** 2
    assert np.abs(A.reshape(n, 2, 2) - want)[t].max() < 1e-5 * scale + floor, (strain, angle)
    # F = I: the reference's 1e-10 in the polar norm leaves a stress of 2 mu * 5e-11 on the diagonal
    assert want[0, 0, 0] != 0 and abs(A[0, 0] - want[0, 0, 0]) < 1e-5 * abs(want[0, 0, 0])
    assert abs(A[0, 1] - want[0, 0, 1]) < 1e-5 * abs(want[0, 0, 0])


def test_2d_fp32_closed_form_declines_next_to_a_reflection(km):
    """r = |(tr F, F10 - F01)| ~ 0 (F close to a reflection: the reference's own R degenerates to ~0 there) or NaN:
    the closed form declines and the kernel takes the fp64 form."""
    F = np.array([[[1.0, 0.0], [0.0, -1.0]], [[0.3, 0.7], [0.7, -0.3]], [[1.0, 0.0], [0.0, 1.0]], [[np.nan, 0.0], [0.0, 1.0]],
                  [[-1.0, 0.0], [0.0, -1.0]]], np.float32)
    n = len(F)
    A = np.zeros((n, 4), np.float32); took = np.zeros(n, np.int32)
    km.km_affine2_f32(C.c_longlong(n), ptr(np.ascontiguousarray(F.reshape(n, 4))), ptr(np.zeros((n, 4), np.float32)),
                      C.c_float(1.0), C.c_float(1.0), C.c_float(1.0), C.c_float(1.0), ptr(A), ptr(took))
    assert took.tolist() == [0, 0, 1, 0, 1]


@pytest.mark.parametrize("dtype,tol", [("f32", 2e-6), ("f64", 1e-14)])
def test_grid_update_walls_clamp_colliders_and_halo_sum(km, dtype, tol):
    """grid_op3_node, the body of every 3D grid-update kernel (three_d/grid_op.py:25-67): momentum -> velocity,
    gravity on y, the +-0.9 dx/dt clamp, per-axis sticky walls on GLOBAL faces only (quirk 5), plane colliders
    with the reference's scalar-added normal -- on the whole grid, and on a two-slab cut whose shared planes are
    summed while loading (ffmpm_grid_op_halo)."""
    rng = np.random.default_rng(17)
    res, G = 12, 13
    np_dt = np.float32 if dtype == "f32" else np.float64
    dx, dt, g = 1.0 / res, 2e-3, -9.8
    mass = np.where(rng.random((G, G, G, 1)) < 0.6, rng.uniform(0.1, 2.0, (G, G, G, 1)), 0.0)
    mom = rng.normal(0, 30.0, (G, G, G, 3)) * (mass > 0)            # large enough to hit the clamp (0.9 dx/dt = 37.5)
    mass, mom = mass.astype(np_dt).astype(np.float64), mom.astype(np_dt).astype(np.float64)
    pts = np.array([[0.5, 0.3, 0.5], [0.2, 0.5, 0.5]]); nrm = np.array([[0.0, 1.0, 0.0], [1.0, 0.2, 0.0]])
    want_v, want_m = mom.copy(), mass.copy()
    O.grid_op_3d(res, dx, dt, g, want_v, want_m)
    O.check_collision_points(pts, nrm, res, dx, want_v)
    stored_n = nrm + 1.0 / np.linalg.norm(nrm, axis=1, keepdims=True)        # ffmpm_set_colliders (grid_op.py:59-60)
    ia = lambda a: np.ascontiguousarray(a, dtype=np.int32)

    def run(i0, i1, grid, halo_lo=None, planes_lo=0, halo_hi=None, planes_hi=0):
        grid = np.array(grid, dtype=np_dt, order="C", copy=True)          # updated in place: never the caller's array
        getattr(km, f"km_grid_op3_{dtype}")(ptr(ia([res] * 3)), ptr(ia([i1 - i0, G, G])), C.c_int(i0), C.c_double(dx),
                                            C.c_double(dt), C.c_double(g), ptr(grid),
                                            ptr(halo_lo) if halo_lo is not None else None, C.c_int(planes_lo),
                                            ptr(halo_hi) if halo_hi is not None else None, C.c_int(planes_hi),
                                            C.c_int(len(pts)), ptr(np.ascontiguousarray(pts)), ptr(np.ascontiguousarray(stored_n)))
        return grid

    whole = run(0, G, np.concatenate([mom, mass], -1))
    scale = np.abs(want_v).max()
    assert np.abs(whole[..., :3] - want_v).max() / scale < tol and np.array_equal(whole[..., 3:], mass.astype(np_dt))
    assert (np.abs(want_v) == 0.9 * dx / dt).any()                     # the clamp was exercised
    # two slabs sharing planes 5..8: each holds a random split of the shared planes' {momentum, mass}
    full = np.concatenate([mom, mass], -1)
    share = rng.uniform(0.2, 0.8, (4, G, G, 1)) * (rng.random((4, G, G, 1)) < 0.7)
    lo_part = full[:9].copy(); lo_part[5:9] *= share
    hi_part = full[5:].copy(); hi_part[0:4] *= (1 - share)
    to = lambda a: np.ascontiguousarray(a, dtype=np_dt)
    lo = run(0, 9, lo_part, halo_hi=to(hi_part[0:4]), planes_hi=4)
    hi = run(5, G, hi_part, halo_lo=to(lo_part[5:9]), planes_lo=4)
    assert np.abs(lo[..., :3] - want_v[:9]).max() / scale < max(tol, 5e-7 if dtype == "f32" else 0)
    assert np.abs(hi[..., :3] - want_v[5:]).max() / scale < max(tol, 5e-7 if dtype == "f32" else 0)
    assert np.array_equal(lo[5:9], hi[0:4])                            # shared planes: a + b == b + a, bit for bit


@pytest.mark.parametrize("res", [16, 37, 64])
def test_tile_major_bin_keys(km, res):
    """bin_key_of: tile-major id of the base cell (4x4x4 cells per tile), n_cells for a stencil that leaves the
    grid or a NaN position; on a slab the key is local to the slab while base_x stays global."""
    rng = np.random.default_rng(res)
    n = 20000
    x = rng.uniform(-0.05, 1.05, size=(n, 3)).astype(np.float32)
    x[0] = [np.nan, 0.5, 0.5]; x[1] = [0.0, 0.0, 0.0]; x[2] = [(res - 1.5) / res] * 3
    with np.errstate(invalid="ignore"):                     # the NaN row: its base is never compared
        base, _ = O.base_and_fx(x.astype(np.float64), float(res))
    km.km_bin_keys_f32.restype = C.c_int
    for origin, n0 in ((0, res + 1), (5, 9)):
        nn = np.array([n0, res + 1, res + 1], np.int32)
        keys = np.zeros(n, np.int32); bx = np.zeros(n, np.int32)
        n_cells = km.km_bin_keys_f32(C.c_int(3), ptr(nn), C.c_int(origin), C.c_double(float(res)), C.c_int(int(res & (res - 1) == 0)),
                                     C.c_longlong(n), ptr(x), ptr(keys), ptr(bx))
        tiles = [(int(v) - 2 + 3) // 4 for v in nn]
        assert n_cells == tiles[0] * tiles[1] * tiles[2] * 64
        b = base - np.array([origin, 0, 0])
        ok = ~np.isnan(x).any(1) & (b >= 0).all(1) & (b + 2 < nn).all(1)
        t = ((b[:, 0] >> 2) * tiles[1] + (b[:, 1] >> 2)) * tiles[2] + (b[:, 2] >> 2)
        want = np.where(ok, t * 64 + ((b[:, 0] & 3) << 4) + ((b[:, 1] & 3) << 2) + (b[:, 2] & 3), n_cells)
        assert np.array_equal(keys, want.astype(np.int32))
        fin = ~np.isnan(x[:, 0])
        assert np.array_equal(bx[fin], base[fin, 0].astype(np.int32))
        assert not ok[0] and (origin > 0 or (ok[1] and ok[2]))


def test_economised_series_table_is_the_generated_one(km):
    """The coefficient tiers embedded in mpm_math.cuh (stress_coef / stress_tier_r) are what
    scripts/series_economized.py generates for the left form, each tier keeps p within 5e-8 of
    (1 + x - sqrt(1 + x)) / x on its interval with fp32 coefficients, and evaluated in fp32 on matrices at the
    tier's upper bound the series sits at rounding level (~2e-7)."""
    import sys
    root = os.path.dirname(HERE)
    sys.path.insert(0, os.path.join(root, "scripts"))
    import series_economized as E
    km.km_stress_coef.restype = C.c_float
    km.km_stress_tier_r.restype = C.c_float
    tab = E.table(left=True)
    assert len(tab) == 4
    for tier, (r, deg, coef) in enumerate(tab):
        assert deg == tier + 2
        assert np.float32(km.km_stress_tier_r(C.c_int(tier))) == np.float32(r)
        mine = np.float32([km.km_stress_coef(C.c_int(tier), C.c_int(i)) for i in range(deg + 1)])
        assert np.array_equal(mine, coef), (tier, mine, coef)
        assert km.km_stress_coef(C.c_int(tier), C.c_int(deg + 1)) == 0.0
        assert E.uniform_error(r, coef, E.p_exact) < 5.1e-8
        assert E.worst_matrix_error(r, coef, n=300, left=True) < 3e-7


@pytest.mark.parametrize("tier", [0, 1, 2, 3])
def test_series_at_a_higher_tier_than_needed(km, tier):
    """The warp-autonomous P2G evaluates all 64 particles of a window at the tier of the most strained one:
    h(G) = G p(G) with the tier-t polynomial must hold the bar for EVERY smaller strain as well (a higher-degree
    interpolant on a wider interval, used well inside it), against the eigendecomposition."""
    rng = np.random.default_rng(tier)
    n = 3000
    r_hi = [0.0136, 0.0436, 0.112, 0.15][tier]
    a = rng.uniform(-1, 1, (n, 3, 3))
    g = (a + np.swapaxes(a, 1, 2)) / 2
    g[::3] = np.einsum("ni,nj->nij", *(2 * [rng.normal(size=(n, 3))[::3]]))            # rank one: spectral radius = norm
    norm = np.linalg.norm(g, axis=(1, 2))
    target = r_hi * 10.0 ** rng.uniform(-4, 0, n) * 0.999
    g = (g / norm[:, None, None] * target[:, None, None]).astype(np.float32)
    G6 = np.ascontiguousarray(np.stack([g[:, 0, 0], g[:, 0, 1], g[:, 0, 2], g[:, 1, 1], g[:, 1, 2], g[:, 2, 2]], 1))
    H6 = np.zeros_like(G6)
    km.km_left_stress_h(C.c_int(tier), C.c_longlong(n), ptr(G6), ptr(H6))
    gs = np.zeros((n, 3, 3))
    for (i, j), col in zip(((0, 0), (0, 1), (0, 2), (1, 1), (1, 2), (2, 2)), range(6)):
        gs[:, i, j] = gs[:, j, i] = G6[:, col].astype(np.float64)
    w, q = np.linalg.eigh(gs)
    exact = np.einsum("nij,nj,nkj->nik", q, 1 + w - np.sqrt(1 + w), q)
    got = np.zeros((n, 3, 3))
    for (i, j), col in zip(((0, 0), (0, 1), (0, 2), (1, 1), (1, 2), (2, 2)), range(6)):
        got[:, i, j] = got[:, j, i] = H6[:, col]
    rel = np.abs(got - exact).max(axis=(1, 2)) / np.abs(exact).max(axis=(1, 2))
    assert rel.max() < 5e-7, (tier, rel.max())


def test_left_strain_and_tier_selection(km):
    """G = F F^T - I = E + E^T + E E^T formed from E = F - I (no cancellation), its Frobenius norm, the tier it
    selects (beyond 0.15 and NaN: the fp64 path) and J - 1 without cancellation, against fp64."""
    rng = np.random.default_rng(3)
    n = 6000
    amp = 10.0 ** rng.uniform(-6, -0.5, (n, 1, 1))
    F = f32(np.eye(3) + amp * rng.uniform(-1, 1, (n, 3, 3)))
    F[7] = np.nan
    Ff = np.ascontiguousarray(F.reshape(n, 9), dtype=np.float32)
    G6 = np.zeros((n, 6), np.float32); r2 = np.zeros(n, np.float32); tier = np.zeros(n, np.int32); jm1 = np.zeros(n, np.float32)
    km.km_left_strain(C.c_longlong(n), ptr(Ff), ptr(G6), ptr(r2), ptr(tier), ptr(jm1))
    Gd = F @ np.swapaxes(F, 1, 2) - np.eye(3)
    want6 = np.stack([Gd[:, 0, 0], Gd[:, 0, 1], Gd[:, 0, 2], Gd[:, 1, 1], Gd[:, 1, 2], Gd[:, 2, 2]], 1)
    ok = np.arange(n) != 7
    scale = np.abs(want6[ok]).max(axis=1)
    assert (np.abs(G6[ok] - want6[ok]).max(axis=1) / scale).max() < 5e-7
    norm = np.linalg.norm(Gd[ok], axis=(1, 2))
    bounds = np.array([0.0136, 0.0436, 0.112, 0.15])
    want_tier = np.searchsorted(bounds, norm, side="right")
    near = np.min(np.abs(norm[:, None] - bounds[None, :]) / bounds[None, :], axis=1) < 1e-5
    assert np.array_equal(tier[ok][~near], want_tier[~near])
    assert tier[7] == 4                                            # NaN: beyond every tier
    J = np.linalg.det(F[ok])
    strain = np.abs(F[ok] - np.eye(3)).max(axis=(1, 2))            # J - 1 = tr E + O(E^2): the scale of its terms
    assert (np.abs(jm1[ok] - (J - 1)) / strain).max() < 1e-6


@pytest.mark.parametrize("angle,strain", [(0.02, 1e-3), (0.1, 1e-3), (0.1, 3e-2), (0.3, 1e-3), (0.3, 3e-2)])
def test_stress_under_rotation(km, angle, strain):
    """F = R(angle) (I + strain): E = F - I is then O(angle), and G = E + E^T + E E^T cancels in fp32; the series
    declines once ||G||_F >= 0.15 and the fp64 path takes over -- either way inside the 1e-5 bar."""
    rng = np.random.default_rng(int(angle * 1000 + strain * 1e6))
    n, res = 2000, 32
    dx = 1.0 / res
    vol = float(f32((dx / 2) ** 3))

    def rot(a, ax):
        c, s = np.cos(a), np.sin(a)
        R = np.eye(3)
        i, j = [(1, 2), (0, 2), (0, 1)][ax]
        R[i, i] = c; R[j, j] = c; R[i, j] = -s; R[j, i] = s
        return R
    Rm = np.stack([rot(rng.uniform(-angle, angle), int(rng.integers(3))) for _ in range(n)])
    F = f32(Rm @ (np.eye(3) + strain * rng.uniform(-1, 1, (n, 3, 3))))
    x = f32(rng.uniform(0.25, 0.75, (n, 3)))
    Z = np.zeros((n, 3, 3)); mass = np.full(n, vol); mu = np.full(n, f32(4166.67)); lam = np.full(n, f32(2777.78))
    want = O.fixed_corotated_stress_3d(F, float(res), mu, lam, 1e-4, vol, mass, Z)
    scale = np.abs(want).max()
    got = prepare3(km, "f32", res, x, np.zeros((n, 3)), Z, F, mass, mu, lam, 1e-4, vol)[2]
    assert np.abs(got - want).max() / scale < 1e-5
