"""ctypes wrapper of oracle/mpm_oracle.c (TEST INFRASTRUCTURE: checker + CPU baseline).

Only tests/, __graft_entry__ and bench.py's cpu_baseline / --impl reference legs
use this module."""
from __future__ import annotations

import ctypes as C
import os
import subprocess
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "libmpm_oracle.so")
_lib = None

_D = np.ctypeslib.ndpointer(dtype=np.float64, flags="C_CONTIGUOUS")


def build(force: bool = False) -> str:
    src = os.path.join(HERE, "mpm_oracle.c")
    if force or not os.path.exists(LIB) or os.path.getmtime(LIB) < os.path.getmtime(src):
        subprocess.run(["make", "-C", HERE, "-B", "libmpm_oracle.so"], check=True, capture_output=True)
    return LIB


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(LIB)
        i64, f64 = C.c_int64, C.c_double
        L.oracle_max_threads.restype = C.c_int
        L.oracle_set_threads.argtypes = [C.c_int]
        L.oracle_p2g_3d.restype = i64
        L.oracle_p2g_3d.argtypes = [i64, i64, f64, f64, f64, f64, f64, _D, _D, _D, _D, _D, _D, _D, _D, _D, C.c_void_p, C.c_int]
        L.oracle_grid_op_3d.argtypes = [i64, f64, f64, f64, _D, _D]
        L.oracle_g2p_3d.restype = i64
        L.oracle_g2p_3d.argtypes = [i64, i64, f64, f64, _D, _D, _D, _D, _D]
        L.oracle_substep_3d.restype = i64
        L.oracle_substep_3d.argtypes = [i64, i64, f64, f64, f64, f64, f64, f64, _D, _D, _D, _D, _D, _D, _D, _D, _D]
        L.oracle_p2g_2d.argtypes = [i64, i64, f64, f64, f64, f64, f64, f64, f64, f64, _D, _D, _D, _D, _D, _D, _D, C.c_int]
        L.oracle_grid_op_2d.argtypes = [i64, f64, f64, _D, _D]
        L.oracle_g2p_2d.argtypes = [i64, i64, f64, f64, _D, _D, _D, _D, _D, _D]
        L.oracle_substep_2d.argtypes = [i64, i64, f64, f64, f64, f64, f64, f64, f64, f64, f64, _D, _D, _D, _D, _D, _D, _D]
        _lib = L
    return _lib


def _c(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def host_threads() -> int:
    """Threads the CPU baseline runs on: every core this process may be scheduled on.  NOT
    omp_get_max_threads(): torch.distributed.run exports OMP_NUM_THREADS=1 to its workers, which would
    silently turn the baseline of an N > 1 launch into a single-threaded one."""
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except AttributeError:
        return max(1, os.cpu_count() or 1)


# ---- reference-shaped helpers (same argument order as oracle/mpm_oracle.py) ----
def p2g_3d(inv_dx, hardening, dx, dt, volume, grid_velocity, grid_mass, x, mass, mu0, lam0, v, F, C_, Jp, model="neo_hookean"):
    n = len(x)
    G = grid_velocity.shape[0]
    jp = _c(Jp).reshape(-1)
    bad = lib().oracle_p2g_3d(n, G, inv_dx, hardening, dx, dt, volume, grid_velocity.reshape(-1), grid_mass.reshape(-1),
                              _c(x), _c(np.broadcast_to(mass, (n,))), _c(np.broadcast_to(mu0, (n,))),
                              _c(np.broadcast_to(lam0, (n,))), _c(v), _c(F), _c(C_), jp.ctypes.data, int(model == "snow"))
    if bad:
        raise RuntimeError("particle stencil leaves the grid")


def grid_op_3d(res, dx, dt, gravity, grid_velocity, grid_mass):
    lib().oracle_grid_op_3d(res, dx, dt, gravity, grid_velocity.reshape(-1), grid_mass.reshape(-1))


def g2p_3d(inv_dx, dt, grid_velocity, x, v, F, C_, Jp=None, model="neo_hookean"):
    assert model == "neo_hookean"
    bad = lib().oracle_g2p_3d(len(x), grid_velocity.shape[0], inv_dx, dt, grid_velocity.reshape(-1), x, v, F, C_)
    if bad:
        raise RuntimeError("particle stencil leaves the grid")


def solve_mls_mpm_3d(res, inv_dx, hardening, dx, dt, volume, gravity, x, mass, mu0, lam0, v, F, C_, scratch=None):
    G = res + 1
    gv = np.empty(G * G * G * 3) if scratch is None else scratch[0]
    gm = np.empty(G * G * G) if scratch is None else scratch[1]
    bad = lib().oracle_substep_3d(len(x), res, inv_dx, hardening, dx, dt, volume, gravity, x, mass, mu0, lam0, v, F, C_, gv, gm)
    if bad:
        raise RuntimeError("particle stencil leaves the grid")
    return gv.reshape(G, G, G, 3), gm.reshape(G, G, G, 1)


def solve_mls_mpm_2d(res, inv_dx, hardening, mu_0, lambda_0, mass, dx, dt, volume, gravity, x, v, F, C_, Jp, scratch=None):
    G = res + 1
    gv = np.empty(G * G * 2) if scratch is None else scratch[0]
    gm = np.empty(G * G) if scratch is None else scratch[1]
    lib().oracle_substep_2d(len(x), res, inv_dx, hardening, mu_0, lambda_0, mass, dx, dt, volume, gravity, x, v, F, C_,
                            Jp.reshape(-1), gv, gm)
    return gv.reshape(G, G, 2), gm.reshape(G, G, 1)


# ---- CPU baseline ----
def time_sample(scene, budget_s: float = 15.0, threads=None, substeps=None):
    """Time the C port on a bounded sample of `scene`: the first m particles (a
    sub-block of the same density on the same grid resolution), m chosen so that a
    few substeps take about `budget_s` seconds."""
    L = lib()
    cores = threads or host_threads()
    L.oracle_set_threads(cores)
    d, res = scene.dim, scene.res
    G = res + 1
    gv = np.empty(G ** d * d)
    gm = np.empty(G ** d)

    def run(m, reps):
        x = _c(scene.x[:m]); v = _c(scene.v[:m]); F = _c(scene.F[:m]); Cc = _c(scene.C[:m])
        mass = np.full(m, scene.mass); mu = np.full(m, scene.mu_0); lam = np.full(m, scene.lambda_0)
        Jp = np.ones(m)
        t0 = time.perf_counter()
        for _ in range(reps):
            if d == 3:
                L.oracle_substep_3d(m, res, float(res), scene.hardening, 1.0 / res, scene.dt, scene.volume, scene.gravity,
                                    x, mass, mu, lam, v, F, Cc, gv, gm)
            else:
                L.oracle_substep_2d(m, res, float(res), scene.hardening, scene.mu_0, scene.lambda_0, scene.mass,
                                    1.0 / res, scene.dt, scene.volume, scene.gravity, x, v, F, Cc, Jp, gv, gm)
        return time.perf_counter() - t0

    m = min(scene.n, 200_000)
    run(min(m, 20_000), 1)                       # warm-up (page faults, thread pool)
    t = run(m, 1)
    reps = substeps or 3
    # scale the sample so reps substeps fill the budget, capped at the full workload
    m2 = int(min(scene.n, max(m, m * budget_s / max(t, 1e-6) / reps * 0.8)))
    t2 = run(m2, reps)
    value = m2 * reps / t2
    return {"value": value, "unit": "particle-substeps/s", "cores": int(cores), "kind": "port",
            "sample": f"first {m2} particles of the workload (same density, full {res}^{d} grid), {reps} substeps, "
                      f"{t2:.1f} s; C port of the reference loops with OpenMP over particles "
                      f"(the reference's numba path is serial: 1 core)",
            "sample_particles": m2, "seconds": t2}


class SampleRunner:
    """Repeated substeps of the C port on a fixed sample of a scene (bench.py --impl reference)."""

    def __init__(self, scene, m: int, threads=None):
        self.L = lib()
        self.cores = threads or host_threads()
        self.L.oracle_set_threads(self.cores)
        self.scene, self.m = scene, int(min(m, scene.n))
        d, res = scene.dim, scene.res
        G = res + 1
        self.gv = np.empty(G ** d * d)
        self.gm = np.empty(G ** d)
        m = self.m
        self.x = _c(scene.x[:m]); self.v = _c(scene.v[:m]); self.F = _c(scene.F[:m]); self.C = _c(scene.C[:m])
        self.mass = np.full(m, scene.mass); self.mu = np.full(m, scene.mu_0); self.lam = np.full(m, scene.lambda_0)
        self.Jp = np.ones(m)

    def step(self) -> float:
        sc, res = self.scene, self.scene.res
        t0 = time.perf_counter()
        if sc.dim == 3:
            self.L.oracle_substep_3d(self.m, res, float(res), sc.hardening, 1.0 / res, sc.dt, sc.volume, sc.gravity,
                                     self.x, self.mass, self.mu, self.lam, self.v, self.F, self.C, self.gv, self.gm)
        else:
            self.L.oracle_substep_2d(self.m, res, float(res), sc.hardening, sc.mu_0, sc.lambda_0, sc.mass, 1.0 / res,
                                     sc.dt, sc.volume, sc.gravity, self.x, self.v, self.F, self.C, self.Jp, self.gv, self.gm)
        return time.perf_counter() - t0
