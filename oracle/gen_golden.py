"""Generate ``tests/golden/*.npz`` by running the REAL reference (numba) here.

Run in the build container only (``/root/reference`` does not exist on the GPU
box):  ``python oracle/gen_golden.py``.  Every array written is an output of the
unmodified reference functions

    femflow.solvers.mpm.three_d.{p2g,grid_op,g2p}      (three_d/*.py)
    femflow.solvers.mpm.two_d.{p2g,grid_op,g2p}        (two_d/*.py)
    femflow.solvers.mpm.mls_mpm.solve_mls_mpm_3d       (mls_mpm.py:40-79)
    femflow.simulation.mpm.primitives.generate_*       (primitives.py:46-76)

on seeded inputs that are stored alongside.  Inputs are rounded to float32 and
then widened to float64 so the fp32 CUDA path and the fp64 reference start from
bit-identical state.
"""
from __future__ import annotations

import os
import sys
import types
import warnings

import numpy as np

REF = os.environ.get("FEMFLOW_REFERENCE", "/root/reference")
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden")


def _import_reference():
    if not os.path.isdir(REF):
        raise SystemExit(f"reference tree not found at {REF}")
    sys.path.insert(0, REF)
    # GUI / meshing dependencies that the MPM path never touches.
    for name in ("igl", "wildmeshing", "skimage", "skimage.measure", "imgui", "glfw",
                 "OpenGL", "OpenGL.GL", "OpenGL.GLU", "ilupp", "matplotlib", "matplotlib.pyplot",
                 "cv2", "jax", "jax.numpy", "loguru"):
        if name not in sys.modules:
            try:
                __import__(name)
            except Exception:
                m = types.ModuleType(name)
                m.__path__ = []  # type: ignore[attr-defined]
                if name == "loguru":
                    class _L:
                        def __getattr__(self, _):
                            return lambda *a, **k: None
                    m.logger = _L()  # type: ignore[attr-defined]
                sys.modules[name] = m
    warnings.filterwarnings("ignore")
    from femflow.solvers.mpm import three_d, two_d
    from femflow.solvers.mpm.mls_mpm import solve_mls_mpm_3d
    from femflow.solvers.mpm.particle import Particle
    from numba.typed import List as NbList
    return three_d, two_d, solve_mls_mpm_3d, Particle, NbList


def f32(a):
    return np.asarray(a, dtype=np.float32).astype(np.float64)


def make_particles(Particle, NbList, x, mass, lam0, mu0):
    lst = NbList()
    for i in range(len(x)):
        # particle.py:20-27 -- argument order is (pos, force, mass, lambda_, mu)
        lst.append(Particle(x[i].copy(), 0.0, float(mass[i]), float(lam0[i]), float(mu0[i])))
    return lst


def positions(particles):
    return np.array([p.pos.copy() for p in particles])


def random_block_3d(rng, n, res, lo=0.3, hi=0.7, strain=0.02, vel=0.1, cmag=0.5):
    x = f32(rng.uniform(lo, hi, size=(n, 3)))
    v = f32(rng.normal(0, vel, size=(n, 3)))
    F = f32(np.eye(3) + rng.normal(0, strain, size=(n, 3, 3)))
    C = f32(rng.normal(0, cmag, size=(n, 3, 3)))
    Jp = np.ones((n, 1))
    return x, v, F, C, Jp


def run_phases_3d(ref, p, x, mass, lam0, mu0, v, F, C, Jp):
    three_d, _, _, Particle, NbList = ref
    G = p["res"] + 1
    particles = make_particles(Particle, NbList, x, mass, lam0, mu0)
    v, F, C, Jp = v.copy(), F.copy(), C.copy(), Jp.copy()
    gv = np.zeros((G, G, G, 3))
    gm = np.zeros((G, G, G, 1))
    three_d.p2g(p["inv_dx"], p["hardening"], p["dx"], p["dt"], p["volume"], gv, gm,
                particles, v, F, C, Jp, "neo_hookean")
    mom = gv.copy()
    three_d.grid_op(p["res"], p["dx"], p["dt"], p["gravity"], gv, gm)
    vel = gv.copy()
    three_d.g2p(p["inv_dx"], p["dt"], gv, particles, v, F, C, Jp, "neo_hookean")
    return dict(grid_momentum=mom, grid_mass=gm, grid_velocity=vel,
                x_out=positions(particles), v_out=v, F_out=F, C_out=C)


def sparse_grid(a, name):
    """Store a mostly-zero grid as (flat index, values)."""
    d = a.shape[-1]
    flat = a.reshape(-1, d)
    nz = np.flatnonzero(np.any(flat != 0, axis=1))
    return {f"{name}_idx": nz.astype(np.int64), f"{name}_val": flat[nz], f"{name}_shape": np.array(a.shape)}


def params_3d(res, dt, volume, hardening, gravity):
    return dict(res=res, dx=1.0 / res, inv_dx=float(res), dt=dt, volume=volume,
                hardening=hardening, gravity=gravity)


def save(name, **arrs):
    os.makedirs(OUT, exist_ok=True)
    path = os.path.join(OUT, name + ".npz")
    np.savez_compressed(path, **arrs)
    print(f"wrote {path}  ({os.path.getsize(path) / 1024:.1f} KiB)")


# --------------------------------------------------------------------------- #
def gen_kat3d(ref):
    """two particles, res 8."""
    p = params_3d(8, 1e-3, 0.5, 0.7, -9.8)
    x = np.array([[0.40, 0.50, 0.60], [0.45, 0.55, 0.52]])
    mass = np.array([1.0, 2.0]); lam0 = np.array([20.0, 40.0]); mu0 = np.array([30.0, 10.0])
    v = np.array([[0.1, -0.2, 0.3], [0.0, 0.5, -0.1]])
    F = np.array([[[1.05, 0.02, 0], [0.01, 0.97, 0.03], [0, -0.02, 1.01]],
                  [[0.9, 0.1, 0], [-0.1, 1.1, 0], [0.05, 0, 1]]])
    C = np.array([[[0.1, 0.2, 0.3], [0, -0.1, 0.2], [0.3, 0.1, 0]], np.zeros((3, 3))])
    Jp = np.ones((2, 1))
    out = run_phases_3d(ref, p, x, mass, lam0, mu0, v, F, C, Jp)
    save("kat3d", x=x, mass=mass, lam0=lam0, mu0=mu0, v=v, F=F, C=C, Jp=Jp,
         **{k: np.float64(val) for k, val in p.items()}, **out)


def gen_block3d(ref):
    """Random perturbed elastic block, per-phase outputs (strain 2e-2)."""
    rng = np.random.default_rng(0)
    n, res = 1500, 32
    p = params_3d(res, 1e-4, 1.0 / res ** 3 / 8, 0.7, -9.8)
    x, v, F, C, Jp = random_block_3d(rng, n, res)
    mass = f32(rng.choice([1.0, 10.0], size=n) * 1e-5)
    lam0 = f32(rng.choice([38.9, 1428.6], size=n))
    mu0 = f32(rng.choice([58.3, 357.1], size=n))
    out = run_phases_3d(ref, p, x, mass, lam0, mu0, v, F, C, Jp)
    grids = {}
    for k in ("grid_momentum", "grid_mass", "grid_velocity"):
        grids.update(sparse_grid(out.pop(k), k))
    save("block3d", x=x, mass=mass, lam0=lam0, mu0=mu0, v=v, F=F, C=C, Jp=Jp,
         **{k: np.float64(val) for k, val in p.items()}, **out, **grids)


def gen_rest3d(ref):
    """Near-rest block (strain 1e-4, v = 0): the stress-dominated regime where
    an all-fp32 constitutive evaluation fails."""
    rng = np.random.default_rng(1)
    n, res = 1000, 32
    p = params_3d(res, 1e-4, 1.0, 1.0, -9.8)
    x, v, F, C, Jp = random_block_3d(rng, n, res, strain=1e-4, vel=0.0, cmag=0.0)
    mass = np.ones(n); lam0 = np.full(n, f32(2777.7778)); mu0 = np.full(n, f32(4166.6665))
    out = run_phases_3d(ref, p, x, mass, lam0, mu0, v, F, C, Jp)
    grids = {}
    for k in ("grid_momentum", "grid_mass", "grid_velocity"):
        grids.update(sparse_grid(out.pop(k), k))
    save("rest3d", x=x, mass=mass, lam0=lam0, mu0=mu0, v=v, F=F, C=C, Jp=Jp,
         **{k: np.float64(val) for k, val in p.items()}, **out, **grids)


def gen_walls3d(ref):
    """Particles hugging all six walls (base 0 via truncation toward zero, base
    R-2 at the high side) moving outward, so clamp + sticky walls are exercised."""
    rng = np.random.default_rng(2)
    res = 16
    p = params_3d(res, 2e-3, 1e-4, 1.0, -9.8)
    dx = 1.0 / res
    pts = []
    for axis in range(3):
        for side in (0, 1):
            q = rng.uniform(0.2, 0.8, size=(40, 3))
            q[:, axis] = rng.uniform(0.0, 1.4 * dx, 40) if side == 0 else \
                rng.uniform(1 - 1.6 * dx, (res - 0.5) * dx - 1e-6, 40)
            pts.append(q)
    x = f32(np.concatenate(pts))
    n = len(x)
    v = f32(rng.normal(0, 40.0, size=(n, 3)))     # beyond v_allowed = 0.9*dx/dt = 28.1
    F = f32(np.eye(3) + rng.normal(0, 0.01, size=(n, 3, 3)))
    C = f32(rng.normal(0, 1.0, size=(n, 3, 3)))
    Jp = np.ones((n, 1))
    mass = np.full(n, f32(1e-3)); lam0 = np.full(n, 40.0); mu0 = np.full(n, 60.0)
    out = run_phases_3d(ref, p, x, mass, lam0, mu0, v, F, C, Jp)
    grids = {}
    for k in ("grid_momentum", "grid_mass", "grid_velocity"):
        grids.update(sparse_grid(out.pop(k), k))
    save("walls3d", x=x, mass=mass, lam0=lam0, mu0=mu0, v=v, F=F, C=C, Jp=Jp,
         **{k: np.float64(val) for k, val in p.items()}, **out, **grids)


def c1_scene_points():
    """paper_1.py:75-114 ``multi_drop_experiment(0)`` particle positions, built
    with the reference's own generators; Mesh stores float32 (viz/mesh.py:26)."""
    from femflow.simulation.mpm.primitives import generate_cube_points, generate_implicit_points
    g = generate_implicit_points("gyroid", 0.2, 0.3, 30).astype(np.float32)
    g[:, 1] += np.float32(0.1)
    lo, hi = g.min(0), g.max(0)
    c = generate_cube_points((lo[0], hi[0]), (lo[1], hi[1]), (lo[2], hi[2]), 30).astype(np.float32)
    c[:, 1] += np.float32(3)
    coeff = 0.05
    xg = (g * coeff).astype(np.float64)      # simulation.py:81-83 (float32 product, then widen)
    xc = (c * coeff).astype(np.float64)
    return g, c, xg, xc


def gen_c1(ref):
    """BASELINE config 1: the paper scene (every 8th particle to keep the
    fixture small; the full scene is rebuilt in the tests from the stored mesh
    vertices), 10 substeps through the reference driver."""
    _, _, solve, Particle, NbList = ref
    from femflow.numerics.fem import Ev_to_lambda, Ev_to_mu
    g, c, xg, xc = c1_scene_points()
    g_s, c_s, xg_s, xc_s = g[::8], c[::8], xg[::8], xc[::8]
    x = np.concatenate([xg_s, xc_s])
    n = len(x)
    mass = np.concatenate([np.full(len(xg_s), 1.0), np.full(len(xc_s), 10.0)])
    lam0 = np.concatenate([np.full(len(xg_s), Ev_to_lambda(140, 0.2)), np.full(len(xc_s), Ev_to_lambda(1000, 0.4))])
    mu0 = np.concatenate([np.full(len(xg_s), Ev_to_mu(140, 0.2)), np.full(len(xc_s), Ev_to_mu(1000, 0.4))])
    particles = make_particles(Particle, NbList, x, mass, lam0, mu0)
    v = np.zeros((n, 3)); F = np.tile(np.eye(3), (n, 1, 1)); C = np.zeros((n, 3, 3)); Jp = np.ones((n, 1))
    res, dt, volume, hardening, gravity = 64, 1e-4, 1.0, 0.7, -9.8
    snaps = {}
    for step in range(1, 11):
        solve(res, float(res), hardening, 1.0 / res, dt, volume, gravity, particles, v, F, C, Jp)
        if step in (1, 10):
            snaps[f"x_{step}"] = positions(particles)
            snaps[f"v_{step}"] = v.copy(); snaps[f"F_{step}"] = F.copy(); snaps[f"C_{step}"] = C.copy()
    save("c1_scene", gyroid_vertices=g_s, collider_vertices=c_s,
         n_gyroid_full=np.int64(len(g)), n_collider_full=np.int64(len(c)),
         gyroid_bbox=np.stack([g.min(0), g.max(0)]), collider_bbox=np.stack([c.min(0), c.max(0)]),
         res=np.float64(res), dt=np.float64(dt), volume=np.float64(volume),
         hardening=np.float64(hardening), gravity=np.float64(gravity),
         tightening_coeff=np.float64(0.05), mass=mass, lam0=lam0, mu0=mu0, **snaps)


def gen_drift3d(ref):
    """1000 substeps of a small falling/compressing block through the reference
    driver; snapshots for the drift curve (reported, not gated)."""
    _, _, solve, Particle, NbList = ref
    rng = np.random.default_rng(3)
    res = 32
    side = 8
    cell = (np.stack(np.meshgrid(*([np.arange(side)] * 3), indexing="ij"), -1).reshape(-1, 3) + 0.5) / 2
    x = f32((cell + rng.uniform(-0.2, 0.2, cell.shape)) / res + np.array([0.4, 0.06, 0.4]))
    n = len(x)
    dx = 1.0 / res
    volume = (dx / 2) ** 3
    mass = np.full(n, f32(volume * 1.0)); E, nu = 1e3, 0.2
    lam0 = np.full(n, f32(E * nu / ((1 + nu) * (1 - 2 * nu)))); mu0 = np.full(n, f32(E / (2 * (1 + nu))))
    particles = make_particles(Particle, NbList, x, mass, lam0, mu0)
    v = np.zeros((n, 3)); F = np.tile(np.eye(3), (n, 1, 1)); C = np.zeros((n, 3, 3)); Jp = np.ones((n, 1))
    dt, hardening, gravity = 2e-4, 1.0, -9.8
    snaps = {}
    for step in range(1, 1001):
        solve(res, float(res), hardening, dx, dt, volume, gravity, particles, v, F, C, Jp)
        if step in (1, 10, 100, 300, 1000):
            snaps[f"x_{step}"] = positions(particles)
            snaps[f"v_{step}"] = v.copy(); snaps[f"F_{step}"] = F.copy(); snaps[f"C_{step}"] = C.copy()
    save("drift3d", x=x, mass=mass, lam0=lam0, mu0=mu0, res=np.float64(res), dt=np.float64(dt),
         volume=np.float64(volume), hardening=np.float64(hardening), gravity=np.float64(gravity), **snaps)


# --------------------------------------------------------------------------- #
TEST2D = dict(res=80, dt=1e-4, gravity=-200.0, mass=1.0, volume=1.0, hardening=10.0,
              mu_0=1e4 / (2 * 1.2), lambda_0=1e4 * 0.2 / (1.2 * 0.6))


def run_phases_2d(ref, p, x, v, F, C, Jp, model="neo_hookean"):
    _, two_d, *_ = ref
    res = p["res"]; G = res + 1; inv_dx = float(res); dx = 1.0 / res
    x, v, F, C, Jp = x.copy(), v.copy(), F.copy(), C.copy(), Jp.copy()
    gv = np.zeros((G, G, 2)); gm = np.zeros((G, G, 1))
    two_d.p2g(inv_dx, p["hardening"], p["mu_0"], p["lambda_0"], p["mass"], dx, p["dt"], p["volume"],
              gv, gm, x, v, F, C, Jp, model)
    mom = gv.copy()
    two_d.grid_op(res, p["dt"], p["gravity"], gv, gm)
    vel = gv.copy()
    two_d.g2p(inv_dx, p["dt"], gv, x, v, F, C, Jp, model)
    return dict(grid_momentum=mom, grid_mass=gm, grid_velocity=vel, x_out=x, v_out=v, F_out=F,
                C_out=C, Jp_out=Jp)


def gen_test2d(ref):
    """The reference's own 3-particle 2D test scene
    (solvers/mpm/tests/test_particle_to_grid.py:11-27,53-80): per-phase
    outputs of one substep plus snapshots of substeps 2-4.  With these constants
    (mass = volume = 1) the reference itself is unstable -- the 1e-10 of
    polar_decomp_2d seeds a stress error that grows ~30x per substep and the
    reference raises LinAlgError at substep 11 -- so longer runs use
    ``drift2d`` instead."""
    _, two_d, *_ = ref
    p = TEST2D
    x = np.array([[0.55, 0.45], [0.45, 0.65], [0.55, 0.85]])
    v = np.zeros((3, 2)); F = np.tile(np.eye(2), (3, 1, 1)); C = np.zeros((3, 2, 2)); Jp = np.ones((3, 1))
    out = run_phases_2d(ref, p, x, v, F, C, Jp)
    res = p["res"]; G = res + 1
    xs, vs, Fs, Cs, Jps = x.copy(), v.copy(), F.copy(), C.copy(), Jp.copy()
    snaps = {}
    for step in range(1, 5):
        gv = np.zeros((G, G, 2)); gm = np.zeros((G, G, 1))
        two_d.p2g(float(res), p["hardening"], p["mu_0"], p["lambda_0"], p["mass"], 1.0 / res, p["dt"],
                  p["volume"], gv, gm, xs, vs, Fs, Cs, Jps, "neo_hookean")
        two_d.grid_op(res, p["dt"], p["gravity"], gv, gm)
        two_d.g2p(float(res), p["dt"], gv, xs, vs, Fs, Cs, Jps, "neo_hookean")
        if step in (2, 3, 4):
            snaps[f"x_{step}"] = xs.copy(); snaps[f"v_{step}"] = vs.copy()
            snaps[f"F_{step}"] = Fs.copy(); snaps[f"C_{step}"] = Cs.copy(); snaps[f"Jp_{step}"] = Jps.copy()
    save("test2d", x=x, v=v, F=F, C=C, Jp=Jp, **{k: np.float64(val) for k, val in p.items()}, **out, **snaps)


def gen_drift2d(ref):
    """1000 substeps of a small, physically scaled 2D block dropping onto the
    floor band (stable: volume = (dx/2)^2, mass = rho*volume)."""
    _, two_d, *_ = ref
    rng = np.random.default_rng(7)
    res = 64; G = res + 1; dx = 1.0 / res
    side = 12
    cell = (np.stack(np.meshgrid(*([np.arange(side)] * 2), indexing="ij"), -1).reshape(-1, 2) + 0.5) / 2
    x = f32((cell + rng.uniform(-0.2, 0.2, cell.shape)) / res + np.array([0.45, 0.08]))
    n = len(x)
    vol = float(f32((dx / 2) ** 2)); E, nu = 1e3, 0.2
    p = dict(res=res, dt=1e-4, gravity=-9.8, mass=vol * 1.0, volume=vol, hardening=1.0,
             mu_0=float(f32(E / (2 * (1 + nu)))), lambda_0=float(f32(E * nu / ((1 + nu) * (1 - 2 * nu)))))
    v = np.zeros((n, 2)); F = np.tile(np.eye(2), (n, 1, 1)); C = np.zeros((n, 2, 2)); Jp = np.ones((n, 1))
    xs, vs, Fs, Cs, Jps = x.copy(), v, F, C, Jp
    snaps = {}
    for step in range(1, 1001):
        gv = np.zeros((G, G, 2)); gm = np.zeros((G, G, 1))
        two_d.p2g(float(res), p["hardening"], p["mu_0"], p["lambda_0"], p["mass"], dx, p["dt"],
                  p["volume"], gv, gm, xs, vs, Fs, Cs, Jps, "neo_hookean")
        two_d.grid_op(res, p["dt"], p["gravity"], gv, gm)
        two_d.g2p(float(res), p["dt"], gv, xs, vs, Fs, Cs, Jps, "neo_hookean")
        if step in (1, 10, 100, 300, 1000):
            snaps[f"x_{step}"] = xs.copy(); snaps[f"v_{step}"] = vs.copy()
            snaps[f"F_{step}"] = Fs.copy(); snaps[f"C_{step}"] = Cs.copy(); snaps[f"Jp_{step}"] = Jps.copy()
    print("  drift2d: final max|F-I| =", np.abs(Fs - np.eye(2)).max(), " min y =", xs[:, 1].min())
    save("drift2d", x=x, **{k: np.float64(val) for k, val in p.items()}, **snaps)


def gen_block2d(ref):
    """Random perturbed 2D block with particles inside the wall and floor bands."""
    rng = np.random.default_rng(4)
    n, res = 2000, 64
    p = dict(TEST2D, res=res, dt=1e-4, gravity=-9.8, mass=f32(1.0 / 64 ** 2 / 4),
             volume=f32(1.0 / 64 ** 2 / 4), hardening=10.0)
    p = {k: float(val) if k != "res" else val for k, val in p.items()}
    x = f32(rng.uniform(0.02, 0.96, size=(n, 2)))
    v = f32(rng.normal(0, 0.5, size=(n, 2)))
    F = f32(np.eye(2) + rng.normal(0, 0.02, size=(n, 2, 2)))
    C = f32(rng.normal(0, 0.5, size=(n, 2, 2)))
    Jp = np.ones((n, 1))
    out = run_phases_2d(ref, p, x, v, F, C, Jp)
    save("block2d", x=x, v=v, F=F, C=C, Jp=Jp, **{k: np.float64(val) for k, val in p.items()}, **out)


def gen_quirk2d(ref):
    """2x2 SVD round trip ``U @ diag(sig) @ Vh.T`` (two_d/g2p.py:37-43) on
    deformation gradients of both determinant signs, grid at rest: documents
    what the reference does when det F < 0."""
    _, two_d, *_ = ref
    rng = np.random.default_rng(5)
    n, res = 64, 16
    x = f32(rng.uniform(0.3, 0.7, size=(n, 2)))
    F = f32(rng.normal(0, 1.0, size=(n, 2, 2)))
    v = np.zeros((n, 2)); C = np.zeros((n, 2, 2)); Jp = np.ones((n, 1))
    G = res + 1
    gv = np.zeros((G, G, 2))
    Fo, Jpo, xo = F.copy(), Jp.copy(), x.copy()
    two_d.g2p(float(res), 1e-4, gv, xo, v, Fo, C, Jpo, "neo_hookean")
    save("quirk2d", x=x, F=F, F_out=Fo, Jp_out=Jpo, res=np.float64(res))


def gen_snow3d(ref):
    """Snow branch of the 3D kernels (unreachable from the 3D driver,
    mls_mpm.py:58, but part of the kernels' signature)."""
    three_d, _, _, Particle, NbList = ref
    rng = np.random.default_rng(6)
    n, res = 300, 16
    p = params_3d(res, 1e-4, 1e-4, 10.0, -9.8)
    x, v, F, C, _ = random_block_3d(rng, n, res, strain=0.03)
    Jp = f32(rng.uniform(0.9, 1.1, size=(n, 1)))
    mass = np.full(n, f32(1e-4)); lam0 = np.full(n, 40.0); mu0 = np.full(n, 60.0)
    G = res + 1
    particles = make_particles(Particle, NbList, x, mass, lam0, mu0)
    vo, Fo, Co, Jpo = v.copy(), F.copy(), C.copy(), Jp.copy()
    gv = np.zeros((G, G, G, 3)); gm = np.zeros((G, G, G, 1))
    three_d.p2g(p["inv_dx"], p["hardening"], p["dx"], p["dt"], p["volume"], gv, gm,
                particles, vo, Fo, Co, Jpo, "snow")
    mom = gv.copy()
    three_d.grid_op(res, p["dx"], p["dt"], p["gravity"], gv, gm)
    three_d.g2p(p["inv_dx"], p["dt"], gv, particles, vo, Fo, Co, Jpo, "snow")
    save("snow3d", x=x, mass=mass, lam0=lam0, mu0=mu0, v=v, F=F, C=C, Jp=Jp,
         **{k: np.float64(val) for k, val in p.items()},
         **sparse_grid(mom, "grid_momentum"), **sparse_grid(gm, "grid_mass"),
         x_out=positions(particles), v_out=vo, F_out=Fo, C_out=Co, Jp_out=Jpo)


def gen_snow2d(ref):
    """2D snow branch: hardening by exp(h (1 - Jp)) in P2G (utils.py:27-49) and the
    singular-value clamp + Jp update in G2P (two_d/g2p.py:37-47), det F > 0 inputs."""
    rng = np.random.default_rng(8)
    n, res = 1500, 64
    p = dict(TEST2D, res=res, dt=1e-4, gravity=-9.8, mass=float(f32(1.0 / 64 ** 2 / 4)),
             volume=float(f32(1.0 / 64 ** 2 / 4)), hardening=10.0)
    x = f32(rng.uniform(0.1, 0.9, size=(n, 2)))
    v = f32(rng.normal(0, 0.5, size=(n, 2)))
    F = f32(np.eye(2) + rng.normal(0, 0.03, size=(n, 2, 2)))
    C = f32(rng.normal(0, 0.5, size=(n, 2, 2)))
    Jp = f32(rng.uniform(0.9, 1.1, size=(n, 1)))
    out = run_phases_2d(ref, p, x, v, F, C, Jp, model="snow")
    save("snow2d", x=x, v=v, F=F, C=C, Jp=Jp, **{k: np.float64(val) for k, val in p.items()}, **out)


def gen_collide3d(ref):
    """Plane colliders, three_d/grid_op.py:50-67 (check_collision_points), on a random velocity grid."""
    three_d = ref[0]
    from femflow.solvers.mpm.three_d.grid_op import check_collision_points
    rng = np.random.default_rng(9)
    res = 12; G = res + 1
    gv = rng.normal(0, 1, size=(G, G, G, 3))
    points = np.array([[0.2, 0.1, 0.3], [0.8, 0.5, 0.5], [0.5, 0.9, 0.1]])
    normals = np.array([[0.0, 1.0, 0.0], [-1.0, 0.2, 0.1], [0.3, -2.0, 0.5]])
    out = gv.copy()
    check_collision_points(points, normals, res, 1.0 / res, out)
    save("collide3d", grid_velocity=gv, points=points, normals=normals, res=np.float64(res), grid_velocity_out=out)


def main(argv):
    ref = _import_reference()
    gens = dict(kat3d=gen_kat3d, block3d=gen_block3d, rest3d=gen_rest3d, walls3d=gen_walls3d,
                c1=gen_c1, drift3d=gen_drift3d, test2d=gen_test2d, drift2d=gen_drift2d, block2d=gen_block2d,
                quirk2d=gen_quirk2d, snow3d=gen_snow3d, snow2d=gen_snow2d, collide3d=gen_collide3d)
    which = argv[1:] or list(gens)
    for name in which:
        print(f"== {name}")
        gens[name](ref)


if __name__ == "__main__":
    main(sys.argv)
