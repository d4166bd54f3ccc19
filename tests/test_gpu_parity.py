"""GPU parity tests: the CUDA path (through the C ABI) against the oracle / goldens.

Tolerances (BASELINE.json north_star): indexing bit-exact; grid mass / momentum /
velocity and particle x, v, C, F within 1e-5 max-norm-relative per substep for the
fp32 build (absolute floors from the reference's own neighbouring fields, SURVEY
8d), 1e-11 for the fp64 build (only summation order differs).
"""
import os

import numpy as np
import pytest

from conftest import dense_grid, load_golden, rel_err
from oracle import mpm_oracle as O

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")

TOL = {"float32": 1e-5, "float64": 1e-11}


@pytest.fixture(scope="module", autouse=True)
def _need_cuda():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    yield
    from femflow_b200.solvers.mpm import _runtime
    _runtime.clear_cache()


@pytest.fixture(params=["float32", "float64"])
def dtype(request):
    from femflow_b200.solvers.mpm import _runtime
    _runtime.set_default_dtype(request.param)
    yield request.param
    _runtime.set_default_dtype("float32")


def _p3(g):
    return dict(res=int(g["res"]), inv_dx=float(g["inv_dx"]), dx=float(g["dx"]), dt=float(g["dt"]),
                volume=float(g["volume"]), hardening=float(g["hardening"]), gravity=float(g["gravity"]))


def _grid(g, name):
    return dense_grid(g, name) if f"{name}_idx" in g else g[name]


def _floors(p, mass, grid_velocity_ref):
    """V = max(|v_grid,ref|_inf, dt*|g|); momentum floor max(m_p)*V; C floor 4*inv_dx*V."""
    V = max(float(np.abs(grid_velocity_ref).max()), p["dt"] * abs(p["gravity"]))
    return dict(vel=V, mom=float(np.max(mass)) * V, C=4 * p["inv_dx"] * V)


# --------------------------------------------------------------------------- #
@pytest.fixture(params=["direct", "production"])
def kernels(request):
    """Kernel path of the reference-signature 3D phase functions (three_d.set_kernels): the thread-per-particle
    kernels on the caller's order, or what ffmpm_substep runs (binning, bulk P2G over cell runs, tiled G2P)."""
    from femflow_b200.solvers.mpm import three_d
    prev = three_d.set_kernels(request.param)
    yield request.param
    three_d.set_kernels(prev)


@pytest.mark.parametrize("name", ["kat3d", "block3d", "rest3d", "walls3d"])
def test_3d_phase_functions_match_reference(name, dtype, kernels):
    """three_d.p2g / grid_op / g2p with the reference's signatures vs the reference's outputs
    (three_d/p2g.py:14-80, grid_op.py:5-47, g2p.py:9-59), through both kernel paths."""
    from femflow_b200.solvers.mpm import three_d
    from femflow_b200.solvers.mpm.particle import ParticleArray
    g = load_golden(name)
    p = _p3(g)
    tol = TOL[dtype]
    G = p["res"] + 1
    fl = _floors(p, g["mass"], _grid(g, "grid_velocity"))
    particles = ParticleArray(g["x"].copy(), g["mass"], g["lam0"], g["mu0"])
    v, F, C, Jp = (g[k].copy() for k in ("v", "F", "C", "Jp"))
    gv = np.zeros((G, G, G, 3)); gm = np.zeros((G, G, G, 1))
    three_d.p2g(p["inv_dx"], p["hardening"], p["dx"], p["dt"], p["volume"], gv, gm, particles, v, F, C, Jp,
                "neo_hookean")
    assert rel_err(gm, _grid(g, "grid_mass")) < tol
    assert rel_err(gv, _grid(g, "grid_momentum"), fl["mom"]) < tol
    # feed the reference's momentum to the grid update so each phase is judged on its own
    gv = _grid(g, "grid_momentum").copy(); gm = _grid(g, "grid_mass").copy()
    three_d.grid_op(p["res"], p["dx"], p["dt"], p["gravity"], gv, gm)
    assert rel_err(gv, _grid(g, "grid_velocity"), fl["vel"]) < tol
    gv = _grid(g, "grid_velocity").copy()
    three_d.g2p(p["inv_dx"], p["dt"], gv, particles, v, F, C, Jp, "neo_hookean")
    assert rel_err(particles.pos, g["x_out"], 1.0) < tol
    assert rel_err(v, g["v_out"], fl["vel"]) < tol
    assert rel_err(F, g["F_out"], 1.0) < tol
    assert rel_err(C, g["C_out"], fl["C"]) < tol


@pytest.mark.parametrize("name", ["kat3d", "block3d", "rest3d", "walls3d"])
def test_3d_production_chain_matches_reference(name, dtype):
    """The phases chained on the device exactly as ffmpm_substep issues them -- binning, bulk P2G, the grid
    update over the node blocks listed by the binning (grid_op3_blocks_kernel only runs when nothing touches
    the grid in between), tiled reordering G2P -- against the reference's grid velocity and particle outputs."""
    from femflow_b200.mpm import MpmSolver
    g = load_golden(name)
    p = _p3(g)
    n = len(g["x"])
    fl = _floors(p, g["mass"], _grid(g, "grid_velocity"))
    s = MpmSolver(3, p["res"], p["dt"], p["volume"], p["gravity"], p["hardening"], capacity=max(n, 1), dx=p["dx"],
                  inv_dx=p["inv_dx"], dtype=getattr(torch, dtype), p2g_mode="tiled")
    s.set_particles(g["x"], g["v"], g["F"], g["C"], None, g["mass"], g["mu0"], g["lam0"])
    s.clear_grid(); s.bin(); s.p2g(); s.grid_op()
    vel = s.grid(readonly=True).double().cpu().numpy()
    # the chain is one substep; its error budget is the sum of the phases'
    assert rel_err(vel[..., :3], _grid(g, "grid_velocity"), fl["vel"]) < 2 * TOL[dtype]
    assert rel_err(vel[..., 3:4], _grid(g, "grid_mass")) < TOL[dtype]
    s.g2p()
    s.check_errors()
    out = {k: t.double().cpu().numpy() for k, t in s.get_particles().items()}
    assert rel_err(out["x"], g["x_out"], 1.0) < 2 * TOL[dtype]
    assert rel_err(out["v"], g["v_out"], fl["vel"]) < 2 * TOL[dtype]
    assert rel_err(out["F"], g["F_out"], 1.0) < 2 * TOL[dtype]
    assert rel_err(out["C"], g["C_out"], fl["C"]) < 2 * TOL[dtype]
    s.close()


@pytest.mark.parametrize("mode", ["scatter", "auto"])
def test_solve_mls_mpm_3d_c1_scene(mode, dtype):
    """BASELINE config 1 (paper scene, every 8th particle): 10 substeps through
    solve_mls_mpm_3d (reference signature) vs the reference's trajectory."""
    from femflow_b200.solvers.mpm.mls_mpm import make_mls_mpm_coefficients, solve_mls_mpm_3d
    from femflow_b200.solvers.mpm.particle import ParticleArray
    g = load_golden("c1_scene")
    x = np.concatenate([g["gyroid_vertices"], g["collider_vertices"]]).astype(np.float32)
    x = (x * np.float32(float(g["tightening_coeff"]))).astype(np.float64)
    particles = ParticleArray(x, g["mass"], g["lam0"], g["mu0"])
    v, F, C, Jp = make_mls_mpm_coefficients(len(x), 3)
    res = int(g["res"])
    tol = TOL[dtype]
    for step in range(1, 11):
        solve_mls_mpm_3d(res, float(res), float(g["hardening"]), 1.0 / res, float(g["dt"]), float(g["volume"]),
                         float(g["gravity"]), particles, v, F, C, Jp, p2g_mode=mode)
        if step in (1, 10):
            k = step  # error may accumulate linearly over substeps
            V = max(np.abs(g[f"v_{step}"]).max(), float(g["dt"]) * 9.8)
            assert rel_err(particles.pos, g[f"x_{step}"], 1.0) < tol * k
            assert rel_err(v, g[f"v_{step}"], V) < tol * k
            assert rel_err(F, g[f"F_{step}"], 1.0) < tol * k
            assert rel_err(C, g[f"C_{step}"], 4 * res * V) < tol * k
    assert np.all(Jp == 1.0)        # quirk 7: the 3D driver never touches Jp


def _tile_keys_3d(base, n_nodes):
    tiles = [(n - 2 + 3) // 4 for n in n_nodes]
    t = ((base[:, 0] >> 2) * tiles[1] + (base[:, 1] >> 2)) * tiles[2] + (base[:, 2] >> 2)
    return t * 64 + ((base[:, 0] & 3) << 4) + ((base[:, 1] & 3) << 2) + (base[:, 2] & 3), tiles


def test_binning_is_bit_exact(dtype):
    """Cell keys from the CUDA binning == keys from the fp64 oracle indexing (any
    resolution, incl. non power of two), and perm is a valid counting-sort order."""
    from femflow_b200.mpm import MpmSolver
    rng = np.random.default_rng(11)
    for res in (80, 64, 37):
        n = 50_000
        x = rng.uniform(0.0, (res - 1.5) / res, size=(n, 3)).astype(np.float32).astype(np.float64)
        # positions one fp32 ulp either side of the (k + 0.5)/res cell boundaries
        k = np.arange(100) % (res - 2)
        edge = ((k + 0.5) / res).astype(np.float32)
        x[:100] = np.nextafter(edge, np.float32(0)).astype(np.float64)[:, None]
        x[100:200] = np.nextafter(edge, np.float32(1)).astype(np.float64)[:, None]
        x[200:300] = edge.astype(np.float64)[:, None]
        s = MpmSolver(3, res, 1e-4, 1.0, -9.8, 1.0, capacity=n, dtype=getattr(torch, dtype))
        s.set_particles(x, mass=1.0, mu0=1.0, lam0=1.0)
        s.bin()
        keys, perm, off, n_cells = s.bin_results()
        keys, perm, off = keys.cpu().numpy(), perm.cpu().numpy(), off.cpu().numpy()
        base, _ = O.base_and_fx(x, float(res))
        want, tiles = _tile_keys_3d(base, [res + 1] * 3)
        assert n_cells == tiles[0] * tiles[1] * tiles[2] * 64
        assert np.array_equal(keys, want.astype(np.int32))
        assert np.array_equal(np.sort(perm), np.arange(n))
        sorted_keys = keys[perm]
        assert np.all(np.diff(sorted_keys) >= 0)
        counts = np.bincount(want, minlength=n_cells + 2)
        assert np.array_equal(off, np.concatenate([[0], np.cumsum(counts)[:-1]]))
        assert s.poll_error() == 0
        s.close()


@pytest.mark.parametrize("strain", [1e-4, 1.5e-3, 1e-2, 4e-2, 0.3])
def test_stress_accuracy_across_series_tiers(strain):
    """fp32 stress (three_d/p2g.py:57-65 via utils.py:75-92): the perturbation series switches degree
    with the strain (3 / 5 / 8 terms, fp64 Newton fallback beyond); with v = C = 0 the grid momentum
    IS the stress term, so its relative error measures the stress alone.  Same 1e-5 bar on every tier."""
    from femflow_b200.mpm import MpmSolver
    rng = np.random.default_rng(11)
    res, n = 32, 40_000
    f32 = lambda a: np.asarray(a, dtype=np.float32).astype(np.float64)
    x = f32(rng.uniform(0.25, 0.75, size=(n, 3)))
    v = np.zeros((n, 3)); C = np.zeros((n, 3, 3))
    F = f32(np.eye(3) + strain * rng.uniform(-1, 1, size=(n, 3, 3)))
    dx = 1.0 / res
    vol = float(f32((dx / 2) ** 3))
    mass = np.full(n, vol); mu0 = np.full(n, f32(4166.67)); lam0 = np.full(n, f32(2777.78))
    gv = np.zeros((res + 1,) * 3 + (3,)); gm = np.zeros((res + 1,) * 3 + (1,))
    O.p2g_3d(float(res), 1.0, dx, 1e-4, vol, gv, gm, x, mass, mu0, lam0, v, F, C, np.ones((n, 1)))
    for mode in ("scatter", "tiled"):
        s = MpmSolver(3, res, 1e-4, vol, 0.0, 1.0, capacity=n, p2g_mode=mode)
        s.set_particles(x, v, F, C, None, mass, mu0, lam0)
        s.clear_grid()
        if mode == "tiled":
            s.bin()
        s.p2g()
        assert s.poll_error() == 0
        g = s.grid().double().cpu().numpy()
        assert rel_err(g[..., :3], gv, np.abs(gv).max()) < TOL["float32"], (mode, strain)
        s.close()


@pytest.mark.parametrize("mode", ["scatter", "tiled"])
def test_large_block_vs_oracle(mode, dtype):
    """200k-particle perturbed block, res 64: full substep vs the NumPy oracle, per phase."""
    from femflow_b200.mpm import MpmSolver
    rng = np.random.default_rng(5)
    res, n = 64, 200_000
    f32 = lambda a: np.asarray(a, dtype=np.float32).astype(np.float64)
    x = f32(rng.uniform(0.2, 0.8, size=(n, 3)))
    v = f32(rng.normal(0, 0.1, size=(n, 3)))
    F = f32(np.eye(3) + rng.normal(0, 0.02, size=(n, 3, 3)))
    C = f32(rng.normal(0, 0.5, size=(n, 3, 3)))
    dx = 1.0 / res
    vol = f32((dx / 2) ** 3)
    mass = np.full(n, float(vol)); mu0 = np.full(n, f32(4166.67)); lam0 = np.full(n, f32(2777.78))
    p = dict(dt=1e-4, gravity=-9.8, inv_dx=float(res))
    s = MpmSolver(3, res, p["dt"], float(vol), p["gravity"], 1.0, capacity=n, dtype=getattr(torch, dtype),
                  p2g_mode=mode)
    s.set_particles(x, v, F, C, None, mass, mu0, lam0)
    xo, vo, Fo, Co, Jp = x.copy(), v.copy(), F.copy(), C.copy(), np.ones((n, 1))
    mom, gmass, vel = O.solve_mls_mpm_3d(res, float(res), 1.0, dx, p["dt"], float(vol), p["gravity"],
                                         xo, mass, mu0, lam0, vo, Fo, Co, Jp, return_grids=True)
    fl = _floors(p, mass, vel)
    tol = TOL[dtype]
    s.clear_grid()
    if mode == "tiled":
        s.bin()
    s.p2g()
    g = s.grid().double().cpu().numpy()
    assert rel_err(g[..., 3:4], gmass) < tol
    assert rel_err(g[..., :3], mom, fl["mom"]) < tol
    s.grid_op()
    g = s.grid().double().cpu().numpy()
    assert rel_err(g[..., :3], vel, fl["vel"]) < tol
    s.g2p()
    assert s.poll_error() == 0
    out = {k: t.double().cpu().numpy() for k, t in s.get_particles().items()}
    assert rel_err(out["x"], xo, 1.0) < tol
    assert rel_err(out["v"], vo, fl["vel"]) < tol
    assert rel_err(out["F"], Fo, 1.0) < tol
    assert rel_err(out["C"], Co, fl["C"]) < tol
    # a second substep runs from the re-ordered buffer
    O.solve_mls_mpm_3d(res, float(res), 1.0, dx, p["dt"], float(vol), p["gravity"], xo, mass, mu0, lam0, vo, Fo, Co, Jp)
    s.substep(1)
    out = {k: t.double().cpu().numpy() for k, t in s.get_particles().items()}
    assert rel_err(out["x"], xo, 1.0) < 2 * tol
    assert rel_err(out["v"], vo, fl["vel"]) < 2 * tol
    assert rel_err(out["F"], Fo, 1.0) < 2 * tol
    s.close()


def test_snow_p2g_3d(dtype):
    """Snow hardening in the 3D scatter (three_d/p2g.py:57-61, utils.py:27-49)."""
    from femflow_b200.solvers.mpm import three_d
    from femflow_b200.solvers.mpm.particle import ParticleArray
    g = load_golden("snow3d")
    p = _p3(g); G = p["res"] + 1
    particles = ParticleArray(g["x"].copy(), g["mass"], g["lam0"], g["mu0"])
    gv = np.zeros((G, G, G, 3)); gm = np.zeros((G, G, G, 1))
    three_d.p2g(p["inv_dx"], p["hardening"], p["dx"], p["dt"], p["volume"], gv, gm, particles,
                g["v"].copy(), g["F"].copy(), g["C"].copy(), g["Jp"].copy(), "snow")
    assert rel_err(gm, dense_grid(g, "grid_mass")) < TOL[dtype]
    assert rel_err(gv, dense_grid(g, "grid_momentum")) < TOL[dtype]


def test_snow_g2p_3d(dtype, kernels):
    """3D snow G2P (three_d/g2p.py:48-58): F <- U clip(sig) Vh^T -- the reference transposes numpy's Vh once more, so
    the result depends on the SIGNS of LAPACK's singular vectors (csrc/mpm_svd3.cuh walks DGESDD's operation sequence) --
    and Jp <- clip(Jp det(F_) / (det F + 1e-10), 0.6, 20), against the reference's own F_out / Jp_out, through the
    thread-per-particle and the reordering kernel."""
    from femflow_b200.solvers.mpm import three_d
    from femflow_b200.solvers.mpm.particle import ParticleArray
    g = load_golden("snow3d")
    p = _p3(g)
    gv = dense_grid(g, "grid_momentum").copy(); gm = dense_grid(g, "grid_mass").copy()
    O.grid_op_3d(p["res"], p["dx"], p["dt"], p["gravity"], gv, gm)          # the pinned oracle: g2p is judged on its own
    fl = _floors(p, g["mass"], gv)
    particles = ParticleArray(g["x"].copy(), g["mass"], g["lam0"], g["mu0"])
    v, F, C, Jp = (g[k].copy() for k in ("v", "F", "C", "Jp"))
    three_d.g2p(p["inv_dx"], p["dt"], gv, particles, v, F, C, Jp, "snow")
    tol = TOL[dtype]
    assert rel_err(particles.pos, g["x_out"], 1.0) < tol
    assert rel_err(v, g["v_out"], fl["vel"]) < tol
    assert rel_err(C, g["C_out"], fl["C"]) < tol
    assert rel_err(F, g["F_out"], 1.0) < tol
    assert rel_err(Jp, g["Jp_out"], 1.0) < tol
    assert np.abs(g["F_out"] - g["F"]).max() > 1e-3        # the return map did something to compare


def test_3d_snow_substeps_vs_oracle(device="cuda"):
    """ffmpm_substep with model = snow in 3D (binning, bulk/runs P2G with snow hardening, grid update, reordering G2P
    that carries F, the fp64 return map of csrc/mpm_svd3.cuh): three substeps against the chained oracle phases
    (three_d/p2g.py, grid_op.py, g2p.py with model="snow" -- a path the reference's own driver never takes,
    mls_mpm.py:58).  fp64 build only: once the clamp has made two singular values equal, the next substep's singular
    vectors hang on a perturbation of size dt*C, and what the reference then computes (U S Vh^T, not U S Vh) amplifies
    storage rounding by 1/gap -- the fp32 build is held to the reference on the single-substep golden instead."""
    from femflow_b200 import scenes
    from femflow_b200.mpm import MpmSolver
    sc = scenes.elastic_block(3, 32, 10, 2, seed=4)
    n = sc.n
    x, v, F, C = (a.astype(np.float64) for a in (sc.x, sc.v, sc.F, sc.C))
    Jp = np.ones((n, 1))
    hard = 3.0
    s = MpmSolver(3, sc.res, sc.dt, sc.volume, sc.gravity, hard, capacity=n, model="snow", dtype=torch.float64, device=device)
    s.set_particles(x, v, F, C, Jp, sc.mass, sc.mu_0, sc.lambda_0)
    m = np.full(n, sc.mass); mu = np.full(n, sc.mu_0); lam = np.full(n, sc.lambda_0)
    G = sc.res + 1
    steps = 3
    s.substep(steps)
    s.check_errors()
    for _ in range(steps):
        gv = np.zeros((G, G, G, 3)); gm = np.zeros((G, G, G, 1))
        O.p2g_3d(float(sc.res), hard, 1.0 / sc.res, sc.dt, sc.volume, gv, gm, x, m, mu, lam, v, F, C, Jp, "snow")
        O.grid_op_3d(sc.res, 1.0 / sc.res, sc.dt, sc.gravity, gv, gm)
        O.g2p_3d(float(sc.res), sc.dt, gv, x, v, F, C, Jp, "snow")
    out = {k: t.double().cpu().numpy() for k, t in s.get_particles().items()}
    V = max(np.abs(v).max(), sc.dt * 9.8)
    assert rel_err(out["x"], x, 1.0) < 1e-10
    assert np.abs(out["v"] - v).max() / V < 1e-8
    assert rel_err(out["F"], F, 1.0) < 1e-8
    assert rel_err(out["Jp"], Jp, 1.0) < 1e-8
    assert np.abs(Jp - 1.0).max() > 1e-4 and np.abs(F - sc.F).max() > 1e-3
    s.close()


def test_oob_raises_runtime_error(dtype):
    """three_d/p2g.py:51-52: a stencil outside [0, R] is a RuntimeError."""
    from femflow_b200.solvers.mpm.mls_mpm import make_mls_mpm_coefficients, solve_mls_mpm_3d
    from femflow_b200.solvers.mpm.particle import ParticleArray
    res = 8
    x = np.array([[0.5, 0.5, 0.5], [0.5, 0.5, (res - 0.4) / res]])
    particles = ParticleArray(x, 1.0, 1.0, 1.0)
    v, F, C, Jp = make_mls_mpm_coefficients(2, 3)
    for mode in ("scatter", "auto"):
        with pytest.raises(RuntimeError):
            solve_mls_mpm_3d(res, float(res), 1.0, 1 / res, 1e-4, 1.0, -9.8, particles, v, F, C, Jp, p2g_mode=mode)
    # empty input is a no-op
    e = ParticleArray(np.zeros((0, 3)), np.zeros(0), np.zeros(0), np.zeros(0))
    ve, Fe, Ce, Jpe = make_mls_mpm_coefficients(0, 3)
    solve_mls_mpm_3d(res, float(res), 1.0, 1 / res, 1e-4, 1.0, -9.8, e, ve, Fe, Ce, Jpe)


# ------------------------------- 2D ---------------------------------------- #
def _p2(g):
    return dict(res=int(g["res"]), dt=float(g["dt"]), gravity=float(g["gravity"]), mass=float(g["mass"]),
                volume=float(g["volume"]), hardening=float(g["hardening"]), mu_0=float(g["mu_0"]),
                lambda_0=float(g["lambda_0"]))


@pytest.mark.parametrize("name", ["test2d", "block2d"])
def test_2d_phase_functions_match_reference(name, dtype):
    from femflow_b200.solvers.mpm import two_d
    g = load_golden(name)
    p = _p2(g)
    tol = TOL[dtype]
    res = p["res"]; G = res + 1
    x, v, F, C, Jp = (g[k].copy() for k in ("x", "v", "F", "C", "Jp"))
    V = max(float(np.abs(g["grid_velocity"]).max()), p["dt"] * abs(p["gravity"]))
    gv = np.zeros((G, G, 2)); gm = np.zeros((G, G, 1))
    two_d.p2g(float(res), p["hardening"], p["mu_0"], p["lambda_0"], p["mass"], 1.0 / res, p["dt"], p["volume"],
              gv, gm, x, v, F, C, Jp)
    assert rel_err(gm, g["grid_mass"]) < tol
    assert rel_err(gv, g["grid_momentum"], p["mass"] * V) < tol
    gv = g["grid_momentum"].copy(); gm = g["grid_mass"].copy()
    two_d.grid_op(res, p["dt"], p["gravity"], gv, gm)
    assert rel_err(gv, g["grid_velocity"], V) < tol
    gv = g["grid_velocity"].copy()
    two_d.g2p(float(res), p["dt"], gv, x, v, F, C, Jp)
    assert rel_err(x, g["x_out"], 1.0) < tol
    assert rel_err(v, g["v_out"], V) < tol
    assert rel_err(F, g["F_out"], 1.0) < tol
    assert rel_err(C, g["C_out"], 4 * res * V) < tol
    assert rel_err(Jp, g["Jp_out"], 1.0) < tol


def test_2d_svd_roundtrip_quirk(dtype):
    """two_d/g2p.py:37-43 ``U @ diag(sig) @ Vh.T`` incl. det F < 0 (tests/golden/quirk2d.npz)."""
    from femflow_b200.solvers.mpm import two_d
    g = load_golden("quirk2d")
    res = int(g["res"]); G = res + 1
    x = g["x"].copy(); F = g["F"].copy(); n = len(x)
    v = np.zeros((n, 2)); C = np.zeros((n, 2, 2)); Jp = np.ones((n, 1))
    two_d.g2p(float(res), 1e-4, np.zeros((G, G, 2)), x, v, F, C, Jp)
    tol = 1e-6 if dtype == "float32" else 1e-12
    assert rel_err(F, g["F_out"]) < tol
    assert rel_err(Jp, g["Jp_out"]) < tol


def test_2d_wall_masks_bit_exact():
    """Quirk 6: the f64 wall predicates on i/R of two_d/grid_op.py:18-23, at R = 1024 and R = 80."""
    from femflow_b200.solvers.mpm import two_d
    for res in (80, 1024):
        G = res + 1
        gv = np.ones((G, G, 2)); gv[..., 1] = -1.0
        gm = np.ones((G, G, 1))
        ref = gv.copy()
        O.grid_op_2d(res, 1.0, 0.0, ref, gm.copy())
        two_d.grid_op(res, 1.0, 0.0, gv, gm)
        assert np.array_equal(gv, ref)


def test_2d_snow_matches_reference(dtype):
    """2D snow model end to end: snow_hardening in P2G (utils.py:27-49), singular-value clamp
    to [1-2.5e-2, 1+7.5e-3] and Jp update in G2P (two_d/g2p.py:37-47)."""
    from femflow_b200.solvers.mpm import two_d
    g = load_golden("snow2d")
    p = _p2(g)
    tol = TOL[dtype]
    res = p["res"]; G = res + 1
    x, v, F, C, Jp = (g[k].copy() for k in ("x", "v", "F", "C", "Jp"))
    V = max(float(np.abs(g["grid_velocity"]).max()), p["dt"] * abs(p["gravity"]))
    gv = np.zeros((G, G, 2)); gm = np.zeros((G, G, 1))
    two_d.p2g(float(res), p["hardening"], p["mu_0"], p["lambda_0"], p["mass"], 1.0 / res, p["dt"], p["volume"],
              gv, gm, x, v, F, C, Jp, "snow")
    assert rel_err(gv, g["grid_momentum"], p["mass"] * V) < tol
    gv = g["grid_velocity"].copy()
    two_d.g2p(float(res), p["dt"], gv, x, v, F, C, Jp, "snow")
    assert rel_err(x, g["x_out"], 1.0) < tol
    assert rel_err(F, g["F_out"], 1.0) < tol
    assert rel_err(C, g["C_out"], 4 * res * V) < tol
    assert rel_err(Jp, g["Jp_out"], 1.0) < tol


@pytest.mark.parametrize("mode", ["auto", "fused", "scatter"])
def test_multi_substep_pipelines_vs_oracle(mode):
    """20 substeps without touching the state in between: exercises the reordering ping-pong,
    the pre-binning emitted by G2P, the overlapped binning, and (mode "fused") the G2P2G kernel
    with its scatter-ahead grid."""
    from femflow_b200 import scenes
    from femflow_b200.mpm import MpmSolver
    from oracle import native as ON
    sc = scenes.elastic_block(3, 64, 20, 2, seed=3)
    n = sc.n
    x, v, F, C = (a.astype(np.float64) for a in (sc.x, sc.v, sc.F, sc.C))
    m = np.full(n, sc.mass); mu = np.full(n, sc.mu_0); lam = np.full(n, sc.lambda_0)
    s = MpmSolver(3, sc.res, sc.dt, sc.volume, sc.gravity, sc.hardening, capacity=n, p2g_mode=mode)
    s.set_particles(sc.x, sc.v, sc.F, sc.C, None, sc.mass, sc.mu_0, sc.lambda_0)
    for chunk in (1, 2, 7, 10):          # odd and even counts: both state buffers and both grids get used
        s.substep(chunk)
        for _ in range(chunk):
            ON.solve_mls_mpm_3d(sc.res, float(sc.res), sc.hardening, 1 / sc.res, sc.dt, sc.volume, sc.gravity,
                                x, m, mu, lam, v, F, C)
    s.check_errors()
    out = {k: t.double().cpu().numpy() for k, t in s.get_particles().items()}
    V = max(np.abs(v).max(), sc.dt * 9.8)
    steps = 20
    assert rel_err(out["x"], x, 1.0) < 1e-5 * steps
    assert np.abs(out["v"] - v).max() / V < 1e-5 * steps
    assert rel_err(out["F"], F, 1.0) < 1e-5 * steps
    assert np.abs(out["C"] - C).max() / (4 * sc.res * V) < 1e-5 * steps
    s.close()


@pytest.mark.parametrize("pipeline", ["binned", "direct"])
@pytest.mark.parametrize("model", ["neo_hookean", "snow"])
def test_2d_multi_substep_pipelines_vs_oracle(pipeline, model, dtype):
    """The binned 2D pipeline (csrc/mpm_2d.cuh: cell-run P2G, reordering G2P that pre-bins the next substep, node-block
    grid update and clear, binning on the internal stream) and the thread-per-particle one, 15 substeps in odd and even
    chunks against the C port of two_d/{p2g,grid_op,g2p}.py; particles moving ~0.2 cells per substep, some inside the
    wall bands of two_d/grid_op.py:18-23."""
    from femflow_b200 import scenes
    from femflow_b200.mpm import MpmSolver
    from oracle import native as ON
    sc = scenes.elastic_block(2, 128, 100, 2, seed=6)         # 100 of 128 cells: reaches into the 5 % wall bands
    n = sc.n
    rng = np.random.default_rng(6)
    x, v, F, C = (a.astype(np.float64) for a in (sc.x, sc.v, sc.F, sc.C))
    speed = 0.2 / sc.res / sc.dt
    v = np.float32(v + speed * np.array([0.7, -1.0]) + rng.normal(0, 0.02 * speed, v.shape)).astype(np.float64)
    Jp = np.ones((n, 1))
    hard = 1.0 if model == "neo_hookean" else 3.0
    s = MpmSolver(2, sc.res, sc.dt, sc.volume, sc.gravity, hard, capacity=n, mass=sc.mass, mu_0=sc.mu_0, lambda_0=sc.lambda_0,
                  model=model, dtype=getattr(torch, dtype), reorder=(pipeline == "binned"))
    assert s.reorder == (pipeline == "binned")
    s.set_particles(x, v, F, C, Jp)
    steps = 15
    for chunk in (1, 2, 5, 7):
        s.substep(chunk)
        for _ in range(chunk):
            if model == "snow":
                O.solve_mls_mpm_2d(sc.res, float(sc.res), hard, sc.mu_0, sc.lambda_0, sc.mass, 1 / sc.res, sc.dt, sc.volume,
                                   sc.gravity, x, v, F, C, Jp, "snow")
            else:
                ON.solve_mls_mpm_2d(sc.res, float(sc.res), hard, sc.mu_0, sc.lambda_0, sc.mass, 1 / sc.res, sc.dt, sc.volume,
                                    sc.gravity, x, v, F, C, Jp)
    s.check_errors()
    out = {k: t.double().cpu().numpy() for k, t in s.get_particles().items()}
    V = max(np.abs(v).max(), sc.dt * 9.8)
    tol = TOL[dtype] * steps
    assert rel_err(out["x"], x, 1.0) < tol
    assert np.abs(out["v"] - v).max() / V < tol
    assert rel_err(out["F"], F, 1.0) < tol
    assert np.abs(out["C"] - C).max() / (4 * sc.res * V) < tol
    assert rel_err(out["Jp"], Jp, 1.0) < tol
    s.close()


@pytest.mark.parametrize("order", ["lattice", "shuffled", "cell_sorted"])
def test_2d_window_kernels_any_particle_order(order, monkeypatch):
    """The warp-window 2D kernels (csrc/mpm_2d_window.cuh, the fp32 default) against the thread-per-particle kernels
    (FFMPM_2D_WINDOW=0) and the C port of two_d/{p2g,grid_op,g2p}.py: 9 particles per cell in lattice order (coherent
    windows: the shared-memory node tile), shuffled (bounding box too large: direct reductions) and sorted by base cell
    (up to 9 lanes share a cell: they take turns on the tile); a particle count that is not a multiple of the window,
    two particles outside the grid, one NaN; the raw grid after P2G and the state after 3 substeps (odd: both grids of
    the ping-pong pair get used)."""
    from femflow_b200 import scenes
    from femflow_b200.mpm import MpmSolver
    from oracle import native as ON
    sc = scenes.elastic_block(2, 128, 90, 3, seed=8)
    rng = np.random.default_rng(8)
    n = sc.n - 37
    x, v, F, C = (a[:n].astype(np.float64) for a in (sc.x, sc.v, sc.F, sc.C))
    speed = 0.15 / sc.res / sc.dt
    v = np.float32(v + speed * np.array([-0.6, 1.0])).astype(np.float64)
    if order == "shuffled":
        perm = rng.permutation(n)
    elif order == "cell_sorted":
        base = np.floor(np.float32(x) * np.float32(sc.res) - np.float32(0.5)).astype(np.int64)
        perm = np.argsort(base[:, 0] * (sc.res + 1) + base[:, 1], kind="stable")
    else:
        perm = np.arange(n)
    x, v, F, C = x[perm], v[perm], F[perm], C[perm]
    Jp = np.ones((n, 1))

    def make(window):
        monkeypatch.setenv("FFMPM_2D_WINDOW", "1" if window else "0")
        s = MpmSolver(2, sc.res, sc.dt, sc.volume, sc.gravity, 1.0, capacity=n, mass=sc.mass, mu_0=sc.mu_0, lambda_0=sc.lambda_0)
        s.set_particles(x, v, F, C, Jp)
        return s
    a, b = make(True), make(False)
    # raw grid {momentum, mass} after P2G
    grids = []
    for s in (a, b):
        s.clear_grid(); s.p2g()
        grids.append(s.grid(readonly=True).double().cpu().numpy().reshape(sc.res + 1, sc.res + 1, 4).copy())
    gv = np.zeros((sc.res + 1, sc.res + 1, 2)); gm = np.zeros((sc.res + 1, sc.res + 1, 1))
    O.p2g_2d(float(sc.res), 1.0, sc.mu_0, sc.lambda_0, sc.mass, 1 / sc.res, sc.dt, sc.volume, gv, gm, x, v, F, C, Jp)
    V = max(np.abs(v).max(), sc.dt * 9.8)
    for g in grids:
        assert rel_err(g[..., 2:3], gm) < 1e-5
        assert rel_err(g[..., :2], gv, gm.max() * V) < 1e-5
        assert np.all(g[..., 3] == 0)
    # three substeps; then out-of-grid particles are flagged, not scattered
    for s in (a, b):
        s.substep(3)
        s.check_errors()
    for _ in range(3):
        ON.solve_mls_mpm_2d(sc.res, float(sc.res), 1.0, sc.mu_0, sc.lambda_0, sc.mass, 1 / sc.res, sc.dt, sc.volume,
                            sc.gravity, x, v, F, C, Jp)
    V = max(np.abs(v).max(), sc.dt * 9.8)
    for s in (a, b):
        out = {k: t.double().cpu().numpy() for k, t in s.get_particles().items()}
        assert rel_err(out["x"], x, 1.0) < 3e-5
        assert np.abs(out["v"] - v).max() / V < 3e-5
        assert rel_err(out["F"], F, 1.0) < 3e-5
        assert np.abs(out["C"] - C).max() / (4 * sc.res * V) < 3e-5
        assert rel_err(out["Jp"], Jp, 1.0) < 3e-5
    bad = x.copy()
    bad[5] = [1.5, 0.5]; bad[n - 1] = [0.5, -0.2]; bad[100] = [np.nan, 0.5]
    for s in (a, b):
        s.set_particles(bad, v, F, C, Jp)
        s.clear_grid(); s.p2g()
        with pytest.raises(RuntimeError):
            s.check_errors()
        s.close()


def test_2d_graph_replay_matches_eager_substeps():
    """2D: a captured pair of substeps (P2G, grid update that also clears the idle grid, G2P -- the two grids ping-pong
    inside the capture) replayed three times against six eager substeps; an odd request is rounded up to an even count."""
    from femflow_b200 import scenes
    from femflow_b200.mpm import MpmSolver
    sc = scenes.elastic_block(2, 256, 128, 2, seed=2)

    def make():
        s = MpmSolver(2, sc.res, sc.dt, sc.volume, sc.gravity, 1.0, capacity=sc.n, mass=sc.mass, mu_0=sc.mu_0, lambda_0=sc.lambda_0)
        s.set_particles(sc.x, sc.v, sc.F, sc.C, np.ones((sc.n, 1)))
        return s
    a, b = make(), make()
    a.substep(9)
    b.substep(3)
    g = b.make_graph(1)
    assert b.graph_substeps == 2 and b.graph_launches == 6
    for _ in range(3):
        g.replay()
    torch.cuda.synchronize()
    a.check_errors(); b.check_errors()
    pa, pb = a.get_particles(), b.get_particles()
    V = max(float(pa["v"].abs().max()), sc.dt * 9.8)
    assert float((pa["x"] - pb["x"]).abs().max()) < 1e-5
    assert float((pa["v"] - pb["v"]).abs().max()) / V < 1e-4
    assert float((pa["F"] - pb["F"]).abs().max()) < 1e-5
    assert float((pa["C"] - pb["C"]).abs().max()) / (4 * sc.res * V) < 1e-4
    a.close(); b.close()


@pytest.mark.parametrize("variant", [0, 1, 5])
@pytest.mark.parametrize("n_materials", [1, 3, 300])
def test_p2g_kernel_variants(variant, n_materials, monkeypatch):
    """The binned P2G kernels selectable with FFMPM_P2G_VARIANT -- 0: through the counting-sort permutation,
    1: physical order without prefetch (p2g_runs3_kernel), 5: the production kernel (p2g_bulk3_kernel: cp.async window
    prefetch, adjacent slot pairs, warp-uniform series degree) -- with one material, a material table and material
    planes: the grid after P2G and 12 chained substeps against the oracle; a particle count that leaves the last
    window ragged (n % 64 != 0, odd)."""
    from femflow_b200 import scenes
    from femflow_b200.mpm import MpmSolver
    from oracle import native as ON
    monkeypatch.setenv("FFMPM_P2G_VARIANT", str(variant))
    sc = scenes.elastic_block(3, 64, 20, 2, seed=4)
    n = sc.n - 37
    rng = np.random.default_rng(0)
    k = rng.integers(0, n_materials, n).astype(np.float64)
    f32 = lambda a: np.asarray(a, dtype=np.float32).astype(np.float64)
    m = f32(sc.mass * (1 + 0.25 * k / n_materials)); mu = f32(sc.mu_0 * (1 + 0.5 * k / n_materials))
    lam = f32(sc.lambda_0 * (1 - 0.125 * k / n_materials))
    x, v, F, C = (a[:n].astype(np.float64) for a in (sc.x, sc.v, sc.F, sc.C))
    # a few particles far beyond the series (fp64 Newton path inside a window of series particles) and at every tier
    F = F.copy()
    F[5] = f32(np.eye(3) + 0.3 * rng.uniform(-1, 1, (3, 3)))
    F[1000:1064] = f32(np.eye(3) + 1e-4 * rng.uniform(-1, 1, (64, 3, 3)))
    F[2000:2064] = f32(np.eye(3) + 4e-3 * rng.uniform(-1, 1, (64, 3, 3)))
    s = MpmSolver(3, sc.res, sc.dt, sc.volume, sc.gravity, sc.hardening, capacity=n)
    s.set_particles(x, v, F, C, None, m, mu, lam)
    s.clear_grid(); s.bin(); s.p2g()
    G = sc.res + 1
    gv = np.zeros((G, G, G, 3)); gm = np.zeros((G, G, G, 1))
    O.p2g_3d(float(sc.res), sc.hardening, 1 / sc.res, sc.dt, sc.volume, gv, gm, x, m, mu, lam, v, F, C, np.ones((n, 1)))
    g = s.grid().double().cpu().numpy()
    assert s.poll_error() == 0
    assert rel_err(g[..., 3:], gm) < 1e-5 and rel_err(g[..., :3], gv) < 1e-5
    F[5] = np.eye(3)              # the chained run without the violently strained particle (its P2G is checked above)
    s.set_particles(x, v, F, C, None, m, mu, lam)
    steps = 12
    s.substep(steps)
    s.check_errors()
    for _ in range(steps):
        ON.solve_mls_mpm_3d(sc.res, float(sc.res), sc.hardening, 1 / sc.res, sc.dt, sc.volume, sc.gravity, x, m, mu, lam, v, F, C)
    out = {k_: t.double().cpu().numpy() for k_, t in s.get_particles().items()}
    V = max(np.abs(v).max(), sc.dt * 9.8)
    assert rel_err(out["x"], x, 1.0) < 1e-5 * steps
    assert np.abs(out["v"] - v).max() / V < 1e-5 * steps
    assert rel_err(out["F"], F, 1.0) < 1e-5 * steps
    assert np.abs(out["C"] - C).max() / (4 * sc.res * V) < 1e-5 * steps
    s.close()


@pytest.mark.parametrize("cells_per_substep", [0.1, 0.45])
def test_fast_moving_particles_vs_oracle(cells_per_substep):
    """Particles that change cell every few substeps (a block moving at 0.1 / 0.45 cells per substep in every
    direction, dt such that the CFL clamp of three_d/grid_op.py:25,34-36 is not reached): the reordering G2P places a
    particle by the cell it was in BEFORE advection, so the physical order P2G walks is fragmented -- the in-warp
    window sort of mpm_p2g_runs.cuh (p2g_sort_window) and the multi-pass run table both run here.  16 substeps
    against the C port."""
    from femflow_b200 import scenes
    from femflow_b200.mpm import MpmSolver
    from oracle import native as ON
    sc = scenes.elastic_block(3, 64, 24, 2, seed=8)
    n = sc.n
    rng = np.random.default_rng(8)
    x, v, F, C = (a.astype(np.float64) for a in (sc.x, sc.v, sc.F, sc.C))
    speed = cells_per_substep / sc.res / sc.dt
    v = np.float32(v + speed * np.array([1.0, -0.6, 0.8]) + rng.normal(0, 0.02 * speed, v.shape)).astype(np.float64)
    m = np.full(n, sc.mass); mu = np.full(n, sc.mu_0); lam = np.full(n, sc.lambda_0)
    s = MpmSolver(3, sc.res, sc.dt, sc.volume, 0.0, sc.hardening, capacity=n)
    s.set_particles(x, v, F, C, None, m, mu, lam)
    steps = 16
    for chunk in (1, 5, 10):
        s.substep(chunk)
        for _ in range(chunk):
            ON.solve_mls_mpm_3d(sc.res, float(sc.res), sc.hardening, 1 / sc.res, sc.dt, sc.volume, 0.0, x, m, mu, lam, v, F, C)
    s.check_errors()
    out = {k: t.double().cpu().numpy() for k, t in s.get_particles().items()}
    V = np.abs(v).max()
    assert rel_err(out["x"], x, 1.0) < 1e-5 * steps
    assert np.abs(out["v"] - v).max() / V < 1e-5 * steps
    assert rel_err(out["F"], F, 1.0) < 1e-5 * steps
    assert np.abs(out["C"] - C).max() / (4 * sc.res * V) < 1e-5 * steps
    s.close()


@pytest.mark.parametrize("ppc_side", [3, 4])
def test_dense_cells_vs_oracle(ppc_side):
    """27 / 64 particles per cell (a settled column, BASELINE configs[4]): a 64-slot P2G window then holds one to three
    long runs, which the production kernel cuts every 8 / 16 slots to keep its (run, slab) lanes busy.  8 substeps
    against the C port, grid after the first P2G included."""
    from femflow_b200 import scenes
    from femflow_b200.mpm import MpmSolver
    from oracle import native as ON
    sc = scenes.elastic_block(3, 32, 10, ppc_side, seed=12)
    n = sc.n
    x, v, F, C = (a.astype(np.float64) for a in (sc.x, sc.v, sc.F, sc.C))
    m = np.full(n, sc.mass); mu = np.full(n, sc.mu_0); lam = np.full(n, sc.lambda_0)
    s = MpmSolver(3, sc.res, sc.dt, sc.volume, sc.gravity, sc.hardening, capacity=n)
    s.set_particles(x, v, F, C, None, m, mu, lam)
    s.substep(1)                                   # the state is cell-sorted from here on: long runs
    ON.solve_mls_mpm_3d(sc.res, float(sc.res), sc.hardening, 1 / sc.res, sc.dt, sc.volume, sc.gravity, x, m, mu, lam, v, F, C)
    s.clear_grid(); s.bin(); s.p2g()
    G = sc.res + 1
    gv = np.zeros((G, G, G, 3)); gm = np.zeros((G, G, G, 1))
    O.p2g_3d(float(sc.res), sc.hardening, 1 / sc.res, sc.dt, sc.volume, gv, gm, x, m, mu, lam, v, F, C, np.ones((n, 1)))
    g = s.grid().double().cpu().numpy()
    assert rel_err(g[..., 3:], gm) < 1e-5 and rel_err(g[..., :3], gv) < 1e-5
    steps = 7
    s.substep(steps)
    s.check_errors()
    for _ in range(steps):
        ON.solve_mls_mpm_3d(sc.res, float(sc.res), sc.hardening, 1 / sc.res, sc.dt, sc.volume, sc.gravity, x, m, mu, lam, v, F, C)
    out = {k_: t.double().cpu().numpy() for k_, t in s.get_particles().items()}
    V = max(np.abs(v).max(), sc.dt * 9.8)
    assert rel_err(out["x"], x, 1.0) < 1e-5 * (steps + 1)
    assert np.abs(out["v"] - v).max() / V < 1e-5 * (steps + 1)
    assert rel_err(out["F"], F, 1.0) < 1e-5 * (steps + 1)
    assert np.abs(out["C"] - C).max() / (4 * sc.res * V) < 1e-5 * (steps + 1)
    s.close()


@pytest.mark.parametrize("substeps", [1, 2])
def test_host_pipeline_returns_callers_order(substeps):
    """femflow_b200.host_pipeline.HostSubstepPipeline: seven independent host-side states through three device slots
    (uploads under downloads), every result in the CALLER's particle order (ffmpm_export_state undoes the cell sort),
    equal to the blocking path (set_particles -> substep -> get_particles) up to the order of the grid atomics, and to
    the oracle within the bar; the solver gets its own buffers back."""
    from femflow_b200 import scenes
    from femflow_b200.host_pipeline import HostSubstepPipeline
    from femflow_b200.mpm import MpmSolver
    from oracle import native as ON
    sc = scenes.elastic_block(3, 64, 16, 2, seed=2)
    n = sc.n
    rng = np.random.default_rng(2)
    order = rng.permutation(n)                       # the caller's order is NOT cell order
    s = MpmSolver(3, sc.res, sc.dt, sc.volume, sc.gravity, sc.hardening, capacity=n)
    s.set_particles(sc.x[order], sc.v[order], sc.F[order], sc.C[order], None, sc.mass, sc.mu_0, sc.lambda_0)
    s.substep(1)
    own = {k: t.clone() for k, t in s.get_particles().items()}
    calls = []
    for k in range(7):
        x = (sc.x[order] + np.float32(0.002 * k)).astype(np.float32)
        v = (sc.v[order] * np.float32(1 + 0.1 * k)).astype(np.float32)
        calls.append((x, v, sc.F[order], sc.C[order]))

    def soa(a):
        return torch.from_numpy(np.ascontiguousarray(a.reshape(n, -1).T)).pin_memory()
    ins = [{"x": soa(x), "v": soa(v), "F": soa(F), "C": soa(C)} for x, v, F, C in calls]
    outs = [{k: torch.empty_like(t).pin_memory() for k, t in d.items()} for d in ins]
    pipe = HostSubstepPipeline(s, depth=3, substeps_per_call=substeps)
    for d_in, d_out in zip(ins, outs):
        pipe.submit(d_in, d_out)
    pipe.drain()
    back = s.get_particles()
    for k in own:
        assert torch.equal(own[k], back[k]), k               # the solver's own state is untouched
    m = np.full(n, sc.mass); mu = np.full(n, sc.mu_0); lam = np.full(n, sc.lambda_0)
    for (x, v, F, C), d_out in zip(calls, outs):
        xo, vo, Fo, Co = (a.astype(np.float64) for a in (x, v, F, C))
        for _ in range(substeps):
            ON.solve_mls_mpm_3d(sc.res, float(sc.res), sc.hardening, 1 / sc.res, sc.dt, sc.volume, sc.gravity, xo, m, mu, lam, vo, Fo, Co)
        V = max(np.abs(vo).max(), sc.dt * 9.8)
        got = {k: t.numpy().T.astype(np.float64) for k, t in d_out.items()}
        assert rel_err(got["x"], xo, 1.0) < 1e-5 * substeps
        assert np.abs(got["v"] - vo).max() / V < 1e-5 * substeps
        assert rel_err(got["F"].reshape(n, 3, 3), Fo, 1.0) < 1e-5 * substeps
        assert np.abs(got["C"].reshape(n, 3, 3) - Co).max() / (4 * sc.res * V) < 1e-5 * substeps
    s.close()


def test_graph_replay_matches_eager_substeps():
    """MpmSolver.make_graph: 3 replays of a captured pair of substeps (internal binning stream and both
    ping-pong halves inside the capture) against the same 6 substeps launched one by one, from the same
    warmed-up state.  Only the order of the atomic sums may differ (scripts/graph_experiment.py,
    profiles/r01q_graph_experiment.json: x 6e-8, v 8e-7, F 2.4e-7 absolute)."""
    from femflow_b200 import scenes
    from femflow_b200.mpm import MpmSolver
    sc = scenes.elastic_block(3, 64, 24, 2, seed=1)

    def make():
        s = MpmSolver(3, sc.res, sc.dt, sc.volume, sc.gravity, sc.hardening, capacity=sc.n)
        s.set_particles(sc.x, sc.v, sc.F, sc.C, None, sc.mass, sc.mu_0, sc.lambda_0)
        return s
    a, b = make(), make()
    a.substep(8)
    b.substep(2)                          # steady state: the live buffer is pre-binned by the last G2P
    g = b.make_graph(2)
    assert b.graph_substeps == 2 and b.graph_launches > 0
    for _ in range(3):
        g.replay()
    torch.cuda.synchronize()
    a.check_errors(); b.check_errors()
    pa, pb = a.get_particles(), b.get_particles()
    V = max(float(pa["v"].abs().max()), sc.dt * 9.8)
    assert float((pa["x"] - pb["x"]).abs().max()) < 1e-5
    assert float((pa["v"] - pb["v"]).abs().max()) / V < 1e-4
    assert float((pa["F"] - pb["F"]).abs().max()) < 1e-5
    assert float((pa["C"] - pb["C"]).abs().max()) / (4 * sc.res * V) < 1e-4
    a.close(); b.close()


@pytest.mark.parametrize("n_materials", [1, 3, 300])
@pytest.mark.parametrize("mode", ["auto", "fused", "scatter"])
def test_material_layouts_agree(n_materials, mode):
    """Particle.mass / mu_0 / lambda_0 (particle.py:7-13) as a material table + 1-byte rows (chosen
    automatically for <= 256 distinct triples; no row plane at all for one) feeds the kernels the same
    values as three scalar planes, through reordering, pre-binning and the prefetching kernels; more
    than 256 triples fall back to planes.  Both layouts are pinned to the oracle (1e-5 per substep);
    against each other they may only differ by the order of the fp32 grid atomics (1e-6 per substep)."""
    from femflow_b200 import scenes
    from femflow_b200.mpm import MpmSolver
    from oracle import native as ON
    sc = scenes.elastic_block(3, 64, 20, 2, seed=5)
    n = sc.n
    rng = np.random.default_rng(7)
    row = rng.integers(0, n_materials, n)
    scale = 1.0 + 0.5 * np.arange(n_materials) / max(n_materials, 1)
    m = (sc.mass * scale[row]).astype(np.float32).astype(np.float64)
    mu = (sc.mu_0 * scale[::-1][row]).astype(np.float32).astype(np.float64)
    lam = (sc.lambda_0 * scale[row]).astype(np.float32).astype(np.float64)
    outs = {}
    for ppm in (None, True):
        s = MpmSolver(3, sc.res, sc.dt, sc.volume, sc.gravity, sc.hardening, capacity=n, p2g_mode=mode,
                      per_particle_material=ppm)
        s.set_particles(sc.x, sc.v, sc.F, sc.C, None, m, mu, lam)
        if ppm is None:
            want = "planes" if n_materials > 256 else f"table[{n_materials}]"
            assert s.material_layout == want
            assert (s.buffers[0].material is not None) == (1 < n_materials <= 256)
            assert (s.buffers[0].mass is not None) == (n_materials > 256)
        else:
            assert s.material_layout == "planes"
        for chunk in (1, 2, 4):
            s.substep(chunk)
        s.check_errors()
        outs[ppm] = {k: t.cpu().numpy() for k, t in s.get_particles().items()}
        s.close()
    x, v, F, C = (a.astype(np.float64) for a in (sc.x, sc.v, sc.F, sc.C))
    for _ in range(7):
        ON.solve_mls_mpm_3d(sc.res, float(sc.res), sc.hardening, 1 / sc.res, sc.dt, sc.volume, sc.gravity,
                            x, m, mu, lam, v, F, C)
    V = max(np.abs(v).max(), sc.dt * 9.8)
    for o in outs.values():
        assert rel_err(o["x"].astype(np.float64), x, 1.0) < 7e-5
        assert np.abs(o["v"] - v).max() / V < 7e-5
        assert rel_err(o["F"].astype(np.float64), F, 1.0) < 7e-5
        assert np.abs(o["C"] - C).max() / (4 * sc.res * V) < 7e-5
    a, b = outs[None], outs[True]
    assert np.abs(a["x"] - b["x"]).max() < 7e-6
    assert np.abs(a["v"] - b["v"]).max() / V < 7e-6
    assert np.abs(a["F"] - b["F"]).max() < 7e-6


def test_block_list_grid_update_matches_oracle(dtype, monkeypatch):
    """three_d/grid_op.py:25-47 over the node blocks listed by the binning (what ffmpm_substep runs)
    and over the whole grid (FFMPM_SPARSE_GRID_OP=0): both against the oracle's velocities, every node
    -- a sparse cloud that touches walls, with whole empty regions in between."""
    from femflow_b200.mpm import MpmSolver
    rng = np.random.default_rng(21)
    res, n = 48, 30_000
    f32 = lambda a: np.asarray(a, dtype=np.float32).astype(np.float64)
    centers = np.array([[0.08, 0.5, 0.5], [0.9, 0.12, 0.85], [0.5, 0.93, 0.3]])
    x = f32(np.clip(centers[rng.integers(0, 3, n)] + rng.normal(0, 0.03, (n, 3)), 0.03, 0.96))
    v = f32(rng.normal(0, 0.5, (n, 3)))
    F = f32(np.eye(3) + rng.normal(0, 0.01, (n, 3, 3)))
    C = f32(rng.normal(0, 0.5, (n, 3, 3)))
    dx = 1.0 / res
    vol = float(f32((dx / 2) ** 3))
    mass = np.full(n, vol); mu0 = np.full(n, f32(4166.67)); lam0 = np.full(n, f32(2777.78))
    dt, grav = 1e-4, -9.8
    gv = np.zeros((res + 1,) * 3 + (3,)); gm = np.zeros((res + 1,) * 3 + (1,))
    O.p2g_3d(float(res), 1.0, dx, dt, vol, gv, gm, x, mass, mu0, lam0, v, F, C, np.ones((n, 1)))
    O.grid_op_3d(res, dx, dt, grav, gv, gm)
    fl = _floors(dict(dt=dt, gravity=grav, inv_dx=float(res)), mass, gv)
    for sparse in ("1", "0"):
        monkeypatch.setenv("FFMPM_SPARSE_GRID_OP", sparse)
        s = MpmSolver(3, res, dt, vol, grav, 1.0, capacity=n, dtype=getattr(torch, dtype), p2g_mode="tiled")
        s.set_particles(x, v, F, C, None, mass, mu0, lam0)
        s.clear_grid()
        s.bin()
        s.p2g()
        s.grid_op()          # no grid access in between: the block-list path when enabled
        assert s.poll_error() == 0
        g = s.grid().double().cpu().numpy()
        assert rel_err(g[..., :3], gv, fl["vel"]) < TOL[dtype], sparse
        assert rel_err(g[..., 3:4], gm) < TOL[dtype]
        s.close()


def test_collision_planes(dtype):
    """Plane colliders fused into the grid update vs the reference (three_d/grid_op.py:50-67)."""
    from femflow_b200.solvers.mpm import three_d
    from femflow_b200.mpm import MpmSolver
    g = load_golden("collide3d")
    res = int(g["res"])
    gv = g["grid_velocity"].copy()
    three_d.check_collision_points(g["points"], g["normals"], res, 1.0 / res, gv)
    want = g["grid_velocity_out"]
    assert np.array_equal(gv == 0, want == 0)
    assert rel_err(gv, want) < (1e-6 if dtype == "float32" else 1e-15)
    # and fused at the end of the grid update: same result as the stand-alone function applied afterwards
    s = MpmSolver(3, 16, 1e-3, 1e-4, -9.8, 1.0, capacity=64, dtype=getattr(torch, dtype))
    x = np.random.default_rng(0).uniform(0.3, 0.7, size=(40, 3))
    s.set_particles(x, v=np.tile([0.3, -1.0, 0.2], (40, 1)), mass=1e-3, mu0=10.0, lam0=10.0)
    planes = ([[0.5, 0.5, 0.5], [0.2, 0.9, 0.2]], [[0.0, 1.0, 0.0], [1.0, -3.0, 0.5]])
    s.clear_grid(); s.p2g(); s.grid_op()
    plain = s.grid().double().cpu().numpy()[..., :3].copy()
    s.set_colliders(*planes)
    s.clear_grid(); s.p2g(); s.grid_op()
    fused = s.grid().double().cpu().numpy()[..., :3]
    O.check_collision_points(planes[0], planes[1], 16, 1.0 / 16, plain)
    assert np.array_equal(fused, plain)
    assert np.any(plain != 0) and np.any((fused == 0).all(-1) & (s.grid().cpu().numpy()[..., 3] > 0))
    with pytest.raises(Exception):
        s.set_colliders(np.zeros((9, 3)), np.ones((9, 3)))      # more than FFMPM_MAX_COLLIDERS
    s.close()


def test_c_abi_error_codes():
    """Error convention of the boundary: negative codes + ffmpm_last_error, never an exception or a crash."""
    import ctypes as C
    from femflow_b200 import _native as N
    lib = N.lib()
    cfg = N.FfMpmConfig()
    cfg.dim = 3
    for i in range(3):
        cfg.res[i] = 16; cfg.n[i] = 17
    cfg.dx, cfg.inv_dx, cfg.dt, cfg.volume, cfg.hardening = 1 / 16, 16.0, 1e-4, 1.0, 1.0
    h = N.H()
    assert lib.ffmpm_create(C.byref(cfg), 99, C.byref(h)) == N.FFMPM_E_INVALID          # no such device
    assert lib.ffmpm_create(C.byref(cfg), 0, C.byref(h)) == 0
    stream = C.c_void_p(0)
    assert lib.ffmpm_substep(h, 1, stream) == N.FFMPM_E_STATE                            # no workspace yet
    nbytes = lib.ffmpm_workspace_bytes(C.byref(cfg), 1000)
    ws = torch.zeros(nbytes, dtype=torch.uint8, device="cuda")
    assert lib.ffmpm_set_workspace(h, ws.data_ptr() + 8, nbytes - 8) == N.FFMPM_E_INVALID   # misaligned
    assert lib.ffmpm_set_workspace(h, ws.data_ptr(), 1024) == N.FFMPM_E_STATE            # too small for the grid
    assert lib.ffmpm_set_workspace(h, ws.data_ptr(), nbytes) == 0
    assert lib.ffmpm_substep(h, 1, stream) == N.FFMPM_E_STATE                            # state not bound
    planes = {k: torch.zeros((m, 1024), device="cuda") for k, m in (("x", 3), ("v", 3), ("C", 9), ("F", 9))}
    st = N.FfMpmState(planes["x"].data_ptr(), planes["v"].data_ptr(), planes["C"].data_ptr(), planes["F"].data_ptr(),
                      None, planes["x"].data_ptr(), None, None, None, None, 1024)         # mass without mu0/lam0
    assert lib.ffmpm_bind_state(h, C.byref(st), None, 10) == N.FFMPM_E_INVALID
    st = N.FfMpmState(planes["x"].data_ptr(), planes["v"].data_ptr(), planes["C"].data_ptr(), planes["F"].data_ptr(),
                      None, None, None, None, None, None, 1024)
    assert lib.ffmpm_bind_state(h, C.byref(st), None, 2000) == N.FFMPM_E_INVALID         # n > stride
    rows = torch.zeros(1024, dtype=torch.uint8, device="cuda")
    st.material = rows.data_ptr()
    assert lib.ffmpm_bind_state(h, C.byref(st), None, 10) == 0
    assert lib.ffmpm_substep(h, 1, stream) == N.FFMPM_E_STATE                            # rows without a table
    one = (C.c_double * 1)(1.0)
    assert lib.ffmpm_set_materials(h, one, one, one, 257) == N.FFMPM_E_INVALID           # > FFMPM_MAX_MATERIALS
    assert lib.ffmpm_set_materials(h, None, one, one, 1) == N.FFMPM_E_INVALID
    st.material = None
    assert lib.ffmpm_bind_state(h, C.byref(st), None, 10) == 0
    planes["F"][[0, 4, 8], :10] = 1.0
    planes["x"][:, :10] = 0.5
    planes["x"][2, 0] = 5.0                                                               # one particle far outside
    assert lib.ffmpm_substep(h, 1, stream) == 0
    code, n_oob = C.c_int32(), C.c_int64()
    assert lib.ffmpm_poll_error(h, stream, C.byref(code), C.byref(n_oob)) == N.FFMPM_E_OOB and n_oob.value >= 1
    assert lib.ffmpm_poll_error(h, stream, C.byref(code), C.byref(n_oob)) == 0           # sticky flag was cleared
    assert b"" != lib.ffmpm_last_error()
    lib.ffmpm_destroy(h)
