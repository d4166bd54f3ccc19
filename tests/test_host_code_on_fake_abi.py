"""The Python host code above the C ABI, on CPU: the real MpmSolver (buffer binding, material layouts, the
reordering ping-pong, id carrying, error convention) driven through tests/fake_abi.py, whose arithmetic is the
NumPy oracle.  What is under test is the HOST logic; the kernels are tested on the GPU and, for their
arithmetic, in test_kernel_math_host.py."""
import numpy as np
import pytest

torch = pytest.importorskip("torch")

import fake_abi  # noqa: E402
from oracle import mpm_oracle as O  # noqa: E402


def scene(n=600, res=16, seed=0, nmat=1):
    rng = np.random.default_rng(seed)
    dx = 1.0 / res
    p = dict(res=res, inv_dx=float(res), dx=dx, dt=1e-3, volume=(dx / 2) ** 3, gravity=-9.8, hardening=0.7)
    x = rng.uniform(0.2, 0.8, size=(n, 3))
    v = rng.normal(0, 0.5, size=(n, 3))
    F = np.eye(3) + rng.normal(0, 0.01, size=(n, 3, 3))
    C = rng.normal(0, 0.2, size=(n, 3, 3))
    k = (np.arange(n) % nmat).astype(np.float64)
    return p, (x, v, F, C, p["volume"] * (1 + 0.25 * k), 40.0 * (1 + 0.5 * k), 30.0 * (1 - 0.1 * k))


@pytest.mark.parametrize("nmat,layout", [(1, "table[1]"), (3, "table[3]"), (300, "planes")])
def test_mpm_solver_host_logic(monkeypatch, nmat, layout):
    """set_particles -> material layout -> 5 substeps over the ping-pong buffers -> get_particles in ORIGINAL order."""
    fake_abi.install(monkeypatch)
    from femflow_b200.mpm import MpmSolver
    p, (x, v, F, C, mass, mu0, lam0) = scene(nmat=nmat)
    n = len(x)
    s = MpmSolver(3, p["res"], p["dt"], p["volume"], p["gravity"], p["hardening"], capacity=n, dtype=torch.float64, device="cpu")
    s.set_particles(x, v, F, C, None, mass, mu0, lam0)
    assert s.material_layout == layout and s.num_particles == n
    s.substep(5)
    s.check_errors()
    assert s.live_index == 1                                     # odd number of reordering substeps
    ids = s.live.id[:n].numpy()
    assert not np.array_equal(ids, np.arange(n)) and np.array_equal(np.sort(ids), np.arange(n))   # really reordered
    out = s.get_particles()
    Jp = np.ones((n, 1))
    for _ in range(5):
        O.solve_mls_mpm_3d(p["res"], p["inv_dx"], p["hardening"], p["dx"], p["dt"], p["volume"], p["gravity"],
                           x, mass, mu0, lam0, v, F, C, Jp)
    for k, ref in (("x", x), ("v", v), ("F", F), ("C", C)):
        assert np.abs(out[k].numpy() - ref).max() <= 1e-12 * max(1.0, np.abs(ref).max()), k
    snap = torch.zeros(3 * n, dtype=torch.float64)
    s.snapshot(0.05, snap)
    assert np.allclose(snap.numpy().reshape(n, 3), x / 0.05, rtol=1e-15)
    # a particle outside the grid: the reference's RuntimeError (three_d/p2g.py:51-52)
    x2 = x.copy(); x2[7] = [0.999, 0.5, 0.5]
    s.set_particles(x2, v, F, C, None, mass, mu0, lam0)
    s.substep(1)
    with pytest.raises(RuntimeError):
        s.check_errors()
    with pytest.raises(ValueError):
        s.set_particles(np.zeros((n + 100, 3)), mass=mass[0], mu0=1.0, lam0=1.0)
    s.close()


def test_reference_signature_wrapper(monkeypatch):
    """femflow_b200.solvers.mpm.mls_mpm.solve_mls_mpm_3d: in-place semantics of the reference call."""
    fake_abi.install(monkeypatch)
    from femflow_b200.solvers.mpm import _runtime
    from femflow_b200.solvers.mpm.mls_mpm import make_mls_mpm_coefficients, solve_mls_mpm_3d
    from femflow_b200.solvers.mpm.particle import Particle
    monkeypatch.setattr(_runtime, "_default_dtype", torch.float64)
    _runtime.clear_cache()
    monkeypatch.setattr(_runtime, "MpmSolver", lambda *a, **k: __import__("femflow_b200.mpm", fromlist=["MpmSolver"]).MpmSolver(*a, device="cpu", **k))
    p, (x, v0, F0, C0, mass, mu0, lam0) = scene(n=200, nmat=2)
    particles = [Particle(x[i].copy(), 0.0, mass[i], lam0[i], mu0[i]) for i in range(len(x))]
    v, F, C, Jp = make_mls_mpm_coefficients(len(x), 3)
    v[:] = v0; F[:] = F0; C[:] = C0
    for _ in range(2):
        solve_mls_mpm_3d(p["res"], p["inv_dx"], p["hardening"], p["dx"], p["dt"], p["volume"], p["gravity"], particles, v, F, C, Jp)
    xr, vr, Fr, Cr = x.copy(), v0.copy(), F0.copy(), C0.copy()
    for _ in range(2):
        O.solve_mls_mpm_3d(p["res"], p["inv_dx"], p["hardening"], p["dx"], p["dt"], p["volume"], p["gravity"],
                           xr, mass, mu0, lam0, vr, Fr, Cr, np.ones((len(x), 1)))
    got_x = np.stack([q.pos for q in particles])
    assert np.abs(got_x - xr).max() < 1e-12 and np.abs(v - vr).max() < 1e-11 and np.abs(F - Fr).max() < 1e-12
    assert np.array_equal(Jp, np.ones_like(Jp))                  # untouched: the driver is neo-hookean only
    _runtime.clear_cache()


class _Mesh:
    def __init__(self, pts):
        self.vertices = np.asarray(pts, dtype=np.float32).reshape(-1)


def test_mpm_simulation_thread_and_snapshot_ring(monkeypatch, tmp_path):
    """MPMSimulation (reference simulation.py:19-149) end to end on the stand-in library: load / start / join, one
    displacement vector per substep in the reference's flat f64 `pos / tightening_coeff` format and ORIGINAL particle
    order (the ring of pinned buffers must neither drop, duplicate nor reorder a step), the .npy dump, final state."""
    fake_abi.install(monkeypatch)
    from femflow_b200.simulation.mpm import MPMSimulation
    rng = np.random.default_rng(1)
    coeff, res, steps = 0.05, 16, 7
    a = rng.uniform(4, 12, size=(150, 3)).astype(np.float32)          # mesh units; * coeff lands inside the unit grid
    b = rng.uniform(5, 11, size=(90, 3)).astype(np.float32)
    sim = MPMSimulation(str(tmp_path / "out"), steps, 1e-3, 1.0, 10.0, 1e-4, -9.8, 140, 1000, 0.2, 0.4, 0.7, res, coeff,
                        device="cpu", dtype="float64", progress=False)
    sim.load(meshes=[_Mesh(a), _Mesh(b)], params=[(1.0, 140, 0.2), (10.0, 1000, 0.4)])
    n = len(sim.particles)
    x0 = sim.particles.pos.copy()
    sim.start(); sim.join(60)
    assert sim.error is None and not sim.running and len(sim.displacements) == steps + 1
    from femflow_b200.solvers.mpm.utils import Ev_to_lambda, Ev_to_mu
    mass = np.r_[np.full(150, 1.0), np.full(90, 10.0)]
    lam = np.r_[np.full(150, Ev_to_lambda(140, 0.2)), np.full(90, Ev_to_lambda(1000, 0.4))]
    mu = np.r_[np.full(150, Ev_to_mu(140, 0.2)), np.full(90, Ev_to_mu(1000, 0.4))]
    x, v, F, C = x0.copy(), np.zeros((n, 3)), np.tile(np.eye(3), (n, 1, 1)), np.zeros((n, 3, 3))
    assert np.array_equal(sim.displacements[0], (x0 / coeff).reshape(-1))
    for k in range(steps):
        O.solve_mls_mpm_3d(res, float(res), 0.7, 1 / res, 1e-3, 1e-4, -9.8, x, mass, mu, lam, v, F, C, np.ones((n, 1)))
        assert np.abs(sim.displacements[k + 1] - (x / coeff).reshape(-1)).max() < 1e-9, k
    assert np.abs(sim.particles.pos - x).max() < 1e-12 and np.abs(sim.F - F).max() < 1e-12
    import os
    files = sorted(os.listdir(sim.outdir), key=lambda f: int(f.split(".")[0]))
    assert files == [f"{i}.npy" for i in range(steps + 1)]
    assert np.array_equal(np.load(os.path.join(sim.outdir, f"{steps}.npy")), sim.displacements[steps])


def test_headless_runner_on_the_stand_in(monkeypatch, tmp_path):
    """femflow_b200.simulation.mpm.headless.run(experiment 0): the whole runner, three substeps of the real scene."""
    fake_abi.install(monkeypatch)
    from femflow_b200.simulation.mpm import headless
    import femflow_b200.simulation.mpm.simulation as simmod
    real = simmod.MPMSimulation
    monkeypatch.setattr(simmod, "MPMSimulation", lambda *a, **k: real(*a, **{**k, "dtype": "float64"}))
    sim, seconds = headless.run(0, steps=3, outdir=str(tmp_path / "o"), device="cpu")
    assert sim.error is None and len(sim.displacements) == 4 and sim.displacements[-1].shape == (3 * 35321,)
    assert np.isfinite(sim.displacements[-1]).all()
    y0, y3 = sim.displacements[0][1::3], sim.displacements[3][1::3]
    assert (y3 < y0).mean() > 0.99                                  # everything is falling


def test_stand_in_grid_update_equals_the_oracle_on_a_cube():
    rng = np.random.default_rng(2)
    res, G = 10, 11
    gm = np.where(rng.random((G, G, G, 1)) < 0.5, rng.uniform(0.1, 2, (G, G, G, 1)), 0.0)
    gv = rng.normal(0, 20, (G, G, G, 3)) * (gm > 0)
    a, b = gv.copy(), gv.copy()
    O.grid_op_3d(res, 0.1, 2e-3, -9.8, a, gm.copy())
    fake_abi.FakeLib.box_grid_op([res] * 3, 0.1, 2e-3, -9.8, b, gm.copy())
    assert np.array_equal(a, b)


@pytest.fixture
def gpu_test_bodies(monkeypatch):
    """The bodies of GPU parity tests, run on the stand-in library in fp64: what they exercise here is the
    reference-signature wrappers' marshalling (typed lists / SoA, caller-owned grids in the reference's two-array
    layout, in-place updates, the error convention) against the goldens of the real reference."""
    fake_abi.install(monkeypatch)
    import test_gpu_parity as TP
    from femflow_b200.solvers.mpm import _runtime
    import femflow_b200.mpm as mpm
    monkeypatch.setattr(_runtime, "_default_dtype", torch.float64)
    monkeypatch.setattr(_runtime, "MpmSolver", lambda *a, **k: mpm.MpmSolver(*a, device="cpu", **k))
    _runtime.clear_cache()
    yield TP
    _runtime.clear_cache()


@pytest.mark.parametrize("kernels", ["direct", "production"])
@pytest.mark.parametrize("name", ["kat3d", "block3d", "rest3d", "walls3d"])
def test_phase_wrappers_against_reference_goldens(gpu_test_bodies, name, kernels):
    from femflow_b200.solvers.mpm import three_d
    prev = three_d.set_kernels(kernels)       # "production": the wrappers' bin + reorder marshalling (ids, ping-pong)
    try:
        gpu_test_bodies.test_3d_phase_functions_match_reference(name, "float64", kernels)
    finally:
        three_d.set_kernels(prev)


def test_solve_wrapper_on_the_paper_scene_and_error_convention(gpu_test_bodies):
    gpu_test_bodies.test_solve_mls_mpm_3d_c1_scene("auto", "float64")
    gpu_test_bodies.test_oob_raises_runtime_error("float64")


def test_graft_entry_smoke_body(monkeypatch, capsys):
    """__graft_entry__.smoke() -- the call the driver makes on the GPU box before the bench -- on the stand-in
    library (fp32 storage): its own code path, comparison and tolerance."""
    fake_abi.install(monkeypatch)
    import femflow_b200.mpm as mpm
    real = mpm.MpmSolver
    monkeypatch.setattr(mpm, "MpmSolver", lambda *a, **k: real(*a, **{**k, "device": "cpu"}))
    import __graft_entry__ as entry
    entry.smoke()
    assert "max-norm relative errors vs oracle" in capsys.readouterr().out


def test_solver_reuse_across_particle_counts_and_material_layouts(monkeypatch):
    """One MpmSolver / one cached wrapper solver fed, in turn, three materials, one material, 300 materials, fewer and
    then more particles than before, and none at all: every re-upload rebinds cleanly (no stale rows, planes or ids)."""
    fake_abi.install(monkeypatch)
    from femflow_b200.mpm import MpmSolver
    p, _ = scene()
    s = MpmSolver(3, p["res"], p["dt"], p["volume"], p["gravity"], p["hardening"], capacity=900, dtype=torch.float64, device="cpu")
    for n, nmat, layout in ((600, 3, "table[3]"), (250, 1, "table[1]"), (900, 300, "planes"), (400, 2, "table[2]"), (0, 1, "table[1]")):
        _, (x, v, F, C, mass, mu0, lam0) = scene(n=n, seed=n + nmat, nmat=nmat)
        s.set_particles(x, v, F, C, None, mass, mu0, lam0)
        assert s.material_layout == layout and s.num_particles == n
        s.substep(2)
        s.check_errors()
        out = s.get_particles()
        Jp = np.ones((n, 1))
        for _ in range(2):
            O.solve_mls_mpm_3d(p["res"], p["inv_dx"], p["hardening"], p["dx"], p["dt"], p["volume"], p["gravity"],
                               x, mass, mu0, lam0, v, F, C, Jp)
        assert out["x"].shape == (n, 3) and out["F"].shape == (n, 3, 3)
        if n:
            assert np.abs(out["x"].numpy() - x).max() < 1e-12 and np.abs(out["C"].numpy() - C).max() < 1e-9, (n, nmat)
    s.close()
    # the wrapper's solver cache: capacity grows, then is reused for a smaller call
    from femflow_b200.solvers.mpm import _runtime
    from femflow_b200.solvers.mpm.mls_mpm import make_mls_mpm_coefficients, solve_mls_mpm_3d
    from femflow_b200.solvers.mpm.particle import ParticleArray
    import femflow_b200.mpm as mpm
    monkeypatch.setattr(_runtime, "_default_dtype", torch.float64)
    monkeypatch.setattr(_runtime, "MpmSolver", lambda *a, **k: mpm.MpmSolver(*a, device="cpu", **k))
    _runtime.clear_cache()
    made = []
    for n in (100, 3000, 50):
        _, (x, v0, F0, C0, mass, mu0, lam0) = scene(n=n, seed=n)
        particles = ParticleArray(x.copy(), mass, lam0, mu0)
        v, F, C, Jp = make_mls_mpm_coefficients(n, 3)
        v[:] = v0; F[:] = F0; C[:] = C0
        solve_mls_mpm_3d(p["res"], p["inv_dx"], p["hardening"], p["dx"], p["dt"], p["volume"], p["gravity"], particles, v, F, C, Jp)
        O.solve_mls_mpm_3d(p["res"], p["inv_dx"], p["hardening"], p["dx"], p["dt"], p["volume"], p["gravity"],
                           x, mass, mu0, lam0, v0, F0, C0, np.ones((n, 1)))
        assert np.abs(particles.pos - x).max() < 1e-12 and np.abs(v - v0).max() < 1e-11
        made.append(id(next(iter(_runtime._cache.values()))))
    assert len(_runtime._cache) == 1 and made[0] != made[1] and made[1] == made[2]
    _runtime.clear_cache()


@pytest.mark.parametrize("name", ["test2d", "block2d"])
def test_2d_wrappers_against_reference_goldens(gpu_test_bodies, name):
    gpu_test_bodies.test_2d_phase_functions_match_reference(name, "float64")


def test_2d_snow_quirk_and_3d_snow_wrappers(gpu_test_bodies):
    gpu_test_bodies.test_2d_snow_matches_reference("float64")
    gpu_test_bodies.test_2d_svd_roundtrip_quirk("float64")
    gpu_test_bodies.test_snow_p2g_3d("float64")
    gpu_test_bodies.test_3d_snow_substeps_vs_oracle(device="cpu")
    from femflow_b200.solvers.mpm import three_d
    for kind in ("direct", "production"):
        prev = three_d.set_kernels(kind)
        try:
            gpu_test_bodies.test_snow_g2p_3d("float64", kind)
        finally:
            three_d.set_kernels(prev)
