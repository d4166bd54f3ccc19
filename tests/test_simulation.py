"""Drop-in surface above the hot path: scene generators and MPMSimulation
(reference femflow/simulation/mpm/{primitives,simulation}.py, paper_1.py:75-114)."""
import os
import time

import numpy as np
import pytest

from conftest import load_golden, rel_err


class FakeMesh:
    """What MPMSimulation.load reads from femflow.viz.mesh.Mesh: a flat float32 vertex vector."""

    def __init__(self, pts):
        self.vertices = np.asarray(pts, dtype=np.float32).reshape(-1)


def test_scene_generators_reproduce_the_paper_scene():
    """paper_1.multi_drop_experiment(0): 8321 gyroid + 27000 collider points."""
    from femflow_b200.simulation.mpm import primitives as P
    g = load_golden("c1_scene")
    gy = P.generate_implicit_points("gyroid", 0.2, 0.3, 30).astype(np.float32)
    gy[:, 1] += np.float32(0.1)
    assert len(gy) == int(g["n_gyroid_full"]) == 8321
    assert np.array_equal(gy[::8], g["gyroid_vertices"])
    lo, hi = gy.min(0), gy.max(0)
    c = P.generate_cube_points((lo[0], hi[0]), (lo[1], hi[1]), (lo[2], hi[2]), 30).astype(np.float32)
    c[:, 1] += np.float32(3)
    assert len(c) == int(g["n_collider_full"]) == 27000
    assert np.array_equal(c[::8], g["collider_vertices"])
    with pytest.raises(ValueError):
        P.generate_implicit_points("nope", 0.2, 0.3, 4)
    # numerics/geometry.py:101-116 lattice: axis 0 fastest, endpoints included
    lat = P.grid(np.array((3, 2, 2)))
    assert lat.shape == (12, 3) and np.allclose(lat[1], (0.5, 0, 0)) and np.allclose(lat[-1], (1, 1, 1))


@pytest.mark.gpu
def test_scene_generators_on_the_gpu():
    """csrc/mpm_scene.cuh (ffmpm_gen_implicit_points / ffmpm_gen_cube_points) against the host generators, which are
    pinned to the reference's own output point for point (test_scene_generators_reproduce_the_paper_scene): every
    implicit function, resolutions that leave ragged last blocks, the paper scene (8 321 + 27 000 points), res 1."""
    torch = pytest.importorskip("torch")
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from femflow_b200.simulation.mpm import primitives as P
    for kind, k, t, res in (("gyroid", 0.2, 0.3, 30), ("gyroid", 0.2, 0.3, 40), ("diamond", 0.35, 0.1, 17),
                            ("primitive", 0.5, -0.2, 23), ("gyroid", 0.2, 0.3, 1), ("primitive", 1.0, 5.0, 9)):
        want = P.generate_implicit_points(kind, k, t, res)
        got = P.generate_implicit_points_gpu(kind, k, t, res).cpu().numpy()
        assert got.shape == want.shape, (kind, res, got.shape, want.shape)
        assert np.array_equal(got, want), (kind, res)
    assert len(P.generate_implicit_points_gpu("gyroid", 0.2, 0.3, 30)) == 8321
    for bounds, res in ((((0.0, 1.0), (0.1, 0.2050000041723251), (-3.0, 2.5)), 30), (((0.25, 0.75),) * 3, 7), (((1.0, 2.0),) * 3, 1)):
        want = P.generate_cube_points(*bounds, res)
        got = P.generate_cube_points_gpu(*bounds, res).cpu().numpy()
        assert np.array_equal(got, want), (bounds, res)
    with pytest.raises(ValueError):
        P.generate_implicit_points_gpu("nope", 0.2, 0.3, 4)
    # a lattice no host loop would want to walk: 512^3 = 1.3e8 points
    big = P.generate_implicit_points_gpu("gyroid", 0.2, 0.3, 512)
    assert 0.2 * 512 ** 3 < len(big) < 0.6 * 512 ** 3 and bool(torch.isfinite(big).all())


@pytest.mark.gpu
def test_mpm_simulation_matches_reference_trajectory(tmp_path):
    torch = pytest.importorskip("torch")
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from femflow_b200.simulation.mpm import MPMSimulation
    g = load_golden("c1_scene")
    outdir = str(tmp_path / "out")
    # constructor arguments exactly as paper_1.py:52-67 passes them
    sim = MPMSimulation(outdir, 10, float(g["dt"]), 1.0, 10.0, float(g["volume"]), float(g["gravity"]), 140, 1000,
                        0.2, 0.4, float(g["hardening"]), int(g["res"]), float(g["tightening_coeff"]), progress=False)
    sim.start()                       # not loaded yet: logs an error and returns (simulation.py:114-116)
    assert not sim.running and not sim.loaded
    meshes = [FakeMesh(g["gyroid_vertices"]), FakeMesh(g["collider_vertices"])]
    with pytest.raises(ValueError):
        sim.load(meshes=meshes, params=[(1.0, 140, 0.2)])
    sim.load(meshes=meshes, params=[(1.0, 140, 0.2), (10.0, 1000, 0.4)])
    assert sim.loaded and len(sim.displacements) == 1
    n = len(sim.particles)
    assert sim.displacements[0].shape == (3 * n,)
    sim.start()
    sim.join(120)
    assert not sim.running and sim.error is None
    assert len(sim.displacements) == 11
    coeff = float(g["tightening_coeff"])
    for step in (1, 10):
        want = (g[f"x_{step}"] / coeff).reshape(-1)
        assert rel_err(sim.displacements[step], want) < 1e-5 * step
    assert rel_err(sim.particles.pos, g["x_10"], 1.0) < 1e-4
    assert rel_err(sim.F, g["F_10"], 1.0) < 1e-4
    files = sorted(os.listdir(outdir))
    assert len(files) == 11 and "0.npy" in files and "10.npy" in files
    assert np.array_equal(np.load(os.path.join(outdir, "10.npy")), sim.displacements[10])
    # reset == load (simulation.py:119-120)
    sim.reset(meshes=meshes, params=[(1.0, 140, 0.2), (10.0, 1000, 0.4)])
    assert len(sim.displacements) == 1


def test_headless_paper_scenes_reproduce_the_reference_setup():
    """femflow_b200.simulation.mpm.headless.multi_drop_scene == paper_1.multi_drop_experiment (paper_1.py:17-144):
    the scene of experiment 0 point for point (golden: every 8th vertex of the reference's meshes) and the
    constructor arguments the reference passes."""
    from femflow_b200.simulation.mpm.headless import PointMesh, multi_drop_scene
    g = load_golden("c1_scene")
    meshes, params, ctor = multi_drop_scene(0)
    gy, co = (m.vertices.reshape(-1, 3) for m in meshes)
    assert gy.dtype == np.float32 and len(gy) == 8321 and len(co) == 27000
    assert np.array_equal(gy[::8], g["gyroid_vertices"]) and np.array_equal(co[::8], g["collider_vertices"])
    assert params == [(1.0, 140, 0.2), (10.0, 1000, 0.4)]
    assert (ctor["steps"], ctor["dt"], ctor["grid_res"], ctor["tightening_coeff"]) == (3500, float(g["dt"]), int(g["res"]),
                                                                                      float(g["tightening_coeff"]))
    assert (ctor["volume"], ctor["force"], ctor["hardening"]) == (float(g["volume"]), float(g["gravity"]), float(g["hardening"]))
    meshes1, _, ctor1 = multi_drop_scene(1)
    assert ctor1["steps"] == 1000 and ctor1["tightening_coeff"] == 0.10 and len(meshes1[1].vertices) == 3 * 40 ** 3
    with pytest.raises(ValueError):
        multi_drop_scene(2)
    m = PointMesh(np.arange(6.0).reshape(2, 3))
    m.translate_y(0.5); m.translate_x(-1); m.translate_z(2)
    assert np.array_equal(m.vertices, np.float32([-1, 1.5, 4, 2, 4.5, 7]))


@pytest.mark.gpu
def test_headless_runner_runs_experiment_0(tmp_path):
    """paper_1.multi_drop_experiment(0) (paper_1.py:75-114): the FULL 35 321-particle scene, 100 substeps through
    the headless runner (MPMSimulation thread, snapshot ring, .npy dump) against the C port of the reference loops
    started from the same float32-vertex positions (simulation.py:81-83)."""
    torch = pytest.importorskip("torch")
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from femflow_b200.simulation.mpm.headless import multi_drop_scene, run
    from femflow_b200.solvers.mpm.utils import Ev_to_lambda, Ev_to_mu
    from oracle import native as ON
    steps = 100
    sim, seconds = run(0, steps=steps, outdir=str(tmp_path / "out"))
    n = 35321
    assert sim.error is None and not sim.running and len(sim.displacements) == steps + 1
    assert sim.displacements[-1].shape == (3 * n,) and np.isfinite(sim.displacements[-1]).all()
    assert sorted(os.listdir(sim.outdir), key=lambda f: int(f.split(".")[0])) == [f"{i}.npy" for i in range(steps + 1)]
    # the port on the same scene
    meshes, params, ctor = multi_drop_scene(0)
    coeff = ctor["tightening_coeff"]
    x = np.concatenate([(m.vertices.reshape(-1, 3) * coeff).astype(np.float64) for m in meshes])
    counts = [len(m.vertices) // 3 for m in meshes]
    mass = np.concatenate([np.full(c, float(p[0])) for c, p in zip(counts, params)])
    mu = np.concatenate([np.full(c, Ev_to_mu(*p[1:])) for c, p in zip(counts, params)])
    lam = np.concatenate([np.full(c, Ev_to_lambda(*p[1:])) for c, p in zip(counts, params)])
    assert len(x) == n
    v = np.zeros((n, 3)); F = np.tile(np.eye(3), (n, 1, 1)); C = np.zeros((n, 3, 3))
    res = ctor["grid_res"]
    mid = None
    for k in range(steps):
        ON.solve_mls_mpm_3d(res, float(res), ctor["hardening"], 1.0 / res, ctor["dt"], ctor["volume"], ctor["force"],
                            x, mass, mu, lam, v, F, C)
        if k == 9:
            mid = (x / coeff).reshape(-1)
    V = max(np.abs(v).max(), ctor["dt"] * 9.8)
    assert rel_err(sim.displacements[10], mid) < 1e-5 * 10
    assert rel_err(sim.displacements[steps], (x / coeff).reshape(-1)) < 1e-5 * steps
    assert rel_err(sim.particles.pos, x, 1.0) < 1e-5 * steps
    assert rel_err(sim.v, v, V) < 1e-5 * steps
    assert rel_err(sim.F, F, 1.0) < 1e-5 * steps
    assert rel_err(sim.C, C, 4 * res * V) < 1e-5 * steps
