"""Size-independent properties of the substep, held by BOTH oracles (NumPy restatement and C port).  They follow from
the reference's formulae alone (quadratic B-splines are a partition of unity with a zero first moment about the
particle, three_d/p2g.py:55,72-80), so they check the checker without a golden file, and they are the properties the
GPU tests use at the headline sizes (tests/test_gpu_parity_at_size.py: sum of grid mass = sum of particle mass)."""
import numpy as np
import pytest

from oracle import mpm_oracle as O
from oracle import native as ON


def _scene(rng, n, res, dim):
    x = rng.uniform(0.25, 0.75, (n, dim))
    v = rng.normal(0, 0.3, (n, dim))
    F = np.eye(dim) + rng.normal(0, 0.03, (n, dim, dim))
    C = rng.normal(0, 2.0, (n, dim, dim))
    mass = rng.uniform(0.5, 2.0, n)
    mu = rng.uniform(50, 100, n)
    lam = rng.uniform(20, 60, n)
    return x, v, F, C, mass, mu, lam


@pytest.mark.parametrize("impl", ["numpy", "c"])
@pytest.mark.parametrize("seed", [0, 1, 2])
def test_p2g_3d_conserves_mass_and_momentum(impl, seed):
    """sum_i w_ip = 1 and sum_i w_ip (x_i - x_p) = 0: the grid holds exactly the particles' mass and momentum, whatever
    the affine matrix (stress + mass C) is."""
    rng = np.random.default_rng(seed)
    res, n = 32, 4000
    x, v, F, C, mass, mu, lam = _scene(rng, n, res, 3)
    G = res + 1
    gv = np.zeros((G, G, G, 3)); gm = np.zeros((G, G, G, 1))
    p2g = O.p2g_3d if impl == "numpy" else ON.p2g_3d
    p2g(float(res), 0.7, 1.0 / res, 1e-4, 1.0 / res ** 3, gv, gm, x, mass, mu, lam, v, F, C, np.ones((n, 1)))
    assert abs(gm.sum() - mass.sum()) < 1e-11 * mass.sum()
    want = (mass[:, None] * v).sum(0)
    assert np.abs(gv.reshape(-1, 3).sum(0) - want).max() < 1e-10 * np.abs(mass[:, None] * v).sum()


@pytest.mark.parametrize("impl", ["numpy", "c"])
def test_uniform_grid_velocity_is_gathered_exactly(impl):
    """G2P of a constant grid velocity u: v_p = u, C_p = 0 (zero first moment), F unchanged, x advected by dt u."""
    rng = np.random.default_rng(3)
    res, n = 16, 500
    x, v, F, C, *_ = _scene(rng, n, res, 3)
    u = np.array([0.3, -0.2, 0.1])
    G = res + 1
    gv = np.broadcast_to(u, (G, G, G, 3)).copy()
    x0, F0 = x.copy(), F.copy()
    g2p = O.g2p_3d if impl == "numpy" else ON.g2p_3d
    g2p(float(res), 1e-3, gv, x, v, F, C, np.ones((n, 1)))
    assert np.abs(v - u).max() < 1e-14
    assert np.abs(C).max() < 1e-11
    assert np.abs(F - F0).max() < 1e-13
    assert np.abs(x - (x0 + 1e-3 * u)).max() < 1e-15


@pytest.mark.parametrize("impl", ["numpy", "c"])
def test_linear_grid_velocity_gives_its_gradient(impl):
    """APIC with quadratic B-splines reproduces an affine field: gv(X) = u + A X  =>  v_p = u + A x_p, C_p = A
    (three_d/g2p.py:37-43: 4 inv_dx sum w gv (x_i - x_p)^T / dx with D = dx^2 / 4)."""
    rng = np.random.default_rng(4)
    res, n = 16, 300
    x, v, F, C, *_ = _scene(rng, n, res, 3)
    A = rng.normal(0, 1.0, (3, 3)); u = rng.normal(0, 1.0, 3)
    G = res + 1
    I = np.stack(np.meshgrid(*[np.arange(G)] * 3, indexing="ij"), -1) / res
    gv = u + I @ A.T
    x0 = x.copy()
    g2p = O.g2p_3d if impl == "numpy" else ON.g2p_3d
    g2p(float(res), 1e-4, gv, x, v, F, C, np.ones((n, 1)))
    assert np.abs(v - (u + x0 @ A.T)).max() < 1e-12
    assert np.abs(C - A).max() < 1e-10


def test_p2g_is_additive_over_particle_sets():
    """The scatter is a sum over particles: grids of two disjoint particle sets add up to the grid of their union
    (what the slab decomposition's halo SUM relies on, distributed.py)."""
    rng = np.random.default_rng(5)
    res, n = 24, 1500
    x, v, F, C, mass, mu, lam = _scene(rng, n, res, 3)
    G = res + 1

    def run(sel):
        gv = np.zeros((G, G, G, 3)); gm = np.zeros((G, G, G, 1))
        O.p2g_3d(float(res), 1.0, 1.0 / res, 1e-4, 1.0 / res ** 3, gv, gm, x[sel], mass[sel], mu[sel], lam[sel], v[sel], F[sel],
                 C[sel], np.ones((int(np.sum(sel)), 1)))
        return gv, gm
    a = rng.random(n) < 0.4
    gva, gma = run(a); gvb, gmb = run(~a); gvu, gmu = run(np.ones(n, bool))
    assert np.abs(gma + gmb - gmu).max() < 1e-13 * gmu.max()
    assert np.abs(gva + gvb - gvu).max() < 1e-12 * np.abs(gvu).max()


@pytest.mark.parametrize("seed", [0, 1])
def test_2d_conservation_and_walls(seed):
    rng = np.random.default_rng(seed)
    res, n = 64, 3000
    x, v, F, C, *_ = _scene(rng, n, res, 2)
    G = res + 1
    gv = np.zeros((G, G, 2)); gm = np.zeros((G, G, 1))
    O.p2g_2d(float(res), 1.0, 80.0, 40.0, 1.5, 1.0 / res, 1e-4, 1.0 / res ** 2, gv, gm, x, v, F, C, np.ones((n, 1)))
    assert abs(gm.sum() - 1.5 * n) < 1e-10 * n
    assert np.abs(gv.reshape(-1, 2).sum(0) - 1.5 * v.sum(0)).max() < 1e-9 * np.abs(v).sum()
    # two_d/grid_op.py:13-24: sticky bands at x < 0.05, x > 0.95, y > 0.95; the floor only stops downward motion
    gv2 = np.ones((G, G, 2)) * np.array([1.0, -1.0]); gm2 = np.ones((G, G, 1))
    O.grid_op_2d(res, 0.0, 0.0, gv2, gm2)
    i = np.arange(G) / res
    sticky = (i < 0.05) | (i > 1 - 0.05)
    assert np.all(gv2[sticky] == 0) and np.all(gv2[:, i > 1 - 0.05] == 0)
    floor = (i < 0.05)
    inner = ~sticky
    assert np.all(gv2[np.ix_(inner, floor)][..., 1] == 0) and np.all(gv2[np.ix_(inner, floor)][..., 0] == 1)
