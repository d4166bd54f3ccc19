"""GPU parity AT THE HEADLINE SIZES (BASELINE configs[1] and [2]) against the C port of the
reference loops (oracle/mpm_oracle.c, pinned to the reference's goldens in test_oracle_golden.py).

One full substep of
  * C3: 3D, 16 777 216 particles, 256^3 grid  (three_d/p2g.py:14-80, grid_op.py:5-47, g2p.py:9-59)
  * C2: 2D,  1 048 576 particles, 1024^2 grid (two_d/p2g.py:11-76, grid_op.py:5-24, g2p.py:5-47)
through the production kernels (what ``ffmpm_substep`` / ``bench.py`` run), checked phase by phase:
cell keys bit-exact; grid mass and momentum after P2G, grid velocity after the grid update,
particle x, v, C, F (Jp in 2D) after G2P within 1e-5 max-norm relative with the floors of SURVEY 8d.
A second substep then runs from the reordered (cell-sorted) buffers -- the state the timed region
of bench.py is in.  The port needs ~2 s per 3D substep on 16 threads and ~6 GB of host memory.
"""
import numpy as np
import pytest

from conftest import rel_err
from oracle import mpm_oracle as O

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")

TOL = 1e-5


@pytest.fixture(scope="module", autouse=True)
def _need_cuda():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")


def _floors(dt, gravity, inv_dx, mass, vel_ref):
    V = max(float(np.abs(vel_ref).max()), dt * abs(gravity))
    return dict(vel=V, mom=float(np.max(mass)) * V, C=4 * inv_dx * V)


def _tile_keys(base, n_nodes, dim):
    if dim == 3:
        tiles = [(n - 2 + 3) // 4 for n in n_nodes]
        t = ((base[:, 0] >> 2) * tiles[1] + (base[:, 1] >> 2)) * tiles[2] + (base[:, 2] >> 2)
        return t * 64 + ((base[:, 0] & 3) << 4) + ((base[:, 1] & 3) << 2) + (base[:, 2] & 3)
    tiles = [(n - 2 + 7) // 8 for n in n_nodes]
    t = (base[:, 0] >> 3) * tiles[1] + (base[:, 1] >> 3)
    return t * 64 + ((base[:, 0] & 7) << 3) + (base[:, 1] & 7)


def _chunked_max_abs_diff(dev_tensor, ref, chunk=1 << 22):
    """max |dev - ref| without materialising a second fp64 copy of a 16.7 M-row array on the host."""
    worst = 0.0
    flat = ref.reshape(len(ref), -1)
    for a in range(0, len(ref), chunk):
        got = dev_tensor[a:a + chunk].double().cpu().numpy().reshape(-1, flat.shape[1])
        worst = max(worst, float(np.abs(got - flat[a:a + chunk]).max()))
    return worst


def test_c3_16m_substep_matches_port():
    from femflow_b200 import scenes
    from femflow_b200.mpm import MpmSolver
    from oracle import native as ON
    sc = scenes.config_3d_16m(seed=0)
    n, res = sc.n, sc.res
    assert n == 16_777_216 and res == 256
    G = res + 1
    x, v, F, C = (a.astype(np.float64) for a in (sc.x, sc.v, sc.F, sc.C))
    mass = np.full(n, sc.mass); mu = np.full(n, sc.mu_0); lam = np.full(n, sc.lambda_0)
    dx = 1.0 / res

    s = MpmSolver(3, res, sc.dt, sc.volume, sc.gravity, sc.hardening, capacity=n, mass=sc.mass, mu_0=sc.mu_0,
                  lambda_0=sc.lambda_0)                       # p2g_mode auto, reorder: the production pipeline
    s.set_particles(sc.x, sc.v, sc.F, sc.C, None, sc.mass, sc.mu_0, sc.lambda_0)

    # ---- binning: keys bit-exact against the fp64 indexing of three_d/p2g.py:50 ----
    s.clear_grid(); s.bin()
    keys, perm, off, n_cells = s.bin_results()
    base, _ = O.base_and_fx(x, float(res))
    want = _tile_keys(base, [G] * 3, 3).astype(np.int32)
    assert np.array_equal(keys.cpu().numpy(), want)
    sk = keys[perm.long()]
    assert bool((sk[1:] >= sk[:-1]).all())                    # perm is a counting-sort order
    del keys, perm, off, sk, base, want

    # ---- P2G (production kernel: p2g_bulk3 over the cp.async-prefetched windows) ----
    s.p2g()
    gv = np.zeros(G * G * G * 3); gm = np.zeros(G * G * G)
    ON.p2g_3d(float(res), sc.hardening, dx, sc.dt, sc.volume, gv.reshape(G, G, G, 3), gm.reshape(G, G, G, 1),
              x, mass, mu, lam, v, F, C, np.ones((n, 1)))
    g = s.grid(readonly=True)
    got_m = g[..., 3].double().cpu().numpy().reshape(-1)
    assert np.abs(got_m - gm).max() / np.abs(gm).max() < TOL
    # conservation (size-independent property): sum of grid mass == sum of particle mass
    assert abs(got_m.sum() - n * sc.mass) / (n * sc.mass) < 1e-6
    got_p = g[..., :3].double().cpu().numpy().reshape(-1)
    mom_ref = gv.copy()
    # grid update of the reference on ITS momentum gives the velocity scale for the floors
    ON.grid_op_3d(res, dx, sc.dt, sc.gravity, gv, gm)
    fl = _floors(sc.dt, sc.gravity, float(res), mass, gv)
    assert np.abs(got_p - mom_ref).max() / max(np.abs(mom_ref).max(), fl["mom"]) < TOL
    del got_m, got_p, mom_ref

    # ---- grid update (production: block-list kernel, no grid access in between) ----
    s.grid_op()
    got_v = s.grid(readonly=True)[..., :3].double().cpu().numpy().reshape(-1)
    assert np.abs(got_v - gv).max() / fl["vel"] < TOL
    del got_v

    # ---- G2P (production: tiled, reordering, pre-binning the next substep) ----
    s.g2p()
    assert s.poll_error() == 0
    ON.g2p_3d(float(res), sc.dt, gv.reshape(G, G, G, 3), x, v, F, C)
    out = s.get_particles()
    assert _chunked_max_abs_diff(out["x"], x) < TOL
    assert _chunked_max_abs_diff(out["v"], v) / fl["vel"] < TOL
    assert _chunked_max_abs_diff(out["F"], F) < TOL
    assert _chunked_max_abs_diff(out["C"], C) / fl["C"] < TOL
    del out

    # ---- second substep: from the cell-sorted buffers, ffmpm_substep as bench.py calls it ----
    s.substep(1)
    assert s.poll_error() == 0
    ON.solve_mls_mpm_3d(res, float(res), sc.hardening, dx, sc.dt, sc.volume, sc.gravity, x, mass, mu, lam, v, F, C,
                        scratch=(gv, gm))
    out = s.get_particles()
    assert _chunked_max_abs_diff(out["x"], x) < 2 * TOL
    assert _chunked_max_abs_diff(out["v"], v) / fl["vel"] < 2 * TOL
    assert _chunked_max_abs_diff(out["F"], F) < 2 * TOL
    assert _chunked_max_abs_diff(out["C"], C) / fl["C"] < 2 * TOL
    s.close()


def test_c2_1m_substep_matches_port():
    from femflow_b200 import scenes
    from femflow_b200.mpm import MpmSolver
    from oracle import native as ON
    sc = scenes.config_2d_1m(seed=0)
    n, res = sc.n, sc.res
    assert n == 1_048_576 and res == 1024
    G = res + 1
    x, v, F, C = (a.astype(np.float64) for a in (sc.x, sc.v, sc.F, sc.C))
    Jp = np.ones((n, 1))
    dx = 1.0 / res
    s = MpmSolver(2, res, sc.dt, sc.volume, sc.gravity, sc.hardening, capacity=n, mass=sc.mass, mu_0=sc.mu_0,
                  lambda_0=sc.lambda_0)                       # the 2D default: thread-per-particle kernels (what bench.py times)
    s.set_particles(sc.x, sc.v, sc.F, sc.C)

    s.clear_grid(); s.bin()
    keys, perm, off, n_cells = s.bin_results()
    base, _ = O.base_and_fx(x, float(res))
    assert np.array_equal(keys.cpu().numpy(), _tile_keys(base, [G, G], 2).astype(np.int32))

    s.p2g()
    gv = np.zeros((G, G, 2)); gm = np.zeros((G, G, 1))
    O.p2g_2d(float(res), sc.hardening, sc.mu_0, sc.lambda_0, sc.mass, dx, sc.dt, sc.volume, gv, gm, x, v, F, C, Jp)
    g = s.grid(readonly=True).double().cpu().numpy()[:, :, 0]
    vel = gv.copy()
    O.grid_op_2d(res, sc.dt, sc.gravity, vel, gm)
    V = max(float(np.abs(vel).max()), sc.dt * abs(sc.gravity))
    assert rel_err(g[..., 2:3], gm) < TOL
    assert rel_err(g[..., :2], gv, sc.mass * V) < TOL
    s.grid_op()
    g = s.grid(readonly=True).double().cpu().numpy()[:, :, 0]
    assert rel_err(g[..., :2], vel, V) < TOL
    s.g2p()
    assert s.poll_error() == 0
    O.g2p_2d(float(res), sc.dt, vel, x, v, F, C, Jp)
    out = {k: t.double().cpu().numpy() for k, t in s.get_particles().items()}
    assert rel_err(out["x"], x, 1.0) < TOL
    assert rel_err(out["v"], v, V) < TOL
    assert rel_err(out["F"], F, 1.0) < TOL
    assert rel_err(out["C"], C, 4 * res * V) < TOL
    assert rel_err(out["Jp"], Jp, 1.0) < TOL
    # two more substeps through ffmpm_substep (the path bench.py --workload 2d1m times)
    s.substep(2)
    assert s.poll_error() == 0
    for _ in range(2):
        ON.solve_mls_mpm_2d(res, float(res), sc.hardening, sc.mu_0, sc.lambda_0, sc.mass, dx, sc.dt, sc.volume,
                            sc.gravity, x, v, F, C, Jp)
    out = {k: t.double().cpu().numpy() for k, t in s.get_particles().items()}
    assert rel_err(out["x"], x, 1.0) < 3 * TOL
    assert rel_err(out["v"], v, V) < 3 * TOL
    assert rel_err(out["F"], F, 1.0) < 3 * TOL
    assert rel_err(out["C"], C, 4 * res * V) < 3 * TOL
    s.close()
