"""The multi-GPU host plumbing on CPU: the REAL ``CudaSlab`` (MpmSolver underneath, tests/fake_abi.py as the
library) driven by ``SlabDriver`` over gloo, world 2 and 3 -- payload packing, the storage-precision leaver
thresholds, migration in both directions, the lagged device-side leaver count, material rows that keep their
meaning across ranks, slab rebalancing with a re-created solver, and the one-sided halo transport.  The
round-end GPU box has one GPU, so these paths otherwise only ever run when someone has two."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from oracle import mpm_oracle as O  # noqa: E402
from test_distributed_cpu import SharedFabric, _free_port, make_lopsided_scene, make_scene  # noqa: E402


def _materials(state, nmat):
    x, v, F, C, mass, mu0, lam0, ids = state
    k = (ids % nmat).astype(np.float64)
    return (x, v, F, C, mass * (1 + 0.25 * k), mu0 * (1 + 0.5 * k), lam0 * (1 - 0.125 * k), ids)


def _worker(rank, world, port, out, cfg):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    import fake_abi
    fake_abi.install()
    from femflow_b200.distributed import CudaSlab, SlabDriver, SlabPlan
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        p, state = (make_lopsided_scene if cfg["lopsided"] else make_scene)()
        x, v, F, C, mass, mu0, lam0, ids = _materials(state, cfg["nmat"])
        plan = SlabPlan.make((p["res"],) * 3, world, rank, cfg["margin"])
        base, _ = O.base_and_fx(x, p["inv_dx"])
        mine = np.flatnonzero((base[:, 0] >= plan.own_lo) & (base[:, 0] < plan.own_hi))
        # lagged=False: a slab without the asynchronous-count interface takes the driver's synchronous migration path
        cls = CudaSlab if cfg["lagged"] else type("CudaSlabSync", (CudaSlab,), {"__getattribute__": _hide_count})
        local = cls(plan, p["dx"], p["dt"], p["volume"], p["gravity"], p["hardening"], capacity=len(x), device="cpu",
                    dtype=torch.float64, p2g_mode=cfg.get("p2g_mode", "auto"))
        local.set_particles(x[mine], v[mine], F[mine], C[mine], mass[mine], mu0[mine], lam0[mine], ids[mine])
        assert local.solver.material_layout == (f"table[{cfg['nmat']}]" if cfg["nmat"] <= 256 else "planes")
        fabric = SharedFabric(rank, world, *cfg["shared"]) if cfg.get("shared") else None
        drv = SlabDriver(plan, local, migrate_every=cfg["migrate_every"], halo="symm" if fabric else "p2p", fabric=fabric)
        assert hasattr(local, "count_leavers_async") == cfg["lagged"]
        drv.substep(cfg["pre"])
        rebalanced = drv.rebalance(layer_cost_per_cell=0.0) if cfg["rebalance"] else False
        drv.substep(cfg["steps"] - cfg["pre"])
        assert local.solver.poll_error() == 0
        got = [t.numpy() for t in local.state_by_id()]
        gathered = [None] * world
        dist.all_gather_object(gathered, (got, drv.migrated, rebalanced, local.num_particles))
        if rank == 0:
            idv = np.concatenate([g[0][0] for g in gathered])
            order = np.argsort(idv)
            res = {k: np.concatenate([g[0][i] for g in gathered])[order] for i, k in ((1, "x"), (2, "v"), (3, "F"), (4, "C"))}
            res.update(ids=idv[order], migrated=sum(g[1] for g in gathered), rebalanced=[g[2] for g in gathered],
                       counts=[g[3] for g in gathered])
            torch.save(res, out)
    finally:
        dist.destroy_process_group()


def _hide_count(self, name):
    if name in ("count_leavers_async", "stage_leaver_count", "read_leaver_count"):
        raise AttributeError(name)
    return object.__getattribute__(self, name)


CASES = {
    "sync-2": dict(world=2, margin=2, migrate_every=2, lagged=False, nmat=1),
    "lagged-3-table": dict(world=3, margin=2, migrate_every=1, lagged=True, nmat=3),
    "lagged-2-planes": dict(world=2, margin=3, migrate_every=2, lagged=True, nmat=300),
    "rebalance-3": dict(world=3, margin=2, migrate_every=1, lagged=True, nmat=3, lopsided=True, rebalance=True, pre=2),
    "symm-2": dict(world=2, margin=2, migrate_every=2, lagged=True, nmat=1, symm=True),
    # the default cadence: particles may stray `slack` cells into the halo margin, a round runs only when one has used it up
    "slack-2-margin4": dict(world=2, margin=4, migrate_every=None, lagged=True, nmat=3, steps=9),
    # the unbinned kernels never fill the device leaver counter: the count must come from the positions instead
    "lagged-2-scatter": dict(world=2, margin=2, migrate_every=1, lagged=True, nmat=3, p2g_mode="scatter"),
}


@pytest.mark.parametrize("case", sorted(CASES))
def test_cuda_slab_plumbing_matches_single_domain(tmp_path, case):
    cfg = dict(lopsided=False, rebalance=False, pre=0, steps=6, symm=False)
    cfg.update(CASES[case])
    world = cfg["world"]
    scene_fn = make_lopsided_scene if cfg["lopsided"] else make_scene
    p, state = scene_fn()
    if cfg["symm"]:
        from femflow_b200.distributed import SlabPlan
        G = p["res"] + 1
        planes = SlabPlan.min_cells(cfg["margin"])
        cfg["shared"] = ([torch.zeros((2, 2, planes, G, G, 4), dtype=torch.float64).share_memory_() for _ in range(world)],
                         [torch.zeros(4 * world, dtype=torch.int32).share_memory_() for _ in range(world)])
    out = str(tmp_path / "res.pt")
    mp.spawn(_worker, args=(world, _free_port(), out, cfg), nprocs=world, join=True)
    got = torch.load(out, weights_only=False)
    x, v, F, C, mass, mu0, lam0, ids = _materials(state, cfg["nmat"])
    Jp = np.ones((len(x), 1))
    for _ in range(cfg["steps"]):
        O.solve_mls_mpm_3d(p["res"], p["inv_dx"], p["hardening"], p["dx"], p["dt"], p["volume"], p["gravity"],
                           x, mass, mu0, lam0, v, F, C, Jp)
    assert np.array_equal(got["ids"], ids)                        # nobody lost, nobody duplicated
    assert got["migrated"] > 0
    if cfg["rebalance"]:
        assert all(got["rebalanced"]) and max(got["counts"]) < len(ids)
    for k, ref in (("x", x), ("v", v), ("F", F), ("C", C)):
        assert np.abs(got[k] - ref).max() <= 1e-11 * max(1.0, np.abs(ref).max()), k


def _dam_worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    import fake_abi
    fake_abi.install()
    import femflow_b200.distributed as D
    real_init = D.CudaSlab.__init__
    D.CudaSlab.__init__ = lambda self, *a, **k: real_init(self, *a, **{**k, "dtype": torch.float64})
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        sol = D.SlabSolver.from_dam_break(rank, world, "cpu", res=16, n_total=4000, margin=2)
        start = [t.numpy().copy() for t in sol.local.state_by_id()]
        counts0 = sol.num_particles
        sol.substep(3)
        did = sol.rebalance(layer_cost_per_cell=0.0)
        sol.substep(4)
        assert sol.poll_error() == 0
        end = [t.numpy() for t in sol.local.state_by_id()]
        cfg = sol.local.solver.cfg
        gathered = [None] * world
        dist.all_gather_object(gathered, (start, end, counts0, sol.num_particles, did, list(sol.plan.all_ranges),
                                          dict(dt=cfg.dt, volume=cfg.volume, gravity=cfg.gravity, hardening=cfg.hardening,
                                               mass=float(sol.local.solver.material_table[0, 0]),
                                               mu=float(sol.local.solver.material_table[0, 1]),
                                               lam=float(sol.local.solver.material_table[0, 2]))))
        if rank == 0:
            torch.save(gathered, out)
    finally:
        dist.destroy_process_group()


def test_dam_break_setup_with_empty_ranks_and_recut(tmp_path):
    """SlabSolver.from_dam_break (BASELINE configs[4]) at toy size on three ranks: the column sits on rank 0, the
    other ranks start EMPTY (and must still take part in the material-table collective), the re-cut spreads it,
    and the union of the slabs follows the single-domain oracle on the union of the initial particles."""
    world = 3
    out = str(tmp_path / "dam.pt")
    mp.spawn(_dam_worker, args=(world, _free_port(), out), nprocs=world, join=True)
    g = torch.load(out, weights_only=False)
    counts0, counts1 = [r[2] for r in g], [r[3] for r in g]
    assert counts0[0] > 0 and counts0[2] == 0 and sum(counts0) == sum(counts1)
    # the column is 8 cell layers wide and a slab at least 2*margin+2 = 6 thick: two ranks can share it, not three
    assert all(r[4] for r in g) and sorted(counts1)[1] > 0 and max(counts1) < 0.6 * sum(counts1)
    assert g[0][5] == g[1][5] == g[2][5] and g[0][5][0][0] == 0 and g[0][5][-1][1] == 16 * world - 1
    prm = g[0][6]
    cat = lambda which, i: np.concatenate([r[which][i] for r in g])
    ids0, order0 = cat(0, 0), np.argsort(cat(0, 0))
    x, v, F, C = (cat(0, i)[order0].astype(np.float64) for i in (1, 2, 3, 4))
    n = len(x)
    assert len(np.unique(ids0)) == n
    # single-domain oracle on the (48, 16, 16) box: embed in the cube the oracle expects; y / z walls restated as in the stand-in
    import fake_abi
    res3, G = [48, 16, 16], 49
    mass, mu, lam = np.full(n, prm["mass"]), np.full(n, prm["mu"]), np.full(n, prm["lam"])
    for _ in range(7):
        gv, gm = np.zeros((G, G, G, 3)), np.zeros((G, G, G, 1))
        O.p2g_3d(16.0, prm["hardening"], 1 / 16, prm["dt"], prm["volume"], gv, gm, x, mass, mu, lam, v, F, C, np.ones((n, 1)))
        fake_abi.FakeLib.box_grid_op(res3, 1 / 16, prm["dt"], prm["gravity"], gv, gm)
        O.g2p_3d(16.0, prm["dt"], gv, x, v, F, C, np.ones((n, 1)))
    order1 = np.argsort(cat(1, 0))
    assert np.array_equal(cat(1, 0)[order1], ids0[order0])
    for i, ref in ((1, x), (2, v), (3, F), (4, C)):
        assert np.abs(cat(1, i)[order1] - ref).max() <= 1e-10 * max(1.0, np.abs(ref).max()), i
