"""Two-GPU slab decomposition (NCCL, CudaSlab) against the single-domain NumPy oracle.
Skipped on boxes with fewer than two GPUs."""
import os
import socket
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def _scene(nmat):
    """The CPU test's scene, stored values rounded to fp32; nmat distinct (mass, mu0, lam0) triples."""
    from test_distributed_cpu import make_scene
    p, state = make_scene(res=32, n=20000, seed=9)
    x, v, F, C, mass, mu0, lam0, ids = state
    k = (ids % nmat).astype(np.float64)
    mass, mu0, lam0 = mass * (1 + 0.25 * k), mu0 * (1 + 0.5 * k), lam0 * (1 - 0.125 * k)
    state = (x, v, F, C, mass, mu0, lam0, ids)
    return p, tuple(np.asarray(a, dtype=np.float32).astype(np.float64) if a.dtype == np.float64 else a for a in state)


def _worker(rank, world, port, steps, out, nmat, halo="p2p"):
    import torch.distributed as dist
    from femflow_b200.distributed import CudaSlab, SlabDriver, SlabPlan
    from oracle import mpm_oracle as O
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    try:
        p, state = _scene(nmat)
        x, v, F, C, mass, mu0, lam0, ids = state
        plan = SlabPlan.make((p["res"],) * 3, world, rank, margin=2)
        base, _ = O.base_and_fx(x, p["inv_dx"])
        mine = np.flatnonzero((base[:, 0] >= plan.own_lo) & (base[:, 0] < plan.own_hi))
        local = CudaSlab(plan, p["dx"], p["dt"], p["volume"], p["gravity"], p["hardening"], capacity=len(x), device=dev)
        local.set_particles(x[mine], v[mine], F[mine], C[mine], mass[mine], mu0[mine], lam0[mine], ids[mine])
        assert local.solver.material_layout == f"table[{nmat}]"
        drv = SlabDriver(plan, local, migrate_every=2, halo=halo)
        drv.substep(steps)
        assert local.solver.poll_error() == 0
        got = [t.cpu().numpy() for t in local.state_by_id()]
        gathered = [None] * world
        dist.all_gather_object(gathered, (got, drv.migrated))
        if rank == 0:
            idv = np.concatenate([g[0][0] for g in gathered])
            order = np.argsort(idv)
            res = {k: np.concatenate([g[0][i] for g in gathered])[order] for i, k in ((1, "x"), (2, "v"), (3, "F"), (4, "C"))}
            res["ids"] = idv[order]
            res["migrated"] = sum(g[1] for g in gathered)
            torch.save(res, out)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("nmat", [1, 3])
def test_two_gpu_slabs_match_oracle(tmp_path, nmat):
    if not torch.cuda.is_available() or torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    import torch.multiprocessing as mp
    from oracle import mpm_oracle as O
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    steps = 6
    out = str(tmp_path / "res.pt")
    mp.spawn(_worker, args=(2, port, steps, out, nmat), nprocs=2, join=True)
    got = torch.load(out, weights_only=False)
    p, (x, v, F, C, mass, mu0, lam0, ids) = _scene(nmat)
    Jp = np.ones((len(x), 1))
    for _ in range(steps):
        O.solve_mls_mpm_3d(p["res"], p["inv_dx"], p["hardening"], p["dx"], p["dt"], p["volume"], p["gravity"],
                           x, mass, mu0, lam0, v, F, C, Jp)
    assert np.array_equal(got["ids"], ids)
    assert got["migrated"] > 0
    V = max(np.abs(v).max(), p["dt"] * 9.8)
    tol = 1e-5 * steps
    assert np.abs(got["x"] - x).max() < tol
    assert np.abs(got["v"] - v).max() / V < tol
    assert np.abs(got["F"] - F).max() < tol
    assert np.abs(got["C"] - C).max() / (4 * p["inv_dx"] * V) < tol


def test_two_gpu_symm_halo_matches_oracle(tmp_path):
    """One-sided halo puts over NVLink peer memory (SymmHalo on torch symmetric memory; green on 2 B200s:
    profiles/r02c_pytest_gpu_2gpus.txt)."""
    if not torch.cuda.is_available() or torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    import torch.multiprocessing as mp
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    steps = 7
    out = str(tmp_path / "res.pt")
    mp.spawn(_worker, args=(2, port, steps, out, 3, "symm"), nprocs=2, join=True)
    got = torch.load(out, weights_only=False)
    p, state = _scene(3)
    assert np.array_equal(got["ids"], state[7]) and got["migrated"] > 0
    _assert_close(p, got, _oracle_run(p, state, steps), steps)


@pytest.mark.parametrize("nmat,dtype", [(1, "float32"), (3, "float32"), (300, "float32"), (3, "float64")])
def test_device_migration_pack_unpack_one_gpu(nmat, dtype):
    """csrc/mpm_migrate.cuh through ffmpm_migrate_pack / _unpack on ONE GPU: the leavers of an owned range land in the
    two outboxes (payload rows = their state, material and id), the holes are back-filled so that the live buffer holds
    exactly the keepers, a small outbox leaves the overflow in place, and feeding the outboxes back in (a loop-back
    exchange) restores the full particle set -- every plane, by particle id."""
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    import ctypes as C
    from femflow_b200 import _native as N
    from femflow_b200.mpm import MpmSolver
    p, (x, v, F, Cm, mass, mu0, lam0, ids) = _scene(nmat)
    n, res = len(x), p["res"]
    tdt = getattr(torch, dtype)
    s = MpmSolver(3, res, p["dt"], p["volume"], p["gravity"], p["hardening"], capacity=2 * n, dtype=tdt, reorder=True)
    s.set_particles(x, v, F, Cm, None, mass, mu0, lam0)
    s.substep(1)                                    # a sorted live buffer with binning state to invalidate
    before = {k: t.clone() for k, t in s.get_particles().items()}
    live = s.live
    base = torch.trunc(live.x[0, :n].double() * p["inv_dx"] - 0.5).long().cpu().numpy()
    own_lo, own_hi = 12, 19
    s.set_owned_range(own_lo, own_hi)
    want_lo, want_hi = int((base < own_lo).sum()), int((base >= own_hi).sum())
    assert want_lo > 100 and want_hi > 100
    rows = int(s.lib.ffmpm_migrate_rows(s._h))
    assert rows == 28
    rec = torch.zeros(8, dtype=torch.int32).pin_memory()

    def box(cap):
        return torch.zeros((rows + 1, cap), dtype=tdt, device="cuda")

    # ---- a round whose outboxes are too small: the overflow stays, nothing is lost ----
    cap = 64
    out_lo, out_hi = box(cap), box(cap)
    N.check(s.lib.ffmpm_migrate_pack(s._h, out_lo.data_ptr(), out_hi.data_ptr(), cap, s._stream()))
    N.check(s.lib.ffmpm_migrate_unpack(s._h, None, None, cap, rec.data_ptr(), s._stream()))
    torch.cuda.synchronize()
    o_lo, o_hi, i_lo, i_hi, n_new, overflow = (int(t) for t in rec[:6])
    assert (o_lo, o_hi, i_lo, i_hi) == (cap, cap, 0, 0) and n_new == n - 2 * cap
    assert overflow == want_lo + want_hi - 2 * cap
    N.check(s.lib.ffmpm_set_num_particles(s._h, n_new)); s.num_particles = n_new
    kept_ids = s.live.id[:n_new].cpu().numpy()
    sent_ids = np.concatenate([out_lo[1 + 27, :cap].view(torch.int32).cpu().numpy() if dtype == "float32" else out_lo[28, :cap].long().cpu().numpy(),
                               out_hi[1 + 27, :cap].view(torch.int32).cpu().numpy() if dtype == "float32" else out_hi[28, :cap].long().cpu().numpy()])
    assert np.array_equal(np.sort(np.concatenate([kept_ids, sent_ids])), np.arange(n))
    # loop back: what left to the left comes back "from the right" and vice versa
    N.check(s.lib.ffmpm_migrate_pack(s._h, None, None, cap, s._stream()))         # no neighbours: nothing leaves
    N.check(s.lib.ffmpm_migrate_unpack(s._h, out_hi.data_ptr(), out_lo.data_ptr(), cap, rec.data_ptr(), s._stream()))
    torch.cuda.synchronize()
    assert [int(t) for t in rec[:6]] == [0, 0, cap, cap, n, 0]
    N.check(s.lib.ffmpm_set_num_particles(s._h, n)); s.num_particles = n

    # ---- a full round ----
    cap = 8192
    out_lo, out_hi = box(cap), box(cap)
    N.check(s.lib.ffmpm_migrate_pack(s._h, out_lo.data_ptr(), out_hi.data_ptr(), cap, s._stream()))
    N.check(s.lib.ffmpm_migrate_unpack(s._h, None, None, cap, rec.data_ptr(), s._stream()))
    torch.cuda.synchronize()
    o_lo, o_hi, i_lo, i_hi, n_new, overflow = (int(t) for t in rec[:6])
    assert (o_lo, o_hi, overflow) == (want_lo, want_hi, 0) and n_new == n - want_lo - want_hi
    assert int(out_lo[0, 0]) == want_lo and int(out_hi[0, 0]) == want_hi
    N.check(s.lib.ffmpm_set_num_particles(s._h, n_new)); s.num_particles = n_new
    kb = torch.trunc(s.live.x[0, :n_new].double() * p["inv_dx"] - 0.5).long()
    assert bool(((kb >= own_lo) & (kb < own_hi)).all())                 # only keepers are left, no holes
    xb = torch.trunc(out_lo[1, :want_lo].double() * p["inv_dx"] - 0.5).long()
    assert bool((xb < own_lo).all())
    # and back again: the state by id is bit-identical to what it was
    N.check(s.lib.ffmpm_migrate_pack(s._h, None, None, cap, s._stream()))
    N.check(s.lib.ffmpm_migrate_unpack(s._h, out_lo.data_ptr(), out_hi.data_ptr(), cap, rec.data_ptr(), s._stream()))
    torch.cuda.synchronize()
    assert [int(t) for t in rec[:6]] == [0, 0, want_lo, want_hi, n, 0]
    N.check(s.lib.ffmpm_set_num_particles(s._h, n)); s.num_particles = n
    after = s.get_particles()
    for k in before:
        assert torch.equal(before[k], after[k]), k
    # the material travelled with its particle: one more substep against the oracle
    s.substep(1)
    s.check_errors()
    got = {k: t.double().cpu().numpy() for k, t in s.get_particles().items()}
    ref = _oracle_run(p, (x, v, F, Cm, mass, mu0, lam0, ids), 2)
    _assert_close(p, got, ref, 2) if dtype == "float32" else None
    s.close()


def _lopsided(nmat):
    """_scene() squeezed into the low-x 40 % of the domain: the even cut leaves rank 1 idle."""
    p, (x, v, F, C, mass, mu0, lam0, ids) = _scene(nmat)
    x = x.copy()
    x[:, 0] = 0.12 + (x[:, 0] - 0.15) * (0.28 / 0.70)
    x = x.astype(np.float32).astype(np.float64)
    return p, (x, v, F, C, mass, mu0, lam0, ids)


def _oracle_run(p, state, steps):
    from oracle import mpm_oracle as O
    x, v, F, C, mass, mu0, lam0, ids = (a.copy() for a in state)
    Jp = np.ones((len(x), 1))
    for _ in range(steps):
        O.solve_mls_mpm_3d(p["res"], p["inv_dx"], p["hardening"], p["dx"], p["dt"], p["volume"], p["gravity"],
                           x, mass, mu0, lam0, v, F, C, Jp)
    return x, v, F, C


def _assert_close(p, got, ref, steps):
    x, v, F, C = ref
    V = max(np.abs(v).max(), p["dt"] * 9.8)
    tol = 1e-5 * steps
    assert np.abs(got["x"] - x).max() < tol
    assert np.abs(got["v"] - v).max() / V < tol
    assert np.abs(got["F"] - F).max() < tol
    assert np.abs(got["C"] - C).max() / (4 * p["inv_dx"] * V) < tol


@pytest.mark.parametrize("nmat", [1, 3])
def test_cuda_slab_rebuild_one_gpu(nmat):
    """CudaSlab.take_all / rebuild / append (the local half of SlabDriver.rebalance) on one GPU: rank 0
    of a two-slab cut is moved from cells [0, 16) to [0, 21) between substeps.  All particles live on
    rank 0 (the absent rank 1 would contribute zero to the shared planes), so no halo exchange is needed
    and the single-domain oracle is the reference."""
    if not torch.cuda.is_available():
        pytest.skip("needs a GPU")
    from femflow_b200.distributed import CudaSlab, SlabPlan
    p, state = _lopsided(nmat)
    x, v, F, C, mass, mu0, lam0, ids = state
    assert x[:, 0].max() * p["inv_dx"] < 13.5           # base cell <= 12 now, <= 14 after 4 substeps: inside both local grids
    res = (p["res"],) * 3
    plan_a = SlabPlan.make(res, 2, 0, margin=2, ranges=[(0, 16), (16, 31)])
    plan_b = SlabPlan.make(res, 2, 0, margin=2, ranges=[(0, 21), (21, 31)])
    dev = torch.device("cuda", 0)
    local = CudaSlab(plan_a, p["dx"], p["dt"], p["volume"], p["gravity"], p["hardening"], capacity=len(x), device=dev)
    local.set_particles(x, v, F, C, mass, mu0, lam0, ids)
    layout = local.solver.material_layout

    def substeps(k):
        for _ in range(k):
            local.scatter()
            local.grid_update(None, 0, None, 0)
            local.gather()

    substeps(2)
    hist = local.layer_histogram(p["res"] - 1)
    assert int(hist.sum()) == len(x) and int(hist[14:].sum()) == 0
    launches = local.solver.launch_count()
    payload, base_x = local.take_all()
    assert local.num_particles == 0 and payload[0].shape == (27, len(x)) and int(base_x.max()) <= 13
    local.rebuild(plan_b, len(x))
    assert local.solver.n[0] == plan_b.n_local_x == 21 + 2 + 2 and local.solver.material_layout == layout
    assert local.launches_carried == launches
    local.append(payload)
    assert local.num_particles == len(x)
    substeps(2)
    assert local.solver.poll_error() == 0
    idv, gx, gv, gF, gC = (t.cpu().numpy() for t in local.state_by_id())
    order = np.argsort(idv)
    assert np.array_equal(idv[order], ids)
    got = {"x": gx[order], "v": gv[order], "F": gF[order], "C": gC[order]}
    _assert_close(p, got, _oracle_run(p, state, 4), 4)


def _rebalance_worker(rank, world, port, out, nmat):
    import torch.distributed as dist
    from femflow_b200.distributed import CudaSlab, SlabDriver, SlabPlan
    from oracle import mpm_oracle as O
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    try:
        p, state = _lopsided(nmat)
        x, v, F, C, mass, mu0, lam0, ids = state
        plan = SlabPlan.make((p["res"],) * 3, world, rank, margin=2)
        base, _ = O.base_and_fx(x, p["inv_dx"])
        mine = np.flatnonzero((base[:, 0] >= plan.own_lo) & (base[:, 0] < plan.own_hi))
        local = CudaSlab(plan, p["dx"], p["dt"], p["volume"], p["gravity"], p["hardening"], capacity=len(x), device=dev)
        local.set_particles(x[mine], v[mine], F[mine], C[mine], mass[mine], mu0[mine], lam0[mine], ids[mine])
        drv = SlabDriver(plan, local, migrate_every=1)
        drv.substep(2)
        _, before = drv.imbalance(0.0)
        assert drv.rebalance(layer_cost_per_cell=0.0)
        _, after = drv.imbalance(0.0)
        count = local.num_particles
        drv.substep(4)
        assert local.solver.poll_error() == 0
        got = [t.cpu().numpy() for t in local.state_by_id()]
        gathered = [None] * world
        dist.all_gather_object(gathered, (got, count, before, after))
        if rank == 0:
            idv = np.concatenate([g[0][0] for g in gathered])
            order = np.argsort(idv)
            res = {k: np.concatenate([g[0][i] for g in gathered])[order] for i, k in ((1, "x"), (2, "v"), (3, "F"), (4, "C"))}
            res.update(ids=idv[order], counts=[g[1] for g in gathered], before=before, after=after)
            torch.save(res, out)
    finally:
        dist.destroy_process_group()


def test_two_gpu_rebalance_matches_oracle(tmp_path):
    if not torch.cuda.is_available() or torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    import torch.multiprocessing as mp
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    out = str(tmp_path / "res.pt")
    mp.spawn(_rebalance_worker, args=(2, port, out, 3), nprocs=2, join=True)
    got = torch.load(out, weights_only=False)
    p, state = _lopsided(3)
    assert np.array_equal(got["ids"], state[7])
    assert got["after"] < got["before"] and min(got["counts"]) > 0      # rank 1 started empty
    _assert_close(p, got, _oracle_run(p, state, 6), 6)
