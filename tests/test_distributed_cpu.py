"""Slab decomposition protocol (femflow_b200.distributed) on CPU: world_size 2 and 3
over gloo, with a NumPy (oracle) stand-in for the rank-local solver.  Halo sums plus
particle migration must reproduce the single-domain oracle run."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from femflow_b200.distributed import LocalSlab, SlabDriver, SlabPlan  # noqa: E402
from oracle import mpm_oracle as O  # noqa: E402


def test_plan_geometry():
    # 64 cells + 1: valid base cells 0..62 -> 63 cells over 3 ranks
    plans = [SlabPlan.make((64, 64, 64), 3, r, margin=2) for r in range(3)]
    assert [(p.own_lo, p.own_hi) for p in plans] == [(0, 21), (21, 42), (42, 63)]
    assert plans[0].g_lo == 0 and plans[2].g_hi == 65
    for a, b in zip(plans[:-1], plans[1:]):
        assert a.planes_hi == b.planes_lo == 2 * 2 + 2          # 2*margin + 2 shared node planes
        assert a.g_hi - a.planes_hi == b.g_lo
    one = SlabPlan.make((64, 64, 64), 1, 0)
    assert (one.g_lo, one.g_hi, one.planes_lo, one.planes_hi) == (0, 65, 0, 0)
    with pytest.raises(ValueError):
        SlabPlan.make((16, 16, 16), 4, 0, margin=2)             # slabs thinner than the halo


class OracleSlab(LocalSlab):
    """Rank-local solver made of the NumPy oracle on the full (small) global grid; only
    the planes of the local range are exchanged / trusted."""

    def __init__(self, plan, p, state):
        self.plan, self.p = plan, p
        self.device, self.dtype = torch.device("cpu"), torch.float64
        self.x, self.v, self.F, self.C, self.mass, self.mu0, self.lam0, self.ids = state
        G = plan.res[0] + 1
        self.grid = torch.zeros((G, G, G, 4), dtype=torch.float64)

    @property
    def num_particles(self):
        return len(self.x)

    def scatter(self):
        G = self.plan.res[0] + 1
        p = self.p
        gv = np.zeros((G, G, G, 3)); gm = np.zeros((G, G, G, 1))
        O.p2g_3d(p["inv_dx"], p["hardening"], p["dx"], p["dt"], p["volume"], gv, gm, self.x, self.mass, self.mu0,
                 self.lam0, self.v, self.F, self.C, np.ones((len(self.x), 1)))
        self.grid[..., :3] = torch.from_numpy(gv)
        self.grid[..., 3:] = torch.from_numpy(gm)

    def grid_planes(self, a, b):
        return self.grid[self.plan.g_lo + a:self.plan.g_lo + b]

    def grid_update(self, recv_lo, planes_lo, recv_hi, planes_hi):
        pl, p = self.plan, self.p
        if planes_lo:
            self.grid[pl.g_lo:pl.g_lo + planes_lo] += recv_lo
        if planes_hi:
            self.grid[pl.g_hi - planes_hi:pl.g_hi] += recv_hi
        gv = self.grid[..., :3].numpy().copy(); gm = self.grid[..., 3:].numpy().copy()
        O.grid_op_3d(pl.res[0], p["dx"], p["dt"], p["gravity"], gv, gm)
        self.gv = gv

    def gather(self):
        p = self.p
        O.g2p_3d(p["inv_dx"], p["dt"], self.gv, self.x, self.v, self.F, self.C, np.ones((len(self.x), 1)))

    def payload_rows(self):
        return 27

    def _pack(self, idx):
        n = len(idx)
        rows = np.concatenate([self.x[idx].T, self.v[idx].T, self.C[idx].reshape(n, 9).T, self.F[idx].reshape(n, 9).T,
                               self.mass[idx][None], self.mu0[idx][None], self.lam0[idx][None]], 0)
        return torch.from_numpy(np.ascontiguousarray(rows)), torch.from_numpy(self.ids[idx].astype(np.int32))

    def extract_leavers(self, own_lo, own_hi):
        base, _ = O.base_and_fx(self.x, self.p["inv_dx"])
        bx = base[:, 0]
        li, ri = np.flatnonzero(bx < own_lo), np.flatnonzero(bx >= own_hi)
        left, right = self._pack(li), self._pack(ri)
        keep = np.flatnonzero((bx >= own_lo) & (bx < own_hi))
        for name in ("x", "v", "F", "C", "mass", "mu0", "lam0", "ids"):
            setattr(self, name, getattr(self, name)[keep])
        return left, right

    # slab rebalancing
    def layer_histogram(self, cells):
        base, _ = O.base_and_fx(self.x, self.p["inv_dx"])
        return torch.from_numpy(np.bincount(np.clip(base[:, 0], 0, cells - 1), minlength=cells).astype(np.int64))

    def take_all(self):
        base, _ = O.base_and_fx(self.x, self.p["inv_dx"])
        payload = self._pack(np.arange(len(self.x)))
        for name in ("x", "v", "F", "C", "mass", "mu0", "lam0", "ids"):
            setattr(self, name, getattr(self, name)[:0])
        return payload, torch.from_numpy(base[:, 0].astype(np.int64))

    def rebuild(self, plan, n_particles):
        self.plan = plan          # the stand-in computes on the whole global grid; only the exchanged planes move
        self.rebuilt_for = n_particles

    # optional asynchronous-count interface (exercises SlabDriver's lagged migration decision)
    lagged = False

    def __getattribute__(self, name):
        if name in ("count_leavers_async", "stage_leaver_count", "read_leaver_count") and not object.__getattribute__(self, "lagged"):
            raise AttributeError(name)
        return object.__getattribute__(self, name)

    def count_leavers_async(self, own_lo, own_hi):
        base, _ = O.base_and_fx(self.x, self.p["inv_dx"])
        return torch.tensor([int(np.count_nonzero((base[:, 0] < own_lo) | (base[:, 0] >= own_hi)))], dtype=torch.int64)

    def stage_leaver_count(self, cnt):
        return cnt.clone()

    def read_leaver_count(self, handle):
        return int(handle[0])

    def append(self, payload):
        data, ids = payload[0].numpy(), payload[1].numpy()
        n = data.shape[1]
        self.x = np.concatenate([self.x, data[0:3].T]); self.v = np.concatenate([self.v, data[3:6].T])
        self.C = np.concatenate([self.C, data[6:15].T.reshape(n, 3, 3)])
        self.F = np.concatenate([self.F, data[15:24].T.reshape(n, 3, 3)])
        self.mass = np.concatenate([self.mass, data[24]]); self.mu0 = np.concatenate([self.mu0, data[25]])
        self.lam0 = np.concatenate([self.lam0, data[26]]); self.ids = np.concatenate([self.ids, ids.astype(np.int64)])


def make_scene(res=24, n=1500, seed=3):
    """Particles spread over the whole x range with velocities of ~0.4 cells/substep both ways."""
    rng = np.random.default_rng(seed)
    dx = 1.0 / res
    p = dict(res=res, inv_dx=float(res), dx=dx, dt=1e-3, volume=(dx / 2) ** 3, gravity=-9.8, hardening=1.0)
    x = rng.uniform(0.15, 0.85, size=(n, 3))
    v = rng.normal(0, 0.3, size=(n, 3))
    v[:, 0] += np.where(rng.random(n) < 0.5, 1, -1) * 0.4 * dx / p["dt"]
    F = np.eye(3) + rng.normal(0, 0.01, size=(n, 3, 3))
    C = rng.normal(0, 0.2, size=(n, 3, 3))
    mass = np.full(n, p["volume"]); mu0 = np.full(n, 40.0); lam0 = np.full(n, 30.0)
    return p, (x, v, F, C, mass, mu0, lam0, np.arange(n, dtype=np.int64))


class SharedFabric:
    """SymmHalo fabric on CPU: the inboxes and signal pads of all ranks are shared-memory tensors made
    by the parent process, so a rank really writes into its neighbour's inbox and raises flags in its
    neighbour's pad -- the semantics of torch's symmetric memory (put_signal: wait for 0, set 1;
    wait_signal: wait for 1, set 0; one flag per (channel, source rank))."""

    def __init__(self, rank, world, inboxes, pads):
        self.rank, self.world, self.inboxes, self.pads = rank, world, inboxes, pads
        self.puts = self.waits = 0

    def empty(self, shape, dtype, device):
        t = self.inboxes[self.rank]
        assert tuple(t.shape) == tuple(shape) and t.dtype == dtype
        return t

    def rendezvous(self, tensor, group):
        return self

    def get_buffer(self, rank, sizes, dtype, storage_offset=0):
        return self.inboxes[rank]

    def barrier(self, channel=0, timeout_ms=0):
        dist.barrier()

    def _spin(self, flag, index, want, timeout_ms):
        import time
        t0 = time.time()
        while int(flag[index]) != want:
            if time.time() - t0 > timeout_ms / 1000:
                raise TimeoutError(f"rank {self.rank}: flag {index} never became {want}")
            time.sleep(0)

    def put_signal(self, dst_rank, channel=0, timeout_ms=0):
        i = channel * self.world + self.rank
        self._spin(self.pads[dst_rank], i, 0, timeout_ms)
        self.pads[dst_rank][i] = 1
        self.puts += 1

    def wait_signal(self, src_rank, channel=0, timeout_ms=0):
        i = channel * self.world + src_rank
        self._spin(self.pads[self.rank], i, 1, timeout_ms)
        self.pads[self.rank][i] = 0
        self.waits += 1


def _worker(rank, world, port, steps, margin, migrate_every, out, lagged=False, shared=None, symm_sync="signal"):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), FFMPM_SYMM_SYNC=symm_sync)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        p, state = make_scene()
        plan = SlabPlan.make((p["res"],) * 3, world, rank, margin)
        base, _ = O.base_and_fx(state[0], p["inv_dx"])
        mine = np.flatnonzero((base[:, 0] >= plan.own_lo) & (base[:, 0] < plan.own_hi))
        local = OracleSlab(plan, p, tuple(a[mine].copy() for a in state))
        local.lagged = lagged
        fabric = SharedFabric(rank, world, *shared) if shared is not None else None
        drv = SlabDriver(plan, local, migrate_every=migrate_every, halo="symm" if fabric else "p2p", fabric=fabric)
        drv.substep(steps)
        if fabric is not None:
            n_nb = (rank > 0) + (rank < world - 1)
            # one flag per neighbour per substep, all consumed (none at all under the barrier protocol)
            assert fabric.puts == fabric.waits == (steps * n_nb if symm_sync == "signal" else 0)
        gathered = [None] * world
        dist.all_gather_object(gathered, (local.ids, local.x, local.v, local.F, local.C, drv.migrated))
        if rank == 0:
            ids = np.concatenate([g[0] for g in gathered])
            order = np.argsort(ids)
            res = {k: np.concatenate([g[i] for g in gathered])[order] for i, k in ((1, "x"), (2, "v"), (3, "F"), (4, "C"))}
            res["ids"] = ids[order]
            res["migrated"] = sum(g[5] for g in gathered)
            torch.save(res, out)
    finally:
        dist.destroy_process_group()


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


@pytest.mark.parametrize("world,margin,migrate_every,lagged", [(2, 2, 2, False), (3, 1, 1, False), (2, 3, 2, False),
                                                                  (2, 3, 2, True), (3, 2, 1, True)])
def test_slabs_match_single_domain(tmp_path, world, margin, migrate_every, lagged):
    steps = 7
    out = str(tmp_path / "res.pt")
    mp.spawn(_worker, args=(world, _free_port(), steps, margin, migrate_every, out, lagged), nprocs=world, join=True)
    got = torch.load(out, weights_only=False)
    p, (x, v, F, C, mass, mu0, lam0, ids) = make_scene()
    Jp = np.ones((len(x), 1))
    for _ in range(steps):
        O.solve_mls_mpm_3d(p["res"], p["inv_dx"], p["hardening"], p["dx"], p["dt"], p["volume"], p["gravity"],
                           x, mass, mu0, lam0, v, F, C, Jp)
    assert np.array_equal(got["ids"], ids)                     # nobody lost, nobody duplicated
    assert got["migrated"] > 0                                 # the scene really exercises migration
    for k, ref in (("x", x), ("v", v), ("F", F), ("C", C)):
        assert np.abs(got[k] - ref).max() <= 1e-12 * max(1.0, np.abs(ref).max()), k


@pytest.mark.parametrize("world,margin,migrate_every,sync", [(2, 2, 2, "signal"), (3, 1, 1, "signal"), (3, 2, 2, "barrier")])
def test_symm_halo_protocol_matches_single_domain(tmp_path, world, margin, migrate_every, sync):
    """SlabDriver(halo="symm"): halo planes put into the neighbour's inbox + signal flags instead of
    matched send/recv (SymmHalo), over a shared-memory stand-in for torch's symmetric memory.  Same
    bar as the p2p transport: the single-domain oracle to 1e-12, after an odd number of substeps so
    that both inbox parities are used an unequal number of times."""
    steps = 7
    p, _ = make_scene()
    G = p["res"] + 1
    planes = SlabPlan.min_cells(margin)
    inboxes = [torch.zeros((2, 2, planes, G, G, 4), dtype=torch.float64).share_memory_() for _ in range(world)]
    pads = [torch.zeros(4 * world, dtype=torch.int32).share_memory_() for _ in range(world)]
    out = str(tmp_path / "res.pt")
    mp.spawn(_worker, args=(world, _free_port(), steps, margin, migrate_every, out, False, (inboxes, pads), sync),
             nprocs=world, join=True)
    assert all(int(pad.abs().sum()) == 0 for pad in pads)           # every raised flag was consumed
    got = torch.load(out, weights_only=False)
    p, (x, v, F, C, mass, mu0, lam0, ids) = make_scene()
    Jp = np.ones((len(x), 1))
    for _ in range(steps):
        O.solve_mls_mpm_3d(p["res"], p["inv_dx"], p["hardening"], p["dx"], p["dt"], p["volume"], p["gravity"],
                           x, mass, mu0, lam0, v, F, C, Jp)
    assert np.array_equal(got["ids"], ids) and got["migrated"] > 0
    for k, ref in (("x", x), ("v", v), ("F", F), ("C", C)):
        assert np.abs(got[k] - ref).max() <= 1e-12 * max(1.0, np.abs(ref).max()), k


def test_balanced_ranges():
    counts = np.zeros(63, dtype=np.int64)
    counts[5:15] = 1000                                         # a column near the low end
    r = SlabPlan.balanced_ranges(counts, 4, min_cells=6)
    assert r[0][0] == 0 and r[-1][1] == 63 and all(a[1] == b[0] for a, b in zip(r[:-1], r[1:]))
    assert all(hi - lo >= 6 for lo, hi in r)
    loads = [counts[lo:hi].sum() for lo, hi in r]
    assert max(loads) == 5000            # optimum under the width constraint (10 column layers, slabs >= 6); even cut: 10000
    # uniform counts reproduce (about) the even split, and the layer cost spreads empty space
    u = SlabPlan.balanced_ranges(np.full(63, 10), 3, 6)
    assert [hi - lo for lo, hi in u] == [21, 21, 21]
    e = SlabPlan.balanced_ranges(np.zeros(63), 3, 6, layer_cost=1.0)
    assert [hi - lo for lo, hi in e] == [21, 21, 21]
    with pytest.raises(ValueError):
        SlabPlan.balanced_ranges(np.zeros(15), 3, 6)
    # explicit ranges round-trip through make(); ranges that do not tile are refused
    plan = SlabPlan.make((64, 64, 64), 4, 1, margin=2, ranges=r)
    assert (plan.own_lo, plan.own_hi) == r[1] and plan.all_ranges == tuple(r)
    assert plan.planes_lo == plan.planes_hi == 6
    with pytest.raises(ValueError):
        SlabPlan.make((64, 64, 64), 2, 0, margin=2, ranges=[(0, 30), (31, 63)])


def make_lopsided_scene(res=32, n=1800, seed=5):
    """All particles in the low-x third of the domain, drifting towards +x: the even cut leaves
    the upper ranks idle."""
    p, (x, v, F, C, mass, mu0, lam0, ids) = make_scene(res=res, n=n, seed=seed)
    x[:, 0] = 0.12 + (x[:, 0] - 0.15) * (0.30 / 0.70)
    v[:, 0] = np.abs(v[:, 0])
    return p, (x, v, F, C, mass, mu0, lam0, ids)


def _rebalance_worker(rank, world, port, margin, lagged, out, migrate_every=1, pre_steps=2):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        p, state = make_lopsided_scene()
        plan = SlabPlan.make((p["res"],) * 3, world, rank, margin)
        base, _ = O.base_and_fx(state[0], p["inv_dx"])
        mine = np.flatnonzero((base[:, 0] >= plan.own_lo) & (base[:, 0] < plan.own_hi))
        local = OracleSlab(plan, p, tuple(a[mine].copy() for a in state))
        local.lagged = lagged
        auto = pre_steps == 0                                     # the driver re-cuts by itself every 2 substeps
        drv = SlabDriver(plan, local, migrate_every=migrate_every, rebalance_every=2 if auto else 0)
        _, before = drv.imbalance(0.0)
        drv.substep(pre_steps if not auto else 2)
        if auto:
            assert drv.rebalanced == 1
            pre_steps = 2
        else:
            assert drv.rebalance(layer_cost_per_cell=0.0)
        _, after = drv.imbalance(0.0)
        assert not drv.rebalance(layer_cost_per_cell=0.0)         # a second call has nothing to gain
        counts_after = local.num_particles
        drv.substep(5 - pre_steps)
        gathered = [None] * world
        dist.all_gather_object(gathered, (local.ids, local.x, local.v, local.F, local.C, counts_after,
                                          (drv.plan.own_lo, drv.plan.own_hi)))
        if rank == 0:
            ids = np.concatenate([g[0] for g in gathered])
            order = np.argsort(ids)
            res = {k: np.concatenate([g[i] for g in gathered])[order] for i, k in ((1, "x"), (2, "v"), (3, "F"), (4, "C"))}
            res.update(ids=ids[order], counts=[g[5] for g in gathered], ranges=[g[6] for g in gathered],
                       before=before, after=after)
            torch.save(res, out)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,margin,lagged,migrate_every,pre_steps",
                         [(3, 2, False, 1, 2), (2, 3, True, 1, 2), (4, 1, False, 1, 2),
                          (2, 3, False, 3, 2),      # re-cut between migrations: particles that strayed into the margin go along
                          (3, 2, False, 2, 3),
                          (3, 2, True, 1, 0)])      # SlabDriver(rebalance_every=2)
def test_rebalanced_slabs_match_single_domain(tmp_path, world, margin, lagged, migrate_every, pre_steps):
    out = str(tmp_path / "res.pt")
    mp.spawn(_rebalance_worker, args=(world, _free_port(), margin, lagged, out, migrate_every, pre_steps), nprocs=world,
             join=True)
    got = torch.load(out, weights_only=False)
    p, (x, v, F, C, mass, mu0, lam0, ids) = make_lopsided_scene()
    Jp = np.ones((len(x), 1))
    for _ in range(5):
        O.solve_mls_mpm_3d(p["res"], p["inv_dx"], p["hardening"], p["dx"], p["dt"], p["volume"], p["gravity"],
                           x, mass, mu0, lam0, v, F, C, Jp)
    assert np.array_equal(got["ids"], ids)
    assert got["after"] < got["before"]                          # the most loaded rank got lighter ...
    assert max(got["counts"]) < len(ids)                         # ... and no longer holds everything
    r = got["ranges"]
    assert r[0][0] == 0 and r[-1][1] == p["res"] - 1 and all(a[1] == b[0] for a, b in zip(r[:-1], r[1:]))
    for k, ref in (("x", x), ("v", v), ("F", F), ("C", C)):
        assert np.abs(got[k] - ref).max() <= 1e-12 * max(1.0, np.abs(ref).max()), k


def test_balanced_ranges_is_optimal_on_small_cases():
    """SlabPlan.balanced_ranges against brute force: over all cuts that respect the minimum slab thickness,
    none has a lighter heaviest slab."""
    import itertools
    rng = np.random.default_rng(0)
    for trial in range(200):
        world = int(rng.integers(2, 5))
        min_cells = int(rng.integers(1, 4))
        cells = int(rng.integers(world * min_cells, world * min_cells + 9))
        counts = rng.integers(0, 50, cells) * (rng.random(cells) < 0.7)
        layer_cost = float(rng.choice([0.0, 0.5, 3.0]))
        got = SlabPlan.balanced_ranges(counts, world, min_cells, layer_cost)
        assert got[0][0] == 0 and got[-1][1] == cells and all(a[1] == b[0] for a, b in zip(got[:-1], got[1:]))
        assert all(hi - lo >= min_cells for lo, hi in got)
        cost = counts + layer_cost

        def worst(cuts):
            edges = (0,) + cuts + (cells,)
            return max(cost[a:b].sum() for a, b in zip(edges[:-1], edges[1:]))
        best = min(worst(c) for c in itertools.combinations(range(1, cells), world - 1)
                   if all(b - a >= min_cells for a, b in zip((0,) + c, c + (cells,))))
        mine = max(cost[lo:hi].sum() for lo, hi in got)
        assert mine <= best + 1e-9 * max(1.0, best), (trial, got, mine, best)


@pytest.mark.parametrize("dtype", [torch.float32, torch.float64])
def test_leaver_threshold_is_the_exact_cell_boundary(dtype):
    """CudaSlab._threshold(cell): the smallest storage-precision position whose base cell is >= cell, so that
    `x < t` / `x >= t` reproduce `base < cell` / `base >= cell` of the kernels' fp64 indexing exactly (any
    resolution; base 0 also owns the (-1, 0) band that truncation sends there, quirk 1)."""
    from femflow_b200.distributed import CudaSlab
    np_dt = np.float32 if dtype == torch.float32 else np.float64
    for res in (37, 64, 80, 256, 1000):
        slab = CudaSlab.__new__(CudaSlab)             # the threshold needs no device: only dtype and inv_dx
        slab.dtype, slab.inv_dx = dtype, float(res)
        assert slab._threshold(0) == float("-inf")
        for cell in (1, 2, 7, res // 3, res // 2, res - 2):
            t = np_dt(slab._threshold(cell))
            below = np.nextafter(t, np_dt(-np.inf))
            base = lambda v: int(np.float64(v) * float(res) - 0.5)
            assert base(t) >= cell and base(below) < cell, (res, cell)
            # and agrees with the oracle's indexing on a cloud of positions around the boundary
            x = (t + np.arange(-50, 50) * np.spacing(t)).astype(np_dt)
            b, _ = O.base_and_fx(x.astype(np.float64)[:, None], float(res))
            assert np.array_equal(x >= t, b[:, 0] >= cell)
