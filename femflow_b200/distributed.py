"""Slab decomposition of the MPM substep across the GPUs of one box (new: the
reference is a single serial process).

The global grid is cut into slabs of node planes along axis 0 (the slowest axis of
the C-order grid, so a node plane is one contiguous block).  A particle belongs to
the rank that owns its base cell ``base.x``.  Per substep each rank

  1. scatters its particles into its local grid (bin + P2G),
  2. exchanges the node planes it shares with its two neighbours and SUMS them
     (the sum is fused into the grid update's load, ``ffmpm_grid_op_halo``),
  3. runs the grid update on every local plane (the shared planes are computed
     redundantly and bit-identically on both sides: a + b == b + a),
  4. gathers (G2P),

and every ``migrate_every`` substeps hands the particles whose base cell left its
range to the +-1 neighbour.  The local grid carries ``margin`` extra cell layers on
each side so that a particle may stray that far between migrations (the reference's
CFL clamp, three_d/grid_op.py:25,34-36, bounds the motion to < 1 cell per substep).

``SlabPlan`` and ``SlabDriver`` are pure host logic over an abstract local solver and
``torch.distributed`` point-to-point ops, so the exchange protocol is unit-tested on
CPU (gloo) with a NumPy stand-in; on GPUs the local solver is ``CudaSlab`` (NCCL).
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import List, Optional, Sequence, Tuple

import numpy as np
import torch
import torch.distributed as dist


# --------------------------------------------------------------------------- #
@dataclass
class SlabPlan:
    """Ownership and halo geometry of one rank."""
    res: Tuple[int, int, int]
    world: int
    rank: int
    margin: int
    own_lo: int      # owned base cells [own_lo, own_hi) along x
    own_hi: int
    g_lo: int        # local node planes [g_lo, g_hi) (global indices)
    g_hi: int
    planes_lo: int   # planes shared with rank-1 (the first planes of the local grid)
    planes_hi: int   # planes shared with rank+1 (the last planes of the local grid)
    all_ranges: Tuple[Tuple[int, int], ...] = ()   # owned cell range of every rank (the cut this plan belongs to)

    @property
    def n_local_x(self) -> int:
        return self.g_hi - self.g_lo

    @staticmethod
    def ranges(res_x: int, world: int) -> List[Tuple[int, int]]:
        """Even split of the valid base cells 0 .. res_x-2 (utils.py:138-150)."""
        cells = res_x - 1
        base, rem = divmod(cells, world)
        out, lo = [], 0
        for r in range(world):
            hi = lo + base + (1 if r < rem else 0)
            out.append((lo, hi))
            lo = hi
        return out

    @staticmethod
    def min_cells(margin: int) -> int:
        """Thinnest slab whose halo planes stay inside the neighbouring slab."""
        return 2 * margin + 2

    @staticmethod
    def balanced_ranges(counts, world: int, min_cells: int, layer_cost: float = 0.0) -> List[Tuple[int, int]]:
        """Cut the base-cell layers 0 .. len(counts)-1 into ``world`` contiguous ranges of about equal
        cost, where a layer costs ``counts[layer] + layer_cost`` (its particles plus a constant for the
        cells the binning scans whether or not they hold particles), each at least ``min_cells`` thick,
        minimising the cost of the heaviest range.  Pure function of its arguments: every rank derives the same cut from the all-reduced counts."""
        cost = np.asarray(counts, dtype=np.float64) + float(layer_cost)
        cells = len(cost)
        if cells < world * min_cells:
            raise ValueError("slabs are thinner than 2*margin+2 cells: halos would reach past the neighbour")
        prefix = np.concatenate([[0.0], np.cumsum(cost)])      # prefix[c] = cost of layers [0, c)
        # Dynamic programme over (slabs used, layers covered): heaviest[r][c] = the lightest possible heaviest
        # slab when the first r+1 slabs cover layers [0, c).  (A greedy "fill each slab up to a bound" is NOT
        # optimal here: with a minimum thickness a fuller slab can force a later one across two heavy layers.)
        inf = np.inf
        heaviest = np.full((world, cells + 1), inf)
        prev_cut = np.zeros((world, cells + 1), dtype=np.int64)
        heaviest[0, min_cells:] = prefix[min_cells:]
        for r in range(1, world):
            for c in range((r + 1) * min_cells, cells + 1):
                starts = np.arange(r * min_cells, c - min_cells + 1)          # where slab r may begin
                worst = np.maximum(heaviest[r - 1, starts], prefix[c] - prefix[starts])
                k = int(np.argmin(worst))
                heaviest[r, c], prev_cut[r, c] = worst[k], starts[k]
        best = [cells]
        for r in range(world - 1, 0, -1):
            best.append(int(prev_cut[r, best[-1]]))
        best.append(0)
        best.reverse()
        return [(best[r], best[r + 1]) for r in range(world)]

    @classmethod
    def make(cls, res: Sequence[int], world: int, rank: int, margin: int = 2,
             ranges: Optional[Sequence[Tuple[int, int]]] = None) -> "SlabPlan":
        """``ranges``: owned base-cell range of every rank (contiguous, covering 0 .. res_x-2);
        default: the even split."""
        res = tuple(int(r) for r in res)
        rng = [(int(a), int(b)) for a, b in ranges] if ranges is not None else cls.ranges(res[0], world)
        if len(rng) != world or rng[0][0] != 0 or rng[-1][1] != res[0] - 1 or \
                any(a[1] != b[0] for a, b in zip(rng[:-1], rng[1:])):
            raise ValueError(f"ranges {rng} do not tile the base cells 0 .. {res[0] - 2} over {world} ranks")
        G = res[0] + 1

        def nodes(r):
            lo, hi = rng[r]
            return max(0, lo - margin), min(G, hi + margin + 2)
        for lo, hi in rng:
            if world > 1 and hi - lo < cls.min_cells(margin):
                raise ValueError("slabs are thinner than 2*margin+2 cells: halos would reach past the neighbour")
        g_lo, g_hi = nodes(rank)
        planes_lo = nodes(rank - 1)[1] - g_lo if rank > 0 else 0
        planes_hi = g_hi - nodes(rank + 1)[0] if rank < world - 1 else 0
        return cls(res, world, rank, margin, rng[rank][0], rng[rank][1], g_lo, g_hi, planes_lo, planes_hi, tuple(rng))


# --------------------------------------------------------------------------- #
class LocalSlab:
    """What SlabDriver needs from a rank-local solver (CudaSlab here, a NumPy
    stand-in in tests/)."""
    device: torch.device
    dtype: torch.dtype
    num_particles: int

    def scatter(self) -> None:                      # clear grid, (bin,) P2G
        raise NotImplementedError

    def grid_planes(self, a: int, b: int) -> torch.Tensor:   # contiguous view of local planes [a, b)
        raise NotImplementedError

    def grid_update(self, recv_lo, planes_lo, recv_hi, planes_hi) -> None:
        raise NotImplementedError

    def gather(self) -> None:                       # G2P
        raise NotImplementedError

    def extract_leavers(self, own_lo: int, own_hi: int):     # -> (left, right) payloads; keeps the rest
        raise NotImplementedError

    def append(self, payload) -> None:
        raise NotImplementedError

    def payload_rows(self) -> int:
        raise NotImplementedError

    # -- slab rebalancing (optional: SlabDriver.rebalance) -----------------------
    def layer_histogram(self, cells: int) -> torch.Tensor:   # (cells,) int64: local particles per global base-cell layer
        raise NotImplementedError

    def take_all(self):                                      # -> (payload, base_x int64 (n,)); leaves the slab empty
        raise NotImplementedError

    def rebuild(self, plan: "SlabPlan", n_particles: int) -> None:   # empty local grid for ``plan``, room for n_particles
        raise NotImplementedError


# --------------------------------------------------------------------------- #
class SymmHalo:
    """Halo planes pushed straight into the neighbour's memory over NVLink (one-sided put into a
    symmetric-memory inbox + a signal), instead of a matched NCCL send/recv pair.

    Every rank owns an inbox ``[parity][side][planes, ny, nz, 4]`` allocated from
    ``torch.distributed._symmetric_memory`` (CUDA IPC / NVLink peer mapping).  Per substep ``k``
    (parity ``q = k & 1``) a rank

      * copies its last shared planes into ``inbox[q][LO]`` OF THE RIGHT NEIGHBOUR and its first
        shared planes into ``inbox[q][HI]`` OF THE LEFT NEIGHBOUR (plain peer stores),
      * raises one flag per neighbour (``put_signal``, release at system scope, stream-ordered
        after the copy), then waits for the two flags raised for it (``wait_signal``),
      * hands ``inbox[q]`` to the grid update, which sums it while loading (``ffmpm_grid_op_halo``).

    Everything is enqueued on the current stream; the host never blocks.  Two parities make reuse
    safe without a second handshake: a neighbour writes ``inbox[q]`` for substep k+2 only after it
    has seen this rank's flag of substep k+1, which this rank raises after its scatter of k+1, i.e.
    (same stream) after its grid update of k consumed ``inbox[q]``.

    ``fabric`` abstracts the allocator / rendezvous so that the protocol is unit-tested on CPU with a
    shared-memory stand-in (tests/test_distributed_cpu.py); on GPUs it is torch's symmetric memory.
    """
    LO, HI = 0, 1

    def __init__(self, plan: SlabPlan, plane_shape: Sequence[int], dtype, device, group=None, fabric=None,
                 timeout_ms: int = 20000, sync: Optional[str] = None):
        """``sync``: ``"signal"`` (default) = one flag per neighbour and substep; ``"barrier"`` = the fabric's
        stream-ordered barrier over all ranks instead (coarser, but the primitive torch's own symmetric-memory
        collectives rely on) -- selectable with FFMPM_SYMM_SYNC for bring-up on new hardware."""
        import os
        self.sync = sync or os.environ.get("FFMPM_SYMM_SYNC", "signal")
        if self.sync not in ("signal", "barrier"):
            raise ValueError(f"unknown SymmHalo sync {self.sync!r}")
        self.rank, self.world = plan.rank, plan.world
        self.left = plan.rank - 1 if plan.rank > 0 else None
        self.right = plan.rank + 1 if plan.rank < plan.world - 1 else None
        self.planes = SlabPlan.min_cells(plan.margin)          # 2*margin + 2 on every interior cut
        self.timeout_ms = int(timeout_ms)
        shape = (2, 2, self.planes) + tuple(int(v) for v in plane_shape)
        fabric = fabric if fabric is not None else _TorchSymmFabric()
        self.inbox = fabric.empty(shape, dtype, device)
        self.inbox.zero_()
        self.handle = fabric.rendezvous(self.inbox, group)
        self.peer = {nb: self.handle.get_buffer(nb, shape, dtype, 0) for nb in (self.left, self.right) if nb is not None}
        self.handle.barrier(0, self.timeout_ms)               # every inbox is zeroed and mapped before the first put

    def exchange(self, step: int, send_lo: Optional[torch.Tensor], send_hi: Optional[torch.Tensor]):
        """``send_lo`` / ``send_hi``: this rank's first / last shared planes (None at a domain end).
        Returns the planes received from the left / right neighbour (None at a domain end)."""
        q, h, t = step & 1, self.handle, self.timeout_ms
        flags = self.sync == "signal"
        if self.right is not None:
            self.peer[self.right][q, self.LO].copy_(send_hi)
            if flags:
                h.put_signal(self.right, 2 * q + self.LO, t)
        if self.left is not None:
            self.peer[self.left][q, self.HI].copy_(send_lo)
            if flags:
                h.put_signal(self.left, 2 * q + self.HI, t)
        if not flags:
            h.barrier(q, t)                      # every rank's puts of this substep have landed
        if flags and self.left is not None:
            h.wait_signal(self.left, 2 * q + self.LO, t)
        if flags and self.right is not None:
            h.wait_signal(self.right, 2 * q + self.HI, t)
        return (self.inbox[q, self.LO] if self.left is not None else None,
                self.inbox[q, self.HI] if self.right is not None else None)


class _TorchSymmFabric:
    """torch.distributed._symmetric_memory as the SymmHalo fabric (GPUs of one NVLink domain)."""

    def empty(self, shape, dtype, device):
        import torch.distributed._symmetric_memory as symm_mem
        return symm_mem.empty(*shape, dtype=dtype, device=device)

    def rendezvous(self, tensor, group):
        import torch.distributed._symmetric_memory as symm_mem
        return symm_mem.rendezvous(tensor, group if group is not None else dist.group.WORLD)


class SlabDriver:
    def __init__(self, plan: SlabPlan, local: LocalSlab, group=None, migrate_every: Optional[int] = None,
                 halo: str = "p2p", fabric=None, rebalance_every: int = 0):
        """``halo``: ``"p2p"`` = grouped send/recv of the shared planes (NCCL on GPUs, gloo in the CPU
        tests); ``"symm"`` = one-sided puts into the neighbour's symmetric-memory inbox (``SymmHalo``).
        ``rebalance_every``: offer a re-cut of the slabs (``rebalance``) every that many substeps; 0 = never."""
        self.plan, self.local, self.group = plan, local, group
        self.rebalance_every = int(rebalance_every)
        if halo not in ("p2p", "symm"):
            raise ValueError(f"unknown halo transport {halo!r}")
        self.migrate_every = migrate_every if migrate_every is not None else max(1, plan.margin)
        if self.migrate_every > max(1, plan.margin) and plan.world > 1:
            raise ValueError("migrate_every must not exceed the halo margin (particles move < 1 cell per substep)")
        self.steps = 0
        p = plan
        self.left = p.rank - 1 if p.rank > 0 else None
        self.right = p.rank + 1 if p.rank < p.world - 1 else None
        probe = local.grid_planes(0, 1)
        shape_lo = (p.planes_lo,) + tuple(probe.shape[1:])
        shape_hi = (p.planes_hi,) + tuple(probe.shape[1:])
        self.recv_lo = torch.zeros(shape_lo, dtype=probe.dtype, device=probe.device)
        self.recv_hi = torch.zeros(shape_hi, dtype=probe.dtype, device=probe.device)
        self.halo_bytes = (self.recv_lo.numel() + self.recv_hi.numel()) * probe.element_size()
        self.symm = None
        self.halo = halo if plan.world > 1 else "none"
        if halo == "symm" and plan.world > 1:
            try:
                self.symm = SymmHalo(plan, tuple(probe.shape[1:]), probe.dtype, probe.device, group, fabric)
            except Exception as e:      # no symmetric memory on this torch build / topology: the NCCL exchange does the same job
                import warnings
                warnings.warn(f"symmetric-memory halo unavailable ({type(e).__name__}: {e}); using NCCL send/recv")
                self.halo = "p2p"
            # every rank must take the same transport: one that could not map its peers sends everybody back to NCCL
            ok = torch.tensor([1 if self.symm is not None else 0], dtype=torch.int32, device=probe.device)
            dist.all_reduce(ok, op=dist.ReduceOp.MIN, group=group)
            if int(ok.item()) == 0:
                self.symm, self.halo = None, "p2p"
        self.migrated = 0
        self.rebalanced = 0
        self._pending = None
        if hasattr(local, "count_leavers_async") and plan.world > 1:
            # the leaver count is acted upon one substep after it was taken: a particle may be margin-1 substeps
            # past its slab when it is finally handed over, so one margin layer is not enough
            if plan.margin < 2:
                raise ValueError("lagged migration (count_leavers_async) needs a halo margin of at least 2 cells")
            if self.migrate_every + 1 > max(2, plan.margin):
                self.migrate_every = max(1, plan.margin - 1)
            # Local solvers that also count the URGENT leavers (more than `slack` cells outside the owned range) let
            # particles stray inside the halo margin and hand them over only when one has used up the slack -- or when
            # so many have left that the load shifts -- instead of every period.  With the count taken every p
            # substeps and acted upon one substep late, a particle is at most slack + p cells out when it is moved:
            # slack + p <= margin - 1.
            if hasattr(local, "set_leaver_slack") and migrate_every is None:
                self.migrate_every = max(1, plan.margin // 2)
                self.slack = max(0, plan.margin - 1 - self.migrate_every)
                local.set_leaver_slack(self.slack)

    # -- halo planes: exchange partial sums with both neighbours ------------------
    def _exchange_halos(self) -> None:
        p, L = self.plan, self.local
        if self.symm is not None:
            lo, hi = self.symm.exchange(self.steps,
                                        L.grid_planes(0, p.planes_lo) if self.left is not None else None,
                                        L.grid_planes(p.n_local_x - p.planes_hi, p.n_local_x) if self.right is not None else None)
            if lo is not None:
                self.recv_lo = lo
            if hi is not None:
                self.recv_hi = hi
            return
        ops = []
        if self.right is not None:
            ops.append(dist.P2POp(dist.isend, L.grid_planes(p.n_local_x - p.planes_hi, p.n_local_x), self.right, self.group))
            ops.append(dist.P2POp(dist.irecv, self.recv_hi, self.right, self.group))
        if self.left is not None:
            ops.append(dist.P2POp(dist.isend, L.grid_planes(0, p.planes_lo), self.left, self.group))
            ops.append(dist.P2POp(dist.irecv, self.recv_lo, self.left, self.group))
        if ops:
            for req in dist.batch_isend_irecv(ops):
                req.wait()

    def substep(self, n: int = 1) -> None:
        for _ in range(n):
            self._mark("start")
            self.local.scatter()
            self._mark("scatter")
            self._exchange_halos()
            self._mark("halo")
            self.local.grid_update(self.recv_lo, self.plan.planes_lo, self.recv_hi, self.plan.planes_hi)
            self._mark("grid")
            self.local.gather()
            self._mark("gather")
            self.steps += 1
            if self.plan.world > 1:
                # a re-cut delivers strays too and stands in for this step's migration round; a declined one does not
                if not (self.rebalance_every and self.steps % self.rebalance_every == 0 and self.rebalance()):
                    self._maybe_migrate()
                self._mark("migrate")

    def _maybe_migrate(self) -> None:
        """Migration cadence.  Local solvers that can count their leavers asynchronously
        (``count_leavers_async``) are polled one substep late: the count of substep k -- max-reduced
        over the ranks on the device, so that all ranks take the same decision -- is read after
        substep k+1 has been queued: the host never drains the GPU."""
        L = self.local
        due = self.steps % self.migrate_every == 0
        if not hasattr(L, "count_leavers_async"):
            if due:
                self.migrate()
            return
        migrated_now = False
        if self._pending is not None:
            k, handle = self._pending
            if self.steps > k:                      # one substep of GPU work is queued behind the count
                self._pending = None
                count = L.read_leaver_count(handle)     # already the max over all ranks
                leavers, urgent = count if isinstance(count, tuple) else (count, count)
                if urgent > 0 or leavers >= getattr(L, "leaver_limit", 1):
                    self.migrate(hint=leavers)
                    migrated_now = True
        # the device counter was filled by this substep's G2P, i.e. BEFORE a migration that has just run: it would
        # still count the particles that were handed over.  Take the next count one period later instead.
        if due and self._pending is None and not migrated_now:
            cnt = L.count_leavers_async(self.plan.own_lo, self.plan.own_hi)      # (1,) int64 on the device
            dist.all_reduce(cnt, op=dist.ReduceOp.MAX, group=self.group)         # every rank takes the same decision
            self._pending = (self.steps, L.stage_leaver_count(cnt))

    # -- optional per-phase CUDA-event timing (bench --slab-timing) ----------------
    timing = None      # dict phase -> accumulated ms when enabled

    def enable_timing(self) -> None:
        self.timing, self._events = {}, []

    def _mark(self, name: str) -> None:
        if self.timing is None or self.local.device.type != "cuda":
            return
        ev = torch.cuda.Event(enable_timing=True)
        ev.record()
        self._events.append((name, ev))

    def collect_timing(self) -> dict:
        torch.cuda.synchronize()
        prev = None
        for name, ev in self._events:
            if name != "start" and prev is not None:
                self.timing[name] = self.timing.get(name, 0.0) + prev.elapsed_time(ev)
            prev = ev
        self._events = []
        return dict(self.timing)

    # -- slab rebalancing (per-slab particle counts every k steps) -------
    def imbalance(self, layer_cost_per_cell: float = 0.03):
        """Collective.  (counts per base-cell layer summed over the ranks, cost of the most loaded
        rank / mean cost) under the present cut."""
        p = self.plan
        cells = p.res[0] - 1
        hist = self.local.layer_histogram(cells)
        dist.all_reduce(hist, op=dist.ReduceOp.SUM, group=self.group)
        counts = hist.cpu().numpy()
        cost = counts + layer_cost_per_cell * p.res[1] * p.res[2]
        loads = [cost[lo:hi].sum() for lo, hi in p.all_ranges]
        return counts, float(max(loads) / max(np.mean(loads), 1e-300))

    def rebalance(self, min_gain: float = 0.1, layer_cost_per_cell: float = 0.03) -> bool:
        """Collective.  Re-cut the slabs so that every rank carries about the same cost -- particles
        plus ``layer_cost_per_cell`` per base cell of its range (the binning scans every cell; 0.03 is
        the measured scan time per cell over the substep time per particle) -- and move every particle
        to the owner of its base cell under the new cut, any number of ranks away.  The cut is a pure
        function of the all-reduced layer histogram, so the ranks agree without further talk.  Nothing
        happens (False) unless the most loaded rank gets at least ``min_gain`` lighter.

        Call it between substeps.  It subsumes a migration round (a particle that strayed into the halo
        margin is delivered to its owner too) and voids a pending lagged leaver count."""
        p, L = self.plan, self.local
        if p.world == 1:
            return False
        counts, _ = self.imbalance(layer_cost_per_cell)
        layer_cost = layer_cost_per_cell * p.res[1] * p.res[2]
        new = SlabPlan.balanced_ranges(counts, p.world, SlabPlan.min_cells(p.margin), layer_cost)
        cost = counts + layer_cost

        def worst(ranges):
            return max(cost[lo:hi].sum() for lo, hi in ranges)
        if tuple(new) == tuple(p.all_ranges) or worst(new) > (1.0 - min_gain) * worst(p.all_ranges):
            return False
        new_plan = SlabPlan.make(p.res, p.world, p.rank, p.margin, ranges=new)
        # split the local particles by destination (base cells outside 0 .. cells-1 are out-of-grid
        # particles: they stay with the edge ranks, as in the kernels' trailing bin)
        (data, ids), base_x = L.take_all()
        cuts = torch.tensor([hi for _, hi in new[:-1]], dtype=torch.int64, device=base_x.device)
        dest = torch.bucketize(base_x, cuts, right=True)
        parts = []
        for r in range(p.world):
            sel = torch.nonzero(dest == r).flatten()
            parts.append((data[:, sel].contiguous(), ids[sel].contiguous()))
        del data, ids, dest
        dev = L.device
        n_out = torch.tensor([q[1].numel() for q in parts], dtype=torch.int64, device=dev)
        table = [torch.zeros_like(n_out) for _ in range(p.world)]
        dist.all_gather(table, n_out, group=self.group)
        n_in = [int(t[p.rank]) for t in table]                 # n_in[r]: particles rank r holds for this rank
        rows = L.payload_rows()
        recv, ops = {}, []
        for r in range(p.world):
            if r == p.rank:
                continue
            if parts[r][1].numel():
                ops += [dist.P2POp(dist.isend, parts[r][0], r, self.group), dist.P2POp(dist.isend, parts[r][1], r, self.group)]
            if n_in[r]:
                recv[r] = (torch.empty((rows, n_in[r]), dtype=L.dtype, device=dev),
                           torch.empty((n_in[r],), dtype=torch.int32, device=dev))
                ops += [dist.P2POp(dist.irecv, recv[r][0], r, self.group), dist.P2POp(dist.irecv, recv[r][1], r, self.group)]
        if ops:
            for req in dist.batch_isend_irecv(ops):
                req.wait()
        L.rebuild(new_plan, sum(n_in))
        L.append(parts[p.rank])
        for r in sorted(recv):
            L.append(recv[r])
            self.migrated += n_in[r]
        self.plan = new_plan
        self._pending = None
        self._cap_limit = None
        self.rebalanced += 1
        return True

    # -- particle migration to the +-1 neighbours -------------------------------
    def _migrate_cap(self, hint) -> int:
        """Outbox capacity per side, the same on every rank (``hint`` is the max-reduced lagged leaver count, taken
        one substep ago: room for another substep of motion and then some; leavers that do not fit stay one round)."""
        want = 4096 if hint is None else 3 * int(hint) + 4096
        cap = 1 << max(12, (want - 1).bit_length())
        if getattr(self, "_cap_limit", None) is None:
            # the smallest local limit (scratch lists of the pack kernel live in the binning workspace), agreed once
            # per cut: every rank must post the same message size
            lim = torch.tensor([self.local.migrate_cap_limit()], dtype=torch.int64, device=self.local.device)
            dist.all_reduce(lim, op=dist.ReduceOp.MIN, group=self.group)
            self._cap_limit = int(lim.item())
        return int(min(cap, 1 << 22, self._cap_limit))

    def _migrate_device(self, hint) -> None:
        """One round with the local solver's pack / unpack kernels (``ffmpm_migrate_pack`` / ``_unpack``): fixed-size
        messages whose headers carry the counts, so nothing between the pack and the unpack needs the host; the
        round's record is read once at the end."""
        L = self.local
        cap = self._migrate_cap(hint)
        out_lo, out_hi, in_lo, in_hi = L.migrate_pack(cap, self.left is not None, self.right is not None)
        ops = []
        if self.right is not None:
            ops += [dist.P2POp(dist.isend, out_hi, self.right, self.group), dist.P2POp(dist.irecv, in_hi, self.right, self.group)]
        if self.left is not None:
            ops += [dist.P2POp(dist.isend, out_lo, self.left, self.group), dist.P2POp(dist.irecv, in_lo, self.left, self.group)]
        if ops:
            for req in dist.batch_isend_irecv(ops):
                req.wait()
        rec = L.migrate_unpack(cap, in_lo, in_hi)
        self.migrated += rec["in_lo"] + rec["in_hi"]
        self.migrate_rounds = getattr(self, "migrate_rounds", 0) + 1
        self.migrate_overflow = getattr(self, "migrate_overflow", 0) + rec["overflow"]

    def migrate(self, hint=None) -> None:
        p, L = self.plan, self.local
        if hasattr(L, "migrate_pack"):
            return self._migrate_device(hint)
        # a domain-end rank has no neighbour on that side: nothing leaves there (a particle outside the global grid
        # stays local and is flagged by the solver, as the reference raises for it)
        left, right = L.extract_leavers(p.own_lo if self.left is not None else -(1 << 62),
                                        p.own_hi if self.right is not None else (1 << 62))
        dev = L.device
        rows = L.payload_rows()
        n_out = torch.tensor([left[0].shape[1] if self.left is not None else 0,
                              right[0].shape[1] if self.right is not None else 0], dtype=torch.int64, device=dev)
        n_in = torch.zeros(2, dtype=torch.int64, device=dev)
        ops = []
        if self.left is not None:
            ops += [dist.P2POp(dist.isend, n_out[0:1], self.left, self.group),
                    dist.P2POp(dist.irecv, n_in[0:1], self.left, self.group)]
        if self.right is not None:
            ops += [dist.P2POp(dist.isend, n_out[1:2], self.right, self.group),
                    dist.P2POp(dist.irecv, n_in[1:2], self.right, self.group)]
        for req in dist.batch_isend_irecv(ops):
            req.wait()
        cnt = n_in.tolist()
        out_cnt = n_out.tolist()
        recv = {}
        ops = []
        for side, nb, payload, k_out, k_in in (("l", self.left, left, out_cnt[0], cnt[0]),
                                              ("r", self.right, right, out_cnt[1], cnt[1])):
            if nb is None:
                continue
            if k_out:
                ops += [dist.P2POp(dist.isend, payload[0].contiguous(), nb, self.group),
                        dist.P2POp(dist.isend, payload[1].contiguous(), nb, self.group)]
            if k_in:
                recv[side] = (torch.empty((rows, k_in), dtype=L.dtype, device=dev),
                              torch.empty((k_in,), dtype=torch.int32, device=dev))
                ops += [dist.P2POp(dist.irecv, recv[side][0], nb, self.group),
                        dist.P2POp(dist.irecv, recv[side][1], nb, self.group)]
        if ops:
            for req in dist.batch_isend_irecv(ops):
                req.wait()
        for side in ("l", "r"):
            if side in recv:
                L.append(recv[side])
                self.migrated += recv[side][0].shape[1]


# --------------------------------------------------------------------------- #
class CudaSlab(LocalSlab):
    """Rank-local CUDA solver for one slab (wraps MpmSolver)."""

    def __init__(self, plan: SlabPlan, dx: float, dt: float, volume: float, gravity: float, hardening: float, *,
                 capacity: int, device, dtype=torch.float32, p2g_mode: str = "auto"):
        self.device = torch.device(device)
        self.dtype = dtype
        self.dx, self.inv_dx = dx, 1.0 / dx
        self._scalars = (dt, volume, gravity, hardening)
        self._p2g_mode = p2g_mode
        import os
        # only the binned, non-fused reordering G2P (g2p_tiled3_kernel) fills the device leaver counter
        self._g2p_counts = p2g_mode in ("auto", "tiled") and os.environ.get("FFMPM_FUSE", "0") in ("", "0")
        self.launches_carried = 0      # kernel launches of the solvers a rebalancing replaced
        self._make_solver(plan, capacity)

    def _make_solver(self, plan: SlabPlan, capacity: int) -> None:
        from .mpm import MpmSolver
        dt, volume, gravity, hardening = self._scalars
        self.plan = plan
        self.solver = MpmSolver(3, list(plan.res), dt, volume, gravity, hardening, capacity=capacity, dx=self.dx,
                                inv_dx=self.inv_dx, dtype=self.dtype, device=self.device,
                                n_nodes=(plan.n_local_x, plan.res[1] + 1, plan.res[2] + 1),
                                origin=(plan.g_lo, 0, 0), per_particle_material=None, p2g_mode=self._p2g_mode,
                                reorder=True)
        # the binned G2P counts the particles that left [own_lo, own_hi) while it advects them
        self.solver.set_owned_range(plan.own_lo if plan.rank > 0 else -(2 ** 31),
                                    plan.own_hi if plan.rank < plan.world - 1 else 2 ** 31 - 1)
        self.solver.set_owned_slack(getattr(self, "_slack", 0))

    @property
    def num_particles(self) -> int:
        return self.solver.num_particles

    def set_particles(self, x, v, F, C_, mass, mu0, lam0, ids) -> None:
        """Collective when torch.distributed is initialised: the ranks agree on ONE material table
        (the union of their distinct (mass, mu0, lam0) triples), so a migrating particle's row index
        means the same thing on the receiving rank."""
        s = self.solver
        n = len(x)
        trip = np.stack([np.broadcast_to(np.asarray(a, dtype=np.float64), (n,)) for a in (mass, mu0, lam0)], 1)
        np_dt = np.float64 if self.dtype == torch.float64 else np.float32
        mine = np.unique(trip.astype(np_dt), axis=0)[:257]
        also = mine
        if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
            parts = [None] * dist.get_world_size()
            dist.all_gather_object(parts, mine)
            also = np.unique(np.concatenate([p.reshape(-1, 3) for p in parts], 0), axis=0)
        s.set_particles(x, v, F, C_, None, mass, mu0, lam0, also_materials=also)
        s.buffers[0].id[:s.num_particles] = torch.as_tensor(np.asarray(ids), device=self.device).to(torch.int32)

    def scatter(self) -> None:
        self.solver.scatter()

    def grid_planes(self, a: int, b: int) -> torch.Tensor:
        return self.solver.grid(readonly=True)[a:b]

    def grid_update(self, recv_lo, planes_lo, recv_hi, planes_hi) -> None:
        from . import _native as N
        s = self.solver
        N.check(s.lib.ffmpm_grid_op_halo(s._h, recv_lo.data_ptr() if planes_lo else None, planes_lo,
                                         recv_hi.data_ptr() if planes_hi else None, planes_hi, s._stream()))

    def gather(self) -> None:
        self.solver.gather()

    def payload_rows(self) -> int:
        return 3 + 3 + 9 + 9 + 3

    def _pack(self, b, idx):
        """27 payload rows: x v C F, then the material -- (mass, mu0, lam0) with planes; with a table
        the row index (exact as a float, the table is the same on every rank) and two unused rows."""
        rows = [b.x[:, idx], b.v[:, idx], b.C[:, idx], b.F[:, idx]]
        if b.mass is not None:
            rows += [b.mass[idx][None], b.mu0[idx][None], b.lam0[idx][None]]
        else:
            mat = torch.zeros((3, idx.numel()), dtype=b.x.dtype, device=b.x.device)
            if b.material is not None:
                mat[0] = b.material[idx].to(b.x.dtype)
            rows.append(mat)
        return torch.cat(rows, 0), b.id[idx]

    # -- migration on the device (csrc/mpm_migrate.cuh) ----------------------------------------
    def migrate_cap_limit(self) -> int:
        return max(1, self.solver.capacity // 2)

    def _mig_buffers(self, cap: int):
        key = (cap, id(self.solver))
        if getattr(self, "_mig_key", None) != key:
            from . import _native as N
            rows = int(self.solver.lib.ffmpm_migrate_rows(self.solver._h))
            mk = lambda: torch.zeros((rows + 1, cap), dtype=self.dtype, device=self.device)
            self._mig = [mk(), mk(), mk(), mk()]                     # out_lo, out_hi, in_lo, in_hi
            self._mig_rec = torch.zeros(8, dtype=torch.int32).pin_memory()
            self._mig_key = key
        return self._mig

    def migrate_pack(self, cap: int, has_left: bool, has_right: bool):
        """Leavers into the two outboxes (``None`` side = domain end: nothing leaves there), holes back-filled from the
        tail; returns (out_lo, out_hi, in_lo, in_hi) message tensors of shape (rows + 1, cap)."""
        from . import _native as N
        s = self.solver
        out_lo, out_hi, in_lo, in_hi = self._mig_buffers(cap)
        N.check(s.lib.ffmpm_migrate_pack(s._h, out_lo.data_ptr() if has_left else None, out_hi.data_ptr() if has_right else None,
                                         cap, s._stream()))
        self._mig_sides = (has_left, has_right)
        return out_lo, out_hi, in_lo, in_hi

    def migrate_unpack(self, cap: int, in_lo, in_hi) -> dict:
        """Append the received particles, read the round's record (the one host synchronisation of a round) and
        take over the new particle count."""
        from . import _native as N
        s = self.solver
        has_left, has_right = self._mig_sides
        N.check(s.lib.ffmpm_migrate_unpack(s._h, in_lo.data_ptr() if has_left else None, in_hi.data_ptr() if has_right else None,
                                           cap, self._mig_rec.data_ptr(), s._stream()))
        ev = torch.cuda.Event()
        ev.record()
        ev.synchronize()
        out_l, out_h, got_l, got_h, n_new, overflow = (int(v) for v in self._mig_rec[:6])
        if overflow >= (1 << 24):
            raise RuntimeError(f"slab capacity {s.capacity} exceeded by migration")
        N.check(s.lib.ffmpm_set_num_particles(s._h, n_new))
        s.num_particles = n_new
        return {"out_lo": out_l, "out_hi": out_h, "in_lo": got_l, "in_hi": got_h, "n": n_new, "overflow": overflow}

    def set_leaver_slack(self, slack: int) -> None:
        self._slack = int(slack)
        self.solver.set_owned_slack(self._slack)

    @property
    def leaver_limit(self) -> int:
        """So many strays inside the margin that handing them over is worth a round whatever the slack says."""
        return max(4096, self.solver.capacity // 64)

    def count_leavers_async(self, own_lo: int, own_hi: int):
        """Device-side counts (2,) of the particles whose base cell left [own_lo, own_hi) and of those more than the
        slack outside it."""
        s = self.solver
        if self._g2p_counts and s.num_particles > 0:
            return s.leaver_count().to(torch.int64)      # filled by the last G2P, no extra pass over x
        x0 = s.live.x[0, :s.num_particles]
        # domain-end ranks have no neighbour on that side: nothing "leaves" there (a particle outside the global
        # grid is reported by the binning, as the reference raises for it) -- the thresholds of _make_solver
        p = self.plan
        k = getattr(self, "_slack", 0)
        t_lo = self._threshold(own_lo) if p.rank > 0 else float("-inf")
        t_hi = self._threshold(own_hi) if p.rank < p.world - 1 else float("inf")
        u_lo = self._threshold(own_lo - k) if p.rank > 0 else float("-inf")
        u_hi = self._threshold(own_hi + k) if p.rank < p.world - 1 else float("inf")
        return torch.stack([torch.count_nonzero((x0 < t_lo) | (x0 >= t_hi)),
                            torch.count_nonzero((x0 < u_lo) | (x0 >= u_hi))]).to(torch.int64)

    def stage_leaver_count(self, cnt):
        if not hasattr(self, "_cnt_host"):
            self._cnt_host = torch.zeros(2, dtype=torch.int64).pin_memory()
        self._cnt_host.copy_(cnt, non_blocking=True)
        ev = torch.cuda.Event()
        ev.record()
        return ev

    def read_leaver_count(self, handle):
        handle.synchronize()
        return int(self._cnt_host[0]), int(self._cnt_host[1])

    def _threshold(self, cell: int):
        """Smallest value of the storage dtype whose base cell is >= ``cell``."""
        np_dt = np.float64 if self.dtype == torch.float64 else np.float32
        if cell <= 0:
            return float("-inf")          # base 0 also holds x*inv_dx - 0.5 in (-1, 0): truncation toward zero (quirk 1)
        x = np_dt((cell + 0.5) / self.inv_dx)

        def base(v):
            return int(np.float64(v) * self.inv_dx - 0.5)
        while base(x) >= cell:
            x = np.nextafter(x, np_dt(-np.inf))
        while base(x) < cell:
            x = np.nextafter(x, np_dt(np.inf))
        return float(x)

    def _store(self, payload, at: int, buf: int) -> None:
        data, ids = payload
        k = data.shape[1]
        b = self.solver.buffers[buf]
        if at + k > b.cap:
            raise RuntimeError(f"slab capacity {b.cap} exceeded by migration ({at + k} particles)")
        b.x[:, at:at + k] = data[0:3]
        b.v[:, at:at + k] = data[3:6]
        b.C[:, at:at + k] = data[6:15]
        b.F[:, at:at + k] = data[15:24]
        if b.mass is not None:
            b.mass[at:at + k] = data[24]
            b.mu0[at:at + k] = data[25]
            b.lam0[at:at + k] = data[26]
        elif b.material is not None:
            b.material[at:at + k] = data[24].to(torch.uint8)
        b.id[at:at + k] = ids

    def append(self, payload) -> None:
        s = self.solver
        n = s.num_particles
        live = s.live_index
        self._store(payload, n, live)
        s._bind(n + payload[0].shape[1], cur=live)

    # -- slab rebalancing ---------------------------------------------------------
    def _base_x(self, n: int) -> torch.Tensor:
        """Global base-cell layer of the first n live particles, by the kernels' own expression
        (trunc toward zero of x*inv_dx - 0.5 evaluated in fp64 from the stored position)."""
        x0 = self.solver.live.x[0, :n]
        return torch.trunc(x0.double() * self.inv_dx - 0.5).to(torch.int64)

    def layer_histogram(self, cells: int) -> torch.Tensor:
        n = self.solver.num_particles
        if n == 0:
            return torch.zeros(cells, dtype=torch.int64, device=self.device)
        return torch.bincount(self._base_x(n).clamp_(0, cells - 1), minlength=cells)

    def take_all(self):
        s = self.solver
        n = s.num_particles
        idx = torch.arange(n, dtype=torch.int64, device=self.device)
        payload, base_x = self._pack(s.live, idx), self._base_x(n)
        s._bind(0, cur=s.live_index)
        return payload, base_x

    def rebuild(self, plan: SlabPlan, n_particles: int) -> None:
        """New local grid (extent, origin, workspace) for ``plan``; the material representation the
        ranks agreed on in ``set_particles`` is carried over, so payload rows keep their meaning."""
        old = self.solver
        b = old.live
        kind = "planes" if b.mass is not None else ("rows" if b.material is not None else "none")
        table = old.material_table
        capacity = max(old.capacity, n_particles + n_particles // 4 + 1024)
        torch.cuda.synchronize(self.device)
        self.launches_carried += old.launch_count()
        old.close()
        old.buffers, old.workspace = [], None
        self.solver = None
        del old, b
        self._make_solver(plan, capacity)
        self.solver.adopt_material_layout(kind, table)

    def state_by_id(self):
        """(ids, x, v, F, C) of the local particles, for gathering / validation."""
        s = self.solver
        b, n = s.live, s.num_particles
        return (b.id[:n].clone(), b.x[:, :n].t().contiguous(), b.v[:, :n].t().contiguous(),
                b.F[:, :n].t().reshape(n, 3, 3).contiguous(), b.C[:, :n].t().reshape(n, 3, 3).contiguous())


class SlabSolver:
    """Convenience bundle used by bench.py: plan + CudaSlab + SlabDriver, with the
    MpmSolver-like surface bench.py needs."""

    def __init__(self, plan, local, driver):
        self.plan, self.local, self.driver = plan, local, driver
        self.reorder = True

    @classmethod
    def from_scene(cls, scene, rank: int, world: int, device, p2g_mode: str = "auto", margin: int = 2,
                   capacity_factor: float = 1.25, halo: str = "p2p"):
        """Weak scaling (BASELINE configs[3]): the scene's block is replicated once per
        rank along x; the global grid is (res*world) x res x res cells, dx = 1/res."""
        res = scene.res
        plan = SlabPlan.make((res * world, res, res), world, rank, margin)
        dx = 1.0 / res
        local = CudaSlab(plan, dx, scene.dt, scene.volume, scene.gravity, scene.hardening,
                         capacity=int(scene.n * capacity_factor), device=device, p2g_mode=p2g_mode)
        x = scene.x.copy()
        x[:, 0] += np.float32(rank)          # shift the block into this rank's slab
        ids = np.arange(scene.n, dtype=np.int64) + rank * scene.n
        local.set_particles(x, scene.v, scene.F, scene.C, scene.mass, scene.mu_0, scene.lambda_0,
                            (ids % (2 ** 31)).astype(np.int32))
        return cls(plan, local, SlabDriver(plan, local, halo=halo))

    @classmethod
    def from_bar(cls, scene, rank: int, world: int, device, p2g_mode: str = "auto", margin: int = 4,
                 capacity_factor: float = 1.3, halo: str = "p2p", drift_cells_per_substep: float = 0.1,
                 yz_cells=None, end_clearance=None):
        """COUPLED weak scaling (BASELINE configs[3] "with halo exchange + particle migration"): ONE elastic bar that
        runs through every slab of the (res*world) x res x res domain -- res x yz_cells cells of 8 particles per GPU,
        i.e. the 16.8 M particles per GPU of the headline block at res 256 -- drifting along +x at
        ``drift_cells_per_substep``, so that every halo plane carries mass and momentum and every migration period
        hands a layer of particles to the next rank.  Each rank generates the particles whose base cell it owns
        (the material, dt and perturbation of ``scene``, the per-GPU block of the same resolution)."""
        from . import scenes
        res = scene.res
        yz_cells = yz_cells or (res // 2, res // 4)                        # res 256: 256 x 128 x 64 cells per GPU
        end_clearance = end_clearance or (max(2, res // 32), max(4, 3 * res // 32))
        plan = SlabPlan.make((res * world, res, res), world, rank, margin)
        dx = 1.0 / res
        gx0, gx1 = end_clearance[0], res * world - end_clearance[1]
        lo, hi = max(gx0, plan.own_lo), min(gx1, plan.own_hi + 1)          # cell c feeds base cells c-1 and c
        cy, cz = yz_cells
        sc = scenes.elastic_block(3, res, 0, 2, seed=1000 + rank, shape=(max(hi - lo, 0), cy, cz),
                                  origin_cell=(lo, (res - cy) // 2, (res - cz) // 2))
        base = np.trunc(sc.x[:, 0].astype(np.float64) * res - 0.5).astype(np.int64)
        keep = (base >= (plan.own_lo if rank > 0 else -1)) & (base < (plan.own_hi if rank < world - 1 else 1 << 40))
        x, v, F, C_ = sc.x[keep], sc.v[keep].copy(), sc.F[keep], sc.C[keep]
        v[:, 0] += np.float32(drift_cells_per_substep * dx / sc.dt)
        n = len(x)
        local = CudaSlab(plan, dx, sc.dt, sc.volume, sc.gravity, sc.hardening,
                         capacity=int(res * cy * cz * 8 * capacity_factor), device=device, p2g_mode=p2g_mode)
        ids = (np.arange(n, dtype=np.int64) + rank * (1 << 27)) % (2 ** 31)
        local.set_particles(x, v, F, C_, sc.mass, sc.mu_0, sc.lambda_0, ids.astype(np.int32))
        obj = cls(plan, local, SlabDriver(plan, local, halo=halo))
        obj.scene_name = (f"3D elastic bar {gx1 - gx0}x{cy}x{cz} cells x 8 ppc through {world} slabs of {res}^3, "
                          f"drifting {drift_cells_per_substep} cells/substep along x")
        obj.scene_dt = sc.dt
        return obj

    @classmethod
    def from_dam_break(cls, rank: int, world: int, device, res: int = 256, n_total: int = 33_554_432,
                       margin: int = 4, capacity: Optional[int] = None, p2g_mode: str = "auto", halo: str = "p2p",
                       late: bool = False):
        """BASELINE configs[4]: soft column at one x-end of a (res*world) x res x res domain;
        every rank generates the particles of its own x interval.  The even cut leaves the ranks
        away from the column empty; ``rebalance()`` re-cuts the slabs by particle count."""
        from . import scenes
        plan = SlabPlan.make((res * world, res, res), world, rank, margin)
        dx = 1.0 / res
        lo = (plan.own_lo + 0.5) * dx if rank > 0 else -1.0
        hi = (plan.own_hi + 0.5) * dx if rank < world - 1 else float(world) + 1.0
        sc = scenes.dam_break_slab(world, rank, res, n_total, x_range=(lo, hi), late=late)
        cap = capacity or max(int(n_total * 0.6), sc.n + 1024)
        local = CudaSlab(plan, dx, sc.dt, sc.volume, sc.gravity, sc.hardening, capacity=cap, device=device,
                         p2g_mode=p2g_mode)
        ids = (np.arange(sc.n, dtype=np.int64) + rank * (2 ** 27)) % (2 ** 31)
        # collective (the ranks agree on the material table): ranks that start empty take part too
        local.set_particles(sc.x, sc.v, sc.F, sc.C, sc.mass, sc.mu_0, sc.lambda_0, ids.astype(np.int32))
        obj = cls(plan, local, SlabDriver(plan, local, halo=halo))
        obj.scene_name = sc.name
        return obj

    @property
    def num_particles(self) -> int:
        return self.local.num_particles

    def substep(self, n: int = 1) -> None:
        self.driver.substep(n)

    def rebalance(self, **kw) -> bool:
        """Collective: ``SlabDriver.rebalance``; the plan of this bundle follows the driver's."""
        done = self.driver.rebalance(**kw)
        self.plan = self.driver.plan
        return done

    def poll_error(self) -> int:
        return self.local.solver.poll_error()

    def launch_count(self) -> int:
        return self.local.solver.launch_count() + self.local.launches_carried
