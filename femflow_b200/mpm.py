"""Host side of the CUDA MPM substep: owns the device tensors (torch is only the
allocator / stream provider here) and drives the C ABI of include/femflow_mpm.h.

The particle state lives on the GPU as SoA planes (one contiguous row per scalar
component).  With ``reorder=True`` (3D default) two state buffers are kept and
every substep G2P writes the particles back in cell order into the other one; the
``id`` plane remembers each particle's original index so that callers always see
the reference's fixed particle order.
"""
from __future__ import annotations

import os
import ctypes as C
from typing import Dict, Optional, Sequence

import numpy as np
import torch

from . import _native as N

_MODEL = {"neo_hookean": N.FFMPM_NEO_HOOKEAN, "snow": N.FFMPM_SNOW}
_P2G = {"auto": N.FFMPM_P2G_AUTO, "scatter": N.FFMPM_P2G_SCATTER, "tiled": N.FFMPM_P2G_TILED,
        "fused": N.FFMPM_P2G_FUSED}


def _as_tensor(a, dtype, device):
    if isinstance(a, torch.Tensor):
        return a.to(device=device, dtype=dtype)
    a = np.array(a, copy=True, order="C") if not getattr(a, "flags", None) or not a.flags.writeable else np.ascontiguousarray(a)
    return torch.as_tensor(a, device=device).to(dtype)


class _StateBuffer:
    """One SoA particle buffer."""

    def __init__(self, dim, cap, dtype, device, per_particle_material, with_id, with_jp):
        d = dim
        self.cap = cap
        self.dtype, self.device = dtype, device
        self.material = None      # uint8 rows of the material table (MpmSolver._set_material_layout)
        # x, v, C, F are views into ONE allocation of 2d + 2d^2 planes at one stride: the P2G prefetch then walks the
        # 24 planes of a window from a single per-lane pointer (csrc/mpm_p2g_bulk.cuh: p2g_planes_contiguous)
        self.planes = torch.zeros((2 * d + 2 * d * d, cap), dtype=dtype, device=device)
        self.x, self.v = self.planes[0:d], self.planes[d:2 * d]
        self.C, self.F = self.planes[2 * d:2 * d + d * d], self.planes[2 * d + d * d:]
        self.Jp = torch.ones((cap,), dtype=dtype, device=device) if with_jp else None
        if per_particle_material:
            self.mass = torch.zeros((cap,), dtype=dtype, device=device)
            self.mu0 = torch.zeros((cap,), dtype=dtype, device=device)
            self.lam0 = torch.zeros((cap,), dtype=dtype, device=device)
        else:
            self.mass = self.mu0 = self.lam0 = None
        self.id = torch.arange(cap, dtype=torch.int32, device=device) if with_id else None

    def c_struct(self) -> N.FfMpmState:
        def p(t):
            return None if t is None else t.data_ptr()
        return N.FfMpmState(p(self.x), p(self.v), p(self.C), p(self.F), p(self.Jp), p(self.mass), p(self.mu0),
                            p(self.lam0), p(self.id), p(self.material), self.cap)

    def nbytes(self) -> int:
        return sum(t.numel() * t.element_size() for t in
                   (self.planes, self.Jp, self.mass, self.mu0, self.lam0, self.id, self.material)
                   if t is not None)

    def set_material_storage(self, kind: str) -> None:
        """``planes``: three scalar planes; ``rows``: one uint8 plane; ``none``: neither."""
        if kind == "planes":
            if self.mass is None:
                self.mass = torch.zeros((self.cap,), dtype=self.dtype, device=self.device)
                self.mu0 = torch.zeros((self.cap,), dtype=self.dtype, device=self.device)
                self.lam0 = torch.zeros((self.cap,), dtype=self.dtype, device=self.device)
            self.material = None
        else:
            self.mass = self.mu0 = self.lam0 = None
            if kind == "rows":
                if self.material is None:
                    self.material = torch.zeros((self.cap,), dtype=torch.uint8, device=self.device)
            else:
                self.material = None


class MpmSolver:
    """GPU-resident MLS-MPM solver for one (slab of a) grid.

    Scalars follow ``solve_mls_mpm_3d`` (reference solvers/mpm/mls_mpm.py:40-53).
    """

    def __init__(self, dim: int, res, dt: float, volume: float, gravity: float, hardening: float, *,
                 capacity: int, dx: Optional[float] = None, inv_dx: Optional[float] = None,
                 model: str = "neo_hookean", dtype: torch.dtype = torch.float32, device="cuda:0",
                 n_nodes: Optional[Sequence[int]] = None, origin: Sequence[int] = (0, 0, 0),
                 mass: float = 0.0, mu_0: float = 0.0, lambda_0: float = 0.0,
                 per_particle_material: Optional[bool] = None, p2g_mode: str = "auto",
                 reorder: Optional[bool] = None):
        if not torch.cuda.is_available():
            raise RuntimeError("femflow_b200 needs a CUDA device; there is no CPU fallback")
        self.lib = N.lib()
        self.dim = int(dim)
        self.device = torch.device(device)
        self.dtype = dtype
        res3 = [int(r) for r in (res if isinstance(res, (list, tuple)) else [res] * self.dim)]
        while len(res3) < 3:
            res3.append(2)
        self.res = res3
        n3 = list(n_nodes) if n_nodes is not None else [r + 1 for r in res3[:self.dim]]
        while len(n3) < 3:
            n3.append(1)
        self.n = [int(v) for v in n3]
        self.origin = [int(o) for o in origin] + [0] * (3 - len(origin))
        self.capacity = (int(capacity) + 63) // 64 * 64      # plane stride: 16-byte aligned plane segments
        self.model = model
        # Per-particle material (reference Particle.mass / mu_0 / lambda_0): True = three scalar planes,
        # False = the config scalars, None (3D default) = chosen from the data in set_particles: a
        # <= 256-row table + 1-byte rows (no row plane at all for a single material), else planes.
        self._material_auto = per_particle_material is None and self.dim == 3
        self.material_layout = "config"
        self.material_table = None     # (R, 3) rows of the table in effect
        if per_particle_material is None:
            per_particle_material = False
        if per_particle_material:
            self.material_layout = "planes"
        if reorder is None:
            # 3D: binned pipeline, G2P writes the state back cell-sorted.  2D (1 M particles, state resident in L2) is
            # fastest unbinned (warp-window kernels, csrc/mpm_2d_window.cuh; measured: profiles/r02s_2d_series.json); the
            # binned 2D pipeline stays selectable with reorder=True (or FFMPM_2D_BINNED=1)
            reorder = (self.dim == 3 or os.environ.get("FFMPM_2D_BINNED") == "1") and p2g_mode != "scatter"
        self.reorder = bool(reorder)
        cfg = N.FfMpmConfig()
        cfg.dim = self.dim
        cfg.dtype = N.FFMPM_F64 if dtype == torch.float64 else N.FFMPM_F32
        cfg.model = _MODEL[model]
        for i in range(3):
            cfg.res[i] = self.res[i]
            cfg.n[i] = self.n[i]
            cfg.origin[i] = self.origin[i]
        r0 = self.res[0]
        cfg.dx = float(dx) if dx is not None else 1.0 / r0
        cfg.inv_dx = float(inv_dx) if inv_dx is not None else 1.0 / cfg.dx
        cfg.dt, cfg.volume, cfg.gravity, cfg.hardening = float(dt), float(volume), float(gravity), float(hardening)
        cfg.mass, cfg.mu_0, cfg.lambda_0 = float(mass), float(mu_0), float(lambda_0)
        cfg.p2g_mode = _P2G[p2g_mode]
        self.cfg = cfg
        self._h = N.H()
        dev_index = self.device.index if self.device.index is not None else torch.cuda.current_device()
        self.dev_index = dev_index
        N.check(self.lib.ffmpm_create(C.byref(cfg), dev_index, C.byref(self._h)))
        ws_bytes = self.lib.ffmpm_workspace_bytes(C.byref(cfg), self.capacity)
        if ws_bytes < 0:
            N.check(int(ws_bytes))
        with torch.cuda.device(self.device):
            self.workspace = torch.zeros((ws_bytes + 255) // 256 * 256, dtype=torch.uint8, device=self.device)
            with_jp = self.dim == 2 or model == "snow"
            self.buffers = [_StateBuffer(self.dim, self.capacity, dtype, self.device, per_particle_material,
                                         self.reorder, with_jp)]
            if self.reorder:
                self.buffers.append(_StateBuffer(self.dim, self.capacity, dtype, self.device,
                                                 per_particle_material, True, with_jp))
        N.check(self.lib.ffmpm_set_workspace(self._h, self.workspace.data_ptr(), self.workspace.numel()))
        self.num_particles = 0
        self._bind(0)

    # ------------------------------------------------------------------ #
    def _bind(self, n: int, cur: int = 0) -> None:
        """(Re)bind the particle buffers; ``cur`` picks which Python-side buffer holds
        the live state (the other one becomes the ping-pong target)."""
        self._order = [cur, 1 - cur] if self.reorder else [0]
        c = self.buffers[self._order[0]].c_struct()
        alt = self.buffers[self._order[1]].c_struct() if self.reorder else None
        N.check(self.lib.ffmpm_bind_state(self._h, C.byref(c), C.byref(alt) if alt is not None else None, n))
        self.num_particles = n

    def close(self) -> None:
        if getattr(self, "_h", None):
            self.lib.ffmpm_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def live_index(self) -> int:
        return self._order[self.lib.ffmpm_live_buffer(self._h)]

    @property
    def live(self) -> _StateBuffer:
        return self.buffers[self.live_index]

    def _stream(self, stream=None):
        if stream is None:
            stream = torch.cuda.current_stream(self.device)
        return C.c_void_p(stream.cuda_stream)

    # ------------------------------------------------------------------ #
    def set_particles(self, x, v=None, F=None, C_=None, Jp=None, mass=None, mu0=None, lam0=None,
                      also_materials=None) -> None:
        """Upload particle state given in the reference's layouts: ``x, v (N, d)``,
        ``F, C (N, d, d)``, ``Jp (N, 1)``; per-particle ``mass, mu0, lam0 (N,)``.
        ``also_materials``: see ``_choose_material_layout``."""
        d = self.dim
        x = _as_tensor(x, self.dtype, self.device).reshape(-1, d)
        n = x.shape[0]
        if n > self.capacity:
            raise ValueError(f"{n} particles exceed the solver capacity {self.capacity}")
        b = self.buffers[0]
        b.x[:, :n] = x.t()
        b.v[:, :n] = 0 if v is None else _as_tensor(v, self.dtype, self.device).reshape(n, d).t()
        if F is None:
            b.F[:, :n] = torch.eye(d, dtype=self.dtype, device=self.device).reshape(d * d, 1)
        else:
            b.F[:, :n] = _as_tensor(F, self.dtype, self.device).reshape(n, d * d).t()
        b.C[:, :n] = 0 if C_ is None else _as_tensor(C_, self.dtype, self.device).reshape(n, d * d).t()
        if b.Jp is not None:
            b.Jp[:n] = 1 if Jp is None else _as_tensor(Jp, self.dtype, self.device).reshape(n)
        if self._material_auto or b.mass is not None:
            given = {}
            for name, val in (("mass", mass), ("mu0", mu0), ("lam0", lam0)):
                if val is None:
                    raise ValueError(f"per-particle {name} is required")
                given[name] = _as_tensor(np.broadcast_to(np.asarray(val, dtype=np.float64), (n,))
                                         if not isinstance(val, torch.Tensor) else val, self.dtype, self.device).reshape(n)
            if self._material_auto:
                self._choose_material_layout(given, n, also_materials)
            if b.mass is not None:
                for name, t in given.items():
                    getattr(b, name)[:n] = t
        if b.id is not None:
            b.id[:n] = torch.arange(n, dtype=torch.int32, device=self.device)
        self._bind(n)

    def set_materials(self, mass, mu0, lam0) -> None:
        """Material table (``ffmpm_set_materials``): row ``i`` = ``(mass[i], mu0[i], lam0[i])``."""
        arrs = [np.ascontiguousarray(a, dtype=np.float64).reshape(-1) for a in (mass, mu0, lam0)]
        dp = C.POINTER(C.c_double)
        N.check(self.lib.ffmpm_set_materials(self._h, *(a.ctypes.data_as(dp) for a in arrs), len(arrs[0])))

    def _choose_material_layout(self, given: Dict[str, torch.Tensor], n: int, also=None) -> None:
        """Pick the cheapest exact representation of the per-particle (mass, mu0, lam0) triples.
        ``also``: (R, 3) triples that must be in the table too (slabs: the triples of the other
        ranks, so that every rank derives the same table and a migrating row index stays valid)."""
        trip = torch.stack([given["mass"], given["mu0"], given["lam0"]], 1)
        k = 0
        if also is not None and len(also):
            also = torch.as_tensor(np.asarray(also), device=self.device).to(self.dtype).reshape(-1, 3)
            k = also.shape[0]
            trip = torch.cat([also, trip], 0)
        rows = inv = None
        if trip.shape[0] == 0:
            rows = torch.tensor([[self.cfg.mass, self.cfg.mu_0, self.cfg.lambda_0]], dtype=self.dtype, device=self.device)
        elif bool((trip == trip[0]).all()):
            rows = trip[:1]
        else:
            u, i = torch.unique(trip, dim=0, return_inverse=True)     # lexicographically sorted rows
            if u.shape[0] <= 256:
                rows, inv = u, i[k:]
        if rows is None:
            kind = "planes"
        else:
            kind = "rows" if inv is not None else "none"
        for b in self.buffers:
            b.set_material_storage(kind)
        self.material_table = rows
        if rows is None:
            N.check(self.lib.ffmpm_set_materials(self._h, None, None, None, 0))
            self.material_layout = "planes"
        else:
            r = rows.double().cpu().numpy()
            self.set_materials(r[:, 0], r[:, 1], r[:, 2])
            if inv is not None:
                self.buffers[0].material[:n] = inv.to(torch.uint8)
            self.material_layout = f"table[{rows.shape[0]}]"

    def adopt_material_layout(self, kind: str, table: Optional[torch.Tensor]) -> None:
        """Take over a material representation decided elsewhere (``kind``: ``planes`` / ``rows`` /
        ``none`` as in ``_StateBuffer.set_material_storage``; ``table``: its (R, 3) rows or None) instead
        of deriving one from the data of ``set_particles`` -- a slab re-created by a rebalancing keeps
        the table all ranks agreed on.  Leaves the solver bound to zero particles."""
        if kind == "planes" and table is not None or kind == "rows" and table is None:
            raise ValueError("material planes come without a table; rows need one")
        for b in self.buffers:
            b.set_material_storage(kind)
        if kind == "none" and table is None:      # nothing was ever decided: the config scalars stay in effect
            self._bind(0)
            return
        self._material_auto = False
        self.material_table = table
        if table is None:
            N.check(self.lib.ffmpm_set_materials(self._h, None, None, None, 0))
            self.material_layout = "planes"
        else:
            r = table.double().cpu().numpy()
            self.set_materials(r[:, 0], r[:, 1], r[:, 2])
            self.material_layout = f"table[{table.shape[0]}]"
        self._bind(0)

    def get_particles(self) -> Dict[str, torch.Tensor]:
        """State in the ORIGINAL particle order, reference layouts, device tensors."""
        b = self.live
        n, d = self.num_particles, self.dim
        if b.id is not None:
            inv = torch.empty(n, dtype=torch.int64, device=self.device)
            inv[b.id[:n].long()] = torch.arange(n, device=self.device)
        else:
            inv = slice(None)
        out = {
            "x": b.x[:, :n].t()[inv].contiguous(),
            "v": b.v[:, :n].t()[inv].contiguous(),
            "F": b.F[:, :n].t()[inv].reshape(n, d, d).contiguous(),
            "C": b.C[:, :n].t()[inv].reshape(n, d, d).contiguous(),
        }
        if b.Jp is not None:
            out["Jp"] = b.Jp[:n][inv].reshape(n, 1).contiguous()
        return out

    # ------------------------------------------------------------------ #
    def clear_grid(self, stream=None):
        N.check(self.lib.ffmpm_clear_grid(self._h, self._stream(stream)))

    def bin(self, stream=None):
        N.check(self.lib.ffmpm_bin(self._h, self._stream(stream)))

    def p2g(self, stream=None):
        N.check(self.lib.ffmpm_p2g(self._h, self._stream(stream)))

    def grid_op(self, stream=None):
        N.check(self.lib.ffmpm_grid_op(self._h, self._stream(stream)))

    def g2p(self, stream=None):
        N.check(self.lib.ffmpm_g2p(self._h, self._stream(stream)))

    def scatter(self, stream=None):
        """Zeroed grid + binning + P2G (first half of a substep)."""
        N.check(self.lib.ffmpm_scatter(self._h, self._stream(stream)))

    def gather(self, stream=None):
        """G2P (second half of a substep)."""
        N.check(self.lib.ffmpm_gather(self._h, self._stream(stream)))

    def substep(self, n_substeps: int = 1, stream=None):
        N.check(self.lib.ffmpm_substep(self._h, int(n_substeps), self._stream(stream)))

    def make_graph(self, n_substeps: int) -> "torch.cuda.CUDAGraph":
        """Capture ``n_substeps`` substeps in a CUDA graph, rounded up to an even count: the state buffers (reordering
        pipelines) and the two grids (every pipeline whose grid update clears the idle grid) ping-pong, so that only an
        even number of substeps ends where it started and can be replayed back to back."""
        if n_substeps % 2:
            n_substeps += 1
        self.graph_substeps = n_substeps
        torch.cuda.synchronize(self.device)
        g = torch.cuda.CUDAGraph()
        before = self.launch_count()
        with torch.cuda.graph(g):
            self.substep(n_substeps)
        self.graph_launches = self.launch_count() - before
        return g

    def set_colliders(self, points, normals) -> None:
        """Plane colliders applied at the end of every grid update (three_d/grid_op.py:50-67)."""
        pts = np.ascontiguousarray(points, dtype=np.float64).reshape(-1, 3)
        nrm = np.ascontiguousarray(normals, dtype=np.float64).reshape(-1, 3)
        assert len(pts) == len(nrm)
        dp = C.POINTER(C.c_double)
        N.check(self.lib.ffmpm_set_colliders(self._h, pts.ctypes.data_as(dp), nrm.ctypes.data_as(dp), len(pts)))

    def set_owned_range(self, own_lo: int, own_hi: int) -> None:
        N.check(self.lib.ffmpm_set_owned_range(self._h, int(own_lo), int(own_hi)))

    def set_owned_slack(self, slack: int) -> None:
        N.check(self.lib.ffmpm_set_owned_slack(self._h, int(slack)))

    def leaver_count(self) -> torch.Tensor:
        """(2,) int32 view of the device counters the binned G2P fills (see ffmpm_set_owned_range):
        [particles outside the owned range, those more than the slack outside it]."""
        ptr = C.c_void_p()
        N.check(self.lib.ffmpm_leaver_count_ptr(self._h, C.byref(ptr)))
        off = ptr.value - self.workspace.data_ptr()
        return self.workspace[off:off + 8].view(torch.int32)

    def collide(self, stream=None) -> None:
        N.check(self.lib.ffmpm_collide(self._h, self._stream(stream)))

    def launch_count(self) -> int:
        return int(self.lib.ffmpm_launch_count(self._h))

    def grid(self, readonly: bool = False) -> torch.Tensor:
        """Node-major grid ``(nx, ny, nz, 4)`` (view into the workspace).  ``readonly=True`` promises
        not to write through the view (``ffmpm_grid_view``): the library then keeps its knowledge of
        which node blocks are non-zero."""
        ptr = C.c_void_p()
        N.check((self.lib.ffmpm_grid_view if readonly else self.lib.ffmpm_grid_ptr)(self._h, C.byref(ptr)))
        off = ptr.value - self.workspace.data_ptr()
        count = self.n[0] * self.n[1] * self.n[2] * 4
        es = 8 if self.dtype == torch.float64 else 4
        return self.workspace[off:off + count * es].view(self.dtype).view(self.n[0], self.n[1], self.n[2], 4)

    def bin_results(self):
        """(keys per live-order particle, perm, cell_offsets) as int32 device tensors."""
        k, p, o = C.c_void_p(), C.c_void_p(), C.c_void_p()
        nc = C.c_int64()
        N.check(self.lib.ffmpm_bin_ptrs(self._h, C.byref(k), C.byref(p), C.byref(o), C.byref(nc)))
        base = self.workspace.data_ptr()

        def view(ptr, count):
            off = ptr.value - base
            return self.workspace[off:off + 4 * count].view(torch.int32)
        n = self.num_particles
        return view(k, n), view(p, n), view(o, nc.value + 2), int(nc.value)

    def poll_error(self, stream=None) -> int:
        """Synchronises; returns the number of out-of-grid particle events (and clears it)."""
        code, n_oob = C.c_int32(), C.c_int64()
        rc = self.lib.ffmpm_poll_error(self._h, self._stream(stream), C.byref(code), C.byref(n_oob))
        if rc not in (N.FFMPM_OK, N.FFMPM_E_OOB):
            N.check(rc)
        return int(n_oob.value)

    def check_errors(self, stream=None) -> None:
        """Mirror of the reference's error convention: a stencil outside the grid is a
        bare ``RuntimeError`` (three_d/p2g.py:51-52)."""
        n_oob = self.poll_error(stream)
        if n_oob:
            raise RuntimeError(f"{n_oob} particle stencil(s) left the grid")

    def snapshot(self, coeff: float, out: torch.Tensor, stream=None) -> None:
        """particle.py:30-33: flat f64 vector of pos / coeff in original particle order."""
        assert out.dtype == torch.float64 and out.numel() >= self.num_particles * self.dim
        N.check(self.lib.ffmpm_snapshot(self._h, float(coeff), out.data_ptr(), self._stream(stream)))
