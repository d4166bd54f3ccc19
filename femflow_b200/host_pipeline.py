"""Substeps for callers that keep their particle state in HOST memory (what ``solve_mls_mpm_3d`` is in the
reference: NumPy arrays in, NumPy arrays updated in place, mls_mpm.py:40-79).

One such call costs a host-to-device copy of the state, one substep and a device-to-host copy of the result; on a
B200 the two copies are ~97 % of it (2 x 1.6 GB over PCIe for 16.8 M particles against 1.5 ms of kernels).  The link
is full duplex, so INDEPENDENT calls -- the frames of a parameter sweep, the scenes of ``paper_1.multi_drop_experiment``
-- can overlap: ``HostSubstepPipeline`` keeps ``depth`` device slots in flight, the upload of call k+1 running under
the download of call k, and every call still moves its whole state both ways.

    pipe = HostSubstepPipeline(solver, depth=3)
    for k in range(n_calls):
        pipe.submit(host_in[k], host_out[k])      # dicts of pinned SoA tensors: x v C F (+ mass mu0 lam0 / material / Jp)
    pipe.drain()                                  # host_out[k] now hold x, v, C, F in the CALLER'S particle order

A dependent chain (call k+1 consumes the output of call k) cannot overlap; keep such state on the device
(``MpmSolver.substep``, ``MPMSimulation``) -- that is what the device-resident throughput measures.
"""
from __future__ import annotations

from typing import Dict, List

import torch

from . import _native as N
from .mpm import MpmSolver, _StateBuffer

import ctypes as C


class HostSubstepPipeline:
    def __init__(self, solver: MpmSolver, depth: int = 3, substeps_per_call: int = 1, step=None):
        """``step``: what advances the bound state by one call (default: ``solver.substep(substeps_per_call)``); a slab
        driver passes its own substep, which puts the halo exchange between the two halves."""
        if depth < 1:
            raise ValueError("depth must be >= 1")
        self.solver, self.depth, self.substeps = solver, depth, int(substeps_per_call)
        self.step = step if step is not None else (lambda: solver.substep(self.substeps))
        s, dev = solver, solver.device
        b0 = s.buffers[0]
        kind = "planes" if b0.mass is not None else ("rows" if b0.material is not None else "none")

        def make():
            sb = _StateBuffer(s.dim, s.capacity, s.dtype, dev, False, s.reorder, b0.Jp is not None)
            sb.set_material_storage(kind)
            return sb
        # per slot: the upload target (also the export target: it is free again once the substep has consumed it)
        # and, when the solver reorders, the buffer its G2P writes into
        self.ins: List[_StateBuffer] = [make() for _ in range(depth)]
        self.outs: List[_StateBuffer] = [make() for _ in range(depth)] if s.reorder else [None] * depth
        self.s_in, self.s_out = torch.cuda.Stream(dev), torch.cuda.Stream(dev)
        self.uploaded = [torch.cuda.Event() for _ in range(depth)]
        self.computed = [torch.cuda.Event() for _ in range(depth)]
        self.drained = [torch.cuda.Event() for _ in range(depth)]
        self._ids = torch.arange(s.capacity, dtype=torch.int32, device=dev) if s.reorder else None
        self.calls = 0
        self._saved = None

    def submit(self, host_in: Dict[str, torch.Tensor], host_out: Dict[str, torch.Tensor]) -> None:
        """Queue one call: ``host_in`` = pinned SoA planes ``x v C F`` of shape (rows, n) (+ per-particle material / Jp
        planes of shape (n,) when the solver carries them), ``host_out`` = pinned tensors for ``x v C F`` (+ ``Jp``).
        Returns immediately; ``drain()`` (or the slot coming round again) waits."""
        s = self.solver
        j, first = self.calls % self.depth, self.calls < self.depth
        n = host_in["x"].shape[-1]
        run = torch.cuda.current_stream(s.device)
        if self._saved is None:
            self._saved = (list(s.buffers), s.num_particles, s.live_index)
        with torch.cuda.stream(self.s_in):
            if not first:
                self.s_in.wait_event(self.drained[j])          # the slot's previous result has left the device
            tgt = self.ins[j]
            for name, t in host_in.items():
                getattr(tgt, name)[..., :n].copy_(t, non_blocking=True)
            self.uploaded[j].record(self.s_in)
        run.wait_event(self.uploaded[j])
        if tgt.id is not None:
            tgt.id[:n].copy_(self._ids[:n])                    # the caller's order: particle i is row i of host_in
        s.buffers = [tgt, self.outs[j]] if s.reorder else [tgt]
        s._bind(n)
        self.step()
        # back into the caller's particle order: the upload buffer is free when the substeps ping-ponged an odd number
        # of times, otherwise the other one is
        live = s.live
        dst = self.outs[j] if (s.reorder and live is tgt) else (tgt if s.reorder else None)
        if dst is not None:
            st = dst.c_struct()
            N.check(s.lib.ffmpm_export_state(s._h, C.byref(st), s._stream()))
        else:
            dst = live
        self.computed[j].record(run)
        with torch.cuda.stream(self.s_out):
            self.s_out.wait_event(self.computed[j])
            for name, t in host_out.items():
                t[..., :n].copy_(getattr(dst, name)[..., :n], non_blocking=True)
            self.drained[j].record(self.s_out)
        self.calls += 1

    def drain(self) -> None:
        """Wait until every queued call has delivered its result; gives the solver its own buffers back."""
        for ev in self.drained[:min(self.calls, self.depth)]:
            ev.synchronize()
        if self._saved is not None:
            s = self.solver
            s.buffers, n, live = self._saved
            s._bind(n, cur=live)                 # the buffer that held the solver's own live state is live again
            self._saved = None
