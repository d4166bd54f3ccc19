"""``MPMSimulation`` with the reference's constructor / load / start / reset /
``displacements`` semantics (femflow/simulation/mpm/simulation.py:19-149), running
the substeps on the GPU with the state resident in HBM.

Differences from the reference, all behind the same public surface:
  * particles live in SoA device planes (``self.solver``) instead of a numba typed
    list of ``Particle`` objects; ``self.particles`` is a ``ParticleArray`` view that
    is refreshed when the run ends;
  * the per-step position snapshot (particle.py:30-33) is one CUDA kernel writing
    ``pos / tightening_coeff`` as f64 in original particle order, copied to the host
    through a ring of pinned buffers while the next substeps run (``displacements``
    therefore trails the GPU by up to two steps during a run; it is complete when
    ``running`` drops);
  * a particle leaving the grid raises ``RuntimeError`` in the simulation thread
    exactly like the reference (three_d/p2g.py:51-52), but ``running`` is reset.
"""
from __future__ import annotations

import logging
import os
import threading
from typing import List, Tuple

import numpy as np

from ...solvers.mpm.particle import ParticleArray, map_particles_to_pos
from ...solvers.mpm.utils import Ev_to_lambda, Ev_to_mu
from ..simulation_base import SimulationBase

try:  # the reference logs through loguru (simulation.py:6); optional here
    from loguru import logger
except Exception:  # pragma: no cover
    logger = logging.getLogger("femflow_b200")
    logger.success = logger.info  # type: ignore[attr-defined]

try:
    from tqdm import tqdm
except Exception:  # pragma: no cover
    def tqdm(it, **_):
        return it


def vector_to_matrix(vec: np.ndarray, cols: int) -> np.ndarray:
    """numerics/linear_algebra.py:19-23"""
    return vec if vec.ndim == 2 else vec.reshape((vec.shape[0] // cols, cols))


class MPMSimulation(SimulationBase):
    def __init__(self, outdir: str, steps: int, dt: float, gyroid_mass: float, collider_mass: float, volume: float,
                 force: float, gyroid_E: float, collider_E: float, gyroid_v: float, collider_v: float,
                 hardening: float, grid_res: int, tightening_coeff: float, save_displacements=True,
                 device: str = "cuda:0", dtype="float32", progress: bool = True):
        super().__init__()
        self.outdir = outdir
        self.save_displacements = save_displacements
        self.loaded = False
        self.running = False
        self.displacements = []
        self.steps = steps
        self.dt = dt
        self.gyroid_mass = gyroid_mass
        self.collider_mass = collider_mass
        self.volume = volume
        self.force = force
        self.gyroid_E = gyroid_E
        self.collider_E = collider_E
        self.gyroid_v = gyroid_v
        self.collider_v = collider_v
        self.hardening = hardening
        self.grid_res = grid_res
        self.tightening_coeff = tightening_coeff

        self.gyroid_mu_0 = self.gyroid_E / (2 * (1 + self.gyroid_v))
        self.gyroid_lambda_0 = self.gyroid_E * self.gyroid_v / ((1 + self.gyroid_v) * (1 - 2 * self.gyroid_v))

        self.dx = 1 / self.grid_res
        self.inv_dx = 1 / self.dx

        self.device = device
        self.dtype = dtype
        self.progress = progress
        self.solver = None
        self.error = None
        self._thread = None

    # simulation.py:68-111
    def load(self, **kwargs):
        import torch
        from ...mpm import MpmSolver
        dim = 3  # "Always 3D for now."
        meshes: List = kwargs["meshes"]
        params: List[Tuple[float, float, float]] = kwargs["params"]
        if len(meshes) != len(params):
            raise ValueError(f"Meshes {len(meshes)} and Params  {len(params)} must be the same length")
        pos, mass, lam, mu = [], [], [], []
        for mesh, param in zip(meshes, params):
            # simulation.py:81-83: float32 vertices * coeff (float32 product), then widened to f64
            verts = np.asarray(mesh.vertices)
            p = (vector_to_matrix(verts.copy(), 3) * self.tightening_coeff).astype(np.float64)
            pos.append(p)
            mass.append(np.full(len(p), float(param[0])))
            lam.append(np.full(len(p), Ev_to_lambda(*param[1:])))
            mu.append(np.full(len(p), Ev_to_mu(*param[1:])))
        pos = np.concatenate(pos) if pos else np.zeros((0, 3))
        self.particles = ParticleArray(pos, np.concatenate(mass) if mass else [], np.concatenate(lam) if lam else [],
                                       np.concatenate(mu) if mu else [], self.force)
        n = len(self.particles)
        if self.save_displacements:
            self.displacements = [map_particles_to_pos(self.particles, self.tightening_coeff)]

        self.v = np.zeros((n, dim), dtype=np.float64)
        self.F = np.tile(np.eye(dim, dtype=np.float64), (n, 1, 1))
        self.C = np.zeros((n, dim, dim), dtype=np.float64)
        self.Jp = np.ones((n, 1), dtype=np.float64)

        if self.solver is not None:
            self.solver.close()
        tdtype = getattr(torch, self.dtype) if isinstance(self.dtype, str) else self.dtype
        self.solver = MpmSolver(3, self.grid_res, self.dt, self.volume, self.force, self.hardening,
                                capacity=max(n, 1), dx=self.dx, inv_dx=self.inv_dx, dtype=tdtype, device=self.device)
        self.solver.set_particles(self.particles.pos, self.v, self.F, self.C, None, self.particles.mass,
                                  self.particles.mu_0, self.particles.lambda_0)
        # the uploads above ran on this thread's current stream; the substeps run on a stream of the simulation thread
        # (start()): make the state visible to whatever stream comes next
        torch.cuda.synchronize(self.solver.device)
        self.loaded = True
        logger.success("Simulation loaded")

    # simulation.py:113-117
    def start(self, **kwargs):
        if not self.loaded:
            logger.error("Please load the simulation first")
            return
        self._thread = threading.Thread(target=self._simulate_offline, daemon=True)
        self._thread.start()

    def join(self, timeout=None):
        """Not in the reference (its GUI polls ``running``): wait for the simulation thread."""
        if self._thread is not None:
            self._thread.join(timeout)

    # simulation.py:119-120
    def reset(self, **kwargs):
        self.load(**kwargs)

    # simulation.py:122-149
    def _simulate_offline(self):
        import torch
        self.running = True
        self.error = None
        try:
            s = self.solver
            n = s.num_particles
            torch.cuda.set_device(s.device)
            stream = torch.cuda.Stream(device=s.device)
            # Snapshot ring: substep k+1 and k+2 are queued while the copy of snapshot k drains and the
            # host appends it, so the per-step snapshot never idles the GPU (the reference spends 0.49 s
            # per step here, particle.py:30-33).
            ring = 3
            snap_dev = [torch.empty(max(3 * n, 1), dtype=torch.float64, device=s.device) for _ in range(ring)]
            snap_host = [torch.empty(max(3 * n, 1), dtype=torch.float64).pin_memory() for _ in range(ring)]
            copied = [torch.cuda.Event() for _ in range(ring)]
            in_flight: List[int] = []

            def drain(keep: int) -> None:
                while len(in_flight) > keep:
                    slot = in_flight.pop(0)
                    copied[slot].synchronize()
                    self.displacements.append(snap_host[slot][:3 * n].numpy().copy())

            it = range(self.steps)
            if self.progress:
                it = tqdm(it)
            with torch.cuda.stream(stream):
                for k in it:
                    s.substep(1)
                    if self.save_displacements:
                        slot = k % ring
                        s.snapshot(self.tightening_coeff, snap_dev[slot])
                        snap_host[slot].copy_(snap_dev[slot], non_blocking=True)
                        copied[slot].record(stream)
                        in_flight.append(slot)
                        drain(ring - 1)      # frees the slot the next step writes
                drain(0)
                s.check_errors()
                out = s.get_particles()
                self.particles.pos[:] = out["x"].double().cpu().numpy()
                self.v[...] = out["v"].double().cpu().numpy()
                self.F[...] = out["F"].double().cpu().numpy()
                self.C[...] = out["C"].double().cpu().numpy()
            logger.info("Saving displacements")
            if not os.path.exists(self.outdir):
                os.mkdir(self.outdir)
            for i, displacement in enumerate(self.displacements):
                np.save(f"{self.outdir}/{i}", displacement)
            logger.success("Simulation done")
        except Exception as e:  # the reference's thread dies silently and leaves running = True
            self.error = e
            raise
        finally:
            self.running = False
