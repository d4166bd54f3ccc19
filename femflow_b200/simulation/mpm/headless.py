"""Headless runner for the reference's drop scenes (``paper_1.multi_drop_experiment``, reference
paper_1.py:17-144): same scene construction and simulation parameters, no window -- the reference can only
run these through its OpenGL GUI (paper_1.py:147-196).

    python -m femflow_b200.simulation.mpm.headless --experiment 0 --outdir tmp            # 3500 substeps, 35 321 particles
    python -m femflow_b200.simulation.mpm.headless --experiment 1 --steps 200 --no-save

Writes ``{outdir}/{i}.npy`` exactly as ``MPMSimulation`` does (simulation.py:143-147).
"""
from __future__ import annotations

import argparse
import time
from typing import List, Tuple

import numpy as np

from . import primitives as P


class PointMesh:
    """The part of ``femflow.viz.mesh.Mesh`` the MPM path touches (viz/mesh.py:15-35, 160-167): a flat
    float32 vertex vector and in-place translations.  (The reference class also carries faces, colours and
    igl-based helpers that the solver never reads.)"""

    def __init__(self, points: np.ndarray):
        self.vertices = np.asarray(points, dtype=np.float64).reshape(-1).astype(np.float32)

    def translate_x(self, amount: float) -> None:
        self.vertices[0::3] += amount

    def translate_y(self, amount: float) -> None:
        self.vertices[1::3] += amount

    def translate_z(self, amount: float) -> None:
        self.vertices[2::3] += amount


# paper_1.py:75-144: what differs between the two experiments; everything else is shared (paper_1.py:75-86)
EXPERIMENTS = {
    0: dict(steps=3500, mesh_res=30, tightening_coeff=0.05, collider_mass=10.0, collider_E=1000, gyroid_E=140),
    1: dict(steps=1000, mesh_res=40, tightening_coeff=0.10, collider_mass=10.0, collider_E=1000, gyroid_E=140),
}
SHARED = dict(outdir="tmp", dt=1e-4, gyroid_mass=1.0, k=0.2, t=0.3, volume=1.0, force=-9.8, gyroid_v=0.2, collider_v=0.4,
              hardening=0.7, grid_res=64, mesh_type="gyroid")


def multi_drop_scene(experiment: int, device=None) -> Tuple[List[PointMesh], List[Tuple[float, float, float]], dict]:
    """Meshes, per-mesh ``(mass, E, nu)`` and the ``MPMSimulation`` constructor arguments of
    ``multi_drop_experiment(experiment)`` (paper_1.py:17-73 ``prepare_sim``).  ``device``: generate the two point
    sets with the CUDA scene kernels (``primitives.generate_*_gpu``) instead of on the host."""
    if experiment not in EXPERIMENTS:
        raise ValueError(f"Experiment {experiment} is not valid")            # paper_1.py:144
    e, s = EXPERIMENTS[experiment], SHARED
    if device is not None:
        gyroid = PointMesh(P.generate_implicit_points_gpu(s["mesh_type"], s["k"], s["t"], e["mesh_res"], device).cpu().numpy())
    else:
        gyroid = PointMesh(P.generate_implicit_points(s["mesh_type"], s["k"], s["t"], e["mesh_res"]))
    gyroid.translate_y(0.1)
    v = gyroid.vertices.reshape(-1, 3)
    lo, hi = v.min(0), v.max(0)
    if device is not None:
        collider = PointMesh(P.generate_cube_points_gpu((lo[0], hi[0]), (lo[1], hi[1]), (lo[2], hi[2]), e["mesh_res"], device).cpu().numpy())
    else:
        collider = PointMesh(P.generate_cube_points((lo[0], hi[0]), (lo[1], hi[1]), (lo[2], hi[2]), e["mesh_res"]))
    collider.translate_y(3)
    params = [(s["gyroid_mass"], e["gyroid_E"], s["gyroid_v"]), (e["collider_mass"], e["collider_E"], s["collider_v"])]
    ctor = dict(outdir=s["outdir"], steps=e["steps"], dt=s["dt"], gyroid_mass=s["gyroid_mass"], collider_mass=e["collider_mass"],
                volume=s["volume"], force=s["force"], gyroid_E=e["gyroid_E"], collider_E=e["collider_E"],
                gyroid_v=s["gyroid_v"], collider_v=s["collider_v"], hardening=s["hardening"], grid_res=s["grid_res"],
                tightening_coeff=e["tightening_coeff"])
    return [gyroid, collider], params, ctor


def run(experiment: int = 0, steps: int = None, outdir: str = None, save: bool = True, device: str = "cuda:0",
        progress: bool = False):
    """Build the scene, run it on the GPU, return the finished ``MPMSimulation`` and the wall time of the run."""
    from .simulation import MPMSimulation
    meshes, params, ctor = multi_drop_scene(experiment, device=device if str(device).startswith("cuda") else None)
    if steps is not None:
        ctor["steps"] = int(steps)
    if outdir is not None:
        ctor["outdir"] = outdir
    sim = MPMSimulation(save_displacements=save, device=device, progress=progress, **ctor)
    sim.load(meshes=meshes, params=params)
    t0 = time.time()
    sim.start()
    sim.join()
    if sim.error is not None:
        raise sim.error
    return sim, time.time() - t0


def main(argv=None) -> None:
    ap = argparse.ArgumentParser(description=__doc__.split("\n\n")[0])
    ap.add_argument("--experiment", type=int, default=0, choices=sorted(EXPERIMENTS))
    ap.add_argument("--steps", type=int, default=None, help="substeps (default: the experiment's own count)")
    ap.add_argument("--outdir", default=None)
    ap.add_argument("--no-save", action="store_true", help="skip the per-step displacement snapshots and the .npy dump")
    ap.add_argument("--device", default="cuda:0")
    args = ap.parse_args(argv)
    sim, seconds = run(args.experiment, args.steps, args.outdir, not args.no_save, args.device, progress=True)
    n = len(sim.particles)
    print(f"experiment {args.experiment}: {n} particles, {sim.steps} substeps in {seconds:.2f} s "
          f"({n * sim.steps / seconds:.3g} particle-substeps/s including snapshots), "
          f"{len(sim.displacements)} displacement vectors" + (f" in {sim.outdir}/" if not args.no_save else ""))


if __name__ == "__main__":
    main()
