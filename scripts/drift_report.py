#!/usr/bin/env python
"""Drift of the CUDA substep against the fp64 oracle over many substeps (reported, not
gated: BASELINE.json north_star "drift reported over 1000 substeps").

    python scripts/drift_report.py [--steps 1000] [--cells 24] [--res 64] [--out profiles/drift.json]

Runs the same perturbed elastic block through (a) the C oracle port (fp64, the
reference's algorithm) and (b) the CUDA path in fp32 and fp64 builds, and records the
max-norm relative error of x, v, F, C at logarithmically spaced substeps, with the
absolute floors of SURVEY 8d.  Needs a GPU; the oracle is only the checker here.
"""
import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=1000)
    ap.add_argument("--cells", type=int, default=24)
    ap.add_argument("--res", type=int, default=64)
    ap.add_argument("--out", default=os.path.join(ROOT, "profiles", "drift.json"))
    args = ap.parse_args()
    import torch
    from femflow_b200 import scenes
    from femflow_b200.mpm import MpmSolver
    from oracle import native as ON

    sc = scenes.elastic_block(3, args.res, args.cells, 2, seed=1)
    n = sc.n
    marks = sorted({1, 3, 10, 30, 100, 300, 1000, args.steps} & set(range(1, args.steps + 1)))
    report = {"scene": sc.name, "particles": n, "dt": sc.dt, "steps": args.steps, "runs": {}}
    for dtype in ("float32", "float64"):
        x, v, F, C = (a.astype(np.float64) for a in (sc.x, sc.v, sc.F, sc.C))
        m = np.full(n, sc.mass); mu = np.full(n, sc.mu_0); lam = np.full(n, sc.lambda_0)
        s = MpmSolver(3, sc.res, sc.dt, sc.volume, sc.gravity, sc.hardening, capacity=n, dtype=getattr(torch, dtype))
        s.set_particles(sc.x, sc.v, sc.F, sc.C, None, sc.mass, sc.mu_0, sc.lambda_0)
        rows = []
        for step in range(1, args.steps + 1):
            s.substep(1)
            ON.solve_mls_mpm_3d(sc.res, float(sc.res), sc.hardening, 1 / sc.res, sc.dt, sc.volume, sc.gravity,
                                x, m, mu, lam, v, F, C)
            if step in marks:
                s.check_errors()
                o = {k: t.double().cpu().numpy() for k, t in s.get_particles().items()}
                V = max(np.abs(v).max(), sc.dt * abs(sc.gravity))
                rows.append({"substep": step,
                             "x": float(np.abs(o["x"] - x).max() / max(np.abs(x).max(), 1.0)),
                             "v": float(np.abs(o["v"] - v).max() / V),
                             "F": float(np.abs(o["F"] - F).max() / max(np.abs(F).max(), 1.0)),
                             "C": float(np.abs(o["C"] - C).max() / max(np.abs(C).max(), 4 * sc.res * V))})
                print(dtype, rows[-1], flush=True)
        report["runs"][dtype] = rows
        s.close()
    os.makedirs(os.path.dirname(args.out), exist_ok=True)
    with open(args.out, "w") as f:
        json.dump(report, f, indent=1)
    print("wrote", args.out)


if __name__ == "__main__":
    main()
