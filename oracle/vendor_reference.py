"""Stage the reference's OWN MPM modules under ``baseline/_ref/`` so that its numba path can be timed on the
GPU box next to the CUDA path (BASELINE.json north_star: "the reference's numba CPU path is timed alongside").

TEST / BENCH INFRASTRUCTURE.  ``/root/reference`` exists only in the build container; ``baseline/_ref/`` is
git-ignored (the reference's sources never enter this repository's history) but travels with the snapshot to
the GPU box, like the built ``.so`` files.  Only the import closure of
``femflow.solvers.mpm.mls_mpm.solve_mls_mpm_3d`` is staged, unmodified, byte for byte.

    python oracle/vendor_reference.py          # no-op when /root/reference is absent
"""
from __future__ import annotations

import os
import shutil

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("FEMFLOW_REFERENCE", "/root/reference")
DST = os.path.join(HERE, "..", "baseline", "_ref")

# import closure of the MPM path (probed with sys.modules after importing mls_mpm, two_d, three_d)
FILES = [
    "femflow/__init__.py",
    "femflow/meshing/__init__.py",
    "femflow/numerics/__init__.py",
    "femflow/numerics/linear_algebra.py",
    "femflow/simulation/__init__.py",
    "femflow/solvers/__init__.py",
    "femflow/solvers/mpm/__init__.py",
    "femflow/solvers/mpm/mls_mpm.py",
    "femflow/solvers/mpm/particle.py",
    "femflow/solvers/mpm/utils.py",
    "femflow/solvers/mpm/three_d/__init__.py",
    "femflow/solvers/mpm/three_d/p2g.py",
    "femflow/solvers/mpm/three_d/grid_op.py",
    "femflow/solvers/mpm/three_d/g2p.py",
    "femflow/solvers/mpm/two_d/__init__.py",
    "femflow/solvers/mpm/two_d/p2g.py",
    "femflow/solvers/mpm/two_d/grid_op.py",
    "femflow/solvers/mpm/two_d/g2p.py",
    "femflow/utils/__init__.py",
    "LICENSE",
]


def stage(verbose: bool = False) -> bool:
    """Copy the files; returns False (and does nothing) when the reference tree is not present."""
    if not os.path.isdir(os.path.join(REF, "femflow")):
        return False
    for rel in FILES:
        src, dst = os.path.join(REF, rel), os.path.join(DST, rel)
        if not os.path.exists(src):
            continue
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        shutil.copyfile(src, dst)
        if verbose:
            print("staged", rel)
    return True


if __name__ == "__main__":
    print("staged" if stage(verbose=True) else f"reference tree not found at {REF}: nothing staged")
