"""Shared pytest plumbing: the ``gpu`` marker, repo-root imports, golden loader."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box via gpurun)")


def load_golden(name):
    with np.load(os.path.join(GOLDEN, name + ".npz")) as z:
        return {k: z[k] for k in z.files}


def dense_grid(g, name):
    """Inverse of oracle/gen_golden.py:sparse_grid."""
    shape = tuple(int(s) for s in g[f"{name}_shape"])
    out = np.zeros((int(np.prod(shape[:-1])), shape[-1]))
    out[g[f"{name}_idx"]] = g[f"{name}_val"]
    return out.reshape(shape)


def rel_err(a, ref, floor=0.0):
    """max-norm relative error with an absolute floor on the denominator
    (SURVEY 8d parity protocol)."""
    a = np.asarray(a, dtype=np.float64)
    ref = np.asarray(ref, dtype=np.float64)
    if a.size == 0:
        return 0.0
    return float(np.max(np.abs(a - ref)) / max(float(np.max(np.abs(ref))), floor, 1e-300))


@pytest.fixture(scope="session")
def golden():
    return load_golden
