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


def _worker(rank, world, port, steps, out, nmat):
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
        drv = SlabDriver(plan, local, migrate_every=2)
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
