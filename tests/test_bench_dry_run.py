"""bench.py's control flow and JSON assembly, dry: the solver, CUDA events / streams and pinned memory are
stand-ins, everything else (argument handling, timed loop, sample spread, phase table, e2e bookkeeping, roofline,
the one JSON line) is the real code.  No number printed here means anything; the point is that a typo in the
script the round's measurement depends on fails HERE, on CPU, not on the GPU box."""
import contextlib
import json
import os
import sys

import pytest

torch = pytest.importorskip("torch")

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


class FakeEvent:
    clock = [0.0]

    def __init__(self, enable_timing=False):
        self.t = None

    def record(self, stream=None):
        FakeEvent.clock[0] += 1.0
        self.t = FakeEvent.clock[0]

    def elapsed_time(self, other):
        return other.t - self.t

    def synchronize(self):
        pass


class FakeStream:
    cuda_stream = 0

    def __init__(self, device=None):
        pass

    def wait_event(self, ev):
        assert ev.t is not None, "waited on an event that was never recorded"

    def wait_stream(self, s):
        pass

    def synchronize(self):
        pass


class FakeBuf:
    def __init__(self, cap):
        self.x = torch.zeros((3, cap)); self.v = torch.zeros((3, cap))
        self.C = torch.zeros((9, cap)); self.F = torch.zeros((9, cap))
        self.Jp = self.mass = self.mu0 = self.lam0 = self.material = None
        self.id = torch.arange(cap, dtype=torch.int32)


class FakeGraph:
    def __init__(self, solver, n):
        self.solver, self.n = solver, n

    def replay(self):
        self.solver.steps_run += self.n


class FakeLibrary:
    def ffmpm_export_state(self, h, dst, stream):
        FakeSolver.exports += 1
        return 0


class FakeSolver:
    """The surface of femflow_b200.mpm.MpmSolver that bench.py (and the host pipeline it drives) touches."""
    instances = []
    exports = 0
    lib, _h = FakeLibrary(), None

    def _stream(self, stream=None):
        return None

    def __init__(self, dim, res, dt, volume, gravity, hardening, *, capacity, device=None, **kw):
        self.dim, self.capacity, self.device, self.dtype = dim, (capacity + 63) // 64 * 64, torch.device("cpu"), torch.float32
        self.reorder = dim == 3
        self.buffers = [FakeBuf(self.capacity), FakeBuf(self.capacity)]
        self.material_layout = "table[1]"
        self.num_particles = 0
        self.steps_run = self.launches = self.binds = 0
        self._live = 0
        FakeSolver.instances.append(self)

    live = property(lambda self: self.buffers[self._live])
    live_index = property(lambda self: self._live)

    def set_particles(self, x, *a, **k):
        self.num_particles = len(x)

    def _bind(self, n, cur=0):
        self.num_particles, self._live, self.binds = n, cur, self.binds + 1

    def substep(self, n=1):
        self.steps_run += n
        self.launches += 10 * n
        if self.reorder:
            self._live ^= n & 1

    def make_graph(self, n):
        self.graph_substeps, self.graph_launches = n, 10 * n
        return FakeGraph(self, n)

    def poll_error(self):
        return 0

    def launch_count(self):
        return self.launches

    def clear_grid(self): pass
    def bin(self): pass
    def p2g(self): pass
    def grid_op(self): pass
    def g2p(self): pass


@pytest.fixture
def dry(monkeypatch):
    import femflow_b200.mpm as mpm
    monkeypatch.setattr(torch.cuda, "is_available", lambda: True)
    monkeypatch.setattr(torch.cuda, "set_device", lambda i: None)
    monkeypatch.setattr(torch.cuda, "synchronize", lambda *a, **k: None)
    monkeypatch.setattr(torch.cuda, "Event", FakeEvent)
    monkeypatch.setattr(torch.cuda, "Stream", FakeStream)
    monkeypatch.setattr(torch.cuda, "current_stream", lambda *a, **k: FakeStream())
    monkeypatch.setattr(torch.cuda, "stream", lambda s: contextlib.nullcontext())
    monkeypatch.setattr(torch.Tensor, "pin_memory", lambda self, *a, **k: self)
    monkeypatch.setattr(mpm, "MpmSolver", FakeSolver)
    monkeypatch.delenv("RANK", raising=False); monkeypatch.delenv("WORLD_SIZE", raising=False)
    FakeSolver.instances.clear()
    FakeSolver.exports = 0
    return monkeypatch


def run_bench(monkeypatch, capsys, *flags):
    import bench
    monkeypatch.setattr(sys, "argv", ["bench.py", "--workload", "3d:32:8", "--no-cpu-baseline", *flags])
    bench.main()
    out = [ln for ln in capsys.readouterr().out.splitlines() if ln.strip()]
    assert len(out) == 1, out                                   # exactly ONE line on stdout
    return json.loads(out[0])


CONTRACT = ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
            "dtype", "data", "config", "clocks", "e2e", "gpu_launches", "roofline", "cpu_baseline")


def test_default_line_carries_the_contract(dry, capsys):
    d = run_bench(dry, capsys, "--steps", "10", "--warmup", "1", "--e2e-steps", "2")
    for k in CONTRACT:
        assert k in d, k
    s = FakeSolver.instances[0]
    assert d["warmup"] == 3 and d["steps"] == 10 and d["n_gpus"] == 1          # warm-up is raised to the required 3
    assert d["gpu_launches"] == 100 and d["config"]["cuda_graph"] is False
    assert len(d["ms_per_step_samples"]) == 5
    assert d["config"]["workload"].startswith("3D elastic block") and d["config"]["parallelism"] == "single GPU"
    e = d["e2e"]
    assert e["h2d_bytes_per_step"] == e["d2h_bytes_per_step"] == 24 * 4 * s.num_particles
    # the headline e2e is the pipelined host-buffer path (12 calls, 3 slots); the blocking call is reported beside it
    assert "pipelined_error" not in e and e["steps"] == 12 and e["serial"]["steps"] == 2 and "HostSubstepPipeline" in e["api"]
    assert FakeSolver.exports == 3 + 12 and len(s.buffers) == 2 and s.buffers[0].x.shape[1] == s.capacity
    r = d["roofline"]
    assert r["bound"] == "hbm" and r["unit"] == "GB/s" and set(r["phase_ms"]) == {"clear", "bin", "p2g", "grid_op", "g2p"}
    assert r["algorithmic_bytes_per_launch"] > 0 and 0 < r["frac"]
    assert s.steps_run == 3 + 10 + 1 + 2 + 3 + 12      # warm-up, timed region, e2e warm step + e2e steps, pipeline warm + timed


def test_graph_and_serial_only_options(dry, capsys):
    d = run_bench(dry, capsys, "--steps", "20", "--warmup", "4", "--e2e-steps", "1", "--graph", "--e2e-serial-only")
    assert d["config"]["cuda_graph"] is True and d["gpu_launches"] == 200 and d["ms_per_step_samples"] == []
    assert "serial" not in d["e2e"] and d["e2e"]["steps"] == 1 and d["e2e"]["value"] > 0
    # steps not a multiple of 10: the graph request is declined, not an error
    d = run_bench(dry, capsys, "--steps", "7", "--graph")
    assert d["config"]["cuda_graph"] is False and d["gpu_launches"] == 70 and len(d["ms_per_step_samples"]) == 7


def test_2d_workload_line(dry, capsys, monkeypatch):
    import bench
    from femflow_b200 import scenes
    monkeypatch.setattr(bench, "make_scene", lambda w, seed=0, **k: scenes.elastic_block(2, 64, 16, 2, seed))
    d = run_bench(dry, capsys, "--steps", "5")
    assert set(d["roofline"]["phase_ms"]) == {"clear", "p2g", "grid_op", "g2p"} and d["e2e"]["value"] > 0


def _two_rank_worker(rank, world, port, outdir):
    """bench.main() as rank `rank` of `world`: gloo instead of NCCL, CPU tensors instead of CUDA ones, the stand-in
    library (tests/fake_abi.py) instead of the .so -- the real bench.py, SlabSolver, CudaSlab and SlabDriver code."""
    import io
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch.distributed as dist
    import fake_abi
    fake_abi.install()
    torch.cuda.Event = FakeEvent
    real_device, real_init = torch.device, dist.init_process_group

    class CpuDevice:                                            # torch.device("cuda", i) -> the CPU
        def __new__(cls, *a, **k):
            return real_device("cpu")
    torch.device = CpuDevice
    dist.init_process_group = lambda backend=None, **k: real_init("gloo", rank=rank, world_size=world)
    import femflow_b200.mpm as mpm
    real_solver = mpm.MpmSolver
    mpm.MpmSolver = lambda *a, **k: real_solver(*a, **{**k, "dtype": torch.float64})      # the stand-in's oracle arithmetic is fp64
    import femflow_b200.distributed as D
    real_slab_init = D.CudaSlab.__init__
    D.CudaSlab.__init__ = lambda self, *a, **k: real_slab_init(self, *a, **{**k, "dtype": torch.float64})
    import bench
    sys.argv = ["bench.py", "--gpus", str(world), "--workload", "3d:32:8", "--steps", "5", "--warmup", "3", "--no-cpu-baseline",
                "--e2e-steps", "1", "--margin", "2", "--drift", "0.4"]
    buf = io.StringIO()
    with contextlib.redirect_stdout(buf):
        bench.main()
    with open(os.path.join(outdir, f"rank{rank}.txt"), "w") as f:
        f.write(buf.getvalue())


def test_two_rank_line(tmp_path):
    """N = 2 as the driver launches it (one process per rank): rank 0 prints the one line, rank 1 nothing; the line
    carries the whole-job aggregate and the slab bookkeeping."""
    import socket
    import torch.multiprocessing as mp
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    mp.spawn(_two_rank_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    lines = [ln for ln in open(tmp_path / "rank0.txt").read().splitlines() if ln.strip()]
    assert len(lines) == 1 and open(tmp_path / "rank1.txt").read().strip() == ""
    d = json.loads(lines[0])
    for k in CONTRACT:
        assert k in d, k
    total = d["config"]["particles_total"]
    assert d["n_gpus"] == 2 and total == sum(d["config"]["slab_particles"]) and min(d["config"]["slab_particles"]) > 0
    assert d["config"]["slab_cells"] == [[0, 32], [32, 63]] and d["config"]["rebalanced"] == 0
    assert "2 slabs along x" in d["config"]["parallelism"] and d["scaling"] == "weak"
    # the default N > 1 workload is the COUPLED bar: it says so, and particles did cross the cut
    assert "coupled" in d["config"]["parallelism"] and "bar" in d["config"]["workload"]
    assert d["config"]["migration"]["particles_received"] > 0 and d["config"]["migration"]["rounds"] > 0
    assert d["e2e"]["h2d_bytes_per_step"] == 24 * 8 * total and d["roofline"] is None and d["config"]["n_oob"] == 0
    assert "pipelined_error" not in d["e2e"] and d["e2e"]["steps"] == 12 and d["e2e"]["serial"]["steps"] == 1


def test_cpu_baseline_leg_and_reference_arm(dry, capsys, monkeypatch):
    """The two places bench.py may execute oracle/: the cpu_baseline leg of the default run (real C port, tiny
    budget here) and `--impl reference`."""
    import bench
    from oracle import native as onative
    real = onative.time_sample
    monkeypatch.setattr(onative, "time_sample", lambda scene, budget_s=15.0, threads=None: real(scene, budget_s=0.2, threads=threads))
    monkeypatch.setattr(sys, "argv", ["bench.py", "--workload", "3d:32:8", "--steps", "5", "--e2e-steps", "1"])
    bench.main()
    d = json.loads([ln for ln in capsys.readouterr().out.splitlines() if ln.strip()][0])
    c = d["cpu_baseline"]
    assert c["kind"] == "port" and c["cores"] >= 1 and c["value"] > 0 and "particles of the workload" in c["sample"]
    monkeypatch.setattr(sys, "argv", ["bench.py", "--workload", "3d:32:8", "--impl", "reference", "--steps", "2", "--warmup", "1"])
    bench.main()
    r = json.loads([ln for ln in capsys.readouterr().out.splitlines() if ln.strip()][0])
    assert r["impl"] == "reference" and r["gpu_launches"] == 0 and r["value"] > 0 and r["metric"] == d["metric"]
    assert r["e2e"] == {"value": r["value"], "unit": r["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert r["cpu_baseline"]["kind"] == "port" and r["config"]["workload"] == d["config"]["workload"]
