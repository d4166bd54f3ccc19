#!/usr/bin/env python
"""Benchmark of the MPM substep (bin + P2G + grid update + G2P).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload 3d16m|2d1m] [--impl reference]

Prints ONE JSON line (rank 0).  ``value`` = particle-substeps/s with the state
resident in HBM; ``e2e`` = the same through host buffers (pinned H2D of the full
particle state + one substep + D2H of x, v, C, F every step); ``roofline`` = the
dominant kernel against the measured HBM peak; ``cpu_baseline`` = the oracle port
timed on this box's host cores on a bounded sample of the same workload.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "particle-substeps/sec (P2G+grid+G2P)"
UNIT = "particle-substeps/s"


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def make_scene(workload: str, seed: int = 0, cells_x=None, res_x=None):
    from femflow_b200 import scenes
    if workload == "2d1m":
        return scenes.config_2d_1m(seed)
    if workload in ("3d16m", "3d16m-blocks"):
        return scenes.config_3d_16m(seed)
    if workload == "3d16m-rest":          # same block as the reference starts it: F = I, v = 0 (simulation.py:76-79)
        sc = scenes.elastic_block(3, 256, 128, 2, seed, perturb=False)
        sc.name += " (at rest)"
        return sc
    if workload.startswith("3d16m-drift:"):   # the headline block moving along +x at <c> cells per substep: every substep a
        c = float(workload.split(":")[1])      # fraction ~c of the particles changes cell (P2G run fragmentation, re-binning)
        sc = scenes.config_3d_16m(seed)
        sc.v[:, 0] += np.float32(c / sc.res / sc.dt)
        sc.name += f" drifting {c} cells/substep"
        return sc
    if workload.startswith("3d:"):        # 3d:<res>:<cells>  (tests / quick runs)
        _, res, cells = workload.split(":")
        return scenes.elastic_block(3, int(res), int(cells), 2, seed)
    raise SystemExit(f"unknown workload {workload}")


# ----------------------------------------------------------------------------- #
class ClockSampler:
    """SM clock and clock-event (throttle) reasons sampled DURING the timed region.

    NVML in-process (pynvml), one sample every ~2 ms from a thread: the driver's line times 20 substeps = 27 ms, which
    an ``nvidia-smi -lms 50`` child (hundreds of ms to start, 50 ms period) cannot resolve -- it is kept as the fallback
    where pynvml is missing, started before the warm-up.  ``mark_begin`` / ``mark_end`` bracket the timed region on
    the host clock; only samples between them count (``window`` says so; if the region was too short for a single
    sample, the ones taken in the 150 ms after it -- clocks do not drop that fast -- are reported and labelled)."""

    REASONS = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap",
               0x80: "hw_power_brake_slowdown"}
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int, uuid=None):
        self.rows = []            # (t, sm_mhz, reasons bitmask)
        self.proc = None
        self.gpu, self.uuid = gpu_index, uuid
        self.nvml = None
        self.max_mhz = None
        self._stop = threading.Event()
        self.t0 = self.t1 = None
        self.source = None

    def _nvml_handle(self):
        import pynvml
        pynvml.nvmlInit()
        h = None
        if self.uuid:
            for cand in (f"GPU-{self.uuid}", str(self.uuid)):
                try:
                    h = pynvml.nvmlDeviceGetHandleByUUID(cand if isinstance(cand, bytes) else cand.encode())
                    break
                except Exception:
                    h = None
        if h is None:
            h = pynvml.nvmlDeviceGetHandleByIndex(self.gpu)
        return pynvml, h

    def start(self):
        try:
            pynvml, h = self._nvml_handle()
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(h, pynvml.NVML_CLOCK_SM))
            reasons_fn = getattr(pynvml, "nvmlDeviceGetCurrentClocksEventReasons", None) or \
                pynvml.nvmlDeviceGetCurrentClocksThrottleReasons

            def poll():
                while not self._stop.is_set():
                    try:
                        self.rows.append((time.perf_counter(), float(pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM)),
                                          int(reasons_fn(h))))
                    except Exception:
                        pass
                    self._stop.wait(0.002)
            self.nvml = pynvml
            self.source = "nvml"
            self.t = threading.Thread(target=poll, daemon=True)
            self.t.start()
            return
        except Exception:
            self.nvml = None
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "20",
                 "-i", str(self.gpu)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.source = "nvidia-smi"
            self.t = threading.Thread(target=self._read_smi, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read_smi(self):
        bits = [0x8, 0x40, 0x20, 0x4]
        for line in self.proc.stdout:
            r = [c.strip() for c in line.split(",")]
            try:
                mask = sum(b for b, val in zip(bits, r[5:9]) if val.lower().startswith("active"))
                self.rows.append((time.perf_counter(), float(r[1]), mask))
                self.max_mhz = max(self.max_mhz or 0.0, float(r[2]))
            except Exception:
                pass

    def mark_begin(self):
        self.t0 = time.perf_counter()

    def mark_end(self):
        self.t1 = time.perf_counter()

    def stop(self):
        if self.nvml is None and self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "samples": 0, "reasons": ["clock sampling unavailable (no pynvml, no nvidia-smi)"]}
        time.sleep(0.15)
        self._stop.set()
        if self.proc is not None:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                self.proc.kill()
        t0 = self.t0 if self.t0 is not None else -float("inf")
        t1 = self.t1 if self.t1 is not None else float("inf")
        inside = [r for r in self.rows if t0 <= r[0] <= t1]
        window = "timed region"
        if not inside:
            inside = [r for r in self.rows if t1 < r[0] <= t1 + 0.15] or self.rows[-3:]
            window = "no sample fell inside the timed region: the 150 ms after it"
        mask = 0
        for r in inside:
            mask |= r[2]
        sm = [r[1] for r in inside]
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_min_mhz": min(sm) if sm else None, "sm_max_mhz": self.max_mhz,
                "samples": len(sm), "reasons": sorted(nm for b, nm in self.REASONS.items() if mask & b),
                "window": window, "source": self.source,
                "region_ms_host": None if self.t0 is None or self.t1 is None else round((self.t1 - self.t0) * 1e3, 2)}


# ----------------------------------------------------------------------------- #
def cpu_baseline(scene, budget_s: float = 15.0, threads=None):
    """Oracle port on the host cores, bounded sample of the same workload: the first
    ``m`` cells-worth of the block (same density, same grid resolution)."""
    from oracle import native as onative
    return onative.time_sample(scene, budget_s=budget_s, threads=threads)


def parity_block(solver, scene):
    """One more substep of the TIMED workload, from the state the timed region left behind, on the GPU and by the
    C port of the reference loops (oracle/mpm_oracle.c) started from the same downloaded state: max-norm errors
    relative to the floors of SURVEY 8d.  The checker, never the thing measured."""
    from oracle import native as onative
    import torch
    d, res, n = scene.dim, scene.res, solver.num_particles
    st = solver.get_particles()
    x, v, F, C = (st[k].double().cpu().numpy() for k in ("x", "v", "F", "C"))
    Jp = st["Jp"].double().cpu().numpy() if "Jp" in st else np.ones((n, 1))
    del st
    solver.substep(1)
    n_oob = solver.poll_error()
    onative.lib().oracle_set_threads(onative.host_threads())
    t0 = time.perf_counter()
    if d == 3:
        mass = np.full(n, scene.mass); mu = np.full(n, scene.mu_0); lam = np.full(n, scene.lambda_0)
        gv, _ = onative.solve_mls_mpm_3d(res, float(res), scene.hardening, 1.0 / res, scene.dt, scene.volume, scene.gravity,
                                         x, mass, mu, lam, v, F, C)
    else:
        gv, _ = onative.solve_mls_mpm_2d(res, float(res), scene.hardening, scene.mu_0, scene.lambda_0, scene.mass, 1.0 / res,
                                         scene.dt, scene.volume, scene.gravity, x, v, F, C, Jp)
    cpu_s = time.perf_counter() - t0
    V = max(float(np.abs(gv).max()), scene.dt * abs(scene.gravity))
    del gv
    out = solver.get_particles()

    def err(name, ref, floor):
        worst, chunk = 0.0, 1 << 22
        flat = ref.reshape(n, -1)
        for a in range(0, n, chunk):
            got = out[name][a:a + chunk].double().cpu().numpy().reshape(-1, flat.shape[1])
            worst = max(worst, float(np.abs(got - flat[a:a + chunk]).max()))
        return worst / floor
    e = {"x": err("x", x, 1.0), "v": err("v", v, V), "F": err("F", F, 1.0), "C": err("C", C, 4 * res * V)}
    if d == 2:
        e["Jp"] = err("Jp", Jp, 1.0)
    return {"against": "C port of the reference loops (oracle/mpm_oracle.c, fp64), one substep from the downloaded state of the "
                       "timed workload, all particles", "particles": int(n), "max_norm_rel_err": e, "tolerance": 1e-5,
            "within_tolerance": bool(max(e.values()) < 1e-5), "n_oob": int(n_oob), "port_seconds": round(cpu_s, 2),
            "floors": "x, F: 1; v: V = max(|v_grid,ref|, dt |g|); C: 4 inv_dx V (SURVEY 8d)"}


def numba_baseline(timeout_s: float = 240.0):
    """The UNMODIFIED reference's numba substep timed on this box (oracle/time_numba_reference.py, a subprocess: numba's
    JIT stays out of this process).  None when baseline/_ref or numba is not there."""
    try:
        proc = subprocess.run([sys.executable, os.path.join(ROOT, "oracle", "time_numba_reference.py")], capture_output=True,
                              text=True, timeout=timeout_s, env={**os.environ, "CUDA_VISIBLE_DEVICES": ""})
        line = [ln for ln in proc.stdout.splitlines() if ln.startswith("{")][-1]
        return json.loads(line)
    except Exception as e:
        return {"unavailable": f"{type(e).__name__}: {e}"}


def run_reference(args):
    """--impl reference: the reference's CPU algorithm on all host threads.  The reference is
    pure Python/numba and cannot travel to the GPU box, so this is the C port of its loops
    (oracle/mpm_oracle.c, kind "port").  Every step is one substep over a fixed, bounded sample
    of the workload (first m particles, full grid), m sized so that warmup + steps take ~90 s."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    scene = make_scene(args.workload)
    from oracle import native as onative
    total = max(1, args.steps + args.warmup)
    probe = onative.SampleRunner(scene, min(scene.n, 1_000_000))
    probe.step()
    t1 = probe.step()
    rate = probe.m / t1
    m = int(min(scene.n, max(1_000_000, rate * 90.0 / total)))
    run = onative.SampleRunner(scene, m)
    times = [run.step() for _ in range(total)][args.warmup:]
    sec = float(np.mean(times))
    v = run.m / sec
    sample = (f"first {run.m} particles of the workload (same density, full {scene.res}^{scene.dim} grid), one substep "
              f"per step; C port of the reference loops, OpenMP over particles on {run.cores} threads "
              f"(the reference's own numba path is serial)")
    line = {
        "impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * sec, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": scene.name, "sample": sample},
        "cpu_baseline": {"value": v, "unit": UNIT, "cores": int(run.cores), "kind": "port", "sample": sample},
        "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


def e2e_pipelined(torch, core, host_in, host_out, n, steps, depth=3, step=None):
    """The host-buffer path as a THROUGHPUT pipeline over independent calls (femflow_b200.host_pipeline: a parameter
    sweep, the scenes of paper_1.multi_drop_experiment): `depth` device slots in flight, the upload of call k+1 under the
    download of call k on the full-duplex link.  Every call still uploads its whole state from pinned memory, runs
    one substep and downloads x, v, C, F -- in the caller's particle order.  Timed on the device."""
    from femflow_b200.host_pipeline import HostSubstepPipeline
    pipe = HostSubstepPipeline(core, depth=depth, step=step)
    outs = [host_out] + [{k: torch.empty_like(t).pin_memory() for k, t in host_out.items()} for _ in range(depth - 1)]
    run = torch.cuda.current_stream(core.device)
    try:
        for k in range(depth):
            pipe.submit(host_in, outs[k % depth])          # warm every slot
        pipe.drain()
        torch.cuda.synchronize(core.device)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(run)
        for k in range(steps):
            pipe.submit(host_in, outs[k % depth])
        for ev in pipe.drained:
            run.wait_event(ev)
        e1.record(run)
        torch.cuda.synchronize(core.device)
        ms = e0.elapsed_time(e1)
    finally:
        pipe.drain()
    return ms, steps


# ----------------------------------------------------------------------------- #
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--workload", default="3d16m")
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-parity", action="store_true", help="skip the parity block (one substep of the timed workload against the C port)")
    ap.add_argument("--no-numba", action="store_true", help="skip timing the reference's own numba path (baseline/_ref, ~60 s)")
    ap.add_argument("--p2g-mode", default="auto")
    ap.add_argument("--e2e-steps", type=int, default=3)
    ap.add_argument("--material", default="auto", choices=["auto", "planes"],
                    help="per-particle material layout (N=1): auto = table/rows when <= 256 distinct triples, planes = 3 scalar planes")
    ap.add_argument("--slab-timing", action="store_true", help="N>1: print per-phase CUDA-event times per rank to stderr")
    ap.add_argument("--margin", type=int, default=4, help="slab halo margin in cells = substeps between migrations")
    ap.add_argument("--drift", type=float, default=0.02,
                    help="N>1 coupled bar: drift along x in cells per substep (0.02 = 26 m/s, a quarter of the material's wave speed "
                         "at the CFL-scaled dt of the block; 0.1 is the stress case kept in profiles/)")
    ap.add_argument("--e2e-serial-only", action="store_true",
                    help="report the blocking host-buffer call as e2e.value instead of the pipelined one (round-1 behaviour)")
    ap.add_argument("--halo", default="symm", choices=["p2p", "symm"],
                    help="N>1: halo planes by one-sided puts into the neighbour's symmetric-memory inbox over NVLink (symm, "
                         "default; falls back to p2p when torch's symmetric memory is unavailable) or by NCCL send/recv (p2p)")
    ap.add_argument("--graph", action="store_true",
                    help="N=1: replay the timed substeps as CUDA graphs of 10 substeps (MpmSolver.make_graph); "
                         "measured -11 %% on 2d1m, -4 %% on a 2M-particle 3D block (profiles/r01q_graph_experiment.json)")
    ap.add_argument("--rebalance", action="store_true",
                    help="dam workloads, N>1: re-cut the slabs by particle count before the warm-up (SlabDriver.rebalance)")
    ap.add_argument("--presteps", type=int, default=0,
                    help="dam workloads: substeps run before the warm-up (untimed) so that a later window of the collapse "
                         "is measured (SURVEY 8d C5: tall column vs pancake); re-cut every --rebalance-every of them")
    ap.add_argument("--rebalance-every", type=int, default=0,
                    help="with --rebalance: also offer a re-cut every K timed substeps (inside the timed region)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup

    if args.impl == "reference":
        return run_reference(args)

    import torch
    import torch.distributed as dist

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback for the product path)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        # stdout carries exactly one JSON line: NCCL's own banner / debug log (NCCL_DEBUG=VERSION prints
        # "NCCL version ..." to stdout at communicator creation) goes to stderr instead
        if os.environ.get("NCCL_DEBUG", "").upper() == "VERSION":
            os.environ["NCCL_DEBUG"] = "WARN"
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
        dist.init_process_group("nccl", device_id=dev)

    from femflow_b200.mpm import MpmSolver
    from femflow_b200._numa import bind_to_gpu
    # pinned host buffers (e2e) are allocated from the CPUs next to this rank's GPU; the CPU-baseline leg gets every core back
    affinity0 = os.sched_getaffinity(0) if hasattr(os, "sched_getaffinity") else None
    numa_note = bind_to_gpu(local_rank)

    dam = args.workload.startswith("dam")
    scene = make_scene("3d:256:8" if dam else args.workload, seed=rank)
    n = scene.n
    if dam:
        # BASELINE configs[4] (report only): --workload dam32m, or dam:<n_total> for smaller runs
        from femflow_b200.distributed import SlabSolver
        late = args.workload.endswith("-late")           # dam32m-late / dam:<n>-late: the settled state as initial condition
        wl = args.workload[:-5] if late else args.workload
        n_total_req = 33_554_432 if wl == "dam32m" else int(wl.split(":")[1])
        if world == 1:
            import torch.distributed as dist_  # noqa: F401  (single rank: the driver never communicates)
        solver = SlabSolver.from_dam_break(rank, world, dev, n_total=n_total_req, margin=args.margin,
                                           p2g_mode=args.p2g_mode, halo=args.halo, late=late)
        scene.name = solver.scene_name
        scene.dt = solver.local.solver.cfg.dt
        if args.rebalance and world > 1:
            solver.rebalance()
        n = solver.num_particles
    elif world > 1:
        from femflow_b200.distributed import SlabSolver
        if args.workload == "3d16m-blocks":
            # round-1 weak scaling: one independent block in the middle of every slab (empty halos, no migration)
            solver = SlabSolver.from_scene(scene, rank, world, dev, p2g_mode=args.p2g_mode, margin=args.margin,
                                           halo=args.halo)
            coupling = "independent block per slab (halo planes carry zeros, nothing migrates)"
        else:
            # BASELINE configs[3]: ONE bar through all slabs, drifting along x: non-zero halos, migration every period
            solver = SlabSolver.from_bar(scene, rank, world, dev, p2g_mode=args.p2g_mode, margin=args.margin,
                                         halo=args.halo, drift_cells_per_substep=args.drift)
            scene.name = solver.scene_name
            coupling = (f"coupled: one bar through all slabs drifting {args.drift} cells/substep "
                        "(halo planes carry mass and momentum, a particle layer migrates every period)")
        n = solver.num_particles
    else:
        solver = MpmSolver(scene.dim, scene.res, scene.dt, scene.volume, scene.gravity, scene.hardening,
                           capacity=n, device=dev, mass=scene.mass, mu_0=scene.mu_0, lambda_0=scene.lambda_0,
                           p2g_mode=args.p2g_mode,
                           per_particle_material=(True if args.material == "planes" and scene.dim == 3 else None))
        solver.set_particles(scene.x, scene.v, scene.F, scene.C, None,
                             *( (scene.mass, scene.mu_0, scene.lambda_0) if scene.dim == 3 else (None, None, None)))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    if dam and args.presteps > 0:
        for k in range(args.presteps):
            solver.substep(1)
            if world > 1 and args.rebalance and args.rebalance_every and (k + 1) % args.rebalance_every == 0:
                solver.rebalance()
        if solver.poll_error():
            core_ = solver.local.solver
            xs = core_.live.x[:, :core_.num_particles]
            print(f"[rank {rank}] pre-steps left particles outside the LOCAL grid: n = {core_.num_particles}, x range "
                  f"{[(float(xs[c].min()), float(xs[c].max())) for c in range(3)]}, local node planes "
                  f"[{solver.plan.g_lo}, {solver.plan.g_hi}), owned cells [{solver.plan.own_lo}, {solver.plan.own_hi}), "
                  f"non-finite: {int((~torch.isfinite(xs)).sum())}", file=sys.stderr, flush=True)
            raise SystemExit("pre-steps left particles outside the grid")

    # ---- warm-up ----
    try:
        gpu_uuid = getattr(torch.cuda.get_device_properties(dev), "uuid", None)
    except Exception:
        gpu_uuid = None
    sampler = ClockSampler(local_rank, uuid=gpu_uuid)
    if rank == 0:
        sampler.start()          # before the warm-up: the sampler is up and streaming when the timed region begins
    for _ in range(args.warmup):
        solver.substep(1)
    if world > 1:
        # one migration round inside the warm-up: its first use sets up the NCCL point-to-point channels and allocates the
        # message buffers (0.5 s once), and with the slack-based cadence the first real round may come 50 substeps later
        solver.driver.migrate(hint=None)
        solver.driver._pending = None
    barrier()
    if solver.poll_error():
        raise SystemExit("warm-up left particles outside the grid")

    # ---- timed region: K substeps, state resident in HBM ----
    graph = None
    if scene.dim == 2 and world == 1 and args.steps % 10 == 0:
        args.graph = True      # a 2D substep is four kernels of ~15 us: replayed as CUDA graphs by default (-11 %)
    if args.graph and world == 1 and not dam and args.steps % 10 == 0:
        graph = solver.make_graph(10)         # captured from the warmed-up (pre-binned) state; does not execute
    l0 = solver.launch_count()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    sampler.mark_begin()
    ev0.record()
    if graph is not None:
        for _ in range(args.steps // 10):
            graph.replay()
    else:
        # SURVEY 8d asks for the spread as well: an event every fifth of the region (recording one costs nothing
        # on the stream) gives five per-substep samples next to the total
        marks, chunk = [], max(1, args.steps // 5)
        for k in range(args.steps):
            solver.substep(1)
            if dam and world > 1 and args.rebalance and args.rebalance_every and (k + 1) % args.rebalance_every == 0:
                solver.rebalance()
            if (k + 1) % chunk == 0 and k + 1 < args.steps:
                marks.append((k + 1, torch.cuda.Event(enable_timing=True)))
                marks[-1][1].record()
    ev1.record()
    barrier()
    sampler.mark_end()
    ms = ev0.elapsed_time(ev1)
    samples = []
    if graph is None:
        try:
            prev_k, prev_ev = 0, ev0
            for k_done, ev in marks + [(args.steps, ev1)]:
                samples.append(prev_ev.elapsed_time(ev) / (k_done - prev_k))
                prev_k, prev_ev = k_done, ev
        except Exception:  # an extra: never at the expense of the line itself
            samples = []
    launches = solver.launch_count() - l0 if graph is None else solver.graph_launches * (args.steps // 10)
    if world > 1:
        t = torch.tensor([ms], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
        drv = solver.driver
        cnt = torch.tensor([solver.num_particles, drv.migrated, getattr(drv, "migrate_rounds", 0),
                            getattr(drv, "migrate_overflow", 0)], device=dev, dtype=torch.int64)
        per_rank = [torch.zeros_like(cnt) for _ in range(world)]
        dist.all_gather(per_rank, cnt)
        slab_counts = [int(t[0].item()) for t in per_rank]
        n_total = sum(slab_counts)
        mig_stats = {"particles_received": sum(int(t[1].item()) for t in per_rank),       # whole job, timed region + warm-up
                     "rounds": max(int(t[2].item()) for t in per_rank),
                     "overflow": sum(int(t[3].item()) for t in per_rank)}
    else:
        n_total = n
    clocks = sampler.stop() if rank == 0 else None
    if world > 1 and args.slab_timing:
        barrier()                        # all ranks enter the instrumented loop together (rank 0 stopped the clock sampler first)
        solver.driver.enable_timing()
        for _ in range(20):
            solver.substep(1)
        tm = solver.driver.collect_timing()
        print(f"[rank {rank}] slab phase ms/substep: " + ", ".join(f"{k} {v / 20:.3f}" for k, v in tm.items()),
              file=sys.stderr, flush=True)
    n_oob = solver.poll_error()
    value = n_total * args.steps / (ms * 1e-3)

    # ---- per-phase timing (same stream, CUDA events) for the roofline of the dominant kernel ----
    phases = {}
    if world == 1 and not dam:
        names = (["bin"] if solver.reorder else []) + ["p2g", "grid_op", "g2p"]
        acc = {k: 0.0 for k in ["clear"] + names}
        reps = max(3, min(args.steps, 10))
        for _ in range(reps):
            evs = [torch.cuda.Event(enable_timing=True) for _ in range(len(names) + 2)]
            evs[0].record()
            solver.clear_grid()
            evs[1].record()
            for i, nm in enumerate(names):
                getattr(solver, nm)()
                evs[i + 2].record()
            torch.cuda.synchronize(dev)
            acc["clear"] += evs[0].elapsed_time(evs[1])
            for i, nm in enumerate(names):
                acc[nm] += evs[i + 1].elapsed_time(evs[i + 2])
        phases = {k: v / reps for k, v in acc.items()}

    # ---- parity of the timed workload (N = 1): one more substep against the C port ----
    parity = None
    if world == 1 and not dam and not args.no_parity:
        try:
            parity = parity_block(solver, scene)
        except Exception as ex:  # the checker never takes the measurement down with it
            parity = {"error": repr(ex)}

    # ---- end to end through host buffers (pinned), every step: H2D state, substep, D2H result ----
    e2e = None
    if not dam:
        d = scene.dim
        core = solver if world == 1 else solver.local.solver      # the MpmSolver that owns the buffers
        n = core.num_particles
        b = core.buffers[0]
        host_in = {k: torch.empty_like(getattr(b, k)[..., :n], device="cpu").pin_memory()
                   for k in ("x", "v", "C", "F") }
        for k in ("mass", "mu0", "lam0", "material"):
            if getattr(b, k) is not None:
                host_in[k] = torch.empty(n, dtype=getattr(b, k).dtype).pin_memory()
        live = core.live
        for k in host_in:
            host_in[k].copy_(getattr(live, k)[..., :n])
        ids0 = live.id[:n].clone() if live.id is not None else None
        host_out = {k: torch.empty_like(host_in[k]).pin_memory() for k in ("x", "v", "C", "F")}
        h2d = sum(t.numel() * t.element_size() for t in host_in.values())
        d2h = sum(t.numel() * t.element_size() for t in host_out.values())

        def e2e_step():
            bb = core.buffers[0]
            for k, t in host_in.items():
                getattr(bb, k)[..., :n].copy_(t, non_blocking=True)
            if bb.id is not None:
                bb.id[:n] = ids0
            core._bind(n)
            solver.substep(1)
            lv = core.live
            m = min(core.num_particles, n)       # N > 1: a migration round inside the step may have changed the count
            for k, t in host_out.items():
                t[..., :m].copy_(getattr(lv, k)[..., :m], non_blocking=True)

        def reduce_ms(ms_):
            if world > 1:
                t_ = torch.tensor([ms_], device=dev, dtype=torch.float64)
                dist.all_reduce(t_, op=dist.ReduceOp.MAX)
                return float(t_.item())
            return ms_

        if world > 1:
            # every call is one independent substep from the uploaded state: nothing strays further than the halo
            # margin, so the migration cadence of the resident run is held off while the host-buffer path is timed
            hold = (solver.driver.migrate_every, solver.driver._pending)
            solver.driver.migrate_every, solver.driver._pending = 1 << 60, None
        e2e_step()
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(args.e2e_steps):
            e2e_step()
        e1.record()
        barrier()
        serial_ms = reduce_ms(e0.elapsed_time(e1))
        if world > 1:       # bytes of the whole job: the ranks' particle counts differ
            t = torch.tensor([h2d, d2h], device=dev, dtype=torch.int64)
            dist.all_reduce(t, op=dist.ReduceOp.SUM)
            h2d, d2h = int(t[0].item()), int(t[1].item())
        serial = {"value": n_total * args.e2e_steps / (serial_ms * 1e-3), "unit": UNIT, "ms_per_step": serial_ms / args.e2e_steps,
                  "steps": args.e2e_steps,
                  "what": "one blocking call at a time: H2D of the state, one substep, D2H of x, v, C, F (storage order)"}
        e2e = {**serial, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
               "api": "ffmpm C ABI via MpmSolver (host SoA pinned buffers in, x/v/C/F out, every step)"}
        if not args.e2e_serial_only and scene.dim == 3:
            try:
                p_steps = max(12, 4 * args.e2e_steps)
                step_fn = (lambda: solver.driver.substep(1)) if world > 1 else None
                from femflow_b200.host_pipeline import HostSubstepPipeline
                barrier()
                ms_p, p_steps = e2e_pipelined(torch, core, host_in, host_out, n, p_steps, step=step_fn)
                barrier()
                ms_p = reduce_ms(ms_p)
                e2e = {"value": n_total * p_steps / (ms_p * 1e-3), "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                       "ms_per_step": ms_p / p_steps, "steps": p_steps,
                       "api": "femflow_b200.host_pipeline.HostSubstepPipeline over the ffmpm C ABI: every call uploads its whole "
                              "state from pinned host SoA buffers, runs one substep and downloads x, v, C, F in the caller's "
                              "particle order; 3 device slots in flight, so the upload of call k+1 runs under the download of "
                              "call k (independent calls; a dependent chain is the `serial` figure)",
                       "serial": serial}
            except Exception as ex:  # the blocking figure stands
                e2e["pipelined_error"] = repr(ex)
        if world > 1:
            solver.driver.migrate_every, solver.driver._pending = hold[0], None

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant kernel ----
    peak, peak_src = peaks()
    roofline = None
    traffic = {}
    tpath = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tpath):
        with open(tpath) as f:
            traffic = json.load(f).get(args.workload, {})
    if phases:
        N_, A = n, scene.active_nodes
        if scene.dim == 3:
            core_ = solver if world == 1 else solver.local.solver
            b0 = core_.buffers[0]
            # per-particle material bytes P2G reads: 12 as planes, 1 as table rows, 0 for a one-row table
            mat_b = 12 if b0.mass is not None else (1 if b0.material is not None else 0)
            alg = {"p2g": (96 + mat_b) * N_ + 16 * A, "g2p": (48 + 96) * N_ + 12 * A, "grid_op": 28 * A, "clear": 16 * A}
        else:
            alg = {"p2g": 48 * N_ + 12 * A, "g2p": (28 + 52) * N_ + 8 * A, "grid_op": 20 * A, "clear": 12 * A}
        dom = max((k for k in phases if k in alg), key=lambda k: phases[k])
        achieved = alg[dom] / (phases[dom] * 1e-3) / 1e9
        total_alg = ((240 + mat_b) * N_ + 72 * A) if scene.dim == 3 else (128 * N_ + 52 * A)
        roofline = {"bound": "hbm", "kernel": dom, "achieved": achieved, "peak": peak, "unit": "GB/s",
                    "frac": achieved / peak, "traffic": traffic.get(dom), "peak_source": peak_src,
                    "traffic_source": traffic.get("source") if traffic.get(dom) else None,
                    "algorithmic_bytes_per_launch": alg[dom],
                    "phase_ms": {k: round(v, 4) for k, v in phases.items()},
                    # every kernel of the step against the same peak (SURVEY 8d also asks for the nominal 8 TB/s)
                    "kernels": {k: {"algorithmic_bytes": alg[k], "achieved": alg[k] / (phases[k] * 1e-3) / 1e9,
                                    "frac": alg[k] / (phases[k] * 1e-3) / 1e9 / peak, "traffic": traffic.get(k)}
                                for k in phases if k in alg and phases[k] > 0},
                    "frac_of_nominal_8TBs": achieved / 8000.0,
                    "substep_frac_of_hbm_roofline": (total_alg / (ms * 1e-3 / args.steps) / 1e9) / peak}

    if affinity0 is not None:
        try:
            os.sched_setaffinity(0, affinity0)
        except Exception:
            pass
    cpu = None
    if not args.no_cpu_baseline:
        try:
            cpu = cpu_baseline(scene)
        except Exception as e:  # the baseline is reported, never the product path
            cpu = {"value": None, "unit": UNIT, "cores": 0, "kind": "port", "sample": f"failed: {e}"}
        if not args.no_numba:
            cpu["numba"] = numba_baseline()

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms / args.steps, "ms_per_step_samples": [round(v, 5) for v in samples], "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32 (cell indexing in f64; stress in f32 perturbation form, f64 fallback at large strain)",
        "data": "synthetic",
        "config": {"workload": scene.name, "particles_per_gpu": n, "particles_total": n_total,
                   "grid": f"{scene.res}^{scene.dim}", "dt": scene.dt, "p2g_mode": args.p2g_mode, "cuda_graph": bool(args.graph and world == 1 and not dam and args.steps % 10 == 0),
                   **({"presteps": args.presteps} if dam else {}),
                   "l2": ("inputs larger than L2 (no flush)" if n * (112 if scene.dim == 3 else 52) > 2 * 126e6
                          else "particle state fits in the 126 MB L2 (flagged: HBM fraction is not meaningful)"),
                   "n_oob": n_oob, "host_numa": numa_note,
                   "material_layout": (solver.local.solver if hasattr(solver, "local") else solver).material_layout,
                   "parallelism": (f"{world} slabs along x, halo sum over "
                                   f"{'NCCL p2p' if solver.driver.halo == 'p2p' else 'symmetric-memory puts over NVLink'} every substep, device-side "
                                   f"migration (pack/unpack kernels) every {solver.driver.migrate_every} substeps; "
                                   + ("dam break" if dam else coupling)) if hasattr(solver, "driver") else "single GPU",
                   **({"migration": mig_stats} if world > 1 else {}),
                   **({"slab_particles": slab_counts, "slab_cells": [list(r) for r in solver.plan.all_ranges],
                       "rebalanced": solver.driver.rebalanced} if world > 1 else {})},
        "clocks": clocks, "e2e": e2e, "gpu_launches": int(launches), "roofline": roofline, "cpu_baseline": cpu,
        "parity": parity,
    }
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
