"""Bind the calling process to the CPUs next to its GPU before it allocates pinned host memory.

``cudaHostAlloc`` places pages on the NUMA node of the allocating thread; with eight ranks streaming 1.6 GB
each way per step, buffers on the wrong socket cross the inter-socket link and share one memory controller
(round 1: 8 GPUs moved 8 GB/s per GPU instead of 55).  Linux only; silently does nothing when sysfs does not
say where the GPU sits."""
from __future__ import annotations

import os
import subprocess


def _parse_cpulist(text: str):
    cpus = set()
    for part in text.strip().split(","):
        if not part:
            continue
        a, _, b = part.partition("-")
        cpus.update(range(int(a), int(b or a) + 1))
    return cpus


def gpu_local_cpus(index: int):
    """CPUs of the NUMA node the GPU hangs off (sysfs ``local_cpulist`` of its PCI function), or None."""
    try:
        out = subprocess.run(["nvidia-smi", "--query-gpu=pci.bus_id", "--format=csv,noheader", "-i", str(index)],
                             capture_output=True, text=True, timeout=10).stdout.strip().splitlines()[0].strip().lower()
        bdf = out[-12:] if len(out) > 12 else out          # nvidia-smi prints an 8-digit domain, sysfs uses 4
        path = f"/sys/bus/pci/devices/{bdf}/local_cpulist"
        cpus = _parse_cpulist(open(path).read())
        return cpus or None
    except Exception:
        return None


def bind_to_gpu(index: int) -> str:
    """Restrict this process to the GPU-local CPUs (intersected with its current affinity).  Returns a one-line
    description of what happened, for the bench record."""
    try:
        local = gpu_local_cpus(index)
        if not local:
            return "no NUMA information for the GPU: affinity unchanged"
        allowed = os.sched_getaffinity(0)
        use = (allowed & local) or allowed
        os.sched_setaffinity(0, use)
        return f"pinned allocations from {len(use)} GPU-local CPUs ({min(use)}-{max(use)})"
    except Exception as e:  # never fatal
        return f"affinity unchanged ({type(e).__name__})"
