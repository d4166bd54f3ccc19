"""Can P2G scatter its halo planes straight into the neighbour GPU's memory?  Two ranks, one symmetric-memory
buffer each; rank 0 issues the 16-byte vector reductions P2G uses (red.global.add.v4.f32, through
ffmpm_debug_red_add4) at rank 1's buffer over NVLink while rank 1 does the same locally, then rank 1 checks
the sums and both report the time per reduction against local memory.

    gpurun --gpus 2 --timeout 300 -- 'python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 \
        --master-addr 127.0.0.1 --master-port 29541 scripts/peer_red_probe.py'
"""
import ctypes as C
import json
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from femflow_b200 import _native as N  # noqa: E402


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", device_id=dev)
    import torch.distributed._symmetric_memory as symm_mem
    lib = N.lib()
    count = 4 * 1024 * 1024                       # nodes: 64 MB of float4
    buf = symm_mem.empty(count * 4, dtype=torch.float32, device=dev)
    buf.zero_()
    hdl = symm_mem.rendezvous(buf, dist.group.WORLD)
    peer = hdl.get_buffer((rank + 1) % world, (count * 4,), torch.float32, 0)
    hdl.barrier(0, 20000)
    val = (C.c_float * 4)(1.0, 2.0, 3.0, 0.5)
    stream = C.c_void_p(torch.cuda.current_stream().cuda_stream)

    def timed(ptr, reps=5):
        N.check(lib.ffmpm_debug_red_add4(C.c_void_p(ptr), val, count, stream))
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            N.check(lib.ffmpm_debug_red_add4(C.c_void_p(ptr), val, count, stream))
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / reps, reps + 1

    out = {"rank": rank}
    t_local, n_local = timed(buf.data_ptr())
    dist.barrier()
    n_remote = 0
    if rank == 0:
        t_peer, n_remote = timed(peer.data_ptr())
        out["peer_ms"] = t_peer
        out["peer_GBps"] = count * 16 / t_peer / 1e6
    out["local_ms"] = t_local
    out["local_GBps"] = count * 16 / t_local / 1e6
    hdl.barrier(0, 20000)
    torch.cuda.synchronize()
    dist.barrier()
    if rank == 1:
        adds = n_local + 6                        # own passes + rank 0's six passes over NVLink
        want = torch.tensor([1.0, 2.0, 3.0, 0.5], device=dev) * adds
        got = buf.view(count, 4)
        out["peer_sums_exact"] = bool((got == want).all())
        out["first_node"] = got[0].tolist()
    gathered = [None] * world
    dist.all_gather_object(gathered, out)
    if rank == 0:
        os.makedirs("gpurun_out", exist_ok=True)
        with open("gpurun_out/peer_red_probe.json", "w") as f:
            json.dump(gathered, f, indent=1)
        print(json.dumps(gathered))
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
