#!/bin/bash
# Two-GPU check of the slab rebalancing (run under `gpurun --gpus 2`): the NCCL parity test, the default
# weak-scaling bench line on a small block, and the dam break with and without a re-cut.
set -u
out=gpurun_out/rebalance
mkdir -p $out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 120 python -m pytest tests/test_gpu_distributed.py -x -q -k "rebalance" > $out/pytest.txt 2>&1
timeout 100 $TR --master-port 29511 bench.py --gpus 2 --steps 50 --warmup 5 --workload 3d:128:32 --no-cpu-baseline \
    > $out/bench_2gpu_small.json 2> $out/bench_2gpu_small.err
timeout 100 $TR --master-port 29512 bench.py --gpus 2 --steps 100 --warmup 5 --workload dam:8388608 --no-cpu-baseline \
    > $out/dam8m_2gpu_static.json 2> $out/dam8m_2gpu_static.err
timeout 100 $TR --master-port 29513 bench.py --gpus 2 --steps 100 --warmup 5 --workload dam:8388608 --no-cpu-baseline --rebalance \
    > $out/dam8m_2gpu_rebalanced.json 2> $out/dam8m_2gpu_rebalanced.err
tail -3 $out/pytest.txt; tail -c 600 $out/*.json; tail -5 $out/*.err
