#!/bin/bash
# One-GPU measurement pass whose outputs are copied into profiles/ (run under gpurun).
set -u
out=gpurun_out/final
mkdir -p $out
timeout 600 python -m pytest tests -m gpu -x -q > $out/pytest_gpu.txt 2>&1
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > $out/smoke.txt 2>&1
timeout 600 python bench.py > $out/bench_3d16m.json 2> $out/bench_3d16m.err
timeout 300 python bench.py --workload 2d1m --steps 400 > $out/bench_2d1m.json 2> $out/bench_2d1m.err
timeout 300 python bench.py --workload 3d16m-rest --no-cpu-baseline > $out/bench_3d16m_rest.json 2> $out/bench_3d16m_rest.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > $out/bench_reference_arm.json 2> $out/bench_reference_arm.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $out/launches.csv \
    python bench.py --steps 8 --warmup 3 --no-cpu-baseline > $out/launches_run.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on \
    -k regex:"p2g_bulk3|g2p_tiled3|grid_op3_blocks|bin_scatter|scan_downsweep" --launch-skip 25 --launch-count 5 -f -o $out/r01n \
    python bench.py --steps 8 --warmup 3 --no-cpu-baseline > $out/ncu_full.log 2>&1
ls -la $out
