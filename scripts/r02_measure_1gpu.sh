#!/bin/bash
# One-GPU measurement recipe of round 2 (what profiles/r02d_* were produced with):
#   gpurun --timeout 2400 -- 'bash scripts/r02_measure_1gpu.sh'
set -u
out=gpurun_out/r02_1gpu
mkdir -p $out
timeout 900 python -m pytest tests -m gpu -x -q --durations=5 > $out/pytest_gpu.txt 2>&1
tail -5 $out/pytest_gpu.txt
timeout 400 python bench.py --steps 100 --warmup 5 > $out/bench.json 2> $out/bench.err
for c in 0.02 0.1 0.3; do
  timeout 120 python bench.py --steps 60 --warmup 5 --workload 3d16m-drift:$c --no-cpu-baseline --e2e-serial-only --e2e-steps 1 > $out/bench_drift$c.json 2> $out/bench_drift$c.err
done
timeout 200 python bench.py --steps 400 --warmup 10 --workload 2d1m --no-cpu-baseline --e2e-steps 1 > $out/bench_2d.json 2> $out/bench_2d.err
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed,lts__t_sector_hit_rate.pct,sm__warps_active.avg.pct_of_peak_sustained_active,launch__registers_per_thread,smsp__inst_executed.sum,smsp__thread_inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active,sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active,lts__t_sectors_op_red.sum,lts__t_sectors_op_atom.sum,smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio,smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio,smsp__average_warps_issue_stalled_wait_per_issue_active.ratio,smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio
timeout 300 ncu --metrics $M --clock-control none -k regex:'p2g_bulk3|g2p_tiled3' -s 12 -c 4 --csv --log-file $out/ncu_kernels.csv \
    python bench.py --steps 6 --warmup 3 --no-cpu-baseline --no-parity --e2e-serial-only --e2e-steps 1 > $out/ncu_run.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:'p2g_bulk3|g2p_tiled3' -s 12 -c 2 -o $out/prof_p2g_g2p \
    python bench.py --steps 6 --warmup 3 --no-cpu-baseline --no-parity --e2e-serial-only --e2e-steps 1 > $out/ncu_full.log 2>&1
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -s 200 -c 120 --csv --log-file $out/launches.csv \
    python bench.py --steps 8 --warmup 3 --no-cpu-baseline --no-parity --e2e-serial-only --e2e-steps 1 > $out/launches_run.log 2>&1
timeout 300 compute-sanitizer --tool memcheck --error-exitcode 3 python -c "import __graft_entry__ as g; g.smoke()" > $out/sanitizer_memcheck.txt 2>&1
timeout 400 compute-sanitizer --tool racecheck --error-exitcode 3 python -c "import __graft_entry__ as g; g.smoke()" > $out/sanitizer_racecheck.txt 2>&1
# here, afterwards:  ncu -i $out/prof_p2g_g2p.ncu-rep --page source --csv --print-source sass > src.csv
#                    python scripts/ncu_hot_regions.py src.csv p2g_bulk3 --chunk 100
