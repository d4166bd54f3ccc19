#!/bin/bash
# Round 2, call 29: ncu launch list and kernel metrics of the final defaults (P2G work items of 8 windows per warp), and the
# drifting block (ragged windows) with work items of 8 against 16.
set -u
out=gpurun_out/r02w
mkdir -p $out
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed,lts__t_sector_hit_rate.pct,sm__warps_active.avg.pct_of_peak_sustained_active,launch__registers_per_thread,smsp__inst_executed.sum,smsp__thread_inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active,lts__t_sectors_op_red.sum,lts__t_sectors_op_atom.sum,smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio,smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio
timeout 120 ncu --metrics gpu__time_duration.sum --clock-control none -s 200 -c 120 --csv --log-file $out/launches.csv \
    python bench.py --steps 8 --warmup 3 --no-cpu-baseline --no-parity --e2e-serial-only --e2e-steps 1 > $out/launches_run.log 2>&1
python scripts/launch_summary.py $out/launches.csv > $out/launches_summary.txt 2>&1; cat $out/launches_summary.txt
timeout 150 ncu --metrics $M --clock-control none -k regex:'p2g_bulk3|g2p_tiled3' -s 12 -c 4 --csv --log-file $out/ncu_kernels.csv \
    python bench.py --steps 6 --warmup 3 --no-cpu-baseline --no-parity --e2e-serial-only --e2e-steps 1 > $out/ncu_run.log 2>&1
for w in 8 16; do
  FFMPM_P2G_WPW=$w timeout 100 python bench.py --steps 60 --warmup 5 --workload 3d16m-drift:0.1 --no-cpu-baseline --e2e-serial-only --e2e-steps 1 > $out/bench_drift0.1_wpw$w.json 2> $out/bench_drift0.1_wpw$w.err
  python -c "import json;d=json.load(open('$out/bench_drift0.1_wpw$w.json'));print('drift 0.1 wpw$w', d['ms_per_step'], d['roofline']['phase_ms'], d['parity']['within_tolerance'])"
done
