#!/bin/bash
# Round-2 GPU call 4 (one B200): GPU suite, headline bench, P2G under particle motion, 2D, ncu counters + full capture.
set -u
out=gpurun_out/r02d
mkdir -p $out
timeout 900 python -m pytest tests -m gpu -x -q --durations=5 > $out/pytest_gpu.txt 2>&1
tail -8 $out/pytest_gpu.txt
timeout 400 python bench.py --steps 100 --warmup 5 > $out/bench.json 2> $out/bench.err
python -c "import json;d=json.load(open('$out/bench.json'));print('headline', d['ms_per_step'], d['roofline']['phase_ms'], d['e2e']['ms_per_step'], d['e2e'].get('serial',{}).get('ms_per_step'), d['config']['host_numa'])"
for c in 0.02 0.1 0.3; do
  timeout 120 python bench.py --steps 60 --warmup 5 --workload 3d16m-drift:$c --no-cpu-baseline --no-parity --e2e-serial-only --e2e-steps 1 > $out/bench_drift$c.json 2> $out/bench_drift$c.err
  python -c "import json;d=json.load(open('$out/bench_drift$c.json'));print('drift $c', d['ms_per_step'], d['roofline']['phase_ms'])"
done
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed,lts__t_sector_hit_rate.pct,sm__warps_active.avg.pct_of_peak_sustained_active,launch__registers_per_thread,smsp__inst_executed.sum,smsp__thread_inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active,sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_lsu.sum,l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum,lts__t_sectors_op_red.sum,lts__t_sectors_op_atom.sum,lts__t_requests_srcunit_tex_op_red.sum,lts__throughput.avg.pct_of_peak_sustained_elapsed,l1tex__throughput.avg.pct_of_peak_sustained_elapsed,smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio,smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio,smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio,smsp__average_warps_issue_stalled_wait_per_issue_active.ratio,smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio,smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio,smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio
timeout 300 ncu --metrics $M --clock-control none -k regex:'p2g_bulk3|g2p_tiled3' -s 12 -c 4 --csv --log-file $out/ncu_kernels.csv \
    python bench.py --steps 6 --warmup 3 --no-cpu-baseline --no-parity --e2e-serial-only --e2e-steps 1 > $out/ncu_run.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:'p2g_bulk3|g2p_tiled3' -s 12 -c 2 -o $out/prof_p2g_g2p \
    python bench.py --steps 6 --warmup 3 --no-cpu-baseline --no-parity --e2e-serial-only --e2e-steps 1 > $out/ncu_full.log 2>&1
ls -la $out/*.ncu-rep
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -s 200 -c 120 --csv --log-file $out/launches.csv \
    python bench.py --steps 8 --warmup 3 --no-cpu-baseline --no-parity --e2e-serial-only --e2e-steps 1 > $out/launches_run.log 2>&1
timeout 200 python bench.py --steps 400 --warmup 10 --workload 2d1m --no-cpu-baseline --e2e-steps 1 > $out/bench_2d.json 2> $out/bench_2d.err
timeout 200 python bench.py --steps 400 --warmup 10 --workload 2d1m --no-cpu-baseline --no-parity --e2e-steps 1 --graph > $out/bench_2d_graph.json 2> $out/bench_2d_graph.err
python -c "
import json
for f in ('bench_2d','bench_2d_graph'):
    d=json.load(open('$out/%s.json'%f)); print(f, d['ms_per_step'], d['roofline']['phase_ms'])"
