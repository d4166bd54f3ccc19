#!/bin/bash
# Round-2 GPU call 2 (one B200): the whole GPU suite incl. parity at the headline sizes, the bench line, and
# one ncu pass over the new P2G kernel (counters the roofline discussion needs).
set -u
out=gpurun_out/r02b
mkdir -p $out
timeout 900 python -m pytest tests -m gpu -x -q --durations=8 > $out/pytest_gpu.txt 2>&1
tail -15 $out/pytest_gpu.txt
timeout 400 python bench.py --steps 100 --warmup 5 > $out/bench.json 2> $out/bench.err
tail -c 1500 $out/bench.json
timeout 120 python bench.py --steps 100 --warmup 5 --workload 3d16m-rest --no-cpu-baseline --no-parity --e2e-steps 1 > $out/bench_rest.json 2> $out/bench_rest.err
timeout 120 python bench.py --steps 200 --warmup 10 --workload 2d1m --no-cpu-baseline --e2e-steps 1 > $out/bench_2d.json 2> $out/bench_2d.err
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed,lts__t_sector_hit_rate.pct,sm__warps_active.avg.pct_of_peak_sustained_active,launch__registers_per_thread,smsp__inst_executed.sum,smsp__thread_inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active,sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_lsu.sum,sm__inst_executed_pipe_fp64.sum,l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum,lts__t_sectors_op_red.sum,lts__t_sectors_op_atom.sum,lts__t_requests_srcunit_tex_op_red.sum,l1tex__t_set_accesses_pipe_lsu_mem_global_op_red.sum,lts__throughput.avg.pct_of_peak_sustained_elapsed,l1tex__throughput.avg.pct_of_peak_sustained_elapsed,smsp__average_warp_latency_issue_stalled_long_scoreboard.ratio,smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio,smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio,smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio,smsp__average_warps_issue_stalled_wait_per_issue_active.ratio,smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio,smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio,smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio,smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio
timeout 300 ncu --metrics $M --clock-control none -k regex:'p2g_bulk3|g2p_tiled3|bin_scatter|scan_' -s 40 -c 12 --csv --log-file $out/ncu_kernels.csv \
    python bench.py --steps 6 --warmup 3 --no-cpu-baseline --no-parity --e2e-steps 1 > $out/ncu_run.log 2>&1
tail -3 $out/ncu_run.log
