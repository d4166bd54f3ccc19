#!/bin/bash
# Round 2, call 19: warp-window 2D kernels after the bank-conflict padding / single-store / approximate-unit changes.
set -u
out=gpurun_out/r02s
mkdir -p $out
timeout 120 python scripts/sanity_2d.py > $out/sanity_2d.txt 2>&1; tail -4 $out/sanity_2d.txt
timeout 300 python -m pytest tests -m gpu -x -q -k "2d or c2 or snow or quirk or wall" > $out/pytest_2d.txt 2>&1
tail -3 $out/pytest_2d.txt
B="python bench.py --workload 2d1m --no-cpu-baseline --e2e-steps 1 --steps 400 --warmup 10"
run() { # name, env...
  name=$1; shift
  env "$@" timeout 120 $B > $out/bench_2d_$name.json 2> $out/bench_2d_$name.err
  python - <<PY
import json
try:
    d=json.load(open('$out/bench_2d_$name.json')); print('$name', round(d['ms_per_step']*1e3,2), 'us', d['roofline'].get('phase_ms'), d.get('parity',{}).get('max_norm_rel_err'))
except Exception as e: print('$name', 'failed', e)
PY
}
run default FFMPM_W2_FAST=1
run nofast FFMPM_W2_FAST=0
M=gpu__time_duration.sum,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum,l1tex__throughput.avg.pct_of_peak_sustained_elapsed,sm__warps_active.avg.pct_of_peak_sustained_active,launch__registers_per_thread
timeout 200 ncu --metrics $M --clock-control none -k regex:'window2|grid_op2' -s 9 -c 6 --csv --log-file $out/ncu_2d_kernels.csv \
    $B --steps 6 --warmup 3 --no-parity > $out/ncu_2d.log 2>&1
python - <<PY
import csv
rows=[r for r in csv.reader(open('$out/ncu_2d_kernels.csv')) if len(r)>10]
h=rows[0]; ik=h.index('Kernel Name'); im=h.index('Metric Name'); iv=h.index('Metric Value'); iid=h.index('ID')
for r in rows[1:]:
    print(r[iid], r[ik][:28], r[im][:60], r[iv])
PY
