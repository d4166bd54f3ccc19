#!/bin/bash
# 2D probe (round 2, call 16): parity of the fp32 closed-form 2D stress / SVD-free G2P on hardware, their effect on the
# 1 M / 1024^2 substep, what bounds the 2D scatter (reductions vs arithmetic), ncu --set full of every 2D kernel.
set -u
out=gpurun_out/r02p
mkdir -p $out
timeout 300 python -m pytest tests -m gpu -x -q -k "2d or c2 or snow or quirk" > $out/pytest_2d.txt 2>&1
tail -3 $out/pytest_2d.txt
B="python bench.py --workload 2d1m --no-cpu-baseline --e2e-steps 1"
timeout 120 $B --steps 400 --warmup 10 > $out/bench_2d.json 2> $out/bench_2d.err
FFMPM_DEBUG_NORED=1 timeout 120 $B --steps 400 --warmup 10 --no-parity > $out/bench_2d_nored.json 2> $out/bench_2d_nored.err
FFMPM_2D_BINNED=1 timeout 120 $B --steps 400 --warmup 10 > $out/bench_2d_binned.json 2> $out/bench_2d_binned.err
for f in bench_2d bench_2d_nored bench_2d_binned; do
python - <<PY
import json
try:
    d=json.load(open('$out/$f.json')); print('$f', round(d['ms_per_step']*1e3,2), 'us', d['roofline'].get('phase_ms'), d.get('parity',{}).get('within_tolerance'))
except Exception as e: print('$f', 'failed', e)
PY
done
timeout 300 ncu --set full --clock-control none --import-source on -k regex:'scatter2|gather2|grid_op2|memset|clear' -s 8 -c 8 -o $out/prof_2d \
    $B --steps 6 --warmup 3 --no-parity > $out/ncu_2d.log 2>&1
FFMPM_2D_BINNED=1 timeout 300 ncu --set full --clock-control none --import-source on -k regex:'runs2|reorder2|grid_op2|bin_|scan_|tiles' -s 22 -c 14 -o $out/prof_2d_binned \
    $B --steps 6 --warmup 3 --no-parity > $out/ncu_2d_binned.log 2>&1
timeout 120 ncu --metrics gpu__time_duration.sum --clock-control none -s 8 -c 60 --csv --log-file $out/launches_2d.csv \
    $B --steps 8 --warmup 3 --no-parity > $out/launches_2d.log 2>&1
ls -la $out
