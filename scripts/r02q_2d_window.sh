#!/bin/bash
# Round 2, call 17: the warp-window 2D kernels on hardware -- parity, sanitizer, the 1 M / 1024^2 substep, ncu.
set -u
out=gpurun_out/r02q
mkdir -p $out
timeout 120 python scripts/sanity_2d.py > $out/sanity_2d.txt 2>&1; tail -4 $out/sanity_2d.txt
timeout 300 python -m pytest tests -m gpu -x -q -k "2d or c2 or snow or quirk or wall" > $out/pytest_2d.txt 2>&1
tail -5 $out/pytest_2d.txt
B="python bench.py --workload 2d1m --no-cpu-baseline --e2e-steps 1"
timeout 120 $B --steps 400 --warmup 10 > $out/bench_2d.json 2> $out/bench_2d.err
for f in bench_2d; do
python - <<PY
import json
try:
    d=json.load(open('$out/$f.json')); print('$f', round(d['ms_per_step']*1e3,2), 'us', d['roofline'].get('phase_ms'), d.get('parity',{}).get('within_tolerance'), d.get('parity',{}).get('max_norm_rel_err'), d['gpu_launches'])
except Exception as e: print('$f', 'failed', e)
PY
done
timeout 200 compute-sanitizer --tool memcheck --error-exitcode 3 python scripts/sanity_2d.py > $out/sanitizer_memcheck_2d.txt 2>&1; echo "memcheck rc $?"; tail -3 $out/sanitizer_memcheck_2d.txt
timeout 300 compute-sanitizer --tool racecheck --error-exitcode 3 python scripts/sanity_2d.py > $out/sanitizer_racecheck_2d.txt 2>&1; echo "racecheck rc $?"; tail -3 $out/sanitizer_racecheck_2d.txt
timeout 200 ncu --set full --clock-control none --import-source on -k regex:'window2|grid_op2' -s 9 -c 3 -o $out/prof_2d_window \
    $B --steps 6 --warmup 3 --no-parity > $out/ncu_2d.log 2>&1
timeout 120 ncu --metrics gpu__time_duration.sum --clock-control none -s 9 -c 45 --csv --log-file $out/launches_2d.csv \
    $B --steps 8 --warmup 3 --no-parity > $out/launches_2d.log 2>&1
ls $out
