#!/bin/bash
# Round 2, call 18: A/B of the warp-window 2D kernels -- distinct-cells fast path, register caps.
set -u
out=gpurun_out/r02r
mkdir -p $out
timeout 120 python scripts/sanity_2d.py > $out/sanity_2d.txt 2>&1; tail -4 $out/sanity_2d.txt
timeout 300 python -m pytest tests -m gpu -x -q -k "2d or c2 or snow or quirk or wall" > $out/pytest_2d.txt 2>&1
tail -3 $out/pytest_2d.txt
B="python bench.py --workload 2d1m --no-cpu-baseline --e2e-steps 1 --steps 400 --warmup 10"
run() { # name, env...
  name=$1; shift
  env "$@" timeout 120 $B --no-parity > $out/bench_2d_$name.json 2> $out/bench_2d_$name.err
  python - <<PY
import json
try:
    d=json.load(open('$out/bench_2d_$name.json')); print('$name', round(d['ms_per_step']*1e3,2), 'us', d['roofline'].get('phase_ms'))
except Exception as e: print('$name', 'failed', e)
PY
}
run default FFMPM_W2_FAST=1
run nofast FFMPM_W2_FAST=0
run p2g6 FFMPM_W2_P2G_MINB=6
run p2g8 FFMPM_W2_P2G_MINB=8
run g2p6 FFMPM_W2_G2P_MINB=6
run g2p8 FFMPM_W2_G2P_MINB=8
run both8 FFMPM_W2_P2G_MINB=8 FFMPM_W2_G2P_MINB=8
timeout 300 compute-sanitizer --tool racecheck --error-exitcode 3 python scripts/sanity_2d.py > $out/sanitizer_racecheck_2d.txt 2>&1; echo "racecheck rc $?"; tail -2 $out/sanitizer_racecheck_2d.txt
