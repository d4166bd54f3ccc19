#!/bin/bash
# Round 2, call 26: full GPU suite with the 3D snow G2P (csrc/mpm_svd3.cuh) and the 2D window kernels as defaults, smoke,
# the driver's bench line, and a launch-configuration sweep of the two dominant kernels (env switches only).
set -u
out=gpurun_out/r02t
mkdir -p $out
timeout 600 python -m pytest tests -m gpu -x -q --durations=5 > $out/pytest_gpu.txt 2>&1
tail -4 $out/pytest_gpu.txt
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > $out/smoke.txt 2>&1; tail -1 $out/smoke.txt
timeout 400 python bench.py --gpus 1 --steps 20 --warmup 5 --no-numba > $out/bench_driver_line.json 2> $out/bench_driver_line.err
python -c "import json;d=json.load(open('$out/bench_driver_line.json'));print('driver line', d['ms_per_step'], d['value'], d['roofline']['phase_ms'], d['roofline']['frac'], d['e2e']['ms_per_step'], d['cpu_baseline']['value'], d['parity']['within_tolerance'], d['clocks'])"
B="python bench.py --no-cpu-baseline --no-parity --e2e-steps 1 --e2e-serial-only --steps 100 --warmup 5"
run() { # name, env...
  name=$1; shift
  env "$@" timeout 100 $B > $out/sweep_$name.json 2> $out/sweep_$name.err
  python - <<PY
import json
try:
    d=json.load(open('$out/sweep_$name.json')); print('$name', round(d['ms_per_step'],4), d['roofline'].get('phase_ms'))
except Exception as e: print('$name', 'failed', e)
PY
}
run base FFMPM_NOP=1
run wpw8 FFMPM_P2G_WPW=8
run wpw32 FFMPM_P2G_WPW=32
run g2p_bps6 FFMPM_G2P_BPS=6
run g2p_bps12 FFMPM_G2P_BPS=12
run g2p_bps16 FFMPM_G2P_BPS=16
