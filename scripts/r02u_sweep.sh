#!/bin/bash
# Round 2, call 27: P2G work-item size sweep (FFMPM_P2G_WPW, windows per warp per work item), the driver's bench line with
# the in-process NVML clock sampler, compute-sanitizer over the 3D snow tests, the 2D line.
set -u
out=gpurun_out/r02u
mkdir -p $out
B="python bench.py --no-cpu-baseline --no-parity --e2e-steps 1 --e2e-serial-only --steps 100 --warmup 5"
run() { # name, env...
  name=$1; shift
  env "$@" timeout 100 $B > $out/sweep_$name.json 2> $out/sweep_$name.err
  python - <<PY
import json
try:
    d=json.load(open('$out/sweep_$name.json')); print('$name', round(d['ms_per_step'],4), d['ms_per_step_samples'], d['roofline'].get('phase_ms'), d['clocks'])
except Exception as e: print('$name', 'failed', e)
PY
}
run wpw16 FFMPM_P2G_WPW=16
run wpw4 FFMPM_P2G_WPW=4
run wpw6 FFMPM_P2G_WPW=6
run wpw8 FFMPM_P2G_WPW=8
run wpw10 FFMPM_P2G_WPW=10
run wpw16b FFMPM_P2G_WPW=16
timeout 200 python bench.py --gpus 1 --steps 20 --warmup 5 --no-numba --no-cpu-baseline > $out/bench_driver_line.json 2> $out/bench_driver_line.err
python -c "import json;d=json.load(open('$out/bench_driver_line.json'));print('driver line', d['ms_per_step'], d['value'], d['e2e']['ms_per_step'], d['parity']['within_tolerance'], d['clocks'])"
timeout 100 python bench.py --workload 2d1m --no-cpu-baseline --e2e-steps 1 --steps 400 --warmup 10 > $out/bench_2d.json 2> $out/bench_2d.err
python -c "import json;d=json.load(open('$out/bench_2d.json'));print('2d', d['ms_per_step'], d['value'], d['parity']['within_tolerance'], d['clocks'])"
timeout 240 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -x -q -k "snow_g2p_3d or 3d_snow_substeps" > $out/sanitizer_memcheck_snow3d.txt 2>&1
echo "sanitizer rc=$?"; tail -4 $out/sanitizer_memcheck_snow3d.txt
