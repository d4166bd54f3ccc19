#!/bin/bash
# Round 2, final validation (call 28): P2G work items of 8 windows per warp as the default, NVML clock sampler, 3D snow G2P.
set -u
out=gpurun_out/r02v
mkdir -p $out
timeout 600 python -m pytest tests -m gpu -x -q --durations=5 > $out/pytest_gpu.txt 2>&1
tail -3 $out/pytest_gpu.txt
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > $out/smoke.txt 2>&1; tail -1 $out/smoke.txt
timeout 400 python bench.py > $out/bench_default.json 2> $out/bench_default.err
python -c "import json;d=json.load(open('$out/bench_default.json'));print('default', d['ms_per_step'], d['value'], d['roofline']['phase_ms'], d['roofline']['frac'], d['e2e']['ms_per_step'], d['e2e']['serial']['ms_per_step'], d['cpu_baseline']['value'], d['cpu_baseline']['numba'].get('value'), d['parity']['within_tolerance'], d['clocks'])"
timeout 200 python bench.py --impl reference --steps 5 --warmup 1 > $out/bench_reference.json 2> $out/bench_reference.err
python -c "import json;d=json.load(open('$out/bench_reference.json'));print('reference', d['value'], d['cpu_baseline']['cores'], d['config']['sample'][:80])"
