#!/bin/bash
# Round 2, last call: the driver's own two bench lines at HEAD (after the traffic.json / numba-script edits).
set -u
out=gpurun_out/r02z
mkdir -p $out
timeout 200 python bench.py --gpus 1 --steps 20 --warmup 5 > $out/bench_driver_line.json 2> $out/bench_driver_line.err
python -c "import json;d=json.load(open('$out/bench_driver_line.json'));print('driver line', d['ms_per_step'], d['value'], d['e2e']['value'], d['roofline']['frac'], d['roofline']['traffic'], d['cpu_baseline']['value'], d['cpu_baseline']['numba'].get('value'), d['parity']['within_tolerance'], d['clocks']['samples'], d['gpu_launches'])"
timeout 100 python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > $out/bench_reference_line.json 2> $out/bench_reference_line.err
python -c "import json;d=json.load(open('$out/bench_reference_line.json'));print('reference line', d['value'], d['steps'], d['cpu_baseline']['cores'])"
