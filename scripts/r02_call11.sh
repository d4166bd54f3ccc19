#!/bin/bash
# call 11 (2 GPUs): chunked phase 2 + slack-based migration cadence + GPU scene generators
set -u
out=gpurun_out/r02k
mkdir -p $out
timeout 900 python -m pytest tests -m gpu -x -q > $out/pytest_gpu_2gpus.txt 2>&1
tail -5 $out/pytest_gpu_2gpus.txt
timeout 200 python bench.py --steps 100 --warmup 5 --no-cpu-baseline --e2e-serial-only --e2e-steps 1 > $out/bench.json 2> $out/bench.err
python -c "import json;d=json.load(open('$out/bench.json'));print('headline', d['ms_per_step'], d['roofline']['phase_ms'], d['parity']['max_norm_rel_err'])"
for c in 0.02 0.1 0.3; do
  timeout 120 python bench.py --steps 60 --warmup 5 --workload 3d16m-drift:$c --no-cpu-baseline --e2e-serial-only --e2e-steps 1 > $out/bench_drift$c.json 2> $out/bench_drift$c.err
  python -c "import json;d=json.load(open('$out/bench_drift$c.json'));print('drift $c', d['ms_per_step'], d['roofline']['phase_ms'], d['parity']['within_tolerance'])"
done
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 300 $TR --master-port 29651 bench.py --gpus 2 --steps 100 --warmup 10 --no-cpu-baseline --slab-timing --e2e-serial-only --e2e-steps 1 > $out/bar_symm.json 2> $out/bar_symm.err
grep -a -o "\[rank [0-9]\] slab phase ms/substep: [a-z0-9., ]*" $out/bar_symm.err
python -c "import json;d=json.load(open('$out/bar_symm.json'));print('bar n2', d['ms_per_step'], d['value'], d['config']['migration'], d['config']['parallelism'])"
