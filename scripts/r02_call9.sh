#!/bin/bash
set -u
out=gpurun_out/r02i
mkdir -p $out
timeout 900 python -m pytest tests -m gpu -x -q > $out/pytest_gpu.txt 2>&1
tail -4 $out/pytest_gpu.txt
timeout 200 python bench.py --steps 100 --warmup 5 --no-cpu-baseline --e2e-serial-only --e2e-steps 1 > $out/bench.json 2> $out/bench.err
python -c "import json;d=json.load(open('$out/bench.json'));print('headline', d['ms_per_step'], d['roofline']['phase_ms'])"
timeout 200 python bench.py --steps 400 --warmup 10 --workload 2d1m --no-cpu-baseline --e2e-steps 1 > $out/bench_2d.json 2> $out/bench_2d.err
python -c "import json;d=json.load(open('$out/bench_2d.json'));print('2d', d['ms_per_step'], d['config']['cuda_graph'], d['roofline']['phase_ms'])"
for wl in dam:4000000 dam:4000000-late; do
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 1 --master-addr 127.0.0.1 --master-port 29631 bench.py --gpus 1 --no-cpu-baseline --steps 50 --warmup 5 --workload $wl > $out/dam1.json 2> $out/dam1.err
python -c "import json;d=json.load(open('$out/dam1.json'));print('$wl n1', d['ms_per_step'], d['config']['workload'], d['config']['particles_total'])" || tail -c 500 $out/dam1.err
done
