#!/bin/bash
# call 12 (2 GPUs): P2G back on the item mapping + window sort; slack-based migration cadence with a warmed first round
set -u
out=gpurun_out/r02l
mkdir -p $out
timeout 600 python -m pytest tests -m gpu -x -q -k "p2g or fast_moving or dense or large_block or multi_substep or c3_16m or two_gpu" > $out/pytest_subset.txt 2>&1
tail -3 $out/pytest_subset.txt
timeout 200 python bench.py --steps 100 --warmup 5 --no-cpu-baseline --e2e-serial-only --e2e-steps 1 > $out/bench.json 2> $out/bench.err
python -c "import json;d=json.load(open('$out/bench.json'));print('headline', d['ms_per_step'], d['roofline']['phase_ms'])"
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 300 $TR --master-port 29661 bench.py --gpus 2 --steps 100 --warmup 10 --no-cpu-baseline --slab-timing --e2e-steps 2 > $out/bar_symm.json 2> $out/bar_symm.err
grep -a -o "\[rank [0-9]\] slab phase ms/substep: [a-z0-9., ]*" $out/bar_symm.err
python -c "import json;d=json.load(open('$out/bar_symm.json'));print('bar n2', d['ms_per_step'], d['value'], d['config']['migration'], d['e2e']['ms_per_step'])"
timeout 300 $TR --master-port 29662 bench.py --gpus 2 --steps 100 --warmup 10 --no-cpu-baseline --drift 0.1 --e2e-serial-only --e2e-steps 1 > $out/bar_symm_drift0.1.json 2> $out/bar_symm_drift0.1.err
python -c "import json;d=json.load(open('$out/bar_symm_drift0.1.json'));print('bar n2 drift 0.1', d['ms_per_step'], d['value'], d['config']['migration'])"
