#!/bin/bash
# call 8 (2 GPUs): fused 2D with coherent L1 loads; dam-break long-run diagnosis (N = 1 physics vs N = 2 slabs)
set -u
out=gpurun_out/r02h
mkdir -p $out
timeout 300 python -m pytest tests -m gpu -x -q -k "2d or c2_1m" > $out/pytest_2d.txt 2>&1
tail -3 $out/pytest_2d.txt
timeout 200 python bench.py --steps 400 --warmup 10 --workload 2d1m --no-cpu-baseline --e2e-steps 1 > $out/bench_2d_fused.json 2> $out/bench_2d_fused.err
python -c "
import json
d=json.load(open('$out/bench_2d_fused.json')); print('2d fused', d['ms_per_step'], d['gpu_launches'], d['parity']['max_norm_rel_err'])"
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
for ps in 3000 12000; do
timeout 300 $TR --nproc-per-node 1 --master-port 29621 bench.py --gpus 1 --no-cpu-baseline --steps 20 --warmup 3 --workload dam:2000000 --presteps $ps > $out/dam_n1_$ps.json 2> $out/dam_n1_$ps.err
python -c "
import json
try:
    d=json.load(open('$out/dam_n1_$ps.json')); print('dam n1 presteps $ps ok', d['ms_per_step'], d['config']['n_oob'])
except Exception as e:
    print('dam n1 $ps FAILED'); print(open('$out/dam_n1_$ps.err', errors='replace').read()[-600:])"
done
for ps in 3000 12000; do
timeout 300 $TR --nproc-per-node 2 --master-port 29622 bench.py --gpus 2 --no-cpu-baseline --steps 20 --warmup 3 --workload dam:2000000 --presteps $ps > $out/dam_n2_$ps.json 2> $out/dam_n2_$ps.err
python -c "
import json
try:
    d=json.load(open('$out/dam_n2_$ps.json')); print('dam n2 presteps $ps ok', d['ms_per_step'], d['config']['slab_particles'], d['config']['migration'])
except Exception as e:
    print('dam n2 $ps FAILED'); print(open('$out/dam_n2_$ps.err', errors='replace').read()[-900:])"
done
