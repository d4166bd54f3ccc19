#!/bin/bash
# Round-2 GPU call 7 (two B200s): fused persistent 2D substep kernel (tests + bench on GPU 0), dam-break late window rehearsal.
set -u
out=gpurun_out/r02g
mkdir -p $out
timeout 600 python -m pytest tests -m gpu -x -q -k "2d or c2_1m" > $out/pytest_2d.txt 2>&1
tail -4 $out/pytest_2d.txt
timeout 200 python bench.py --steps 400 --warmup 10 --workload 2d1m --no-cpu-baseline --e2e-steps 1 > $out/bench_2d_fused.json 2> $out/bench_2d_fused.err
python -c "
import json
d=json.load(open('$out/bench_2d_fused.json')); print('2d fused', d['ms_per_step'], d['gpu_launches'], d['config'].get('substeps_per_launch'), d['parity']['max_norm_rel_err'])"
for bps in 1 2 8; do
FFMPM_FUSE2D_BPS=$bps timeout 200 python bench.py --steps 400 --warmup 10 --workload 2d1m --no-cpu-baseline --no-parity --e2e-steps 1 > $out/bench_2d_fused_bps$bps.json 2> $out/bench_2d_fused_bps$bps.err
python -c "
import json
d=json.load(open('$out/bench_2d_fused_bps$bps.json')); print('2d fused bps $bps', d['ms_per_step'])"
done
FFMPM_FUSE2D=0 timeout 200 python bench.py --steps 400 --warmup 10 --workload 2d1m --no-cpu-baseline --no-parity --e2e-steps 1 --graph > $out/bench_2d_graph.json 2> $out/bench_2d_graph.err
python -c "
import json
d=json.load(open('$out/bench_2d_graph.json')); print('2d separate kernels, graph', d['ms_per_step'])"
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 400 $TR --master-port 29611 bench.py --gpus 2 --no-cpu-baseline --steps 100 --warmup 10 --workload dam:8388608 --presteps 12000 > $out/dam_static_late.json 2> $out/dam_static_late.err
timeout 400 $TR --master-port 29612 bench.py --gpus 2 --no-cpu-baseline --steps 100 --warmup 10 --workload dam:8388608 --presteps 12000 --rebalance --rebalance-every 500 > $out/dam_recut_late.json 2> $out/dam_recut_late.err
python - <<PY
import json
for t in ("dam_static_late", "dam_recut_late"):
    try:
        d = json.load(open("$out/%s.json" % t)); print(t, d["ms_per_step"], d["config"].get("slab_particles"), d["config"].get("rebalanced"), d["config"].get("slab_cells"))
    except Exception as e:
        print(t, "FAILED", e); print(open("$out/%s.err" % t, errors="replace").read()[-800:])
PY
