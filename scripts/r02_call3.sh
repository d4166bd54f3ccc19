#!/bin/bash
# Round-2 GPU call 3 (two B200s): the multi-GPU tests on hardware (NCCL halo + device migration, rebalance, symmetric-memory
# halo), the re-measured single-GPU P2G, and the coupled-bar weak scaling at N = 2 (p2p vs symm halo, phase tables).
set -u
out=gpurun_out/r02c
mkdir -p $out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
FFMPM_TEST_SYMM=1 timeout 600 python -m pytest tests -m gpu -q --durations=5 > $out/pytest_gpu_2gpus.txt 2>&1
tail -12 $out/pytest_gpu_2gpus.txt
timeout 200 python bench.py --steps 100 --warmup 5 --no-cpu-baseline --e2e-steps 1 > $out/bench_1gpu.json 2> $out/bench_1gpu.err
python -c "import json;d=json.load(open('$out/bench_1gpu.json'));print('1gpu', d['ms_per_step'], d['roofline']['phase_ms'], d['parity']['max_norm_rel_err'])"
for halo in p2p symm; do
  timeout 300 $TR --master-port 29521 bench.py --gpus 2 --steps 100 --warmup 10 --no-cpu-baseline --halo $halo --slab-timing \
      > $out/bench_bar_${halo}.json 2> $out/bench_bar_${halo}.err
  grep "slab phase" $out/bench_bar_${halo}.err
done
timeout 300 $TR --master-port 29522 bench.py --gpus 2 --steps 100 --warmup 10 --no-cpu-baseline --workload 3d16m-blocks --slab-timing \
    > $out/bench_blocks_p2p.json 2> $out/bench_blocks_p2p.err
grep "slab phase" $out/bench_blocks_p2p.err
timeout 120 $TR --master-port 29541 scripts/peer_red_probe.py > $out/peer_red_probe.json 2> $out/peer_red_probe.err
tail -c 400 $out/peer_red_probe.json
python - <<PY
import glob, json
for f in sorted(glob.glob("$out/bench_b*.json")):
    try:
        d = json.load(open(f))
        print(f, round(d["ms_per_step"], 4), "%.4g" % d["value"], d["config"].get("slab_particles"), d["config"].get("migration"))
    except Exception as e:
        print(f, "failed:", e)
PY
