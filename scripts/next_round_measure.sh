#!/bin/bash
# First GPU calls of the next round (everything here was written after this round's GPU budget ran out).
#   gpurun --gpus 2 --timeout 600 -- 'bash scripts/next_round_measure.sh 2'
#   gpurun --gpus 8 --timeout 900 -- 'bash scripts/next_round_measure.sh 8'
#   gpurun --timeout 600 -- 'bash scripts/next_round_measure.sh 1'     (packed-fp32 P2G variants: parity, then A/B)
set -u
N=${1:-2}
out=gpurun_out/next_n$N
mkdir -p $out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
if [ "$N" = "1" ]; then
  FFMPM_TEST_EXPERIMENTAL=1 timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -k "packed_fp32" > $out/pytest_packed.txt 2>&1
  tail -3 $out/pytest_packed.txt
  for v in 5 7 10 8 11 12 9; do
    FFMPM_P2G_VARIANT=$v timeout 120 python bench.py --steps 100 --warmup 5 --no-cpu-baseline --e2e-steps 1 \
        > $out/bench_v$v.json 2> $out/bench_v$v.err
  done
  timeout 180 python bench.py --steps 100 --warmup 5 --no-cpu-baseline --e2e-streamed > $out/bench_e2e_streamed.json 2> $out/bench_e2e_streamed.err
  timeout 120 python bench.py --steps 100 --warmup 6 --no-cpu-baseline --e2e-steps 1 --graph > $out/bench_graph.json 2> $out/bench_graph.err
  for v in 8 11; do   # packed stress: economised coefficients (5 products instead of 7 here), and the left form (4, none with F)
    FFMPM_P2G_VARIANT=$v FFMPM_FP32_STRESS=2 timeout 120 python bench.py --steps 100 --warmup 5 --no-cpu-baseline --e2e-steps 1 \
        > $out/bench_v${v}econ.json 2> $out/bench_v${v}econ.err
    FFMPM_P2G_VARIANT=$v FFMPM_FP32_STRESS=3 timeout 120 python bench.py --steps 100 --warmup 5 --no-cpu-baseline --e2e-steps 1 \
        > $out/bench_v${v}left.json 2> $out/bench_v${v}left.err
  done
  FFMPM_G2P_PACKED=1 timeout 120 python bench.py --steps 100 --warmup 5 --no-cpu-baseline --e2e-steps 1 \
      > $out/bench_vg2p.json 2> $out/bench_vg2p.err
  FFMPM_G2P_PACKED=2 timeout 120 python bench.py --steps 100 --warmup 5 --no-cpu-baseline --e2e-steps 1 \
      > $out/bench_vg2p6.json 2> $out/bench_vg2p6.err
  python - <<PY
import json
for v in (5, 7, 10, 8, 11, 12, 9, "8econ", "11econ", "8left", "11left", "g2p", "g2p6"):
    try:
        d = json.load(open("$out/bench_v%s.json" % v))
        print("variant", v, d["ms_per_step"], d["roofline"]["phase_ms"])
    except Exception as e:
        print("variant", v, "failed:", e)
PY
  # dense cells (column settling): long runs leave phase-2 lanes idle; run cap on the packed kernel
  for cfgs in "5:0" "7:0" "7:8" "7:16"; do
    v=${cfgs%%:*}; cap=${cfgs##*:}
    FFMPM_P2G_VARIANT=$v FFMPM_P2G_RUN_CAP=$cap timeout 200 python bench.py --workload dam:8388608 --steps 50 --warmup 5 \
        --presteps 2000 --no-cpu-baseline > $out/dam8m_v${v}_cap${cap}.json 2> $out/dam8m_v${v}_cap${cap}.err
    python -c "import json;d=json.load(open('$out/dam8m_v${v}_cap${cap}.json'));print('dam8m variant $v cap $cap', d['ms_per_step'])" || true
  done
  exit 0
fi
if [ "$N" = "2" ]; then
  # SymmHalo on hardware: parity first (signals time out after 20 s instead of hanging), then p2p vs symm
  FFMPM_TEST_SYMM=1 timeout 180 python -m pytest tests/test_gpu_distributed.py -x -q -k "symm" > $out/pytest_symm.txt 2>&1
  tail -3 $out/pytest_symm.txt
  # if the flag protocol misbehaves on this torch build, the coarser barrier protocol isolates the copy path
  FFMPM_TEST_SYMM=1 FFMPM_SYMM_SYNC=barrier timeout 180 python -m pytest tests/test_gpu_distributed.py -x -q -k "symm" > $out/pytest_symm_barrier.txt 2>&1
  tail -3 $out/pytest_symm_barrier.txt
  # vector REDs at peer memory over NVLink: the precondition for P2G scattering halo planes into the neighbour's inbox
  timeout 120 $TR --master-port 29541 scripts/peer_red_probe.py > $out/peer_red_probe.json 2> $out/peer_red_probe.err
  tail -c 600 $out/peer_red_probe.json
fi
for halo in p2p symm; do
  timeout 300 $TR --master-port 2952$N bench.py --gpus $N --steps 100 --warmup 10 --no-cpu-baseline --halo $halo --slab-timing \
      > $out/bench_${halo}.json 2> $out/bench_${halo}.err
done
# the p2p exchange moves 2 x (2*margin+2) planes of 1.06 MB per neighbour: 21 MB in ~0.18 ms looks bandwidth- as much as
# latency-bound for NCCL send/recv, so a thinner margin (more frequent, still asynchronous, migration checks) may pay
for m in 2 3; do
  timeout 300 $TR --master-port 2954$N bench.py --gpus $N --steps 100 --warmup 10 --no-cpu-baseline --margin $m \
      > $out/bench_p2p_margin$m.json 2> $out/bench_p2p_margin$m.err
done
if [ "$N" = "8" ]; then
  for mode in "" "--rebalance"; do
    tag=$([ -z "$mode" ] && echo static || echo rebalanced)
    timeout 300 $TR --master-port 2953$N bench.py --gpus $N --steps 100 --warmup 10 --workload dam32m --no-cpu-baseline $mode \
        > $out/dam32m_${tag}.json 2> $out/dam32m_${tag}.err
  done
fi
python - <<PY
import glob, json
for f in sorted(glob.glob("$out/*.json")):
    try:
        d = json.load(open(f))
        print(f, d["ms_per_step"], d["value"], d["config"].get("slab_particles"))
    except Exception as e:
        print(f, "failed:", e)
PY
