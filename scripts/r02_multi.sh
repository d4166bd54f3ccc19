#!/bin/bash
# Round-2 multi-GPU measurements (N = 2 rehearsal, N = 8 for the record):
#   coupled-bar weak scaling (BASELINE configs[3]) with phase tables, at N and at the smaller counts on the same box;
#   the decoupled blocks of round 1 for comparison; p2p vs symmetric-memory halo;
#   BASELINE configs[4]: 32 M-particle dam break, even cut vs re-cut, early (tall column) and late (spread) windows.
set -u
N=${1:-2}
out=gpurun_out/r02_n$N
mkdir -p $out
run() {  # run <n> <tag> <bench args...>
  local n=$1 tag=$2; shift 2
  timeout 420 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29600 + RANDOM % 300)) \
      bench.py --gpus $n --no-cpu-baseline "$@" > $out/$tag.json 2> $out/$tag.err
  grep -a -o "\[rank [0-9]\] slab phase ms/substep: [a-z0-9., ]*" $out/$tag.err > $out/$tag.phases.txt
  python - <<PY
import json
try:
    d = json.load(open("$out/$tag.json"))
    print("$tag", "ms/substep %.4f" % d["ms_per_step"], "value %.4g" % d["value"], "e2e ms %.2f" % d["e2e"]["ms_per_step"] if d.get("e2e") else "",
          d["config"].get("slab_particles"), d["config"].get("migration"), d["config"].get("rebalanced"))
except Exception as e:
    print("$tag FAILED:", e); print(open("$out/$tag.err", errors="replace").read()[-1500:])
PY
}
run $N bar_symm --steps 100 --warmup 10 --slab-timing
run $N bar_p2p --steps 100 --warmup 10 --halo p2p --slab-timing --e2e-serial-only --e2e-steps 1
run $N blocks_symm --steps 100 --warmup 10 --workload 3d16m-blocks --slab-timing --e2e-serial-only --e2e-steps 1
if [ "$N" = "8" ]; then
  run 4 bar_symm_n4 --steps 100 --warmup 10 --e2e-steps 2

  run 8 bar_symm_drift0.1 --steps 100 --warmup 10 --drift 0.1 --slab-timing --e2e-serial-only --e2e-steps 1
fi
DAM=dam32m
[ "$N" = "2" ] && DAM=dam:8388608
run $N dam_static_early --steps 100 --warmup 10 --workload $DAM
run $N dam_recut_early --steps 100 --warmup 10 --workload $DAM --rebalance
run $N dam_static_late --steps 100 --warmup 10 --workload $DAM-late
run $N dam_recut_late --steps 100 --warmup 10 --workload $DAM-late --rebalance
