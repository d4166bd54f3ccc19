#!/bin/bash
# call 13 (8 GPUs): coupled bar with the slack-based migration cadence (final weak-scaling numbers), N = 8 and N = 4
set -u
out=gpurun_out/r02m
mkdir -p $out
run() {
  local n=$1 tag=$2; shift 2
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29700 + RANDOM % 200)) \
      bench.py --gpus $n --no-cpu-baseline "$@" > $out/$tag.json 2> $out/$tag.err
  grep -a -o "\[rank [0-9]\] slab phase ms/substep: [a-z0-9., ]*" $out/$tag.err | sort > $out/$tag.phases.txt
  python -c "
import json
d = json.load(open('$out/$tag.json'))
print('$tag', 'ms/substep %.4f' % d['ms_per_step'], 'value %.4g' % d['value'], d['config'].get('slab_particles'), d['config'].get('migration'), 'e2e', d['e2e']['ms_per_step'])"
}
run 8 bar_symm --steps 100 --warmup 10 --slab-timing --e2e-serial-only --e2e-steps 1
run 4 bar_symm_n4 --steps 100 --warmup 10 --e2e-serial-only --e2e-steps 1
cat $out/bar_symm.phases.txt
