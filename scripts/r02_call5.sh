#!/bin/bash
# Round-2 GPU call 5 (one B200): binned 2D pipeline, G2P with contiguous addressing + packed sums (8 vs 7 CTAs/SM),
# compute-sanitizer passes over smoke().
set -u
out=gpurun_out/r02e
mkdir -p $out
timeout 900 python -m pytest tests -m gpu -x -q --durations=5 > $out/pytest_gpu.txt 2>&1
tail -8 $out/pytest_gpu.txt
for mb in 8 7; do
  FFMPM_G2P_MINB=$mb timeout 200 python bench.py --steps 100 --warmup 5 --no-cpu-baseline --e2e-serial-only --e2e-steps 1 > $out/bench_minb$mb.json 2> $out/bench_minb$mb.err
  python -c "import json;d=json.load(open('$out/bench_minb$mb.json'));print('minb $mb', d['ms_per_step'], d['roofline']['phase_ms'], d['parity']['max_norm_rel_err'])"
done
timeout 200 python bench.py --steps 400 --warmup 10 --workload 2d1m --no-cpu-baseline --e2e-steps 1 > $out/bench_2d.json 2> $out/bench_2d.err
python -c "
import json
d=json.load(open('$out/bench_2d.json')); print('2d', d['ms_per_step'], d['roofline']['phase_ms'], d['config']['cuda_graph'], d['parity']['max_norm_rel_err'])"
FFMPM_OVERLAP=0 timeout 200 python bench.py --steps 400 --warmup 10 --workload 2d1m --no-cpu-baseline --no-parity --e2e-steps 1 > $out/bench_2d_nooverlap.json 2> $out/bench_2d_nooverlap.err
python -c "
import json
d=json.load(open('$out/bench_2d_nooverlap.json')); print('2d no overlap', d['ms_per_step'], d['roofline']['phase_ms'])"
timeout 200 python bench.py --steps 401 --warmup 10 --workload 2d1m --no-cpu-baseline --no-parity --e2e-steps 1 > $out/bench_2d_eager.json 2> $out/bench_2d_eager.err
python -c "
import json
d=json.load(open('$out/bench_2d_eager.json')); print('2d eager', d['ms_per_step'], d['config']['cuda_graph'])"
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -s 300 -c 60 --csv --log-file $out/launches_2d.csv \
    python bench.py --steps 21 --warmup 5 --workload 2d1m --no-cpu-baseline --no-parity --e2e-steps 1 > $out/launches_2d_run.log 2>&1
timeout 300 compute-sanitizer --tool memcheck --error-exitcode 3 python -c "import __graft_entry__ as g; g.smoke()" > $out/sanitizer_memcheck.txt 2>&1
echo "memcheck rc=$?"; tail -4 $out/sanitizer_memcheck.txt
timeout 400 compute-sanitizer --tool racecheck --error-exitcode 3 python -c "import __graft_entry__ as g; g.smoke()" > $out/sanitizer_racecheck.txt 2>&1
echo "racecheck rc=$?"; tail -4 $out/sanitizer_racecheck.txt
