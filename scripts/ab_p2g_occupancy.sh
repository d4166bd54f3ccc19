#!/bin/bash
# One-GPU A/B of the 20-warps/SM P2G (FFMPM_P2G_VARIANT=6) against the default, parity first.
set -u
out=gpurun_out/ab_occ
mkdir -p $out
FFMPM_P2G_VARIANT=6 timeout 50 python -m pytest tests/test_gpu_parity.py -x -q -k "large_block or multi_substep or material_layouts" > $out/pytest_v6.txt 2>&1
FFMPM_P2G_VARIANT=6 timeout 60 python bench.py --steps 100 --warmup 5 --no-cpu-baseline --e2e-steps 1 > $out/bench_v6.json 2> $out/bench_v6.err
timeout 60 python bench.py --steps 100 --warmup 5 --no-cpu-baseline --e2e-steps 1 > $out/bench_v5.json 2> $out/bench_v5.err
tail -2 $out/pytest_v6.txt
python - <<'PY'
import json
for v in ("v6", "v5"):
    try:
        d = json.load(open(f"gpurun_out/ab_occ/bench_{v}.json"))
        print(v, d["ms_per_step"], d["roofline"]["phase_ms"])
    except Exception as e:
        print(v, "failed", e)
PY
