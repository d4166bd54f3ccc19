#!/bin/bash
# The driver's own launch lines at N = 2 (our arm and the reference arm under torchrun), the 1000-substep drift report
# with the left-form stress, and the full paper scene (3500 substeps) through the headless runner.
set -u
out=gpurun_out/r02o
mkdir -p $out
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29801 bench.py --impl reference --gpus 2 --steps 5 --warmup 1 > $out/scale2_reference.json 2> $out/scale2_reference.err
python -c "import json;d=json.load(open('$out/scale2_reference.json'));print('reference arm under torchrun:', d['value'], 'cores', d['cpu_baseline']['cores'])"
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29802 bench.py --gpus 2 --steps 20 --warmup 3 > $out/scale2.json 2> $out/scale2.err
python -c "import json;d=json.load(open('$out/scale2.json'));print('scale N=2 (driver line):', d['ms_per_step'], d['value'], d['e2e']['ms_per_step'], d['config']['migration'], d['cpu_baseline'] is not None)"
timeout 400 python scripts/drift_report.py --steps 1000 --out $out/drift_1000_substeps.json > $out/drift.log 2>&1; tail -4 $out/drift.log
timeout 300 python -m femflow_b200.simulation.mpm.headless --experiment 0 --no-save > $out/headless_c1.txt 2>&1; tail -2 $out/headless_c1.txt
timeout 300 python -m femflow_b200.simulation.mpm.headless --experiment 0 --outdir $out/c1_frames --steps 1000 > $out/headless_c1_saving.txt 2>&1; tail -1 $out/headless_c1_saving.txt; rm -rf $out/c1_frames
