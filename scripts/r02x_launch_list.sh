#!/bin/bash
# Round 2, call 30: ncu launch list of the bench command (final defaults), from the first launch on.
set -u
out=gpurun_out/r02x
mkdir -p $out
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $out/launches.csv \
    python bench.py --steps 8 --warmup 3 --no-cpu-baseline --no-parity --e2e-serial-only --e2e-steps 1 > $out/launches_run.log 2>&1
python scripts/launch_summary.py $out/launches.csv 3 > $out/launches_summary.txt 2>&1; cat $out/launches_summary.txt; wc -l $out/launches.csv
