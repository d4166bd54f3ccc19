#!/bin/bash
# Round 2, call 31: the reference's own numba path at the sizes SURVEY 8d lists besides the paper scene's
# (3D 262 144 synthetic, 2D 65 536), timed on the GPU box's host.  No GPU work.
set -u
out=gpurun_out/r02y
mkdir -p $out
timeout 200 python oracle/time_numba_reference.py 262144 2 > $out/numba_3d_262144.json 2> $out/numba_3d.err &
timeout 200 python oracle/time_numba_reference.py 2d 65536 3 > $out/numba_2d_65536.json 2> $out/numba_2d.err &
wait
cat $out/numba_3d_262144.json $out/numba_2d_65536.json
