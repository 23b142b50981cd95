#!/bin/bash
# Runs on the GPU box (under gpurun): timing of the lattice kernels at cfg5 (default, staging-only, compute-only), then one
# ncu --set full capture of each (forward, backward).  Outputs land in gpurun_out/.
set -u
TAG=${1:-r02}
python tools/time_kernels.py 333 1024 5 default env:FEO_DEBUG_MODE=1 env:FEO_DEBUG_MODE=2 2>&1 | grep cfg > gpurun_out/tk_lat_${TAG}.log || exit 1
cut -c1-140 gpurun_out/tk_lat_${TAG}.log
python tools/time_kernels.py 333 1024 1 > gpurun_out/plain_lat_${TAG}.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:residual_lattice -s 6 -c 2 -f -o gpurun_out/prof_lat_${TAG} python tools/time_kernels.py 333 1024 1 > gpurun_out/ncu_lat_${TAG}.log 2>&1
tail -3 gpurun_out/ncu_lat_${TAG}.log | cut -c1-200
