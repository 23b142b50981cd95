#!/bin/bash
# Runs on the GPU box (under gpurun): timing of the patch kernels at cfg5 (default, staging-only, compute-only), then one
# ncu --set full capture of each (forward, backward).  Outputs land in gpurun_out/.
set -u
TAG=${1:-r02}
python tools/time_kernels.py 333 1024 5 2>&1 | grep cfg > gpurun_out/tk_${TAG}.log
for m in 1 2; do FEO_DEBUG_MODE=$m python tools/time_kernels.py 333 1024 3 2>&1 | grep cfg | sed "s/^/mode $m: /" >> gpurun_out/tk_${TAG}.log; done
cat gpurun_out/tk_${TAG}.log
if [ "${2:-ncu}" = "ncu" ]; then
  ncu --set full --clock-control none --import-source on -k regex:residual_patch -s 6 -c 2 -f -o gpurun_out/prof_patch_${TAG} python tools/time_kernels.py 333 1024 1 > gpurun_out/ncu_patch_${TAG}.log 2>&1
  tail -3 gpurun_out/ncu_patch_${TAG}.log | cut -c1-200
fi
