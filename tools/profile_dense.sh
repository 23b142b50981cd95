#!/bin/bash
# Runs on the GPU box (under gpurun): plain timing of the preconditioner GEMM, then one ncu --set full capture of
# the tcgen05 kernel (tensor-pipe utilisation) and of the fp32 FMA comparison kernel.  Outputs land in gpurun_out/.
set -u
TAG=${1:-r01}
CMD="python tools/time_dense.py 2549 1024 20"
$CMD > gpurun_out/dense_plain_${TAG}.log 2>&1 || exit 1
FEO_DENSE_SIMT=1 $CMD >> gpurun_out/dense_plain_${TAG}.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:dense_apply_tc -s 6 -c 1 -o gpurun_out/prof_dense_${TAG} $CMD > gpurun_out/ncu_dense_${TAG}.log 2>&1
ncu -i gpurun_out/prof_dense_${TAG}.ncu-rep --page raw --csv > gpurun_out/prof_dense_${TAG}_raw.csv 2>/dev/null
cat gpurun_out/dense_plain_${TAG}.log
tail -2 gpurun_out/ncu_dense_${TAG}.log | cut -c1-200
