#!/bin/bash
# Runs on the GPU box (under gpurun): plain timing of the preconditioner GEMM, the launch list of one apply (pre-pass,
# tensor-core kernel, loss finalize), then one ncu --set full capture of the tcgen05 kernel (tensor-pipe utilisation).
# Outputs land in gpurun_out/.
set -u
TAG=${1:-r02}
CMD="python tools/time_dense.py 2549 1024 20"
$CMD > gpurun_out/dense_plain_${TAG}.log 2>&1 || exit 1
FEO_DENSE_GEN=1 $CMD >> gpurun_out/dense_plain_${TAG}.log 2>&1
FEO_DENSE_SIMT=1 $CMD >> gpurun_out/dense_plain_${TAG}.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"dense_|finalize" -s 15 -c 9 --csv --log-file gpurun_out/dense_launches_${TAG}.csv $CMD > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:dense_apply_tc -s 6 -c 1 -o gpurun_out/prof_dense_${TAG} $CMD > gpurun_out/ncu_dense_${TAG}.log 2>&1
ncu -i gpurun_out/prof_dense_${TAG}.ncu-rep --page raw --csv > gpurun_out/prof_dense_${TAG}_raw.csv 2>/dev/null
ncu -i gpurun_out/prof_dense_${TAG}.ncu-rep --page source --csv > gpurun_out/prof_dense_${TAG}_src.csv 2>/dev/null
cat gpurun_out/dense_plain_${TAG}.log
grep -v "^==" gpurun_out/dense_launches_${TAG}.csv | cut -d, -f5,15- | cut -c1-160
tail -2 gpurun_out/ncu_dense_${TAG}.log | cut -c1-200
