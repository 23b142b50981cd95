#!/bin/bash
# Runs on the GPU box (under gpurun): plain bench, then the ncu launch list of the same command, then one
# full capture of the two fused kernels.  Outputs land in gpurun_out/ and are summarised into profiles/ by
# tools/summarise_profiles.py on the authoring box.
set -u
TAG=${1:-r01}
CMD="python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --no-train-step --no-parity --no-precond-gemm"
$CMD > gpurun_out/bench_prof_${TAG}.json 2> gpurun_out/bench_prof_${TAG}.err || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches_${TAG}.csv $CMD > gpurun_out/ncu_launches_${TAG}.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:residual_ -s 6 -c 2 -o gpurun_out/prof_${TAG} $CMD > gpurun_out/ncu_full_${TAG}.log 2>&1
tail -2 gpurun_out/ncu_full_${TAG}.log | cut -c1-200
