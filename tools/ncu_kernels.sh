#!/bin/bash
# Profiling recipe of this repo (run on the GPU box via gpurun): a plain run first, then one
# `ncu --set full` capture of the column kernels of the same command.
#   tools/ncu_kernels.sh <out-tag> [workload] [kernel-regex]
TAG=${1:-prof}
WL=${2:-tx_sample}
RE=${3:-'tracer_column|momentum_column|impvmixt_kernel|momentum_finish'}
CMD="python bench.py --workload $WL --steps 2 --warmup 3 --no-cpu-baseline"
mkdir -p gpurun_out
$CMD > gpurun_out/plain_$TAG.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/plain_$TAG.log; exit 1; }
ncu --set full --clock-control none --import-source on -k "regex:$RE" -s 4 -c 5 -f -o gpurun_out/$TAG $CMD > gpurun_out/ncu_$TAG.log 2>&1
echo "ncu rc=$?"
tail -3 gpurun_out/ncu_$TAG.log
tail -c 1200 gpurun_out/plain_$TAG.log
