#!/bin/bash
# one `ncu --set full` capture of selected kernels of a short bench run (after the same command ran clean)
#   tools/ncu_one.sh <out-tag> <kernel-regex> [skip] [count] [workload]
TAG=$1; RE=$2; SKIP=${3:-2}; CNT=${4:-2}; WL=${5:-tx_sample}
CMD="python bench.py --workload $WL --steps 1 --warmup 1 --no-cpu-baseline --no-e2e"
mkdir -p gpurun_out
$CMD > gpurun_out/plain_$TAG.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/plain_$TAG.log; exit 1; }
ncu --set full --clock-control none --import-source on -k "regex:$RE" -s $SKIP -c $CNT -f -o gpurun_out/$TAG $CMD > gpurun_out/ncu_$TAG.log 2>&1
echo "ncu rc=$?"; tail -2 gpurun_out/ncu_$TAG.log; tail -c 600 gpurun_out/plain_$TAG.log
