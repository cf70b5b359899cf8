#!/bin/bash
# Profiling recipe of this repo (run on the GPU box via gpurun): a plain run first, then
#   (1) the launch list (gpu__time_duration per launch) of the same command,
#   (2) one `ncu --set full` capture of the column kernels,
#   (3) one `ncu --set full` capture of a few solver-iteration kernels.
#   tools/ncu_kernels.sh <out-tag> [workload] [column-kernel-regex]
TAG=${1:-prof}
WL=${2:-tx_sample}
RE=${3:-'tracer_column|momentum_column|impvmixt_kernel|momentum_finish|state_3d'}
CMD="python bench.py --workload $WL --steps 1 --warmup 1 --no-cpu-baseline --no-e2e"
mkdir -p gpurun_out
$CMD > gpurun_out/plain_$TAG.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/plain_$TAG.log; exit 1; }
# launch list: skip the set-up + warm-up step launches is not possible by count alone (solver iteration
# counts vary), so list everything after the first 200 launches and let the summariser pick the last step
ncu --metrics gpu__time_duration.sum --clock-control none -s 200 -c 12000 --csv \
    --log-file gpurun_out/launches_$TAG.csv $CMD > gpurun_out/ncu_list_$TAG.log 2>&1
echo "ncu list rc=$?"
ncu --set full --clock-control none --import-source on -k "regex:$RE" -s 6 -c 7 -f -o gpurun_out/$TAG $CMD > gpurun_out/ncu_$TAG.log 2>&1
echo "ncu cols rc=$?"
ncu --set full --clock-control none --import-source on -k "regex:bt_|halo" -s 400 -c 8 -f -o gpurun_out/${TAG}_solver $CMD > gpurun_out/ncu_${TAG}_solver.log 2>&1
echo "ncu solver rc=$?"
tail -3 gpurun_out/ncu_$TAG.log
tail -c 1500 gpurun_out/plain_$TAG.log
