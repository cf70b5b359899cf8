#!/bin/bash
# Profiling recipe of this repo (run on the GPU box via gpurun): a plain run first, then
#   (1) the launch list (gpu__time_duration per launch of this library's kernels) of the same command,
#   (2) one `ncu --set full` capture of the column kernels (tx_sample workload),
#   (3) one `ncu --set full` capture of the P-CSI pass kernel at the full tx0.1v3 size.
#   tools/ncu_kernels.sh <out-tag>
TAG=${1:-prof}
OURS='regex:^(add_baro|avg_|bt_|dhdt|diag_|div_|grad_|halo_|impvmixt|momentum_|pguess|sfc_|state_|sum_|t2u|tracer_|vmix_|pcsi_)'
CMD="python bench.py --workload tx_sample --steps 1 --warmup 1 --no-cpu-baseline --no-e2e"
CMDF="python bench.py --workload tx0.1v3 --steps 1 --warmup 1 --no-cpu-baseline --no-e2e"
mkdir -p gpurun_out
$CMD > gpurun_out/plain_$TAG.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/plain_$TAG.log; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none -k "$OURS" -c 6000 --csv \
    --log-file gpurun_out/launches_$TAG.csv $CMD > gpurun_out/ncu_list_$TAG.log 2>&1
echo "ncu list rc=$?"
ncu --set full --clock-control none --import-source on -k "regex:tracer_fast|tracer_column|momentum_column|impvmixt_kernel|momentum_finish|state_3d" -s 5 -c 8 -f -o gpurun_out/$TAG $CMD > gpurun_out/ncu_$TAG.log 2>&1
echo "ncu cols rc=$?"
$CMDF > gpurun_out/plain_${TAG}_full.log 2>&1 || { echo "plain full run failed"; exit 1; }
ncu --set full --clock-control none --import-source on -k "regex:pcsi_iter2|pcsi_iter_kernel" -s 100 -c 2 -f -o gpurun_out/${TAG}_solver $CMDF > gpurun_out/ncu_${TAG}_solver.log 2>&1
echo "ncu solver rc=$?"
tail -c 800 gpurun_out/plain_$TAG.log; echo; tail -c 800 gpurun_out/plain_${TAG}_full.log
