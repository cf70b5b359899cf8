#!/bin/bash
# launch list only (gpu__time_duration per launch of this library's kernels) of one tx_sample step; the full recipe is tools/ncu_kernels.sh
OURS='regex:^(add_baro|avg_|bt_|dhdt|diag_|div_|grad_|halo_|impvmixt|momentum_|pguess|sfc_|state_|sum_|t2u|tracer_|vmix_|pcsi_|gm_|rf_|convad)'
CMD="python bench.py --workload tx_sample --steps 1 --warmup 1 --no-cpu-baseline --no-e2e"
$CMD > gpurun_out/plain_r1final.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/plain_r1final.log; exit 1; }
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k "$OURS" -c 6000 --csv --log-file gpurun_out/launches_r1final.csv $CMD > gpurun_out/ncu_list_r1final.log 2>&1
echo "ncu list rc=$?"; wc -l gpurun_out/launches_r1final.csv; tail -c 300 gpurun_out/plain_r1final.log
