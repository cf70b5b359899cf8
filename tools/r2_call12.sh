#!/bin/bash
# round-2 GPU call 12: lw_lim after the first optimisation pass: bench (sample, nt=8) + ncu of the lw kernels
mkdir -p gpurun_out
run() { tag=$1; shift; envs=(); while [ "$1" != "--" ]; do envs+=("$1"); shift; done; shift
  env "${envs[@]}" timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e "$@" > gpurun_out/r2c12_$tag.json 2> gpurun_out/r2c12_$tag.err
  python - "$tag" <<'PY'
import json, sys
v = sys.argv[1]
try:
    d = json.loads([l for l in open("gpurun_out/r2c12_%s.json" % v) if l.startswith("{")][-1])
    ph = d["phases_ms_per_step"]
    print("%-10s step %.2f  TR %.2f  MOMCOL %.2f  VMIX %.2f  STATE %.2f  FIN %.2f  SOLVER %.2f iters %s" % (v, d["ms_per_step"], ph["TRACER_UPDATE"], ph["MOMENTUM_COLUMN"], ph["VMIX_TRACER_IMPLICIT"], ph["STATE"], ph["MOMENTUM_FINISH"], ph["SOLVER"], d.get("solver_iterations")))
except Exception as e:
    print(v, "FAILED", e); import subprocess; print(subprocess.run(["tail", "-2", "gpurun_out/r2c12_%s.err" % v], capture_output=True, text=True).stdout)
PY
}
run s_nt8cen X=1 -- --workload tx_sample --nt 8
run s_nt8lw X=1 -- --workload tx_sample --nt 8 --passive-advect lw_lim
run s_nt8lwnofast POP_B200_NO_FAST_TRACER=1 -- --workload tx_sample --nt 8 --passive-advect lw_lim
CMD="python bench.py --workload tx_sample --nt 4 --passive-advect lw_lim --steps 1 --warmup 1 --no-cpu-baseline --no-e2e"
ncu --set full --clock-control none --import-source on -k "regex:lw_lim_kernel|lw_flux_kernel|tracer_fast_kernel|tracer_column_kernel" -s 4 -c 5 -f -o gpurun_out/r2d_lw $CMD > gpurun_out/ncu_r2d_lw.log 2>&1
echo "ncu rc=$?"; tail -2 gpurun_out/ncu_r2d_lw.log
