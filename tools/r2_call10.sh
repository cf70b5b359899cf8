#!/bin/bash
# round-2 GPU call 10: lw_lim benches (passive tracers on lw_lim), full size and sample
mkdir -p gpurun_out
run() { tag=$1; shift; envs=(); while [ "$1" != "--" ]; do envs+=("$1"); shift; done; shift
  env "${envs[@]}" timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e "$@" > gpurun_out/r2c10_$tag.json 2> gpurun_out/r2c10_$tag.err
  python - "$tag" <<'PY'
import json, sys
v = sys.argv[1]
try:
    d = json.loads([l for l in open("gpurun_out/r2c10_%s.json" % v) if l.startswith("{")][-1])
    ph = d["phases_ms_per_step"]
    print("%-10s step %.2f  TR %.2f  LWFLUX %.2f LWLIM %.2f MOMCOL %.2f  VMIX %.2f  STATE %.2f  FIN %.2f  SOLVER %.2f iters %s" % (v, d["ms_per_step"], ph["TRACER_UPDATE"], ph.get("LW_FLUX_VEL", 0), ph.get("ADVT_LW_LIM", 0), ph["MOMENTUM_COLUMN"], ph["VMIX_TRACER_IMPLICIT"], ph["STATE"], ph["MOMENTUM_FINISH"], ph["SOLVER"], d.get("solver_iterations")))
except Exception as e:
    print(v, "FAILED", e); import subprocess; print(subprocess.run(["tail", "-2", "gpurun_out/r2c10_%s.err" % v], capture_output=True, text=True).stdout)
PY
}
run s_nt8cen X=1 -- --workload tx_sample --nt 8
run s_nt8lw X=1 -- --workload tx_sample --nt 8 --passive-advect lw_lim
run s_nt8lwpbc X=1 -- --workload tx_sample --nt 8 --passive-advect lw_lim --pbc
run nt8lw X=1 -- --nt 8 --passive-advect lw_lim
run nt6lw X=1 -- --nt 6 --passive-advect lw_lim
