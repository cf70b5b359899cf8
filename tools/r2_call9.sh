#!/bin/bash
# round-2 GPU call 9: full GPU suite with lw_lim, benches: default, nt=8 centred vs nt=8 with lw_lim passive tracers
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -8 ) > gpurun_out/r2c9_pytest.log 2>&1
tail -6 gpurun_out/r2c9_pytest.log
run() { tag=$1; shift; envs=(); while [ "$1" != "--" ]; do envs+=("$1"); shift; done; shift
  env "${envs[@]}" timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e "$@" > gpurun_out/r2c9_$tag.json 2> gpurun_out/r2c9_$tag.err
  python - "$tag" <<'PY'
import json, sys
v = sys.argv[1]
try:
    d = json.loads([l for l in open("gpurun_out/r2c9_%s.json" % v) if l.startswith("{")][-1])
    ph = d["phases_ms_per_step"]
    print("%-10s step %.2f  TR %.2f  LWFLUX %.2f LWLIM %.2f MOMCOL %.2f  VMIX %.2f  STATE %.2f  FIN %.2f  SOLVER %.2f iters %s" % (v, d["ms_per_step"], ph["TRACER_UPDATE"], ph.get("LW_FLUX_VEL", 0), ph.get("ADVT_LW_LIM", 0), ph["MOMENTUM_COLUMN"], ph["VMIX_TRACER_IMPLICIT"], ph["STATE"], ph["MOMENTUM_FINISH"], ph["SOLVER"], d.get("solver_iterations")))
except Exception as e:
    print(v, "FAILED", e)
PY
}
run base X=1 --
run nt8cen X=1 -- --nt 8
run nt8lw X=1 -- --nt 8 --passive-advect lw_lim
run nt8lwpbc X=1 -- --nt 8 --passive-advect lw_lim --pbc
