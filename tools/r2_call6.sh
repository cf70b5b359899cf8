#!/bin/bash
# round-2 GPU call 6: full GPU suite (incl. partial bottom cells, EVP, TMA-Thomas opt-in), pbc bench, PDL A/B at full size
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -8 ) > gpurun_out/r2c6_pytest.log 2>&1
tail -6 gpurun_out/r2c6_pytest.log
run() { # tag, env..., -- args
  tag=$1; shift
  envs=(); while [ "$1" != "--" ]; do envs+=("$1"); shift; done; shift
  env "${envs[@]}" timeout 400 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e "$@" > gpurun_out/r2c6_$tag.json 2> gpurun_out/r2c6_$tag.err
  python - "$tag" <<'PY'
import json, sys
v = sys.argv[1]
try:
    d = json.loads([l for l in open("gpurun_out/r2c6_%s.json" % v) if l.startswith("{")][-1])
    ph = d["phases_ms_per_step"]
    print("%-10s step %.2f  TR %.2f  MOMCOL %.2f  VMIX %.2f  STATE %.2f  FIN %.2f  SOLVER %.2f iters %s" % (v, d["ms_per_step"], ph["TRACER_UPDATE"], ph["MOMENTUM_COLUMN"], ph["VMIX_TRACER_IMPLICIT"], ph["STATE"], ph["MOMENTUM_FINISH"], ph["SOLVER"], d.get("solver_iterations")))
except Exception as e:
    print(v, "FAILED", e)
PY
}
run base X=1 --
run pdl POP_B200_PDL=1 --
run pbc X=1 -- --pbc
