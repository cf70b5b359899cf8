#!/bin/bash
# round-2 GPU call 26: partial bottom cells after the fast-kernel / staged-DZU / Thomas changes: parity + bench
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests/test_gpu_ops.py tests/test_gpu_step.py -m gpu -q -k "pbc or partial_bottom or coupled" 2>&1 | tail -5 ) > gpurun_out/r2c26_pytest.log 2>&1
tail -4 gpurun_out/r2c26_pytest.log
run() { tag=$1; shift; envs=(); while [ "$1" != "--" ]; do envs+=("$1"); shift; done; shift
  env "${envs[@]}" timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e "$@" > gpurun_out/r2c26_$tag.json 2> gpurun_out/r2c26_$tag.err
  python - "$tag" <<'PY'
import json, sys
v = sys.argv[1]
try:
    d = json.loads([l for l in open("gpurun_out/r2c26_%s.json" % v) if l.startswith("{")][-1])
    ph = d["phases_ms_per_step"]
    print("%-10s step %.2f  TR %.2f  MOMCOL %.2f  VMIX %.2f  STATE %.2f  FIN %.2f  SOLVER %.2f iters %s" % (v, d["ms_per_step"], ph["TRACER_UPDATE"], ph["MOMENTUM_COLUMN"], ph["VMIX_TRACER_IMPLICIT"], ph["STATE"], ph["MOMENTUM_FINISH"], ph["SOLVER"], d.get("solver_iterations")))
except Exception as e:
    print(v, "FAILED", e); import subprocess; print(subprocess.run(["tail", "-2", "gpurun_out/r2c26_%s.err" % v], capture_output=True, text=True).stdout)
PY
}
run pbc X=1 -- --pbc
run pbcnofast POP_B200_NO_PBC_FAST=1 -- --pbc
run base X=1 --
