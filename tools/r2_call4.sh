#!/bin/bash
# round-2 GPU call 4: EVP GPU tests, base timing with the restructured Thomas kernels, EVP timing at tx0.1v3
mkdir -p gpurun_out
( timeout 900 python -m pytest tests/test_gpu_evp.py tests/test_gpu_ops.py tests/test_gpu_benchmark_shapes.py -m gpu -q 2>&1 | tail -6 ) > gpurun_out/r2c4_pytest.log 2>&1
tail -4 gpurun_out/r2c4_pytest.log
for v in diagonal evp; do
  timeout 400 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --precond $v > gpurun_out/r2c4_$v.json 2> gpurun_out/r2c4_$v.err
  python - "$v" <<'PY'
import json, sys
v = sys.argv[1]
try:
    d = json.loads([l for l in open("gpurun_out/r2c4_%s.json" % v) if l.startswith("{")][-1])
    ph = d["phases_ms_per_step"]
    print("%-10s step %.2f  TR %.2f  MOMCOL %.2f  VMIX %.2f  STATE %.2f  FIN %.2f  SOLVER %.2f iters %s" % (v, d["ms_per_step"], ph["TRACER_UPDATE"], ph["MOMENTUM_COLUMN"], ph["VMIX_TRACER_IMPLICIT"], ph["STATE"], ph["MOMENTUM_FINISH"], ph["SOLVER"], d.get("solver_iterations")))
except Exception as e:
    print(v, "FAILED", e)
PY
  tail -2 gpurun_out/r2c4_$v.err
done
python - <<'PY'
# iteration counts of the two runs
import json
for v in ("diagonal", "evp"):
    pass
PY
