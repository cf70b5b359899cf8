#!/bin/bash
# round-2 GPU call 1: full GPU test suite, default bench (with the parity check), finish-kernel A/B, ncu of the Thomas kernels
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm --format=csv,noheader
( time timeout 1500 python -m pytest tests -m gpu -x -q -rP 2>&1 | tail -40 ) > gpurun_out/r2c1_pytest.log 2>&1
tail -5 gpurun_out/r2c1_pytest.log
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/r2c1_bench.json 2> gpurun_out/r2c1_bench.err
echo "bench rc=$?"; tail -c 1500 gpurun_out/r2c1_bench.json; tail -3 gpurun_out/r2c1_bench.err
for v in FUSED OVERLAP NOOVERLAP; do
  case $v in FUSED) E="";; OVERLAP) E="POP_B200_OVERLAP_FINISH=1";; NOOVERLAP) E="POP_B200_NO_OVERLAP=1";; esac
  env $E timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/r2c1_ab_$v.json 2> gpurun_out/r2c1_ab_$v.err
  python - "$v" <<'PY'
import json, sys
v = sys.argv[1]
try:
    d = json.loads([l for l in open("gpurun_out/r2c1_ab_%s.json" % v) if l.startswith("{")][-1])
    ph = d["phases_ms_per_step"]
    print("%-10s step %.2f  TR %.2f  MOMCOL %.2f  VMIX %.2f  STATE %.2f  FIN %.2f  SOLVER %.2f" % (v, d["ms_per_step"], ph["TRACER_UPDATE"], ph["MOMENTUM_COLUMN"], ph["VMIX_TRACER_IMPLICIT"], ph["STATE"], ph["MOMENTUM_FINISH"], ph["SOLVER"]))
except Exception as e:
    print(v, "FAILED", e)
PY
done
bash tools/ncu_one.sh r2a "impvmixt_kernel|momentum_finish" 2 3 tx_sample
