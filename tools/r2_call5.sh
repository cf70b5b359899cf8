#!/bin/bash
# round-2 GPU call 5: TMA-staged Thomas kernels: tests, A/B against the register-ring kernels, ring-shape variants, ncu
mkdir -p gpurun_out
( timeout 900 python -m pytest tests/test_gpu_ops.py tests/test_gpu_benchmark_shapes.py tests/test_gpu_step.py -m gpu -q -x 2>&1 | tail -6 ) > gpurun_out/r2c5_pytest.log 2>&1
tail -3 gpurun_out/r2c5_pytest.log
POP_B200_NO_THOMAS_TMA=1 timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/var_noTMA.json 2>/dev/null
python - <<'PY'
import json
d = json.loads([l for l in open("gpurun_out/var_noTMA.json") if l.startswith("{")][-1]); ph = d["phases_ms_per_step"]
print("regring    step %.2f  VMIX %.2f  FIN %.2f  SOLVER %.2f" % (d["ms_per_step"], ph["VMIX_TRACER_IMPLICIT"], ph["MOMENTUM_FINISH"], ph["SOLVER"]))
PY
BENCH_ARGS="" bash tools/run_variants.sh tv83 tv210 tv46 mt82 mt28 2>&1 | tee gpurun_out/r2c5_variants.log
bash tools/ncu_one.sh r2c "impvmixt_tma|momentum_finish_tma" 2 3 tx_sample
