#!/bin/bash
# round-2 GPU call 28: direct U1/V1 stores: coupled parity + end-to-end time (A/B)
mkdir -p gpurun_out
( timeout 600 python -m pytest tests/test_gpu_step.py -m gpu -q -k "coupled or variants" 2>&1 | tail -3 ) > gpurun_out/r2c28_pytest.log 2>&1; tail -2 gpurun_out/r2c28_pytest.log
for tag in direct nodirect; do
  if [ $tag = nodirect ]; then export POP_B200_NO_DIRECT_SFC=1; fi
  timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r2c28_$tag.json 2> gpurun_out/r2c28_$tag.err
  python - $tag <<'PY'
import json, sys
t = sys.argv[1]
try:
    d = json.loads([l for l in open("gpurun_out/r2c28_%s.json" % t) if l.startswith("{")][-1])
    print(t, "step %.2f e2e %.2f FIN %.2f" % (d["ms_per_step"], d["e2e"]["ms_per_step"], d["phases_ms_per_step"]["MOMENTUM_FINISH"]), d["state_checksum_after_e2e"]["UVEL"])
except Exception as e:
    print(t, "FAILED", e)
PY
done
