#!/bin/bash
# tuning aid (GPU box): bench each library variant built by tools/build_variant.py;  tools/run_variants.sh v1 v2 ...
L=pop2-cesm_b200/csrc/libpop_b200.so
cp $L /tmp/libpop_b200_base.so
for v in base "$@"; do
  if [ "$v" = base ]; then cp /tmp/libpop_b200_base.so $L; else cp pop2-cesm_b200/csrc/variants/libpop_b200_$v.so $L; fi
  timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e ${BENCH_ARGS} > gpurun_out/var_$v.json 2> gpurun_out/var_$v.err
  python - "$v" <<'PY'
import json, sys
v = sys.argv[1]
try:
    d = json.loads([l for l in open("gpurun_out/var_%s.json" % v) if l.startswith("{")][-1])
    ph = d["phases_ms_per_step"]
    print("%-10s step %.2f  TR %.2f  MOMCOL %.2f  VMIX %.2f  STATE %.2f  FIN %.2f  SOLVER %.2f" % (v, d["ms_per_step"], ph["TRACER_UPDATE"], ph["MOMENTUM_COLUMN"], ph["VMIX_TRACER_IMPLICIT"], ph["STATE"], ph["MOMENTUM_FINISH"], ph["SOLVER"]))
except Exception as e:
    print(v, "FAILED", e)
PY
done
cp /tmp/libpop_b200_base.so $L
