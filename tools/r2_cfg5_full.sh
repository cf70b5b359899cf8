#!/bin/bash
# BASELINE config 5 at full size: tx0.1v3, nt = 34 (T, S + 32 passive tracers on lw_lim), N GPUs: tools/r2_cfg5_full.sh N
n=$1
mkdir -p gpurun_out
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29731 bench.py --gpus $n --nt 34 --passive-advect lw_lim --steps 3 --warmup 2 --no-cpu-baseline --no-e2e > gpurun_out/r2_cfg5_n$n.json 2> gpurun_out/r2_cfg5_n$n.err
python - $n <<'PY'
import json, sys
n = sys.argv[1]
try:
    d = json.loads([l for l in open("gpurun_out/r2_cfg5_n%s.json" % n) if l.startswith("{")][-1])
    print(json.dumps({"workload": "tx0.1v3 3600x2400x62 nt=34, 32 passive tracers on lw_lim (config 5)", "n_gpus": int(n), "ms_per_step": round(d["ms_per_step"], 2), "solver_iterations": d.get("solver_iterations"), "phases_ms": {k: round(v, 2) for k, v in d["phases_ms_per_step"].items()}}))
except Exception as e:
    print(n, "FAILED", e)
PY
