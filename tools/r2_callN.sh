#!/bin/bash
# final scaling line at N GPUs: tools/r2_callN.sh N
n=$1
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29729 bench.py --gpus $n --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r2final_n$n.json 2> gpurun_out/r2final_n$n.err
python - $n <<'PY'
import json, sys
n = sys.argv[1]
try:
    d = json.loads([l for l in open("gpurun_out/r2final_n%s.json" % n) if l.startswith("{")][-1])
    print(n, "step %.2f e2e %.2f SOLVER %.2f" % (d["ms_per_step"], d["e2e"]["ms_per_step"], d["phases_ms_per_step"]["SOLVER"]), d["state_checksum"]["PSURF"], d["state_checksum_after_e2e"]["PSURF"])
except Exception as e:
    print(n, "FAILED", e)
PY
