#!/bin/bash
# round-2 GPU call 16 (4 GPUs): multirank suite at world 2 and 4, scaling bench at N=4 and N=2
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests/test_gpu_multirank.py -m gpu -q 2>&1 | tail -6 ) > gpurun_out/r2c16_pytest.log 2>&1
tail -5 gpurun_out/r2c16_pytest.log
for n in 4 2; do
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29715 bench.py --gpus $n --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r2c16_n$n.json 2> gpurun_out/r2c16_n$n.err
  python - $n <<'PY'
import json, sys
n = sys.argv[1]
try:
    d = json.loads([l for l in open("gpurun_out/r2c16_n%s.json" % n) if l.startswith("{")][-1])
    ph = d["phases_ms_per_step"]
    print("n=%s step %.2f e2e %.2f SOLVER %.2f" % (n, d["ms_per_step"], d["e2e"]["ms_per_step"], ph["SOLVER"]))
except Exception as e:
    print(n, "FAILED", e)
PY
done
