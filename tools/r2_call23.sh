#!/bin/bash
# round-2 GPU call 23 (2 GPUs): state checksums at N = 1 and N = 2 on the full tx0.1v3 grid (final scaling lines)
mkdir -p gpurun_out
timeout 900 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r2c23_n1.json 2> gpurun_out/r2c23_n1.err
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29725 bench.py --gpus 2 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r2c23_n2.json 2> gpurun_out/r2c23_n2.err
python - <<'PY'
import json
for n in (1, 2):
    try:
        d = json.loads([l for l in open("gpurun_out/r2c23_n%d.json" % n) if l.startswith("{")][-1])
        print(n, "step %.2f e2e %.2f" % (d["ms_per_step"], d["e2e"]["ms_per_step"]), d["state_checksum"], d["state_checksum_after_e2e"])
    except Exception as e:
        print(n, "FAILED", e)
PY
