#!/bin/bash
# round-2 GPU call 24 (2 GPUs): coupled strips bitwise; checksums at N = 1 and 2 on the full grid
mkdir -p gpurun_out
( timeout 600 python -m pytest tests/test_gpu_multirank.py tests/test_gpu_step.py -m gpu -q -k "(2- and (coupled or pcsi])) or test_step_coupled" 2>&1 | tail -4 ) > gpurun_out/r2c24_pytest.log 2>&1
tail -3 gpurun_out/r2c24_pytest.log
timeout 900 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r2c24_n1.json 2> gpurun_out/r2c24_n1.err
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29727 bench.py --gpus 2 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r2c24_n2.json 2> gpurun_out/r2c24_n2.err
python - <<'PY'
import json
for n in (1, 2):
    try:
        d = json.loads([l for l in open("gpurun_out/r2c24_n%d.json" % n) if l.startswith("{")][-1])
        print(n, "step %.2f e2e %.2f" % (d["ms_per_step"], d["e2e"]["ms_per_step"]), d["state_checksum"]["PSURF"], d["state_checksum_after_e2e"])
    except Exception as e:
        print(n, "FAILED", e)
PY
