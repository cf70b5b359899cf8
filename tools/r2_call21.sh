#!/bin/bash
# round-2 GPU call 21 (8 GPUs): final scaling line at N=8 (with e2e), world-8 parity of the P-CSI strips
mkdir -p gpurun_out
( time timeout 600 python -m pytest tests/test_gpu_multirank.py -m gpu -q -k "8-pcsi]" 2>&1 | tail -4 ) > gpurun_out/r2c21_pytest.log 2>&1
tail -3 gpurun_out/r2c21_pytest.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29721 bench.py --gpus 8 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r2c21_n8.json 2> gpurun_out/r2c21_n8.err
python - <<'PY'
import json
try:
    d = json.loads([l for l in open("gpurun_out/r2c21_n8.json") if l.startswith("{")][-1])
    ph = d["phases_ms_per_step"]
    print("n=8 step %.2f e2e %.2f SOLVER %.2f TR %.2f MOMCOL %.2f VMIX %.2f FIN %.2f STATE %.2f HALO %.2f" % (d["ms_per_step"], d["e2e"]["ms_per_step"], ph["SOLVER"], ph["TRACER_UPDATE"], ph["MOMENTUM_COLUMN"], ph["VMIX_TRACER_IMPLICIT"], ph["MOMENTUM_FINISH"], ph["STATE"], ph["HALO"]))
except Exception as e:
    print("FAILED", e)
PY
