#!/bin/bash
# round-2 GPU call 8 (2 GPUs): P-CSI multirank tests after the side-stream checks, 2-GPU bench
mkdir -p gpurun_out
( time timeout 1200 python -m pytest tests/test_gpu_multirank.py tests/test_gpu_step.py -m gpu -q -k "(strips and (pcsi or gm or pbc)) or deep_strip" 2>&1 | tail -12 ) > gpurun_out/r2c8_pytest.log 2>&1
tail -8 gpurun_out/r2c8_pytest.log
show() {
  python - "$1" <<'PY'
import json, sys
v = sys.argv[1]
try:
    d = json.loads([l for l in open("gpurun_out/r2c8_%s.json" % v) if l.startswith("{")][-1])
    ph = d["phases_ms_per_step"]
    print("%-10s step %.2f  TR %.2f  MOMCOL %.2f  VMIX %.2f  STATE %.2f  FIN %.2f  SOLVER %.2f HALO %.2f iters %s" % (v, d["ms_per_step"], ph["TRACER_UPDATE"], ph["MOMENTUM_COLUMN"], ph["VMIX_TRACER_IMPLICIT"], ph["STATE"], ph["MOMENTUM_FINISH"], ph["SOLVER"], ph.get("HALO", 0), d.get("solver_iterations")))
except Exception as e:
    print(v, "FAILED", e)
PY
}
run1() { tag=$1; shift; envs=(); while [ "$1" != "--" ]; do envs+=("$1"); shift; done; shift
  env "${envs[@]}" timeout 400 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e "$@" > gpurun_out/r2c8_$tag.json 2> gpurun_out/r2c8_$tag.err; show $tag; }
run2() { tag=$1; shift; envs=(); while [ "$1" != "--" ]; do envs+=("$1"); shift; done; shift
  env "${envs[@]}" timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29711 bench.py --gpus 2 --steps 3 --warmup 3 --no-cpu-baseline --no-e2e "$@" > gpurun_out/r2c8_$tag.json 2> gpurun_out/r2c8_$tag.err; show $tag; }
run1 base1 X=1 --
run2 deep12 X=1 --
run2 sync POP_B200_SYNC_CHECKS=1 --
run2 plain POP_B200_NO_DEEP_HALO=1 --
