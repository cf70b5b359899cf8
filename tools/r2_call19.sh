#!/bin/bash
# round-2 GPU call 19 (2 GPUs): ghost depth by wave count on 300-row strips; pbc branches with shared reciprocals
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests/test_gpu_ops.py tests/test_gpu_step.py tests/test_gpu_multirank.py -m gpu -q -k "pbc or (2- and pcsi) or deep_strip" 2>&1 | tail -6 ) > gpurun_out/r2c19_pytest.log 2>&1
tail -4 gpurun_out/r2c19_pytest.log
show() {
  python - "$1" <<'PY'
import json, sys
v = sys.argv[1]
try:
    d = json.loads([l for l in open("gpurun_out/r2c19_%s.json" % v) if l.startswith("{")][-1])
    ph = d["phases_ms_per_step"]
    print("%-10s step %.2f  TR %.2f MOMCOL %.2f VMIX %.2f FIN %.2f SOLVER %.2f HALO %.2f iters %s" % (v, d["ms_per_step"], ph["TRACER_UPDATE"], ph["MOMENTUM_COLUMN"], ph["VMIX_TRACER_IMPLICIT"], ph["MOMENTUM_FINISH"], ph["SOLVER"], ph.get("HALO", 0), d.get("solver_iterations")))
except Exception as e:
    print(v, "FAILED", e)
PY
}
run2() { tag=$1; wl=$2; shift; shift; envs=(); while [ "$1" != "--" ]; do envs+=("$1"); shift; done; shift
  env "${envs[@]}" timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29719 bench.py --gpus 2 --workload $wl --steps 3 --warmup 3 --no-cpu-baseline --no-e2e "$@" > gpurun_out/r2c19_$tag.json 2> gpurun_out/r2c19_$tag.err; show $tag; }
run1() { tag=$1; shift; envs=(); while [ "$1" != "--" ]; do envs+=("$1"); shift; done; shift
  env "${envs[@]}" timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e "$@" > gpurun_out/r2c19_$tag.json 2> gpurun_out/r2c19_$tag.err; show $tag; }
run2 s_auto tx_2strips POP_B200_TRACE=1 --
run2 s_d12 tx_2strips POP_B200_DEEP_HALO_EXACT=1 --
run2 s_auto_noov tx_2strips POP_B200_NO_DEEP_OVERLAP=1 --
run2 s_d8 tx_2strips POP_B200_DEEP_HALO=8 POP_B200_DEEP_HALO_EXACT=1 --
run1 strip8 X=1 -- --workload tx_strip8
run1 pbc X=1 -- --pbc
