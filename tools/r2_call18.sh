#!/bin/bash
# round-2 GPU call 18 (2 GPUs): exchange overlapped with the interior tiles: bitwise tests + A/B
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests/test_gpu_multirank.py -m gpu -q -k "2- and (pcsi or lwlim or pbc)" 2>&1 | tail -6 ) > gpurun_out/r2c18_pytest.log 2>&1
tail -4 gpurun_out/r2c18_pytest.log
show() {
  python - "$1" <<'PY'
import json, sys
v = sys.argv[1]
try:
    d = json.loads([l for l in open("gpurun_out/r2c18_%s.json" % v) if l.startswith("{")][-1])
    ph = d["phases_ms_per_step"]
    print("%-10s step %.2f  SOLVER %.2f HALO %.2f iters %s" % (v, d["ms_per_step"], ph["SOLVER"], ph.get("HALO", 0), d.get("solver_iterations")))
except Exception as e:
    print(v, "FAILED", e)
PY
}
run2() { tag=$1; wl=$2; shift; shift; envs=(); while [ "$1" != "--" ]; do envs+=("$1"); shift; done; shift
  env "${envs[@]}" timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29717 bench.py --gpus 2 --workload $wl --steps 3 --warmup 3 --no-cpu-baseline --no-e2e "$@" > gpurun_out/r2c18_$tag.json 2> gpurun_out/r2c18_$tag.err; show $tag; }
run2 ov tx0.1v3 X=1 --
run2 noov tx0.1v3 POP_B200_NO_DEEP_OVERLAP=1 --
run2 s_ov tx_2strips X=1 --
run2 s_noov tx_2strips POP_B200_NO_DEEP_OVERLAP=1 --
run2 s_ov22 tx_2strips POP_B200_DEEP_HALO=22 --
