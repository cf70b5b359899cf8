#!/bin/bash
# round-2 GPU call 11 (8 GPUs): strips bitwise = single strip at world=8 (subset), scaling bench, config 5 at full size
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests/test_gpu_multirank.py -m gpu -q -k "8- and (pcsi- or pcsi] or lwlim or evp)" 2>&1 | tail -8 ) > gpurun_out/r2c11_pytest.log 2>&1
tail -6 gpurun_out/r2c11_pytest.log
show() {
  python - "$1" <<'PY'
import json, sys
v = sys.argv[1]
try:
    d = json.loads([l for l in open("gpurun_out/r2c11_%s.json" % v) if l.startswith("{")][-1])
    ph = d["phases_ms_per_step"]
    print("%-10s n=%d step %.2f e2e %s TR %.2f LWF %.2f LWL %.2f MOMCOL %.2f  VMIX %.2f  STATE %.2f  FIN %.2f  SOLVER %.2f HALO %.2f iters %s" % (v, d["n_gpus"], d["ms_per_step"], d.get("e2e", {}).get("ms_per_step"), ph["TRACER_UPDATE"], ph.get("LW_FLUX_VEL", 0), ph.get("ADVT_LW_LIM", 0), ph["MOMENTUM_COLUMN"], ph["VMIX_TRACER_IMPLICIT"], ph["STATE"], ph["MOMENTUM_FINISH"], ph["SOLVER"], ph.get("HALO", 0), d["config"].get("solver_iterations_per_step")))
except Exception as e:
    print(v, "FAILED", e); import subprocess; print(subprocess.run(["tail", "-3", "gpurun_out/r2c11_%s.err" % v], capture_output=True, text=True).stdout)
PY
}
runN() { n=$1; tag=$2; shift; shift; envs=(); while [ "$1" != "--" ]; do envs+=("$1"); shift; done; shift
  env "${envs[@]}" timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29713 bench.py --gpus $n --steps 5 --warmup 3 --no-cpu-baseline "$@" > gpurun_out/r2c11_$tag.json 2> gpurun_out/r2c11_$tag.err; show $tag; }
runN 8 n8 X=1 --
runN 8 n8plain POP_B200_NO_DEEP_HALO=1 -- --no-e2e
runN 8 n8cfg5 X=1 -- --nt 34 --passive-advect lw_lim --no-e2e --steps 3
