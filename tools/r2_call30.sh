#!/bin/bash
# round-2 GPU call 30: last check of the final tree -- whole GPU suite, smoke(), default bench line
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -4 ) > gpurun_out/r2c30_pytest.log 2>&1; tail -3 gpurun_out/r2c30_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -1
( time timeout 900 python bench.py ) > gpurun_out/r2c30_bench.json 2> gpurun_out/r2c30_bench.err
python - <<'PY'
import json
d = json.loads([l for l in open("gpurun_out/r2c30_bench.json") if l.startswith("{")][-1])
print("step %.2f e2e %.2f frac %.3f parity %s cpu %.3e" % (d["ms_per_step"], d["e2e"]["ms_per_step"], d["roofline"]["frac"], d["parity_check"]["ok"], d["cpu_baseline"]["value"]), d["state_checksum"]["T"], d["gpu_launches"], d["clocks"])
PY
