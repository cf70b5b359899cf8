#!/bin/bash
# round-2 GPU call 17: final checks -- whole GPU suite, smoke(), default bench line, reference arm
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -6 ) > gpurun_out/r2c17_pytest.log 2>&1
tail -5 gpurun_out/r2c17_pytest.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/r2c17_smoke.log 2>&1; tail -2 gpurun_out/r2c17_smoke.log
( time timeout 1500 python bench.py ) > gpurun_out/r2c17_bench.json 2> gpurun_out/r2c17_bench.err; tail -c 1500 gpurun_out/r2c17_bench.json; tail -4 gpurun_out/r2c17_bench.err
( time timeout 1500 python bench.py --impl reference ) > gpurun_out/r2c17_ref.json 2> gpurun_out/r2c17_ref.err; tail -c 1200 gpurun_out/r2c17_ref.json; tail -4 gpurun_out/r2c17_ref.err
