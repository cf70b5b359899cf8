#!/bin/bash
# round-2 GPU call 27: final state -- whole GPU suite, smoke(), profiling recipe (launch list + ncu), default bench line,
# reference arm
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -6 ) > gpurun_out/r2c27_pytest.log 2>&1
tail -5 gpurun_out/r2c27_pytest.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/r2c27_smoke.log 2>&1; tail -2 gpurun_out/r2c27_smoke.log
bash tools/ncu_kernels.sh r2final2 > gpurun_out/r2c27_ncu.log 2>&1; grep "rc=" gpurun_out/r2c27_ncu.log
( time timeout 1500 python bench.py ) > gpurun_out/r2c27_bench.json 2> gpurun_out/r2c27_bench.err; tail -c 600 gpurun_out/r2c27_bench.json; tail -3 gpurun_out/r2c27_bench.err
( time timeout 1500 python bench.py --impl reference ) > gpurun_out/r2c27_ref.json 2> gpurun_out/r2c27_ref.err; tail -c 400 gpurun_out/r2c27_ref.json; tail -3 gpurun_out/r2c27_ref.err
