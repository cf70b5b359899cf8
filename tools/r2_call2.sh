#!/bin/bash
# round-2 GPU call 2: full GPU test suite, then the Thomas-kernel prefetch-depth variants (tools/build_variant.py)
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -q -rP 2>&1 | grep -E "passed|failed|FAILED|Error|sums\]|relerr field|real" | tail -60 ) > gpurun_out/r2c2_pytest.log 2>&1
tail -12 gpurun_out/r2c2_pytest.log
BENCH_ARGS="" bash tools/run_variants.sh t82 t83 t46 t28 t216 m82 m28 m44b 2>&1 | tee gpurun_out/r2c2_variants.log
