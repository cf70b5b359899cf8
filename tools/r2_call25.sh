#!/bin/bash
# round-2 GPU call 25: compute-sanitizer memcheck on the kernels added this round (small cases)
mkdir -p gpurun_out
CS=/usr/local/cuda/bin/compute-sanitizer
( time timeout 1200 $CS --tool memcheck --error-exitcode 9 --print-limit 20 python -m pytest tests/test_gpu_ops.py tests/test_gpu_step.py tests/test_gpu_evp.py -m gpu -q -x -k "lw_lim_level or (sequence and (lw_lim_chrongear or pbc_tripole) and r16) or (deep_strip and 64) or (evp and pcsi)" 2>&1 | tail -25 ) > gpurun_out/r2c25_memcheck.log 2>&1
echo "memcheck rc=$?"; tail -12 gpurun_out/r2c25_memcheck.log
