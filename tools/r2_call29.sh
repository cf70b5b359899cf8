#!/bin/bash
# round-2 GPU call 29 (2 GPUs): coupled strips with pinned buffers (direct STF read / U1,V1 store), bitwise
mkdir -p gpurun_out
( timeout 600 python -m pytest tests/test_gpu_multirank.py -m gpu -q -k "2-coupled or 2-pcsi]" 2>&1 | tail -12 ) > gpurun_out/r2c29_pytest.log 2>&1; tail -8 gpurun_out/r2c29_pytest.log
