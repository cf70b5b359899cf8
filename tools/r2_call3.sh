#!/bin/bash
# round-2 GPU call 3: L2-prefetch-distance variants of the Thomas kernels + ncu of the base
mkdir -p gpurun_out
BENCH_ARGS="" bash tools/run_variants.sh tpd0 tpd8 tpd32 tpd16c8 mpd0 mpd8 mpd32 2>&1 | tee gpurun_out/r2c3_variants.log
bash tools/ncu_one.sh r2b "impvmixt_kernel|momentum_finish" 2 3 tx_sample
