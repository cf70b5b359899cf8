#!/bin/bash
# round-2 GPU call 13: the profiling recipe at the final kernels (launch list, ncu --set full of the column kernels on
# tx_sample, of the P-CSI pass at full size), after plain runs of the same commands
bash tools/ncu_kernels.sh r2final
