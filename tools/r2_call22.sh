#!/bin/bash
# round-2 GPU call 22 (2 GPUs): hunt for the decomposition dependence the full-size state checksums show
mkdir -p gpurun_out
for wl in tx_sample; do
  POP_BENCH_DEBUG=1 timeout 300 python bench.py --workload $wl --steps 1 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/r2c22_${wl}_1.json 2> gpurun_out/r2c22_${wl}_1.err
  grep DEBUG gpurun_out/r2c22_${wl}_1.err
  POP_BENCH_DEBUG=1 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29723 bench.py --gpus 2 --workload $wl --steps 1 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/r2c22_${wl}_2.json 2> gpurun_out/r2c22_${wl}_2.err
  grep DEBUG gpurun_out/r2c22_${wl}_2.err
  POP_BENCH_DEBUG=1 POP_B200_NO_DEEP_HALO=1 POP_B200_SYNC_CHECKS=1 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29723 bench.py --gpus 2 --workload $wl --steps 1 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/r2c22_${wl}_2p.json 2> gpurun_out/r2c22_${wl}_2p.err
  echo "plain layout, sync checks:"; grep DEBUG gpurun_out/r2c22_${wl}_2p.err
done
