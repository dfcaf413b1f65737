#!/usr/bin/env bash
# Scaling table on one 8-GPU box: C2 (bench contract workload, weak scaling) and C5 (strong scaling) at N = 1, 2, 4, 8.
# Usage (under gpurun --gpus 8): bash tools/scaling_run.sh > gpurun_out/scaling.jsonl
set -u
port=29600
for wl in c2 c5; do
  for n in 1 2 4 8; do
    port=$((port + 1))
    if [ "$n" = 1 ]; then
      python bench.py --workload $wl --gpus 1 --steps 20 --warmup 3 --no-cpu-baseline 2>/dev/null | tail -1
    else
      python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $port bench.py --workload $wl --gpus $n --steps 20 --warmup 3 --no-cpu-baseline 2>/dev/null | tail -1
    fi
  done
done
