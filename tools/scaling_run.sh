#!/usr/bin/env bash
# Strong scaling of the headline workload (C5: 1e9 docs / 64 segments) on one box: N ranks, one per GPU.
# usage (under gpurun --gpus N): tools/scaling_run.sh N
set -u
n=$1
out=gpurun_out
if [ "$n" = 1 ]; then
  python bench.py --gpus 1 --steps 20 --warmup 5 --no-per-config --no-cpu-baseline > $out/r2_scale_n$n.json 2> $out/r2_scale_n$n.err
else
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $n --steps 20 --warmup 5 > $out/r2_scale_n$n.json 2> $out/r2_scale_n$n.err
fi
tail -c 600 $out/r2_scale_n$n.err
python - <<P
import json
d = json.loads([l for l in open("$out/r2_scale_n$n.json") if l.startswith("{")][-1])
print("N=$n ms/step %.3f kernel %.3f value %.4g e2e ms %.3f frac %.3f" % (d["ms_per_step"], d["kernel_ms_per_step"], d["value"], d["e2e"]["ms_per_step"], d["roofline"]["frac"]), d.get("parity", "")[:60])
P
