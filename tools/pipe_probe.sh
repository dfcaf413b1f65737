for gs in ""; do
  for i in 1 2; do
  TAGG_STREAM_GS=$gs python bench.py --steps 20 --warmup 5 --no-per-config --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads([l for l in sys.stdin if l.startswith('{')][-1])
print('GS=$gs', 'seq', round(d['ms_per_step'],3), 'k', round(d['kernel_ms_per_step'],3), 'pipe', {k:round(v,3) for k,v in d['pipelined'].items() if 'ms' in k})
"
  done
done
