#!/usr/bin/env bash
# Regenerates the evidence under profiles/ on one B200 (run under gpurun; results land in gpurun_out/).
#   1. bench.py without a profiler (the number), then its ncu launch list (per-launch durations)
#   2. one `ncu --set full` capture of the dominant kernel of every BASELINE config, reduced on the box to a metric
#      summary (raw page) and a source-line hot-spot list (tools/ncu_hot.py); the .ncu-rep files are too big to bring back
#   3. kernel time / algorithmic GB/s of all five configs (tools/configs_bench.py)
set -u
out=gpurun_out
python bench.py --steps 20 --warmup 3 > $out/r1_bench_n1.json 2> $out/r1_bench_n1.err
python bench.py --impl reference --steps 3 --warmup 1 > $out/r1_bench_reference_arm.json 2>> $out/r1_bench_n1.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $out/r1_launches_bench.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > $out/ncu_launches.log 2>&1
cap() {  # name, kernel regex, launch-skip, cubin, command...
  local name=$1 rx=$2 skip=$3 cub=$4; shift 4
  ncu --set full --clock-control none --import-source on -k "regex:$rx" -s $skip -c 1 -f -o $out/prof_$name "$@" > $out/ncu_$name.log 2>&1
  python tools/ncu_hot.py $out/prof_$name.ncu-rep ${rx//[^a-z_]/} $cub 40 > $out/r1_ncu_hot_$name.txt 2>&1
  ncu -i $out/prof_$name.ncu-rep --page raw --csv > $out/r1_ncu_raw_$name.csv 2>&1
  rm -f $out/prof_$name.ncu-rep
}
cap C2 k_stream 3 stream python bench.py --steps 2 --warmup 1 --no-cpu-baseline
cap C1x k_stream 1 stream python tools/configs_bench.py c1x
cap C3 k_stream 1 stream python tools/configs_bench.py c3pair
cap C4 '^k_mterms$' 1 mterms python tools/configs_bench.py c4
cap C5 k_stream 1 stream python tools/configs_bench.py c5
python tools/configs_bench.py c1,c1x,c3,c4,c5 > $out/r1_configs.txt 2>&1
