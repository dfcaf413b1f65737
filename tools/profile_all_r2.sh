#!/usr/bin/env bash
# Regenerates the round-2 evidence under profiles/ on one B200 (run under gpurun; results land in gpurun_out/).
#   1. bench.py without a profiler (the numbers), the reference arm, then the ncu launch list of the same command
#   2. one `ncu --set full` capture of the dominant kernel of every BASELINE config, reduced on the box to a metric
#      summary (raw page) and a source-line hot-spot list (tools/ncu_hot.py); the .ncu-rep files are too big to bring back
#   3. compute-sanitizer memcheck + racecheck over small parity tests (every kernel family, TMA pipelines included)
set -u
out=gpurun_out
python bench.py > $out/r2_bench_n1.json 2> $out/r2_bench_n1.err
python bench.py --impl reference --steps 3 --warmup 1 > $out/r2_bench_reference_arm.json 2>> $out/r2_bench_n1.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $out/r2_launches_bench.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > $out/ncu_launches.log 2>&1
tools/prof_one.sh C5 k_stream 1 stream_inst_ct_terms python tools/configs_bench.py c5
tools/prof_one.sh C2 k_stream 1 stream_inst_ct_terms python tools/configs_bench.py c2
tools/prof_one.sh C1x k_stream 1 stream_inst_ct_root python tools/configs_bench.py c1x
tools/prof_one.sh C3 k_stream 1 stream_inst_ct_hist python tools/configs_bench.py c3pair
tools/prof_one.sh C3h k_stream 1 stream_inst_ct_hist python tools/configs_bench.py c3hist
tools/prof_one.sh C4 '^k_mterms$' 1 mterms python tools/configs_bench.py c4d
python tools/configs_bench.py c1,c1x,c2,c3,c4,c5 > $out/r2_configs.txt 2>&1
# (compute-sanitizer is closed on this GPU pool: profiles/r2_sanitizer_closed_on_pool.log is the tool's own refusal)
