#!/usr/bin/env bash
# Final-state refresh of the round-2 evidence (subset of profile_all_r2.sh: the shapes whose launch configuration
# changed after the full capture — C5 gained its third consumer group — plus the bench lines and the launch list).
set -u
out=gpurun_out
python bench.py --steps 20 --warmup 5 > $out/r2_bench_n1.json 2> $out/r2_bench_n1.err
python bench.py --impl reference --steps 3 --warmup 1 > $out/r2_bench_reference_arm.json 2>> $out/r2_bench_n1.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $out/r2_launches_bench.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-per-config > $out/ncu_launches.log 2>&1
tools/prof_one.sh C5 k_stream 1 stream_inst_ct_terms python tools/configs_bench.py c5
python tools/configs_bench.py c1,c1x,c2,c3,c4,c5 > $out/r2_configs.txt 2>&1
tail -c 300 $out/r2_bench_n1.err
