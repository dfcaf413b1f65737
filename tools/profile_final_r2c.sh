#!/usr/bin/env bash
# Final refresh of the round-2 evidence after the last k_stream changes (predicate tables cached per segment, compile-time
# shapes for every leaf mask): GPU suite, bench lines, launch list, one ncu capture of the headline kernel, per-config table.
set -u
out=gpurun_out
( time python -m pytest tests -m gpu -q -x ) > $out/r2_gpu_tests.log 2>&1
python bench.py --steps 20 --warmup 5 > $out/r2_bench_n1.json 2> $out/r2_bench_n1.err
python bench.py --impl reference --steps 3 --warmup 1 > $out/r2_bench_reference_arm.json 2>> $out/r2_bench_n1.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $out/r2_launches_bench.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-per-config > $out/ncu_launches.log 2>&1
HOT_LINES=60 HOT_BY_LINE=1 tools/prof_one.sh C5 k_stream 1 stream_inst_ct_terms python tools/configs_bench.py c5
python tools/configs_bench.py c1,c1x,c2,c3,c4,c5 > $out/r2_configs.txt 2>&1
tail -3 $out/r2_gpu_tests.log; tail -c 300 $out/r2_bench_n1.err; cat $out/r2_configs.txt
