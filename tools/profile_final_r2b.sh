#!/usr/bin/env bash
# End-of-round-2 evidence on one B200 (run under gpurun; results land in gpurun_out/): the full GPU parity suite, the bench
# lines (product arm with per_config, reference arm), the ncu launch list of the bench command, `ncu --set full` captures of
# the k_mterms shapes that changed last (C4 dense / hashed), and the per-config kernel table.
set -u
out=gpurun_out
( time python -m pytest tests -m gpu -q -x ) > $out/r2_gpu_tests.log 2>&1
python bench.py --steps 20 --warmup 5 > $out/r2_bench_n1.json 2> $out/r2_bench_n1.err
python bench.py --impl reference --steps 3 --warmup 1 > $out/r2_bench_reference_arm.json 2>> $out/r2_bench_n1.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $out/r2_launches_bench.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-per-config > $out/ncu_launches.log 2>&1
HOT_LINES=50 tools/prof_one.sh C4 '^k_mterms$' 1 mterms python tools/configs_bench.py c4d
HOT_LINES=50 tools/prof_one.sh C4h '^k_mterms$' 1 mterms python tools/configs_bench.py c4h
python tools/configs_bench.py c1,c1x,c2,c3,c4,c5 > $out/r2_configs.txt 2>&1
tail -3 $out/r2_gpu_tests.log; tail -c 300 $out/r2_bench_n1.err; cat $out/r2_configs.txt
