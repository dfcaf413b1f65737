#!/usr/bin/env bash
# One `ncu --set full` capture of a kernel, reduced on the box to a metric summary and a source-line hot-spot list.
# usage: tools/prof_one.sh <name> <kernel regex> <launch-skip> <cubin> <command...>
set -u
out=gpurun_out
name=$1 rx=$2 skip=$3 cub=$4; shift 4
ncu --set full --clock-control none --import-source on -k "regex:$rx" -s $skip -c 1 -f -o $out/prof_$name "$@" > $out/ncu_$name.log 2>&1
python tools/ncu_hot.py $out/prof_$name.ncu-rep ${rx//[^a-z_]/} $cub ${HOT_LINES:-60} > $out/r2_ncu_hot_$name.txt 2>&1
ncu -i $out/prof_$name.ncu-rep --page raw --csv > $out/r2_ncu_raw_$name.csv 2>&1
rm -f $out/prof_$name.ncu-rep
