#!/usr/bin/env bash
# Targeted re-capture of the matcher's per-candidate instruction counts (one GPU):
#   gpurun -- 'bash profiles/recapture_filtered.sh r2h'   then here:
#   python profiles/update_inst_json.py gpurun_out/r2h/filtered_raw.csv gpurun_out/r2h/prof_plain.log r2h
set -u
TAG=${1:-r2h}
OUT=gpurun_out/$TAG
mkdir -p $OUT
timeout 300 python profiles/prof_filtered.py > $OUT/prof_plain.log 2> $OUT/prof_plain.err || { echo "plain run failed"; tail -5 $OUT/prof_plain.err; exit 1; }
timeout 500 ncu --set full --metrics smsp__inst_executed_pipe_alu.sum,smsp__inst_executed_pipe_fma.sum,smsp__inst_executed_pipe_lsu.sum,smsp__thread_inst_executed.sum \
    --clock-control none -k regex:nr_match_filtered_kernel --launch-count 2 -f -o $OUT/filtered python profiles/prof_filtered.py > $OUT/ncu.log 2>&1
echo "ncu rc=$?"; tail -2 $OUT/ncu.log
ncu -i $OUT/filtered.ncu-rep --page raw --csv > $OUT/filtered_raw.csv 2> /dev/null
rm -f $OUT/filtered.ncu-rep
ls -la $OUT
