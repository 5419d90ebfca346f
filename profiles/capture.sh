#!/usr/bin/env bash
# Profile capture recipe (run on the GPU box through gpurun, one GPU):
#   gpurun --timeout 1200 -- 'bash profiles/capture.sh r1'
# 1. plain run of the exact bench command (must exit 0 before anything is profiled)
# 2. launch list: ncu --metrics gpu__time_duration.sum --clock-control none over the same command
# 3. one ncu --set full capture of the dominant kernel (nr_match_filtered_kernel)
# Outputs go to gpurun_out/; profiles/summarise.py turns them into the tracked files under
# profiles/ (launch share table, kernel summary, inst_per_candidate.json).
set -u
TAG=${1:-r1}
OUT=gpurun_out
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-dp-gcups"   # the default workload (4 Mi candidates per step)
mkdir -p $OUT
$CMD > $OUT/plain_$TAG.log 2> $OUT/plain_$TAG.err || { echo "plain run failed"; tail -5 $OUT/plain_$TAG.err; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv \
    --log-file $OUT/launches_$TAG.csv $CMD > $OUT/ncu_l_$TAG.log 2>&1
[ "${LIST_ONLY:-0}" = 1 ] && exit 0
ncu --set full --metrics smsp__inst_executed_pipe_alu.sum,smsp__inst_executed_pipe_fma.sum,smsp__inst_executed_pipe_lsu.sum,smsp__inst_executed_pipe_xu.sum,smsp__inst_executed_pipe_uniform.sum,smsp__thread_inst_executed.sum,sm__cycles_active.avg,sm__cycles_elapsed.avg \
    --clock-control none --import-source on -k regex:nr_match_filtered_kernel \
    --launch-skip 3 -c 1 -f -o $OUT/prof_filtered_$TAG $CMD > $OUT/ncu_f_$TAG.log 2>&1
ls -la $OUT | tail -8
