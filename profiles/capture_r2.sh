#!/usr/bin/env bash
# Round-2 capture (one GPU, through gpurun):  gpurun --timeout 1500 -- 'bash profiles/capture_r2.sh'
#   1. plain run of profiles/prof_all.py (must exit 0 before anything is profiled)
#   2. ncu --set full over every launch of our kernels in that script (library sort/scan kernels
#      are left out by the -k filter); its raw page -> gpurun_out/r2/prof_all_raw.csv
# profiles/summarise_r2.py turns the report into profiles/r2_kernels.md and
# profiles/inst_per_candidate.json here on the CPU box.
set -u
OUT=gpurun_out/r2
mkdir -p $OUT
python profiles/prof_all.py > $OUT/prof_all_plain.log 2> $OUT/prof_all_plain.err || { echo "plain run failed"; tail -5 $OUT/prof_all_plain.err; exit 1; }
ncu --set full --metrics smsp__inst_executed_pipe_alu.sum,smsp__inst_executed_pipe_fma.sum,smsp__inst_executed_pipe_lsu.sum,smsp__thread_inst_executed.sum \
    --clock-control none \
    -k regex:'nr_pack_kernel|nr_match_|nr_deep_|k_cluster|k_large|k_rec_|k_part_|nr_hw_search' \
    -f -o $OUT/prof_all python profiles/prof_all.py > $OUT/prof_all_ncu.log 2>&1
tail -3 $OUT/prof_all_ncu.log
# the report itself is too large to travel (gpurun_out is capped at 64 MiB): keep the raw page
ncu -i $OUT/prof_all.ncu-rep --page raw --csv > $OUT/prof_all_raw.csv 2> /dev/null
ls -la $OUT/prof_all.ncu-rep $OUT/prof_all_raw.csv
rm -f $OUT/prof_all.ncu-rep
