#!/usr/bin/env bash
# Round-2c GPU pass (one GPU):  gpurun --timeout 900 -- 'bash profiles/run_r2c.sh'
#   word-directory variant of the filtered kernel: parity, A/B timing on the 3M-sized list and on
#   the 737K list, ncu of the main pass on the 3M-sized list
set -u
OUT=gpurun_out/r2c
mkdir -p $OUT
timeout 500 python -m pytest tests/test_gpu_dir.py "tests/test_gpu_match.py::test_dense_index_3M_sized_whitelist_vs_oracle" -x -q > $OUT/pytest.log 2>&1; echo "pytest rc=$?"; tail -15 $OUT/pytest.log
timeout 300 python tools/time_3m.py > $OUT/t3m_dir.log 2>&1; echo "3M dir rc=$?"; tail -2 $OUT/t3m_dir.log
NR_FILTER_DIR=0 timeout 300 python tools/time_3m.py > $OUT/t3m_rank.log 2>&1; echo "3M rank rc=$?"; tail -2 $OUT/t3m_rank.log
NR_FILTER_DIR=1 timeout 300 python tools/time_auto.py 4194304 0.1 1e-3 > $OUT/t737_dir.log 2>&1; echo "737K dir rc=$?"; grep "cand/s" $OUT/t737_dir.log
timeout 300 python tools/time_auto.py 4194304 0.1 1e-3 > $OUT/t737_rank.log 2>&1; echo "737K rank rc=$?"; grep "cand/s" $OUT/t737_rank.log
timeout 400 ncu --set full --clock-control none --import-source on -k regex:nr_match_filtered_kernel \
    --launch-skip 2 --launch-count 1 -f -o $OUT/f3m python tools/time_3m.py 1048576 > $OUT/ncu.log 2>&1
echo "ncu rc=$?"; tail -2 $OUT/ncu.log
ncu -i $OUT/f3m.ncu-rep --page raw --csv > $OUT/f3m_raw.csv 2> /dev/null
ls -la $OUT
