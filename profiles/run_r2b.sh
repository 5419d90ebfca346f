#!/usr/bin/env bash
# Round-2b GPU pass (one GPU, through gpurun):  gpurun --timeout 900 -- 'bash profiles/run_r2b.sh'
#   1. the GPU parity suite; 2. brute-force kernel timings, bit-parallel and DPX (tools/time_exhaustive.py);
#   3. ncu --set full of the bit-parallel kernel (737K and slide-seq launches) -> raw page csv
set -u
OUT=gpurun_out/r2b
mkdir -p $OUT
timeout 700 python -m pytest tests -m gpu -x -q > $OUT/pytest.log 2>&1; echo "pytest rc=$?"; tail -3 $OUT/pytest.log
timeout 300 python tools/time_exhaustive.py > $OUT/time_bs.log 2>&1; echo "bit-parallel rc=$?"; tail -4 $OUT/time_bs.log
NR_EXHAUSTIVE_DPX=1 timeout 300 python tools/time_exhaustive.py > $OUT/time_dpx.log 2>&1; echo "dpx rc=$?"; tail -4 $OUT/time_dpx.log
timeout 400 ncu --set full --clock-control none --import-source on -k regex:nr_match_bitsliced \
    --launch-skip 1 --launch-count 2 -f -o $OUT/bs python tools/time_exhaustive.py 296 2368 > $OUT/ncu.log 2>&1
echo "ncu rc=$?"; tail -2 $OUT/ncu.log
ncu -i $OUT/bs.ncu-rep --page raw --csv > $OUT/bs_raw.csv 2> /dev/null
ncu -i $OUT/bs.ncu-rep --page details --csv > $OUT/bs_details.csv 2> /dev/null
ls -la $OUT
