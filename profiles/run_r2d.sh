#!/usr/bin/env bash
# Round-2d GPU pass (one GPU):  gpurun --timeout 1500 -- 'bash profiles/run_r2d.sh'
#   the GPU suite, smoke(), the default bench line, slide-seq AUTO with / without the deep tier, and
#   the ncu launch list of the bench command (after its plain run has exited 0)
set -u
OUT=gpurun_out/r2d
mkdir -p $OUT
timeout 800 python -m pytest tests -m gpu -x -q > $OUT/pytest.log 2>&1; echo "pytest rc=$?"; tail -3 $OUT/pytest.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/smoke.log 2>&1; echo "smoke rc=$?"; tail -1 $OUT/smoke.log
SECONDS=0
timeout 900 python bench.py > $OUT/bench.json 2> $OUT/bench.err; echo "bench rc=$? in ${SECONDS}s"; tail -c 600 $OUT/bench.json
timeout 300 python tools/time_slideseq.py 262144 > $OUT/slideseq.log 2>&1; echo "slideseq rc=$?"; cat $OUT/slideseq.log
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extras"
timeout 300 $CMD > $OUT/plain.json 2> $OUT/plain.err; echo "plain rc=$?"
timeout 500 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file $OUT/launches.csv $CMD > $OUT/ncu_l.log 2>&1
echo "ncu rc=$?"; tail -c 300 $OUT/ncu_l.log
ls -la $OUT
