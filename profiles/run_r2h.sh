#!/usr/bin/env bash
# Round-2h GPU pass (one GPU): validation of the N-pass rework -- the GPU suite (stop on failure),
# the default bench line, then the targeted ncu re-capture of the matcher
set -u
OUT=gpurun_out/r2h
mkdir -p $OUT
timeout 800 python -m pytest tests -m gpu -x -q > $OUT/pytest.log 2>&1; rc=$?; echo "pytest rc=$rc"; tail -12 $OUT/pytest.log
[ $rc -ne 0 ] && exit 1
SECONDS=0
timeout 600 python bench.py > $OUT/bench.json 2> $OUT/bench.err; echo "bench rc=$? in ${SECONDS}s"
python - <<'PY'
import json
d = json.load(open("gpurun_out/r2h/bench.json"))
print("value", d["value"], "e2e", d["e2e"]["value"], "auto", d["auto_mode"]["value"], "share", d["roofline"]["kernel_share_of_step"],
      "ms_match", d["roofline"]["kernel_ms_per_launch"], "wl3m", d["wl3m"]["value"], "kinnex", d["kinnex"]["value"])
PY
bash profiles/recapture_filtered.sh r2h
