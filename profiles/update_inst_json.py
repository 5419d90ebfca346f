#!/usr/bin/env python
"""raw page of profiles/recapture_filtered.sh -> profiles/inst_per_candidate.json (same fields as
profiles/summarise_r2.py writes):  python profiles/update_inst_json.py <raw.csv> <plain.log> <tag>"""
import csv
import hashlib
import json
import os
import re
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from summarise_r2 import num, short  # noqa: E402


def main():
    raw, log, tag = sys.argv[1], sys.argv[2], sys.argv[3]
    rows = list(csv.reader(open(raw)))
    H, U = rows[0], rows[1]
    u = dict(zip(H, U))
    text = open(log).read()
    B = int(re.search(r"PHASE flanks-filtered (\d+)", text).group(1))
    counters = eval(re.search(r"COUNTERS (\{.*\})", text).group(1))
    main_d = n_d = None
    for r in rows[2:]:
        d = dict(zip(H, r))
        nm = short(d.get("Kernel Name", ""))
        if nm.startswith("nr_match_filtered_kernel<0, 0>") and main_d is None:
            main_d = d
        if nm.startswith("nr_match_filtered_kernel<0, 1>") and n_d is None:
            n_d = d
    assert main_d is not None and n_d is not None, "both passes must be in the capture"

    def tot(key, alt=None):
        s = 0.0
        for d in (main_d, n_d):
            v = num(d, u, key)
            if v is None and alt:
                v = num(d, u, alt)
            s += v or 0.0
        return s
    alu = tot("smsp__inst_executed_pipe_alu.sum", "sm__inst_executed_pipe_alu.sum")
    fma = tot("smsp__inst_executed_pipe_fma.sum", "sm__inst_executed_pipe_fma.sum")
    winst = tot("smsp__inst_executed.sum")
    tinst = tot("smsp__thread_inst_executed.sum")
    dram = tot("dram__bytes_read.sum") + tot("dram__bytes_write.sum")
    h = hashlib.sha256()
    for f in ("nr_match_filtered.cu", "nr_filter_core.h"):
        h.update(open(os.path.join(ROOT, "nanoranger_b200", "csrc", f), "rb").read())
    js = {"737K-5p": {
        "source": f"profiles/{tag}_filtered_recapture.md (ncu --set full, {B} candidates per launch, p_n 1e-3; main pass + N pass)",
        "kernel_source_hash": h.hexdigest()[:16],
        "candidates_per_launch": B,
        "alu_warp_inst_per_candidate": alu / B, "fma_warp_inst_per_candidate": fma / B,
        "warp_inst_per_candidate": winst / B, "thread_inst_per_candidate": tinst / B,
        "dram_bytes_per_candidate": dram / B,
        "alu_pipe_pct_of_peak_under_ncu_main_pass": num(main_d, u, "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active"),
        "alu_pipe_pct_of_peak_under_ncu_n_pass": num(n_d, u, "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active"),
        "ms_under_ncu": {"main_pass": num(main_d, u, "gpu__time_duration.sum"), "n_pass": num(n_d, u, "gpu__time_duration.sum")},
        "probes_per_candidate": counters.get("probes"), "verifications_per_candidate": counters.get("verifications"),
    }}
    json.dump(js, open(os.path.join(ROOT, "profiles", "inst_per_candidate.json"), "w"), indent=1)
    keys = ["gpu__time_duration.sum", "smsp__inst_executed.sum", "smsp__inst_executed_pipe_alu.sum",
            "smsp__thread_inst_executed_per_inst_executed.ratio",
            "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
            "sm__inst_issued.avg.pct_of_peak_sustained_active", "dram__bytes_read.sum", "dram__bytes_write.sum",
            "lts__t_sector_hit_rate.pct", "launch__registers_per_thread", "launch__grid_size"]
    out = [f"# {tag}: re-capture of the bench step's matcher after the N-pass rework", "",
           f"`profiles/recapture_filtered.sh` (`ncu --set full --clock-control none`, {B} candidates per launch).", "",
           "| metric | main pass `<0, 0>` | N pass `<0, 1>` |", "|---|---|---|"]
    for k in keys:
        out.append(f"| `{k}` ({u.get(k, '')}) | {main_d.get(k, '')} | {n_d.get(k, '')} |")
    out += ["", "```json", json.dumps(js, indent=1), "```"]
    open(os.path.join(ROOT, "profiles", f"{tag}_filtered_recapture.md"), "w").write("\n".join(out) + "\n")
    print("\n".join(out))


if __name__ == "__main__":
    main()
