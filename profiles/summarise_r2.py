#!/usr/bin/env python
"""profiles/capture_r2.sh -> tracked summaries (run here, on the CPU box):

  python profiles/summarise_r2.py [gpurun_out/r2/prof_all_raw.csv] [gpurun_out/r2/prof_all_plain.log]

  profiles/r2_kernels.md            one row per launch of profiles/prof_all.py: duration, the pipe
                                    or memory figure that bounds it, its roofline fraction
  profiles/inst_per_candidate.json  per-candidate instruction counts of the bench step's matcher
                                    (main pass + N pass), the hash of the kernel sources they were
                                    captured with, and the work counters of that run
"""
import csv
import hashlib
import json
import os
import re
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
RAW = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "gpurun_out", "r2", "prof_all_raw.csv")
LOG = sys.argv[2] if len(sys.argv) > 2 else os.path.join(ROOT, "gpurun_out", "r2", "prof_all_plain.log")
HBM_PEAK = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"] \
    if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6650.0


def num(d, u, k):
    v = d.get(k, "")
    if v in ("", "n/a"):
        return None
    x = float(v.replace(",", ""))
    unit = u.get(k, "")
    return x * {"Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "usecond": 1e-3, "us": 1e-3, "msecond": 1.0, "ms": 1.0,
                "nsecond": 1e-6, "ns": 1e-6, "second": 1e3, "s": 1e3}.get(unit, 1.0)


def short(name):
    name = re.sub(r"^void ", "", name)
    name = name.replace("<unnamed>::", "")
    return re.sub(r"\(.*", "", name)


# launches in the order of prof_all.py, with the phase they belong to and the units they processed
PHASES = [
    ("flanks-filtered", ["nr_pack_kernel", "nr_match_filtered_kernel<0, 0>", "nr_match_filtered_kernel<0, 1>",
                         "nr_match_deep_kernel", "nr_match_deep_kernel", "nr_deep_finalize_kernel",
                         "nr_match_bitsliced_kernel",
                         "nr_match_filtered_kernel<1, 0>", "nr_match_filtered_kernel<1, 1>", "nr_match_deep_kernel",
                         "nr_match_deep_kernel", "nr_deep_finalize_kernel", "nr_match_bitsliced_kernel"]),
]


def main():
    rows = list(csv.reader(open(RAW)))
    H, U = rows[0], rows[1]
    u = dict(zip(H, U))
    log = open(LOG).read() if os.path.exists(LOG) else ""
    phases = re.findall(r"PHASE (\S+) ?(\d*)", log)
    counters = re.search(r"COUNTERS (\{.*\})", log)
    counters = eval(counters.group(1)) if counters else {}
    B = int(phases[0][1]) if phases else 1 << 22
    out = ["# ncu --set full over every kernel with a throughput claim (round 2)", "",
           "Captured by `profiles/capture_r2.sh` (`profiles/prof_all.py` under `ncu --set full --clock-control none`,",
           "one B200); durations are cold-cache, serialised, under the profiler: they rank kernels and give the",
           "per-launch counters, they are not bench values.  ALU % = `sm__inst_executed_pipe_alu` of peak;",
           "issue % = `smsp__issue_active`; thr/inst = active threads per warp instruction; DRAM = read + written.",
           f"HBM fraction = DRAM bytes / duration over the measured copy bandwidth ({HBM_PEAK:.0f} GB/s, MEASURED_PEAKS.json).",
           "", "Phases of the script, in launch order: " + ", ".join(f"{p} ({n})" for p, n in phases), "",
           "| # | kernel | grid x block | ms | ALU % | FMA % | LSU % | issue % | thr/inst | warps act. % | L1 hit % | L2 hit % | DRAM MB | HBM frac | regs | top stall |",
           "|---|---|---|---|---|---|---|---|---|---|---|---|---|---|---|---|"]
    stalls = [h for h in H if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio")]
    agg = {}
    for k, r in enumerate(rows[2:]):
        d = dict(zip(H, r))
        nm = short(d.get("Kernel Name", "?"))
        ms = num(d, u, "gpu__time_duration.sum")
        dram = (num(d, u, "dram__bytes_read.sum") or 0) + (num(d, u, "dram__bytes_write.sum") or 0)
        st = sorted(((num(d, u, s) or 0, s.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", ""))
                     for s in stalls if "selected" not in s), reverse=True)[:1]
        def g(key):
            v = num(d, u, key)
            return "" if v is None else f"{v:.1f}"
        out.append(
            f"| {k} | `{nm[:60]}` | {d.get('launch__grid_size', '?')} x {d.get('launch__block_size', '?')} | {ms:.3f} | "
            f"{g('sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active')} | "
            f"{g('sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active')} | "
            f"{g('sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active')} | "
            f"{g('smsp__issue_active.avg.pct_of_peak_sustained_active')} | "
            f"{g('smsp__thread_inst_executed_per_inst_executed.ratio')} | "
            f"{g('sm__warps_active.avg.pct_of_peak_sustained_active')} | {g('l1tex__t_sector_hit_rate.pct')} | "
            f"{g('lts__t_sector_hit_rate.pct')} | {dram / 1e6:.1f} | {dram / (ms * 1e-3) / 1e9 / HBM_PEAK:.3f} | "
            f"{d.get('launch__registers_per_thread', '?')} | {st[0][1] if st else ''} {st[0][0]:.1f} |")
        a = agg.setdefault((k, nm), d)
    # per-candidate figures of the bench step's matcher: first main-pass + first N-pass launch
    main_d = n_d = None
    for r in rows[2:]:
        d = dict(zip(H, r))
        nm = short(d.get("Kernel Name", ""))
        if nm.startswith("nr_match_filtered_kernel<0, 0>") and main_d is None:
            main_d = d
        if nm.startswith("nr_match_filtered_kernel<0, 1>") and n_d is None:
            n_d = d
    js = None
    if main_d is not None and n_d is not None:
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
            "source": f"profiles/r2_kernels.md (ncu --set full, {B} candidates per launch, p_n 1e-3; main pass + N pass)",
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
        p = os.path.join(ROOT, "profiles", "inst_per_candidate.json")
        json.dump(js, open(p, "w"), indent=1)
        out += ["", "## Bench step, per candidate (main pass + N pass)", "", "```json", json.dumps(js, indent=1), "```"]
    open(os.path.join(ROOT, "profiles", "r2_kernels.md"), "w").write("\n".join(out) + "\n")
    print("\n".join(out))


if __name__ == "__main__":
    main()
