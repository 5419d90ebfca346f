#!/usr/bin/env python
"""Turn the scratch files of profiles/capture.sh (gpurun_out/) into the tracked summaries:

  python profiles/summarise.py r1 [candidates_per_launch]

  profiles/<tag>_launches.md            per-kernel launch counts and time share of one bench run
  profiles/<tag>_filtered_kernel_ncu.md selected counters of the ncu --set full capture
  profiles/inst_per_candidate.json      per-candidate instruction counts bench.py multiplies by
                                        its live candidates/s to get roofline.achieved
"""
import csv
import io
import json
import os
import subprocess
import sys
from collections import defaultdict

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from extract_ncu import WANT  # noqa: E402

EXTRA = [
    "smsp__inst_executed_pipe_alu.sum", "smsp__inst_executed_pipe_fma.sum",
    "smsp__inst_executed_pipe_lsu.sum", "smsp__inst_executed_pipe_xu.sum",
    "smsp__inst_executed_pipe_uniform.sum", "smsp__inst_executed_pipe_cbu.sum",
    "smsp__inst_executed_pipe_adu.sum", "sm__inst_executed_pipe_alu.sum",
    "sm__inst_executed_pipe_fma.sum", "sm__inst_executed_pipe_lsu.sum",
    "sm__inst_executed_pipe_xu.sum", "smsp__thread_inst_executed.sum",
    "sm__cycles_elapsed.avg", "sm__cycles_active.avg", "smsp__cycles_active.avg",
    "lts__t_sectors.sum", "lts__t_bytes.sum", "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum",
    "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum", "sm__warps_active.avg.per_cycle_active",
    "launch__waves_per_multiprocessor", "sm__maximum_warps_per_active_cycle_pct",
    "l1tex__throughput.avg.pct_of_peak_sustained_active",
    "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed",
    "l1tex__m_l1tex2xbar_req_cycles_active.avg.pct_of_peak_sustained_elapsed",
    "sm__inst_executed_pipe_adu.avg.pct_of_peak_sustained_active",
    "idc__request_cycles_active.avg.pct_of_peak_sustained_elapsed",
]


def launches(tag):
    p = os.path.join(ROOT, "gpurun_out", f"launches_{tag}.csv")
    if not os.path.exists(p):
        return None
    lines = open(p).read().splitlines()
    start = next(i for i, ln in enumerate(lines) if ln.startswith('"ID"'))
    rows = list(csv.DictReader(io.StringIO("\n".join(lines[start:]))))
    agg = defaultdict(lambda: [0, 0.0])
    for r in rows:
        if r["Metric Name"] != "gpu__time_duration.sum":
            continue
        k = r["Kernel Name"].split("(")[0].replace("void ", "").replace("<unnamed>::", "")
        v = float(r["Metric Value"].replace(",", ""))
        if r["Metric Unit"] in ("us", "usecond"):
            v *= 1e3
        elif r["Metric Unit"] in ("ms", "msecond"):
            v *= 1e6
        agg[k][0] += 1
        agg[k][1] += v
    tot = sum(v[1] for v in agg.values())
    step_k = ("nr_pack_kernel", "nr_match_filtered_kernel<0>", "nr_match_exhaustive16_kernel")
    step_tot = sum(v[1] for k, v in agg.items() if k.startswith(step_k))
    out = [f"# Launch list of `python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-dp-gcups` ({tag})",
           "", "`ncu --metrics gpu__time_duration.sum --clock-control none` (cold-cache, serialised launches:",
           "the share column is what must agree with bench.py, not the absolute times).", "",
           "The whole process is listed: index build (nr_index_*, cub), the ALU-peak probe",
           "(nr_int_peak_kernel, measurement support) and the counted run (kernel<1>) are outside the",
           "timed step; `step %` is the share among the three kernels of a step.", "",
           "| kernel | launches | total ms | share % | step % |", "|---|---|---|---|---|"]
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        sp = f"{100 * v[1] / step_tot:.2f}" if k.startswith(step_k) and step_tot else ""
        out.append(f"| `{k[:80]}` | {v[0]} | {v[1] / 1e6:.3f} | {100 * v[1] / tot:.2f} | {sp} |")
    out.append(f"| total | {sum(v[0] for v in agg.values())} | {tot / 1e6:.3f} | 100 | |")
    open(os.path.join(ROOT, "profiles", f"{tag}_launches.md"), "w").write("\n".join(out) + "\n")
    return agg


def full(tag, n_cand):
    rep = os.path.join(ROOT, "gpurun_out", f"prof_filtered_{tag}.ncu-rep")
    if not os.path.exists(rep):
        return
    raw = subprocess.check_output(["ncu", "-i", rep, "--page", "raw", "--csv"], text=True)
    rows = list(csv.reader(io.StringIO(raw)))
    H, U = rows[0], rows[1]
    d = dict(zip(H, rows[2]))
    u = dict(zip(H, U))
    out = [f"# ncu --set full of nr_match_filtered_kernel ({tag}), {n_cand} candidates vs 737K", "",
           f"kernel: `{d.get('Kernel Name', '?')[:100]}`", "", "| metric | unit | value |", "|---|---|---|"]
    for h in H:
        if h in WANT or h in EXTRA:
            out.append(f"| {h} | {u[h]} | {d[h]} |")

    def f(name):
        v = d.get(name)
        return float(v.replace(",", "")) if v not in (None, "", "n/a") else None

    def scaled(name):
        v, unit = f(name), u.get(name, "")
        if v is None:
            return None
        return v * {"Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "byte": 1}.get(unit, 1)

    alu = f("smsp__inst_executed_pipe_alu.sum") or f("sm__inst_executed_pipe_alu.sum")
    fma = f("smsp__inst_executed_pipe_fma.sum") or f("sm__inst_executed_pipe_fma.sum")
    winst = f("smsp__inst_executed.sum")
    tinst = f("smsp__thread_inst_executed.sum")
    if tinst is None and winst is not None:
        tinst = winst * f("smsp__thread_inst_executed_per_inst_executed.ratio")
    dram = (scaled("dram__bytes_read.sum") or 0) + (scaled("dram__bytes_write.sum") or 0)
    js = {
        "737K-5p": {
            "source": f"profiles/{tag}_filtered_kernel_ncu.md (ncu --set full, {n_cand} candidates per launch)",
            "candidates_per_launch": n_cand,
            "alu_warp_inst_per_candidate": alu / n_cand if alu else None,
            "fma_warp_inst_per_candidate": fma / n_cand if fma else None,
            "warp_inst_per_candidate": winst / n_cand if winst else None,
            "thread_inst_per_candidate": tinst / n_cand if tinst else None,
            "dram_bytes_per_candidate": dram / n_cand,
            "alu_pipe_pct_of_peak_under_ncu": f("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active"),
        }
    }
    out += ["", "Derived per candidate:", "", "```json", json.dumps(js, indent=1), "```"]
    open(os.path.join(ROOT, "profiles", f"{tag}_filtered_kernel_ncu.md"), "w").write("\n".join(out) + "\n")
    p = os.path.join(ROOT, "profiles", "inst_per_candidate.json")
    old = json.load(open(p)) if os.path.exists(p) else {}
    old.update(js)
    json.dump(old, open(p, "w"), indent=1)
    print(json.dumps(js, indent=1))


if __name__ == "__main__":
    tag = sys.argv[1] if len(sys.argv) > 1 else "r1"
    n = int(sys.argv[2]) if len(sys.argv) > 2 else 1 << 22
    a = launches(tag)
    if a:
        print(open(os.path.join(ROOT, "profiles", f"{tag}_launches.md")).read())
    full(tag, n)
