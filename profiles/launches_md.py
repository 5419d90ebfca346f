#!/usr/bin/env python
"""ncu launch list (--metrics gpu__time_duration.sum --csv --log-file X) -> a tracked table:
  python profiles/launches_md.py gpurun_out/r2d/launches.csv profiles/r2d_launches.md "<command>"
Kernels of the timed step get their share among themselves in the last column."""
import csv
import io
import sys
from collections import defaultdict

STEP = ("nr_pack_kernel", "nr_match_filtered_kernel<0, 0>", "nr_match_filtered_kernel<0, 1>",
        "nr_match_deep_kernel", "nr_deep_finalize_kernel", "nr_match_bitsliced_kernel", "nr_match_exhaustive")


def main():
    src, dst, cmd = sys.argv[1], sys.argv[2], sys.argv[3]
    lines = open(src).read().splitlines()
    start = next(i for i, ln in enumerate(lines) if ln.startswith('"ID"'))
    agg = defaultdict(lambda: [0, 0.0])
    for r in csv.DictReader(io.StringIO("\n".join(lines[start:]))):
        if r["Metric Name"] != "gpu__time_duration.sum":
            continue
        k = r["Kernel Name"].split("(")[0].replace("void ", "").replace("<unnamed>::", "")
        v = float(r["Metric Value"].replace(",", ""))
        v *= {"us": 1e3, "usecond": 1e3, "ms": 1e6, "msecond": 1e6, "s": 1e9, "second": 1e9}.get(r["Metric Unit"], 1.0)
        if k.startswith("nr_match_bitsliced_kernel") and v > 5e6:
            # not part of a step: bench.py's dp_gcups sample (2 048 candidates through the brute force)
            k = "[dp_gcups sample] " + k
        agg[k][0] += 1
        agg[k][1] += v
    tot = sum(v[1] for v in agg.values())
    step_tot = sum(v[1] for k, v in agg.items() if k.startswith(STEP))
    out = [f"# Launch list of `{cmd}`", "",
           "`ncu --metrics gpu__time_duration.sum --clock-control none` (cold-cache, serialised launches: the",
           "share columns are what must agree with bench.py, not the absolute times).  The whole process is",
           "listed -- index build (nr_index_*, cub / thrust), the ALU-peak probe (nr_int_peak_kernel), the",
           "counted run (kernel<1, ...>), the e2e pass through nr_match_host (same kernels, 1 Mi-candidate",
           "chunks) and the brute-force sample (dp_gcups); `step %` is the share among the kernels a step",
           "launches.", "", "| kernel | launches | total ms | share % | step % |", "|---|---|---|---|---|"]
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        sp = f"{100 * v[1] / step_tot:.2f}" if k.startswith(STEP) and step_tot else ""
        out.append(f"| `{k[:90]}` | {v[0]} | {v[1] / 1e6:.3f} | {100 * v[1] / tot:.2f} | {sp} |")
    out.append(f"| total | {sum(v[0] for v in agg.values())} | {tot / 1e6:.3f} | 100 | |")
    open(dst, "w").write("\n".join(out) + "\n")


if __name__ == "__main__":
    main()
