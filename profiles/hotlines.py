#!/usr/bin/env python
"""Per-source-line instruction counts from an .ncu-rep captured with --import-source on
(kernels compiled with -lineinfo):  python profiles/hotlines.py <rep> [top]"""
import csv
import io
import subprocess
import sys


def main():
    rep = sys.argv[1]
    top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
    raw = subprocess.check_output(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source",
                                   "cuda,sass"], text=True, stderr=subprocess.DEVNULL)
    cur, H, out = None, None, []
    for r in csv.reader(io.StringIO(raw)):
        if not r:
            continue
        if r[0] == "File Path":
            cur, H = r[1].split("/")[-1], None
            continue
        if r[0] == "Function Name":
            continue
        if r[0] == "Line No":
            H = r
            continue
        if H is None or r[2] != "-":          # rows with Address "-" are the per-line totals
            continue
        ie = float(r[H.index("Instructions Executed")] or 0)
        te = float(r[H.index("Thread Instructions Executed")] or 0)
        sm = float(r[H.index("# Samples")] or 0)
        if ie > 0:
            out.append((ie, te, sm, cur, r[0], r[1].strip()[:100]))
    tot = sum(o[0] for o in out)
    tots = sum(o[2] for o in out)
    print(f"total warp instructions {tot:.0f}, samples {tots:.0f}")
    out.sort(reverse=True)
    print("| inst % | thr/inst | samples % | where | source |\n|---|---|---|---|---|")
    for o in out[:top]:
        print(f"| {o[0] / tot * 100:.1f} | {o[1] / o[0]:.1f} | {o[2] / max(tots, 1) * 100:.1f} | {o[3]}:{o[4]} | `{o[5]}` |")


if __name__ == "__main__":
    main()
