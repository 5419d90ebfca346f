#!/usr/bin/env python
"""Generates the fixtures in this directory.  Run in the build container, where /root/reference
exists (the GPU box has no /root/reference; tests read only the committed outputs).

Inputs  : /root/reference/sample_fastq/*.fastq.gz, /root/reference/data/*.gz (read-only data)
Outputs : <name>.fa.gz            barcode+UMI candidates cut by the lite extractor below
          <name>.oracle.npz       oracle results for them (oracle/nr_oracle.c, exhaustive)
          slideseq_whitelist.txt.gz   14-nt slide-seq barcodes (reference data file, '-1' stripped)

Lite extractor: minimap2/pysam/edlib are not installed, so instead of
utils.decon_* the exact TSO 'TTTCTTATATG' is located on both strands and the 40 nt upstream plus
the first 10 nt of the TSO are kept (shape of decon_5p10X* output, utils.py:129-139); for
slide-seq the exact linker (utils.py:14) is located and 16 nt before / 22 nt after are kept
(utils.py:443-448).  These are inputs only; expected outputs come from the oracle.
"""
import gzip
import os
import re
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from oracle import oracle as O  # noqa: E402
from nanoranger_b200 import whitelists  # noqa: E402

REF = "/root/reference"
TSO = "TTTCTTATATG"
LINKER = "TCTTCAGCGTTCCCGAGA"


def read_fastq(path):
    with gzip.open(path, "rt") as f:
        while True:
            h = f.readline()
            if not h:
                return
            s = f.readline().strip()
            f.readline(); f.readline()
            yield h[1:].split()[0], s


def lite_5p(path):
    out = []
    for name, seq in read_fastq(path):
        for flag, s in ((0, seq), (16, O.revcomp(seq))):
            for mt in re.finditer(TSO, s):
                a = mt.start()
                if a < 26:
                    continue
                cand = s[max(0, a - 40):a + 10]
                out.append((f"{name}_{max(0, a - 40)}_{a + 10}_{flag}_lite", cand))
    return out


def lite_slideseq(path):
    out = []
    for name, seq in read_fastq(path):
        for flag, s in ((0, seq), (16, O.revcomp(seq))):
            for mt in re.finditer(LINKER, s):
                a, b = mt.start(), mt.end()
                if a < 16 or b + 22 > len(s):
                    continue
                out.append((f"{name}_slideseq_{a - 16}_{b + 22}_{flag}_lite", s[a - 16:b + 22]))
    return out


def write_fa(path, recs):
    with gzip.open(path, "wt") as f:
        for n, s in recs:
            f.write(f">{n}\n{s}\n")


def run(name, recs, wl_codes, pad_l, pad_r):
    seqs = [s for _, s in recs]
    cc, cl = O.encode_many(seqs, 64)
    r = O.match(wl_codes, pad_l, pad_r, cc, cl)
    np.savez_compressed(os.path.join(HERE, f"{name}.oracle.npz"), pad_l=pad_l, pad_r=pad_r, **r)
    write_fa(os.path.join(HERE, f"{name}.fa.gz"), recs)
    hi = r["best_score"] >= (14 if wl_codes.shape[1] == 16 else 30)
    print(name, len(recs), "candidates;", int(hi.sum()), "at/above threshold;",
          int((hi & (r["n_best"] == 1) & (r["strand"] == 0)).sum()), "assigned")


def main():
    wl737 = O._CODE[whitelists.load_737k()]
    for name, fq in (("tcr3", "TCR3.fastq.gz"), ("mtdna1026", "1026_mtDNA_ASXL1_NRAS_SF3B1.fastq.gz")):
        recs = lite_5p(os.path.join(REF, "sample_fastq", fq))
        run(name, recs, wl737, 30, 40)
    bcs = sorted({ln.strip().split("-")[0] for ln in
                  gzip.open(os.path.join(REF, "data", "slideseq.matched.barcodes.tsv.gz"), "rt")})
    with gzip.open(os.path.join(HERE, "slideseq_whitelist.txt.gz"), "wt") as f:
        f.write("\n".join(bcs) + "\n")
    cores = [b[:8] + LINKER + b[8:] for b in bcs]
    wlc, _ = O.encode_many(cores, 32)
    run("slideseq", lite_slideseq(os.path.join(REF, "sample_fastq", "slideseq_XCR.fastq.gz")), wlc, 15, 24)


if __name__ == "__main__":
    main()
