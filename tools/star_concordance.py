#!/usr/bin/env python
"""STAR concordance harness: pins the CPU oracle against the reference's own aligner when a
`STAR` binary exists on the box (it does not in the build image -- then the harness says so).

TEST INFRASTRUCTURE (this module may import oracle/): called by `bench.py --impl reference`
and by tests/test_star_concordance.py.

What it runs, per golden fixture (tests/golden/{tcr3,mtdna1026,slideseq}.fa.gz, candidates cut
from the reference's sample_fastq by tests/golden/make_golden.py):

  1. the padded whitelist FASTA, written by nanoranger_b200.utils.write_bc_5p10X /
     write_bc_slideseq exactly as /root/reference/utils.py:604-622 / 584-601 write it;
  2. STAR --runMode genomeGenerate with the flags of scripts/barcode_ref.sh:11-18;
  3. STAR alignment with the flags of scripts/barcode_align.sh:14-35, then the rename of
     scripts/barcode_align.sh:41.
     When $NANORANGER_REF points at a reference checkout its two scripts are executed verbatim;
     otherwise the same argv is issued from the tables below (one entry per script line, cited)
     -- the scripts themselves are not copied into this repo, and nothing here reads
     /root/reference on its own;
  4. parse of `<out>.sam`: QNAME, FLAG, RNAME, AS:i;
  5. comparison with the oracle's answer for the same candidates (tests/golden/*.oracle.npz):
       identical_barcode_as  STAR reports the read and (barcode, AS, strand) equal the oracle's
                             unique best pair
       star_only             STAR reports a read the oracle calls ambiguous (tie) or puts elsewhere
       oracle_only           the oracle has a unique best pair, STAR reports nothing (expected for
                             reads STAR's seed search cannot anchor: STAR is a heuristic, the
                             oracle the exhaustive optimum of the same scoring)
     and wall-clock of index build and alignment (`cpu_baseline.kind = "reference"` timings).
"""
from __future__ import annotations

import gzip
import os
import shutil
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

ABSENT = "STAR absent — concordance not measured"

# scripts/barcode_ref.sh:11-18, one flag per script line
GENOME_GENERATE_FLAGS = [
    ("--runMode", "genomeGenerate"),            # :12
    ("--runThreadN", "4"),                      # :13
    ("--genomeDir", "{genome_dir}"),            # :14
    ("--genomeFastaFiles", "{ref_fasta}"),      # :15
    ("--genomeSAindexNbases", "6"),             # :16
    ("--genomeChrBinNbits", "7"),               # :17
    ("--limitGenomeGenerateRAM", "92000000000"),  # :18
]

# scripts/barcode_align.sh:14-35
ALIGN_FLAGS = [
    ("--runThreadN", "{threads}"),              # :15
    ("--readFilesIn", "{input}"),               # :16
    ("--genomeDir", "{genome_dir}"),            # :17
    ("--alignIntronMax", "1"),                  # :18
    ("--outFileNamePrefix", "{out}"),           # :19
    ("--outSAMmode", "NoQS"),                   # :20
    ("--outSAMattributes", "AS", "nM", "MD"),   # :21
    ("--outFilterMultimapNmax", "1"),           # :22
    ("--outFilterMultimapScoreRange", "0"),     # :23
    ("--outFilterScoreMinOverLread", "0"),      # :24
    ("--outFilterMatchNminOverLread", "0"),     # :25
    ("--scoreGenomicLengthLog2scale", "0"),     # :26
    ("--scoreDelBase", "-1"),                   # :27
    ("--scoreDelOpen", "0"),                    # :28
    ("--scoreInsOpen", "0"),                    # :29
    ("--scoreInsBase", "-1"),                   # :30
    ("--seedSearchStartLmax", "4"),             # :31
    ("--seedSearchStartLmaxOverLread", "0.9"),  # :32
    ("--alignEndsType", "EndToEnd"),            # :33
    ("--readNameSeparator", "space"),           # :34
    ("--readFilesCommand", "zcat"),             # :35
]

FIXTURES = {
    # name: (candidate FASTA, oracle answers, whitelist kind, pads, threshold)
    "tcr3": ("tcr3.fa.gz", "tcr3.oracle.npz", "737K", (30, 40), 14),
    "mtdna1026": ("mtdna1026.fa.gz", "mtdna1026.oracle.npz", "737K", (30, 40), 14),
    "slideseq": ("slideseq.fa.gz", "slideseq.oracle.npz", "slideseq", (15, 24), 30),
}


def _argv(star, table, **kw):
    out = [star]
    for row in table:
        out += [x.format(**kw) for x in row]
    return out


def reference_scripts_dir():
    base = os.environ.get("NANORANGER_REF")
    if base and os.path.isfile(os.path.join(base, "scripts", "barcode_align.sh")):
        return os.path.join(base, "scripts")
    return None


def star_genome_generate(star, ref_fasta, genome_dir, log=None):
    """scripts/barcode_ref.sh <ref_fasta> <genome_dir>."""
    os.makedirs(genome_dir, exist_ok=True)
    sd = reference_scripts_dir()
    t = time.perf_counter()
    if sd:
        cmd = ["bash", os.path.join(sd, "barcode_ref.sh"), ref_fasta, genome_dir]
    else:
        cmd = _argv(star, GENOME_GENERATE_FLAGS, genome_dir=genome_dir, ref_fasta=ref_fasta)
    subprocess.run(cmd, check=True, stdout=log, stderr=log, cwd=os.path.dirname(genome_dir) or ".")
    return time.perf_counter() - t, ("reference script" if sd else "restated argv")


def star_align(star, fasta_gz, genome_dir, out_prefix, threads, log=None):
    """scripts/barcode_align.sh <input.fa.gz> <genome_dir> <out_prefix> <threads> -> <out_prefix>.sam"""
    sd = reference_scripts_dir()
    t = time.perf_counter()
    if sd:
        subprocess.run(["bash", os.path.join(sd, "barcode_align.sh"), fasta_gz, genome_dir, out_prefix,
                        str(threads)], check=True, stdout=log, stderr=log)
    else:
        subprocess.run(_argv(star, ALIGN_FLAGS, threads=threads, input=fasta_gz,
                             genome_dir=genome_dir, out=out_prefix), check=True, stdout=log, stderr=log)
        os.replace(out_prefix + "Aligned.out.sam", out_prefix + ".sam")      # barcode_align.sh:41
    return time.perf_counter() - t


def parse_sam(path):
    """-> {qname: (flag, rname, AS)} (the fields utils.process_matching_* read, utils.py:697-702)."""
    out = {}
    with open(path) as f:
        for ln in f:
            if ln.startswith("@"):
                continue
            c = ln.rstrip("\n").split("\t")
            a = None
            for tag in c[11:]:
                if tag.startswith("AS:i:"):
                    a = int(tag[5:])
            out[c[0]] = (int(c[1]), c[2], a)
    return out


def read_fasta_gz(path):
    names, seqs = [], []
    with gzip.open(path, "rt") as f:
        for ln in f:
            ln = ln.strip()
            if ln.startswith(">"):
                names.append(ln[1:].split()[0])
            elif ln:
                seqs.append(ln)
    return names, seqs


def write_padded_whitelist(kind, workdir):
    """-> (fasta path, list of reference names in whitelist order)"""
    from nanoranger_b200 import utils as U, whitelists
    if kind == "737K":
        wl = whitelists.load_737k()
        src = os.path.join(workdir, "737K.txt")
        with open(src, "w") as f:
            for row in wl:
                f.write(row.tobytes().decode() + "-1\n")       # the list's own '-1' suffix (utils.py:613)
        U.write_bc_5p10X("conc", workdir, src)
    else:
        src = os.path.join(workdir, "slideseq.matched.barcodes.tsv")
        with gzip.open(os.path.join(ROOT, "tests", "golden", "slideseq_whitelist.txt.gz"), "rt") as g, \
                open(src, "w") as f:
            for ln in g:
                f.write(ln.strip() + "-1\n")
        U.write_bc_slideseq("conc", workdir, src)
    fa = os.path.join(workdir, "conc_bcreads.fasta")
    names = [ln[1:].strip() for ln in open(fa) if ln.startswith(">")]
    return fa, names


def compare(star_records, names, ref_names, oracle_npz):
    """star_records: parse_sam output; oracle_npz: dict with best_idx / best_score / n_best / strand."""
    o = oracle_npz
    n = len(names)
    ident = star_only = oracle_only = both_absent = differs = 0
    for i, q in enumerate(names):
        uniq = int(o["n_best"][i]) == 1
        rec = star_records.get(q)
        if rec is None:
            if uniq:
                oracle_only += 1
            else:
                both_absent += 1
            continue
        flag, rname, a = rec
        if uniq and ref_names[int(o["best_idx"][i])] == rname and a == int(o["best_score"][i]) and \
                (flag == 16) == (int(o["strand"][i]) == 1):
            ident += 1
        elif uniq:
            differs += 1
        else:
            star_only += 1
    return {"n": n, "identical_barcode_as": ident, "star_only": star_only, "oracle_only": oracle_only,
            "star_differs_from_unique_oracle_pair": differs, "absent_in_both": both_absent,
            "identical_fraction_of_star_records": ident / max(1, len(star_records))}


def run(fixtures=None, threads=None, keep=False):
    """-> dict for the bench JSON line.  Never raises because STAR is missing."""
    star = shutil.which("STAR")
    if star is None:
        return {"status": ABSENT, "star_binary": "absent"}
    threads = threads or os.cpu_count() or 1
    gold = os.path.join(ROOT, "tests", "golden")
    out = {"status": "measured", "star_binary": star, "threads": threads, "fixtures": {}}
    try:
        v = subprocess.run([star, "--version"], capture_output=True, text=True, timeout=30)
        out["star_version"] = (v.stdout or v.stderr).strip().splitlines()[0] if (v.stdout or v.stderr) else "?"
    except Exception as e:      # a stub or a broken binary: record, go on
        out["star_version"] = f"unknown ({type(e).__name__})"
    work = tempfile.mkdtemp(prefix="nr_star_")
    genomes = {}
    try:
        with open(os.path.join(work, "star.log"), "w") as log:
            for name in (fixtures or FIXTURES):
                fa_gz, npz, kind, pads, thr = FIXTURES[name]
                if kind not in genomes:
                    wd = os.path.join(work, kind)
                    os.makedirs(wd, exist_ok=True)
                    ref_fa, ref_names = write_padded_whitelist(kind, wd)
                    gdir = os.path.join(wd, "conc_ref")
                    t_idx, how = star_genome_generate(star, ref_fa, gdir, log)
                    genomes[kind] = (gdir, ref_names, t_idx, how)
                gdir, ref_names, t_idx, how = genomes[kind]
                cand = os.path.join(gold, fa_gz)
                prefix = os.path.join(work, f"{name}_matching")
                t_al = star_align(star, cand, gdir, prefix, threads, log)
                recs = parse_sam(prefix + ".sam")
                names, _ = read_fasta_gz(cand)
                o = dict(np.load(os.path.join(gold, npz)))
                r = compare(recs, names, ref_names, o)
                r.update(index_build_s=t_idx, align_s=t_al, invocation=how,
                         star_candidates_per_s=len(names) / t_al if t_al > 0 else None,
                         star_records=len(recs), threshold=thr)
                out["fixtures"][name] = r
    except (subprocess.CalledProcessError, OSError, KeyError, ValueError) as e:
        out["status"] = f"STAR present but the run failed: {type(e).__name__}: {e}"
    finally:
        if not keep:
            shutil.rmtree(work, ignore_errors=True)
    return out


if __name__ == "__main__":
    import json
    print(json.dumps(run(sys.argv[1:] or None), indent=1))
