"""Candidate extraction: the step right before the barcode matcher.

The reference walks the minimap2 transcriptome SAM with pysam and calls
``edlib.align(const, window, "HW", "locations", k[, ad_seq])`` on a flank window of every
alignment (utils.decon_*, utils.py:95-189, 192-310, 360-483).  Here the SAM is read as text, the
windows of ALL records are searched in one batch by the bit-parallel kernel
(``nr_hw_search_*``, nanoranger_b200/csrc/nr_hwsearch.cu), and the same files are written:
``{sample}_deconcat.fastq.gz`` / ``{sample}_VDJ.fastq``, ``{sample}_BCUMI.fasta(.gz)``,
``{sample}_eds.csv``.  Same function names and arguments as the reference; no pysam, no edlib.
"""
from __future__ import annotations

import gzip
import re

import numpy as np
import pandas as pd

from . import _lib
from .matcher import pack_ascii
from .samio import revcomp_bytes
from .whitelists import LINKER_SLIDESEQ

ad_seq = [("N", "A"), ("N", "T"), ("N", "G"), ("N", "C")]      # utils.py:15
_CIG = re.compile(r"(\d+)([MIDNSHP=X])")


def rev(seq: str) -> str:
    """utils.py:18-19 (Bio.Seq reverse_complement for ACGTN)."""
    return revcomp_bytes(seq.encode("ascii")).decode("ascii")


# ---- the edlib call ---------------------------------------------------------------------------------

def hw_search(windows, pattern: str, k: int, wildcard: bool = True, device: int = 0):
    """Batch form of ``edlib.align(pattern, w, "HW", "locations", k[, ad_seq])``.
    windows: list of str, or (u8 buffer, u64 offsets).  -> dict of numpy arrays:
    ed [n] int8 (-1: more than k), first [n,2], last [n,2] int32 (start, end inclusive; the
    reference's ``ed["locations"][0]`` / ``[-1]``), nloc [n]."""
    if isinstance(windows, tuple):
        buf, off = windows
        buf = np.ascontiguousarray(buf, np.uint8)
        off = np.ascontiguousarray(off, np.uint64)
    else:
        buf, off = pack_ascii(windows)
    n = len(off) - 1
    ed = np.empty(n, np.int8)
    first = np.empty((n, 2), np.int32)
    last = np.empty((n, 2), np.int32)
    nloc = np.empty(n, np.int32)
    if n:
        if len(buf) == 0:
            buf = np.zeros(1, np.uint8)
        _lib.check(_lib.lib().nr_hw_search_host(
            buf.ctypes.data, off.ctypes.data, n, pattern.encode("ascii"), len(pattern), k,
            1 if wildcard else 0, ed.ctypes.data, first.ctypes.data, last.ctypes.data,
            nloc.ctypes.data, device), "nr_hw_search_host")
    return {"ed": ed, "first": first, "last": last, "nloc": nloc}


def hw_search_device(d_text, d_offsets, pattern: str, k: int, wildcard: bool = True):
    """Device-resident form (torch uint8 text, int64 offsets); stream-ordered, no sync."""
    import torch
    n = d_offsets.numel() - 1
    dev = d_text.device
    ed = torch.empty(n, dtype=torch.int8, device=dev)
    first = torch.empty((n, 2), dtype=torch.int32, device=dev)
    last = torch.empty((n, 2), dtype=torch.int32, device=dev)
    nloc = torch.empty(n, dtype=torch.int32, device=dev)
    st = torch.cuda.current_stream(dev).cuda_stream
    with torch.cuda.device(dev):
        _lib.check(_lib.lib().nr_hw_search_device(
            d_text.data_ptr(), d_offsets.data_ptr(), n, pattern.encode("ascii"), len(pattern), k,
            1 if wildcard else 0, ed.data_ptr(), first.data_ptr(), last.data_ptr(),
            nloc.data_ptr(), st), "nr_hw_search_device")
    return {"ed": ed, "first": first, "last": last, "nloc": nloc}


# ---- minimap2 SAM records with the pysam attributes the extractors use -----------------------------

class Aln:
    """One alignment line.  Attribute names follow pysam.AlignedSegment as the reference uses
    them (utils.py:112-127): qname, flag, reference_name, seq, qual, rlen (= len(SEQ)),
    qlen (= aligned query length), query_alignment_start/end, reference_start/end."""
    __slots__ = ("qname", "flag", "reference_name", "reference_start", "reference_end", "seq",
                 "qual", "rlen", "qlen", "query_alignment_start", "query_alignment_end", "tags")

    def __init__(self, t):
        self.qname, self.flag, self.reference_name = t[0], int(t[1]), t[2]
        self.reference_start = int(t[3]) - 1
        self.seq = "" if t[9] == "*" else t[9]
        self.qual = t[10]
        ops = [(int(n), op) for n, op in _CIG.findall(t[5])]
        lead = trail = 0
        k = 0
        while k < len(ops) and ops[k][1] in "SH":
            if ops[k][1] == "S":
                lead += ops[k][0]
            k += 1
        k = len(ops) - 1
        while k >= 0 and ops[k][1] in "SH":
            if ops[k][1] == "S":
                trail += ops[k][0]
            k -= 1
        ref_len = sum(n for n, op in ops if op in "MDN=X")
        self.reference_end = self.reference_start + ref_len
        self.rlen = len(self.seq)
        self.query_alignment_start = lead
        self.query_alignment_end = self.rlen - trail
        self.qlen = self.query_alignment_end - self.query_alignment_start
        self.tags = t[11:]

    def get_tag(self, name):
        pre = name + ":"
        for x in self.tags:
            if x.startswith(pre):
                v = x.split(":", 2)
                return int(v[2]) if v[1] == "i" else (float(v[2]) if v[1] == "f" else v[2])
        raise KeyError(f"tag '{name}' not present")


def read_alignments(path: str):
    """Mapped alignment lines of a SAM text file in file order (what `samfile.fetch()` yields
    for the reference: minimap2 runs with --sam-hit-only --secondary=no,
    scripts/align_trns.sh)."""
    op = gzip.open if path.endswith(".gz") else open
    out = []
    with op(path, "rt") as f:
        for ln in f:
            if ln.startswith("@"):
                continue
            t = ln.rstrip("\n").split("\t")
            if len(t) >= 11 and not (int(t[1]) & 4):
                out.append(Aln(t))
    return out


def _gz_write(path, text_iter):
    """the reference writes plain text and then runs `pigz -f` (utils.py:187-188)"""
    with gzip.open(path + ".gz", "wt", compresslevel=6) as f:
        for s in text_iter:
            f.write(s)


def _mod_coords(r):
    if r.flag == 16 or r.flag == 2064:
        return r.rlen - r.query_alignment_end, r.rlen - r.query_alignment_start
    return r.query_alignment_start, r.query_alignment_end


# ---- 5' 10x ------------------------------------------------------------------------------------------

def _decon_5p(sample, outdir, n_umi, lclip, rclip, tcr, device):
    const = "CGCTCTTCCGATCT" + n_umi * "N" + "TTTCTTATATG"
    recs = read_alignments(f"{outdir}/{sample}_trns.sam")
    wins = []
    for r in recs:
        r.get_tag("AS")
        qs = r.query_alignment_start
        wins.append(r.seq[qs - lclip: qs + rclip] if qs > lclip else r.seq[: qs + rclip])
    res = hw_search(wins, const, 6, True, device)
    fq, fa, eds = [], [], []
    lclipV, rclipV, r_hang = 60, 80, 0
    for i, r in enumerate(recs):
        qs, qe = r.query_alignment_start, r.query_alignment_end
        qsm, qem = _mod_coords(r)
        name = f"{r.qname}_{qsm}_{qem}_{r.flag}_{r.reference_name}"
        if tcr:                                                     # utils.py:247-269
            sub_s = qe - lclipV if r.qlen > lclipV else qs
            sub_e = qe + rclipV if r.rlen - qe > rclipV else r.rlen
            fq.append(f"@{name}\n{r.seq[sub_s:sub_e]}\n+\n{r.qual[sub_s:sub_e]}\n")
        d = int(res["ed"][i])
        if -1 < d < 7:
            start, end = (int(x) for x in res["last"][i])
            bcumi = wins[i][start:end]
            start = lclip - start if qs > lclip else qs - start
            if tcr:
                eds.append([start, r.reference_name, len(bcumi), d])
            else:                                                   # utils.py:146-168
                eds.append([start, len(bcumi), d])
                fq.append(f"@{name}\n{r.seq[qs:qe + r_hang]}\n+\n{r.qual[qs:qe + r_hang]}\n")
            fa.append(f">{name}\n{bcumi}\n")
    _gz_write(f"{outdir}/{sample}_deconcat.fastq", fq)
    _gz_write(f"{outdir}/{sample}_BCUMI.fasta", fa)
    if tcr:
        pd.DataFrame(np.array(eds)).to_csv(f"{outdir}/{sample}_eds.csv")
    return len(fa)


def decon_5p10XGEX(sample, outdir, device: int = 0):
    """utils.py:95-189: window seq[qstrt-80 : qstrt+20], motif with 26 N, k = 6, LAST location."""
    return _decon_5p(sample, outdir, 26, 80, 20, False, device)


def decon_5p10XTCR(sample, outdir, device: int = 0):
    """utils.py:192-310: window seq[qstrt-200 : qstrt+20], motif with 28 N (GEM-X), k = 6, LAST
    location; every alignment goes to the FASTQ, hits to the FASTA and `{sample}_eds.csv`."""
    return _decon_5p(sample, outdir, 28, 200, 20, True, device)


# ---- 3' slide-seq -------------------------------------------------------------------------------------

def decon_3pXCR_slideseq(sample, outdir, device: int = 0):
    """utils.py:360-483: linker search (k = 2, no wildcard, FIRST location) in 40-nt windows every
    20 nt of the 200 nt after the alignment; first window with a hit wins."""
    const = rev(LINKER_SLIDESEQ)
    rclip, lclip, rclipA, lclipA, r_search = 80, 200, 22, 16, 200
    recs = read_alignments(f"{outdir}/{sample}_trns.sam")
    wins, owner, meta = [], [], []
    fq = []
    for ri, r in enumerate(recs):
        r.get_tag("AS")
        qs, qe = r.query_alignment_start, r.query_alignment_end
        accept = (r.reference_end - r.reference_start) > 400
        dd = r.seq[qe: qe + r_search] if r.rlen - qe > r_search else r.seq[qe:]
        sub_e = qs + rclip
        sub_s = qs - lclip if qs > lclip else 0
        sub_seq = r.seq[sub_s:sub_e]
        name = f"{r.qname}_{sample}_{sub_s}_{sub_e}_{r.flag}_{r.reference_name}"
        meta.append((name, dd))
        if len(sub_seq) > 100 and accept:
            fq.append(f"@{name}\n{sub_seq}\n+\n{r.qual[sub_s:sub_e]}\n")
            for i in range(int(len(dd) / 20)):
                wins.append(dd[20 * i: 20 * i + 40])
                owner.append((ri, i))
    res = hw_search(wins, const, 2, False, device)
    newnames, c_hangs, polyAs, c_eds = [], [], [], []
    done = set()
    for w, (ri, i) in enumerate(owner):
        d = int(res["ed"][w])
        if ri in done or not (-1 < d < 4):
            continue
        done.add(ri)
        name, dd = meta[ri]
        start = int(res["first"][w][0]) + 20 * i
        end = int(res["first"][w][1]) + 20 * i
        upstart = 0 if start - rclipA < 0 else start - rclipA
        upend = end + lclipA
        c_hangs.append(rev(dd[upstart:upend]))
        polyAs.append(dd[: upstart + 5])
        c_eds.append(d)
        newnames.append(">" + name)
    _gz_write(f"{outdir}/{sample}_VDJ.fastq", fq)                      # utils.py:483-486 (pigz -f)
    with gzip.open(f"{outdir}/{sample}_eds_names.csv.gz", "wt") as f:
        pd.DataFrame([newnames, c_eds]).T.to_csv(f, index=None)
    fa, pa_out = [], []
    for nm, hang, pa in zip(newnames, c_hangs, polyAs):
        accept = len(hang) > 45 and len(pa) < 70
        if len(hang) > 45 and len(pa) > 70 and pa.count("A") / len(pa) > 0.5:
            accept = True
        if accept:
            fa.append(f"{nm}\n{hang}\n")
            pa_out.append(f"{nm}\n{pa}\n")
    _gz_write(f"{outdir}/{sample}_BCUMI.fasta", fa)
    _gz_write(f"{outdir}/{sample}_polyA.fasta", pa_out)
    return len(fa)


# ---- 3' 10x ------------------------------------------------------------------------------------------

def decon_3p10XTCR(sample, outdir, device: int = 0):
    """utils.py:313-373: 6A + 28N + TruSeq motif (k = 5, N wildcard, FIRST location) in the 150 nt
    after the alignment; `{sample}_VDJ.fastq.gz`, `{sample}_BCUMI.fasta.gz`, `{sample}_eds.csv`."""
    from .utils import sort_cnt
    const = 6 * "A" + 28 * "N" + "AGATCGGAAGAGCGTCGTGT"
    lclip, r_search, rclip = 350, 150, 100
    recs = read_alignments(f"{outdir}/{sample}_trns.sam")
    wins = []
    for r in recs:
        qe = r.query_alignment_end
        wins.append(r.seq[qe: qe + r_search] if r.rlen - qe > r_search else r.seq[qe:])
    res = hw_search(wins, const, 5, True, device)
    fq, fa, eds = [], [], []
    for i, r in enumerate(recs):
        qs = r.query_alignment_start
        sub_e = qs + rclip
        sub_s = qs - lclip if qs > lclip else 0
        sub_seq = r.seq[sub_s:sub_e]
        dist = int(res["ed"][i])
        eds.append(dist)
        name = (f"{r.qname[-10:]}_q{r.qlen}_d{dist}_s{sub_s}_e{sub_e}_f{r.flag}_"
                f"{r.reference_name.split('-')[0]}")
        if -1 < dist < 6 and len(sub_seq) > 100 and r.qlen > 100:
            fq.append(f"@{name}\n{sub_seq}\n+\n{r.qual[sub_s:sub_e]}\n")
            a, b = (int(x) for x in res["first"][i])
            fa.append(f">{name}\n{rev(wins[i][a:b])[14:]}\n")
    _gz_write(f"{outdir}/{sample}_VDJ.fastq", fq)
    _gz_write(f"{outdir}/{sample}_BCUMI.fasta", fa)
    sort_cnt(eds).to_csv(f"{outdir}/{sample}_eds.csv")
    return len(fa)


def _stepped_first_hit(tails, const, k, accept_below, device, step=200, overlap=70):
    """windows end_qu[step*i : step*(i+1)+overlap] for i = 0..len//step of every tail (None:
    record not searched); -> per tail (i, start, end, ed) of the first window with
    editDistance < accept_below, or None (utils.py:1045-1083, 1359-1383)."""
    wins, owner = [], []
    for ti, t in enumerate(tails):
        if t is None:
            continue
        for i in range(int(len(t) / step) + 1):
            wins.append(t[step * i: step * (i + 1) + overlap])
            owner.append((ti, i))
    res = hw_search(wins, const, k, False, device)
    hit = [None] * len(tails)
    for w, (ti, i) in enumerate(owner):
        d = int(res["ed"][w])
        if hit[ti] is None and -1 < d < accept_below:
            hit[ti] = (i, int(res["first"][w][0]), int(res["first"][w][1]), d)
    return hit


def decon_3p10XTCR_nuc(sample, outdir, device: int = 0):
    """utils.py:982-1113: TruSeq adapter (k = 2, FIRST location) in 270-nt windows every 200 nt of
    the 2000 nt after the alignment; candidate = rev(end_qu[start-35 : end-12]), kept if > 30 nt."""
    const = "AGATCGGAAGAGCGTCGTGT"
    r_search, rclip = 2000, 100
    recs = read_alignments(f"{outdir}/{sample}_trns.sam")
    tails, fq, names = [], [], []
    for r in recs:
        qs, qe = r.query_alignment_start, r.query_alignment_end
        end_qu = r.seq[qe: qe + r_search] if r.rlen - qe > r_search else r.seq[qe:]
        sub_e = qe + rclip if r.rlen - qe > rclip else r.rlen
        sub_seq = r.seq[qs:sub_e]
        name = f"{r.qname}_{sample}_{qs}_{sub_e}_{r.flag}_{r.reference_name.split('-')[0]}"
        names.append(name)
        if len(sub_seq) > 100:
            fq.append(f"@{name}\n{sub_seq}\n+\n{r.qual[qs:sub_e]}\n")
            tails.append(end_qu)
        else:
            tails.append(None)
    hits = _stepped_first_hit(tails, const, 2, 3, device)
    fa, short_bc = [], 0
    for ti, h in enumerate(hits):
        if h is None:
            continue
        i, a, b, _ = h
        start, end = a + 200 * i, b + 200 * i
        bcumi = rev(tails[ti][start - 35: end - 12])
        if len(bcumi) > 30:
            fa.append(f">{names[ti]}\n{bcumi}\n")
        else:
            short_bc += 1
    print(short_bc)
    _gz_write(f"{outdir}/{sample}_VDJ.fastq", fq)
    _gz_write(f"{outdir}/{sample}_BCUMI.fasta", fa)
    return len(fa)


def decon_3p10XGEX(sample, outdir, device: int = 0):
    """utils.py:1283-1410: TruSeq adapter (k = 3, FIRST location) in stepped windows of
    seq[qend-70 : qend+700]; candidate = rev(end_qu[start-32 : start+3]) (35 nt), raw barcode
    counts to `{sample}_bc_count.json`."""
    import json
    import os
    bc_count_json = f"{outdir}/{sample}_bc_count.json"
    if os.path.isfile(bc_count_json):
        print(bc_count_json, " exists, skip")
        return
    const = "AGATCGGAAGAGCGTCGTGT"
    r_search, rclip, lclip = 700, 1, 1
    recs = read_alignments(f"{outdir}/{sample}_trns.sam")
    tails, fq, names = [], [], []
    for r in recs:
        qs, qe = r.query_alignment_start, r.query_alignment_end
        end_qu = r.seq[qe - 70: qe + r_search] if r.rlen - qe > r_search else r.seq[qe - 70:]
        sub_e = qe + rclip if r.rlen - qe > rclip else r.rlen
        sub_s = 0 if qs < lclip else qs - lclip
        sub_seq = r.seq[sub_s:sub_e]
        qsm, qem = _mod_coords(r)
        name = f"{r.qname}_{qsm}_{qem}_{r.flag}_{r.reference_name}"
        names.append(name)
        if len(sub_seq) > 50:
            fq.append(f"@{name}\n{sub_seq}\n+\n{r.qual[sub_s:sub_e]}\n")
            tails.append(end_qu)
        else:
            tails.append(None)
    hits = _stepped_first_hit(tails, const, 3, 4, device)
    fa, short_bc, counts = [], 0, {}
    for ti, h in enumerate(hits):
        if h is None:
            continue
        i, a, _, _ = h
        start = a + 200 * i
        bcumi = rev(tails[ti][start - 16 - 12 - 4: start + 3])
        raw = bcumi[3:3 + 16]
        counts[raw] = counts.get(raw, 0) + 1
        if len(bcumi) > 30:
            fa.append(f">{names[ti]}\n{bcumi}\n")
        else:
            short_bc += 1
    print("number of short BCUMIs = ", short_bc)
    _gz_write(f"{outdir}/{sample}_deconcat.fastq", fq)
    _gz_write(f"{outdir}/{sample}_BCUMI.fasta", fa)
    with open(bc_count_json, "w") as f:
        json.dump(counts, f)
    return len(fa)
