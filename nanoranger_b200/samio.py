"""Minimal SAM text I/O for `{sample}_matching.sam` (the file scripts/barcode_align.sh:41 of the
reference produces and utils.process_matching_* consume through pysam).

Fields the reference consumes: QNAME, FLAG, RNAME, POS+CIGAR (only to find the query index
aligned to reference position padL+L, utils.py:705-708), SEQ, AS:i.  The matcher does not trace
alignments back, so POS/CIGAR are *anchored*: an ungapped M run placed so that reference column
padL+L pairs with the UMI start the kernel reports (leading/trailing I where the run would leave
the padded reference); when no optimal alignment reaches that column the run ends at column
padL+L-1, so the column is absent from aligned_pairs exactly as pysam would report for STAR's
record (the reference then takes its `except` branch, utils.py:709-710).
"""
from __future__ import annotations

import re

import numpy as np

_CIG = re.compile(r"(\d+)([MIDNSHP=X])")
_COMP = bytes.maketrans(b"ACGTNacgtn", b"TGCANtgcan")


def anchored_alignment(m: int, umi_q: int, pad_l: int, core_len: int, pad_r: int):
    """-> (pos 1-based, cigar) for a read of length m whose base umi_q pairs with reference
    column pad_l+core_len (0-based); umi_q < 0: read ends at column pad_l+core_len-1."""
    ref_len = pad_l + core_len + pad_r
    if m == 0:
        return 1, "*"
    if umi_q < 0:
        umi_col, q_anchor = pad_l + core_len, m        # one past the last read base
    else:
        umi_col, q_anchor = pad_l + core_len, umi_q
    start = umi_col - q_anchor                            # reference column of read base 0
    lead = max(0, -start)
    end = start + m
    trail = max(0, end - ref_len)
    mid = m - lead - trail
    if mid <= 0:
        return 1, f"{m}I"
    cig = (f"{lead}I" if lead else "") + f"{mid}M" + (f"{trail}I" if trail else "")
    return max(start, 0) + 1, cig


def revcomp_bytes(b: bytes) -> bytes:
    return b.translate(_COMP)[::-1]


def write_sam(path: str, ref_names, ref_len: int, records, header: str = "used") -> int:
    """records: iterable of (qname, flag, rname_idx, pos, cigar, seq, AS).
    header: 'used' writes @SQ only for referenced entries, 'full' for every entry (STAR's layout:
    one @SQ per barcode, 737 280 lines for the 10x list)."""
    records = list(records)
    with open(path, "w") as f:
        f.write("@HD\tVN:1.4\n")
        if header == "full":
            idxs = range(len(ref_names))
        else:
            idxs = sorted({r[2] for r in records})
        for i in idxs:
            f.write(f"@SQ\tSN:{ref_names[i]}\tLN:{ref_len}\n")
        f.write("@PG\tID:nanoranger_b200\tPN:nanoranger_b200\n")
        f.write("@CO\tuser command line: nanoranger_b200.utils.barcode_align\n")
        # STAR is asked for "AS nM MD" (scripts/barcode_align.sh:20); nM/MD need a traceback the
        # matcher does not do and no reference code reads them, so only AS is written
        for q, flag, ri, pos, cig, seq, a_s in records:
            f.write(f"{q}\t{flag}\t{ref_names[ri]}\t{pos}\t255\t{cig}\t*\t0\t0\t{seq}\t*\t"
                    f"NH:i:1\tHI:i:1\tAS:i:{a_s}\n")
    return len(records)


def query_index_at(pos: int, cigar: str, ref_col: int):
    """query index paired with 0-based reference column ref_col (pysam aligned_pairs semantics:
    only M/=/X columns pair), or None."""
    r, q = pos - 1, 0
    for n, op in _CIG.findall(cigar):
        n = int(n)
        if op in "M=X":
            if r <= ref_col < r + n:
                return q + (ref_col - r)
            r += n
            q += n
        elif op in "IS":
            q += n
        elif op in "DN":
            r += n
    return None


def read_sam(path: str):
    """-> list of dict(qname, flag, rname, pos, cigar, seq, AS) for alignment lines."""
    out = []
    with open(path) as f:
        for ln in f:
            if ln.startswith("@"):
                continue
            t = ln.rstrip("\n").split("\t")
            if len(t) < 11:
                continue
            a_s = None
            for tag in t[11:]:
                if tag.startswith("AS:i:"):
                    a_s = int(tag[5:])
            out.append({"qname": t[0], "flag": int(t[1]), "rname": t[2], "pos": int(t[3]),
                        "cigar": t[5], "seq": t[9], "AS": a_s})
    return out
