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


def read_sam_table(path: str):
    """Column-wise reader for large SAM files (pandas' C parser): -> dict of numpy arrays
    qname, flag, rname, pos, cigar, seq (object / int64) and AS (int64; -1 where the tag is
    missing).  Same content as read_sam()."""
    import pandas as pd
    n_hdr = 0
    first = None
    with open(path) as f:
        for ln in f:
            if ln.startswith("@"):
                n_hdr += 1
                continue
            first = ln.rstrip("\n").split("\t")
            break
    empty = {k: np.zeros(0, object) for k in ("qname", "rname", "cigar", "seq")}
    empty.update({k: np.zeros(0, np.int64) for k in ("flag", "pos", "AS")})
    if first is None:
        return empty
    as_col = next((j for j in range(11, len(first)) if first[j].startswith("AS:i:")), None)
    cols = [0, 1, 2, 3, 5, 9] + ([as_col] if as_col is not None else [])
    try:
        df = pd.read_csv(path, sep="\t", header=None, skiprows=n_hdr, usecols=cols, quoting=3,
                         dtype={0: str, 1: np.int64, 2: str, 3: np.int64, 5: str, 9: str},
                         na_filter=False, engine="c")
        AS = (df[as_col].str.slice(5).astype(np.int64).to_numpy() if as_col is not None
              else np.full(len(df), -1, np.int64))
        if as_col is not None and not df[as_col].str.startswith("AS:i:").all():
            raise ValueError("AS tag not in a fixed column")
    except Exception:
        recs = read_sam(path)                       # irregular file: record-wise reader
        return {"qname": np.array([r["qname"] for r in recs], object),
                "flag": np.array([r["flag"] for r in recs], np.int64),
                "rname": np.array([r["rname"] for r in recs], object),
                "pos": np.array([r["pos"] for r in recs], np.int64),
                "cigar": np.array([r["cigar"] for r in recs], object),
                "seq": np.array([r["seq"] for r in recs], object),
                "AS": np.array([-1 if r["AS"] is None else r["AS"] for r in recs], np.int64)}
    return {"qname": df[0].to_numpy(object), "flag": df[1].to_numpy(), "rname": df[2].to_numpy(object),
            "pos": df[3].to_numpy(), "cigar": df[5].to_numpy(object), "seq": df[9].to_numpy(object),
            "AS": AS}


_SIMPLE = re.compile(r"^(?:(\d+)I)?(\d+)M(?:(\d+)I)?$")


def query_index_at_many(pos, cigar, ref_col: int):
    """query_index_at() for arrays; -1 where the column is not paired.  CIGARs of the anchored
    form [aI]bM[cI] (everything this package writes) are handled without a per-record parse."""
    out = np.full(len(pos), -1, np.int64)
    for i, (p, c) in enumerate(zip(pos, cigar)):
        m = _SIMPLE.match(c)
        if m:
            lead = int(m.group(1) or 0)
            mid = int(m.group(2))
            r = int(p) - 1
            if r <= ref_col < r + mid:
                out[i] = lead + ref_col - r
        else:
            q = query_index_at(int(p), c, ref_col)
            out[i] = -1 if q is None else q
    return out
