"""Host layer with the reference's function signatures for the barcode-match + UMI path.

Mirrors /root/reference/utils.py and the two STAR scripts, same names, same arguments, same files
in `outdir`:

  write_bc_5p10X / write_bc_slideseq / write_bc_3p10XTCR_nuc / write_bc_3p10XGEX
        utils.py:604-622, 584-601, 1116-1132, 1412-1458   whitelist -> `{sample}_bcreads.fasta`
  barcode_ref(ref_fasta, genome_dir)
        scripts/barcode_ref.sh:11-18                       padded FASTA -> packed whitelist dir
  barcode_align(input_fasta_gz, genome_dir, out_name, threads, *ignored)
        scripts/barcode_align.sh:14-41                     candidates -> `{out_name}.sam`
  process_matching_5p10X / _5p10XTCR / _3p10XTCR_nuc / _3p10XGEX / _slideseq_XCR
        utils.py:685, 830, 1135, 1461, 625                 SAM -> CSV / JSON outputs
  sort_cnt                                                 utils.py:36-41

The compute (matching, UMI dedup) runs in the CUDA library; there is no CPU fallback.
Out of scope (SURVEY.md section 8f / DESIGN.md section 9): plots (matplotlib absent); MiXCR clone
tables are merged when their input files exist and skipped with a message otherwise -- never faked.
"""
from __future__ import annotations

import gzip
import json
import os

import numpy as np
import pandas as pd

from . import fastx, samio
from ._lib import NR_FLAG_RC, NR_FLAG_TOO_LONG, NR_MODE_AUTO, NR_UMI_NONE
from .h5lite import read_10x_h5_barcodes
from .matcher import Whitelist
from .whitelists import LINKER_SLIDESEQ

linker = LINKER_SLIDESEQ   # utils.py:14

_GENOME_FILE = "nr_whitelist.npz"


def sort_cnt(arr):
    """utils.py:36-41: value counts as a two-column frame sorted by count, descending."""
    vals, cnts = np.unique(np.asarray(arr), return_counts=True)
    dfcnt = pd.DataFrame({0: vals, 1: cnts.astype("int")})
    return dfcnt.sort_values(by=1, ascending=False, kind="stable")


# ---- whitelist -> padded FASTA ---------------------------------------------------------------------

def _write_padded(path, names, cores, left, right):
    with open(path, "w") as f:
        for n, c in zip(names, cores):
            f.write(f">{n}\n{'N' * left}{c}{'N' * right}\n")


def _read_bc_table(bc_file):
    return pd.read_table(bc_file, names=["bc"])


def write_bc_5p10X(sample, outdir, bc_file):
    """utils.py:604-622: `N`*30 + barcode + `N`*40 per whitelist line (text lists; `-1`
    suffixes stripped).  `.h5` inputs (:606-610: `sc.read_10x_h5` + `sc.pp.filter_cells(min_genes=20)`,
    first 16 characters of the cell names) are read by nanoranger_b200.h5lite -- no scanpy / h5py."""
    if ".h5" in bc_file:
        bcs = [a[:16] for a in read_10x_h5_barcodes(bc_file, 20)]
    elif ".txt" in bc_file:
        bcs = _read_bc_table(bc_file).bc.apply(lambda x: x.split("-")[0]).to_list()
    else:
        raise ValueError(f"unsupported whitelist file {bc_file}")
    _write_padded(f"{outdir}/{sample}_bcreads.fasta", bcs, bcs, 30, 40)


def write_bc_slideseq(sample, outdir, bc_file):
    """utils.py:584-601: N*15 + bc[:8] + linker + bc[8:] + N*24, names = np.unique'd barcodes."""
    left, right = 56 - 41, 56 - 32
    barcodes = _read_bc_table(bc_file)
    if "BeadBarcodes" in bc_file:
        bcs = np.unique(barcodes.bc.apply(lambda x: "".join(x.split(","))).to_list())
    elif "matched" in bc_file:
        bcs = np.unique(barcodes.bc.apply(lambda x: x.split("-")[0]).to_list())
    else:
        raise ValueError(f"unsupported slide-seq barcode file {bc_file}")
    cores = [b[:8] + linker + b[8:] for b in bcs]
    _write_padded(f"{outdir}/{sample}_bcreads.fasta", bcs, cores, left, right)


def write_bc_3p10XTCR_nuc(sample, outdir, bc_file):
    """utils.py:1116-1132: cells of a 10x .h5 with >= 4 genes (`sc.read_10x_h5` +
    `sc.pp.filter_cells(min_genes=4)` there; nanoranger_b200.h5lite here), N*16 + bc + N*28."""
    bcs = [a[:16] for a in read_10x_h5_barcodes(bc_file, 4)]
    _write_padded(f"{outdir}/{sample}_bcreads.fasta", bcs, bcs, 16, 28)


def write_bc_3p10XGEX(sample, outdir, barcodes):
    """utils.py:1412-1458: sum `split/*bc_count.json`, keep raw 16-mers seen in > 20 reads that are
    on the whitelist, N*4 + bc + N*17.  File contract kept as the reference writes it:
    `{sample}_bc_read_count.csv` is the pandas Series named `read_count` (barcode index, header
    `,read_count`) in the insertion order of the aggregated counts (:1434-1436); the FASTA lists
    the shared barcodes in that same order (:1446-1447); whitelist lines are taken verbatim -- no
    suffix stripping here (:1442)."""
    out = f"{outdir}/{sample}_bcreads.fasta"
    if os.path.isfile(out):
        print(out, " exists, skip")
        return
    split = f"{outdir}/split/"
    agg: dict[str, int] = {}
    for fn in sorted(f for f in os.listdir(split) if f.endswith("bc_count.json")):
        with open(os.path.join(split, fn)) as fh:
            for k, v in json.load(fh).items():
                agg[k] = agg.get(k, 0) + v
    read_cnt = pd.Series(agg, dtype="int64" if agg else "float64")
    read_cnt.name = "read_count"
    read_cnt.to_csv(f"{outdir}/{sample}_bc_read_count.csv")
    raw_bcs = read_cnt[read_cnt > 20]
    shared = set(_read_bc_table(barcodes)["bc"]) & set(raw_bcs.index)
    bcs = list(raw_bcs[raw_bcs.index.isin(shared)].index)
    _write_padded(out, bcs, bcs, 3 + 1, 12 + 4 + 1)


# ---- barcode_ref.sh -------------------------------------------------------------------------------

def barcode_ref(ref_fasta, genome_dir):
    """scripts/barcode_ref.sh <ref_fasta> <genome_dir>: instead of a STAR suffix-array index the
    'genome directory' holds the pad geometry and the core sequences; the GPU seed index is
    rebuilt from them in ~0.1 s when barcode_align opens the directory."""
    names, seqs, offsets = fastx.read_fasta(ref_fasta)
    if not names:
        raise ValueError(f"{ref_fasta}: no sequences")
    o = offsets.astype(np.int64)
    lens = np.diff(o)
    if not (lens == lens[0]).all():
        raise ValueError("padded whitelist records must all have the same length")
    mat = seqs.reshape(len(names), int(lens[0]))
    is_n = (mat == ord("N")) | (mat == ord("n"))
    # pads = all-N columns at both ends (identical for every record by construction)
    col_all_n = is_n.all(axis=0)
    left = 0
    while left < mat.shape[1] and col_all_n[left]:
        left += 1
    right = 0
    while right < mat.shape[1] - left and col_all_n[mat.shape[1] - 1 - right]:
        right += 1
    cores = np.ascontiguousarray(mat[:, left:mat.shape[1] - right])
    if cores.shape[1] == 0 or cores.shape[1] > 32:
        raise ValueError(f"core length {cores.shape[1]} not supported (1..32)")
    os.makedirs(genome_dir, exist_ok=True)
    np.savez_compressed(os.path.join(genome_dir, _GENOME_FILE), cores=cores,
                        names=np.array(names), pad_l=left, pad_r=right)
    return genome_dir


def load_genome(genome_dir, device: int = 0):
    d = np.load(os.path.join(genome_dir, _GENOME_FILE), allow_pickle=False)
    wl = Whitelist(d["cores"], int(d["pad_l"]), int(d["pad_r"]), device=device)
    return wl, [str(x) for x in d["names"]]


# ---- barcode_align.sh -----------------------------------------------------------------------------

def match_records(wl: Whitelist, names, seqs, offsets, ref_names, mode=NR_MODE_AUTO,
                  min_score=None):
    """Runs the matcher and yields the SAM records STAR would print: one per candidate whose best
    score is reached by exactly one (entry, strand) pair (--outFilterMultimapNmax 1,
    --outFilterMultimapScoreRange 0, scripts/barcode_align.sh:21-22); other candidates are absent
    (default --outSAMunmapped None)."""
    if min_score is None:
        min_score = wl.core_len - 2
    res = wl.match_host(seqs, offsets, min_score=min_score, mode=mode)
    o = offsets.astype(np.int64)
    raw = seqs.tobytes()
    keep = np.flatnonzero((res.nbest == 1) & ((res.flags & NR_FLAG_TOO_LONG) == 0))
    for i in keep:
        s = raw[o[i]:o[i + 1]]
        rc = bool(res.flags[i] & NR_FLAG_RC)
        u = int(res.umi_q[i])
        pos, cig = samio.anchored_alignment(len(s), -1 if (u == NR_UMI_NONE or rc) else u,
                                            wl.pad_l, wl.core_len, wl.pad_r)
        if rc:
            s = samio.revcomp_bytes(s)
        yield (names[i], 16 if rc else 0, int(res.idx[i]), pos, cig, s.decode("ascii"),
               int(res.score[i]))


def barcode_align(input_fastq, genome_dir, out_name, threads=1, *ignored, device: int = 0,
                  header: str = "used", mode=NR_MODE_AUTO, alignments: str = "traceback"):
    """scripts/barcode_align.sh <input.fa.gz> <genome_dir> <out_prefix> <threads> [ignored]:
    writes `<out_prefix>.sam`.  `threads` is accepted for call compatibility (the work runs on
    the GPU).

    mode: NR_MODE_AUTO (default, at any input size) resolves EVERY candidate exactly, as the
    reference's STAR call does (--outFilterScoreMinOverLread 0): reads the seed filter cannot
    decide go through the deep tier (exact optimum over the whole whitelist), so the low-score
    tail of `_barcode_scores.csv` (utils.py:698, 728-730) is the oracle's.  NR_MODE_FILTERED is
    opt-in: it resolves exactly everything the reference keeps (AS >= core length - 2) and leaves
    the rest out of the SAM.

    alignments: "traceback" (default) writes real POS / CIGAR and the `AS nM MD` attributes
    scripts/barcode_align.sh:21 asks STAR for (host tracebacks of the kept records, `threads`
    workers, all cores when threads <= 1); "anchored" writes the fast form -- an ungapped M run
    placed so that the reference's aligned_pairs lookup (utils.py:705-708) finds the same UMI."""
    import ctypes as C
    from . import _lib
    from ._lib import NR_MODE_FILTERED
    wl, ref_names = load_genome(genome_dir, device)
    try:
        # no per-record Python from here to the file: raw FASTA buffers -> nr_match_host ->
        # nr_sam_write (same bytes as samio.write_sam(match_records(...)), ~50x faster)
        nbuf, noff, seqs, offsets = fastx.read_fasta_raw(input_fastq)
        n = len(offsets) - 1
        res = wl.match_host(seqs, offsets, min_score=wl.core_len - 2, mode=mode)
        rbuf, roff = fastx._pack_names(ref_names)
        if n == 0:
            nbuf = seqs = np.zeros(1, np.uint8)
        written = C.c_uint64(0)
        if alignments == "traceback":
            try:
                nthr = int(threads)
            except (TypeError, ValueError):
                nthr = 0
            _lib.check(_lib.lib().nr_sam_write_aligned(
                wl.handle, f"{out_name}.sam".encode(), 1 if header == "full" else 0, nbuf.ctypes.data,
                noff.ctypes.data, seqs.ctypes.data, offsets.ctypes.data, n, res.idx.ctypes.data,
                res.score.ctypes.data, res.nbest.ctypes.data, res.flags.ctypes.data,
                res.umi_q.ctypes.data, rbuf.ctypes.data, roff.ctypes.data, len(ref_names),
                nthr if nthr > 1 else 0, C.byref(written)), "nr_sam_write_aligned")
        elif alignments == "anchored":
            _lib.check(_lib.lib().nr_sam_write(
                f"{out_name}.sam".encode(), 1 if header == "full" else 0, nbuf.ctypes.data,
                noff.ctypes.data, seqs.ctypes.data, offsets.ctypes.data, n, res.idx.ctypes.data,
                res.score.ctypes.data, res.nbest.ctypes.data, res.flags.ctypes.data,
                res.umi_q.ctypes.data, rbuf.ctypes.data, roff.ctypes.data, len(ref_names), wl.pad_l,
                wl.core_len, wl.pad_r, C.byref(written)), "nr_sam_write")
        else:
            raise ValueError(f"alignments must be 'traceback' or 'anchored', not {alignments!r}")
        n = int(written.value)
    finally:
        wl.close()
    return n


# ---- SAM -> assignments --------------------------------------------------------------------------

def _parse_matching(sample, outdir, thr, umi_ref_col, umi_len, exact_len):
    """Common front half of process_matching_* (utils.py:697-718 and twins): AS list over all
    records, accepted (name, bc, umi) triples, short-UMI count."""
    t = samio.read_sam_table(f"{outdir}/{sample}_matching.sam")
    all_AS = np.stack([t["AS"], t["flag"]], axis=1).astype(np.int64).reshape(-1, 2)
    sel = np.flatnonzero((t["AS"] >= thr) & (t["flag"] == 0))
    q = samio.query_index_at_many(t["pos"][sel], t["cigar"][sel], umi_ref_col)
    triples, bad = [], 0
    for i, qi in zip(sel, q):
        # aligned_pairs lookup failed -> the reference's `except: umi = "N"` (utils.py:709-710)
        umi = t["seq"][i][qi:qi + umi_len] if qi >= 0 else "N"
        if (len(umi) != umi_len) if exact_len else (len(umi) < umi_len):
            bad += 1
        else:
            triples.append((t["qname"][i], t["rname"][i], umi))
    print("number of short UMI reads = ", bad)
    return all_AS, triples


def _write_scores(all_AS, path):
    fwd = all_AS[all_AS[:, 1] == 0][:, 0] if len(all_AS) else np.zeros(0, np.int64)
    scores = sort_cnt(fwd)
    scores.columns = ["score", "count"]
    scores.to_csv(path, index=None)
    return scores


def umi_dedup_table(bcs, umis, umi_len, device: int = 0):
    """utils.py:759-773: per barcode, number of distinct UMIs, number of reads, their ratio;
    sorted by umi_cnt descending.  Distinct counting runs in nr_umi_collapse_device
    (max_dist 0, gene 0 == np.unique per barcode)."""
    from . import umi as U
    if not bcs:
        return pd.DataFrame({"umi_cnt": [], "read_cnt": [], "dup_rate": []},
                            index=pd.Index([], name="bc"))
    names, inv = np.unique(np.array(bcs), return_inverse=True)
    first = {}
    for i, b in enumerate(bcs):               # dict insertion order of the reference
        first.setdefault(b, i)
    codes, ok = U.pack_umis(list(umis), umi_len)
    # UMIs with N cannot be packed: give each distinct such string its own code above 2^(2*len)
    if not ok.all():
        if umi_len > 15:
            raise ValueError("UMIs containing N need umi_len <= 15 to be representable")
        extra = {s: k for k, s in enumerate(sorted({umis[i] for i in np.flatnonzero(~ok)}))}
        for i in np.flatnonzero(~ok):
            codes[i] = np.uint32((1 << (2 * umi_len)) + extra[umis[i]])
    # (nr_umi_collapse_device sorts all 32 bits of the UMI word, so the escape codes need no
    # widened umi_len; the device-resident records path, nr_umi_records_device, cannot carry such
    # UMIs in 2 bit/base and counts them in its stats[2] instead -- the count matrix path therefore
    # leaves N-containing UMIs out, this per-barcode table keeps them like the reference does)
    r = U.collapse_host(inv.astype(np.uint32), np.zeros(len(inv), np.uint32), codes, umi_len, 0, device)
    umi_cnt = np.bincount(r["g_bc"].astype(np.int64), minlength=len(names))
    read_cnt = np.bincount(r["g_bc"].astype(np.int64), weights=r["g_reads"],
                           minlength=len(names)).astype(np.int64)
    order = sorted(range(len(names)), key=lambda k: first[names[k]])
    df = pd.DataFrame({"bc": names[order], "umi_cnt": umi_cnt[order].astype("int"),
                       "read_cnt": read_cnt[order].astype("int")})
    df = df.sort_values(by="umi_cnt", ascending=False, kind="stable").set_index("bc")
    df["dup_rate"] = df.read_cnt / df.umi_cnt
    return df


def _per_name(triples):
    """utils.py:718: `read_bcumi_dic[name] = [bc, umi]` -- last record per name wins, dict keeps
    first-insertion order."""
    d = {}
    for n, b, u in triples:
        d[n] = (b, u)
    return d


def _skip(what):
    print(f"nanoranger_b200: {what} -- skipped (out of scope this round, see DESIGN.md)")


def _clone_merge(sample, outdir, triples, clone_first):
    path = f"{outdir}/{sample}_cloneID_filtered.csv.gz"
    if not os.path.isfile(path):
        _skip(f"{path} missing (MiXCR output), clone_bcumi table")
        return
    cloneID = pd.read_csv(path, index_col=0)
    df = pd.DataFrame(triples, columns=["ID", "bc", "umi"]).set_index("ID")
    merged = (pd.merge(cloneID, df, how="inner", left_index=True, right_index=True) if clone_first
              else pd.merge(df, cloneID, how="inner", left_index=True, right_index=True))
    merged = merged.sort_values(by=["cloneId", "bc", "umi"])
    with gzip.open(f"{outdir}/{sample}_clone_bcumi.csv.gz", "wt") as f:
        merged.to_csv(f, index=None)


def process_matching_5p10X(sample, outdir, device: int = 0):
    """utils.py:685-827 (AS >= 14, UMI 10 nt at reference column 46)."""
    all_AS, triples = _parse_matching(sample, outdir, 14, 46, 10, exact_len=False)
    _write_scores(all_AS, f"{outdir}/{sample}_barcode_scores.csv")
    per = _per_name(triples)
    bcs = [v[0] for v in per.values()]
    umis = [v[1] for v in per.values()]
    ded = umi_dedup_table(bcs, umis, 10, device)
    ded[ded.umi_cnt > 0].to_csv(f"{outdir}/{sample}_bcumi_dedup.csv")
    # name -> (CB, UB, XT) as the reference builds it before tagging the genome BAM
    table = {n: (b, u, n.split("_")[4] if n.count("_") >= 4 else "") for n, (b, u) in per.items()}
    with open(f"{outdir}/{sample}_read_tags.tsv", "w") as f:
        for n, (b, u, t) in table.items():
            f.write(f"{n}\t{b}\t{u}\t{t}\n")
    # utils.py:801-827: tag the genome alignments of the assigned reads, count their transcripts
    if os.path.isfile(f"{outdir}/{sample}_genome.bam"):
        from . import bamio
        all_trns = bamio.tag_genome_bam(f"{outdir}/{sample}_genome.bam",
                                        f"{outdir}/{sample}_genome_tagged.bam", table)
        sort_cnt(all_trns).to_csv(f"{outdir}/{sample}_trns_ct.csv", index=None)
    else:
        _skip(f"{outdir}/{sample}_genome.bam missing, CB/UB/XT tagging")
    return table


def process_matching_5p10XTCR(sample, outdir, device: int = 0):
    """utils.py:830-979 (AS >= 14, UMI 12 nt at reference column 46, clone merge)."""
    all_AS, triples = _parse_matching(sample, outdir, 14, 46, 12, exact_len=False)
    _write_scores(all_AS, f"{outdir}/{sample}_barcode_scores.csv")
    per = _per_name(triples)
    ded = umi_dedup_table([v[0] for v in per.values()], [v[1] for v in per.values()], 12, device)
    ded[ded.umi_cnt > 0].to_csv(f"{outdir}/{sample}_bcumi_dedup.csv")
    _clone_merge(sample, outdir, triples, clone_first=False)


def process_matching_3p10XTCR_nuc(sample, outdir, device: int = 0):
    """utils.py:1135-1280 (AS >= 14, UMI 12 nt at reference column 32)."""
    all_AS, triples = _parse_matching(sample, outdir, 14, 32, 12, exact_len=False)
    _write_scores(all_AS, f"{outdir}/{sample}_barcode_scores.csv")
    per = _per_name(triples)
    ded = umi_dedup_table([v[0] for v in per.values()], [v[1] for v in per.values()], 12, device)
    ded[ded.umi_cnt > 0].to_csv(f"{outdir}/{sample}_bcumi_dedup.csv")
    _clone_merge(sample, outdir, triples, clone_first=False)


def process_matching_3p10XGEX(sample, outdir):
    """utils.py:1461-1520 (AS >= 14, UMI 12 nt at reference column 20): `{sample}_quads.json` =
    {barcode: [[umi, transcript], ...]} and `{sample}_barcode_scores.csv`."""
    quads_json = f"{outdir}/{sample}_quads.json"
    if os.path.isfile(quads_json):
        print(quads_json, " exists, skip")
        return
    all_AS, triples = _parse_matching(sample, outdir, 14, 20, 12, exact_len=False)
    quad: dict[str, list] = {}
    for name, bc, umi in triples:
        trns = "_".join(name.split("/ccs_")[-1].split("_")[3:])
        if "|" in trns:
            trns = trns.split("|")[-1].split("_")[0]
        quad.setdefault(bc, []).append([umi, trns])
    with open(quads_json, "w") as f:
        json.dump(quad, f)
    _write_scores(all_AS, f"{outdir}/{sample}_barcode_scores.csv")


def process_matching_slideseq_XCR(sample, outdir, cloneID):
    """utils.py:625-682 (AS >= 30, UMI exactly 9 nt at reference column 47, merged with the
    cloneID frame; the score histogram only feeds a plot there, written here as CSV instead)."""
    all_AS, triples = _parse_matching(sample, outdir, 30, 47, 9, exact_len=True)
    _write_scores(all_AS, f"{outdir}/{sample}_barcode_scores.csv")
    df = pd.DataFrame(triples, columns=["ID", "bc", "umi"]).set_index("ID")
    if cloneID is None:
        _skip("cloneID frame not given, clone_bcumi table")
        return df
    merged = pd.merge(cloneID, df, how="inner", left_index=True, right_index=True)
    merged = merged.sort_values(by=["cloneId", "bc", "umi"])
    with gzip.open(f"{outdir}/{sample}_clone_bcumi.csv.gz", "wt") as f:
        merged.to_csv(f, index=None)
    return merged


def make_count_mtx_3p10XGEX(sample, outdir, max_dist: int = 0, device: int = 0):
    """utils.py:1523-1548 is unfinished in the reference (merges `split/*quads.json`, then stops).
    This finishes it: (barcode, transcript) UMI counts after collapse, written as
    `{sample}_gex.mtx.gz` (MatrixMarket, rows = transcripts, columns = barcodes) with
    `{sample}_gex_barcodes.tsv` / `{sample}_gex_features.tsv`."""
    from . import umi as U
    mtx_file = f"{outdir}/{sample}_gex.mtx.gz"
    if os.path.isfile(mtx_file):
        print(mtx_file, " exists, skip")
        return
    split = f"{outdir}/split/"
    agg: dict[str, list] = {}
    for fn in sorted(f for f in os.listdir(split) if f.endswith("quads.json")):
        with open(os.path.join(split, fn)) as fh:
            for k, v in json.load(fh).items():
                agg.setdefault(k, []).extend(v)
    bcs = sorted(agg)
    genes = sorted({t for v in agg.values() for _, t in v})
    gi = {g: i for i, g in enumerate(genes)}
    b_idx, g_idx, umis = [], [], []
    for bi, b in enumerate(bcs):
        for u, t in agg[b]:
            if set(u) <= set("ACGT"):
                b_idx.append(bi); g_idx.append(gi[t]); umis.append(u)
    codes, _ = U.pack_umis(umis, 12)
    r = U.collapse_host(np.array(b_idx, np.uint32), np.array(g_idx, np.uint32), codes, 12,
                        max_dist, device)
    pair = r["g_bc"].astype(np.int64) * len(genes) + r["g_gene"].astype(np.int64)
    up, cnt = np.unique(pair, return_counts=True)
    with gzip.open(mtx_file, "wt") as f:
        f.write("%%MatrixMarket matrix coordinate integer general\n")
        f.write(f"{len(genes)} {len(bcs)} {len(up)}\n")
        for p, c in zip(up, cnt):
            f.write(f"{p % len(genes) + 1} {p // len(genes) + 1} {c}\n")
    with open(f"{outdir}/{sample}_gex_barcodes.tsv", "w") as f:
        f.write("\n".join(bcs) + "\n")
    with open(f"{outdir}/{sample}_gex_features.tsv", "w") as f:
        f.write("\n".join(genes) + "\n")
