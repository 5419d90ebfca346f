"""Adapter-motif search kernel (nr_hw_search_*) vs the oracle, and the decon_* extractors on a
synthetic minimap2-style SAM vs the reference's loop restated with the oracle's search."""
import gzip

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

MOTIF_GEX = "CGCTCTTCCGATCT" + 26 * "N" + "TTTCTTATATG"     # utils.py:103
MOTIF_TCR = "CGCTCTTCCGATCT" + 28 * "N" + "TTTCTTATATG"     # utils.py:202


def _rs(rng, n, alpha="ACGT"):
    return "".join(alpha[i] for i in rng.integers(0, len(alpha), n))


def _mutate(rng, s, nops):
    c = list(s)
    for _ in range(nops):
        if not c:
            break
        j = int(rng.integers(0, len(c)))
        op = rng.random()
        if op < 0.34:
            c[j] = "ACGT"[rng.integers(0, 4)]
        elif op < 0.67:
            del c[j]
        else:
            c.insert(j, "ACGT"[rng.integers(0, 4)])
    return "".join(c)


def _windows(rng, pat, n):
    out = []
    for i in range(n):
        inst = "".join(ch if ch != "N" else "ACGT"[rng.integers(0, 4)] for ch in pat)
        kind = rng.random()
        alpha = "ACGTN" if i % 5 == 0 else "ACGT"
        if kind < 0.15:
            w = _rs(rng, int(rng.integers(0, 200)), alpha)
        elif kind < 0.3:       # motif twice
            w = _rs(rng, 10, alpha) + _mutate(rng, inst, 1) + _rs(rng, 7, alpha) + _mutate(rng, inst, 1) + _rs(rng, 5)
        else:
            w = (_rs(rng, int(rng.integers(0, 120)), alpha) + _mutate(rng, inst, int(rng.integers(0, 9)))
                 + _rs(rng, int(rng.integers(0, 40)), alpha))
        if kind > 0.9:
            w = w[: int(rng.integers(0, len(w) + 1))]     # truncated
        out.append(w)
    return out


@pytest.mark.parametrize("pat,k,wild", [
    (MOTIF_GEX, 6, True), (MOTIF_TCR, 6, True), ("TCTCGGGAACGCTGAAGA", 2, False),
    ("AGATCGGAAGAGCGTCGTGT", 3, False), ("AAAAAA" + 28 * "N" + "AGATCGGAAGAGCGTCGTGT", 5, True),
    ("ACGT" * 16, 10, False), ("ACGTN", 1, True), ("A", 0, False)])
def test_hw_search_vs_oracle(cuda_device, oracle, pat, k, wild):
    from nanoranger_b200 import extract
    rng = np.random.default_rng(len(pat) * 7 + k)
    wins = _windows(rng, pat, 1500) + ["", "N" * 30, pat.replace("N", "A")]
    res = extract.hw_search(wins, pat, k, wild)
    for i, w in enumerate(wins):
        ref = oracle.hw_search(pat, w, k, wild)
        got = {"editDistance": int(res["ed"][i]), "n_locations": int(res["nloc"][i]),
               "first": tuple(int(x) for x in res["first"][i]), "last": tuple(int(x) for x in res["last"][i])}
        assert got == ref, (i, w, got, ref)
    # device-resident entry point gives the same arrays
    import torch
    from nanoranger_b200 import pack_ascii
    buf, off = pack_ascii(wins)
    dev = torch.device("cuda:0")
    r2 = extract.hw_search_device(torch.from_numpy(buf.copy()).to(dev),
                                  torch.from_numpy(off.view(np.int64).copy()).to(dev), pat, k, wild)
    for key in ("ed", "first", "last", "nloc"):
        assert np.array_equal(r2[key].cpu().numpy(), res[key])


def test_hw_search_argument_errors(cuda_device):
    from nanoranger_b200 import extract
    with pytest.raises(RuntimeError):
        extract.hw_search(["ACGT"], "A" * 65, 2)
    with pytest.raises(RuntimeError):
        extract.hw_search(["ACGT"], "ACGT", 4)           # k must be < len(pattern)
    assert len(extract.hw_search([], "ACGT", 1)["ed"]) == 0


def _synthetic_trns_sam(path, rng, n, motif_n):
    """minimap2 -aY --eqx style records: [junk][adapter bc umi TSO][aligned part][tail], soft
    clips on both sides, forward / reverse / supplementary flags."""
    lines = ["@HD\tVN:1.6\tSO:unsorted\n", "@SQ\tSN:TRBC1-201|ENST0001.1_900\tLN:900\n"]
    for i in range(n):
        bc_umi = _rs(rng, motif_n)
        flank = "CGCTCTTCCGATCT" + bc_umi + "TTTCTTATATG"
        flank = _mutate(rng, flank, int(rng.integers(0, 5)))
        if rng.random() < 0.15:
            flank = _rs(rng, len(flank))                    # no motif at all
        junk = _rs(rng, int(rng.integers(0, 260)))
        body = _rs(rng, int(rng.integers(40, 500)))
        tail = _rs(rng, int(rng.integers(0, 120)))
        seq = junk + flank + body + tail
        lead, trail = len(junk) + len(flank), len(tail)
        if rng.random() < 0.2:                              # alignment swallowed part of the flank
            lead -= int(rng.integers(1, 15))
        cigar = (f"{lead}S" if lead else "") + f"{len(seq) - lead - trail}=" + (f"{trail}S" if trail else "")
        flag = [0, 16, 2048, 2064][int(rng.integers(0, 4))]
        qual = "".join(chr(33 + int(x)) for x in rng.integers(2, 40, len(seq)))
        lines.append(f"read{i:05d}-{_rs(rng, 4).lower()}\t{flag}\tTRBC1-201|ENST0001.1_900\t{int(rng.integers(1, 300))}\t60\t"
                     f"{cigar}\t*\t0\t0\t{seq}\t{qual}\tNM:i:0\tAS:i:{len(body)}\n")
    with open(path, "w") as f:
        f.writelines(lines)


def _ref_decon_5p(oracle, extract, sam, const, lclip, rclip, tcr):
    """utils.py:95-189 / 192-310 restated record by record with the oracle's search."""
    fq, fa, eds = [], [], []
    for r in extract.read_alignments(sam):
        qs, qe = r.query_alignment_start, r.query_alignment_end
        beg = r.seq[qs - lclip: qs + rclip] if qs > lclip else r.seq[: qs + rclip]
        if r.flag in (16, 2064):
            qsm, qem = r.rlen - qe, r.rlen - qs
        else:
            qsm, qem = qs, qe
        name = f"{r.qname}_{qsm}_{qem}_{r.flag}_{r.reference_name}"
        if tcr:
            ss = qe - 60 if r.qlen > 60 else qs
            se = qe + 80 if r.rlen - qe > 80 else r.rlen
            fq.append(f"@{name}\n{r.seq[ss:se]}\n+\n{r.qual[ss:se]}\n")
        ed = oracle.hw_search(const, beg, 6, True)
        if -1 < ed["editDistance"] < 7:
            start, end = ed["last"]
            bcumi = beg[start:end]
            start = lclip - start if qs > lclip else qs - start
            eds.append([start, r.reference_name, len(bcumi), ed["editDistance"]])
            if not tcr:
                fq.append(f"@{name}\n{r.seq[qs:qe]}\n+\n{r.qual[qs:qe]}\n")
            fa.append(f">{name}\n{bcumi}\n")
    return "".join(fq), "".join(fa), eds


@pytest.mark.parametrize("mode", ["5p10XGEX", "5p10XTCR"])
def test_decon_5p_files(cuda_device, oracle, tmp_path, mode):
    from nanoranger_b200 import extract
    rng = np.random.default_rng(31 if mode == "5p10XGEX" else 32)
    tcr = mode == "5p10XTCR"
    _synthetic_trns_sam(tmp_path / "s_trns.sam", rng, 700, 28 if tcr else 26)
    fn = extract.decon_5p10XTCR if tcr else extract.decon_5p10XGEX
    n = fn("s", str(tmp_path))
    fq, fa, eds = _ref_decon_5p(oracle, extract, str(tmp_path / "s_trns.sam"), MOTIF_TCR if tcr else MOTIF_GEX,
                                200 if tcr else 80, 20, tcr)
    assert gzip.open(tmp_path / "s_deconcat.fastq.gz", "rt").read() == fq
    assert gzip.open(tmp_path / "s_BCUMI.fasta.gz", "rt").read() == fa
    assert n == fa.count(">") and n > 300
    if tcr:
        import pandas as pd
        got = pd.read_csv(tmp_path / "s_eds.csv", index_col=0)
        assert got.values.tolist() == [[a, b, c, d] for a, b, c, d in eds]
    # the FASTA feeds the matcher's reader
    from nanoranger_b200 import fastx
    names, seqs, off = fastx.read_fasta(str(tmp_path / "s_BCUMI.fasta.gz"))
    assert len(names) == n and names[0].split("_")[4].startswith("TRBC1-201|ENST0001.1")


def test_decon_slideseq_files(cuda_device, oracle, tmp_path):
    from nanoranger_b200 import extract
    from nanoranger_b200.whitelists import LINKER_SLIDESEQ
    rng = np.random.default_rng(40)
    lines = ["@HD\tVN:1.6\n", "@SQ\tSN:TRAC\tLN:2000\n"]
    for i in range(500):
        bc = _rs(rng, 14)
        struct = extract.rev(_rs(rng, 22)[:22] + bc[:8] + LINKER_SLIDESEQ + bc[8:] + _rs(rng, 9) + "T" * 12)
        struct = _mutate(rng, struct, int(rng.integers(0, 3)))
        pre = _rs(rng, int(rng.integers(0, 300)))
        body = _rs(rng, int(rng.integers(350, 600)))
        polya = "A" * int(rng.integers(0, 90))
        seq = pre + body + polya + struct + _rs(rng, int(rng.integers(0, 60)))
        trail = len(seq) - len(pre) - len(body)
        cigar = (f"{len(pre)}S" if pre else "") + f"{len(body)}=" + f"{trail}S"
        qual = "I" * len(seq)
        lines.append(f"r{i}\t{[0, 16][i % 2]}\tTRAC\t5\t60\t{cigar}\t*\t0\t0\t{seq}\t{qual}\tAS:i:300\n")
    (tmp_path / "x_trns.sam").write_text("".join(lines))
    n = extract.decon_3pXCR_slideseq("x", str(tmp_path))
    # restated reference loop (utils.py:393-481) with the oracle's search
    const = extract.rev(LINKER_SLIDESEQ)
    exp = []
    for r in extract.read_alignments(str(tmp_path / "x_trns.sam")):
        qs, qe = r.query_alignment_start, r.query_alignment_end
        dd = r.seq[qe: qe + 200] if r.rlen - qe > 200 else r.seq[qe:]
        sub_s = qs - 200 if qs > 200 else 0
        if len(r.seq[sub_s:qs + 80]) > 100 and r.reference_end - r.reference_start > 400:
            for i in range(int(len(dd) / 20)):
                ed = oracle.hw_search(const, dd[20 * i: 20 * i + 40], 2, False)
                if -1 < ed["editDistance"] < 4:
                    start, end = ed["first"][0] + 20 * i, ed["first"][1] + 20 * i
                    up = max(0, start - 22)
                    hang, pa = extract.rev(dd[up:end + 16]), dd[: up + 5]
                    ok = len(hang) > 45 and (len(pa) < 70 or (len(pa) > 70 and pa.count("A") / len(pa) > 0.5))
                    if ok:
                        exp.append(f">{r.qname}_x_{sub_s}_{qs + 80}_{r.flag}_{r.reference_name}\n{hang}\n")
                    break
    assert gzip.open(tmp_path / "x_BCUMI.fasta.gz", "rt").read() == "".join(exp)
    assert n == len(exp) and n > 200


def _tail_sam(path, rng, n, adapter, before_len):
    """records whose 3' soft clip holds [polyA][before_len random nt][adapter][junk], sometimes
    far downstream (second stepped window) or absent."""
    lines = ["@HD\tVN:1.6\n", "@SQ\tSN:TRBV7-9-201|ENST9.1_700\tLN:700\n"]
    for i in range(n):
        pre = _rs(rng, int(rng.integers(0, 30)))
        body = _rs(rng, int(rng.integers(30, 400)))
        gap = _rs(rng, int(rng.choice([0, 5, 150, 260])))
        ad = _mutate(rng, adapter, int(rng.integers(0, 4))) if rng.random() > 0.1 else ""
        tail = "A" * int(rng.integers(0, 25)) + gap + _rs(rng, before_len) + ad + _rs(rng, int(rng.integers(0, 50)))
        seq = pre + body + tail
        cigar = (f"{len(pre)}S" if pre else "") + f"{len(body)}=" + (f"{len(tail)}S" if tail else "")
        flag = [0, 16, 2048, 2064][i % 4]
        lines.append(f"m64012_{i}/ccs\t{flag}\tTRBV7-9-201|ENST9.1_700\t3\t60\t{cigar}\t*\t0\t0\t{seq}\t{'F' * len(seq)}\tAS:i:50\n")
    with open(path, "w") as f:
        f.writelines(lines)


def test_decon_3p10XGEX_files(cuda_device, oracle, tmp_path):
    import json
    from nanoranger_b200 import extract
    rng = np.random.default_rng(50)
    ad = "AGATCGGAAGAGCGTCGTGT"
    _tail_sam(tmp_path / "p_trns.sam", rng, 600, ad, 32)
    n = extract.decon_3p10XGEX("p", str(tmp_path))
    fa, counts = [], {}
    for r in extract.read_alignments(str(tmp_path / "p_trns.sam")):        # utils.py:1313-1383
        qs, qe = r.query_alignment_start, r.query_alignment_end
        end_qu = r.seq[qe - 70: qe + 700] if r.rlen - qe > 700 else r.seq[qe - 70:]
        sub_e = qe + 1 if r.rlen - qe > 1 else r.rlen
        sub_s = 0 if qs < 1 else qs - 1
        qsm, qem = (r.rlen - qe, r.rlen - qs) if r.flag in (16, 2064) else (qs, qe)
        if len(r.seq[sub_s:sub_e]) > 50:
            for i in range(int(len(end_qu) / 200) + 1):
                ed = oracle.hw_search(ad, end_qu[200 * i: 200 * (i + 1) + 70], 3, False)
                if -1 < ed["editDistance"] < 4:
                    start = ed["first"][0] + 200 * i
                    bcumi = extract.rev(end_qu[start - 32: start + 3])
                    counts[bcumi[3:19]] = counts.get(bcumi[3:19], 0) + 1
                    if len(bcumi) > 30:
                        fa.append(f">{r.qname}_{qsm}_{qem}_{r.flag}_{r.reference_name}\n{bcumi}\n")
                    break
    assert gzip.open(tmp_path / "p_BCUMI.fasta.gz", "rt").read() == "".join(fa)
    assert json.load(open(tmp_path / "p_bc_count.json")) == counts
    assert n == len(fa) and n > 300
    assert extract.decon_3p10XGEX("p", str(tmp_path)) is None              # "exists, skip"


def test_decon_3p10XTCR_variants_files(cuda_device, oracle, tmp_path):
    from nanoranger_b200 import extract
    rng = np.random.default_rng(51)
    ad = "AGATCGGAAGAGCGTCGTGT"
    _tail_sam(tmp_path / "n_trns.sam", rng, 500, ad, 35)
    n = extract.decon_3p10XTCR_nuc("n", str(tmp_path))
    fa = []
    for r in extract.read_alignments(str(tmp_path / "n_trns.sam")):        # utils.py:998-1083
        qs, qe = r.query_alignment_start, r.query_alignment_end
        end_qu = r.seq[qe: qe + 2000] if r.rlen - qe > 2000 else r.seq[qe:]
        sub_e = qe + 100 if r.rlen - qe > 100 else r.rlen
        if len(r.seq[qs:sub_e]) > 100:
            for i in range(int(len(end_qu) / 200) + 1):
                ed = oracle.hw_search(ad, end_qu[200 * i: 200 * (i + 1) + 70], 2, False)
                if -1 < ed["editDistance"] < 3:
                    start, end = ed["first"][0] + 200 * i, ed["first"][1] + 200 * i
                    bcumi = extract.rev(end_qu[start - 35: end - 12])
                    if len(bcumi) > 30:
                        fa.append(f">{r.qname}_n_{qs}_{sub_e}_{r.flag}_{r.reference_name.split('-')[0]}\n{bcumi}\n")
                    break
    assert gzip.open(tmp_path / "n_BCUMI.fasta.gz", "rt").read() == "".join(fa)
    assert n == len(fa) and n > 150
    # legacy 3p10XTCR: one window, wildcard motif
    const = 6 * "A" + 28 * "N" + ad
    _tail_sam(tmp_path / "t_trns.sam", rng, 300, ad, 40)
    extract.decon_3p10XTCR("t", str(tmp_path))
    fa = []
    for r in extract.read_alignments(str(tmp_path / "t_trns.sam")):        # utils.py:316-357
        qs, qe = r.query_alignment_start, r.query_alignment_end
        end_qu = r.seq[qe: qe + 150] if r.rlen - qe > 150 else r.seq[qe:]
        sub_s = qs - 350 if qs > 350 else 0
        ed = oracle.hw_search(const, end_qu, 5, True)
        d = ed["editDistance"]
        if -1 < d < 6 and len(r.seq[sub_s:qs + 100]) > 100 and r.qlen > 100:
            nm = f"{r.qname[-10:]}_q{r.qlen}_d{d}_s{sub_s}_e{qs + 100}_f{r.flag}_{r.reference_name.split('-')[0]}"
            fa.append(f">{nm}\n{extract.rev(end_qu[ed['first'][0]:ed['first'][1]])[14:]}\n")
    assert gzip.open(tmp_path / "t_BCUMI.fasta.gz", "rt").read() == "".join(fa)
