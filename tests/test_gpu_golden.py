"""Committed fixtures (tests/golden, made by make_golden.py from the reference's sample FASTQs):
candidates through the reference-shaped host layer (barcode_ref -> barcode_align -> SAM ->
process_matching_*) on the GPU, compared with the stored oracle results.  Bit-exact."""
import gzip
import os

import numpy as np
import pandas as pd
import pytest

from helpers import compare

pytestmark = pytest.mark.gpu
G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _load(name):
    from nanoranger_b200 import fastx
    names, seqs, off = fastx.read_fasta(os.path.join(G, f"{name}.fa.gz"))
    ref = dict(np.load(os.path.join(G, f"{name}.oracle.npz")))
    return names, seqs, off, ref


@pytest.mark.parametrize("name", ["tcr3", "mtdna1026"])
def test_5p_fixtures_all_modes(cuda_device, name):
    from nanoranger_b200 import (NR_MODE_AUTO, NR_MODE_EXHAUSTIVE, NR_MODE_FILTERED, Whitelist,
                                 whitelists)
    names, seqs, off, ref = _load(name)
    wl = Whitelist(whitelists.load_737k(), 30, 40)
    r = wl.match_host(seqs, off, min_score=14, mode=NR_MODE_FILTERED)
    assert compare(ref, r, 14, exact_below=False, label=f"{name} filtered") > 2000
    r = wl.match_host(seqs, off, min_score=14, mode=NR_MODE_AUTO)
    compare(ref, r, 14, exact_below=True, label=f"{name} auto")
    sub = {k: (v[:300] if getattr(v, "ndim", 0) else v) for k, v in ref.items()}
    r = wl.match_host(seqs[:int(off[300])], off[:301], min_score=14, mode=NR_MODE_EXHAUSTIVE)
    compare(sub, r, 14, exact_below=True, label=f"{name} exhaustive")


def test_slideseq_fixture(cuda_device):
    from nanoranger_b200 import NR_MODE_AUTO, Whitelist
    from nanoranger_b200.whitelists import LINKER_SLIDESEQ
    names, seqs, off, ref = _load("slideseq")
    bcs = gzip.open(os.path.join(G, "slideseq_whitelist.txt.gz"), "rt").read().split()
    wl = Whitelist([b[:8] + LINKER_SLIDESEQ + b[8:] for b in bcs], 15, 24)
    r = wl.match_host(seqs, off, min_score=30, mode=NR_MODE_AUTO)
    assert compare(ref, r, 30, exact_below=True, label="slideseq") > 1000


def _expected_from_oracle(names, seqs, off, ref, wl_names, thr, umi_len, exact_len):
    """What the reference's process_matching_* computes, restated on the oracle arrays."""
    o = off.astype(np.int64)
    raw = seqs.tobytes()
    uniq = ref["n_best"] == 1
    fwd_scores = ref["best_score"][uniq & (ref["strand"] == 0)]
    triples = []
    for i in np.flatnonzero(uniq & (ref["strand"] == 0) & (ref["best_score"] >= thr)):
        u = int(ref["umi_q"][i])
        s = raw[o[i]:o[i + 1]].decode()
        umi = s[u:u + umi_len] if u >= 0 else "N"
        if (len(umi) != umi_len) if exact_len else (len(umi) < umi_len):
            continue
        triples.append((names[i], wl_names[ref["best_idx"][i]], umi))
    return fwd_scores, triples


def test_pipeline_tcr3_files(cuda_device, tmp_path):
    """config 1 (TCR3, 5' 10x TCR mode): write_bc_5p10X -> barcode_ref -> barcode_align ->
    process_matching_5p10XTCR, files compared with the oracle-derived expectation."""
    from nanoranger_b200 import utils, whitelists
    names, seqs, off, ref = _load("tcr3")
    out = str(tmp_path)
    wl_a = whitelists.load_737k()
    wl_names = whitelists.ascii_to_strings(wl_a)
    with open(f"{out}/wl.txt", "w") as f:
        f.write("\n".join(n + "-1" for n in wl_names) + "\n")
    os.link(os.path.join(G, "tcr3.fa.gz"), f"{out}/s_BCUMI.fasta.gz") if hasattr(os, "link") else None
    utils.write_bc_5p10X("s", out, f"{out}/wl.txt")
    utils.barcode_ref(f"{out}/s_bcreads.fasta", f"{out}/s_ref/")
    n = utils.barcode_align(f"{out}/s_BCUMI.fasta.gz", f"{out}/s_ref/", f"{out}/s_matching", 4, "-1")
    assert n == int((ref["n_best"] == 1).sum())
    utils.process_matching_5p10XTCR("s", out)
    fwd_scores, triples = _expected_from_oracle(names, seqs, off, ref, wl_names, 14, 12, False)
    got = pd.read_csv(f"{out}/s_barcode_scores.csv")
    v, c = np.unique(fwd_scores, return_counts=True)
    assert set(zip(got.score.tolist(), got["count"].tolist())) == set(zip(v.tolist(), c.tolist()))
    per = {}
    for nme, b, u in triples:
        per[nme] = (b, u)
    exp = {}
    for b, u in per.values():
        exp.setdefault(b, []).append(u)
    ded = pd.read_csv(f"{out}/s_bcumi_dedup.csv", index_col=0)
    assert len(ded) == len(exp)
    for b, us in exp.items():
        assert ded.loc[b, "umi_cnt"] == len(set(us)) and ded.loc[b, "read_cnt"] == len(us)
        assert abs(ded.loc[b, "dup_rate"] - len(us) / len(set(us))) < 1e-12
    assert (np.diff(ded.umi_cnt.values) <= 0).all()


def test_pipeline_5p10X_tags(cuda_device, tmp_path):
    """config 3 geometry (5p10XGEX: UMI 10 nt): name -> (CB, UB, XT) table."""
    from nanoranger_b200 import utils, whitelists
    names, seqs, off, ref = _load("mtdna1026")
    out = str(tmp_path)
    wl_a = whitelists.load_737k()
    wl_names = whitelists.ascii_to_strings(wl_a)
    np.savez_compressed(f"{out}/nr_whitelist.npz", cores=wl_a, names=np.array(wl_names),
                        pad_l=30, pad_r=40)
    utils.barcode_align(os.path.join(G, "mtdna1026.fa.gz"), out, f"{out}/s_matching", 8)
    # a genome BAM of the same reads (every third one secondary: must not be tagged), plus a stranger
    from nanoranger_b200 import bamio
    recs = [bamio.make_record(n, [0, 16, 256][i % 3], 0, 100 + i, 60, "20M", "ACGT" * 5)
            for i, n in enumerate(names)]
    recs.append(bamio.make_record("not_a_candidate", 0, 0, 5, 60, "20M", "ACGT" * 5))
    bamio.write_bam(f"{out}/s_genome.bam", "@HD\tVN:1.6\n@SQ\tSN:chrM\tLN:16569\n", [("chrM", 16569)], recs)
    table = utils.process_matching_5p10X("s", out)
    _, triples = _expected_from_oracle(names, seqs, off, ref, wl_names, 14, 10, False)
    exp = {n: (b, u, "lite") for n, b, u in triples}
    assert table == exp and len(table) > 3000
    # utils.py:801-827: tagged BAM holds the assigned reads with flag < 20, CB/UB/XT set
    tagged = list(bamio.BamReader(f"{out}/s_genome_tagged.bam"))
    want = [n for i, n in enumerate(names) if n in exp and [0, 16, 256][i % 3] < 20]
    assert [bamio.rec_qname(r) for r in tagged] == want
    for r in tagged[:200]:
        b, u, t = exp[bamio.rec_qname(r)]
        assert (bamio.get_tag(r, "CB"), bamio.get_tag(r, "UB"), bamio.get_tag(r, "XT")) == (b, u, t)
    ct = pd.read_csv(f"{out}/s_trns_ct.csv")
    assert ct.iloc[0, 0] == "lite" and int(ct.iloc[0, 1]) == len(want)


# ---- real alignments in the SAM (nr_sam_write_aligned) ---------------------------------------------

def _walk(pos, cigar, seq, ref):
    """pysam-style aligned pairs + score / mismatches / MD recomputed from POS, CIGAR, SEQ and the
    padded reference (the scoring of scripts/barcode_align.sh:18-33)."""
    import re
    r, q = pos - 1, 0
    pairs, score, nm, md, run, in_del = {}, 0, 0, "", 0, False
    for n, op in re.findall(r"(\d+)([MID])", cigar):
        n = int(n)
        if op == "M":
            for _ in range(n):
                a, b = seq[q], ref[r]
                pairs[r] = q
                if a == "N" or b == "N":
                    run += 1
                elif a == b:
                    score += 1
                    run += 1
                else:
                    score -= 1
                    nm += 1
                    md += f"{run}{b}"
                    run = 0
                r += 1
                q += 1
            in_del = False
        elif op == "I":
            score -= n
            q += n
        else:
            score -= n
            md += f"{run}^" + ref[r:r + n]
            run = 0
            r += n
    md += str(run)
    return pairs, score, nm, md, q, r


def _check_sam(path, names, ref, ref_strs, pad_l, L, pad_r):
    from nanoranger_b200 import samio
    by_name = {n: i for i, n in enumerate(names)}
    n_fwd = n_rev = n_gapped = 0
    with open(path) as f:
        for ln in f:
            if ln.startswith("@"):
                continue
            t = ln.rstrip("\n").split("\t")
            i = by_name[t[0]]
            assert ref["n_best"][i] == 1
            tags = t[11:]
            assert [x[:5] for x in tags] == ["AS:i:", "nM:i:", "MD:Z:"], tags      # barcode_align.sh:21
            a_s, n_m, md = int(tags[0][5:]), int(tags[1][5:]), tags[2][5:]
            flag, rname, pos, cigar, seq = int(t[1]), t[2], int(t[3]), t[5], t[9]
            core = ref_strs[ref["best_idx"][i]]
            padded = "N" * pad_l + core + "N" * pad_r
            pairs, score, nm, md2, qlen, rend = _walk(pos, cigar, seq, padded)
            assert qlen == len(seq) and rend <= len(padded)                 # EndToEnd, inside the chromosome
            assert score == a_s == int(ref["best_score"][i]), (t[0], cigar, score, a_s)
            assert nm == n_m and md2 == md, (t[0], cigar, md, md2)
            assert (flag == 16) == (ref["strand"][i] == 1)
            n_gapped += ("I" in cigar) or ("D" in cigar)
            if flag == 0:
                # utils.py:705-708: query index paired with reference column padL+L
                u = pairs.get(pad_l + L, -1)
                assert u == int(ref["umi_q"][i]), (t[0], cigar, u, ref["umi_q"][i])
                assert samio.query_index_at(pos, cigar, pad_l + L) == (None if u < 0 else u)
                n_fwd += 1
            else:
                n_rev += 1
    return n_fwd, n_rev, n_gapped


def test_sam_records_are_real_alignments_tcr3(cuda_device, tmp_path):
    from nanoranger_b200 import utils, whitelists
    names, seqs, off, ref = _load("tcr3")
    out = str(tmp_path)
    wl_a = whitelists.load_737k()
    wl_names = whitelists.ascii_to_strings(wl_a)
    np.savez_compressed(f"{out}/nr_whitelist.npz", cores=wl_a, names=np.array(wl_names), pad_l=30, pad_r=40)
    n = utils.barcode_align(os.path.join(G, "tcr3.fa.gz"), out, f"{out}/t_matching", 4)
    assert n == int((ref["n_best"] == 1).sum())
    n_fwd, n_rev, n_gapped = _check_sam(f"{out}/t_matching.sam", names, ref, wl_names, 30, 16, 40)
    assert n_fwd > 1500 and n_rev > 10 and n_gapped > 100
    # the anchored (fast) form stays available and yields the same UMI lookups
    n2 = utils.barcode_align(os.path.join(G, "tcr3.fa.gz"), out, f"{out}/a_matching", 4, alignments="anchored")
    assert n2 == n


def test_sam_records_are_real_alignments_slideseq(cuda_device, tmp_path):
    """32-column cores with N columns, pads 15/24 (utils.py:584-601)."""
    from nanoranger_b200 import utils
    from nanoranger_b200.whitelists import LINKER_SLIDESEQ
    names, seqs, off, ref = _load("slideseq")
    out = str(tmp_path)
    bcs = gzip.open(os.path.join(G, "slideseq_whitelist.txt.gz"), "rt").read().split()
    cores = [b[:8] + LINKER_SLIDESEQ + b[8:] for b in bcs]
    np.savez_compressed(f"{out}/nr_whitelist.npz",
                        cores=np.frombuffer("".join(cores).encode(), np.uint8).reshape(len(cores), 32),
                        names=np.array(bcs), pad_l=15, pad_r=24)
    n = utils.barcode_align(os.path.join(G, "slideseq.fa.gz"), out, f"{out}/s_matching", 4)
    assert n == int((ref["n_best"] == 1).sum())
    n_fwd, n_rev, n_gapped = _check_sam(f"{out}/s_matching.sam", names, ref, cores, 15, 32, 24)
    assert n_fwd > 500 and n_gapped > 50
