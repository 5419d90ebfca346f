"""The oracle against itself and against hand-written known answers.

The reference ships no golden vectors for this path (its aligner, STAR, is external and absent),
so the oracle is pinned by: (1) the literal Appendix-C DP in C (tier 0) == its numpy twin,
(2) tier 1 (core columns + closed-form pads) == tier 0, (3) tier 2 (exhaustive SIMD scan) ==
tier 1 on every pair, (4) the known-answer vectors below.
"""
import numpy as np
import pytest

from helpers import mutate, rs


def test_known_answers(oracle):
    O = oracle
    bc = "ACGTTGCATCGATTGA"
    ref = "N" * 30 + bc + "N" * 40
    pre, umi, tso = "CGCTCTTCCGATCT", "AACCGGTTAACC", "TTTCTTATAT"
    q = pre + bc + umi + tso
    assert O.as_padded(q, ref) == 16                                  # exact barcode
    assert O.pair(q, bc, 30, 40) == (16, len(pre) + 16)               # UMI starts right after it
    sub = pre + bc[:5] + "A" + bc[6:] + umi + tso
    assert bc[5] != "A" and O.as_padded(sub, ref) == 14               # one substitution
    ins = pre + bc[:7] + "G" + bc[7:] + umi + tso
    assert O.as_padded(ins, ref) == 15                                # one extra read base
    assert O.pair(ins, bc, 30, 40) == (15, len(pre) + 17)
    dele = pre + bc[:9] + bc[10:] + umi + tso
    assert O.as_padded(dele, ref) == 14                               # one missing base
    ins2 = pre + bc[:3] + "T" + bc[3:11] + "C" + bc[11:] + umi + tso
    assert O.as_padded(ins2, ref) == 14                               # two extra bases
    # barcode starting at query offset 31 with left pad 30: one forced insertion
    q31 = rs(np.random.default_rng(0), 31) + bc + umi[:3]
    assert O.as_padded(q31, ref) == 15
    # core hanging over the read start by two columns costs 1 per column
    assert O.as_padded(bc[2:] + umi + tso, ref) == 14
    # N in the candidate scores 0 at that column
    qn = pre + bc[:4] + "N" + bc[5:] + umi + tso
    assert O.as_padded(qn, ref) == 15
    # whitelist entry with an internal N (slide-seq): that column scores 0
    core = "ACGTTGCA" + "TCTTCAGCGTTCCCGAGA" + "TCNATT"
    rd = "GG" + core.replace("N", "G") + "ACGTACGTA"
    assert O.as_padded(rd, "N" * 15 + core + "N" * 24) == 31


def test_tie_between_hamming2_neighbours(oracle):
    O = oracle
    a = "ACGTTGCATCGATTGA"
    b = "ACGTTGCTTCGATTGC"        # Hamming 2 from a
    mid = "ACGTTGCATCGATTGC"      # Hamming 1 from both
    wl, _ = O.encode_many([a, b, "TTTTGGGGCCCCAAAA"], 16)
    cc, cl = O.encode_many(["CGCTCTTCCGATCT" + mid + "GGTTAACCGGTTTTTCTTATAT"], 64)
    r = O.match(wl, 30, 40, cc, cl)
    assert r["best_score"][0] == 14 and r["n_best"][0] == 2 and r["best_idx"][0] == 0


def test_reverse_strand_is_flagged(oracle):
    O = oracle
    a = "ACGTTGCATCGATTGA"
    wl, _ = O.encode_many([a, "TTTTGGGGCCCCAAAA"], 16)
    q = O.revcomp("CGCTCTTCCGATCT" + a + "AACCGGTTAACCTTTCTTATAT")
    cc, cl = O.encode_many([q], 64)
    r = O.match(wl, 30, 40, cc, cl)
    assert r["best_score"][0] == 16 and r["strand"][0] == 1 and r["umi_q"][0] == -1


def test_tier0_c_equals_numpy_twin(oracle):
    O = oracle
    rng = np.random.default_rng(11)
    for _ in range(60):
        core = rs(rng, int(rng.integers(8, 33)))
        if rng.random() < 0.3:
            j = int(rng.integers(0, len(core)))
            core = core[:j] + "N" + core[j + 1:]
        ref = "N" * int(rng.integers(0, 31)) + core + "N" * int(rng.integers(0, 41))
        q = rs(rng, int(rng.integers(0, 10))) + mutate(rng, core.replace("N", "A"), int(rng.integers(0, 4))) + \
            rs(rng, int(rng.integers(0, 20)))
        q = q[:64] or "A"
        assert O.as_padded(q, ref) == O.as_padded_numpy(q, ref)


def test_tier1_equals_tier0(oracle):
    O = oracle
    rng = np.random.default_rng(12)
    for it in range(4000):
        pad_l = int(rng.choice([30, 4, 16, 15, 0, 2]))
        pad_r = int(rng.choice([40, 17, 28, 24, 0, 3]))
        L = int(rng.choice([16, 16, 32, 20]))
        core = rs(rng, L)
        if rng.random() < 0.2:
            j = int(rng.integers(0, L))
            core = core[:j] + "N" + core[j + 1:]
        mid = mutate(rng, core.replace("N", "C"), int(rng.integers(0, 4)))
        mode = rng.integers(0, 6)
        pre, suf = rs(rng, int(rng.integers(0, 36))), rs(rng, int(rng.integers(0, 30)))
        q = (mid[int(rng.integers(1, 4)):] + suf) if mode == 0 else \
            (pre + mid[:-int(rng.integers(1, 4))]) if mode == 1 else (pre + mid + suf)
        q = q[:64] or "G"
        a1, u1 = O.pair(q, core, pad_l, pad_r)
        assert a1 == O.as_padded(q, "N" * pad_l + core + "N" * pad_r), (q, core, pad_l, pad_r)
        assert -1 <= u1 <= len(q)


def test_tier2_equals_tier1(oracle):
    O = oracle
    rng = np.random.default_rng(13)
    for (L, pad_l, pad_r) in [(16, 30, 40), (32, 15, 24), (16, 4, 17)]:
        wl = sorted({rs(rng, L) for _ in range(300)})
        if L == 32:
            wl = [w if rng.random() > 0.2 else w[:5] + "N" + w[6:] for w in wl]
        qs = []
        for _ in range(60):
            c = wl[rng.integers(0, len(wl))].replace("N", "T")
            qs.append((rs(rng, int(rng.integers(0, 20))) + mutate(rng, c, int(rng.integers(0, 3))) +
                       rs(rng, 20))[:int(rng.integers(20, 65))])
        qs += ["", "A", "N" * 10]
        wlc, _ = O.encode_many(wl, L)
        cc, cl = O.encode_many(qs, 64)
        r = O.match(wlc, pad_l, pad_r, cc, cl, threads=3)
        for i, q in enumerate(qs):
            best, cnt, arg = -999, 0, None
            for e, w in enumerate(wl):
                for s, qq in enumerate((q, O.revcomp(q))):
                    a, _ = O.pair(qq, w, pad_l, pad_r) if qq else (0, 0)
                    if a > best:
                        best, cnt, arg = a, 1, (e, s)
                    elif a == best:
                        cnt += 1
            assert r["best_score"][i] == best and r["n_best"][i] == cnt
            assert (r["best_idx"][i], r["strand"][i]) == arg
            if arg[1] == 0 and q:
                assert r["umi_q"][i] == O.pair(q, wl[arg[0]], pad_l, pad_r)[1]


def test_umi_cluster_oracle_basics(oracle):
    O = oracle
    bc = np.array([1, 1, 1, 1, 1, 2, 2], np.uint32)
    gene = np.zeros(7, np.uint32)
    u = 0b1101_1000
    umi = np.array([u, u, u, u ^ 1, 0xFFFF, u, u ^ 1], np.uint32)
    k0, rep0 = O.umi_cluster(bc, gene, umi, 0)
    assert k0 == 5 and np.array_equal(rep0, umi)                # exact dedup: np.unique per barcode
    k1, rep1 = O.umi_cluster(bc, gene, umi, 1)
    assert k1 == 3 and rep1[3] == u                             # 1-read neighbour joins the 3-read UMI
    assert rep1[5] == u and rep1[6] == u                        # 1 vs 1: 1 >= 2*1-1, joins the smaller UMI
    assert rep1[4] == 0xFFFF


def test_cpu_filter_port_equals_exhaustive_oracle(oracle):
    """oracle/nr_filter_cpu.cpp (the GPU path's algorithm on the host, bench baseline
    "port-filtered") against the exhaustive scorer on the real 737K list."""
    import time
    from nanoranger_b200 import synth, whitelists
    O = oracle
    wl_a = whitelists.load_737k()
    d = synth.make_candidates(wl_a, 1500, seed=9, p_n=0.004)
    seqs = synth.to_strings(d["seqs"], d["offsets"])
    cc, cl = O.encode_many(seqs, 64)
    wlc = O._CODE[wl_a]
    ref = O.match(wlc, 30, 40, cc, cl)
    f = O.FilterCPU(wlc, 30, 40)
    t = time.perf_counter()
    out = f.match(cc, cl)
    dt = time.perf_counter() - t
    hi = (ref["best_score"] >= 14) & (out["took"] == 1)
    assert hi.sum() > 1000
    for k in ("best_idx", "best_score", "n_best", "strand", "umi_q"):
        assert np.array_equal(out[k][hi], ref[k][hi]), k
    lo = (ref["best_score"] < 14) & (out["took"] == 1)
    assert (out["best_score"][lo] == -128).all()
    assert dt < 30
