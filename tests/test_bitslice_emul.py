"""Exactness of the bit-parallel brute-force DP's arithmetic, checked on the CPU.

tests/emul/bitslice_emul.cpp is compiled with g++ from the header the sm_100a kernel includes
(nanoranger_b200/csrc/nr_bitslice_core.h): the delta-encoded cell, the plane adders, the transpose
and the lane minimum are the shipped code.  Compared with the oracle on EVERY (entry, strand)
pair's AS (not only the best one), then on best score / tie count / smallest entry / strand.
"""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from helpers import mixed_candidates, rs, tie_rich_whitelist
from test_deep_emul import P, pack_cores

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "..", "nanoranger_b200", "csrc")


@pytest.fixture(scope="module")
def emul():
    bd = os.path.join(HERE, "emul", "_build")
    os.makedirs(bd, exist_ok=True)
    so = os.path.join(bd, "libbitslice_emul.so")
    deps = [os.path.join(HERE, "emul", "bitslice_emul.cpp"), os.path.join(CSRC, "nr_bitslice_core.h")]
    if not os.path.exists(so) or os.path.getmtime(so) < max(os.path.getmtime(d) for d in deps):
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-Wall",
                               "-Wno-unknown-pragmas", "-o", so, deps[0]])
    return C.CDLL(so)


def test_cell_adders_transpose(emul):
    """every (a, b, s) input of the cell against d = min(s, a, b), a' = d - b + 3, b' = d - a + 3;
    the sign-extended adders, the per-lane minimum and the 32 x 32 transpose on random words"""
    assert emul.nr_emul_bitslice_cell_check() == 0


def run_all_pairs(E, O, wl, cands, pad_l, pad_r):
    L = len(wl[0])
    wlc, _ = O.encode_many(wl, L)
    lo, hi, nm = pack_cores(wlc)
    has_n = int(nm.any())
    n = len(wl)
    cc, cl = O.encode_many(cands, 64)
    ref = O.match(wlc, pad_l, pad_r, cc, cl)
    for k, q in enumerate(cands):
        m = len(q)
        qf = np.ascontiguousarray(cc[k, :m])
        qr = np.ascontiguousarray(np.where(qf[::-1] > 3, 4, 3 - qf[::-1]).astype(np.uint8))
        as_f = np.zeros(n, np.int8)
        as_r = np.zeros(n, np.int8)
        best = np.zeros(4, np.int32)
        rc = E.nr_emul_bitslice(P(lo, C.c_uint32), P(hi, C.c_uint32), P(nm, C.c_uint32), C.c_int64(n), L,
                                has_n, pad_l, pad_r, P(qf, C.c_uint8), P(qr, C.c_uint8), m,
                                P(as_f, C.c_int8), P(as_r, C.c_int8), P(best, C.c_int32))
        assert rc == 0
        ef = O.scores(wlc, pad_l, pad_r, q)
        er = O.scores(wlc, pad_l, pad_r, O.revcomp(q))
        bad = np.flatnonzero((as_f != ef) | (as_r != er))
        assert len(bad) == 0, (q, wl[bad[0]], int(as_f[bad[0]]), int(ef[bad[0]]), int(as_r[bad[0]]), int(er[bad[0]]))
        assert (best[0], best[1], best[2], best[3]) == (ref["best_score"][k], ref["n_best"][k],
                                                       ref["best_idx"][k], ref["strand"][k]), (q, best)


@pytest.mark.parametrize("pads", [(30, 40), (4, 17), (0, 0), (16, 28)])
def test_bitslice_16_all_pairs(emul, pads):
    from oracle import oracle as O
    rng = np.random.default_rng(100 + pads[0])
    wl = tie_rich_whitelist(rng, 300)[:293]          # a ragged last word
    cands = mixed_candidates(rng, wl, 60, pads[0], 50, with_n=0.3)
    cands += [rs(rng, int(k)) for k in (1, 2, 3, 15, 16, 17, 63, 64)]
    cands += [wl[5], wl[7][3:], wl[9][:11], "N" * 20, "A" * 64]
    run_all_pairs(emul, O, wl, cands, *pads)


def test_bitslice_16_entries_with_n(emul):
    from oracle import oracle as O
    rng = np.random.default_rng(7)
    wl = []
    for _ in range(130):
        s = list(rs(rng, 16))
        for _ in range(int(rng.integers(0, 3))):
            s[int(rng.integers(0, 16))] = "N"
        wl.append("".join(s))
    wl = sorted(set(wl))
    cands = mixed_candidates(rng, [w.replace("N", "A") for w in wl], 40, 30, 48, with_n=0.3)
    run_all_pairs(emul, O, wl, cands, 30, 40)


@pytest.mark.parametrize("with_n", [False, True])
def test_bitslice_32_slideseq_geometry(emul, with_n):
    """8 + linker + 6 cores (reference utils.py:584-601), pads 15 / 24, with and without N columns"""
    from oracle import oracle as O
    from nanoranger_b200.whitelists import LINKER_SLIDESEQ
    rng = np.random.default_rng(11 + with_n)
    wl = []
    for _ in range(150):
        b = list(rs(rng, 14))
        if with_n and rng.random() < 0.2:
            b[int(rng.integers(0, 14))] = "N"
        b = "".join(b)
        wl.append(b[:8] + LINKER_SLIDESEQ + b[8:])
    wl = sorted(set(wl))
    cands = mixed_candidates(rng, [w.replace("N", "C") for w in wl], 40, 15, 60, with_n=0.2)
    cands += [rs(rng, int(k)) for k in (1, 5, 31, 32, 33, 64)]
    run_all_pairs(emul, O, wl, cands, 15, 24)


def test_bitslice_golden_slideseq_fixture(emul):
    """the reference's own slide-seq list (17 753 x 32 columns, 2 584 entries with N) and reads cut
    from its sample FASTQ (tests/golden, make_golden.py): every pair's AS equals the oracle's and
    best / ties / smallest entry / strand equal the committed golden answers"""
    import gzip
    from oracle import oracle as O
    from nanoranger_b200.whitelists import LINKER_SLIDESEQ
    G = os.path.join(HERE, "golden")
    bcs = gzip.open(os.path.join(G, "slideseq_whitelist.txt.gz"), "rt").read().split()
    wl = [b[:8] + LINKER_SLIDESEQ + b[8:] for b in bcs]
    seqs = [ln.strip() for ln in gzip.open(os.path.join(G, "slideseq.fa.gz"), "rt") if not ln.startswith(">")]
    gold = np.load(os.path.join(G, "slideseq.oracle.npz"))
    pick = list(range(0, len(seqs), max(1, len(seqs) // 12)))[:12]
    run_all_pairs(emul, O, wl, [seqs[i] for i in pick], int(gold["pad_l"]), int(gold["pad_r"]))
    # run_all_pairs compared with the oracle computed now; the committed answers agree with it
    wlc, _ = O.encode_many(wl, 32)
    cc, cl = O.encode_many([seqs[i] for i in pick], 64)
    ref = O.match(wlc, int(gold["pad_l"]), int(gold["pad_r"]), cc, cl)
    for k in ("best_idx", "best_score", "n_best", "strand"):
        assert np.array_equal(ref[k], gold[k][pick]), k
