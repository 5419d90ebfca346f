"""Losslessness of the anchored seed filter (slide-seq cores: 8 + linker 18 + 6 columns) checked
on the CPU: tests/emul/anchor_emul.cpp is compiled with g++ from the headers the sm_100a kernel
includes (nr_anchor_core.h, nr_anchor_index.h).  Compared bit-exactly with the oracle: score,
entry, tie count, strand, UMI column for every candidate whose best score is >= 30 (cost <= 2),
and "nothing found" for all others."""
import ctypes as C
import gzip
import os
import subprocess

import numpy as np
import pytest

from helpers import mutate, rs
from test_deep_emul import P, pack_cores

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "..", "nanoranger_b200", "csrc")
LINKER = "TCTTCAGCGTTCCCGAGA"


@pytest.fixture(scope="module")
def emul():
    bd = os.path.join(HERE, "emul", "_build")
    os.makedirs(bd, exist_ok=True)
    so = os.path.join(bd, "libanchor_emul.so")
    deps = [os.path.join(HERE, "emul", "anchor_emul.cpp")] + [os.path.join(CSRC, f) for f in
            ("nr_anchor_core.h", "nr_anchor_index.h", "nr_deep_core.h")]
    if not os.path.exists(so) or os.path.getmtime(so) < max(os.path.getmtime(d) for d in deps):
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-Wall", "-Wno-unknown-pragmas",
                               "-o", so, deps[0]])
    return C.CDLL(so)


def run_emul(E, O, wl, cands, pad_l, pad_r):
    L = len(wl[0])
    wlc, _ = O.encode_many(wl, L)
    cc, cl = O.encode_many(cands, 64)
    lo, hi, nm = pack_cores(wlc)
    n = len(cands)
    out = dict(idx=np.zeros(n, np.int32), score=np.zeros(n, np.int8), nbest=np.zeros(n, np.int32),
               strand=np.zeros(n, np.uint8), umi=np.zeros(n, np.int16), took=np.zeros(n, np.uint8))
    info = np.zeros(4, np.int32)
    cnt = np.zeros(2, np.int64)
    rc = E.nr_emul_anchored(P(lo, C.c_uint32), P(hi, C.c_uint32), P(nm, C.c_uint32), C.c_int64(len(wl)), L,
                            pad_l, pad_r, P(cc, C.c_uint8), P(cl.astype(np.uint8), C.c_uint8), C.c_int64(n),
                            P(out["idx"], C.c_int32), P(out["score"], C.c_int8), P(out["nbest"], C.c_int32),
                            P(out["strand"], C.c_uint8), P(out["umi"], C.c_int16), P(out["took"], C.c_uint8),
                            P(info, C.c_int32), P(cnt, C.c_int64))
    assert rc == 0
    ref = O.match(wlc, pad_l, pad_r, cc, cl)
    return ref, out, info, cnt


def check(ref, out, L, cands, wl):
    hi = ref["best_score"] >= L - 2
    exp = np.where(hi, ref["best_score"], -128)
    ok = (out["score"] == exp) & np.where(
        hi, (out["idx"] == ref["best_idx"]) & (out["nbest"] == ref["n_best"]) &
        (out["strand"] == ref["strand"]) & (out["umi"] == ref["umi_q"]), True)
    ok |= out["took"] == 0
    bad = np.flatnonzero(~ok)
    assert len(bad) == 0, (len(bad), cands[bad[0]], wl[ref["best_idx"][bad[0]]],
                           {k: v[bad[0]] for k, v in ref.items()}, {k: v[bad[0]] for k, v in out.items()})
    return int((hi & (out["took"] == 1)).sum())


def slide_whitelist(rng, n, p_n=0.15, near=0.3):
    bcs = set()
    while len(bcs) < n:
        b = list(rs(rng, 14))
        if rng.random() < p_n:
            b[int(rng.integers(0, 14))] = "N"
        bcs.add("".join(b))
        if rng.random() < near:                          # a neighbour at Hamming distance 1 or 2: ties
            b2 = list(b)
            for j in rng.choice(14, int(rng.integers(1, 3)), replace=False):
                b2[int(j)] = "ACGT"[int(rng.integers(0, 4))]
            bcs.add("".join(b2))
    bcs = sorted(bcs)
    return [b[:8] + LINKER + b[8:] for b in bcs]


def slide_candidates(rng, O, wl, n, with_n=0.0):
    out = []
    for _ in range(n):
        core = wl[int(rng.integers(0, len(wl)))]
        core = "".join(c if c != "N" else "ACGT"[int(rng.integers(0, 4))] for c in core)
        mid = mutate(rng, core, int(rng.choice([0, 0, 1, 1, 2, 2, 3])))
        mode = int(rng.integers(0, 8))
        if mode == 0:
            q = mid[int(rng.integers(1, 3)):] + rs(rng, int(rng.integers(5, 25)))      # core over the read start
        elif mode == 1:
            q = rs(rng, int(rng.integers(8, 18))) + mid[:len(mid) - int(rng.integers(1, 3))]   # over the read end
        elif mode == 2:
            q = rs(rng, int(rng.integers(40, 60)))
        else:
            q = rs(rng, int(rng.integers(0, 20))) + mid + rs(rng, int(rng.integers(0, 28)))
        q = q[:63]
        if rng.random() < 0.12:
            q = O.revcomp(q)
        if with_n and rng.random() < with_n and len(q) > 2:
            q = list(q)
            for _k in range(int(rng.choice([1, 1, 2, 3]))):
                q[int(rng.integers(0, len(q)))] = "N"
            q = "".join(q)
        out.append(q)
    return out


@pytest.mark.parametrize("pad_l,pad_r", [(15, 24), (0, 0), (3, 2), (30, 40)])
def test_anchored_filter_lossless_random(oracle, emul, pad_l, pad_r):
    rng = np.random.default_rng(40 + pad_l)
    wl = slide_whitelist(rng, 1500)
    cands = slide_candidates(rng, oracle, wl, 2500)
    ref, out, info, cnt = run_emul(emul, oracle, wl, cands, pad_l, pad_r)
    assert tuple(info[:3]) == (8, 18, 6)
    # (with tiny pads the flanks of the read cost 1 per base: few candidates reach AS >= 30)
    assert check(ref, out, 32, cands, wl) > (600 if pad_l >= 15 else 0)


def test_anchored_filter_reads_with_n(oracle, emul):
    """one or two N anywhere in the read (inside the 8 columns in front of the linker, inside the
    linker, in the tail): substituted windows + wildcard linker walk stay lossless."""
    rng = np.random.default_rng(77)
    wl = slide_whitelist(rng, 1500)
    cands = slide_candidates(rng, oracle, wl, 3000, with_n=1.0)
    ref, out, info, cnt = run_emul(emul, oracle, wl, cands, 15, 24)
    n_n = np.array([c.count("N") for c in cands])
    assert (out["took"][n_n <= 2] == 1).all() and (out["took"][n_n > 2] == 0).all()
    assert check(ref, out, 32, cands, wl) > 400


def _variants(rng, core):
    out = {core}
    L = len(core)
    for i in range(L):
        for b in "ACGT":
            if b != core[i]:
                out.add(core[:i] + b + core[i + 1:])
        out.add(core[:i] + core[i + 1:])
    for i in range(L + 1):
        for b in "ACGT":
            out.add(core[:i] + b + core[i:])
    for i in range(L + 1):
        for j in range(i, L + 1):
            b1, b2 = "ACGT"[rng.integers(0, 4)], "ACGT"[rng.integers(0, 4)]
            out.add(core[:i] + b1 + core[i:j] + b2 + core[j:])
    return sorted(out)


def test_anchored_filter_all_small_cost_variants(oracle, emul):
    """every <= 1-edit and two-insertion variant of a few cores, interior and flush with either
    read end (overhangs): all placements of cost <= 2 must be found."""
    rng = np.random.default_rng(6)
    wl = slide_whitelist(rng, 600, p_n=0.1)
    n_hi = 0
    for rep in range(3):
        core = wl[int(rng.integers(0, len(wl)))].replace("N", "ACGT"[rep])
        cands = []
        for v in _variants(rng, core):
            for where in range(4):
                pre = rs(rng, [9, 0, 14, 0][where])
                suf = rs(rng, [12, 20, 0, 0][where])
                for cut in ((0, 0), (1, 0), (2, 0), (0, 1), (0, 2), (1, 1)) if where else ((0, 0),):
                    q = pre + v + suf
                    if where == 1:
                        q = q[cut[0]:]                    # core hangs over the read start
                    if where == 2:
                        q = q[:len(q) - cut[1]]           # ... over the read end
                    if where == 3:
                        q = q[cut[0]:len(q) - cut[1]]
                    cands.append(q[:63])
        ref, out, info, cnt = run_emul(emul, oracle, wl, cands, 15, 24)
        n_hi += check(ref, out, 32, cands, wl)
    assert n_hi > 1500


def test_anchored_filter_real_slideseq_fixture(oracle, emul):
    """the reference's slide-seq list and the candidates cut from its sample FASTQ."""
    G = os.path.join(HERE, "golden")
    bcs = gzip.open(os.path.join(G, "slideseq_whitelist.txt.gz"), "rt").read().split()
    wl = [b[:8] + LINKER + b[8:] for b in bcs]
    seqs = [ln.strip() for ln in gzip.open(os.path.join(G, "slideseq.fa.gz"), "rt") if not ln.startswith(">")]
    ref, out, info, cnt = run_emul(emul, oracle, wl, seqs, 15, 24)
    assert check(ref, out, 32, seqs, wl) > 700
    stored = dict(np.load(os.path.join(G, "slideseq.oracle.npz")))
    assert np.array_equal(stored["best_score"], ref["best_score"])
