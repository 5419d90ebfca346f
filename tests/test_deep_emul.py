"""Exactness of the deep matcher tier's arithmetic, checked on the CPU.

tests/emul/deep_emul.cpp is compiled with g++ from the headers the sm_100a kernel includes
(nanoranger_b200/csrc/nr_deep_core.h, nr_deep_index.h): the plane automaton, the prefix/suffix
grouping and the join are the shipped code.  Compared bit-exactly with the oracle's exhaustive
scan: best score, smallest entry among the co-optimal pairs, their number and the strand, for
every candidate whose best cost is <= K -- and "not taken" exactly for the others.
"""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from helpers import mixed_candidates, mutate, rs, tie_rich_whitelist

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "..", "nanoranger_b200", "csrc")


@pytest.fixture(scope="module")
def emul():
    bd = os.path.join(HERE, "emul", "_build")
    os.makedirs(bd, exist_ok=True)
    so = os.path.join(bd, "libdeep_emul.so")
    deps = [os.path.join(HERE, "emul", "deep_emul.cpp"), os.path.join(CSRC, "nr_deep_core.h"),
            os.path.join(CSRC, "nr_deep_index.h")]
    if not os.path.exists(so) or os.path.getmtime(so) < max(os.path.getmtime(d) for d in deps):
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-Wall", "-Wno-unknown-pragmas", "-o", so, deps[0]])
    return C.CDLL(so)


def P(a, t):
    return a.ctypes.data_as(C.POINTER(t)) if a is not None else None


def pack_cores(wlc):
    n, L = wlc.shape
    lo = np.zeros(n, np.uint32)
    hi = np.zeros(n, np.uint32)
    nm = np.zeros(n, np.uint32)
    for j in range(L):
        c = wlc[:, j].astype(np.uint32)
        isn = c > 3
        c = np.where(isn, 0, c).astype(np.uint32)
        if j < 16:
            lo |= c << np.uint32(2 * j)
        else:
            hi |= c << np.uint32(2 * (j - 16))
        nm |= isn.astype(np.uint32) << np.uint32(j)
    return lo, hi, nm


def run_emul(E, O, wl, cands, pad_l, pad_r, K, force_s=0):
    L = len(wl[0])
    wlc, _ = O.encode_many(wl, L)
    cc, cl = O.encode_many(cands, 64)
    lo, hi, nm = pack_cores(wlc)
    n = len(cands)
    out = dict(idx=np.zeros(n, np.int32), score=np.zeros(n, np.int8), nbest=np.zeros(n, np.int32),
               strand=np.zeros(n, np.uint8), took=np.zeros(n, np.uint8))
    info = np.zeros(7, np.int32)
    rc = E.nr_emul_deep(P(lo, C.c_uint32), P(hi, C.c_uint32), P(nm, C.c_uint32), C.c_int64(len(wl)),
                        L, pad_l, pad_r, K, force_s, P(cc, C.c_uint8), P(cl.astype(np.uint8), C.c_uint8),
                        C.c_int64(n), P(out["idx"], C.c_int32), P(out["score"], C.c_int8),
                        P(out["nbest"], C.c_int32), P(out["strand"], C.c_uint8),
                        P(out["took"], C.c_uint8), P(info, C.c_int32))
    assert rc == 0
    ref = O.match(wlc, pad_l, pad_r, cc, cl)
    return ref, out, info, cl


def check(ref, out, cl, L, K, cands, wl):
    within = (L - ref["best_score"].astype(np.int64) <= K) & (cl >= 1) & (cl <= 63)
    ok = out["took"] == within.astype(np.uint8)
    ok &= np.where(within, (out["score"] == ref["best_score"]) & (out["idx"] == ref["best_idx"]) &
                   (out["nbest"] == ref["n_best"]) & (out["strand"] == ref["strand"]), True)
    bad = np.flatnonzero(~ok)
    assert len(bad) == 0, (len(bad), cands[bad[0]], wl[ref["best_idx"][bad[0]]],
                           {k: v[bad[0]] for k, v in ref.items()},
                           {k: v[bad[0]] for k, v in out.items()})
    return int(within.sum())


@pytest.mark.parametrize("K", [3, 5, 8])
def test_deep_equals_oracle_random_geometries(oracle, emul, K):
    rng = np.random.default_rng(10 + K)
    taken = 0
    for rnd in range(12):
        pad_l = int(rng.choice([30, 4, 16, 0, 2, 40]))
        pad_r = int(rng.choice([40, 17, 28, 0, 3]))
        wl = tie_rich_whitelist(rng, int(rng.integers(50, 400)))
        qlen = int(rng.choice([24, 35, 50, 56, 63]))
        cands = mixed_candidates(rng, wl, 150, pad_l, qlen, with_n=0.2)
        cands += [rs(rng, int(rng.integers(1, 20))) for _ in range(20)]       # very short reads
        cands += [mutate(rng, wl[0], 4) for _ in range(10)]
        ref, out, info, cl = run_emul(emul, oracle, wl, cands, pad_l, pad_r, K,
                                      force_s=int(rng.choice([0, 0, 3, 8, 13])))
        taken += check(ref, out, cl, 16, K, cands, wl)
    assert taken > 300


def test_deep_slideseq_cores_with_n_columns(oracle, emul):
    """32-column cores (8 + linker 18 + 6) with N inside some entries (utils.py:584-601)."""
    rng = np.random.default_rng(5)
    linker = "TCTTCAGCGTTCCCGAGA"
    bcs = set()
    while len(bcs) < 300:
        b = list(rs(rng, 14))
        if rng.random() < 0.2:
            b[int(rng.integers(0, 14))] = "N"
        bcs.add("".join(b))
    bcs = sorted(bcs)
    wl = [b[:8] + linker + b[8:] for b in bcs]
    cands = []
    for _ in range(300):
        core = wl[int(rng.integers(0, len(wl)))].replace("N", "ACGT"[int(rng.integers(0, 4))])
        mid = mutate(rng, core, int(rng.choice([0, 1, 2, 3, 4])))
        q = rs(rng, int(rng.integers(0, 18))) + mid + rs(rng, int(rng.integers(0, 14)))
        q = q[:63]
        if rng.random() < 0.1:
            j = int(rng.integers(0, len(q)))
            q = q[:j] + "N" + q[j + 1:]
        if rng.random() < 0.15:
            q = oracle.revcomp(q)
        cands.append(q)
    for K in (3, 8):
        ref, out, info, cl = run_emul(emul, oracle, wl, cands, 15, 24, K)
        assert check(ref, out, cl, 32, K, cands, wl) > 100
        assert 1 <= info[0] < 32


def test_deep_real_737k_sample(oracle, emul):
    """A slice of the real list (sorted: many shared prefixes) and ONT-profile candidates."""
    from nanoranger_b200 import synth, whitelists
    wl_a = whitelists.load_737k()[::37][:6000]
    wl = ["".join(chr(c) for c in row) for row in wl_a]
    d = synth.make_candidates(wl_a, 400, seed=4, frac_negative=0.5)
    cands = synth.to_strings(d["seqs"], d["offsets"])
    ref, out, info, cl = run_emul(emul, oracle, wl, cands, 30, 40, 5)
    assert check(ref, out, cl, 16, 5, cands, wl) > 350
    assert info[1] < len(wl) and info[2] < len(wl)
    assert 0 < info[3] < info[0] and 0 < info[5] < info[1], info      # prefix side shares columns


def test_umi_row_by_planes_equals_oracle_pair(oracle, emul):
    """the deep tier's finaliser derives the UMI column from the plane automaton; the oracle's
    pair DP (tier 1) is the definition."""
    O = oracle
    rng = np.random.default_rng(21)
    n_checked = n_none = 0
    for _ in range(6000):
        L = int(rng.choice([16, 16, 32, 12]))
        pad_l = int(rng.choice([30, 4, 16, 0, 2, 15]))
        pad_r = int(rng.choice([40, 17, 28, 0, 3, 24]))
        core = list(rs(rng, L))
        if rng.random() < 0.2:
            core[int(rng.integers(0, L))] = "N"
        core = "".join(core)
        mid = mutate(rng, core.replace("N", "A"), int(rng.integers(0, 5)))
        pre, suf = rs(rng, int(rng.integers(0, 36))), rs(rng, int(rng.integers(0, 30)))
        mode = rng.integers(0, 6)
        q = (mid[int(rng.integers(1, 4)):] + suf) if mode == 0 else \
            (pre + mid[:-int(rng.integers(1, 4))]) if mode == 1 else (pre + mid + suf)
        q = list(q[:63])
        if not q:
            continue
        if rng.random() < 0.15:
            q[int(rng.integers(0, len(q)))] = "N"
        q = "".join(q)
        a1, u1 = O.pair(q, core, pad_l, pad_r)
        c = L - a1
        if c > 8:
            continue
        wlc, _ = O.encode_many([core], L)
        lo, hi, nm = pack_cores(wlc)
        qc = np.ascontiguousarray(O.encode(q))
        u = emul.nr_emul_deep_umi(P(qc, C.c_uint8), len(qc), int(lo[0]), int(hi[0]), int(nm[0]), L,
                                  pad_l, pad_r, int(c))
        assert u == u1, (q, core, pad_l, pad_r, a1, u1, u)
        n_checked += 1
        n_none += u1 < 0
    assert n_checked > 2500 and n_none > 50
