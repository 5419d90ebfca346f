"""Losslessness of the filtered matcher's arithmetic, checked on the CPU.

tests/emul/filter_emul.cpp is compiled with g++ from the very header the sm_100a kernel includes
(nanoranger_b200/csrc/nr_filter_core.h): the probe table, the key extraction, the index layout
and the cost<=2 automaton are therefore the shipped code; only the warp choreography differs.
Compared bit-exactly with the oracle: score, entry, tie count, strand, UMI column for every
candidate whose best score is >= 14, and "nothing found" for all others.
"""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from helpers import mixed_candidates, mutate, rs, tie_rich_whitelist

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope="module")
def emul():
    bd = os.path.join(HERE, "emul", "_build")
    os.makedirs(bd, exist_ok=True)
    so = os.path.join(bd, "libfilter_emul.so")
    src = os.path.join(HERE, "emul", "filter_emul.cpp")
    hdr = os.path.join(HERE, "..", "nanoranger_b200", "csrc", "nr_filter_core.h")
    if not os.path.exists(so) or os.path.getmtime(so) < max(os.path.getmtime(src), os.path.getmtime(hdr)):
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-Wall", "-o", so, src])
    return C.CDLL(so)


def P(a, t):
    return a.ctypes.data_as(C.POINTER(t))


def run_emul(E, O, wl, cands, pad_l, pad_r, windowed=1):
    wlc, _ = O.encode_many(wl, 16)
    cc, cl = O.encode_many(cands, 64)
    lo = np.zeros(len(wl), np.uint32)
    for j in range(16):
        lo |= wlc[:, j].astype(np.uint32) << np.uint32(2 * j)
    n = len(cands)
    out = dict(idx=np.zeros(n, np.int32), score=np.zeros(n, np.int8), nbest=np.zeros(n, np.int32),
               strand=np.zeros(n, np.uint8), umi=np.zeros(n, np.int16), took=np.zeros(n, np.uint8))
    cnt = np.zeros(3, np.int64)
    clu = cl.astype(np.uint8)
    rc = E.nr_emul_filtered(P(lo, C.c_uint32), C.c_int64(len(wl)), pad_l, pad_r, P(cc, C.c_uint8),
                            P(clu, C.c_uint8), C.c_int64(n), windowed, 24, P(out["idx"], C.c_int32),
                            P(out["score"], C.c_int8), P(out["nbest"], C.c_int32),
                            P(out["strand"], C.c_uint8), P(out["umi"], C.c_int16),
                            P(out["took"], C.c_uint8), P(cnt, C.c_int64))
    assert rc == 0
    ref = O.match(wlc, pad_l, pad_r, cc, cl)
    return ref, out, cnt


def check(ref, out, cands, wl):
    hi = ref["best_score"] >= 14
    exp = np.where(hi, ref["best_score"], -128)
    ok = (out["score"] == exp) & np.where(
        hi, (out["idx"] == ref["best_idx"]) & (out["nbest"] == ref["n_best"]) &
        (out["strand"] == ref["strand"]) & (out["umi"] == ref["umi_q"]), True)
    ok |= out["took"] == 0
    bad = np.flatnonzero(~ok)
    assert len(bad) == 0, (cands[bad[0]], wl[ref["best_idx"][bad[0]]],
                           {k: v[bad[0]] for k, v in ref.items()},
                           {k: v[bad[0]] for k, v in out.items()})
    return int((hi & (out["took"] == 1)).sum())


def test_automaton_equals_oracle_pair(oracle, emul):
    O = oracle
    rng = np.random.default_rng(1)
    n_low = 0
    for _ in range(6000):
        pad_l = int(rng.choice([30, 4, 16, 0, 2]))
        pad_r = int(rng.choice([40, 17, 28, 0, 3]))
        core = rs(rng, 16)
        mid = mutate(rng, core, int(rng.integers(0, 4)))
        pre, suf = rs(rng, int(rng.integers(0, 36))), rs(rng, int(rng.integers(0, 30)))
        mode = rng.integers(0, 6)
        q = (mid[int(rng.integers(1, 4)):] + suf) if mode == 0 else \
            (pre + mid[:-int(rng.integers(1, 4))]) if mode == 1 else (pre + mid + suf)
        q = q[:64]
        if not q:
            continue
        a1, u1 = O.pair(q, core, pad_l, pad_r)
        qc, cc = np.ascontiguousarray(O.encode(q)), np.ascontiguousarray(O.encode(core))
        u = C.c_int(-9)
        cost = emul.nr_emul_nfa(P(qc, C.c_uint8), len(qc), P(cc, C.c_uint8), pad_l, pad_r, C.byref(u))
        if 16 - a1 <= 2:
            n_low += 1
            assert (cost, u.value) == (16 - a1, u1), (q, core, pad_l, pad_r)
        else:
            assert cost == 3, (q, core, pad_l, pad_r, a1)
    assert n_low > 600


def test_revcomp_words(emul):
    rng = np.random.default_rng(2)
    comp = np.array([3, 2, 1, 0], np.uint8)
    for m in list(range(0, 65, 7)) + [1, 15, 16, 17, 31, 32, 33, 48, 63, 64]:
        q = rng.integers(0, 4, 64).astype(np.uint8)
        q[m:] = 0
        out = np.zeros(64, np.uint8)
        emul.nr_emul_revcomp(P(q, C.c_uint8), m, P(out, C.c_uint8))
        exp = np.zeros(64, np.uint8)
        exp[:m] = comp[q[:m][::-1]]
        assert np.array_equal(out, exp), m


@pytest.mark.parametrize("pad_l,pad_r,qlen", [(30, 40, 50), (4, 17, 35), (16, 28, 41), (30, 40, 64)])
def test_filter_lossless_random(oracle, emul, pad_l, pad_r, qlen):
    rng = np.random.default_rng(300 + qlen)
    wl = tie_rich_whitelist(rng, 2500)
    cands = mixed_candidates(rng, wl, 2500, pad_l, qlen)
    for windowed in (0, 1):
        ref, out, cnt = run_emul(emul, oracle, wl, cands, pad_l, pad_r, windowed)
        assert check(ref, out, cands, wl) > 300


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


@pytest.mark.parametrize("pad_l,pad_r,qlen", [(30, 40, 50), (4, 17, 35), (16, 28, 41)])
def test_filter_lossless_all_cost2_variants(oracle, emul, pad_l, pad_r, qlen):
    """every substitution, deletion, single insertion and every pair of insertion slots of a few
    cores (repeats included), placed in the interior, at the pad limit, hanging over the read
    start and ending inside the core."""
    rng = np.random.default_rng(7 + qlen)
    cores = [rs(rng, 16) for _ in range(2)] + ["ACACACACACACACAC", "AAAACCCCGGGGTTTT",
                                                "AAAAAAAACAAAAAAA"]
    wl = set(cores)
    for c in cores:
        for _ in range(3):
            i, j = sorted(rng.choice(16, 2, replace=False))
            s = list(c)
            s[i] = "ACGT"[("ACGT".index(s[i]) + 1) % 4]
            s[j] = "ACGT"[("ACGT".index(s[j]) + 2) % 4]
            wl.add("".join(s))
    while len(wl) < 300:
        wl.add(rs(rng, 16))
    wl = sorted(wl)
    cands = []
    lim = min(pad_l, qlen - 19)
    for c in cores:
        for v in _variants(rng, c):
            o = max(0, int(rng.choice([0, 1, 2, lim - 1, lim, lim + 1, lim + 2, lim + 3])))
            cands.append((rs(rng, o) + v + rs(rng, 64))[:qlen])
            k = int(rng.integers(1, 3))
            cands.append((v[k:] + rs(rng, 64))[:qlen])
            k = int(rng.integers(0, 3))
            o2 = int(rng.integers(10, max(11, pad_l + 3)))
            cands.append((rs(rng, o2) + v[:len(v) - k])[-64:])
    ref, out, cnt = run_emul(emul, oracle, wl, cands, pad_l, pad_r, 1)
    assert check(ref, out, cands, wl) > 1000


def test_filter_lossless_737k(oracle, emul):
    """the real 737K list: heavy-tailed key multiplicities exercise the kstart row ranges."""
    from nanoranger_b200 import synth, whitelists
    wl_a = whitelists.load_737k()
    d = synth.make_candidates(wl_a, 200, seed=3)
    cands = synth.to_strings(d["seqs"], d["offsets"])
    wl = whitelists.ascii_to_strings(wl_a)
    ref, out, cnt = run_emul(emul, oracle, wl, cands, 30, 40, 1)
    assert check(ref, out, cands, wl) > 150


@pytest.mark.parametrize("limit,max_cost", [(1, 0), (16, 1)])
def test_probe_prefixes_complete_for_small_costs(oracle, emul, limit, max_cost):
    """NR_PROBES_COST0 / NR_PROBES_COST1 (nr_filter_core.h): with only that prefix of the probe
    table every candidate whose best cost is <= max_cost still gets the exact answer -- all pairs
    at the best cost are found.  The kernel relies on this after its first chunk of slots."""
    rng = np.random.default_rng(900 + limit)
    emul.nr_emul_set_probe_limit(limit)
    try:
        total = 0
        for pad_l, pad_r, qlen in [(30, 40, 50), (4, 17, 35), (2, 3, 30), (30, 40, 64)]:
            wl = tie_rich_whitelist(rng, 1500)
            cands = mixed_candidates(rng, wl, 2500, pad_l, qlen, with_n=0.0)
            # every single-event variant of some cores, at several offsets
            for core in wl[:40]:
                for v in _variants(rng, core):
                    a = int(rng.integers(0, min(pad_l + 2, 20)))
                    cands.append((rs(rng, a) + v + rs(rng, 40))[:qlen])
                # core hanging over the read end / the read start by one column
                cands.append(rs(rng, 14) + core[:-1])
                cands.append(core[1:] + rs(rng, 14))
                cands.append(rs(rng, 10) + core)
                cands.append(core + rs(rng, 10))
            ref, out, _ = run_emul(emul, oracle, wl, cands, pad_l, pad_r)
            sel = (16 - ref["best_score"] <= max_cost) & (out["took"] == 1)
            ok = ((out["score"] == ref["best_score"]) & (out["idx"] == ref["best_idx"]) &
                  (out["nbest"] == ref["n_best"]) & (out["strand"] == ref["strand"]) &
                  (out["umi"] == ref["umi_q"]))
            bad = np.flatnonzero(sel & ~ok)
            assert len(bad) == 0, (cands[bad[0]], {k: v[bad[0]] for k, v in ref.items()},
                                   {k: v[bad[0]] for k, v in out.items()})
            total += int(sel.sum())
        assert total > 1500
    finally:
        emul.nr_emul_set_probe_limit(0)


# ---- reads with one or two N: substituted variants + N-aware automaton --------------------------

def test_automaton_with_n_rows_equals_oracle_pair(oracle, emul):
    O = oracle
    rng = np.random.default_rng(11)
    n_low = 0
    for _ in range(8000):
        pad_l = int(rng.choice([30, 4, 16, 0, 2]))
        pad_r = int(rng.choice([40, 17, 28, 0, 3]))
        core = rs(rng, 16)
        mid = mutate(rng, core, int(rng.integers(0, 3)))
        pre, suf = rs(rng, int(rng.integers(0, 36))), rs(rng, int(rng.integers(0, 30)))
        mode = rng.integers(0, 6)
        q = (mid[int(rng.integers(1, 4)):] + suf) if mode == 0 else \
            (pre + mid[:-int(rng.integers(1, 4))]) if mode == 1 else (pre + mid + suf)
        q = list(q[:64])
        if not q:
            continue
        for _k in range(int(rng.integers(1, 3))):
            # N mostly inside / next to the core
            j = int(np.clip(len(pre) + rng.integers(-2, 19), 0, len(q) - 1)) if mode >= 2 else int(rng.integers(0, len(q)))
            q[j] = "N"
        q = "".join(q)
        a1, u1 = O.pair(q, core, pad_l, pad_r)
        qc, cc = np.ascontiguousarray(O.encode(q)), np.ascontiguousarray(O.encode(core))
        u = C.c_int(-9)
        cost = emul.nr_emul_nfa_n(P(qc, C.c_uint8), len(qc), P(cc, C.c_uint8), pad_l, pad_r, C.byref(u))
        if 16 - a1 <= 2:
            n_low += 1
            assert (cost, u.value) == (16 - a1, u1), (q, core, pad_l, pad_r)
        else:
            assert cost == 3, (q, core, pad_l, pad_r, a1)
    assert n_low > 600


def run_emul_n(E, O, wl, cands, pad_l, pad_r):
    wlc, _ = O.encode_many(wl, 16)
    cc, cl = O.encode_many(cands, 64)
    lo = np.zeros(len(wl), np.uint32)
    for j in range(16):
        lo |= wlc[:, j].astype(np.uint32) << np.uint32(2 * j)
    n = len(cands)
    out = dict(idx=np.zeros(n, np.int32), score=np.zeros(n, np.int8), nbest=np.zeros(n, np.int32),
               strand=np.zeros(n, np.uint8), umi=np.zeros(n, np.int16), took=np.zeros(n, np.uint8))
    cnt = np.zeros(2, np.int64)
    rc = E.nr_emul_filtered_n(P(lo, C.c_uint32), C.c_int64(len(wl)), pad_l, pad_r, P(cc, C.c_uint8),
                              P(cl.astype(np.uint8), C.c_uint8), C.c_int64(n), 24,
                              P(out["idx"], C.c_int32), P(out["score"], C.c_int8),
                              P(out["nbest"], C.c_int32), P(out["strand"], C.c_uint8),
                              P(out["umi"], C.c_int16), P(out["took"], C.c_uint8), P(cnt, C.c_int64))
    assert rc == 0
    ref = O.match(wlc, pad_l, pad_r, cc, cl)
    return ref, out, cnt


def _with_n(rng, q, k, lo, hi):
    q = list(q)
    for j in rng.choice(np.arange(max(0, lo), min(len(q), hi)), size=k, replace=False):
        q[int(j)] = "N"
    return "".join(q)


@pytest.mark.parametrize("pad_l,pad_r,qlen", [(30, 40, 50), (4, 17, 35), (16, 28, 41), (30, 40, 64)])
def test_filter_n_reads_lossless_random(oracle, emul, pad_l, pad_r, qlen):
    rng = np.random.default_rng(700 + qlen)
    wl = tie_rich_whitelist(rng, 2500)
    base = mixed_candidates(rng, wl, 2500, pad_l, qlen, with_n=0.0)
    cands = []
    for q in base:
        if len(q) < 2:
            continue
        k = int(rng.choice([1, 1, 2]))
        cands.append(_with_n(rng, q, k, 0, len(q)))
    ref, out, cnt = run_emul_n(emul, oracle, wl, cands, pad_l, pad_r)
    assert check(ref, out, cands, wl) > 200


def test_filter_n_reads_all_small_cost_variants(oracle, emul):
    """every <= 1-edit variant of a few cores, one or two N placed in / around the core, interior
    and at both read ends: all placements of true cost <= 2 must be found."""
    rng = np.random.default_rng(8)
    wl = tie_rich_whitelist(rng, 400)
    n_hi = 0
    for rep in range(3):
        core = wl[int(rng.integers(0, len(wl)))]
        cands = []
        for v in _variants(rng, core):
            for where in range(3):
                pre = rs(rng, [14, 0, 30][where])
                suf = rs(rng, [20, 25, 0][where])
                q = (pre + v + suf)[:64]
                a = len(pre)
                for k in (1, 2):
                    cands.append(_with_n(rng, q, k, a - 1, a + len(v) + 1))
        ref, out, cnt = run_emul_n(emul, oracle, wl, cands, 30, 40)
        n_hi += check(ref, out, cands, wl)
    assert n_hi > 500
