"""Shared builders of parity inputs (seeded) for the CPU and GPU tests."""
from __future__ import annotations

import numpy as np

B = "ACGT"


def rs(rng, n):
    return "".join(B[i] for i in rng.integers(0, 4, n))


def mutate(rng, core, nops, ins_bias=0.5):
    s = list(core)
    for _ in range(nops):
        r = rng.random()
        if r < ins_bias:
            i = rng.integers(0, len(s) + 1)
            s.insert(i, B[rng.integers(0, 4)])
        elif r < ins_bias + (1 - ins_bias) / 2 and s:
            i = rng.integers(0, len(s))
            s[i] = B[(B.index(s[i]) + rng.integers(1, 4)) % 4]
        elif s:
            i = rng.integers(0, len(s))
            del s[i]
    return "".join(s)


def tie_rich_whitelist(rng, n, L=16):
    """random L-mers plus Hamming-2 neighbours of a third of them (ties at cost 2)."""
    wl = set()
    while len(wl) < n:
        c = rs(rng, L)
        wl.add(c)
        if rng.random() < 0.3:
            i, j = sorted(rng.choice(L, 2, replace=False))
            s = list(c)
            s[i] = B[(B.index(s[i]) + 1) % 4]
            s[j] = B[(B.index(s[j]) + 2) % 4]
            wl.add("".join(s))
    return sorted(wl)


def mixed_candidates(rng, wl, n, pad_l, qlen, with_n=0.02, rc=0.1):
    """candidates around whitelist entries: 0-3 edits, interior / overhanging / random /
    reverse-complemented / N-containing."""
    from oracle import oracle as O
    out = []
    for _ in range(n):
        core = wl[rng.integers(0, len(wl))]
        mid = mutate(rng, core, int(rng.choice([0, 1, 1, 2, 2, 2, 3])))
        mode = rng.integers(0, 8)
        a = int(rng.integers(0, min(pad_l + 5, max(1, qlen - 14)) + 1))
        if mode == 0:
            q = mid[int(rng.integers(1, 3)):] + rs(rng, qlen)
        elif mode == 1:
            o = int(rng.integers(8, pad_l + 3)) if pad_l > 6 else int(rng.integers(0, pad_l + 3))
            q = rs(rng, o) + mid[:len(mid) - int(rng.integers(0, 3))]
        elif mode == 2:
            q = rs(rng, qlen)
        else:
            q = rs(rng, a) + mid + rs(rng, max(0, qlen - a - len(mid)))
        q = q[:min(64, max(qlen, 1))] if mode != 1 else q[:64]
        if rng.random() < rc:
            q = O.revcomp(q)
        if rng.random() < with_n and len(q) > 0:
            j = rng.integers(0, len(q))
            q = q[:j] + "N" + q[j + 1:]
        out.append(q)
    return out


def compare(ref: dict, res, min_score: int, exact_below: bool, label=""):
    """ref: oracle.match dict; res: MatchResult with numpy arrays.  Bit-exact on every field for
    candidates whose best score >= min_score; below it either bit-exact too (exhaustive / AUTO)
    or flagged NR_FLAG_BELOW with an unresolved score (filtered)."""
    from nanoranger_b200 import _lib as K
    n = len(ref["best_score"])
    idx, score, nbest, flags, umi = (np.asarray(x) for x in
                                     (res.idx, res.score, res.nbest, res.flags, res.umi_q))
    hi = ref["best_score"] >= min_score
    exp_nbest = np.minimum(ref["n_best"], 255).astype(np.uint8)
    exp_umi = np.where(ref["umi_q"] < 0, K.NR_UMI_NONE, ref["umi_q"]).astype(np.uint8)
    full = hi | exact_below
    resolved = (flags & K.NR_FLAG_EXHAUSTIVE) != 0
    full = full | resolved
    ok = np.ones(n, bool)
    ok &= np.where(full, idx == ref["best_idx"], True)
    ok &= np.where(full, score == ref["best_score"], True)
    ok &= np.where(full, nbest == exp_nbest, True)
    ok &= np.where(full, umi == exp_umi, True)
    ok &= np.where(full, ((flags & K.NR_FLAG_TIE) != 0) == (ref["n_best"] > 1), True)
    ok &= np.where(full, ((flags & K.NR_FLAG_RC) != 0) == (ref["strand"] == 1), True)
    ok &= np.where(full, ((flags & K.NR_FLAG_BELOW) != 0) == (ref["best_score"] < min_score), True)
    ok &= np.where(full, ((flags & K.NR_FLAG_NO_UMI) != 0) == (ref["umi_q"] < 0), True)
    # unresolved: must be flagged below, and must really be below
    unres = ~full
    ok &= np.where(unres, ((flags & K.NR_FLAG_BELOW) != 0) & (score == K.NR_SCORE_BELOW) &
                   (idx == -1) & (nbest == 0), True)
    bad = np.flatnonzero(~ok)
    msg = ""
    if len(bad):
        i = bad[0]
        msg = (f"{label}: {len(bad)}/{n} mismatches; first at {i}: oracle "
               f"(idx {ref['best_idx'][i]}, AS {ref['best_score'][i]}, n {ref['n_best'][i]}, "
               f"strand {ref['strand'][i]}, umi {ref['umi_q'][i]}) vs cuda (idx {idx[i]}, AS {score[i]}, "
               f"n {nbest[i]}, flags {flags[i]:#x}, umi {umi[i]})")
    assert len(bad) == 0, msg
    return int(hi.sum())
