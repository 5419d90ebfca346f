"""Parity of the deep matcher tier (nr_match_deep.cu) through the C ABI: NR_MODE_AUTO must be
bit-exact to the oracle at EVERY score (the reference histograms all uniquely mapped forward
reads, utils.py:698, 728-730), and the tier -- not the brute-force kernel -- must be what
resolved the low-scoring reads."""
import numpy as np
import pytest

from helpers import compare, mixed_candidates, mutate, rs, tie_rich_whitelist
from test_gpu_match import _oracle, _run_device

pytestmark = pytest.mark.gpu


def test_auto_737k_every_score_exact(cuda_device, oracle):
    from nanoranger_b200 import NR_MODE_AUTO, Whitelist, synth, whitelists
    wl_a = whitelists.load_737k()
    d = synth.make_candidates(wl_a, 2500, seed=21, frac_negative=0.5, p_sub=0.04, p_ins=0.04, p_del=0.04)
    seqs = synth.to_strings(d["seqs"], d["offsets"])
    rng = np.random.default_rng(3)
    seqs += [rs(rng, int(rng.integers(1, 24))) for _ in range(60)]            # shorter than the filter takes
    seqs += [s[:k] + "N" + s[k + 1:] for s, k in zip(seqs[:80], rng.integers(0, 30, 80))]
    seqs += ["N" * 30, "ACGT" * 16, "A" * 63, rs(rng, 64), rs(rng, 63)]
    wl = Whitelist(wl_a, 30, 40)
    cc, cl = oracle.encode_many(seqs, 64)
    ref = oracle.match(oracle._CODE[wl_a], 30, 40, cc, cl)
    res, ws = _run_device(wl, seqs, 14, NR_MODE_AUTO)
    compare(ref, res, 14, exact_below=True, label="auto 737K")
    t = wl.tier_counts(ws)
    below = int((ref["best_score"] < 14).sum())
    assert t["left_by_filter"] >= below
    assert t["deep_k3"] + t["deep_k5"] + t["brute_force"] == t["left_by_filter"]
    # nearly everything the filter leaves has a pair at cost <= 3 somewhere in 737K barcodes
    assert t["deep_k3"] > 0.8 * t["left_by_filter"], t
    assert t["brute_force"] <= 80, t          # the very short reads


@pytest.mark.parametrize("pad_l,pad_r,qlen", [(30, 40, 50), (4, 17, 35), (2, 3, 30), (30, 40, 63)])
def test_auto_small_whitelists_all_tiers(cuda_device, oracle, pad_l, pad_r, qlen, monkeypatch):
    """small whitelists: most random reads have no pair at cost <= 3, so K = 5 and the brute-force
    kernel get work too.  (Left to itself the API skips the deep tier on a list that shares this
    little -- nr_deep_usable's cost rule -- hence NR_DEEP_TIER=always; the rule has its own test.)"""
    from nanoranger_b200 import NR_MODE_AUTO, Whitelist
    monkeypatch.setenv("NR_DEEP_TIER", "always")
    rng = np.random.default_rng(300 + qlen)
    wl_strs = tie_rich_whitelist(rng, 3000)
    seqs = mixed_candidates(rng, wl_strs, 2500, pad_l, qlen, with_n=0.1)
    seqs += [mutate(rng, wl_strs[int(rng.integers(0, len(wl_strs)))], int(rng.integers(2, 7))) for _ in range(300)]
    wl = Whitelist(wl_strs, pad_l, pad_r)
    ref = _oracle(oracle, wl_strs, pad_l, pad_r, seqs)
    res, ws = _run_device(wl, seqs, 14, NR_MODE_AUTO)
    compare(ref, res, 14, exact_below=True, label="auto small")
    t = wl.tier_counts(ws)
    assert t["deep_k3"] > 0 and t["deep_k5"] > 0
    assert t["deep_k3"] + t["deep_k5"] + t["brute_force"] == t["left_by_filter"]


def test_auto_without_seed_index_goes_through_deep_tier(cuda_device, oracle, monkeypatch):
    """cores no seed filter covers (20 columns, N columns in some entries): no seed index, AUTO =
    deep tier over every candidate (forced: 4 000 random cores share too little for the cost rule)."""
    from nanoranger_b200 import NR_MODE_AUTO, Whitelist
    monkeypatch.setenv("NR_DEEP_TIER", "always")
    rng = np.random.default_rng(9)
    bcs = set()
    while len(bcs) < 4000:
        b = list(rs(rng, 20))
        if rng.random() < 0.15:
            b[int(rng.integers(0, 20))] = "N"
        bcs.add("".join(b))
    wl_strs = sorted(bcs)
    seqs = []
    for _ in range(2000):
        core = wl_strs[int(rng.integers(0, len(wl_strs)))].replace("N", "ACGT"[int(rng.integers(0, 4))])
        mid = mutate(rng, core, int(rng.choice([0, 0, 1, 2, 3, 5])))
        q = (rs(rng, int(rng.integers(0, 18))) + mid + rs(rng, int(rng.integers(0, 14))))[:64]
        if rng.random() < 0.1:
            q = oracle.revcomp(q)
        seqs.append(q)
    wl = Whitelist(wl_strs, 15, 24)
    assert not wl.has_index
    ref = _oracle(oracle, wl_strs, 15, 24, seqs)
    res, ws = _run_device(wl, seqs, 18, NR_MODE_AUTO)
    compare(ref, res, 18, exact_below=True, label="auto 20-column cores")
    t = wl.tier_counts(ws)
    assert t["deep_k3"] > 1000, t


def test_deep_tier_cost_rule(cuda_device, oracle, monkeypatch):
    """nr_deep_usable: a small random list shares too little between its entries for meeting in the
    middle to beat the bit-parallel brute-force kernel -- AUTO sends what the filter leaves straight
    there; the 737K list keeps the deep tier; NR_DEEP_TIER=never / always override.  Bit-exact in
    every combination."""
    from nanoranger_b200 import NR_MODE_AUTO, Whitelist, synth, whitelists
    rng = np.random.default_rng(12)
    wl_strs = tie_rich_whitelist(rng, 3000)
    seqs = mixed_candidates(rng, wl_strs, 1200, 30, 50, with_n=0.1)
    wl = Whitelist(wl_strs, 30, 40)
    ref = _oracle(oracle, wl_strs, 30, 40, seqs)
    for setting, deep in ((None, False), ("always", True), ("never", False)):
        if setting is None:
            monkeypatch.delenv("NR_DEEP_TIER", raising=False)
        else:
            monkeypatch.setenv("NR_DEEP_TIER", setting)
        res, ws = _run_device(wl, seqs, 14, NR_MODE_AUTO)
        compare(ref, res, 14, exact_below=True, label=f"small list, NR_DEEP_TIER={setting}")
        t = wl.tier_counts(ws)
        assert (t["deep_k3"] + t["deep_k5"] > 0) == deep, (setting, t)
        assert t["deep_k3"] + t["deep_k5"] + t["brute_force"] == t["left_by_filter"]
    wl_a = whitelists.load_737k()
    d = synth.make_candidates(wl_a, 400, seed=5, frac_negative=0.5)
    seqs = synth.to_strings(d["seqs"], d["offsets"])
    wl = Whitelist(wl_a, 30, 40)
    cc, cl = oracle.encode_many(seqs, 64)
    ref = oracle.match(oracle._CODE[wl_a], 30, 40, cc, cl)
    for setting, deep in ((None, True), ("never", False)):
        if setting is None:
            monkeypatch.delenv("NR_DEEP_TIER", raising=False)
        else:
            monkeypatch.setenv("NR_DEEP_TIER", setting)
        res, ws = _run_device(wl, seqs, 14, NR_MODE_AUTO)
        compare(ref, res, 14, exact_below=True, label=f"737K, NR_DEEP_TIER={setting}")
        t = wl.tier_counts(ws)
        assert (t["deep_k3"] + t["deep_k5"] > 0) == deep, (setting, t)


@pytest.mark.parametrize("mode_name", ["filtered", "auto"])
def test_reads_with_n_through_the_n_pass(cuda_device, oracle, mode_name):
    """reads with one or two N are resolved by the filter's N pass (substituted variants), reads
    with more by the deep tier; bit-exact either way."""
    from nanoranger_b200 import NR_MODE_AUTO, NR_MODE_FILTERED, Whitelist, synth, whitelists
    mode = NR_MODE_AUTO if mode_name == "auto" else NR_MODE_FILTERED
    wl_a = whitelists.load_737k()
    d = synth.make_candidates(wl_a, 3000, seed=33, p_n=0.03)
    seqs = synth.to_strings(d["seqs"], d["offsets"])
    n_with_n = sum("N" in s for s in seqs)
    assert n_with_n > 1500
    wl = Whitelist(wl_a, 30, 40)
    cc, cl = oracle.encode_many(seqs, 64)
    ref = oracle.match(oracle._CODE[wl_a], 30, 40, cc, cl)
    res, ws = _run_device(wl, seqs, 14, mode)
    nhi = compare(ref, res, 14, exact_below=(mode == NR_MODE_AUTO), label=f"N reads {mode_name}")
    assert nhi > 1000
    t = wl.tier_counts(ws)
    # only reads with > 2 N (and in AUTO the sub-threshold ones) may reach the deep tier
    many = sum(s.count("N") > 2 for s in seqs)
    if mode == NR_MODE_FILTERED:
        assert t["left_by_filter"] <= many + 5, (t, many)


def test_reads_with_n_small_whitelist_random_geometries(cuda_device, oracle):
    from nanoranger_b200 import NR_MODE_FILTERED, Whitelist
    rng = np.random.default_rng(77)
    for pad_l, pad_r, qlen in [(30, 40, 50), (4, 17, 35), (16, 28, 41), (30, 40, 64)]:
        wl_strs = tie_rich_whitelist(rng, 3000)
        base = mixed_candidates(rng, wl_strs, 3000, pad_l, qlen, with_n=0.0)
        seqs = []
        for q in base:
            q = list(q)
            for _ in range(int(rng.choice([0, 1, 1, 2, 3]))):
                if q:
                    q[int(rng.integers(0, len(q)))] = "N"
            seqs.append("".join(q))
        wl = Whitelist(wl_strs, pad_l, pad_r)
        ref = _oracle(oracle, wl_strs, pad_l, pad_r, seqs)
        res, _ = _run_device(wl, seqs, 14, NR_MODE_FILTERED)
        compare(ref, res, 14, exact_below=False, label=f"N reads {pad_l}/{pad_r}/{qlen}")
