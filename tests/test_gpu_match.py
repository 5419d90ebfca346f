"""Parity of the CUDA matcher (through the C ABI) against the CPU oracle.  Bit-exact."""
import numpy as np
import pytest

from helpers import compare, mixed_candidates, mutate, rs, tie_rich_whitelist

pytestmark = pytest.mark.gpu


def _run_device(wl, seqs, min_score, mode, counted=False):
    import torch
    from nanoranger_b200 import pack_ascii
    buf, off = pack_ascii(seqs)
    dev = torch.device("cuda:0")
    d_seqs = torch.from_numpy(buf.copy()).to(dev) if len(buf) else torch.zeros(1, dtype=torch.uint8, device=dev)
    d_off = torch.from_numpy(off.view(np.int64).copy()).to(dev)
    bases, meta, nmask = wl.pack_device(d_seqs, d_off)
    ws = wl.workspace(len(seqs), dev)
    # poisoned outputs and workspace: a kernel that reads a result slot nobody has written yet (or
    # relies on a zeroed workspace beyond the header the API clears) fails here, not once in a while
    out = wl.alloc_result(len(seqs), dev)
    out.idx.fill_(0x7FFFFFF0); out.score.fill_(-77); out.nbest.fill_(254); out.flags.fill_(0xFF)
    out.umi_q.fill_(254)
    ws.fill_(0xA5)
    res = wl.match_device(bases, meta, nmask, min_score=min_score, mode=mode, workspace=ws,
                          counted=counted, out=out)
    torch.cuda.synchronize()
    from nanoranger_b200 import MatchResult
    out = MatchResult(*(t.cpu().numpy() for t in (res.idx, res.score, res.nbest, res.flags, res.umi_q)))
    return out, ws


def _oracle(oracle, wl_strs, pad_l, pad_r, seqs):
    L = len(wl_strs[0])
    wlc, _ = oracle.encode_many(wl_strs, L)
    keep = [s if len(s) <= 64 else "" for s in seqs]
    cc, cl = oracle.encode_many(keep, 64)
    return oracle.match(wlc, pad_l, pad_r, cc, cl)


GEOMS = [(30, 40, 50), (4, 17, 35), (16, 28, 41), (30, 40, 64), (2, 3, 30)]


@pytest.mark.parametrize("pad_l,pad_r,qlen", GEOMS)
def test_exhaustive_vs_oracle(cuda_device, oracle, pad_l, pad_r, qlen):
    from nanoranger_b200 import Whitelist, NR_MODE_EXHAUSTIVE
    rng = np.random.default_rng(100 + qlen)
    wl_strs = tie_rich_whitelist(rng, 2500)
    seqs = mixed_candidates(rng, wl_strs, 1500, pad_l, qlen)
    seqs += ["A", "ACGT" * 16, "N" * 20, rs(rng, 1), rs(rng, 2), rs(rng, 15), rs(rng, 16)]
    wl = Whitelist(wl_strs, pad_l, pad_r)
    ref = _oracle(oracle, wl_strs, pad_l, pad_r, seqs)
    res, _ = _run_device(wl, seqs, 14, NR_MODE_EXHAUSTIVE)
    compare(ref, res, 14, exact_below=True, label="exhaustive")


@pytest.mark.parametrize("pad_l,pad_r,qlen", GEOMS)
def test_filtered_and_auto_vs_oracle(cuda_device, oracle, pad_l, pad_r, qlen):
    from nanoranger_b200 import Whitelist, NR_MODE_AUTO, NR_MODE_FILTERED
    rng = np.random.default_rng(200 + qlen)
    wl_strs = tie_rich_whitelist(rng, 3000)
    seqs = mixed_candidates(rng, wl_strs, 4000, pad_l, qlen)
    wl = Whitelist(wl_strs, pad_l, pad_r)
    assert wl.has_index
    ref = _oracle(oracle, wl_strs, pad_l, pad_r, seqs)
    res, ws = _run_device(wl, seqs, 14, NR_MODE_FILTERED, counted=True)
    nhi = compare(ref, res, 14, exact_below=False, label="filtered")
    cnt = wl.counters(ws)
    assert cnt["probes"] > 0 and cnt["verifications"] >= cnt["passes"]
    res2, _ = _run_device(wl, seqs, 14, NR_MODE_FILTERED)
    for a, b in zip((res.idx, res.score, res.nbest, res.flags, res.umi_q),
                    (res2.idx, res2.score, res2.nbest, res2.flags, res2.umi_q)):
        assert np.array_equal(a, b)
    res3, _ = _run_device(wl, seqs, 14, NR_MODE_AUTO)
    compare(ref, res3, 14, exact_below=True, label="auto")
    assert nhi > 200 or pad_l < 4


def test_737k_synthetic_vs_oracle(cuda_device, oracle):
    """C4-shaped inputs at a size the oracle finishes in seconds: real 737K list, ONT errors."""
    from nanoranger_b200 import Whitelist, NR_MODE_AUTO, NR_MODE_FILTERED, synth, whitelists
    wl_a = whitelists.load_737k()
    d = synth.make_candidates(wl_a, 600, seed=1)
    seqs = synth.to_strings(d["seqs"], d["offsets"])
    wl_strs = whitelists.ascii_to_strings(wl_a)
    wl = Whitelist(wl_a, 30, 40)
    ref = _oracle(oracle, wl_strs, 30, 40, seqs)
    res, _ = _run_device(wl, seqs, 14, NR_MODE_FILTERED)
    compare(ref, res, 14, exact_below=False, label="737K filtered")
    res, _ = _run_device(wl, seqs[:150], 14, NR_MODE_AUTO)
    sub = {k: v[:150] for k, v in ref.items()}
    compare(sub, res, 14, exact_below=True, label="737K auto")


def test_slideseq_geometry_generic_kernel(cuda_device, oracle):
    """32-column cores with N inside (write_bc_slideseq, utils.py:584-601): the brute-force generic
    kernel (NR_MODE_EXHAUSTIVE) and AUTO (anchored filter + deep tier) against the oracle."""
    from nanoranger_b200 import Whitelist, NR_MODE_AUTO, NR_MODE_EXHAUSTIVE
    from nanoranger_b200.whitelists import LINKER_SLIDESEQ
    rng = np.random.default_rng(5)
    bcs = sorted({rs(rng, 14) for _ in range(1500)})
    bcs = [b if rng.random() > 0.15 else b[:3] + "N" + b[4:] for b in bcs]
    cores = [b[:8] + LINKER_SLIDESEQ + b[8:] for b in bcs]
    seqs = []
    for _ in range(800):
        c = cores[rng.integers(0, len(cores))].replace("N", "ACGT"[rng.integers(0, 4)])
        from helpers import mutate
        mid = mutate(rng, c, int(rng.integers(0, 3)))
        a = int(rng.integers(0, 18))
        seqs.append((rs(rng, a) + mid + rs(rng, 30))[:int(rng.integers(40, 60))])
    wl = Whitelist(cores, 15, 24)
    assert wl.has_index                      # the anchored index (8 + linker + 6)
    ref = _oracle(oracle, cores, 15, 24, seqs)
    res, _ = _run_device(wl, seqs, 30, NR_MODE_EXHAUSTIVE)
    compare(ref, res, 30, exact_below=True, label="slideseq brute force")
    res, _ = _run_device(wl, seqs, 30, NR_MODE_AUTO)
    compare(ref, res, 30, exact_below=True, label="slideseq auto")


def test_host_entry_point_matches_device(cuda_device, oracle):
    from nanoranger_b200 import Whitelist, NR_MODE_AUTO, NR_MODE_FILTERED
    rng = np.random.default_rng(9)
    wl_strs = tie_rich_whitelist(rng, 2000)
    seqs = mixed_candidates(rng, wl_strs, 3000, 30, 50) + ["ACGT" * 20, ""]
    wl = Whitelist(wl_strs, 30, 40)
    for mode in (NR_MODE_FILTERED, NR_MODE_AUTO):
        a, _ = _run_device(wl, seqs, 14, mode)
        b = wl.match_host(seqs, min_score=14, mode=mode)
        for x, y in zip((a.idx, a.score, a.nbest, a.flags, a.umi_q),
                        (b.idx, b.score, b.nbest, b.flags, b.umi_q)):
            assert np.array_equal(x, y)
    from nanoranger_b200 import _lib as K
    assert b.flags[-2] & K.NR_FLAG_TOO_LONG
    empty = wl.match_host([], min_score=14)
    assert len(empty.idx) == 0


def test_filtered_mode_refused_when_unsupported(cuda_device):
    from nanoranger_b200 import Whitelist, NR_MODE_FILTERED
    wl = Whitelist(["ACGTACGTACGTACGTACGT", "ACGTACGTACGTACGTACGA"], 4, 4)
    with pytest.raises(RuntimeError):
        wl.match_host(["ACGTACGTACGTACGTACGTAA"], min_score=18, mode=NR_MODE_FILTERED)
    wl16 = Whitelist(["ACGTACGTACGTACGT", "ACGTACGTACGTACGA"], 4, 4)
    with pytest.raises(RuntimeError):
        wl16.match_host(["ACGTACGTACGTACGTACGTAA"], min_score=10, mode=NR_MODE_FILTERED)


@pytest.mark.parametrize("n", [1, 3, 37, 295, 297])
def test_exhaustive_small_batches_split_over_whitelist(cuda_device, oracle, n):
    """Batches smaller than the grid: the whitelist scan of each candidate is cut into slices
    handled by different blocks and merged by the last block to arrive."""
    from nanoranger_b200 import Whitelist, NR_MODE_AUTO, NR_MODE_EXHAUSTIVE
    rng = np.random.default_rng(700 + n)
    wl_strs = tie_rich_whitelist(rng, 5000)
    seqs = mixed_candidates(rng, wl_strs, n, 30, 50, with_n=0.3)
    wl = Whitelist(wl_strs, 30, 40)
    ref = _oracle(oracle, wl_strs, 30, 40, seqs)
    for mode in (NR_MODE_EXHAUSTIVE, NR_MODE_AUTO):
        for rep in range(2):      # twice: the arrival counters must be back at zero
            res, _ = _run_device(wl, seqs, 14, mode)
            compare(ref, res, 14, exact_below=True, label=f"split n={n} mode={mode} rep={rep}")


def test_dense_index_3M_sized_whitelist_vs_oracle(cuda_device, oracle):
    """C4's whitelist size (6 794 880 synthetic 16-mers): a third of all 24-bit keys are set, items
    overflow the hit queue and are queued in probe ranges; every answer still equals the oracle's."""
    from nanoranger_b200 import Whitelist, NR_MODE_FILTERED, synth, whitelists
    wl_a = whitelists.synthetic_whitelist(6794880)
    d = synth.make_candidates(wl_a, 160, seed=4)
    seqs = synth.to_strings(d["seqs"], d["offsets"])
    wl = Whitelist(wl_a, 30, 40)
    wlc = oracle._CODE[wl_a]
    cc, cl = oracle.encode_many(seqs, 64)
    ref = oracle.match(wlc, 30, 40, cc, cl)
    res, ws = _run_device(wl, seqs, 14, NR_MODE_FILTERED, counted=True)
    nhi = compare(ref, res, 14, exact_below=False, label="3M filtered")
    assert nhi > 40
    assert wl.counters(ws)["hits"] > 300 * len(seqs)


def test_adversarial_cases_vs_oracle(cuda_device, oracle):
    """barcode at offset 0 / exactly padL / beyond the pad, homopolymer
    and low-complexity barcodes (every probe of a slot hits), N in the candidate, lengths 1..64."""
    from nanoranger_b200 import Whitelist, NR_MODE_AUTO, NR_MODE_FILTERED
    rng = np.random.default_rng(77)
    wl_strs = ["A" * 16, "C" * 16, "ACAC" * 4, "AAAAAAAACCCCCCCC", "ACGT" * 4, "A" * 15 + "C", "C" + "A" * 15,
               "AAAACCCCGGGGTTTT", "TTTTGGGGCCCCAAAA"] + tie_rich_whitelist(rng, 400)
    wl_strs = sorted(set(wl_strs))
    pad_l, pad_r = 30, 40
    seqs = []
    for bc in wl_strs[:60]:
        for off in (0, 1, 14, 29, 30, 31, 32, 33, 40):
            seqs.append((rs(rng, off) + bc + rs(rng, 20))[:64])
            seqs.append((rs(rng, off) + mutate(rng, bc, 1) + rs(rng, 20))[:64])
        seqs.append(bc[3:] + rs(rng, 30))          # start overhang
        seqs.append(rs(rng, 30) + bc[:13])         # end overhang
        seqs.append("A" * 50)
        seqs.append(("A" * 20 + bc + "A" * 20)[:64])
        seqs.append(bc[:8] + "N" + bc[8:] + rs(rng, 20))
    seqs += [rs(rng, n) for n in range(0, 65)]
    wl = Whitelist(wl_strs, pad_l, pad_r)
    ref = _oracle(oracle, wl_strs, pad_l, pad_r, seqs)
    res, _ = _run_device(wl, seqs, 14, NR_MODE_FILTERED)
    compare(ref, res, 14, exact_below=False, label="adversarial filtered")
    res, _ = _run_device(wl, seqs, 14, NR_MODE_AUTO)
    compare(ref, res, 14, exact_below=True, label="adversarial auto")
    assert (ref["n_best"] > 32).sum() >= 0


@pytest.mark.parametrize("L", [16, 20, 32])
def test_generic_kernel_core_lengths_with_n(cuda_device, oracle, L):
    """the three instantiations of the generic exhaustive kernel (compile-time 16 and 32 columns,
    run-time otherwise), whitelist entries with N inside"""
    from nanoranger_b200 import Whitelist, NR_MODE_AUTO
    rng = np.random.default_rng(300 + L)
    cores = sorted({rs(rng, L) for _ in range(700)})
    cores = [c if rng.random() > 0.2 else c[:5] + "N" + c[6:] for c in cores]
    seqs = []
    for _ in range(500):
        c = cores[rng.integers(0, len(cores))].replace("N", "ACGT"[rng.integers(0, 4)])
        seqs.append((rs(rng, int(rng.integers(0, 12))) + mutate(rng, c, int(rng.integers(0, 3))) + rs(rng, 25))[:64])
    seqs += ["", "A", rs(rng, 64), "N" * 40]
    wl = Whitelist(cores, 10, 20)
    assert not wl.has_index
    wlc, _ = oracle.encode_many(cores, L)
    cc, cl = oracle.encode_many(seqs, 64)
    ref = oracle.match(wlc, 10, 20, cc, cl)
    res, _ = _run_device(wl, seqs, L - 2, NR_MODE_AUTO)
    compare(ref, res, L - 2, exact_below=True, label=f"generic L={L}")
