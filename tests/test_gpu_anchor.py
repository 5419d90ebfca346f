"""Parity of the anchored seed filter (nr_match_anchored.cu: slide-seq cores, 8 + linker 18 + 6
columns, utils.py:584-601; threshold AS >= 30, utils.py:638) through the C ABI against the oracle.
Bit-exact."""
import gzip
import os

import numpy as np
import pytest

from helpers import compare
from test_anchor_emul import LINKER, slide_candidates, slide_whitelist
from test_gpu_match import _oracle, _run_device

pytestmark = pytest.mark.gpu
G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.mark.parametrize("pad_l,pad_r", [(15, 24), (0, 0), (3, 2), (30, 40)])
def test_anchored_filtered_and_auto_vs_oracle(cuda_device, oracle, pad_l, pad_r):
    from nanoranger_b200 import NR_MODE_AUTO, NR_MODE_FILTERED, Whitelist
    rng = np.random.default_rng(90 + pad_l)
    wl_strs = slide_whitelist(rng, 3000)
    seqs = slide_candidates(rng, oracle, wl_strs, 4000)
    seqs += slide_candidates(rng, oracle, wl_strs, 1500, with_n=1.0)              # one to three N anywhere
    seqs += [s[:k] + "N" + s[k + 1:] for s, k in zip(seqs[:200], rng.integers(0, 30, 200))]
    seqs += ["A", "N" * 40, LINKER, LINKER * 3, "ACGT" * 16]
    wl = Whitelist(wl_strs, pad_l, pad_r)
    assert wl.has_index
    ref = _oracle(oracle, wl_strs, pad_l, pad_r, seqs)
    res, ws = _run_device(wl, seqs, 30, NR_MODE_FILTERED, counted=True)
    nhi = compare(ref, res, 30, exact_below=False, label="anchored filtered")
    assert nhi > (1000 if pad_l >= 15 else 0)
    c = wl.counters(ws)
    assert c["probes"] > 0 and c["verifications"] > 0
    res, ws = _run_device(wl, seqs, 30, NR_MODE_AUTO)
    compare(ref, res, 30, exact_below=True, label="anchored auto")
    t = wl.tier_counts(ws)
    assert t["deep_k3"] + t["deep_k5"] + t["brute_force"] == t["left_by_filter"]


def test_anchored_real_slideseq_list_and_fixture(cuda_device):
    from nanoranger_b200 import NR_MODE_FILTERED, Whitelist, fastx
    names, seqs, off = fastx.read_fasta(os.path.join(G, "slideseq.fa.gz"))
    ref = dict(np.load(os.path.join(G, "slideseq.oracle.npz")))
    bcs = gzip.open(os.path.join(G, "slideseq_whitelist.txt.gz"), "rt").read().split()
    wl = Whitelist([b[:8] + LINKER + b[8:] for b in bcs], 15, 24)
    assert wl.has_index
    r = wl.match_host(seqs, off, min_score=30, mode=NR_MODE_FILTERED)
    assert compare(ref, r, 30, exact_below=False, label="slideseq filtered") > 1000
