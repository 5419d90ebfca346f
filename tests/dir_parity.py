"""Run by tests/test_gpu_dir.py in a process of its own with NR_FILTER_DIR=1, so that every
whitelist -- not only those of millions of entries -- goes through the word-directory variant of
the filtered kernel (nr_match_filtered.cu, DIR = true).  Compared with the oracle field by field."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, HERE)
assert os.environ.get("NR_FILTER_DIR") == "1"

from helpers import compare, mixed_candidates, rs, tie_rich_whitelist  # noqa: E402
from test_gpu_match import _oracle, _run_device  # noqa: E402
from nanoranger_b200 import NR_MODE_AUTO, NR_MODE_FILTERED, Whitelist, synth, whitelists  # noqa: E402
from oracle import oracle as O  # noqa: E402

B = "ACGT"
rng = np.random.default_rng(41)

# crowded keys: for every quarter j, families of 5..40 entries that differ in quarter j only (their
# key with quarter j dropped has that many rows: words of the directory marked irregular next to
# regular ones), on top of a tie-rich list
wl = set(tie_rich_whitelist(rng, 2500))
for j in range(4):
    for fam in range(12):
        base = rs(rng, 16)
        for _ in range(int(rng.integers(5, 41))):
            wl.add(base[:4 * j] + rs(rng, 4) + base[4 * j + 4:])
wl = sorted(wl)
for pad_l, pad_r, qlen in ((30, 40, 50), (4, 17, 35), (2, 3, 30)):
    seqs = mixed_candidates(rng, wl, 3000, pad_l, qlen, with_n=0.05)
    w = Whitelist(wl, pad_l, pad_r)
    ref = _oracle(O, wl, pad_l, pad_r, seqs)
    res, ws = _run_device(w, seqs, 14, NR_MODE_FILTERED, counted=True)
    nhi = compare(ref, res, 14, exact_below=False, label=f"dir filtered {pad_l}/{pad_r}")
    assert w.counters(ws)["hits"] > 0 and nhi > 100
    res, _ = _run_device(w, seqs, 14, NR_MODE_AUTO)
    compare(ref, res, 14, exact_below=True, label=f"dir auto {pad_l}/{pad_r}")
    w.close()

wl_a = whitelists.load_737k()
d = synth.make_candidates(wl_a, 800, seed=6, p_n=0.01)
seqs = synth.to_strings(d["seqs"], d["offsets"])
w = Whitelist(wl_a, 30, 40)
cc, cl = O.encode_many(seqs, 64)
ref = O.match(O._CODE[wl_a], 30, 40, cc, cl)
res, _ = _run_device(w, seqs, 14, NR_MODE_FILTERED)
compare(ref, res, 14, exact_below=False, label="dir 737K filtered")
w.close()
print("dir_parity ok")
