#!/usr/bin/env python
"""The bench step's matcher alone (4 Mi flanks with N against 737K, FILTERED; then the counted
build for the work counters) for a targeted re-capture of profiles/inst_per_candidate.json after a
kernel edit: profiles/recapture_filtered.sh runs it under ncu, profiles/update_inst_json.py reads
the raw page."""
import os
import sys

R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, R)
import numpy as np
import torch

from nanoranger_b200 import NR_MODE_FILTERED, Whitelist, synth, whitelists

dev = torch.device("cuda:0")
B = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 22
wl_a = whitelists.load_737k()
wl = Whitelist(wl_a, 30, 40)
d = synth.make_candidates(wl_a, B, seed=2, p_n=1e-3)
d_seqs = torch.from_numpy(d["seqs"]).to(dev)
d_off = torch.from_numpy(d["offsets"].view(np.int64).copy()).to(dev)
ws = wl.workspace(B, dev)
out = wl.alloc_result(B, dev)
print("PHASE flanks-filtered", B, flush=True)
bases, meta, nmask = wl.pack_device(d_seqs, d_off)
wl.match_device(bases, meta, nmask, min_score=14, mode=NR_MODE_FILTERED, out=out, workspace=ws)
torch.cuda.synchronize()
wl.match_device(bases, meta, nmask, min_score=14, out=out, workspace=ws, counted=True)
torch.cuda.synchronize()
c = wl.counters(ws)
print("COUNTERS", {k: v / B for k, v in c.items()}, flush=True)
