"""Times NR_MODE_AUTO against NR_MODE_FILTERED on the bench workload (device-resident) and prints
where AUTO resolved its candidates.  usage: time_auto.py [n] [frac_negative] [p_n]"""
import os, sys
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, R)
import numpy as np, torch
from nanoranger_b200 import Whitelist, NR_MODE_AUTO, NR_MODE_FILTERED, synth, whitelists
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 20
fneg = float(sys.argv[2]) if len(sys.argv) > 2 else 0.10
p_n = float(sys.argv[3]) if len(sys.argv) > 3 else 0.0
wl_a = whitelists.load_737k()
d = synth.make_candidates(wl_a, n, seed=2, frac_negative=fneg, p_n=p_n)
wl = Whitelist(wl_a, 30, 40)
dev = torch.device("cuda:0")
d_seqs = torch.from_numpy(d["seqs"]).to(dev)
d_off = torch.from_numpy(d["offsets"].view(np.int64).copy()).to(dev)
bases, meta, nmask = wl.pack_device(d_seqs, d_off)
ws = wl.workspace(n, dev)
for mode, name in ((NR_MODE_FILTERED, "filtered"), (NR_MODE_AUTO, "auto")):
    # a single call right after the host-side set-up runs at idle clocks: warm up, then average
    for _ in range(3):
        out = wl.match_device(bases, meta, nmask, min_score=14, mode=mode, workspace=ws)
    torch.cuda.synchronize()
    REPS = 5
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(REPS):
        out = wl.match_device(bases, meta, nmask, min_score=14, mode=mode, workspace=ws)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / REPS
    t = wl.tier_counts(ws)
    sc = out.score.cpu().numpy()
    hist = {int(v): int(c) for v, c in zip(*np.unique(sc, return_counts=True))}
    print(f"{name}: {n} candidates {ms:.1f} ms -> {n / ms * 1e3:.3e} cand/s; tiers {t}; "
          f"deep rate {(t['deep_k3'] + t['deep_k5']) / max(ms, 1e-9) * 1e3:.3e}/s (upper bound: whole call time)")
    print("   score histogram", hist)
