"""FILTERED rate against the 3M-sized synthetic whitelist (BASELINE config 4), device-resident.
usage: time_3m.py [n]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from nanoranger_b200 import Whitelist, NR_MODE_FILTERED, synth, whitelists
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 21
wl_a = whitelists.synthetic_whitelist(6794880)
wl = Whitelist(wl_a, 30, 40)
d = synth.make_candidates(wl_a, n, seed=33, p_n=1e-3)
dev = torch.device("cuda:0")
ds, do = torch.from_numpy(d["seqs"]).to(dev), torch.from_numpy(d["offsets"].view(np.int64).copy()).to(dev)
b, m, nm = wl.pack_device(ds, do)
ws = wl.workspace(n, dev)
# a single call right after the host-side set-up runs at idle clocks: warm up, then average
for _ in range(3):
    out = wl.match_device(b, m, nm, min_score=14, mode=NR_MODE_FILTERED, workspace=ws)
torch.cuda.synchronize()
REPS = 5
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(REPS):
    out = wl.match_device(b, m, nm, min_score=14, mode=NR_MODE_FILTERED, workspace=ws)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / REPS
print(f"3M-sized list: {n} candidates {ms:.1f} ms -> {n / ms * 1e3:.3e} cand/s, assigned {float(out.assigned(14).float().mean()):.3f}")
