"""End-to-end rate of nr_match_host (pinned host buffers) on the bench workload.
usage: time_host.py [n] [p_n]   (NR_HOST_CHUNK_LOG2 selects the chunk size)"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from nanoranger_b200 import MatchResult, NR_MODE_AUTO, NR_MODE_FILTERED, Whitelist, synth, whitelists
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 22
p_n = float(sys.argv[2]) if len(sys.argv) > 2 else 1e-3
wl_a = whitelists.load_737k()
d = synth.make_candidates(wl_a, n, seed=2, p_n=p_n)
wl = Whitelist(wl_a, 30, 40)
def pinned(a):
    t = torch.empty(a.shape, dtype=torch.from_numpy(a[:0]).dtype, pin_memory=True); t.numpy()[...] = a; return t
ps, po = pinned(d["seqs"]), pinned(d["offsets"].view(np.int64))
outs = [torch.empty(n, dtype=dt, pin_memory=True) for dt in (torch.int32, torch.int8, torch.uint8, torch.uint8, torch.uint8)]
res = MatchResult(*(t.numpy() for t in outs))
for mode, name in ((NR_MODE_FILTERED, "filtered"), (NR_MODE_AUTO, "auto")):
    wl.match_host(ps.numpy(), po.numpy().view(np.uint64), min_score=14, mode=mode, out=res)
    t0 = time.perf_counter()
    for _ in range(4):
        wl.match_host(ps.numpy(), po.numpy().view(np.uint64), min_score=14, mode=mode, out=res)
    dt = (time.perf_counter() - t0) / 4
    print(f"chunk 2^{os.environ.get('NR_HOST_CHUNK_LOG2', '20')} {name}: {n / dt:.3e} cand/s ({dt * 1e3:.1f} ms)")
