import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__)))); sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np, torch
from nanoranger_b200 import Whitelist, NR_MODE_AUTO, NR_MODE_FILTERED, pack_ascii
from nanoranger_b200.whitelists import LINKER_SLIDESEQ
from helpers import rs, mutate
rng = np.random.default_rng(5)
bcs = sorted({rs(rng, 14) for _ in range(17753)})
bcs = [b if rng.random() > 0.15 else b[:3] + "N" + b[4:] for b in bcs]
cores = [b[:8] + LINKER_SLIDESEQ + b[8:] for b in bcs]
n = int(sys.argv[1]) if len(sys.argv) > 1 else 20000
seqs = []
for _ in range(n):
    c = cores[rng.integers(0, len(cores))].replace("N", "ACGT"[rng.integers(0, 4)])
    mid = mutate(rng, c, int(rng.integers(0, 3)))
    a = int(rng.integers(0, 18))
    seqs.append((rs(rng, a) + mid + rs(rng, 30))[:int(rng.integers(46, 57))])
wl = Whitelist(cores, 15, 24)
buf, off = pack_ascii(seqs)
dev = torch.device("cuda:0")
d_seqs = torch.from_numpy(buf.copy()).to(dev); d_off = torch.from_numpy(off.view(np.int64).copy()).to(dev)
bases, meta, nmask = wl.pack_device(d_seqs, d_off)
ws = wl.workspace(n, dev)
def timed(mode, reps=4, warm=2):
    for _ in range(warm):
        out = wl.match_device(bases, meta, nmask, min_score=30, mode=mode, workspace=ws)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        out = wl.match_device(bases, meta, nmask, min_score=30, mode=mode, workspace=ws)
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps, out


# AUTO three ways: the API's own choice (nr_deep_usable's cost rule), deep tier forced, deep tier off
for mode, name, deep in ((NR_MODE_FILTERED, "filtered", None), (NR_MODE_AUTO, "auto", None),
                         (NR_MODE_AUTO, "auto", "always"), (NR_MODE_AUTO, "auto", "never")):
    if deep is None:
        os.environ.pop("NR_DEEP_TIER", None)
    else:
        os.environ["NR_DEEP_TIER"] = deep
    ms, out = timed(mode)
    print(f"slide-seq {name} (NR_DEEP_TIER={deep}): {n} candidates x {len(cores)} entries: {ms:.1f} ms, "
          f"{n/ms*1e3:.3e} cand/s, assigned {float(out.assigned(30).float().mean()):.3f}, "
          f"tiers {wl.tier_counts(ws)}", flush=True)
