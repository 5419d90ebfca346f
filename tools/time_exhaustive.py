"""Times the brute-force DP kernel (NR_MODE_EXHAUSTIVE) against the 737K list and the slide-seq
list, device-resident, and checks a sample of each against the oracle.  The kernel is the
bit-parallel one unless NR_EXHAUSTIVE_DPX is set (then the DPX kernels of round 1): run the script
once each way for the A/B figure.  usage: time_exhaustive.py [n_737k] [n_slideseq]"""
import gzip, os, sys
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, R); sys.path.insert(0, os.path.join(R, "tests"))
import numpy as np, torch
from helpers import compare
from nanoranger_b200 import Whitelist, NR_MODE_EXHAUSTIVE, synth, whitelists
from nanoranger_b200.whitelists import LINKER_SLIDESEQ
from oracle import oracle as O

which = "dpx" if os.environ.get("NR_EXHAUSTIVE_DPX") else "bit-parallel"
dev = torch.device("cuda:0")


def run(name, wl_a, pads, n, geometry, min_score, n_check):
    L = wl_a.shape[1]
    d = synth.make_candidates(wl_a, n, seed=9, geometry=geometry, p_n=2e-3)
    wl = Whitelist(wl_a, pads[0], pads[1], device=0)
    ds = torch.from_numpy(d["seqs"]).to(dev)
    do = torch.from_numpy(d["offsets"].view(np.int64).copy()).to(dev)
    b, m, nm = wl.pack_device(ds, do)
    ws = wl.workspace(n, dev, NR_MODE_EXHAUSTIVE)
    out = wl.match_device(b, m, nm, min_score=min_score, mode=NR_MODE_EXHAUSTIVE, workspace=ws)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    out = wl.match_device(b, m, nm, min_score=min_score, mode=NR_MODE_EXHAUSTIVE, workspace=ws)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    cells = float(m.to(torch.int64).bitwise_and(0x7F).sum().item()) * len(wl_a) * L * 2
    seqs = synth.to_strings(d["seqs"], d["offsets"])[:n_check]
    cc, cl = O.encode_many(seqs, 64)
    ref = O.match(O._CODE[wl_a], pads[0], pads[1], cc, cl)

    class Sub:
        pass
    sub = Sub()
    for k in ("idx", "score", "nbest", "flags", "umi_q"):
        setattr(sub, k, getattr(out, k)[:n_check].cpu().numpy())
    compare(ref, sub, min_score, exact_below=True, label=name)
    print(f"{which} {name}: {n} candidates x {len(wl_a)} entries x {L} columns: {ms:.1f} ms, "
          f"{n / ms * 1e3:.3e} cand/s, {cells / ms * 1e3:.3e} cells/s; first {n_check} equal the oracle", flush=True)
    wl.close()


n1 = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
n2 = int(sys.argv[2]) if len(sys.argv) > 2 else 32768
run("737K", whitelists.load_737k(), (30, 40), n1, "5p", 14, 96)
bcs = gzip.open(os.path.join(R, "tests", "golden", "slideseq_whitelist.txt.gz"), "rt").read().split()
cores = [x[:8] + LINKER_SLIDESEQ + x[8:] for x in bcs]
wl_s = np.frombuffer("".join(cores).encode(), np.uint8).reshape(len(cores), 32)
run("slide-seq", wl_s, (15, 24), n2, "slideseq", 30, 256)
