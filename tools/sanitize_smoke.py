"""Small run of every kernel for compute-sanitizer (memcheck / racecheck / initcheck)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from nanoranger_b200 import NR_MODE_AUTO, NR_MODE_FILTERED, NR_MODE_EXHAUSTIVE, Whitelist, synth, whitelists, pack_ascii, extract
from nanoranger_b200 import umi as U
wl_a = whitelists.load_737k()[::8]                       # 92K entries: quick index, same code paths
d = synth.make_candidates(wl_a, 3000, seed=3)
seqs = synth.to_strings(d["seqs"], d["offsets"])
seqs[5] = seqs[5][:10] + "N" + seqs[5][11:]
seqs += ["ACGT" * 20, "", "ACGTACGTAC"]
wl = Whitelist(wl_a, 30, 40)
dev = torch.device("cuda:0")
buf, off = pack_ascii(seqs)
d_seqs = torch.from_numpy(buf.copy()).to(dev); d_off = torch.from_numpy(off.view(np.int64).copy()).to(dev)
bases, meta, nmask = wl.pack_device(d_seqs, d_off)
for mode in (NR_MODE_FILTERED, NR_MODE_AUTO):
    res = wl.match_device(bases, meta, nmask, min_score=14, mode=mode)
res_c = wl.match_device(bases, meta, nmask, min_score=14, counted=True)
ex = wl.match_device(bases[:40], meta[:40], nmask[:40], min_score=14, mode=NR_MODE_EXHAUSTIVE)
h = wl.match_host(seqs, min_score=14, mode=NR_MODE_FILTERED)
# 32-column cores with N columns: anchored filter, then the bit-parallel brute force (<32, N>);
# a ragged whitelist (n % 32 != 0) and a batch smaller than the grid (whitelist slicing)
import gzip
from nanoranger_b200.whitelists import LINKER_SLIDESEQ
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
bcs = gzip.open(os.path.join(R, "tests", "golden", "slideseq_whitelist.txt.gz"), "rt").read().split()[:3001]
wl_s = np.frombuffer("".join(b[:8] + LINKER_SLIDESEQ + b[8:] for b in bcs).encode(), np.uint8).reshape(len(bcs), 32)
ds = synth.make_candidates(wl_s, 500, seed=5, geometry="slideseq", p_n=0.01)
sseqs = synth.to_strings(ds["seqs"], ds["offsets"])
wls = Whitelist(wl_s, 15, 24)
for mode in (NR_MODE_AUTO, NR_MODE_EXHAUSTIVE):
    wls.match_host(sseqs, min_score=30, mode=mode)
wls.match_host(sseqs[:7], min_score=30, mode=NR_MODE_EXHAUSTIVE)
wls.close()
gene = torch.arange(len(seqs), dtype=torch.int32, device=dev) % 5
rec = U.records_device(bases, meta, nmask, res, 14, 12, gene=gene)
rows, counts = U.partition_device(rec["bc"], rec["gene"], rec["umi"], 4)
U.unzip_device(rows)
rng = np.random.default_rng(2)
n = 40000
bc = np.where(rng.random(n) < 0.6, 3, rng.integers(0, 40, n)).astype(np.uint32)
um = rng.integers(0, 1 << 13, n).astype(np.uint32)
for md in (0, 1):
    U.collapse_host(bc, np.zeros(n, np.uint32), um, 12, md)
extract.hw_search(seqs, "CGCTCTTCCGATCT" + 26 * "N" + "TTTCTTATATG", 6, True)
torch.cuda.synchronize()
print("sanitize_smoke ok", int(res.assigned(14).sum()), rec["n_records"])
