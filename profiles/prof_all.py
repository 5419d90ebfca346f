#!/usr/bin/env python
"""One pass over every kernel with a throughput claim, on the workload that claim is quoted on
(run under `ncu --set full` by profiles/capture_r2.sh; prints one `PHASE ...` line per phase so the
launch list can be read against it).  No timing here: numbers under a profiler are not bench values."""
import os
import sys

R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, R)
sys.path.insert(0, os.path.join(R, "tests"))
import numpy as np
import torch

from nanoranger_b200 import (NR_MODE_AUTO, NR_MODE_EXHAUSTIVE, NR_MODE_FILTERED, Whitelist, extract,
                             pack_ascii, synth, whitelists)
from nanoranger_b200 import umi as U
from nanoranger_b200.whitelists import LINKER_SLIDESEQ

dev = torch.device("cuda:0")
B = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 22


def to_dev(d):
    return torch.from_numpy(d["seqs"]).to(dev), torch.from_numpy(d["offsets"].view(np.int64).copy()).to(dev)


# 1. the bench step: 4 Mi flanks with N against 737K, FILTERED (pack, main pass, N pass), counted
wl_a = whitelists.load_737k()
wl = Whitelist(wl_a, 30, 40)
d = synth.make_candidates(wl_a, B, seed=2, p_n=1e-3)
d_seqs, d_off = to_dev(d)
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
# 2. AUTO on the first Mi: deep tier K=3, K=5, finaliser
n2 = min(B, 1 << 20)
off2 = d_off[:n2 + 1].contiguous()
b2, m2, nm2 = wl.pack_device(d_seqs, off2)
print("PHASE flanks-auto", n2, flush=True)
wl.match_device(b2, m2, nm2, min_score=14, mode=NR_MODE_AUTO, workspace=ws)
torch.cuda.synchronize()
print("TIERS", wl.tier_counts(ws), flush=True)
# 3. brute-force DP, 1024 candidates
n3 = 1024
off3 = d_off[:n3 + 1].contiguous()
b3, m3, nm3 = wl.pack_device(d_seqs, off3)
print("PHASE flanks-bruteforce", n3, flush=True)
wl.match_device(b3, m3, nm3, min_score=14, mode=NR_MODE_EXHAUSTIVE)
torch.cuda.synchronize()
wl.close()
del d_seqs, d_off, ws, out

# 4. slide-seq geometry (config 2): 17 752 x 32 columns with N
from helpers import mutate, rs  # noqa: E402
rng = np.random.default_rng(5)
bcs = sorted({rs(rng, 14) for _ in range(17753)})
bcs = [b if rng.random() > 0.15 else b[:3] + "N" + b[4:] for b in bcs]
cores = [b[:8] + LINKER_SLIDESEQ + b[8:] for b in bcs]
seqs = []
for _ in range(20000):
    cc = cores[rng.integers(0, len(cores))].replace("N", "ACGT"[rng.integers(0, 4)])
    seqs.append((rs(rng, int(rng.integers(0, 18))) + mutate(rng, cc, int(rng.integers(0, 3))) + rs(rng, 30))[:int(rng.integers(46, 57))])
wls = Whitelist(cores, 15, 24)
buf, off = pack_ascii(seqs)
ds, do = torch.from_numpy(buf.copy()).to(dev), torch.from_numpy(off.view(np.int64).copy()).to(dev)
bs, ms_, ns = wls.pack_device(ds, do)
print("PHASE slideseq-auto", len(seqs), flush=True)
wss = wls.workspace(len(seqs), dev)
wls.match_device(bs, ms_, ns, min_score=30, mode=NR_MODE_AUTO, workspace=wss)
torch.cuda.synchronize()
print("TIERS", wls.tier_counts(wss), flush=True)
print("PHASE slideseq-bruteforce", 2048, flush=True)
o4 = do[:2049].contiguous()
b4, m4, n4 = wls.pack_device(ds, o4)
wls.match_device(b4, m4, n4, min_score=30, mode=NR_MODE_EXHAUSTIVE)
torch.cuda.synchronize()
wls.close()

# 5. kinnex (config 5): records + UMI collapse (max_dist 1) of 1 Mi sub-reads, 10 K cells
n_cells, n_genes = 10000, 20000
r0 = np.random.Generator(np.random.PCG64(20180201))
cells = np.sort(r0.choice(len(wl_a), n_cells, replace=False))
wlk_a = wl_a[cells]
wlk = Whitelist(wlk_a, 4, 17)
Bk = 1 << 20
r2 = np.random.Generator(np.random.PCG64(2))
w = 1.0 / np.arange(1, n_cells + 1) ** 0.8
cell = r2.choice(n_cells, size=Bk, p=w / w.sum())
gene = (r2.zipf(1.4, Bk) % n_genes).astype(np.uint32)
mol = r2.integers(0, 6, Bk).astype(np.uint64)
h = (cell.astype(np.uint64) * np.uint64(0x9E3779B97F4A7C15) ^ gene.astype(np.uint64) * np.uint64(0xC2B2AE3D27D4EB4F)
     ^ mol * np.uint64(0x165667B19E3779F9))
h ^= h >> np.uint64(29)
umi_codes = np.stack([((h >> np.uint64(2 * k + 7)) & np.uint64(3)).astype(np.uint8) for k in range(12)], 1)
dk = synth.make_candidates(wlk_a, Bk, seed=9, geometry="3p", umi_len=12, cell_idx=cell, umi_codes=umi_codes)
ks, ko = to_dev(dk)
kg = torch.from_numpy(gene.view(np.int32)).to(dev)
print("PHASE kinnex", Bk, flush=True)
kb, km, kn = wlk.pack_device(ks, ko)
res = wlk.match_device(kb, km, kn, min_score=14, mode=NR_MODE_FILTERED)
rec = U.records_device(kb, km, kn, res, 14, 12, gene=kg, with_src=False)
rows, counts = U.partition_device(rec["bc"], rec["gene"], rec["umi"], 8)
r = U.collapse_device(rec["bc"], rec["gene"], rec["umi"], 12, 1, bc_bits=U.key_bits(n_cells), gene_bits=U.key_bits(n_genes),
                      umi_bits=24)
torch.cuda.synchronize()
print("KINNEX records", rec["n_records"], "molecules", r["n_groups"], flush=True)
wlk.close()

# 6. adapter search (decon_*): 1 Mi windows of 100 nt, the 5' motif with N wildcards, k = 6
nw, Lw = 1 << 20, 100
bufw = np.frombuffer(b"ACGT", np.uint8)[np.random.default_rng(1).integers(0, 4, nw * Lw)]
offw = np.arange(nw + 1, dtype=np.int64) * Lw
print("PHASE hwsearch", nw, flush=True)
extract.hw_search_device(torch.from_numpy(bufw).to(dev), torch.from_numpy(offw).to(dev),
                         "CGCTCTTCCGATCT" + 26 * "N" + "TTTCTTATATG", 6, True)
torch.cuda.synchronize()
print("PHASE done", flush=True)
