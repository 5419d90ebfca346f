"""UMI collapse alone, at the group depth one rank owns in an N-rank kinnex run, on ONE GPU.
The synthetic library of bench.py (10 000 cells, zipf genes, <= 6 molecules per (cell, gene)) is
drawn `depth` times as deep, a substitution error lands in `p_err` of the UMIs, and the records of
the cells rank 0 of `depth` ranks would own are collapsed (max_dist 1) and, with --check, compared
with the CPU oracle's sequential walk.
  python tools/time_umi.py [depth] [--check] [--keyed]"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
from nanoranger_b200 import umi as U

depth = int(sys.argv[1]) if len(sys.argv) > 1 and sys.argv[1].isdigit() else 8
check = "--check" in sys.argv
B = 3_000_000 * depth
n_cells, n_genes = 10000, 20000
rng = np.random.Generator(np.random.PCG64(11))
w = 1.0 / np.arange(1, n_cells + 1) ** 0.8
cell = rng.choice(n_cells, size=B, p=w / w.sum()).astype(np.uint32)
own = U.owner_rank(cell, depth) == 0
cell = cell[own]
n = len(cell)
gene = (rng.zipf(1.4, n) % n_genes).astype(np.uint32)
mol = rng.integers(0, 6, n).astype(np.uint64)
h = (cell.astype(np.uint64) * np.uint64(0x9E3779B97F4A7C15) ^ gene.astype(np.uint64) * np.uint64(0xC2B2AE3D27D4EB4F)
     ^ mol * np.uint64(0x165667B19E3779F9))
h ^= h >> np.uint64(29)
umi = ((h >> np.uint64(7)) & np.uint64((1 << 24) - 1)).astype(np.uint32)
p_err = 0.35
for _ in range(2):                                # one or two substitutions
    flip = rng.random(n) < p_err
    umi = np.where(flip, umi ^ (rng.integers(1, 4, n).astype(np.uint32) << (2 * rng.integers(0, 12, n)).astype(np.uint32)), umi)
    p_err *= 0.3
umi = umi.astype(np.uint32)
key = (cell.astype(np.uint64) << np.uint64(32)) | gene
_, cnt = np.unique(key, return_counts=True)
print(f"depth {depth}: {n} records, {len(cnt)} (cell, gene) groups, deepest {cnt.max()} records", flush=True)

dev = torch.device("cuda:0")
d = [torch.from_numpy(a.view(np.int32)).to(dev) for a in (cell, gene, umi)]
widths = dict(bc_bits=U.key_bits(n_cells), gene_bits=U.key_bits(n_genes), umi_bits=24) if "--keyed" in sys.argv else {}
for md in (1, 0):
    r = U.collapse_device(*d, 12, md, **widths)
    torch.cuda.synchronize()
    ts = []
    for _ in range(20):
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
        ev[0].record()
        r = U.collapse_device(*d, 12, md, **widths)
        ev[1].record()
        torch.cuda.synchronize()
        ts.append(ev[0].elapsed_time(ev[1]))
    print(f"max_dist {md}: median {np.median(ts):.3f} ms, min {min(ts):.3f} ms per collapse (20 runs), "
          f"{r['n_groups']} molecules", flush=True)
    if check and md == 1:
        sys.path.insert(0, ROOT)
        from oracle import oracle as O
        t0 = time.time()
        k, rep = O.umi_cluster(cell, gene, umi, 1)
        assert k == r["n_groups"], (k, r["n_groups"])
        assert np.array_equal(rep, r["rep_umi"].cpu().numpy().view(np.uint32))
        print(f"oracle agrees ({time.time() - t0:.1f} s on the CPU)", flush=True)
