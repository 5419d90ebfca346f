"""Throughput of the adapter-search kernel (nr_hw_search_device) on device-resident windows."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from nanoranger_b200 import extract
n, L = int(sys.argv[1]) if len(sys.argv) > 1 else 2_000_000, int(sys.argv[2]) if len(sys.argv) > 2 else 100
rng = np.random.default_rng(1)
buf = np.frombuffer(b"ACGT", np.uint8)[rng.integers(0, 4, n * L)]
motif = "CGCTCTTCCGATCT" + 26 * "N" + "TTTCTTATATG"
inst = np.frombuffer(motif.replace("N", "A").encode(), np.uint8)
for i in range(0, n, 3):                       # plant the motif in a third of the windows
    a = i * L + 20
    buf[a:a + len(inst)] = inst
off = (np.arange(n + 1, dtype=np.int64) * L)
dev = torch.device("cuda:0")
d_buf, d_off = torch.from_numpy(buf).to(dev), torch.from_numpy(off).to(dev)
for pat, k, wild in ((motif, 6, True), ("AGATCGGAAGAGCGTCGTGT", 3, False)):
    r = extract.hw_search_device(d_buf, d_off, pat, k, wild); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); r = extract.hw_search_device(d_buf, d_off, pat, k, wild); e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    print(f"pattern {len(pat)} nt, k={k}: {n} windows x {L} nt in {ms:.2f} ms = {n / ms * 1e3:.3g} windows/s, "
          f"{n * L * len(pat) / ms / 1e6:.0f} GCUPS, {n * L / ms / 1e6:.1f} GB/s of text, hits {(r['ed'] >= 0).float().mean().item():.3f}")
