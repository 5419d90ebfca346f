import os, sys, time, gzip, tempfile, cProfile, pstats
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from nanoranger_b200 import synth, whitelists, utils, fastx
n = 1000000
wl = whitelists.load_737k()
d = synth.make_candidates(wl, n, seed=5)
seqs = synth.to_strings(d["seqs"], d["offsets"])
out = tempfile.mkdtemp()
with gzip.open(f"{out}/s_BCUMI.fasta.gz", "wt", compresslevel=1) as f:
    for i, s in enumerate(seqs):
        f.write(f">read{i:08d}-uuid_{i}_{i+500}_0_GENE{i%50}-201|ENST{i%50}.1_900\n{s}\n")
with open(f"{out}/wl.txt", "w") as f:
    f.write("\n".join(x + "-1" for x in whitelists.ascii_to_strings(wl)) + "\n")
utils.write_bc_5p10X("s", out, f"{out}/wl.txt")
utils.barcode_ref(f"{out}/s_bcreads.fasta", f"{out}/ref/")
for fn, args in ((utils.barcode_align, (f"{out}/s_BCUMI.fasta.gz", f"{out}/ref/", f"{out}/s_matching", 8)),
                 (utils.process_matching_5p10X, ("s", out))):
    pr = cProfile.Profile(); pr.enable(); fn(*args); pr.disable()
    st = pstats.Stats(pr); st.sort_stats("cumulative").print_stats(18)
