import os, sys, time, gzip, tempfile
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from nanoranger_b200 import synth, whitelists, utils, fastx
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1000000
wl = whitelists.load_737k()
d = synth.make_candidates(wl, n, seed=5)
seqs = synth.to_strings(d["seqs"], d["offsets"])
out = tempfile.mkdtemp()
t = time.time()
with gzip.open(f"{out}/s_BCUMI.fasta.gz", "wt", compresslevel=1) as f:
    for i, s in enumerate(seqs):
        f.write(f">read{i:08d}-uuid_{i}_{i+500}_0_GENE{i%50}-201|ENST{i%50}.1_900\n{s}\n")
print("write fasta", time.time() - t)
with open(f"{out}/wl.txt", "w") as f:
    f.write("\n".join(x + "-1" for x in whitelists.ascii_to_strings(wl)) + "\n")
t = time.time(); utils.write_bc_5p10X("s", out, f"{out}/wl.txt"); print("write_bc", time.time() - t)
t = time.time(); utils.barcode_ref(f"{out}/s_bcreads.fasta", f"{out}/ref/"); print("barcode_ref", time.time() - t)
t = time.time(); names, sq, off = fastx.read_fasta(f"{out}/s_BCUMI.fasta.gz"); print("read_fasta", time.time() - t)
t = time.time(); k = utils.barcode_align(f"{out}/s_BCUMI.fasta.gz", f"{out}/ref/", f"{out}/s_matching", 8); print("barcode_align", time.time() - t, k)
t = time.time(); utils.process_matching_5p10X("s", out); print("process_matching", time.time() - t)
