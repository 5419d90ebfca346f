"""The whole stage on synthetic reads, file to file: minimap2-style SAM -> decon_5p10XGEX ->
write_bc_5p10X -> barcode_ref -> barcode_align -> process_matching_5p10X (+ genome BAM tagging).
Error-free reads must come out with exactly the barcode, UMI and transcript that were planted."""
import gzip

import numpy as np
import pandas as pd
import pytest

pytestmark = pytest.mark.gpu


def _rs(rng, n):
    return "".join("ACGT"[i] for i in rng.integers(0, 4, n))


def test_5p10XGEX_chain_recovers_planted_barcodes(cuda_device, tmp_path):
    from nanoranger_b200 import bamio, extract, utils, whitelists
    rng = np.random.default_rng(2024)
    out = str(tmp_path)
    wl_names = whitelists.ascii_to_strings(whitelists.load_737k())
    with open(f"{out}/wl.txt", "w") as f:
        f.write("\n".join(n + "-1" for n in wl_names) + "\n")
    trans = ["NRAS-201|ENST00000369535.5_4326", "MT-CO1-201|ENST00000361624.2_1542"]
    lines = ["@HD\tVN:1.6\n"] + [f"@SQ\tSN:{t}\tLN:5000\n" for t in trans]
    planted, bam_recs = {}, []
    n = 400
    for i in range(n):
        bc = wl_names[int(rng.integers(0, len(wl_names)))]
        umi = _rs(rng, 10)
        flank = "CGCTCTTCCGATCT" + bc + umi + "TTTCTTATATG"
        junk, body, tail = _rs(rng, int(rng.integers(0, 120))), _rs(rng, int(rng.integers(60, 300))), _rs(rng, int(rng.integers(0, 30)))
        seq = junk + flank + body + tail
        lead = len(junk) + len(flank)
        t = trans[i % 2]
        cigar = f"{lead}S{len(body)}=" + (f"{len(tail)}S" if tail else "")
        qname = f"{i:08x}-aaaa-bbbb-cccc-{i:012x}"                  # ONT read ids: hyphens, no underscores
        lines.append(f"{qname}\t0\t{t}\t{100 + i}\t60\t{cigar}\t*\t0\t0\t{seq}\t{'I' * len(seq)}\tAS:i:{len(body)}\n")
        cand_name = f"{qname}_{lead}_{lead + len(body)}_0_{t}"
        planted[cand_name] = (bc, umi, t.split("_")[0])
        bam_recs.append(bamio.make_record(cand_name, 0, 0, 1000 + i, 60, f"{len(body)}M", body))
    with open(f"{out}/s_trns.sam", "w") as f:
        f.writelines(lines)
    bamio.write_bam(f"{out}/s_genome.bam", "@HD\tVN:1.6\n@SQ\tSN:chr1\tLN:248956422\n", [("chr1", 248956422)], bam_recs)

    assert extract.decon_5p10XGEX("s", out) == n                     # every read has its motif
    fa = gzip.open(f"{out}/s_BCUMI.fasta.gz", "rt").read().split("\n")
    assert fa[0][1:] in planted and len(fa[1]) == 50                 # adapter + bc + umi + TSO minus its last base
    utils.write_bc_5p10X("s", out, f"{out}/wl.txt")
    utils.barcode_ref(f"{out}/s_bcreads.fasta", f"{out}/s_matching/")
    utils.barcode_align(f"{out}/s_BCUMI.fasta.gz", f"{out}/s_matching/", f"{out}/s_matching", 8, "-1")
    table = utils.process_matching_5p10X("s", out)
    # error-free reads: assigned unless another barcode is as close somewhere in the 50 nt (rare)
    assert len(table) > 0.95 * n
    for name, (cb, ub, xt) in table.items():
        assert (cb, ub, xt) == planted[name]
    tagged = list(bamio.BamReader(f"{out}/s_genome_tagged.bam"))
    assert len(tagged) == len(table)
    for r in tagged:
        cb, ub, xt = planted[bamio.rec_qname(r)]
        assert (bamio.get_tag(r, "CB"), bamio.get_tag(r, "UB"), bamio.get_tag(r, "XT")) == (cb, ub, xt)
    ct = pd.read_csv(f"{out}/s_trns_ct.csv")
    assert set(ct.iloc[:, 0]) == {"NRAS-201|ENST00000369535.5", "MT-CO1-201|ENST00000361624.2"}
    ded = pd.read_csv(f"{out}/s_bcumi_dedup.csv", index_col=0)
    assert int(ded.read_cnt.sum()) == len(table) and (ded.dup_rate >= 1).all()
