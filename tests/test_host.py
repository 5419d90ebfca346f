"""Host-side logic that needs no GPU: file formats, geometry, partitioning, generators."""
import gzip
import os

import numpy as np
import pytest


def test_read_fasta_two_line_multimember_gzip(tmp_path):
    from nanoranger_b200 import fastx
    p = tmp_path / "x_BCUMI.fasta.gz"
    with open(p, "wb") as f:                        # two gzip members, as `cat part_*` produces
        f.write(gzip.compress(b">r1_0_50_0_TRBV1 extra words\nACGTN\n>r2_1_2_16_X\nGG\n"))
        f.write(gzip.compress(b">r3\nTTTT\n"))
    names, seqs, off = fastx.read_fasta(str(p))
    assert names == ["r1_0_50_0_TRBV1", "r2_1_2_16_X", "r3"]
    assert seqs.tobytes() == b"ACGTNGGTTTT" and off.tolist() == [0, 5, 7, 11]


def test_read_fasta_multiline_and_empty(tmp_path):
    from nanoranger_b200 import fastx
    p = tmp_path / "m.fa"
    p.write_text(">a\nAC\nGT\n>b\n>c\nT\r\n")
    names, seqs, off = fastx.read_fasta(str(p))
    assert names == ["a", "b", "c"] and seqs.tobytes() == b"ACGTT" and off.tolist() == [0, 4, 4, 5]
    q = tmp_path / "e.fa"
    q.write_text("")
    assert fastx.read_fasta(str(q))[0] == []


def test_anchored_alignment_roundtrip():
    from nanoranger_b200 import samio
    for (m, u, pl, L, pr) in [(50, 30, 30, 16, 40), (50, 46, 30, 16, 40), (64, 60, 30, 16, 40),
                              (35, 19, 4, 16, 17), (35, 34, 4, 16, 17), (20, 0, 30, 16, 40),
                              (60, 15, 4, 16, 17)]:
        pos, cig = samio.anchored_alignment(m, u, pl, L, pr)
        assert samio.query_index_at(pos, cig, pl + L) == u, (m, u, cig)
        qlen = sum(int(n) for n, op in samio._CIG.findall(cig) if op in "MI")
        assert qlen == m and pos >= 1
    pos, cig = samio.anchored_alignment(20, -1, 30, 16, 40)
    assert samio.query_index_at(pos, cig, 46) is None           # the reference's `except` branch
    assert samio.query_index_at(pos, cig, 45) == 19


def test_sam_write_read(tmp_path):
    from nanoranger_b200 import samio
    p = str(tmp_path / "a.sam")
    recs = [("q1", 0, 2, 17, "50M", "A" * 50, 16), ("q2", 16, 0, 1, "3I47M", "C" * 50, 13)]
    assert samio.write_sam(p, ["b0", "b1", "b2"], 86, recs) == 2
    txt = open(p).read()
    assert "@SQ\tSN:b2\tLN:86" in txt and "@SQ\tSN:b1" not in txt
    back = samio.read_sam(p)
    assert [(r["qname"], r["flag"], r["rname"], r["AS"]) for r in back] == \
        [("q1", 0, "b2", 16), ("q2", 16, "b0", 13)]


def test_write_bc_and_barcode_ref(tmp_path):
    from nanoranger_b200 import utils
    wl = tmp_path / "wl.txt"
    wl.write_text("ACGTACGTACGTACGT-1\nTTTTCCCCGGGGAAAA-1\n")
    utils.write_bc_5p10X("s", str(tmp_path), str(wl))
    lines = open(tmp_path / "s_bcreads.fasta").read().split("\n")
    assert lines[0] == ">ACGTACGTACGTACGT" and lines[1] == "N" * 30 + "ACGTACGTACGTACGT" + "N" * 40
    g = utils.barcode_ref(str(tmp_path / "s_bcreads.fasta"), str(tmp_path / "s_ref"))
    d = np.load(os.path.join(g, "nr_whitelist.npz"))
    assert int(d["pad_l"]) == 30 and int(d["pad_r"]) == 40 and d["cores"].shape == (2, 16)
    assert bytes(d["cores"][1]) == b"TTTTCCCCGGGGAAAA" and str(d["names"][0]) == "ACGTACGTACGTACGT"


def test_write_bc_slideseq_geometry(tmp_path):
    from nanoranger_b200 import utils
    f = tmp_path / "x.matched.barcodes.tsv"
    f.write_text("TTTTTTTTAAAAAA-1\nACGTACGTNCGTAC-1\n")
    utils.write_bc_slideseq("s", str(tmp_path), str(f))
    lines = open(tmp_path / "s_bcreads.fasta").read().split("\n")
    assert lines[0] == ">ACGTACGTNCGTAC"                       # np.unique sorts
    assert lines[1] == "N" * 15 + "ACGTACGT" + "TCTTCAGCGTTCCCGAGA" + "NCGTAC" + "N" * 24
    utils.barcode_ref(str(tmp_path / "s_bcreads.fasta"), str(tmp_path / "ref"))
    d = np.load(tmp_path / "ref" / "nr_whitelist.npz")
    assert (int(d["pad_l"]), int(d["pad_r"]), d["cores"].shape[1]) == (15, 24, 32)


def test_write_bc_3p10XGEX(tmp_path):
    import json
    from nanoranger_b200 import utils
    os.makedirs(tmp_path / "split")
    json.dump({"A" * 16: 15, "C" * 16: 30}, open(tmp_path / "split" / "part_1_bc_count.json", "w"))
    json.dump({"A" * 16: 10, "G" * 16: 50}, open(tmp_path / "split" / "part_2_bc_count.json", "w"))
    wl = tmp_path / "3M.txt"
    wl.write_text("A" * 16 + "\n" + "C" * 16 + "\n" + "T" * 16 + "\n")
    utils.write_bc_3p10XGEX("s", str(tmp_path), str(wl))
    txt = open(tmp_path / "s_bcreads.fasta").read().split("\n")
    assert txt[0] == ">" + "A" * 16 and txt[1] == "NNNN" + "A" * 16 + "N" * 17     # 25 reads, listed
    assert txt[2] == ">" + "C" * 16 and len(txt) == 5                              # G*16 not whitelisted
    # byte-level: the pandas Series the reference writes (utils.py:1434-1436): barcode index, the
    # column named read_count, insertion order of the aggregated dict
    assert open(tmp_path / "s_bc_read_count.csv").read() == \
        ",read_count\n" + "A" * 16 + ",25\n" + "C" * 16 + ",30\n" + "G" * 16 + ",50\n"


def test_write_bc_3p10XGEX_order_and_verbatim_whitelist(tmp_path):
    """FASTA order = order of the counts (not of the whitelist); whitelist lines are compared
    verbatim (a `-1` suffix in the file makes nothing match, as in the reference: utils.py:1442-1447)."""
    import json
    from nanoranger_b200 import utils
    os.makedirs(tmp_path / "split")
    json.dump({"T" * 16: 40, "C" * 16: 30, "A" * 16: 21, "G" * 16: 20},
              open(tmp_path / "split" / "part_1_bc_count.json", "w"))
    wl = tmp_path / "3M.txt"
    wl.write_text("A" * 16 + "\n" + "C" * 16 + "\n" + "T" * 16 + "\n" + "G" * 16 + "\n")
    utils.write_bc_3p10XGEX("s", str(tmp_path), str(wl))
    names = [ln[1:] for ln in open(tmp_path / "s_bcreads.fasta").read().split("\n") if ln.startswith(">")]
    assert names == ["T" * 16, "C" * 16, "A" * 16]                 # > 20 reads, count order
    os.remove(tmp_path / "s_bcreads.fasta")
    wl.write_text("A" * 16 + "-1\n" + "C" * 16 + "-1\n")
    utils.write_bc_3p10XGEX("s", str(tmp_path), str(wl))
    assert open(tmp_path / "s_bcreads.fasta").read() == ""


def test_sort_cnt_and_pack_umis():
    from nanoranger_b200 import umi, utils
    df = utils.sort_cnt([16, 14, 16, 15, 16, 14])
    assert df.iloc[0].tolist() == [16, 3] and set(map(tuple, df.values.tolist())) == {(16, 3), (14, 2), (15, 1)}
    codes, ok = umi.pack_umis(["ACGT", "TTTT", "ANGT"], 4)
    assert ok.tolist() == [True, True, False] and codes[0] == 0b11100100 and codes[1] == 0xFF
    assert umi.unpack_umis(codes[:2], 4) == ["ACGT", "TTTT"]


def test_partition_and_shards():
    from nanoranger_b200 import umi
    rng = np.random.default_rng(0)
    bc = rng.integers(0, 5000, 20000).astype(np.uint32)
    own = umi.owner_rank(bc, 8)
    assert own.min() == 0 and own.max() == 7
    assert np.bincount(own, minlength=8).min() > 1500                    # balanced
    rec, counts = umi.partition_records(bc, bc * 0, bc * 3, 8)
    assert counts.sum() == len(bc) and rec.shape == (20000, 3)
    assert (np.diff(umi.owner_rank(rec[:, 0], 8)) >= 0).all()            # grouped by owner
    assert np.array_equal(rec[:, 2], rec[:, 0] * 3)
    cover = [umi.shard_bounds(1001, 8, r) for r in range(8)]
    assert cover[0][0] == 0 and cover[-1][1] == 1001
    assert all(cover[i][1] == cover[i + 1][0] for i in range(7))
    assert umi.shard_bounds(3, 8, 7) == (3, 3)


def test_synth_is_seeded_and_shaped():
    from nanoranger_b200 import synth, whitelists
    wl = whitelists.load_737k()
    assert wl.shape == (737280, 16) and bytes(wl[0]) == b"AAACCTGAGAAACCAT"
    a = synth.make_candidates(wl, 2000, seed=5)
    b = synth.make_candidates(wl, 2000, seed=5)
    assert np.array_equal(a["seqs"], b["seqs"]) and np.array_equal(a["offsets"], b["offsets"])
    lens = np.diff(a["offsets"].astype(np.int64))
    assert lens.max() <= 64 and 49 < lens.mean() < 55
    assert 0.05 < (a["true_idx"] < 0).mean() < 0.15
    clean = synth.make_candidates(wl, 50, seed=1, p_sub=0, p_ins=0, p_del=0, frac_negative=0)
    s = synth.to_strings(clean["seqs"], clean["offsets"])
    for i, q in enumerate(s):
        assert q[:14] == "CGCTCTTCCGATCT" and q[14:30] == bytes(wl[clean["true_idx"][i]]).decode()
    sw = whitelists.synthetic_whitelist(5000, seed=3)
    assert len({bytes(r) for r in sw}) == 5000


def test_sam_alignment_attributes_follow_pysam(tmp_path):
    """extract.Aln: the pysam attributes the reference's extractors read (utils.py:112-127)."""
    from nanoranger_b200 import extract
    seq = "A" * 10 + "C" * 50 + "G" * 7
    sam = ("@HD\tVN:1.6\n@SQ\tSN:T1\tLN:500\n"
           f"r1\t0\tT1\t101\t60\t10S30=2X5I10=3D3=7S\t*\t0\t0\t{seq}\t{'I' * len(seq)}\tNM:i:10\tAS:i:77\ttp:A:P\n"
           f"r2\t2064\tT1\t5\t60\t5H20=5H\t*\t0\t0\t{'T' * 20}\t{'#' * 20}\tAS:i:40\n"
           "r3\t4\t*\t0\t0\t*\t*\t0\t0\tACGT\tIIII\n")
    p = tmp_path / "x_trns.sam"
    p.write_text(sam)
    recs = extract.read_alignments(str(p))
    assert [r.qname for r in recs] == ["r1", "r2"]                    # unmapped record skipped
    a = recs[0]
    assert (a.query_alignment_start, a.query_alignment_end, a.qlen, a.rlen) == (10, 60, 50, 67)
    assert (a.reference_start, a.reference_end) == (100, 100 + 30 + 2 + 10 + 3 + 3)
    assert a.get_tag("AS") == 77 and a.get_tag("tp") == "P" and a.flag == 0
    with pytest.raises(KeyError):
        a.get_tag("XX")
    b = recs[1]                                                       # hard clips are not in SEQ
    assert (b.query_alignment_start, b.query_alignment_end, b.rlen) == (0, 20, 20)
    assert extract.rev("AACGTN") == "NACGTT"
    assert extract._mod_coords(b) == (0, 20)


def test_native_sam_writer_equals_python_writer(tmp_path):
    """nr_sam_write (C++) writes the bytes samio.write_sam(match_records-style records) writes."""
    import ctypes as C
    from nanoranger_b200 import _lib, fastx, samio
    from nanoranger_b200._lib import NR_FLAG_RC, NR_FLAG_TOO_LONG, NR_UMI_NONE
    rng = np.random.default_rng(12)
    n, n_ref, pad_l, L, pad_r = 3000, 40, 30, 16, 40
    ref_names = ["".join("ACGT"[i] for i in rng.integers(0, 4, 16)) for _ in range(n_ref)]
    seqs = ["".join("ACGTN"[i] for i in rng.integers(0, 5, int(rng.integers(0, 65)))) for _ in range(n)]
    names = [f"read{i}-x_{i}_{i + 7}_0_G{i % 3}|T.1_9" for i in range(n)]
    idx = rng.integers(0, n_ref, n).astype(np.int32)
    score = rng.integers(-5, 17, n).astype(np.int8)
    nbest = rng.choice([0, 1, 1, 1, 2], n).astype(np.uint8)
    flags = (rng.choice([0, 0, 0, NR_FLAG_RC, NR_FLAG_TOO_LONG], n)).astype(np.uint8)
    umi_q = np.where(rng.random(n) < 0.2, NR_UMI_NONE, rng.integers(0, 64, n)).astype(np.uint8)
    recs = []
    for i in range(n):
        if nbest[i] == 1 and not flags[i] & NR_FLAG_TOO_LONG:
            rc = bool(flags[i] & NR_FLAG_RC)
            u = int(umi_q[i])
            pos, cig = samio.anchored_alignment(len(seqs[i]), -1 if (u == NR_UMI_NONE or rc) else u, pad_l, L, pad_r)
            s = samio.revcomp_bytes(seqs[i].encode()).decode() if rc else seqs[i]
            recs.append((names[i], 16 if rc else 0, int(idx[i]), pos, cig, s, int(score[i])))
    for header in ("used", "full"):
        samio.write_sam(str(tmp_path / "py.sam"), ref_names, pad_l + L + pad_r, recs, header=header)
        nb, no = fastx._pack_names(names)
        sb, so = fastx._pack_names(seqs)
        rb, ro = fastx._pack_names(ref_names)
        w = C.c_uint64()
        rc_ = _lib.lib().nr_sam_write(str(tmp_path / "c.sam").encode(), header == "full", nb.ctypes.data,
                                      no.ctypes.data, sb.ctypes.data, so.ctypes.data, n, idx.ctypes.data,
                                      score.ctypes.data, nbest.ctypes.data, flags.ctypes.data,
                                      umi_q.ctypes.data, rb.ctypes.data, ro.ctypes.data, n_ref, pad_l, L,
                                      pad_r, C.byref(w))
        assert rc_ == 0 and w.value == len(recs)
        assert (tmp_path / "c.sam").read_bytes() == (tmp_path / "py.sam").read_bytes()
    # the raw FASTA reader agrees with the record reader
    fa = tmp_path / "x.fa.gz"
    with gzip.open(fa, "wt") as f:
        for nm, s in zip(names, seqs):
            f.write(f">{nm} trailing words\n{s}\n" if s else f">{nm}\nN\n")
    a_names, a_seq, a_off = fastx.read_fasta(str(fa))
    nb, no, sb, so = fastx.read_fasta_raw(str(fa))
    assert np.array_equal(a_seq, sb) and np.array_equal(a_off, so)
    raw = nb.tobytes().decode()
    assert [raw[int(no[i]):int(no[i + 1])] for i in range(len(no) - 1)] == a_names


def test_sam_table_reader_equals_record_reader(tmp_path):
    from nanoranger_b200 import samio
    recs = [("a_1_2_0_T", 0, 0, 17, "3I40M2I", "ACGT" * 11 + "A", 15), ("b", 16, 1, 1, "50M", "C" * 50, 14),
            ("c x", 0, 1, 31, "12M", "G" * 12, -3)]
    p = str(tmp_path / "m.sam")
    samio.write_sam(p, ["BC1", "BC2"], 86, recs)
    t = samio.read_sam_table(p)
    r = samio.read_sam(p)
    assert [x["qname"] for x in r] == list(t["qname"]) and [x["AS"] for x in r] == list(t["AS"])
    assert [x["cigar"] for x in r] == list(t["cigar"]) and [x["pos"] for x in r] == list(t["pos"])
    assert [x["seq"] for x in r] == list(t["seq"]) and [x["rname"] for x in r] == list(t["rname"])
    for col in (0, 16, 20, 45, 46, 60):
        many = samio.query_index_at_many(t["pos"], t["cigar"], col)
        one = [samio.query_index_at(x["pos"], x["cigar"], col) for x in r]
        assert list(many) == [-1 if v is None else v for v in one]
    # STAR-style line: AS not at the same column as ours, complex CIGAR
    with open(p, "a") as f:
        f.write("d\t0\tBC1\t5\t255\t10M2D5M1I3M\t*\t0\t0\t" + "T" * 19 + "\t*\tAS:i:12\tnM:i:1\tMD:Z:10^AC8\n")
    t = samio.read_sam_table(p)
    assert list(t["AS"]) == [15, 14, -3, 12] and t["cigar"][3] == "10M2D5M1I3M"
    assert samio.query_index_at_many(t["pos"][3:], t["cigar"][3:], 17)[0] == samio.query_index_at(5, "10M2D5M1I3M", 17)
    empty = str(tmp_path / "e.sam")
    samio.write_sam(empty, ["BC1"], 86, [])
    assert len(samio.read_sam_table(empty)["AS"]) == 0
