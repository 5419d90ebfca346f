"""BAM I/O without pysam (nanoranger_b200/bamio.py): BGZF framing cross-checked with Python's gzip
module, record layout round trips, and the CB/UB/XT tagging of utils.py:801-824."""
import gzip
import struct

import numpy as np

from nanoranger_b200 import bamio


def _records(rng, n):
    recs, names = [], []
    for i in range(n):
        L = int(rng.integers(1, 400))
        seq = "".join("ACGTN"[j] for j in rng.integers(0, 5, L))
        qual = "".join(chr(33 + int(x)) for x in rng.integers(0, 60, L))
        name = f"read{i}_{i*3}_{i*3+L}_{[0, 16][i % 2]}_GENE{i % 5}-201|ENST{i % 5}.1_900"
        flag = [0, 16, 256, 2048, 4][i % 5]
        ref = -1 if flag == 4 else i % 2
        aux = b"NMi" + struct.pack("<i", i) + b"XTZold\x00" + b"MDZ" + f"{L}".encode() + b"\x00"
        if i % 3 == 0:
            aux += b"B0BS" + struct.pack("<i", 3) + struct.pack("<HHH", 1, 2, 3)
        recs.append(bamio.make_record(name, flag, ref, int(rng.integers(0, 1 << 20)), 60, f"{L}M", seq, qual, aux))
        names.append(name)
    return recs, names


def test_bgzf_and_record_roundtrip(tmp_path):
    rng = np.random.default_rng(1)
    recs, names = _records(rng, 3000)                       # > 1 BGZF block
    text = "@HD\tVN:1.6\tSO:coordinate\n@SQ\tSN:chr1\tLN:248956422\n@SQ\tSN:chrM\tLN:16569\n"
    p = str(tmp_path / "a.bam")
    bamio.write_bam(p, text, [("chr1", 248956422), ("chrM", 16569)], recs)
    raw = open(p, "rb").read()
    assert raw.endswith(bamio._EOF) and raw[:4] == b"\x1f\x8b\x08\x04"
    # an independent gzip implementation reads the same byte stream
    plain = gzip.open(p, "rb").read()
    rd = bamio.BamReader(p)
    assert plain == rd.data and plain[:4] == b"BAM\x01"
    assert rd.text.decode() == text and rd.refs == [("chr1", 248956422), ("chrM", 16569)]
    got = list(rd)
    assert got == recs
    assert [bamio.rec_qname(r) for r in got] == names
    assert bamio.rec_flag(got[1]) == 16 and bamio.rec_refid(got[4]) == -1
    assert bamio.get_tag(got[7], "NM") == 7 and bamio.get_tag(got[7], "XT") == "old"
    assert [t for t, _, _ in bamio.aux_items(got[0])] == ["NM", "XT", "MD", "B0"]


def test_tag_genome_bam(tmp_path):
    rng = np.random.default_rng(2)
    recs, names = _records(rng, 500)
    p, q = str(tmp_path / "s_genome.bam"), str(tmp_path / "s_genome_tagged.bam")
    bamio.write_bam(p, "@HD\tVN:1.6\n@SQ\tSN:chr1\tLN:1000\n@SQ\tSN:chrM\tLN:16569\n",
                    [("chr1", 1000), ("chrM", 16569)], recs)
    table = {n: ("ACGTACGTACGTACGT", "TTTTGGGGCC", n.split("_")[4]) for n in names[::2]}
    trns = bamio.tag_genome_bam(p, q, table)
    out = list(bamio.BamReader(q))
    # reference semantics: in the table, flag < 20, placed on a reference; file order kept
    exp = [i for i in range(0, 500, 2) if [0, 16, 256, 2048, 4][i % 5] < 20 and [0, 16, 256, 2048, 4][i % 5] != 4]
    assert [bamio.rec_qname(r) for r in out] == [names[i] for i in exp]
    assert trns == [names[i].split("_")[4] for i in exp]
    for r, i in zip(out, exp):
        assert bamio.get_tag(r, "CB") == "ACGTACGTACGTACGT" and bamio.get_tag(r, "UB") == "TTTTGGGGCC"
        assert bamio.get_tag(r, "XT") == names[i].split("_")[4]          # old XT replaced, not duplicated
        tags = [t for t, _, _ in bamio.aux_items(r)]
        assert tags.count("XT") == 1 and tags[-3:] == ["CB", "UB", "XT"]
        a0 = bamio._aux_start(r)
        assert r[:a0] == recs[i][:a0]                                       # fixed part untouched
    assert bamio.BamReader(q).header_bytes == bamio.BamReader(p).header_bytes
