"""BAM reading / writing without pysam: just enough to do what the
reference does with the genome BAM after barcode assignment (utils.py:801-824):

    for read in samfile.fetch():                      # mapped records, file order
        if name in table and read.flag < 20:
            read.set_tag("CB", bc); read.set_tag("UB", umi); read.set_tag("XT", transcript)
            tagged_bam.write(read)

BGZF = concatenated gzip members with a 'BC' extra field (SAM spec section 4.1); BAM records are
copied byte for byte, only the auxiliary block is edited.  zlib does the (de)compression; this is
file-format plumbing around the hot path, not part of it.
"""
from __future__ import annotations

import struct
import zlib

_EOF = bytes.fromhex("1f8b08040000000000ff0600424302001b0003000000000000000000")
_BLOCK = 0xFF00            # uncompressed bytes per BGZF block (htslib's choice)
_FIXED = {"A": 1, "c": 1, "C": 1, "s": 2, "S": 2, "i": 4, "I": 4, "f": 4}


def _bgzf_blocks(raw: bytes):
    """yield the uncompressed payload of every BGZF block of a file image"""
    p, n = 0, len(raw)
    while p < n:
        if raw[p:p + 4] != b"\x1f\x8b\x08\x04":
            raise ValueError("not a BGZF stream (bad gzip member header)")
        xlen = struct.unpack_from("<H", raw, p + 10)[0]
        q, end, bsize = p + 12, p + 12 + xlen, None
        while q < end:
            si1, si2, slen = raw[q], raw[q + 1], struct.unpack_from("<H", raw, q + 2)[0]
            if si1 == 66 and si2 == 67 and slen == 2:
                bsize = struct.unpack_from("<H", raw, q + 4)[0]
            q += 4 + slen
        if bsize is None:
            raise ValueError("gzip member without BGZF 'BC' field")
        cdata = raw[end: p + bsize + 1 - 8]
        crc, isize = struct.unpack_from("<II", raw, p + bsize + 1 - 8)
        data = zlib.decompress(cdata, -15) if isize else b""
        if len(data) != isize or (zlib.crc32(data) & 0xFFFFFFFF) != crc:
            raise ValueError("BGZF block fails its CRC/length check")
        yield data
        p += bsize + 1


class BgzfWriter:
    def __init__(self, path: str, level: int = 6):
        self.f = open(path, "wb")
        self.level = level
        self.buf = bytearray()

    def _emit(self, chunk: bytes):
        co = zlib.compressobj(self.level, zlib.DEFLATED, -15)
        c = co.compress(chunk) + co.flush()
        bsize = 12 + 6 + len(c) + 8 - 1
        if bsize > 0xFFFF:                     # incompressible: store
            co = zlib.compressobj(0, zlib.DEFLATED, -15)
            c = co.compress(chunk) + co.flush()
            bsize = 12 + 6 + len(c) + 8 - 1
        self.f.write(b"\x1f\x8b\x08\x04\x00\x00\x00\x00\x00\xff\x06\x00BC\x02\x00" +
                     struct.pack("<H", bsize) + c +
                     struct.pack("<II", zlib.crc32(chunk) & 0xFFFFFFFF, len(chunk)))

    def write(self, b: bytes):
        self.buf += b
        while len(self.buf) >= _BLOCK:
            self._emit(bytes(self.buf[:_BLOCK]))
            del self.buf[:_BLOCK]

    def close(self):
        if self.buf:
            self._emit(bytes(self.buf))
            self.buf.clear()
        self.f.write(_EOF)
        self.f.close()


class BamReader:
    """Iterates the alignment records of a BAM file as raw bytes (without the block_size word)."""

    def __init__(self, path: str):
        with open(path, "rb") as f:
            raw = f.read()
        self.data = b"".join(_bgzf_blocks(raw))
        d = self.data
        if d[:4] != b"BAM\x01":
            raise ValueError(f"{path}: not a BAM file")
        l_text = struct.unpack_from("<i", d, 4)[0]
        self.text = d[8:8 + l_text]
        p = 8 + l_text
        n_ref = struct.unpack_from("<i", d, p)[0]
        p += 4
        self.refs = []
        for _ in range(n_ref):
            l_name = struct.unpack_from("<i", d, p)[0]
            name = d[p + 4:p + 4 + l_name - 1].decode()
            l_ref = struct.unpack_from("<i", d, p + 4 + l_name)[0]
            self.refs.append((name, l_ref))
            p += 8 + l_name
        self.header_bytes = d[:p]
        self._start = p

    def __iter__(self):
        d, p, n = self.data, self._start, len(self.data)
        while p + 4 <= n:
            bs = struct.unpack_from("<i", d, p)[0]
            yield d[p + 4:p + 4 + bs]
            p += 4 + bs


def rec_refid(rec: bytes) -> int:
    return struct.unpack_from("<i", rec, 0)[0]


def rec_flag(rec: bytes) -> int:
    return struct.unpack_from("<H", rec, 14)[0]


def rec_qname(rec: bytes) -> str:
    return rec[32:32 + rec[8] - 1].decode("ascii", "replace")


def _aux_start(rec: bytes) -> int:
    l_name, n_cig = rec[8], struct.unpack_from("<H", rec, 12)[0]
    l_seq = struct.unpack_from("<i", rec, 16)[0]
    return 32 + l_name + 4 * n_cig + (l_seq + 1) // 2 + l_seq


def aux_items(rec: bytes):
    """-> list of (tag str, type char, raw bytes of the whole field incl. tag+type)"""
    p, n, out = _aux_start(rec), len(rec), []
    while p + 3 <= n:
        tag, typ = rec[p:p + 2].decode(), chr(rec[p + 2])
        q = p + 3
        if typ in _FIXED:
            q += _FIXED[typ]
        elif typ in "ZH":
            q = rec.index(b"\x00", q) + 1
        elif typ == "B":
            sub, cnt = chr(rec[q]), struct.unpack_from("<i", rec, q + 1)[0]
            q += 5 + cnt * _FIXED[sub]
        else:
            raise ValueError(f"unknown BAM aux type {typ!r}")
        out.append((tag, typ, rec[p:q]))
        p = q
    return out


def get_tag(rec: bytes, tag: str):
    for t, typ, raw in aux_items(rec):
        if t == tag:
            if typ == "Z":
                return raw[3:-1].decode()
            if typ == "A":
                return chr(raw[3])
            fmt = {"c": "<b", "C": "<B", "s": "<h", "S": "<H", "i": "<i", "I": "<I", "f": "<f"}.get(typ)
            return struct.unpack_from(fmt, raw, 3)[0] if fmt else raw[3:]
    raise KeyError(tag)


def set_tags_z(rec: bytes, tags: dict) -> bytes:
    """pysam's read.set_tag(name, str) for every item: an existing field of that name is removed,
    the new Z field is appended."""
    a0 = _aux_start(rec)
    keep = b"".join(raw for t, _, raw in aux_items(rec) if t not in tags)
    new = b"".join(k.encode() + b"Z" + v.encode() + b"\x00" for k, v in tags.items())
    return rec[:a0] + keep + new


def write_bam(path: str, header_text: str, refs, records, level: int = 6):
    """records: iterable of raw record bytes (without block_size)."""
    w = BgzfWriter(path, level)
    t = header_text.encode()
    hb = b"BAM\x01" + struct.pack("<i", len(t)) + t + struct.pack("<i", len(refs))
    for name, ln in refs:
        nb = name.encode() + b"\x00"
        hb += struct.pack("<i", len(nb)) + nb + struct.pack("<i", ln)
    w.write(hb)
    for r in records:
        w.write(struct.pack("<i", len(r)) + r)
    w.close()


_NT16 = {c: i for i, c in enumerate("=ACMGRSVTWYHKDBN")}
_CIGOP = {c: i for i, c in enumerate("MIDNSHP=X")}


def make_record(qname, flag, ref_id, pos0, mapq, cigar, seq, qual=None, aux=b"") -> bytes:
    """Build one raw record (test helper and minimal SAM->BAM path)."""
    import re
    ops = [(int(n), _CIGOP[o]) for n, o in re.findall(r"(\d+)([MIDNSHP=X])", cigar)]
    ref_len = sum(n for n, o in ops if o in (0, 2, 3, 7, 8)) or 1
    end = pos0 + ref_len

    def reg2bin(beg, end):
        end -= 1
        for sh, off in ((14, 4681), (17, 585), (20, 73), (23, 9), (26, 1)):
            if beg >> sh == end >> sh:
                return off + (beg >> sh)
        return 0

    nb = qname.encode() + b"\x00"
    packed = bytearray((len(seq) + 1) // 2)
    for i, c in enumerate(seq):
        packed[i >> 1] |= _NT16.get(c.upper(), 15) << (4 if i % 2 == 0 else 0)
    q = bytes([0xFF] * len(seq)) if qual is None else bytes(ord(c) - 33 for c in qual)
    fixed = struct.pack("<iiBBHHHiiii", ref_id, pos0, len(nb), mapq, reg2bin(max(pos0, 0), max(end, 1)),
                        len(ops), flag, len(seq), -1, -1, 0)
    cig = b"".join(struct.pack("<I", (n << 4) | o) for n, o in ops)
    return fixed + nb + cig + bytes(packed) + q + aux


def tag_genome_bam(in_bam: str, out_bam: str, table: dict):
    """utils.py:801-824.  table: qname -> (CB, UB, XT).  Writes the records whose name is in the
    table and whose flag < 20, tagged, in file order; returns the list of XT values written
    (the reference's `all_trns`).  `samfile.fetch()` on an indexed BAM yields the records placed
    on a reference, so records with refID -1 are skipped."""
    rd = BamReader(in_bam)
    w = BgzfWriter(out_bam)
    w.write(rd.header_bytes)
    all_trns = []
    for rec in rd:
        if rec_refid(rec) < 0:
            continue
        hit = table.get(rec_qname(rec))
        if hit is not None and rec_flag(rec) < 20:
            new = set_tags_z(rec, {"CB": hit[0], "UB": hit[1], "XT": hit[2]})
            w.write(struct.pack("<i", len(new)) + new)
            all_trns.append(hit[2])
    w.close()
    return all_trns
