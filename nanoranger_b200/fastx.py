"""FASTA(.gz) reading for the matcher: `{sample}_BCUMI.fasta.gz` as written by the reference's
utils.decon_* (2-line records, possibly several gzip members concatenated by
pipeline.py:187-190) and `{sample}_bcreads.fasta` as written by utils.write_bc_*."""
from __future__ import annotations

import gzip

import numpy as np


def _read_bytes(path: str) -> bytes:
    with open(path, "rb") as f:
        head = f.read(2)
    if head == b"\x1f\x8b":
        with gzip.open(path, "rb") as f:      # handles multi-member files
            return f.read()
    with open(path, "rb") as f:
        return f.read()


def read_fasta(path: str):
    """-> (names list[str], seqs u8 buffer, offsets u64 [n+1]).  Names are cut at the first
    space (STAR --readNameSeparator space, scripts/barcode_align.sh:33)."""
    raw = _read_bytes(path)
    if not raw:
        return [], np.zeros(0, np.uint8), np.zeros(1, np.uint64)
    buf = np.frombuffer(raw, dtype=np.uint8)
    if buf[-1] != 10:
        buf = np.concatenate([buf, np.array([10], np.uint8)])
    nl = np.flatnonzero(buf == 10)
    starts = np.concatenate([[0], nl[:-1] + 1])
    ends = nl.copy()
    # strip \r
    cr = (ends > starts) & (buf[np.maximum(ends - 1, 0)] == 13)
    ends = ends - cr
    nonempty = ends > starts
    starts, ends = starts[nonempty], ends[nonempty]
    is_hdr = buf[starts] == ord(">")
    two_line = len(starts) % 2 == 0 and is_hdr[0::2].all() and not is_hdr[1::2].any()
    if two_line:
        hs, he = starts[0::2], ends[0::2]
        ss, se = starts[1::2], ends[1::2]
        lens = (se - ss).astype(np.uint64)
        offsets = np.zeros(len(ss) + 1, np.uint64)
        np.cumsum(lens, out=offsets[1:])
        keep = np.zeros(len(buf) + 1, np.int8)
        np.add.at(keep, ss, 1)
        np.add.at(keep, se, -1)
        seqs = buf[np.cumsum(keep[:-1]) > 0]
        names = [raw[a + 1:b].split(b" ", 1)[0].decode("ascii", "replace") for a, b in zip(hs, he)]
        return names, np.ascontiguousarray(seqs), offsets
    # general multi-line FASTA
    names, chunks, cur = [], [], []
    for a, b, h in zip(starts, ends, is_hdr):
        if h:
            if names:
                chunks.append(b"".join(cur))
            names.append(raw[a + 1:b].split(b" ", 1)[0].decode("ascii", "replace"))
            cur = []
        else:
            cur.append(raw[a:b])
    if names:
        chunks.append(b"".join(cur))
    lens = np.fromiter((len(c) for c in chunks), np.uint64, len(chunks))
    offsets = np.zeros(len(chunks) + 1, np.uint64)
    np.cumsum(lens, out=offsets[1:])
    return names, np.frombuffer(b"".join(chunks), np.uint8).copy(), offsets


def read_fasta_raw(path: str):
    """2-line FASTA(.gz) without any per-record Python work: -> (names u8 buffer, name offsets
    u64 [n+1], seqs u8 buffer, seq offsets u64 [n+1]); names are cut at the first space.  Falls
    back to read_fasta() for multi-line files."""
    raw = _read_bytes(path)
    if not raw:
        z = np.zeros(0, np.uint8)
        return z, np.zeros(1, np.uint64), z, np.zeros(1, np.uint64)
    buf = np.frombuffer(raw, dtype=np.uint8)
    if buf[-1] != 10:
        buf = np.concatenate([buf, np.array([10], np.uint8)])
    nl = np.flatnonzero(buf == 10)
    starts = np.concatenate([[0], nl[:-1] + 1])
    ends = nl.copy()
    ends = ends - ((ends > starts) & (buf[np.maximum(ends - 1, 0)] == 13))
    nonempty = ends > starts
    starts, ends = starts[nonempty], ends[nonempty]
    is_hdr = buf[starts] == ord(">")
    if not (len(starts) % 2 == 0 and is_hdr[0::2].all() and not is_hdr[1::2].any()):
        names, seqs, off = read_fasta(path)
        nb, no = _pack_names(names)
        return nb, no, seqs, off
    hs, he = starts[0::2] + 1, ends[0::2].copy()
    # cut names at the first space: position of the first space at or after the header start
    sp = np.flatnonzero(buf == 32)
    if len(sp):
        k = np.searchsorted(sp, hs)
        first = np.where(k < len(sp), sp[np.minimum(k, len(sp) - 1)], np.iinfo(np.int64).max)
        he = np.minimum(he, first)

    def gather(a, b):
        lens = (b - a).astype(np.uint64)
        off = np.zeros(len(a) + 1, np.uint64)
        np.cumsum(lens, out=off[1:])
        # +1 at every start, -1 at every end (starts and ends never coincide: a < b < next a)
        keep = np.zeros(len(buf) + 1, np.int8)
        keep[a] = 1
        keep[b] -= 1
        return np.ascontiguousarray(buf[np.cumsum(keep[:-1], dtype=np.int8) > 0]), off

    nb, no = gather(hs, he)
    sb, so = gather(starts[1::2], ends[1::2])
    return nb, no, sb, so


def _pack_names(names):
    b = [n.encode("ascii", "replace") for n in names]
    off = np.zeros(len(b) + 1, np.uint64)
    np.cumsum(np.fromiter((len(x) for x in b), np.uint64, len(b)), out=off[1:])
    return (np.frombuffer(b"".join(b), np.uint8).copy() if b else np.zeros(0, np.uint8)), off


def write_fasta(path: str, names, seqs, gz: bool | None = None) -> None:
    gz = path.endswith(".gz") if gz is None else gz
    op = gzip.open if gz else open
    with op(path, "wt") as f:
        for n, s in zip(names, seqs):
            f.write(f">{n}\n{s}\n")
