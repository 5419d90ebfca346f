"""Seeded synthetic barcode/UMI candidates with an ONT-like error profile.

Templates follow what the reference's extractors emit:
  5'  : CGCTCTTCCGATCT + barcode(16) + UMI(umi_len) + TTTCTTATAT      (utils.py:105, 137-139, 202)
  3'  : 3 adapter bases + barcode(16) + UMI(12) + TTTT                 (utils.py:1374-1376)
  slide-seq : 8 nt + bc[:8] + linker(18) + bc[8:] + 16 nt               (utils.py:443-448, 584-601)
Per-base iid errors: substitution, insertion, deletion (default 2 % each, 6 % total); a fraction
of candidates carries a random non-whitelist 16-mer (negatives).  Fully vectorised numpy.
"""
from __future__ import annotations

import numpy as np

_ASCII = np.frombuffer(b"ACGT", dtype=np.uint8)
_CODE = np.zeros(256, dtype=np.uint8)
for _i, _c in enumerate(b"ACGT"):
    _CODE[_c] = _i

ADAPTER_5P = b"CGCTCTTCCGATCT"
TSO_5P = b"TTTCTTATAT"


def make_candidates(whitelist_ascii: np.ndarray, n: int, seed: int = 2, geometry: str = "5p",
                    umi_len: int = 12, p_sub: float = 0.02, p_ins: float = 0.02,
                    p_del: float = 0.02, frac_negative: float = 0.10, max_len: int = 64,
                    cell_idx: np.ndarray | None = None, umi_codes: np.ndarray | None = None,
                    p_n: float = 0.0):
    """-> dict(seqs u8 buffer, offsets u64 [n+1], true_idx int64 [n] (-1 for negatives),
    umi [n, umi_len] codes).  p_n: per-base probability that the basecaller emitted N (drawn
    last, so p_n = 0 reproduces the earlier streams)."""
    rng = np.random.Generator(np.random.PCG64(seed))
    n_wl = len(whitelist_ascii)
    if cell_idx is None:
        cell_idx = rng.integers(0, n_wl, size=n)
    neg = rng.random(n) < frac_negative
    bc = _CODE[whitelist_ascii[cell_idx]]                       # [n, 16] codes
    rnd = rng.integers(0, 4, size=(n, bc.shape[1]), dtype=np.uint8)
    bc = np.where(neg[:, None], rnd, bc)
    if umi_codes is None:
        umi_codes = rng.integers(0, 4, size=(n, umi_len), dtype=np.uint8)
    if geometry == "5p":
        left = np.broadcast_to(_CODE[np.frombuffer(ADAPTER_5P, np.uint8)], (n, len(ADAPTER_5P)))
        right = np.broadcast_to(_CODE[np.frombuffer(TSO_5P, np.uint8)], (n, len(TSO_5P)))
    elif geometry == "3p":
        left = rng.integers(0, 4, size=(n, 3), dtype=np.uint8)
        right = np.full((n, 4), 3, dtype=np.uint8)
    elif geometry == "slideseq":
        # decon_3pXCR_slideseq keeps 22 nt before / 16 nt after the linker hit of the reverse strand
        # (utils.py:443-448), i.e. 16 + 8 before and 6 + 16 after the linker in barcode orientation;
        # whitelist_ascii holds the 32-column cores (N columns read as A); no separate UMI segment
        left = rng.integers(0, 4, size=(n, 8), dtype=np.uint8)
        right = rng.integers(0, 4, size=(n, 16), dtype=np.uint8)
        umi_codes = umi_codes[:, :0]
    else:
        raise ValueError(geometry)
    tmpl = np.concatenate([left, bc, umi_codes, right], axis=1)   # [n, T]
    T = tmpl.shape[1]
    u = rng.random((n, T))
    is_del = u < p_del
    is_ins = (u >= p_del) & (u < p_del + p_ins)
    is_sub = (u >= p_del + p_ins) & (u < p_del + p_ins + p_sub)
    sub = (tmpl + rng.integers(1, 4, size=(n, T), dtype=np.uint8)) & 3
    base = np.where(is_sub, sub, tmpl)
    ins_base = rng.integers(0, 4, size=(n, T), dtype=np.uint8)
    emit = (~is_del).astype(np.int64) + is_ins.astype(np.int64)     # bases emitted per position
    lens = emit.sum(axis=1)
    lens_c = np.minimum(lens, max_len)
    offsets = np.zeros(n + 1, dtype=np.uint64)
    np.cumsum(lens_c, out=offsets[1:])
    # position of each emitted base inside its candidate
    start = np.cumsum(emit, axis=1) - emit                          # [n, T]
    out = np.zeros(int(offsets[-1]), dtype=np.uint8)
    row0 = offsets[:-1].astype(np.int64)[:, None]
    # inserted base comes first, then the template base
    pos_ins = start
    ok = is_ins & (pos_ins < max_len)
    out[(row0 + pos_ins)[ok]] = _ASCII[ins_base[ok]]
    pos_base = start + is_ins
    ok = (~is_del) & (pos_base < max_len)
    out[(row0 + pos_base)[ok]] = _ASCII[base[ok]]
    if p_n > 0:
        out[rng.random(out.shape[0]) < p_n] = ord("N")
    true_idx = np.where(neg, -1, cell_idx).astype(np.int64)
    return {"seqs": out, "offsets": offsets, "true_idx": true_idx, "umi": umi_codes,
            "lens": lens_c}


def to_strings(seqs: np.ndarray, offsets: np.ndarray) -> list[str]:
    b = seqs.tobytes()
    o = offsets.astype(np.int64)
    return [b[o[i]:o[i + 1]].decode("ascii") for i in range(len(o) - 1)]
