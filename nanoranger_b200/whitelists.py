"""Whitelists shipped with the package.

``load_737k`` returns the 10x 737K-august-2016 list (the reference's data/737K-august-2016.txt.gz,
used by pipeline.py for the 5' modes) from a delta-coded copy: the list is sorted, so the
2-bit big-endian codes are stored as successive differences (24 KB compressed).
"""
from __future__ import annotations

import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_ASCII = np.frombuffer(b"ACGT", dtype=np.uint8)

LINKER_SLIDESEQ = "TCTTCAGCGTTCCCGAGA"   # utils.py:14 of the reference


def codes_to_ascii(be: np.ndarray, length: int = 16) -> np.ndarray:
    """big-endian 2-bit codes (base 0 in the top bits) -> [n, length] ASCII bytes."""
    out = np.empty((len(be), length), dtype=np.uint8)
    for j in range(length):
        out[:, j] = _ASCII[(be >> np.uint32(2 * (length - 1 - j))) & np.uint32(3)]
    return out


def load_737k() -> np.ndarray:
    """-> [737280, 16] uint8 ASCII, in the file's (lexicographic) order."""
    d = np.load(os.path.join(_HERE, "data", "737K-august-2016.delta.npz"))["delta"]
    be = np.cumsum(d.astype(np.int64)).astype(np.uint32)
    return codes_to_ascii(be, 16)


def synthetic_whitelist(n: int, seed: int = 20180201, length: int = 16) -> np.ndarray:
    """n distinct random `length`-mers with pairwise Hamming distance >= 2 (stand-in for the
    3M-february-2018 list, which is missing from the reference checkout: SURVEY.md section 8d).
    Distance >= 2 is obtained with a parity base: the last base is the sum of the others mod 4.
    -> [n, length] uint8 ASCII, sorted."""
    rng = np.random.Generator(np.random.PCG64(seed))
    bits = 2 * (length - 1)
    need = int(n * 1.02) + 16
    vals = np.unique(rng.integers(0, 1 << bits, size=need, dtype=np.int64))
    while len(vals) < n:
        vals = np.unique(np.concatenate([vals, rng.integers(0, 1 << bits, size=need, dtype=np.int64)]))
    vals = np.sort(rng.permutation(vals)[:n])
    body = np.empty((n, length - 1), dtype=np.uint8)
    for j in range(length - 1):
        body[:, j] = (vals >> (2 * (length - 2 - j))) & 3
    parity = body.sum(axis=1, dtype=np.int64) & 3
    codes = np.concatenate([body, parity[:, None].astype(np.uint8)], axis=1)
    return _ASCII[codes]


def ascii_to_strings(a: np.ndarray) -> list[str]:
    return [bytes(r).decode("ascii") for r in a]
