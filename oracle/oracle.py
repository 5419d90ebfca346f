"""ctypes front-end of oracle/libnr_oracle.so plus a pure-numpy twin of the literal DP.

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's CPU
baseline legs.  The product package (nanoranger_b200/) never imports this module.

PARITY UNPINNED at the STAR boundary (see nr_oracle.c): the reference ships no golden
vectors and its aligner (STAR) is absent; the scoring restated here follows
/root/reference/scripts/barcode_align.sh:18-33 and SURVEY.md Appendix C.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libnr_oracle.so")
_lib = None

_CODE = np.full(256, 4, dtype=np.uint8)
for _i, _ch in enumerate("ACGT"):
    _CODE[ord(_ch)] = _i
    _CODE[ord(_ch.lower())] = _i


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "nr_oracle.c")
    if force or not os.path.exists(_SO) or (
        os.path.exists(src) and os.path.getmtime(src) > os.path.getmtime(_SO)
    ):
        subprocess.check_call(["make", "-s", "-C", _HERE, "-B", "libnr_oracle.so"])
    return _SO


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_SO)
        u8p, i8p, i32p, i16p, u32p = (C.POINTER(t) for t in
                                      (C.c_uint8, C.c_int8, C.c_int32, C.c_int16, C.c_uint32))
        L.nr_oracle_as_padded.argtypes = [u8p, C.c_int, u8p, C.c_int]
        L.nr_oracle_as_padded.restype = C.c_int
        L.nr_oracle_pair.argtypes = [u8p, C.c_int, u8p, C.c_int, C.c_int, C.c_int,
                                     C.POINTER(C.c_int)]
        L.nr_oracle_pair.restype = C.c_int
        L.nr_oracle_match.argtypes = [u8p, C.c_int64, C.c_int, C.c_int, C.c_int, u8p, u8p,
                                      C.c_int64, C.c_int, i32p, i8p, i32p, u8p, i16p]
        L.nr_oracle_match.restype = C.c_int
        L.nr_oracle_scores.argtypes = [u8p, C.c_int64, C.c_int, C.c_int, C.c_int, u8p, C.c_int,
                                       i8p]
        L.nr_oracle_scores.restype = C.c_int
        L.nr_oracle_umi_cluster.argtypes = [u32p, u32p, u32p, C.c_int64, C.c_int, u32p]
        L.nr_oracle_umi_cluster.restype = C.c_int64
        L.nr_oracle_hw_search.argtypes = [C.c_char_p, C.c_int, C.c_char_p, C.c_int, C.c_int, C.c_int,
                                          C.POINTER(C.c_int)]
        L.nr_oracle_hw_search.restype = C.c_int
        _lib = L
    return _lib


def _p(a, t):
    return a.ctypes.data_as(C.POINTER(t))


def encode(seq: str) -> np.ndarray:
    return _CODE[np.frombuffer(seq.encode("ascii"), dtype=np.uint8)]


def encode_many(seqs, width: int) -> tuple[np.ndarray, np.ndarray]:
    """-> (codes [n, width] padded with 4, lengths [n])."""
    n = len(seqs)
    out = np.full((n, width), 4, dtype=np.uint8)
    lens = np.zeros(n, dtype=np.int64)
    for i, s in enumerate(seqs):
        c = encode(s)
        if len(c) > width:
            raise ValueError(f"sequence {i} longer than {width}")
        out[i, : len(c)] = c
        lens[i] = len(c)
    return out, lens


def as_padded(q: str, ref: str) -> int:
    """tier 0: literal Appendix-C DP of candidate q against the padded reference string."""
    qc, rc = np.ascontiguousarray(encode(q)), np.ascontiguousarray(encode(ref))
    return lib().nr_oracle_as_padded(_p(qc, C.c_uint8), len(qc), _p(rc, C.c_uint8), len(rc))


def pair(q: str, core: str, pad_l: int, pad_r: int) -> tuple[int, int]:
    """tier 1: (AS, umi_q) of one pair on the core columns."""
    qc, cc = np.ascontiguousarray(encode(q)), np.ascontiguousarray(encode(core))
    u = C.c_int(-1)
    a = lib().nr_oracle_pair(_p(qc, C.c_uint8), len(qc), _p(cc, C.c_uint8), len(cc), pad_l,
                             pad_r, C.byref(u))
    return a, u.value


def match(wl_codes: np.ndarray, pad_l: int, pad_r: int, cand_codes: np.ndarray,
          cand_len: np.ndarray, threads: int | None = None) -> dict:
    """tier 2: exhaustive best / tie count / argmin / strand / UMI column per candidate."""
    wl_codes = np.ascontiguousarray(wl_codes, dtype=np.uint8)
    n, L = wl_codes.shape
    cand = np.full((len(cand_codes), 64), 4, dtype=np.uint8)
    cand[:, : cand_codes.shape[1]] = cand_codes
    clen = np.ascontiguousarray(cand_len, dtype=np.uint8)
    N = len(clen)
    out = dict(
        best_idx=np.full(N, -1, np.int32), best_score=np.zeros(N, np.int8),
        n_best=np.zeros(N, np.int32), strand=np.zeros(N, np.uint8),
        umi_q=np.full(N, -1, np.int16))
    if N == 0:
        return out
    threads = threads or os.cpu_count() or 1
    rc = lib().nr_oracle_match(
        _p(wl_codes, C.c_uint8), n, L, pad_l, pad_r, _p(cand, C.c_uint8), _p(clen, C.c_uint8),
        N, threads, _p(out["best_idx"], C.c_int32), _p(out["best_score"], C.c_int8),
        _p(out["n_best"], C.c_int32), _p(out["strand"], C.c_uint8), _p(out["umi_q"], C.c_int16))
    if rc != 0:
        raise ValueError("nr_oracle_match: bad arguments")
    return out


def scores(wl_codes: np.ndarray, pad_l: int, pad_r: int, q: str) -> np.ndarray:
    wl_codes = np.ascontiguousarray(wl_codes, dtype=np.uint8)
    n, L = wl_codes.shape
    qc = np.ascontiguousarray(encode(q))
    out = np.zeros(n, np.int8)
    rc = lib().nr_oracle_scores(_p(wl_codes, C.c_uint8), n, L, pad_l, pad_r, _p(qc, C.c_uint8),
                                len(qc), _p(out, C.c_int8))
    if rc != 0:
        raise ValueError("nr_oracle_scores: bad arguments")
    return out


def umi_cluster(bc: np.ndarray, gene: np.ndarray, umi: np.ndarray, max_dist: int = 1):
    bc = np.ascontiguousarray(bc, np.uint32)
    gene = np.ascontiguousarray(gene, np.uint32)
    umi = np.ascontiguousarray(umi, np.uint32)
    out = np.zeros(len(bc), np.uint32)
    k = lib().nr_oracle_umi_cluster(_p(bc, C.c_uint32), _p(gene, C.c_uint32),
                                    _p(umi, C.c_uint32), len(bc), max_dist,
                                    _p(out, C.c_uint32))
    if k < 0:
        raise MemoryError
    return int(k), out


# ---------------------------------------------------------------------------------------
# numpy twin of the literal DP -- slow, for small cross-checks only.

def as_padded_numpy(q: str, ref: str) -> int:
    qc, rc = encode(q).astype(np.int64), encode(ref).astype(np.int64)
    m, n = len(qc), len(rc)
    prev = np.zeros(n + 1, dtype=np.int64)
    for i in range(1, m + 1):
        s = np.where((qc[i - 1] > 3) | (rc > 3), 0, np.where(rc == qc[i - 1], 1, -1))
        cand = np.maximum(prev[:-1] + s, prev[1:] - 1)          # diag / up
        cur = np.empty(n + 1, dtype=np.int64)
        cur[0] = -i
        # left-to-right dependency: cur[j] = max(cand[j-1], cur[j-1] - 1)
        run = cur[0]
        for j in range(1, n + 1):
            run = max(cand[j - 1], run - 1)
            cur[j] = run
        prev = cur
    return int(prev.max())


def revcomp(s: str) -> str:
    return s.translate(str.maketrans("ACGTNacgtn", "TGCANtgcan"))[::-1]


def hw_search(query: str, target: str, k: int, wildcard: bool = True) -> dict:
    """What the reference asks of edlib.align(query, target, "HW", "locations", k[, ad_seq])
    (utils.py:134 ...): {"editDistance": d or -1, "locations": [first (start, end), last (start,
    end)], "n_locations": count}; end inclusive, as edlib reports it."""
    out = (C.c_int * 6)()
    lib().nr_oracle_hw_search(query.encode(), len(query), target.encode(), len(target), k,
                              1 if wildcard else 0, out)
    return {"editDistance": out[0], "n_locations": out[1], "first": (out[2], out[3]),
            "last": (out[4], out[5])}


def hw_search_numpy(query: str, target: str, k: int, wildcard: bool = True) -> dict:
    """Independent twin of nr_oracle_hw_search: full DP matrix in numpy, every optimal start
    enumerated by brute force over substrings (small inputs only)."""
    def eq(p, c):
        if p == c:
            return True
        return wildcard and ((p == "N" and c in "ACGT") or (c == "N" and p in "ACGT"))

    def dist(a, b):          # global unit-cost edit distance
        D = np.arange(len(b) + 1)
        for i in range(1, len(a) + 1):
            prev, D = D, np.empty(len(b) + 1, np.int64)
            D[0] = i
            for j in range(1, len(b) + 1):
                D[j] = min(prev[j - 1] + (0 if eq(a[i - 1], b[j - 1]) else 1), prev[j] + 1, D[j - 1] + 1)
        return int(D[-1])

    n = len(target)
    best, ends = None, {}
    for e in range(n):
        for s in range(e + 2):                       # s == e + 1: empty substring ending at e
            d = dist(query, target[s:e + 1])
            if best is None or d < best:
                best, ends = d, {}
            if d == best:
                ends.setdefault(e, s)                 # smallest start first
    if best is None or best > k:
        return {"editDistance": -1, "n_locations": 0, "first": (-1, -1), "last": (-1, -1)}
    es = sorted(ends)
    return {"editDistance": best, "n_locations": len(es), "first": (ends[es[0]], es[0]),
            "last": (ends[es[-1]], es[-1])}


# ---------------------------------------------------------------------------------------
# CPU build of the GPU path's own algorithm (nr_filter_cpu.cpp): bench baseline "port-filtered".

_fc = None


def filter_cpu_lib():
    global _fc
    if _fc is None:
        so = os.path.join(_HERE, "libnr_filter_cpu.so")
        src = os.path.join(_HERE, "nr_filter_cpu.cpp")
        hdr = os.path.join(_HERE, "..", "nanoranger_b200", "csrc", "nr_filter_core.h")
        if not os.path.exists(so) or (os.path.exists(src) and os.path.exists(hdr) and
                                      os.path.getmtime(so) < max(os.path.getmtime(src), os.path.getmtime(hdr))):
            subprocess.check_call(["make", "-s", "-C", _HERE, "-B", "libnr_filter_cpu.so"])
        L = C.CDLL(so)
        L.nr_cpu_filter_create.restype = C.c_void_p
        L.nr_cpu_filter_create.argtypes = [C.c_void_p, C.c_int64, C.c_int, C.c_int]
        L.nr_cpu_filter_destroy.argtypes = [C.c_void_p]
        L.nr_cpu_filter_match.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int, C.c_int] + \
            [C.c_void_p] * 6
        L.nr_cpu_filter_match.restype = C.c_int
        _fc = L
    return _fc


class FilterCPU:
    """the seed filter on the host cores, 16-column N-free whitelists."""

    def __init__(self, wl_codes: np.ndarray, pad_l: int, pad_r: int):
        wl_codes = np.ascontiguousarray(wl_codes, dtype=np.uint8)
        assert wl_codes.shape[1] == 16 and (wl_codes < 4).all()
        lo = np.zeros(len(wl_codes), np.uint32)
        for j in range(16):
            lo |= wl_codes[:, j].astype(np.uint32) << np.uint32(2 * j)
        self._lo = lo
        self._h = filter_cpu_lib().nr_cpu_filter_create(lo.ctypes.data, len(lo), pad_l, pad_r)

    def match(self, cand_codes, cand_len, threads=None, min_len=24):
        cand = np.full((len(cand_codes), 64), 4, dtype=np.uint8)
        cand[:, : cand_codes.shape[1]] = cand_codes
        clen = np.ascontiguousarray(cand_len, dtype=np.uint8)
        N = len(clen)
        out = dict(best_idx=np.full(N, -1, np.int32), best_score=np.zeros(N, np.int8),
                   n_best=np.zeros(N, np.int32), strand=np.zeros(N, np.uint8),
                   umi_q=np.full(N, -1, np.int16), took=np.zeros(N, np.uint8))
        rc = filter_cpu_lib().nr_cpu_filter_match(
            self._h, cand.ctypes.data, clen.ctypes.data, N, min_len, threads or os.cpu_count() or 1,
            out["best_idx"].ctypes.data, out["best_score"].ctypes.data, out["n_best"].ctypes.data,
            out["strand"].ctypes.data, out["umi_q"].ctypes.data, out["took"].ctypes.data)
        if rc != 0:
            raise ValueError("nr_cpu_filter_match: bad arguments")
        return out

    def close(self):
        if self._h:
            filter_cpu_lib().nr_cpu_filter_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
