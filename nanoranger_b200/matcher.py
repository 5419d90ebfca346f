"""Whitelist handle and matcher calls on top of the C ABI.

``Whitelist`` replaces the STAR index directory built by scripts/barcode_ref.sh:11-18 of the
reference; ``Whitelist.match_host`` / ``match_device`` replace the STAR run of
scripts/barcode_align.sh:14-41 plus the per-record geometry utils.process_matching_* derives
from the SAM (AS tag, flag, RNAME, query index aligned to reference column padL+L:
utils.py:697-708).  torch is used for device memory and streams only.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass

import numpy as np

from . import _lib
from ._lib import (NR_FLAG_BELOW, NR_FLAG_RC, NR_FLAG_TOO_LONG, NR_MODE_AUTO)


def pack_ascii(seqs) -> tuple[np.ndarray, np.ndarray]:
    """list of str/bytes -> (concatenated ASCII bytes u8, offsets u64 [n+1])."""
    bs = [s if isinstance(s, (bytes, bytearray)) else s.encode("ascii") for s in seqs]
    lens = np.fromiter((len(b) for b in bs), dtype=np.uint64, count=len(bs))
    offsets = np.zeros(len(bs) + 1, dtype=np.uint64)
    np.cumsum(lens, out=offsets[1:])
    buf = np.frombuffer(b"".join(bs), dtype=np.uint8) if bs else np.zeros(0, np.uint8)
    return buf, offsets


@dataclass
class MatchResult:
    """Per-candidate outputs of the matcher (numpy on the host, torch tensors on the device)."""
    idx: object      # int32  entry index of the best pair (smallest among ties), -1 none
    score: object    # int8   best AS over all entries and both strands
    nbest: object    # uint8  number of (entry, strand) pairs attaining it (saturated)
    flags: object    # uint8  NR_FLAG_* bits
    umi_q: object    # uint8  query index aligned to reference column padL+L, 255 = none

    def assigned(self, min_score: int):
        """Candidates STAR would report with flag 0 and the reference would keep
        (unique best pair, forward strand, AS >= threshold: utils.py:699)."""
        bad = NR_FLAG_RC | NR_FLAG_BELOW | NR_FLAG_TOO_LONG
        return (self.nbest == 1) & ((self.flags & bad) == 0) & (self.score >= min_score)


class Whitelist:
    """Packed whitelist + seed index resident on one GPU."""

    def __init__(self, cores, pad_l: int, pad_r: int, device: int = 0):
        """cores: sequence of equal-length core strings (ACGTN), the FASTA sequences written by
        utils.write_bc_* without their N pads."""
        L = _lib.lib()
        if isinstance(cores, np.ndarray) and cores.dtype == np.uint8 and cores.ndim == 2:
            n, core_len = cores.shape
            blob = np.ascontiguousarray(cores).tobytes()
        else:
            cores = list(cores)
            if not cores:
                raise ValueError("empty whitelist")
            core_len = len(cores[0])
            if any(len(c) != core_len for c in cores):
                raise ValueError("whitelist cores must all have the same length")
            n = len(cores)
            blob = "".join(cores).encode("ascii")
        self.n, self.core_len, self.pad_l, self.pad_r, self.device = n, core_len, pad_l, pad_r, device
        h = C.c_void_p()
        _lib.check(L.nr_whitelist_create(blob, n, core_len, pad_l, pad_r, device, C.byref(h)),
                   "nr_whitelist_create")
        self._h = h

    def close(self):
        if getattr(self, "_h", None):
            _lib.lib().nr_whitelist_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def handle(self):
        return self._h

    @property
    def has_index(self) -> bool:
        return bool(_lib.lib().nr_whitelist_has_index(self._h))

    @property
    def device_bytes(self) -> int:
        return int(_lib.lib().nr_whitelist_device_bytes(self._h))

    # ---- host buffers: the call a reference-side stub binds ---------------------------------
    def match_host(self, seqs, offsets=None, min_score: int = 14, mode: int = NR_MODE_AUTO,
                   out: MatchResult | None = None) -> MatchResult:
        """seqs: list of str, or (u8 buffer, u64 offsets).  H2D, pack, match, D2H inside."""
        if offsets is None:
            buf, offsets = pack_ascii(seqs)
        else:
            buf = np.ascontiguousarray(seqs, dtype=np.uint8)
            offsets = np.ascontiguousarray(offsets, dtype=np.uint64)
        n = len(offsets) - 1
        if out is None:
            out = MatchResult(np.empty(n, np.int32), np.empty(n, np.int8), np.empty(n, np.uint8),
                              np.empty(n, np.uint8), np.empty(n, np.uint8))
        if n == 0:
            return out
        _lib.check(_lib.lib().nr_match_host(
            self._h, buf.ctypes.data, offsets.ctypes.data, n, min_score, mode,
            out.idx.ctypes.data, out.score.ctypes.data, out.nbest.ctypes.data,
            out.flags.ctypes.data, out.umi_q.ctypes.data), "nr_match_host")
        return out

    # ---- device buffers (torch tensors), stream-ordered, no synchronisation --------------------
    def pack_device(self, d_seqs, d_offsets):
        """ASCII batch on the GPU -> (bases [n,16] u8, meta [n] u8, nmask [n] i64)."""
        import torch
        n = d_offsets.numel() - 1
        dev = d_seqs.device
        bases = torch.empty((n, 16), dtype=torch.uint8, device=dev)
        meta = torch.empty(n, dtype=torch.uint8, device=dev)
        nmask = torch.empty(n, dtype=torch.int64, device=dev)
        st = torch.cuda.current_stream(dev).cuda_stream
        _lib.check(_lib.lib().nr_pack_device(d_seqs.data_ptr(), d_offsets.data_ptr(), n,
                                             bases.data_ptr(), meta.data_ptr(), nmask.data_ptr(),
                                             st), "nr_pack_device")
        return bases, meta, nmask

    def alloc_result(self, n: int, device) -> MatchResult:
        import torch
        return MatchResult(torch.empty(n, dtype=torch.int32, device=device),
                           torch.empty(n, dtype=torch.int8, device=device),
                           torch.empty(n, dtype=torch.uint8, device=device),
                           torch.empty(n, dtype=torch.uint8, device=device),
                           torch.empty(n, dtype=torch.uint8, device=device))

    def workspace(self, n: int, device, mode: int = NR_MODE_AUTO):
        import torch
        nb = _lib.lib().nr_match_workspace_bytes(self._h, n, mode)
        return torch.empty(nb, dtype=torch.uint8, device=device)

    def match_device(self, bases, meta, nmask, min_score: int = 14, mode: int = NR_MODE_AUTO,
                     out: MatchResult | None = None, workspace=None, counted: bool = False):
        import torch
        n = meta.numel()
        dev = meta.device
        if out is None:
            out = self.alloc_result(n, dev)
        if workspace is None:
            workspace = self.workspace(n, dev, mode)
        st = torch.cuda.current_stream(dev).cuda_stream
        L = _lib.lib()
        if counted:
            _lib.check(L.nr_match_device_counted(
                self._h, bases.data_ptr(), meta.data_ptr(), nmask.data_ptr(), n, min_score,
                out.idx.data_ptr(), out.score.data_ptr(), out.nbest.data_ptr(),
                out.flags.data_ptr(), out.umi_q.data_ptr(), workspace.data_ptr(),
                workspace.numel(), st), "nr_match_device_counted")
        else:
            _lib.check(L.nr_match_device(
                self._h, bases.data_ptr(), meta.data_ptr(), nmask.data_ptr(), n, min_score, mode,
                out.idx.data_ptr(), out.score.data_ptr(), out.nbest.data_ptr(),
                out.flags.data_ptr(), out.umi_q.data_ptr(), workspace.data_ptr(),
                workspace.numel(), st), "nr_match_device")
        return out

    def counters(self, workspace) -> dict:
        import torch
        c = (C.c_uint64 * 5)()
        st = torch.cuda.current_stream(workspace.device).cuda_stream
        _lib.check(_lib.lib().nr_match_counters(workspace.data_ptr(), c, st), "nr_match_counters")
        return dict(zip(("probes", "hits", "verifications", "passes", "listed"), map(int, c)))

    def tier_counts(self, workspace) -> dict:
        """Where the last match_device call on `workspace` resolved its candidates."""
        import torch
        c = (C.c_uint64 * 4)()
        st = torch.cuda.current_stream(workspace.device).cuda_stream
        _lib.check(_lib.lib().nr_match_tier_counts(workspace.data_ptr(), c, st), "nr_match_tier_counts")
        return dict(zip(("left_by_filter", "deep_k3", "deep_k5", "brute_force"), map(int, c)))


def int_peak(device: int = 0, iters: int = 2000) -> dict:
    """ALU-pipe roofline denominator measured on this GPU (thread-ops/s)."""
    L = _lib.lib()
    a, ms = C.c_double(), C.c_double()
    _lib.check(L.nr_int_peak(device, iters, C.byref(a), C.byref(ms)), "nr_int_peak")
    b, ms2 = C.c_double(), C.c_double()
    _lib.check(L.nr_int_peak_dual(device, iters, C.byref(b), C.byref(ms2)), "nr_int_peak_dual")
    return {"alu_ops_per_s": a.value, "alu_ms": ms.value, "dual_ops_per_s": b.value,
            "dual_ms": ms2.value}
