"""ctypes binding of libnanoranger_b200.so (the C ABI in include/nanoranger_b200.h).

There is no CPU fallback: if the library is missing, or a call fails, a RuntimeError is raised.
The library is built in-tree by ``make -C nanoranger_b200/csrc`` (``__graft_entry__.build()``).
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("NANORANGER_B200_LIB") or os.path.join(_HERE, "libnanoranger_b200.so")

NR_MAX_QUERY = 64
NR_MAX_CORE = 32

NR_FLAG_TIE = 0x01
NR_FLAG_RC = 0x02
NR_FLAG_BELOW = 0x04
NR_FLAG_NO_UMI = 0x08
NR_FLAG_TOO_LONG = 0x10
NR_FLAG_EXHAUSTIVE = 0x20
NR_SCORE_BELOW = -128
NR_UMI_NONE = 255

NR_MODE_AUTO = 0
NR_MODE_EXHAUSTIVE = 1
NR_MODE_FILTERED = 2

# every symbol include/nanoranger_b200.h declares (tests/test_abi.py checks header == this list)
SYMBOLS = (
    "nr_last_error", "nr_version", "nr_whitelist_create", "nr_whitelist_destroy",
    "nr_whitelist_size", "nr_whitelist_has_index", "nr_whitelist_device_bytes", "nr_pack_device",
    "nr_match_device", "nr_match_workspace_bytes", "nr_match_host", "nr_host_alloc",
    "nr_host_free", "nr_umi_collapse_device", "nr_umi_collapse_device_keyed", "nr_umi_workspace_bytes", "nr_int_peak",
    "nr_int_peak_dual", "nr_match_device_counted", "nr_match_counters",
    "nr_umi_records_device", "nr_umi_records_workspace_bytes", "nr_umi_partition_device",
    "nr_umi_unzip_device", "nr_hw_search_device", "nr_hw_search_host",
    "nr_sam_write", "nr_match_tier_counts", "nr_sam_write_aligned",
)

_lib = None


def lib() -> C.CDLL:
    """Load the library once; raise loudly when it is absent (no fallback path exists)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"{LIB_PATH} not found: build it with `make -C nanoranger_b200/csrc` "
            "(nvcc, sm_100a). nanoranger_b200 has no CPU fallback.")
    L = C.CDLL(LIB_PATH)
    vp, u64, u32, i32 = C.c_void_p, C.c_uint64, C.c_uint32, C.c_int
    sz = C.c_size_t
    L.nr_last_error.restype = C.c_char_p
    L.nr_version.restype = C.c_char_p
    L.nr_whitelist_create.argtypes = [C.c_char_p, u64, u32, u32, u32, i32, C.POINTER(vp)]
    L.nr_whitelist_create.restype = i32
    L.nr_whitelist_destroy.argtypes = [vp]
    L.nr_whitelist_destroy.restype = None
    L.nr_whitelist_size.argtypes = [vp]
    L.nr_whitelist_size.restype = u64
    L.nr_whitelist_has_index.argtypes = [vp]
    L.nr_whitelist_has_index.restype = i32
    L.nr_whitelist_device_bytes.argtypes = [vp]
    L.nr_whitelist_device_bytes.restype = u64
    L.nr_pack_device.argtypes = [vp, vp, u64, vp, vp, vp, vp]
    L.nr_pack_device.restype = i32
    L.nr_match_device.argtypes = [vp, vp, vp, vp, u64, i32, i32, vp, vp, vp, vp, vp, vp, sz, vp]
    L.nr_match_device.restype = i32
    L.nr_match_device_counted.argtypes = [vp, vp, vp, vp, u64, i32, vp, vp, vp, vp, vp, vp, sz, vp]
    L.nr_match_device_counted.restype = i32
    L.nr_match_workspace_bytes.argtypes = [vp, u64, i32]
    L.nr_match_workspace_bytes.restype = sz
    L.nr_match_host.argtypes = [vp, vp, vp, u64, i32, i32, vp, vp, vp, vp, vp]
    L.nr_match_host.restype = i32
    L.nr_host_alloc.argtypes = [sz]
    L.nr_host_alloc.restype = vp
    L.nr_host_free.argtypes = [vp]
    L.nr_host_free.restype = None
    L.nr_umi_collapse_device.argtypes = [vp, vp, vp, u64, i32, i32, vp, vp, vp, vp, vp, vp, vp, sz, vp]
    L.nr_umi_collapse_device.restype = i32
    L.nr_umi_collapse_device_keyed.argtypes = [vp, vp, vp, u64, i32, i32, i32, i32, i32, vp, vp, vp, vp, vp, vp,
                                               vp, sz, vp]
    L.nr_umi_collapse_device_keyed.restype = i32
    L.nr_umi_workspace_bytes.argtypes = [u64]
    L.nr_umi_workspace_bytes.restype = sz
    L.nr_umi_records_device.argtypes = [vp] * 9 + [u64, i32, i32] + [vp] * 6 + [sz, vp]
    L.nr_umi_records_device.restype = i32
    L.nr_umi_records_workspace_bytes.argtypes = [u64]
    L.nr_umi_records_workspace_bytes.restype = sz
    L.nr_umi_partition_device.argtypes = [vp, vp, vp, vp, u64, i32, vp, vp, vp, vp]
    L.nr_umi_partition_device.restype = i32
    L.nr_umi_unzip_device.argtypes = [vp, u64, vp, vp, vp, vp, vp]
    L.nr_umi_unzip_device.restype = i32
    L.nr_hw_search_device.argtypes = [vp, vp, u64, C.c_char_p, i32, i32, i32, vp, vp, vp, vp, vp]
    L.nr_hw_search_device.restype = i32
    L.nr_hw_search_host.argtypes = [vp, vp, u64, C.c_char_p, i32, i32, i32, vp, vp, vp, vp, i32]
    L.nr_hw_search_host.restype = i32
    L.nr_sam_write.argtypes = [C.c_char_p, i32, vp, vp, vp, vp, u64, vp, vp, vp, vp, vp, vp, vp, u64,
                               u32, u32, u32, C.POINTER(u64)]
    L.nr_sam_write.restype = i32
    L.nr_sam_write_aligned.argtypes = [vp, C.c_char_p, i32, vp, vp, vp, vp, u64, vp, vp, vp, vp, vp, vp, vp,
                                       u64, i32, C.POINTER(u64)]
    L.nr_sam_write_aligned.restype = i32
    L.nr_int_peak.argtypes = [i32, i32, C.POINTER(C.c_double), C.POINTER(C.c_double)]
    L.nr_int_peak.restype = i32
    L.nr_int_peak_dual.argtypes = [i32, i32, C.POINTER(C.c_double), C.POINTER(C.c_double)]
    L.nr_int_peak_dual.restype = i32
    L.nr_match_counters.argtypes = [vp, C.POINTER(u64), vp]
    L.nr_match_counters.restype = i32
    L.nr_match_tier_counts.argtypes = [vp, C.POINTER(u64), vp]
    L.nr_match_tier_counts.restype = i32
    _lib = L
    return L


def check(rc: int, what: str) -> None:
    """Turn a negative NR_E* return code into a RuntimeError carrying nr_last_error()."""
    if rc != 0:
        msg = lib().nr_last_error()
        raise RuntimeError(f"{what} failed (rc={rc}): {msg.decode() if msg else ''}")
