"""The C-ABI library loads and exports every symbol include/nanoranger_b200.h declares.
No compute calls here (no GPU on the CPU test box)."""
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def built():
    so = os.path.join(ROOT, "nanoranger_b200", "libnanoranger_b200.so")
    if not os.path.exists(so):
        subprocess.check_call(["make", "-s", "-j", "8", "-C", os.path.join(ROOT, "nanoranger_b200", "csrc")])
    return so


def header_symbols():
    txt = open(os.path.join(ROOT, "include", "nanoranger_b200.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(nr_[a-z0-9_]+)\s*\(", txt)))


def test_header_symbols_exported(built):
    from nanoranger_b200 import _lib
    L = _lib.lib()
    hs = header_symbols()
    assert len(hs) >= 18
    missing = [s for s in hs if not hasattr(L, s)]
    assert not missing, missing
    assert sorted(_lib.SYMBOLS) == hs          # the Python binding covers the header, no more, no less
    nm = subprocess.check_output(["nm", "-D", "--defined-only", built], text=True)
    exported = set(re.findall(r" T (nr_[a-z0-9_]+)", nm))
    assert set(hs) <= exported


def test_version_and_error_string(built):
    from nanoranger_b200 import _lib
    L = _lib.lib()
    assert b"sm_100a" in L.nr_version()
    assert isinstance(L.nr_last_error(), bytes)


def test_argument_errors_without_gpu(built):
    """argument validation happens before any CUDA call"""
    import ctypes as C
    from nanoranger_b200 import _lib
    L = _lib.lib()
    h = C.c_void_p()
    assert L.nr_whitelist_create(b"ACGT", 0, 4, 1, 1, 0, C.byref(h)) == -1          # n == 0
    assert L.nr_whitelist_create(b"A" * 40, 1, 40, 1, 1, 0, C.byref(h)) == -1       # core too long
    assert b"bad arguments" in L.nr_last_error()
    assert L.nr_match_host(None, None, None, 5, 14, 0, None, None, None, None, None) == -1
    assert L.nr_match_workspace_bytes(None, 1000, 0) == 65536 + 16128          # header + four lists
    assert L.nr_umi_records_workspace_bytes(0) > 0
    assert L.nr_umi_partition_device(None, None, None, None, 5, 0, None, None, None, None) == -1   # world 0


def test_missing_library_fails_loudly(monkeypatch, built):
    from nanoranger_b200 import _lib
    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "LIB_PATH", "/nonexistent/libnanoranger_b200.so")
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        _lib.lib()
