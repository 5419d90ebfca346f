"""nanoranger_b200/h5lite.py (pure-Python HDF5 subset) and the ``.h5`` whitelist inputs of the
reference (utils.py:606-610 write_bc_5p10X, utils.py:1116-1132 write_bc_3p10XTCR_nuc), which the
reference reads through scanpy (absent here).  The reader is pinned to a file libhdf5 itself wrote
(scipy ships a MATLAB v7.3 file: superblock 0 behind a 512-byte user block, symbol-table group,
object header v1, contiguous layout; known content 0 : pi/4 : 2 pi) and exercised on 10x-style
fixtures from tests/h5write.py (chunk B-trees of one and two levels, shuffle + deflate)."""
import os

import numpy as np
import pytest

from h5write import write_10x
from nanoranger_b200 import utils
from nanoranger_b200.h5lite import H5Error, H5Lite, read_10x_h5_barcodes


def test_reads_a_file_written_by_libhdf5():
    import scipy.io
    p = os.path.join(os.path.dirname(scipy.io.__file__), "matlab", "tests", "data", "testhdf5_7.4_GLNX86.mat")
    if not os.path.exists(p):
        pytest.skip("scipy's MATLAB v7.3 test file is not installed")
    with H5Lite(p) as f:
        assert f.base == 512 and list(f.members()) == ["testdouble"]
        x = f.read("testdouble")
    assert x.dtype == np.float64 and x.shape == (9, 1)
    assert np.allclose(x.ravel(), np.arange(9) * np.pi / 4)      # scipy's `testdouble` vector


def _barcodes(rng, n):
    return ["".join("ACGT"[i] for i in rng.integers(0, 4, 16)) + "-1" for _ in range(n)]


@pytest.mark.parametrize("version,userblock,zeros", [(3, 0, 0), (2, 0, 0), (3, 512, 2)])
def test_10x_barcodes_filter_cells(tmp_path, version, userblock, zeros):
    rng = np.random.default_rng(version + userblock)
    bcs = _barcodes(rng, 53)
    genes = [int(g) for g in rng.integers(0, 41, len(bcs))]
    genes[:4] = [19, 20, 3, 4]                                   # both thresholds' boundaries
    p = str(tmp_path / "m.h5")
    write_10x(p, bcs, genes, version=version, userblock=userblock, explicit_zeros=zeros)
    with H5Lite(p) as f:
        grp = "matrix" if version == 3 else "GRCh38"
        assert [b.decode() for b in f.read(grp + "/barcodes")] == bcs
        assert list(f.read(grp + "/shape")) == [50, 53]
        ip = f.read(grp + "/indptr")
        assert ip.dtype == np.int64 and list(np.diff(ip)) == [g + zeros for g in genes]
    for thr in (20, 4):
        got = read_10x_h5_barcodes(p, thr)
        # stored zeros do not count as detected genes (sc.pp.filter_cells counts X > 0)
        assert got == [b for b, g in zip(bcs, genes) if g >= thr]


def test_10x_new_style_groups(tmp_path):
    """superblock 2, object headers v2 with link messages (HDF5 1.8 'latest' groups, compact)"""
    rng = np.random.default_rng(9)
    bcs = _barcodes(rng, 21)
    genes = [int(g) for g in rng.integers(0, 30, len(bcs))]
    p = str(tmp_path / "n.h5")
    write_10x(p, bcs, genes, new_groups=True)
    with H5Lite(p) as f:
        assert list(f.members()) == ["matrix"]
        assert sorted(f.members(f.lookup("matrix"))) == ["barcodes", "data", "indices", "indptr", "shape"]
    assert read_10x_h5_barcodes(p, 4) == [b for b, g in zip(bcs, genes) if g >= 4]


def test_write_bc_from_h5(tmp_path):
    """W1's .h5 branch and W3: cells with >= 20 / >= 4 genes, first 16 characters, pads 30/40 and
    16/28 (reference utils.py:606-622, 1116-1132), byte for byte."""
    rng = np.random.default_rng(5)
    bcs = _barcodes(rng, 30)
    genes = [int(g) for g in rng.integers(0, 30, len(bcs))]
    p = str(tmp_path / "filtered_feature_bc_matrix.h5")
    write_10x(p, bcs, genes)
    utils.write_bc_5p10X("s", str(tmp_path), p)
    exp = "".join(f">{b[:16]}\n{'N' * 30}{b[:16]}{'N' * 40}\n" for b, g in zip(bcs, genes) if g >= 20)
    assert open(tmp_path / "s_bcreads.fasta").read() == exp and exp
    utils.write_bc_3p10XTCR_nuc("t", str(tmp_path), p)
    exp = "".join(f">{b[:16]}\n{'N' * 16}{b[:16]}{'N' * 28}\n" for b, g in zip(bcs, genes) if g >= 4)
    assert open(tmp_path / "t_bcreads.fasta").read() == exp and exp


def test_not_hdf5(tmp_path):
    p = tmp_path / "x.h5"
    p.write_bytes(b"not an hdf5 file" * 100)
    with pytest.raises(H5Error):
        H5Lite(str(p))
