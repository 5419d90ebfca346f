"""UMI collapse kernel vs the oracle twin (oracle/nr_oracle.c: nr_oracle_umi_cluster)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _check(oracle, bc, gene, umi, umi_len, max_dist):
    from nanoranger_b200 import umi as U
    r = U.collapse_host(bc, gene, umi, umi_len, max_dist)
    k, rep = oracle.umi_cluster(bc, gene, umi, max_dist)
    assert r["n_groups"] == k
    assert np.array_equal(r["rep_umi"], rep)
    # group table == (bc, gene, rep) value counts, sorted
    key = np.stack([bc, gene, rep], axis=1).astype(np.int64)
    uk, cnt = np.unique(key, axis=0, return_counts=True)
    assert np.array_equal(np.stack([r["g_bc"], r["g_gene"], r["g_umi"]], axis=1).astype(np.int64), uk)
    assert np.array_equal(r["g_reads"].astype(np.int64), cnt)


@pytest.mark.parametrize("max_dist", [0, 1])
def test_umi_collapse_random(cuda_device, oracle, max_dist):
    rng = np.random.default_rng(3 + max_dist)
    n = 200000
    bc = rng.zipf(1.3, n).astype(np.uint32) % 3000
    gene = rng.integers(0, 20, n).astype(np.uint32)
    base = rng.integers(0, 1 << 24, 4000).astype(np.uint32)
    umi = base[rng.integers(0, 4000, n)]
    flip = rng.random(n) < 0.08                                   # sequencing errors in the UMI
    umi = np.where(flip, umi ^ (rng.integers(1, 4, n).astype(np.uint32) << (2 * rng.integers(0, 12, n)).astype(np.uint32)), umi)
    _check(oracle, bc, gene, umi.astype(np.uint32), 12, max_dist)


def test_umi_collapse_big_groups_and_edges(cuda_device, oracle):
    rng = np.random.default_rng(8)
    # one barcode with thousands of UMIs (TCR / slide-seq style: gene = 0), tight UMI space
    n = 30000
    bc = np.where(rng.random(n) < 0.7, 7, rng.integers(0, 50, n)).astype(np.uint32)
    gene = np.zeros(n, np.uint32)
    umi = rng.integers(0, 1 << 12, n).astype(np.uint32)            # 6-nt space: many neighbours
    for md in (0, 1):
        _check(oracle, bc, gene, umi, 6, md)
    # single record, all identical, 16-nt UMIs
    _check(oracle, np.array([5], np.uint32), np.array([1], np.uint32), np.array([9], np.uint32), 12, 1)
    _check(oracle, np.full(100, 3, np.uint32), np.zeros(100, np.uint32), np.full(100, 77, np.uint32), 12, 1)
    u16 = rng.integers(0, 1 << 32, 5000, dtype=np.uint64).astype(np.uint32)
    _check(oracle, rng.integers(0, 10, 5000).astype(np.uint32), np.zeros(5000, np.uint32), u16, 16, 1)
    from nanoranger_b200 import umi as U
    r = U.collapse_host(np.zeros(0, np.uint32), np.zeros(0, np.uint32), np.zeros(0, np.uint32), 12, 1)
    assert r["n_groups"] == 0 and len(r["rep_umi"]) == 0


def test_exact_dedup_equals_numpy_unique_per_barcode(cuda_device):
    """the reference's U1 block (utils.py:759-773): np.unique of the UMI strings per barcode."""
    from nanoranger_b200 import utils
    rng = np.random.default_rng(4)
    bcs = ["BC%03d" % i for i in rng.integers(0, 40, 3000)]
    umis = ["".join("ACGT"[j] for j in rng.integers(0, 4, 10)) if rng.random() > 0.02 else "ACGTNACGTA"
            for _ in range(3000)]
    umis = [u if rng.random() > 0.5 else umis[0] for u in umis]
    df = utils.umi_dedup_table(bcs, umis, 10)
    exp = {}
    for b, u in zip(bcs, umis):
        exp.setdefault(b, []).append(u)
    assert len(df) == len(exp)
    for b, us in exp.items():
        assert df.loc[b, "umi_cnt"] == len(np.unique(us)) and df.loc[b, "read_cnt"] == len(us)
