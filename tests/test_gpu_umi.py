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


def test_records_from_match_results(cuda_device, oracle):
    """nr_umi_records_device == the per-record loop of process_matching_* on the oracle's
    answers: assigned candidates only, umi = seq[q:q+len], short and N-containing UMIs dropped."""
    import torch
    from nanoranger_b200 import NR_MODE_FILTERED, Whitelist, pack_ascii, synth, whitelists
    from nanoranger_b200 import umi as U
    wl_a = whitelists.load_737k()
    d = synth.make_candidates(wl_a, 3000, seed=11)
    seqs = synth.to_strings(d["seqs"], d["offsets"])
    rng = np.random.default_rng(1)
    for i in rng.integers(0, len(seqs), 150):              # sprinkle N into some UMIs / barcodes
        j = int(rng.integers(0, len(seqs[i])))
        seqs[i] = seqs[i][:j] + "N" + seqs[i][j + 1:]
    seqs += [s[:34] for s in seqs[:200]]                   # cut inside the UMI: short
    wl = Whitelist(wl_a, 30, 40)
    buf, off = pack_ascii(seqs)
    dev = torch.device("cuda:0")
    bases, meta, nmask = wl.pack_device(torch.from_numpy(buf.copy()).to(dev),
                                        torch.from_numpy(off.view(np.int64).copy()).to(dev))
    res = wl.match_device(bases, meta, nmask, min_score=14, mode=NR_MODE_FILTERED)
    gene = torch.arange(len(seqs), dtype=torch.int32, device=dev) % 7
    for umi_len in (10, 12, 16):
        r = U.records_device(bases, meta, nmask, res, 14, umi_len, gene=gene)
        cc, cl = oracle.encode_many(seqs, 64)
        ref = oracle.match(oracle._CODE[wl_a], 30, 40, cc, cl)
        exp, short, with_n = [], 0, 0
        for i, s in enumerate(seqs):
            if ref["n_best"][i] == 1 and ref["strand"][i] == 0 and ref["best_score"][i] >= 14:
                q = int(ref["umi_q"][i])
                u = s[q:q + umi_len] if q >= 0 else ""
                if len(u) < umi_len:
                    short += 1
                elif "N" in u:
                    with_n += 1
                else:
                    exp.append((int(ref["best_idx"][i]), i % 7, U.pack_umis([u], umi_len)[0][0], i))
        got = list(zip(*(r[k].cpu().numpy().view(np.uint32).tolist() for k in ("bc", "gene", "umi", "src"))))
        assert got == [tuple(int(x) for x in e) for e in exp]
        assert (r["n_records"], r["n_short_umi"], r["n_umi_with_n"]) == (len(exp), short, with_n)
        assert with_n > 0 and (short > 0 or umi_len < 12)


@pytest.mark.parametrize("world", [1, 2, 8, 200])
def test_partition_by_owner(cuda_device, world):
    import torch
    from nanoranger_b200 import umi as U
    rng = np.random.default_rng(world)
    n = 100003
    bc = rng.integers(0, 5000, n).astype(np.uint32)
    gene = rng.integers(0, 9, n).astype(np.uint32)
    um = rng.integers(0, 1 << 24, n).astype(np.uint32)
    dev = torch.device("cuda:0")
    t = [torch.from_numpy(x.view(np.int32)).to(dev) for x in (bc, gene, um)]
    rows, counts = U.partition_device(t[0], t[1], t[2], world)
    rows = rows.cpu().numpy().view(np.uint32)
    own = U.owner_rank(bc, world)
    assert counts.tolist() == np.bincount(own, minlength=world).tolist()
    got_own = U.owner_rank(rows[:, 0], world)
    assert (np.diff(got_own) >= 0).all()                         # ordered by owner
    src = rows[:, 3].astype(np.int64)
    assert np.array_equal(np.sort(src), np.arange(n))            # a permutation of the input
    assert np.array_equal(rows[:, 0], bc[src]) and np.array_equal(rows[:, 1], gene[src])
    assert np.array_equal(rows[:, 2], um[src])
    b2, g2, u2, s2 = (x.cpu().numpy().view(np.uint32) for x in U.unzip_device(
        torch.from_numpy(rows.view(np.int32)).to(dev)))
    assert np.array_equal(np.stack([b2, g2, u2, s2], 1), rows)


def test_umi_collapse_millions_of_records(cuda_device):
    """workspace sizing at C5 scale (CUB's sort storage grows with n) + exact dedup invariants."""
    import torch
    from nanoranger_b200 import umi as U
    rng = np.random.default_rng(77)
    n = 5_000_000
    bc = rng.integers(0, 10000, n).astype(np.uint32)
    gene = rng.integers(0, 2000, n).astype(np.uint32)
    um = rng.integers(0, 1 << 6, n).astype(np.uint32)
    r = U.collapse_host(bc, gene, um, 12, 0)
    key = (bc.astype(np.uint64) << np.uint64(40)) | (gene.astype(np.uint64) << np.uint64(24)) | um
    uk, cnt = np.unique(key, return_counts=True)
    assert r["n_groups"] == len(uk)
    assert int(r["g_reads"].sum()) == n and np.array_equal(r["g_reads"], cnt.astype(np.uint32))
    assert np.array_equal(r["rep_umi"], um)


def test_umi_collapse_huge_group_hash_path(cuda_device, oracle):
    """one (barcode, gene) group with ~15 000 distinct 12-nt UMIs (deeply sequenced cell x highly
    expressed gene, 40 % of the reads carrying a UMI error): the representatives are found through
    the hash set; the answer equals the sequential walk of the oracle."""
    rng = np.random.default_rng(21)
    n = 60000
    true = rng.integers(0, 1 << 24, 3000).astype(np.uint32)
    um = true[(rng.zipf(1.2, n) % 3000)]
    flip = rng.random(n) < 0.4
    um = np.where(flip, um ^ (rng.integers(1, 4, n).astype(np.uint32) << (2 * rng.integers(0, 12, n)).astype(np.uint32)), um)
    bc = np.where(rng.random(n) < 0.9, 5, rng.integers(0, 30, n)).astype(np.uint32)
    gene = np.zeros(n, np.uint32)
    _check(oracle, bc, gene, um.astype(np.uint32), 12, 1)
    _check(oracle, bc, gene, um.astype(np.uint32), 12, 0)


def test_umi_collapse_group_size_boundaries(cuda_device, oracle):
    """groups of exactly 1..10, 31..34, 63..66, 95..98, 300, 2048 and 2049 distinct UMIs (one thread /
    one warp with 1, 2, 3 register slots / one block in shared memory / the grid-wide hash-set
    rounds), drawn from a tight UMI space so that most
    UMIs have neighbours, with many equal read counts (the walk order then falls back on the UMI
    value) and a few dominant UMIs that absorb their neighbours."""
    rng = np.random.default_rng(99)
    sizes = list(range(1, 11)) + [31, 32, 33, 34, 63, 64, 65, 66, 95, 96, 97, 98, 300, 2048, 2049]
    bc, gene, um = [], [], []
    for rep in range(4):
        for gi, nd in enumerate(sizes):
            space = 1 << (2 * (4 if nd <= 98 else (5 if nd <= 300 else 6)))   # 4 / 5 / 6-nt UMIs inside a 12-nt word
            u = rng.choice(space, size=nd, replace=False).astype(np.uint32)
            reads = np.where(rng.random(nd) < 0.15, rng.integers(5, 40, nd), rng.integers(1, 4, nd))
            for x, k in zip(u, reads):
                bc += [rep * 100 + gi] * int(k)
                gene += [gi % 3] * int(k)
                um += [int(x)] * int(k)
    p = rng.permutation(len(bc))
    bc, gene, um = (np.asarray(a, np.uint32)[p] for a in (bc, gene, um))
    _check(oracle, bc, gene, um, 12, 1)
    _check(oracle, bc, gene, um, 12, 0)


def test_umi_collapse_declared_key_widths(cuda_device, oracle):
    """nr_umi_collapse_device_keyed: one 64-bit sort (widths add up to <= 64), two narrower sorts
    (> 64), both equal to the oracle; a record wider than declared is reported, not mis-sorted."""
    from nanoranger_b200 import umi as U
    rng = np.random.default_rng(5)
    n = 200_000
    bc = rng.integers(0, 3000, n).astype(np.uint32)
    gene = rng.integers(0, 500, n).astype(np.uint32)
    um = rng.integers(0, 1 << 10, n).astype(np.uint32) << np.uint32(7)       # 12-nt words, 17 bits used
    k, rep = oracle.umi_cluster(bc, gene, um, 1)
    for widths in (dict(bc_bits=12, gene_bits=9, umi_bits=24),                # 45 bits: one sort
                   dict(bc_bits=32, gene_bits=9, umi_bits=24),                # 65 bits: two sorts
                   dict(bc_bits=12, gene_bits=32, umi_bits=32)):
        r = U.collapse_host(bc, gene, um, 12, 1, **widths)
        assert r["n_groups"] == k and np.array_equal(r["rep_umi"], rep), widths
    full = U.collapse_host(bc, gene, um, 12, 1)
    one = U.collapse_host(bc, gene, um, 12, 1, bc_bits=12, gene_bits=9, umi_bits=24)
    for f in ("g_bc", "g_gene", "g_umi", "g_reads"):
        assert np.array_equal(full[f], one[f]), f
    for widths in (dict(bc_bits=11, gene_bits=9, umi_bits=24), dict(bc_bits=12, gene_bits=8, umi_bits=24),
                   dict(bc_bits=32, gene_bits=8, umi_bits=32)):
        with pytest.raises(ValueError):
            U.collapse_host(bc, gene, um, 12, 1, **widths)
    with pytest.raises(Exception):
        U.collapse_host(bc, gene, um, 12, 1, umi_bits=16)                     # < 2 * umi_len
