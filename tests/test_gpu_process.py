"""process_matching_* / make_count_mtx_3p10XGEX on small hand-made SAM files: outputs compared with
the reference's per-record rules restated in the test (utils.py:625-682, 830-979, 1135-1280,
1461-1548).  Needs the GPU only for the UMI dedup kernel."""
import gzip
import json

import numpy as np
import pandas as pd
import pytest

pytestmark = pytest.mark.gpu


def _sam(path, ref_names, ref_len, recs):
    from nanoranger_b200 import samio
    samio.write_sam(path, ref_names, ref_len, recs, header="full")


def _recs(rng, n, ref_names, pad_l, L, umi_len, name_fn):
    """records whose UMI start pairs with reference column pad_l+L through an anchored CIGAR;
    some below threshold, some reverse strand, some too short for a whole UMI"""
    from nanoranger_b200 import samio
    out, truth = [], []
    for i in range(n):
        ri = int(rng.integers(0, len(ref_names)))
        pre = int(rng.integers(0, pad_l + 1))
        umi = "".join("ACGT"[j] for j in rng.integers(0, 4, umi_len))
        tail = "".join("ACGT"[j] for j in rng.integers(0, 4, int(rng.integers(0, 6))))
        seq = "A" * pre + "C" * L + umi + tail
        if i % 11 == 0:
            seq = seq[:pre + L + umi_len - 3]                      # short UMI
        a_s = int(rng.choice([L, L - 1, L - 2, L - 3, L - 6]))
        flag = 16 if i % 13 == 0 else 0
        u = pre + L
        pos, cig = samio.anchored_alignment(len(seq), u if i % 17 else -1, pad_l, L, 40)
        name = name_fn(i)
        out.append((name, flag, ri, pos, cig, seq, a_s))
        truth.append((name, ref_names[ri], seq, u if i % 17 else None, a_s, flag))
    return out, truth


def _expected(truth, thr, umi_len, exact):
    trip = []
    for name, bc, seq, u, a_s, flag in truth:
        if a_s >= thr and flag == 0:
            umi = seq[u:u + umi_len] if u is not None else "N"
            if (len(umi) != umi_len) if exact else (len(umi) < umi_len):
                continue
            trip.append((name, bc, umi))
    return trip


def test_process_matching_5p10XTCR_with_clone_merge(cuda_device, tmp_path):
    from nanoranger_b200 import utils
    rng = np.random.default_rng(3)
    refs = ["".join("ACGT"[j] for j in rng.integers(0, 4, 16)) for _ in range(12)]
    recs, truth = _recs(rng, 400, refs, 30, 16, 12, lambda i: f"r{i // 2}-x_{i}_{i + 9}_0_TRB|T.1_5")
    out = str(tmp_path)
    _sam(f"{out}/s_matching.sam", refs, 86, recs)
    trip = _expected(truth, 14, 12, False)
    names = sorted({t[0] for t in trip})[::2]
    clone = pd.DataFrame({"chains": ["TRB"] * len(names), "cloneId": np.arange(len(names)) % 7}, index=names)
    with gzip.open(f"{out}/s_cloneID_filtered.csv.gz", "wt") as f:
        clone.to_csv(f)
    utils.process_matching_5p10XTCR("s", out)
    got = pd.read_csv(f"{out}/s_clone_bcumi.csv.gz")
    df = pd.DataFrame(trip, columns=["ID", "bc", "umi"]).set_index("ID")
    exp = pd.merge(df, clone, how="inner", left_index=True, right_index=True).sort_values(by=["cloneId", "bc", "umi"])
    assert list(got.columns) == ["bc", "umi", "chains", "cloneId"]
    assert got.values.tolist() == exp.values.tolist() and len(got) > 20
    # dedup table: last record per name wins (utils.py:718), exact-distinct UMIs per barcode
    per = {}
    for n, b, u in trip:
        per[n] = (b, u)
    by = {}
    for b, u in per.values():
        by.setdefault(b, []).append(u)
    ded = pd.read_csv(f"{out}/s_bcumi_dedup.csv", index_col=0)
    assert {b: (len(set(u)), len(u)) for b, u in by.items()} == {b: (int(r.umi_cnt), int(r.read_cnt)) for b, r in ded.iterrows()}
    sc = pd.read_csv(f"{out}/s_barcode_scores.csv")
    fwd = [t[4] for t in truth if t[5] == 0]
    v, c = np.unique(fwd, return_counts=True)
    assert dict(zip(sc.score, sc["count"])) == dict(zip(v.tolist(), c.tolist()))


def test_process_matching_3p10XTCR_nuc_and_slideseq(cuda_device, tmp_path):
    from nanoranger_b200 import utils
    rng = np.random.default_rng(4)
    out = str(tmp_path)
    refs = ["".join("ACGT"[j] for j in rng.integers(0, 4, 16)) for _ in range(9)]
    recs, truth = _recs(rng, 300, refs, 16, 16, 12, lambda i: f"n{i}_s_{i}_{i + 5}_0_TRBV7")
    _sam(f"{out}/a_matching.sam", refs, 60, recs)
    utils.process_matching_3p10XTCR_nuc("a", out)
    trip = _expected(truth, 14, 12, False)
    ded = pd.read_csv(f"{out}/a_bcumi_dedup.csv", index_col=0)
    assert int(ded.read_cnt.sum()) == len({t[0] for t in trip})
    assert (np.diff(ded.umi_cnt.values) <= 0).all()
    # slide-seq: 32-column cores, UMI exactly 9 nt at column 47, threshold 30, merged with cloneID
    refs2 = ["".join("ACGT"[j] for j in rng.integers(0, 4, 14)) for _ in range(9)]
    recs2, truth2 = _recs(rng, 300, refs2, 15, 32, 9, lambda i: f"q{i}_x_{i}_{i + 5}_0_TRAC")
    truth2 = [(n, b, s, u, a + 16, f) for n, b, s, u, a, f in truth2]           # scores around 32
    recs2 = [(n, f, r, p, c, s, a + 16) for n, f, r, p, c, s, a in recs2]
    _sam(f"{out}/b_matching.sam", refs2, 71, recs2)
    trip2 = _expected(truth2, 30, 9, True)
    clone = pd.DataFrame({"chains": "TRA", "cloneId": np.arange(len(trip2)) % 5}, index=[t[0] for t in trip2])
    merged = utils.process_matching_slideseq_XCR("b", out, clone)
    got = pd.read_csv(f"{out}/b_clone_bcumi.csv.gz")
    assert list(got.columns) == ["chains", "cloneId", "bc", "umi"] and len(got) == len(trip2) > 50
    exp = pd.merge(clone, pd.DataFrame(trip2, columns=["ID", "bc", "umi"]).set_index("ID"), how="inner",
                   left_index=True, right_index=True).sort_values(by=["cloneId", "bc", "umi"])
    assert got.values.tolist() == exp.values.tolist() and len(merged) == len(got)


def test_process_matching_3p10XGEX_and_count_matrix(cuda_device, tmp_path):
    import os
    from nanoranger_b200 import utils
    rng = np.random.default_rng(6)
    out = str(tmp_path)
    os.makedirs(f"{out}/split")
    refs = ["".join("ACGT"[j] for j in rng.integers(0, 4, 16)) for _ in range(6)]
    genes = ["ACTB-201|ENST1.1_1200", "GAPDH-202|ENST2.4_900", "novelgene"]
    quads_all = {}
    for part in (1, 2):
        recs, truth = _recs(rng, 250, refs, 4, 16, 12,
                            lambda i: f"m64012_1/{i}/ccs_{i}_{i + 40}_0_{genes[i % 3]}")
        recs = recs + recs[:40]                                             # PCR duplicates: same UMI again
        truth = truth + truth[:40]
        _sam(f"{out}/split/part_{part}_matching.sam", refs, 37, recs)
        utils.process_matching_3p10XGEX(f"part_{part}", f"{out}/split")
        q = json.load(open(f"{out}/split/part_{part}_quads.json"))
        exp = {}
        for name, bc, umi in _expected(truth, 14, 12, False):
            t = "_".join(name.split("/ccs_")[-1].split("_")[3:])
            if "|" in t:
                t = t.split("|")[-1].split("_")[0]
            exp.setdefault(bc, []).append([umi, t])
        assert q == exp
        for b, v in exp.items():
            quads_all.setdefault(b, []).extend(v)
        assert utils.process_matching_3p10XGEX(f"part_{part}", f"{out}/split") is None   # exists, skip
    utils.make_count_mtx_3p10XGEX("s", out)
    bcs = open(f"{out}/s_gex_barcodes.tsv").read().split()
    feats = open(f"{out}/s_gex_features.tsv").read().split()
    lines = gzip.open(f"{out}/s_gex.mtx.gz", "rt").read().splitlines()
    assert lines[0].startswith("%%MatrixMarket") and lines[1].split() == [str(len(feats)), str(len(bcs)), str(len(lines) - 2)]
    got = {(bcs[int(c) - 1], feats[int(r) - 1]): int(v) for r, c, v in (ln.split() for ln in lines[2:])}
    exp = {}
    for b, v in quads_all.items():
        for g in {t for _, t in v}:
            exp[(b, g)] = len({u for u, t in v if t == g})                 # distinct UMIs per (barcode, gene)
    assert got == exp and sorted(feats) == sorted({"ENST1.1", "ENST2.4", "novelgene"})
