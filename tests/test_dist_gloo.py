"""N > 1 host logic on CPU: world_size 2 over gloo.  Candidate shards cover the batch in order;
the UMI exchange puts every record of a barcode on one rank (one variable-count all-to-all)."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, tmp):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from nanoranger_b200 import umi
    from oracle import oracle as O
    rng = np.random.default_rng(42)                       # same global data on every rank
    n = 6001
    bc = rng.integers(0, 300, n).astype(np.uint32)
    gene = rng.integers(0, 4, n).astype(np.uint32)
    um = rng.integers(0, 64, n).astype(np.uint32)
    lo, hi = umi.shard_bounds(n, world, rank)             # this rank's candidate shard
    rec, counts = umi.partition_records(bc[lo:hi], gene[lo:hi], um[lo:hi], world)
    got = umi.exchange_records(torch.from_numpy(rec.view(np.int32)), counts).numpy().view(np.uint32)
    assert (umi.owner_rank(got[:, 0], world) == rank).all()
    # every record of the barcodes this rank owns arrived, nothing else
    mine = umi.owner_rank(bc, world) == rank
    exp = np.stack([bc[mine], gene[mine], um[mine]], axis=1)
    key = lambda a: a[np.lexsort((a[:, 2], a[:, 1], a[:, 0]))]
    assert np.array_equal(key(got), key(exp))
    # local collapse (CPU oracle twin here; the GPU kernel is checked against it in -m gpu tests)
    k_local, _ = O.umi_cluster(got[:, 0], got[:, 1], got[:, 2], 1)
    t = torch.tensor([k_local], dtype=torch.int64)
    dist.all_reduce(t)
    k_global, _ = O.umi_cluster(bc, gene, um, 1)
    assert int(t.item()) == k_global                      # partitioning by barcode loses nothing
    # AS histogram of the flag-0 records, summed over the ranks (the match's only collective)
    from nanoranger_b200._lib import NR_FLAG_RC, NR_SCORE_BELOW
    score = rng.integers(5, 17, n).astype(np.int8)
    score[::13] = NR_SCORE_BELOW
    nbest = rng.integers(1, 3, n).astype(np.uint8)
    flags = np.where(rng.random(n) < 0.1, NR_FLAG_RC, 0).astype(np.uint8)
    h = umi.score_histogram(torch.from_numpy(score[lo:hi]), torch.from_numpy(nbest[lo:hi]),
                            torch.from_numpy(flags[lo:hi])).numpy()
    keep = (nbest == 1) & (flags == 0) & (score != NR_SCORE_BELOW)
    assert np.array_equal(h, np.bincount(score[keep].astype(np.int64), minlength=64))
    np.save(os.path.join(tmp, f"ok{rank}.npy"), np.array([hi - lo]))
    dist.destroy_process_group()


def test_world2_shard_and_exchange(tmp_path):
    port = _free_port()
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    sizes = [int(np.load(tmp_path / f"ok{r}.npy")[0]) for r in range(2)]
    assert sum(sizes) == 6001
