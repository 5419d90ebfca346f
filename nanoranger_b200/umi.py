"""UMI collapse on the GPU (nr_umi_collapse_device) and its multi-GPU partitioning.

Replaces the exact per-barcode dedup of utils.py:759-777 of the reference (np.unique over the UMI
strings of each barcode) and provides the (barcode, transcript, UMI-cluster) table that
utils.make_count_mtx_3p10XGEX (utils.py:1523-1548) was meant to produce.
"""
from __future__ import annotations

import numpy as np

from . import _lib

_CODE = np.full(256, 255, dtype=np.uint8)
for _i, _c in enumerate(b"ACGT"):
    _CODE[_c] = _i
_ASCII = np.frombuffer(b"ACGT", dtype=np.uint8)


def pack_umis(umis, umi_len: int) -> tuple[np.ndarray, np.ndarray]:
    """list of equal-length ACGT strings -> (u32 codes, ok mask).  Base k at bits 2k.
    UMIs containing anything but ACGT are not packable (ok False)."""
    n = len(umis)
    if n == 0:
        return np.zeros(0, np.uint32), np.zeros(0, bool)
    a = np.frombuffer("".join(umis).encode("ascii"), np.uint8).reshape(n, umi_len)
    c = _CODE[a]
    ok = (c < 4).all(axis=1)
    c = np.where(c < 4, c, 0).astype(np.uint32)
    out = np.zeros(n, np.uint32)
    for k in range(umi_len):
        out |= c[:, k] << np.uint32(2 * k)
    return out, ok


def unpack_umis(codes: np.ndarray, umi_len: int) -> list[str]:
    a = np.empty((len(codes), umi_len), np.uint8)
    for k in range(umi_len):
        a[:, k] = _ASCII[(codes >> np.uint32(2 * k)) & np.uint32(3)]
    return [bytes(r).decode("ascii") for r in a]


def key_bits(n_values: int) -> int:
    """bits that hold every id in [0, n_values)"""
    return max(1, int(n_values - 1).bit_length())


def collapse_device(d_bc, d_gene, d_umi, umi_len: int, max_dist: int = 0, bc_bits: int = 32,
                    gene_bits: int = 32, umi_bits: int = 32):
    """torch uint32-as-int32 tensors on one GPU -> dict of device tensors:
    rep_umi [n], n_groups (int), g_bc / g_gene / g_umi / g_reads [n_groups] sorted by
    (bc, gene, umi).  Stream-ordered; the only synchronisation is reading n_groups.
    bc_bits / gene_bits / umi_bits: declared key widths (key_bits(len(whitelist)),
    key_bits(n_genes), 2 * umi_len when no escape codes are used): the sorts run over those bits
    only; a record that does not fit raises ValueError."""
    import torch
    n = d_bc.numel()
    dev = d_bc.device
    L = _lib.lib()
    rep = torch.empty(n, dtype=torch.int32, device=dev)
    ng = torch.zeros(1, dtype=torch.int64, device=dev)
    g = [torch.empty(max(n, 1), dtype=torch.int32, device=dev) for _ in range(4)]
    ws = torch.empty(max(int(L.nr_umi_workspace_bytes(n)), 1), dtype=torch.uint8, device=dev)
    st = torch.cuda.current_stream(dev).cuda_stream
    with torch.cuda.device(dev):
        _lib.check(L.nr_umi_collapse_device_keyed(
            d_bc.data_ptr(), d_gene.data_ptr(), d_umi.data_ptr(), n, umi_len, max_dist,
            bc_bits, gene_bits, umi_bits,
            rep.data_ptr(), ng.data_ptr(), g[0].data_ptr(), g[1].data_ptr(), g[2].data_ptr(),
            g[3].data_ptr(), ws.data_ptr(), ws.numel(), st), "nr_umi_collapse_device_keyed")
    k = int(ng.item())
    if k < 0:
        raise ValueError(f"collapse_device: a record does not fit the declared key widths "
                         f"(bc_bits={bc_bits}, gene_bits={gene_bits}, umi_bits={umi_bits})")
    return {"rep_umi": rep, "n_groups": k, "g_bc": g[0][:k], "g_gene": g[1][:k],
            "g_umi": g[2][:k], "g_reads": g[3][:k]}


def collapse_host(bc, gene, umi, umi_len: int, max_dist: int = 0, device: int = 0, **widths):
    """numpy in, numpy out (uint32 arrays).  widths: bc_bits / gene_bits / umi_bits of collapse_device."""
    import torch
    dev = torch.device("cuda", device)
    t = [torch.from_numpy(np.ascontiguousarray(x, np.uint32).view(np.int32)).to(dev)
         for x in (bc, gene, umi)]
    r = collapse_device(t[0], t[1], t[2], umi_len, max_dist, **widths)
    out = {"n_groups": r["n_groups"]}
    for k in ("rep_umi", "g_bc", "g_gene", "g_umi", "g_reads"):
        out[k] = r[k].cpu().numpy().view(np.uint32)
    return out


# ---- multi-GPU: all records of one barcode must land on one rank ---------------------------------

def owner_rank(bc: np.ndarray, world: int) -> np.ndarray:
    """hash(barcode idx) % world (Fibonacci hashing so neighbouring indices spread)."""
    h = (bc.astype(np.uint64) * np.uint64(0x9E3779B97F4A7C15)) >> np.uint64(40)
    return (h % np.uint64(world)).astype(np.int64)


def partition_records(bc, gene, umi, world: int):
    """-> (packed [n,3] int64-free uint32 records ordered by owner, send_counts [world])."""
    own = owner_rank(bc, world)
    order = np.argsort(own, kind="stable")
    counts = np.bincount(own, minlength=world).astype(np.int64)
    rec = np.stack([np.asarray(x, np.uint32)[order] for x in (bc, gene, umi)], axis=1)
    return rec, counts


def exchange_records(rec_t, send_counts, group=None):
    """One variable-count all-to-all of (bc, gene, umi[, src]) records (torch int32 [n,3|4] on the
    process group's device: NCCL over NVLink on GPUs, gloo in the CPU tests).
    send_counts: per-destination row counts, a list / array on the host or -- as
    partition_device returns them -- an int64 tensor on the device.  The split sizes both
    directions need come from ONE all-gather of the counts (world x world matrix) and one read of
    it on the host; the record exchange itself is a single all_to_all_single.
    -> received records [m,3|4]."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    sc = torch.as_tensor(send_counts, dtype=torch.int64).to(rec_t.device).contiguous()
    allc = torch.empty(world * world, dtype=torch.int64, device=rec_t.device)
    dist.all_gather_into_tensor(allc, sc, group=group)
    m = allc.view(world, world).tolist()               # the only host synchronisation
    send = [int(x) for x in m[rank]]
    recv = [int(m[r][rank]) for r in range(world)]
    out = torch.empty((sum(recv), rec_t.shape[1]), dtype=rec_t.dtype, device=rec_t.device)
    dist.all_to_all_single(out, rec_t[:sum(send)].contiguous(), output_split_sizes=recv,
                           input_split_sizes=send, group=group)
    return out


def score_histogram(score, nbest, flags, group=None):
    """AS histogram of the records STAR would write with flag 0 (unique best pair on the forward
    strand, any score: what `_barcode_scores.csv` counts, reference utils.py:698, 728-730) as 64
    int64 bins over AS 0..63, summed over the ranks of `group` when torch.distributed is
    initialised (the barcode match needs no other collective).  score / nbest / flags:
    this rank's shard as torch tensors (any device)."""
    import torch
    from ._lib import NR_FLAG_RC, NR_FLAG_TOO_LONG, NR_SCORE_BELOW
    flags = flags.to(torch.int32)
    sc = score.to(torch.int64)
    # NR_FLAG_BELOW alone does not exclude a record (AUTO resolves it, with its true score);
    # unresolved scores of NR_MODE_FILTERED carry NR_SCORE_BELOW and are not records
    keep = (nbest == 1) & ((flags & (NR_FLAG_RC | NR_FLAG_TOO_LONG)) == 0) & (sc != NR_SCORE_BELOW) & (sc >= 0)
    h = torch.bincount(sc[keep].clamp(max=63), minlength=64)[:64]
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(h, group=group)
    return h


def shard_bounds(n: int, world: int, rank: int) -> tuple[int, int]:
    """contiguous candidate shard of a rank (keeps output order)."""
    per = (n + world - 1) // world
    lo = min(n, rank * per)
    return lo, min(n, lo + per)


# ---- device-resident pipeline: matcher results -> records -> (all-to-all) -> collapse -------------

def records_device(bases, meta, nmask, res, min_score: int, umi_len: int, gene=None,
                   with_src: bool = True):
    """Matcher outputs (torch tensors on one GPU, see Whitelist.match_device) -> UMI records of
    the assigned candidates, in candidate order (device twin of the loop in
    utils.process_matching_*, utils.py:697-718).  -> dict(bc, gene, umi, src [k] int32 tensors,
    n_records, n_short_umi, n_umi_with_n).  Reading the counts is the only synchronisation."""
    import torch
    n = meta.numel()
    dev = meta.device
    L = _lib.lib()
    out = [torch.empty(max(n, 1), dtype=torch.int32, device=dev) for _ in range(4)]
    stats = torch.zeros(3, dtype=torch.int64, device=dev)
    ws = torch.empty(int(L.nr_umi_records_workspace_bytes(n)), dtype=torch.uint8, device=dev)
    st = torch.cuda.current_stream(dev).cuda_stream
    with torch.cuda.device(dev):
        _lib.check(L.nr_umi_records_device(
            bases.data_ptr(), meta.data_ptr(), nmask.data_ptr(), res.idx.data_ptr(),
            res.score.data_ptr(), res.nbest.data_ptr(), res.flags.data_ptr(), res.umi_q.data_ptr(),
            gene.data_ptr() if gene is not None else None, n, min_score, umi_len,
            out[0].data_ptr(), out[1].data_ptr(), out[2].data_ptr(),
            out[3].data_ptr() if with_src else None, stats.data_ptr(), ws.data_ptr(), ws.numel(),
            st), "nr_umi_records_device")
    k, short, with_n = (int(x) for x in stats.tolist())
    return {"bc": out[0][:k], "gene": out[1][:k], "umi": out[2][:k], "src": out[3][:k],
            "n_records": k, "n_short_umi": short, "n_umi_with_n": with_n}


def partition_device(d_bc, d_gene, d_umi, world: int, d_src=None):
    """Device records -> ([n,4] int32 rows (bc, gene, umi, src) ordered by owner rank,
    send_counts int64 tensor [world] ON THE DEVICE: no host synchronisation here).  Same owner
    hash as owner_rank()."""
    import torch
    n = d_bc.numel()
    dev = d_bc.device
    L = _lib.lib()
    rows = torch.empty((max(n, 1), 4), dtype=torch.int32, device=dev)
    counts = torch.zeros(world, dtype=torch.int64, device=dev)
    cursor = torch.empty(world, dtype=torch.int64, device=dev)
    st = torch.cuda.current_stream(dev).cuda_stream
    with torch.cuda.device(dev):
        _lib.check(L.nr_umi_partition_device(
            d_bc.data_ptr(), d_gene.data_ptr() if d_gene is not None else None, d_umi.data_ptr(),
            d_src.data_ptr() if d_src is not None else None, n, world, rows.data_ptr(),
            counts.data_ptr(), cursor.data_ptr(), st), "nr_umi_partition_device")
    return rows[:n], counts


def unzip_device(rows):
    """[m,4] int32 rows -> (bc, gene, umi, src) int32 tensors."""
    import torch
    m = rows.shape[0]
    dev = rows.device
    out = [torch.empty(max(m, 1), dtype=torch.int32, device=dev) for _ in range(4)]
    st = torch.cuda.current_stream(dev).cuda_stream
    with torch.cuda.device(dev):
        _lib.check(_lib.lib().nr_umi_unzip_device(
            rows.data_ptr(), m, out[0].data_ptr(), out[1].data_ptr(), out[2].data_ptr(),
            out[3].data_ptr(), st), "nr_umi_unzip_device")
    return tuple(t[:m] for t in out)


def collapse_distributed(d_bc, d_gene, d_umi, umi_len: int, max_dist: int = 0, group=None, **widths):
    """UMI collapse over all ranks of `group` (one process per GPU, NCCL): partition the local
    records by owner rank on the device, ONE variable-count all-to-all of 16-byte rows, local
    collapse of the barcodes this rank owns.  -> collapse_device() dict for the owned barcodes.
    With no process group (single GPU) this is collapse_device().  widths: bc_bits / gene_bits /
    umi_bits of collapse_device."""
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return collapse_device(d_bc, d_gene, d_umi, umi_len, max_dist, **widths)
    world = dist.get_world_size(group)
    rows, counts = partition_device(d_bc, d_gene, d_umi, world)
    got = exchange_records(rows, counts, group)
    bc, gene, umi, _ = unzip_device(got)
    return collapse_device(bc, gene, umi, umi_len, max_dist, **widths)
