"""2+ GPU check of the distributed UMI collapse (run by tests/test_gpu_multi.py through torchrun,
one process per GPU, NCCL): the union of the per-rank group tables equals the single-GPU table."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    from nanoranger_b200 import umi as U
    rng = np.random.default_rng(5)                       # same global data on every rank
    n = 400000
    bc = (rng.zipf(1.3, n) % 4000).astype(np.uint32)
    gene = rng.integers(0, 30, n).astype(np.uint32)
    um = rng.integers(0, 1 << 14, n).astype(np.uint32)
    lo, hi = U.shard_bounds(n, world, rank)
    t = [torch.from_numpy(x[lo:hi].view(np.int32).copy()).to(dev) for x in (bc, gene, um)]
    for max_dist in (0, 1):
        r = U.collapse_distributed(t[0], t[1], t[2], 12, max_dist)
        tab = torch.stack([r["g_bc"], r["g_gene"], r["g_umi"], r["g_reads"]], 1).contiguous()
        assert (U.owner_rank(r["g_bc"].cpu().numpy().view(np.uint32), world) == rank).all()
        sizes = [torch.zeros(1, dtype=torch.int64, device=dev) for _ in range(world)]
        dist.all_gather(sizes, torch.tensor([tab.shape[0]], dtype=torch.int64, device=dev))
        mx = max(int(s.item()) for s in sizes)
        pad = torch.zeros((mx, 4), dtype=torch.int32, device=dev)
        pad[:tab.shape[0]] = tab
        allt = [torch.empty_like(pad) for _ in range(world)]
        dist.all_gather(allt, pad)
        if rank == 0:
            full = U.collapse_device(*[torch.from_numpy(x.view(np.int32)).to(dev) for x in (bc, gene, um)],
                                     12, max_dist)
            exp = torch.stack([full["g_bc"], full["g_gene"], full["g_umi"], full["g_reads"]], 1).cpu().numpy()
            got = np.concatenate([a[:int(s.item())].cpu().numpy() for a, s in zip(allt, sizes)])
            got = got[np.lexsort((got[:, 2].view(np.uint32), got[:, 1].view(np.uint32), got[:, 0].view(np.uint32)))]
            assert np.array_equal(got, exp), f"max_dist {max_dist}: distributed table differs"
    dist.barrier()
    if rank == 0:
        print("dist_umi_nccl ok", world)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
