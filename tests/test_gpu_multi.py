"""Multi-GPU paths on real devices (skipped when the box has one GPU): NCCL all-to-all of the
UMI records and the sharded bench line."""
import json
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _ngpu():
    import torch
    return torch.cuda.device_count()


def _torchrun(n, script, *args, port=29541):
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={n}",
           "--master-addr", "127.0.0.1", "--master-port", str(port), script, *args]
    return subprocess.run(cmd, cwd=ROOT, capture_output=True, text=True, timeout=600)


def test_distributed_umi_collapse_nccl(cuda_device):
    if _ngpu() < 2:
        pytest.skip("needs >= 2 GPUs")
    r = _torchrun(2, os.path.join(ROOT, "tests", "dist_umi_nccl.py"))
    assert r.returncode == 0 and "dist_umi_nccl ok" in r.stdout, r.stderr[-2000:]


def test_bench_two_ranks(cuda_device):
    if _ngpu() < 2:
        pytest.skip("needs >= 2 GPUs")
    r = _torchrun(2, os.path.join(ROOT, "bench.py"), "--gpus", "2", "--steps", "2", "--warmup", "3",
                  "--batch", "262144", "--no-cpu-baseline", port=29542)
    assert r.returncode == 0, r.stderr[-2000:]
    line = [ln for ln in r.stdout.splitlines() if ln.startswith("{")][-1]
    d = json.loads(line)
    assert d["n_gpus"] == 2 and d["value"] > 0 and d["e2e"]["matches_device_path"]
