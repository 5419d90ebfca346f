#!/usr/bin/env python
"""bench.py -- barcode candidates assigned per second against the 737K whitelist on B200.

One step = one pass of the hot path (ASCII batch resident in HBM -> 2-bit pack -> filtered
matcher + exhaustive fallback for the candidates the filter hands over) over one batch of
synthetic ONT-error-profile 5' flanks per GPU.  Candidates are sharded over ranks, whitelist
replicated, no data-path collective ("scaling": "weak").

  python bench.py [--gpus N] [--steps K] [--warmup W]            # our arm
  python bench.py --impl reference ...                          # CPU arm (oracle port; STAR absent)

Under torchrun (N > 1) every rank runs its shard; rank 0 prints ONE JSON line.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

CELLS_PER_CANDIDATE_737K = 737280 * 16 * 50      # SURVEY.md section 8d "equivalent cell updates"
KERNELS_PER_STEP = 3                             # nr_pack_kernel, nr_match_filtered_kernel, nr_match_exhaustive16_kernel


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=1 << 22, help="candidates per GPU per step")
    ap.add_argument("--whitelist", default="737K", choices=["737K", "3M-synthetic"])
    ap.add_argument("--cpu-sample", type=int, default=8000)
    ap.add_argument("--ref-sample", type=int, default=1024, help="candidates per step of the CPU arm")
    ap.add_argument("--seed", type=int, default=2)
    ap.add_argument("--workload", default="flanks", choices=["flanks", "kinnex"],
                    help="flanks: BASELINE metric config (5' flanks vs 737K); kinnex: config 5, 16 "
                         "sub-reads per read, 3' geometry, barcode match + UMI collapse")
    ap.add_argument("--max-dist", type=int, default=1, help="kinnex: UMI clustering distance")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-dp-gcups", action="store_true", help="skip the exhaustive-kernel GCUPS sample")
    return ap.parse_args()


class ClockSampler:
    """nvidia-smi clocks and throttle reasons during the timed region (B200_PROFILING.md)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.gpu = gpu_index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                 "-lms", "100", "-i", str(self.gpu)], stdout=subprocess.PIPE,
                stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for ln in self.proc.stdout:
            self.lines.append(ln.strip())

    def stop(self) -> dict:
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=3)
        except Exception:
            self.proc.kill()
        sm, smax, reasons, power = [], [], set(), []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); smax.append(float(f[2])); power.append(float(f[3]))
            except ValueError:
                continue
            for nm, v in zip(names, f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None,
                "sm_max_mhz": float(max(smax)) if smax else None,
                "power_w_max": float(max(power)) if power else None,
                "samples": len(sm), "reasons": sorted(reasons)}


def load_whitelist(name: str):
    from nanoranger_b200 import whitelists
    if name == "737K":
        return whitelists.load_737k(), "737K-august-2016 (real list, 737280 x 16 nt)"
    return (whitelists.synthetic_whitelist(6794880),
            "synthetic stand-in for 3M-february-2018 (6794880 x 16 nt, min Hamming 2, seed 20180201; "
            "the real file is missing from the reference checkout)")


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return json.load(open(p)), "measured"
    return {"hbm_gbs": 6650.0}, "fallback"


def inst_per_candidate(workload_key: str):
    """ncu-measured thread instructions of the matcher kernels per candidate (profiles/)."""
    p = os.path.join(ROOT, "profiles", "inst_per_candidate.json")
    if os.path.exists(p):
        d = json.load(open(p))
        if workload_key in d:
            return d[workload_key]
    return None


def cpu_oracle_rate(wl_ascii, pad_l, pad_r, seqs, offsets, n_sample, threads):
    from oracle import oracle as O
    from nanoranger_b200 import synth
    s = synth.to_strings(seqs, offsets[:n_sample + 1])
    cc, cl = O.encode_many(s, 64)
    wlc = O._CODE[wl_ascii]
    O.match(wlc, pad_l, pad_r, cc[:8], cl[:8], threads=threads)     # warm (build, page in)
    t = time.perf_counter()
    O.match(wlc, pad_l, pad_r, cc, cl, threads=threads)
    dt = time.perf_counter() - t
    return n_sample / dt, dt


def run_reference(args, rank, world):
    """CPU arm: the reference's own implementation of this path is the external STAR binary,
    which is neither in /root/reference nor installed; the arm therefore times the oracle port
    (oracle/nr_oracle.c: the exhaustive scorer of the scoring the reference configures STAR
    with) on all host threads, on bounded samples of the same workload."""
    if rank != 0:
        return
    from nanoranger_b200 import synth
    wl_ascii, wl_desc = load_whitelist(args.whitelist)
    threads = os.cpu_count() or 1
    S = args.ref_sample
    d = synth.make_candidates(wl_ascii, S * (args.steps + args.warmup), seed=args.seed)
    from oracle import oracle as O
    strs = synth.to_strings(d["seqs"], d["offsets"])
    wlc = O._CODE[wl_ascii]
    times = []
    for k in range(args.steps + args.warmup):
        cc, cl = O.encode_many(strs[k * S:(k + 1) * S], 64)
        t = time.perf_counter()
        O.match(wlc, 30, 40, cc, cl, threads=threads)
        dt = time.perf_counter() - t
        if k >= args.warmup:
            times.append(dt)
    total = sum(times)
    v = S * args.steps / total
    star = "present" if any(os.access(os.path.join(p, "STAR"), os.X_OK)
                            for p in os.environ.get("PATH", "").split(":")) else "absent"
    line = {
        "impl": "reference", "metric": "barcode_candidates_per_sec", "value": v,
        "unit": "candidates/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * total / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "i8", "data": "synthetic",
        "config": {"workload": f"synthetic-ont6pct-5p-flanks-vs-{args.whitelist}",
                   "whitelist": wl_desc, "pads": [30, 40], "min_score": 14,
                   "sample_per_step": S, "star_binary": star},
        "cpu_baseline": {"value": v, "unit": "candidates/s", "cores": threads, "kind": "port",
                         "sample": f"{S} candidates per step x {args.steps} steps, exhaustive "
                                   "both-strand DP against every whitelist entry"},
        "e2e": {"value": v, "unit": "candidates/s", "h2d_bytes_per_step": 0,
                "d2h_bytes_per_step": 0},
        "gcups_equivalent": v * len(wl_ascii) * 16 * 50 / 1e9,
    }
    print(json.dumps(line), flush=True)


def run_kinnex(args, rank, local_rank, world):
    """BASELINE config 5: synthetic MAS-ISO-seq/Kinnex concatemers, 16 sub-reads per read, 10x 3'
    GEX geometry (35-nt candidates, pads 4/17, UMI 12 at reference column 20: utils.py:1374,
    1451-1452, 1490-1491), whitelist = the observed cells (write_bc_3p10XGEX keeps raw 16-mers
    with > 20 reads that are on the 737K list).  One step = pack -> match -> records ->
    partition by barcode hash -> ONE all-to-all -> UMI collapse of the owned barcodes."""
    import torch
    import torch.distributed as dist
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    from nanoranger_b200 import NR_MODE_FILTERED, Whitelist, synth, whitelists
    from nanoranger_b200 import umi as U
    n_cells, n_genes, sub = 10000, 20000, 16
    rng = np.random.Generator(np.random.PCG64(20180201))
    wl_all = whitelists.load_737k()
    cells = np.sort(rng.choice(len(wl_all), n_cells, replace=False))
    wl_ascii = wl_all[cells]
    wl = Whitelist(wl_ascii, 4, 17, device=local_rank)
    reads = args.batch // sub
    B = reads * sub
    r2 = np.random.Generator(np.random.PCG64(args.seed + 1000 * rank))
    w = 1.0 / np.arange(1, n_cells + 1) ** 0.8
    cell = r2.choice(n_cells, size=B, p=w / w.sum())
    gene = (r2.zipf(1.4, B) % n_genes).astype(np.uint32)
    mol = r2.integers(0, 6, B).astype(np.uint64)                 # few molecules per (cell, gene): PCR duplicates
    h = (cell.astype(np.uint64) * np.uint64(0x9E3779B97F4A7C15) ^ gene.astype(np.uint64) * np.uint64(0xC2B2AE3D27D4EB4F)
         ^ mol * np.uint64(0x165667B19E3779F9))
    h ^= h >> np.uint64(29)
    umi_codes = np.stack([((h >> np.uint64(2 * k + 7)) & np.uint64(3)).astype(np.uint8) for k in range(12)], 1)
    d = synth.make_candidates(wl_ascii, B, seed=args.seed + 7 + 1000 * rank, geometry="3p", umi_len=12,
                              cell_idx=cell, umi_codes=umi_codes)
    d_seqs = torch.from_numpy(d["seqs"]).to(dev)
    d_off = torch.from_numpy(d["offsets"].view(np.int64)).to(dev)
    d_gene = torch.from_numpy(gene.view(np.int32)).to(dev)
    ws = wl.workspace(B, dev, NR_MODE_FILTERED)
    out = wl.alloc_result(B, dev)
    info = {}

    phases = {}

    def step(timed=False):
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(6)] if timed else None
        mark = (lambda k: ev[k].record()) if timed else (lambda k: None)
        mark(0)
        bases, meta, nmask = wl.pack_device(d_seqs, d_off)
        wl.match_device(bases, meta, nmask, min_score=14, mode=NR_MODE_FILTERED, out=out, workspace=ws)
        mark(1)
        rec = U.records_device(bases, meta, nmask, out, 14, 12, gene=d_gene, with_src=False)
        mark(2)
        if world > 1:
            rows, counts = U.partition_device(rec["bc"], rec["gene"], rec["umi"], world)
            mark(3)
            got = U.exchange_records(rows, counts)
            mark(4)
            bc, gene_r, umi_r, _ = U.unzip_device(got)
        else:
            mark(3)
            mark(4)
            bc, gene_r, umi_r = rec["bc"], rec["gene"], rec["umi"]
        r = U.collapse_device(bc, gene_r, umi_r, 12, args.max_dist)
        mark(5)
        info.update(n_records=rec["n_records"], n_groups=r["n_groups"], short=rec["n_short_umi"])
        if timed:
            torch.cuda.synchronize()
            for k, nm in enumerate(("pack+match", "records", "partition", "all_to_all", "collapse")):
                phases[nm] = ev[k].elapsed_time(ev[k + 1])

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        step()
    barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for _ in range(args.steps):
        step()
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1) / args.steps
    clocks = sampler.stop() if rank == 0 else None
    step(timed=True)                       # one extra, untimed-for-the-metric step with phase events
    barrier()
    t = torch.tensor([ms], dtype=torch.float64, device=dev)
    tot = torch.tensor([float(info["n_records"]), float(info["n_groups"])], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(tot, op=dist.ReduceOp.SUM)
    if rank == 0:
        ms = float(t.item())
        line = {
            "metric": "barcode_candidates_per_sec", "value": world * B / (ms * 1e-3), "unit": "candidates/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u32",
            "data": "synthetic",
            "config": {"workload": "synthetic-kinnex16-3p-gex-match+umi-collapse",
                       "reads_per_gpu_per_step": reads, "subreads_per_read": sub,
                       "whitelist": f"{n_cells} observed cells drawn from 737K-august-2016", "pads": [4, 17],
                       "min_score": 14, "umi_len": 12, "umi_max_dist": args.max_dist, "genes": n_genes,
                       "exchange": "one variable-count all-to-all of 16 B records (NCCL)" if world > 1 else "none (1 GPU)",
                       "l2": "inputs larger than L2 (%.0f MB ASCII per GPU)" % (d["seqs"].nbytes / 1e6)},
            "reads_per_sec": world * reads / (ms * 1e-3),
            "umi": {"records": int(tot[0].item()), "molecules": int(tot[1].item()),
                    "short_umi_rank0": info["short"]},
            "phases_ms_rank0": phases,
            "gpu_launches": None, "clocks": clocks,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    args = parse()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import torch
    import torch.distributed as dist
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py needs a CUDA device (no CPU fallback exists)")
    if args.workload == "kinnex":
        run_kinnex(args, rank, local_rank, world)
        return
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    from nanoranger_b200 import NR_MODE_FILTERED, Whitelist, int_peak, synth
    from nanoranger_b200 import _lib as K

    pad_l, pad_r, min_score = 30, 40, 14
    wl_ascii, wl_desc = load_whitelist(args.whitelist)
    t0 = time.perf_counter()
    wl = Whitelist(wl_ascii, pad_l, pad_r, device=local_rank)
    torch.cuda.synchronize()
    t_index = time.perf_counter() - t0
    B = args.batch
    d = synth.make_candidates(wl_ascii, B, seed=args.seed + 1000 * rank)
    h_seqs, h_off = d["seqs"], d["offsets"]
    d_seqs = torch.from_numpy(h_seqs).to(dev)
    d_off = torch.from_numpy(h_off.view(np.int64)).to(dev)
    n_bytes_in = h_seqs.nbytes + h_off.nbytes
    ws = wl.workspace(B, dev, NR_MODE_FILTERED)
    out = wl.alloc_result(B, dev)

    def step():
        bases, meta, nmask = wl.pack_device(d_seqs, d_off)
        wl.match_device(bases, meta, nmask, min_score=min_score, mode=NR_MODE_FILTERED, out=out,
                        workspace=ws)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        step()
    barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps + 1)]
    barrier()
    ev[0].record()
    for k in range(args.steps):
        step()
        ev[k + 1].record()
    barrier()
    ms_total = ev[0].elapsed_time(ev[-1])
    clocks = sampler.stop() if rank == 0 else None

    # per-kernel time of the dominant kernel: counted run tells how much work each stage did;
    # a separate timed loop over the matcher alone (pack excluded) gives its launch duration
    bases, meta, nmask = wl.pack_device(d_seqs, d_off)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        wl.match_device(bases, meta, nmask, min_score=min_score, mode=NR_MODE_FILTERED, out=out,
                        workspace=ws)
    e1.record()
    torch.cuda.synchronize()
    ms_match = e0.elapsed_time(e1) / args.steps
    wl.match_device(bases, meta, nmask, min_score=min_score, out=out, workspace=ws, counted=True)
    torch.cuda.synchronize()
    counters = wl.counters(ws)
    assigned = int(out.assigned(min_score).sum().item())
    # accuracy against the generator's ground truth (reported, not a parity criterion: the scoring
    # the reference configures lets a read with barcode errors sit closer to a neighbouring entry)
    a_np = out.assigned(min_score).cpu().numpy()
    t_np = d["true_idx"]
    pos = a_np & (t_np >= 0)
    accuracy = {"assigned_to_true_barcode": float((out.idx.cpu().numpy()[pos] == t_np[pos]).mean()) if pos.any() else None,
                "negatives_assigned": float(a_np[t_np < 0].mean()) if (t_np < 0).any() else None}

    # end to end through the host-buffer C-ABI call (H2D + pack + match + D2H inside the timed
    # region); inputs and outputs live in pinned host memory, as the contract asks
    from nanoranger_b200 import MatchResult

    def pinned(a):
        t = torch.empty(a.shape, dtype=torch.from_numpy(a[:0]).dtype, pin_memory=True)
        t.numpy()[...] = a
        return t

    p_seqs, p_off = pinned(h_seqs), pinned(h_off.view(np.int64))
    p_out = [torch.empty(B, dtype=dt, pin_memory=True)
             for dt in (torch.int32, torch.int8, torch.uint8, torch.uint8, torch.uint8)]
    h_out = MatchResult(*(t.numpy() for t in p_out))
    e2e_steps = max(2, min(args.steps, 5))
    e2e_seqs, e2e_off = p_seqs.numpy(), p_off.numpy().view(np.uint64)
    wl.match_host(e2e_seqs, e2e_off, min_score=min_score, mode=NR_MODE_FILTERED, out=h_out)
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        wl.match_host(e2e_seqs, e2e_off, min_score=min_score, mode=NR_MODE_FILTERED, out=h_out)
    barrier()
    e2e_s = (time.perf_counter() - t0) / e2e_steps
    same = bool(np.array_equal(h_out.idx, out.idx.cpu().numpy()))
    # same call with pageable (plain numpy) buffers: staged through the library's pinned slots
    g_out = MatchResult(np.empty(B, np.int32), np.empty(B, np.int8), np.empty(B, np.uint8),
                        np.empty(B, np.uint8), np.empty(B, np.uint8))
    wl.match_host(h_seqs, h_off, min_score=min_score, mode=NR_MODE_FILTERED, out=g_out)
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        wl.match_host(h_seqs, h_off, min_score=min_score, mode=NR_MODE_FILTERED, out=g_out)
    e2e_pageable_s = (time.perf_counter() - t0) / e2e_steps

    # the exhaustive DP kernel on a bounded sample: real (executed) cell updates per second
    dp = None
    if not args.no_dp_gcups:
        from nanoranger_b200 import NR_MODE_EXHAUSTIVE
        n_dp = min(B, 2048)
        dp_off = d_off[:n_dp + 1].contiguous()
        dp_bases, dp_meta, dp_nmask = wl.pack_device(d_seqs, dp_off)
        dp_out = wl.alloc_result(n_dp, dev)
        dp_ws = wl.workspace(n_dp, dev, NR_MODE_EXHAUSTIVE)
        wl.match_device(dp_bases, dp_meta, dp_nmask, min_score=min_score, mode=NR_MODE_EXHAUSTIVE,
                        out=dp_out, workspace=dp_ws)
        torch.cuda.synchronize()
        e0.record()
        wl.match_device(dp_bases, dp_meta, dp_nmask, min_score=min_score, mode=NR_MODE_EXHAUSTIVE,
                        out=dp_out, workspace=dp_ws)
        e1.record()
        torch.cuda.synchronize()
        ms_dp = e0.elapsed_time(e1)
        dp_cells = float(dp_meta.to(torch.int64).bitwise_and(0x7F).sum().item()) * len(wl_ascii) * 16 * 2
        dp_mask = out.assigned(min_score)[:n_dp]
        dp = {"value": dp_cells / (ms_dp * 1e-3) / 1e9 * world, "kernel": "nr_match_exhaustive16_kernel",
              "sample": f"first {n_dp} candidates of the batch per GPU, both strands, every whitelist entry",
              "ms": ms_dp, "candidates_per_sec": world * n_dp / (ms_dp * 1e-3),
              "agrees_with_filtered_on_assigned": bool(torch.equal(dp_out.idx[dp_mask], out.idx[:n_dp][dp_mask])),
              "note": "executed cell updates of the exhaustive DP (NR_MODE_EXHAUSTIVE); "
                      "gcups_equivalent is the filtered path's candidates/s x cells a brute force would do"}

    # max over ranks
    t = torch.tensor([ms_total, ms_match, e2e_s], dtype=torch.float64, device=dev)
    tot = torch.tensor([float(assigned)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(tot, op=dist.ReduceOp.SUM)
    ms_total, ms_match, e2e_s = (float(x) for x in t.tolist())
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    ms_step = ms_total / args.steps
    value = world * B / (ms_step * 1e-3)
    e2e_value = world * B / e2e_s
    pk, pk_kind = peaks()
    ip = int_peak(local_rank, 2000)
    wkey = f"{args.whitelist}-5p"
    ipc = inst_per_candidate(wkey)
    n_wl = len(wl_ascii)
    lanes_peak = ip["alu_ops_per_s"] / 1e9            # G lane-ops/s the ALU pipe can issue
    if ipc is not None and ipc.get("alu_warp_inst_per_candidate"):
        # ALU-pipe warp instructions per candidate counted by ncu for this kernel on this
        # workload (profiles/), x 32 lanes an issued warp instruction occupies, x live rate
        lane_ops = ipc["alu_warp_inst_per_candidate"] * 32.0
        thread_ops = ipc.get("thread_inst_per_candidate")
        ops_src = ipc["source"]
        traffic = ipc.get("dram_bytes_per_candidate")
        traffic = traffic * B if traffic is not None else None
    else:
        # fallback: 14 warp-wide instructions per probe slot, 28 per verified row, from SASS
        lane_ops = (counters["probes"] * 14 + counters["verifications"] * 28) / B
        thread_ops = None
        ops_src = "estimate from kernel counters (profiles/inst_per_candidate.json missing)"
        traffic = None
    achieved = lane_ops * B / (ms_match * 1e-3) / 1e9
    alg_bytes = n_bytes_in / B + 25 + 8                # ASCII + offsets in, packed record r/w, results out
    roofline = {
        "bound": "alu", "achieved": achieved, "peak": lanes_peak, "unit": "G lane-ops/s",
        "frac": achieved / lanes_peak, "traffic": traffic,
        "definition": "ALU-pipe issue slots used by nr_match_filtered_kernel: ncu-counted ALU warp "
                      "instructions per candidate x 32 lanes x candidates/s of the kernel timed live "
                      "with CUDA events, over the LOP3/SHF issue rate nr_int_peak measures on the same "
                      "GPU in the same run (= 148 SM x 64 lanes x SM clock)",
        "peak_source": "nr_int_peak, measured in this run (MEASURED_PEAKS.json has no integer peak)",
        "peak_dual_issue": ip["dual_ops_per_s"] / 1e9,
        "alu_lane_ops_per_candidate": lane_ops, "active_thread_inst_per_candidate": thread_ops,
        "per_candidate_source": ops_src,
        "kernel": "nr_match_filtered_kernel", "kernel_ms_per_launch": ms_match,
        "kernel_share_of_step": ms_match / ms_step,
        "hbm": {"algorithmic_bytes_per_candidate": alg_bytes,
                "achieved_gbs": alg_bytes * B / (ms_step * 1e-3) / 1e9,
                "peak_gbs": pk.get("hbm_gbs"), "peak_kind": pk_kind,
                "frac": alg_bytes * B / (ms_step * 1e-3) / 1e9 / pk.get("hbm_gbs")},
    }
    line = {
        "metric": "barcode_candidates_per_sec", "value": value, "unit": "candidates/s",
        "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_step,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u32",
        "data": "synthetic",
        "config": {"workload": f"synthetic-ont6pct-5p-flanks-vs-{args.whitelist}",
                   "whitelist": wl_desc, "candidates_per_gpu_per_step": B, "pads": [pad_l, pad_r],
                   "min_score": min_score, "mode": "filtered+exhaustive-fallback",
                   "error_profile": "2% sub / 2% ins / 2% del iid, 10% negatives, UMI 12",
                   "l2": "inputs larger than L2 (ASCII batch %.0f MB per GPU)" % (n_bytes_in / 1e6),
                   "index_build_s": t_index},
        "gcups_equivalent": value * n_wl * 16 * 50 / 1e9,
        "dp_gcups": dp,
        "assigned_per_sec": float(tot.item()) / (ms_step * 1e-3),
        "assigned_fraction": float(tot.item()) / (world * B),
        "accuracy_rank0": accuracy,
        "counters_per_candidate": {k: v / B for k, v in counters.items()},
        "e2e": {"value": e2e_value, "unit": "candidates/s", "h2d_bytes_per_step": int(n_bytes_in),
                "d2h_bytes_per_step": int(8 * B), "steps": e2e_steps,
                "host_buffers": "pinned", "matches_device_path": same,
                "pageable_buffers_value": world * B / e2e_pageable_s},
        "gpu_launches": KERNELS_PER_STEP * args.steps,
        "clocks": clocks, "roofline": roofline,
    }
    if world == 1 and not args.no_cpu_baseline:
        threads = os.cpu_count() or 1
        n_s = min(args.cpu_sample, B)
        v, dt = cpu_oracle_rate(wl_ascii, pad_l, pad_r, h_seqs, h_off, n_s, threads)
        line["cpu_baseline"] = {
            "value": v, "unit": "candidates/s", "cores": threads, "kind": "port",
            "sample": f"first {n_s} candidates of the batch, {dt:.1f} s; oracle/nr_oracle.c exhaustive "
                      "scorer (STAR, the reference's implementation of this path, is not installed)"}
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
