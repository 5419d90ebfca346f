#!/usr/bin/env python
"""bench.py -- barcode candidates assigned per second against the 737K whitelist on B200.

One step = one pass of the hot path (ASCII batch resident in HBM -> 2-bit pack -> seed filter,
its N pass for reads with N, the deep tier / brute-force DP for what the filter hands over) over
one batch of synthetic ONT-error-profile 5' flanks per GPU (6 % indel/mismatch, 1e-3 N per base).
Candidates are sharded over ranks, whitelist replicated, no data-path collective ("scaling":
"weak").  After the timed loop every rank also runs, on smaller batches, the other BASELINE
configs so that the driver's records carry them: NR_MODE_AUTO (every score exact), config 5
(Kinnex: match -> UMI records -> partition -> ONE NCCL all-to-all -> UMI collapse) and config 4
(the 3M-sized whitelist); they are sub-objects of the one JSON line.

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
# kernels one step launches (nr_pack_device + nr_match_device in NR_MODE_FILTERED), in stream order
STEP_KERNELS = ("nr_pack_kernel", "nr_match_filtered_kernel<main pass>", "nr_match_filtered_kernel<N pass>",
                "nr_match_deep_kernel<K=3>", "nr_match_deep_kernel<K=5>", "nr_match_bitsliced_kernel<16>",
                "nr_deep_finalize_kernel")
P_N = 1e-3                                       # per-base probability of an N in the synthetic flanks


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
    ap.add_argument("--no-extras", action="store_true",
                    help="skip the auto-mode / kinnex / 3M-list sub-measurements (profiling runs)")
    ap.add_argument("--kinnex-batch", type=int, default=1 << 21, help="sub-reads per GPU per kinnex step")
    ap.add_argument("--wl3m-batch", type=int, default=1 << 20, help="candidates per GPU per 3M-list step")
    ap.add_argument("--p-n", type=float, default=P_N, help="per-base N probability of the synthetic flanks")
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


def kernel_source_hash() -> str:
    """sha256 over the sources of the dominant kernel: profiles/inst_per_candidate.json records the
    hash its ncu capture was taken with, so a kernel edit without a re-capture is visible."""
    import hashlib
    h = hashlib.sha256()
    for f in ("nr_match_filtered.cu", "nr_filter_core.h"):
        h.update(open(os.path.join(ROOT, "nanoranger_b200", "csrc", f), "rb").read())
    return h.hexdigest()[:16]


def inst_per_candidate(workload_key: str):
    """ncu-measured instructions of the matcher kernel per candidate (profiles/)."""
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


def cpu_filtered_rate(wl_ascii, pad_l, pad_r, seqs, offsets, n_sample, threads):
    """the GPU path's own algorithm (seed filter + exact verification, N variants) compiled for
    the host from the same header (oracle/nr_filter_cpu.cpp), all host threads."""
    from oracle import oracle as O
    from nanoranger_b200 import synth
    s = synth.to_strings(seqs, offsets[:n_sample + 1])
    cc, cl = O.encode_many(s, 64)
    t = time.perf_counter()
    f = O.FilterCPU(O._CODE[wl_ascii], pad_l, pad_r)
    t_index = time.perf_counter() - t
    f.match(cc[:256], cl[:256], threads=threads)
    t = time.perf_counter()
    out = f.match(cc, cl, threads=threads)
    dt = time.perf_counter() - t
    f.close()
    return n_sample / dt, dt, t_index, out


def run_reference(args, rank, world):
    """CPU arm: the reference's own implementation of this path is the external STAR binary,
    which is neither in /root/reference nor installed; the arm therefore times the oracle port
    (oracle/nr_oracle.c: the exhaustive scorer of the scoring the reference configures STAR
    with) on all host threads, on bounded samples of the same workload.  When a STAR binary IS on
    PATH, tools/star_concordance.py runs the reference's two invocations on the golden fixtures
    and the line carries the concordance with the oracle and STAR's own timings."""
    if rank != 0:
        return
    from nanoranger_b200 import synth
    wl_ascii, wl_desc = load_whitelist(args.whitelist)
    threads = os.cpu_count() or 1
    S = args.ref_sample
    d = synth.make_candidates(wl_ascii, S * (args.steps + args.warmup), seed=args.seed, p_n=args.p_n)
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
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    import star_concordance as SC
    conc = SC.run(threads=threads)
    line = {
        "impl": "reference", "metric": "barcode_candidates_per_sec", "value": v,
        "unit": "candidates/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * total / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "i8", "data": "synthetic",
        "config": {"workload": f"synthetic-ont6pct-5p-flanks-vs-{args.whitelist}",
                   "whitelist": wl_desc, "pads": [30, 40], "min_score": 14,
                   "sample_per_step": S, "p_n": args.p_n, "star_binary": conc.get("star_binary")},
        "cpu_baseline": {"value": v, "unit": "candidates/s", "cores": threads, "kind": "port",
                         "sample": f"{S} candidates per step x {args.steps} steps, exhaustive "
                                   "both-strand DP against every whitelist entry (oracle/nr_oracle.c)"},
        "e2e": {"value": v, "unit": "candidates/s", "h2d_bytes_per_step": 0,
                "d2h_bytes_per_step": 0},
        "star_concordance": conc,
    }
    if conc.get("status") == "measured" and conc.get("fixtures"):
        # STAR itself on this box: the reference's implementation of the path, timed
        fx = conc["fixtures"]
        n_c = sum(f["n"] for f in fx.values())
        t_al = sum(f["align_s"] for f in fx.values())
        line["cpu_baseline_star"] = {
            "value": n_c / t_al if t_al > 0 else None, "unit": "candidates/s", "cores": threads,
            "kind": "reference",
            "sample": f"STAR on the golden fixtures ({n_c} candidates), index build "
                      f"{max(f['index_build_s'] for f in fx.values()):.1f} s timed separately"}
    # the same algorithm as the GPU path on the host cores, beside the brute force
    n_f = min(len(strs), 20 * S)
    vf, dtf, t_idx, _ = cpu_filtered_rate(wl_ascii, 30, 40, d["seqs"], d["offsets"], n_f, threads)
    line["cpu_baseline_filtered"] = {
        "value": vf, "unit": "candidates/s", "cores": threads, "kind": "port-filtered",
        "sample": f"{n_f} candidates, {dtf:.2f} s; oracle/nr_filter_cpu.cpp = nr_filter_core.h (the "
                  f"GPU kernel's header) compiled for the host; index build {t_idx:.2f} s"}
    print(json.dumps(line), flush=True)


# ---- sub-measurements (every rank runs them; rank 0 keeps the dict) ------------------------------

def all_max(ctx, vals):
    import torch
    import torch.distributed as dist
    t = torch.tensor(vals, dtype=torch.float64, device=ctx["dev"])
    if ctx["world"] > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return [float(x) for x in t.tolist()]


def all_sum(ctx, vals):
    import torch
    import torch.distributed as dist
    t = torch.tensor(vals, dtype=torch.float64, device=ctx["dev"])
    if ctx["world"] > 1:
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return [float(x) for x in t.tolist()]


def barrier(ctx):
    import torch
    import torch.distributed as dist
    if ctx["world"] > 1:
        dist.barrier()
    torch.cuda.synchronize()


def timed_steps(ctx, step, steps, warmup):
    """W untimed, K timed steps bracketed by barrier + synchronize; max over ranks; ms per step."""
    import torch
    for _ in range(warmup):
        step()
    barrier(ctx)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        step()
    e1.record()
    barrier(ctx)
    return all_max(ctx, [e0.elapsed_time(e1) / steps])[0]


def measure_kinnex(args, ctx, batch, steps, warmup):
    """BASELINE config 5: synthetic MAS-ISO-seq/Kinnex concatemers, 16 sub-reads per read, 10x 3'
    GEX geometry (35-nt candidates, pads 4/17, UMI 12 at reference column 20: utils.py:1374,
    1451-1452, 1490-1491), whitelist = the observed cells (write_bc_3p10XGEX keeps raw 16-mers
    with > 20 reads that are on the 737K list).  One step = pack -> match -> records ->
    partition by barcode hash -> ONE all-to-all -> UMI collapse of the owned barcodes."""
    import torch
    from nanoranger_b200 import NR_MODE_FILTERED, Whitelist, synth, whitelists
    from nanoranger_b200 import umi as U
    rank, world, dev = ctx["rank"], ctx["world"], ctx["dev"]
    n_cells, n_genes, sub = 10000, 20000, 16
    rng = np.random.Generator(np.random.PCG64(20180201))
    wl_all = whitelists.load_737k()
    cells = np.sort(rng.choice(len(wl_all), n_cells, replace=False))
    wl_ascii = wl_all[cells]
    wl = Whitelist(wl_ascii, 4, 17, device=ctx["local_rank"])
    reads = batch // sub
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
                              cell_idx=cell, umi_codes=umi_codes, p_n=args.p_n)
    d_seqs = torch.from_numpy(d["seqs"]).to(dev)
    d_off = torch.from_numpy(d["offsets"].view(np.int64)).to(dev)
    d_gene = torch.from_numpy(gene.view(np.int32)).to(dev)
    ws = wl.workspace(B, dev, NR_MODE_FILTERED)
    out = wl.alloc_result(B, dev)
    info, phases = {}, {}

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
        r = U.collapse_device(bc, gene_r, umi_r, 12, args.max_dist, bc_bits=U.key_bits(n_cells),
                              gene_bits=U.key_bits(n_genes), umi_bits=24)
        mark(5)
        info.update(n_records=rec["n_records"], n_groups=r["n_groups"], short=rec["n_short_umi"],
                    received=int(bc.numel()))
        if timed:
            torch.cuda.synchronize()
            for k, nm in enumerate(("pack+match", "records", "partition", "all_to_all", "collapse")):
                phases[nm] = ev[k].elapsed_time(ev[k + 1])

    ms = timed_steps(ctx, step, steps, warmup)
    step(timed=True)                       # one extra step with phase events (not in the metric)
    barrier(ctx)
    ph = all_max(ctx, [phases[k] for k in ("pack+match", "records", "partition", "all_to_all", "collapse")])
    tot = all_sum(ctx, [float(info["n_records"]), float(info["n_groups"])])
    mx = all_max(ctx, [float(info["received"])])
    wl.close()
    if rank != 0:
        return None
    return {
        "value": world * B / (ms * 1e-3), "unit": "sub-reads/s", "ms_per_step": ms, "steps": steps,
        "reads_per_sec": world * reads / (ms * 1e-3),
        "config": {"workload": "synthetic-kinnex16-3p-gex-match+umi-collapse",
                   "subreads_per_gpu_per_step": B, "subreads_per_read": sub,
                   "whitelist": f"{n_cells} observed cells drawn from 737K-august-2016", "pads": [4, 17],
                   "min_score": 14, "umi_len": 12, "umi_max_dist": args.max_dist, "genes": n_genes,
                   "p_n": args.p_n,
                   "exchange": "one variable-count all-to-all of 16 B records (NCCL), split sizes from one "
                               "all-gather of the per-destination counts" if world > 1 else "none (1 GPU)"},
        "umi": {"records": int(tot[0]), "molecules": int(tot[1]),
                "max_records_owned_by_one_rank": int(mx[0])},
        "phases_ms_max_over_ranks": dict(zip(("pack+match", "records", "partition", "all_to_all", "collapse"), ph)),
    }


def measure_wl3m(args, ctx, batch, steps, warmup):
    """BASELINE config 4's whitelist size: 6 794 880 entries (synthetic stand-in, see
    load_whitelist), FILTERED and AUTO, device-resident."""
    import torch
    from nanoranger_b200 import NR_MODE_AUTO, NR_MODE_FILTERED, Whitelist, synth
    rank, world, dev = ctx["rank"], ctx["world"], ctx["dev"]
    wl_ascii, wl_desc = load_whitelist("3M-synthetic")
    t0 = time.perf_counter()
    wl = Whitelist(wl_ascii, 30, 40, device=ctx["local_rank"])
    torch.cuda.synchronize()
    t_index = time.perf_counter() - t0
    d = synth.make_candidates(wl_ascii, batch, seed=args.seed + 31 + 1000 * rank, p_n=args.p_n)
    d_seqs = torch.from_numpy(d["seqs"]).to(dev)
    d_off = torch.from_numpy(d["offsets"].view(np.int64)).to(dev)
    ws = wl.workspace(batch, dev, NR_MODE_AUTO)
    out = wl.alloc_result(batch, dev)

    def mk(mode):
        def step():
            bases, meta, nmask = wl.pack_device(d_seqs, d_off)
            wl.match_device(bases, meta, nmask, min_score=14, mode=mode, out=out, workspace=ws)
        return step

    ms_f = timed_steps(ctx, mk(NR_MODE_FILTERED), steps, warmup)
    assigned = all_sum(ctx, [float(out.assigned(14).sum().item())])[0]
    bases, meta, nmask = wl.pack_device(d_seqs, d_off)
    wl.match_device(bases, meta, nmask, min_score=14, out=out, workspace=ws, counted=True)
    torch.cuda.synchronize()
    counters = wl.counters(ws)
    ms_a = timed_steps(ctx, mk(NR_MODE_AUTO), max(1, steps // 2), 1)
    tiers = wl.tier_counts(ws)
    nbytes = wl.device_bytes
    wl.close()
    if rank != 0:
        return None
    return {
        "value": world * batch / (ms_f * 1e-3), "unit": "candidates/s", "ms_per_step": ms_f, "steps": steps,
        "auto_mode_value": world * batch / (ms_a * 1e-3),
        "assigned_fraction": assigned / (world * batch),
        "config": {"workload": "synthetic-ont6pct-5p-flanks-vs-3M-synthetic", "whitelist": wl_desc,
                   "candidates_per_gpu_per_step": batch, "pads": [30, 40], "min_score": 14, "p_n": args.p_n,
                   "index_build_s": t_index, "index_device_bytes": int(nbytes)},
        "counters_per_candidate": {k: v / batch for k, v in counters.items()},
        "auto_tiers_rank0": tiers,
    }


def measure_slideseq(args, ctx, batch, steps, warmup):
    """BASELINE config 2's whitelist: the reference's slide-seq list (17 753 barcodes of 14 nt,
    2 584 with N, as 8 + linker 18 + 6 = 32 scored columns, pads 15/24, threshold AS >= 30:
    utils.py:584-601, 638), synthetic ONT-profile candidates; FILTERED (anchored seed filter) and
    AUTO (every score exact), device-resident."""
    import gzip
    import torch
    from nanoranger_b200 import NR_MODE_AUTO, NR_MODE_FILTERED, Whitelist, synth
    from nanoranger_b200.whitelists import LINKER_SLIDESEQ
    rank, world, dev = ctx["rank"], ctx["world"], ctx["dev"]
    path = os.path.join(ROOT, "tests", "golden", "slideseq_whitelist.txt.gz")
    bcs = gzip.open(path, "rt").read().split()
    cores = [b[:8] + LINKER_SLIDESEQ + b[8:] for b in bcs]
    wl_ascii = np.frombuffer("".join(cores).encode(), np.uint8).reshape(len(cores), 32)
    wl = Whitelist(wl_ascii, 15, 24, device=ctx["local_rank"])
    d = synth.make_candidates(wl_ascii, batch, seed=args.seed + 53 + 1000 * rank, geometry="slideseq",
                              p_n=args.p_n)
    d_seqs = torch.from_numpy(d["seqs"]).to(dev)
    d_off = torch.from_numpy(d["offsets"].view(np.int64).copy()).to(dev)
    ws = wl.workspace(batch, dev, NR_MODE_AUTO)
    out = wl.alloc_result(batch, dev)

    def mk(mode):
        def step():
            bases, meta, nmask = wl.pack_device(d_seqs, d_off)
            wl.match_device(bases, meta, nmask, min_score=30, mode=mode, out=out, workspace=ws)
        return step

    ms_f = timed_steps(ctx, mk(NR_MODE_FILTERED), steps, warmup)
    assigned = all_sum(ctx, [float(out.assigned(30).sum().item())])[0]
    ms_a = timed_steps(ctx, mk(NR_MODE_AUTO), max(1, steps // 2), 1)
    tiers = wl.tier_counts(ws)
    wl.close()
    if rank != 0:
        return None
    return {
        "value": world * batch / (ms_f * 1e-3), "unit": "candidates/s", "ms_per_step": ms_f, "steps": steps,
        "auto_mode_value": world * batch / (ms_a * 1e-3),
        "assigned_fraction": assigned / (world * batch),
        "config": {"workload": "synthetic-ont6pct-slideseq-candidates-vs-slideseq.matched.barcodes",
                   "whitelist": f"{len(cores)} x 32 columns (8 + linker 18 + 6), the reference's "
                                "data/slideseq.matched.barcodes.tsv.gz", "pads": [15, 24], "min_score": 30,
                   "candidates_per_gpu_per_step": batch, "p_n": args.p_n},
        "auto_tiers_rank0": tiers,
    }


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
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    ctx = {"rank": rank, "local_rank": local_rank, "world": world, "dev": dev}
    try:
        if args.workload == "kinnex":
            sampler = ClockSampler(local_rank)
            if rank == 0:
                sampler.start()
            k = measure_kinnex(args, ctx, args.batch, args.steps, args.warmup)
            clocks = sampler.stop() if rank == 0 else None
            if rank == 0:
                line = {"metric": "barcode_candidates_per_sec", "value": k["value"], "unit": "candidates/s",
                        "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                        "ms_per_step": k["ms_per_step"], "higher_is_better": True, "scaling": "weak",
                        "vs_baseline": None, "dtype": "u32", "data": "synthetic", "config": k["config"],
                        "kinnex": k, "gpu_launches": None, "clocks": clocks}
                print(json.dumps(line), flush=True)
            return
        run_flanks(args, ctx)
    finally:
        if world > 1:
            dist.destroy_process_group()


def run_flanks(args, ctx):
    import torch
    from nanoranger_b200 import NR_MODE_AUTO, NR_MODE_FILTERED, MatchResult, Whitelist, int_peak, synth
    rank, local_rank, world, dev = ctx["rank"], ctx["local_rank"], ctx["world"], ctx["dev"]
    pad_l, pad_r, min_score = 30, 40, 14
    wl_ascii, wl_desc = load_whitelist(args.whitelist)
    t0 = time.perf_counter()
    wl = Whitelist(wl_ascii, pad_l, pad_r, device=local_rank)
    torch.cuda.synchronize()
    t_index = time.perf_counter() - t0
    B = args.batch
    d = synth.make_candidates(wl_ascii, B, seed=args.seed + 1000 * rank, p_n=args.p_n)
    h_seqs, h_off = d["seqs"], d["offsets"]
    d_seqs = torch.from_numpy(h_seqs).to(dev)
    d_off = torch.from_numpy(h_off.view(np.int64)).to(dev)
    n_bytes_in = h_seqs.nbytes + h_off.nbytes
    ws = wl.workspace(B, dev, NR_MODE_AUTO)
    out = wl.alloc_result(B, dev)

    def step():
        bases, meta, nmask = wl.pack_device(d_seqs, d_off)
        wl.match_device(bases, meta, nmask, min_score=min_score, mode=NR_MODE_FILTERED, out=out,
                        workspace=ws)

    for _ in range(args.warmup):
        step()
    barrier(ctx)
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier(ctx)
    e0.record()
    for k in range(args.steps):
        step()
    e1.record()
    barrier(ctx)
    ms_total = e0.elapsed_time(e1)
    clocks = sampler.stop() if rank == 0 else None
    tiers_filtered = wl.tier_counts(ws)

    # the matcher alone (pack excluded): its launches' duration, live, same stream
    bases, meta, nmask = wl.pack_device(d_seqs, d_off)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(args.steps):
        wl.match_device(bases, meta, nmask, min_score=min_score, mode=NR_MODE_FILTERED, out=out,
                        workspace=ws)
    e1.record()
    torch.cuda.synchronize()
    ms_match = e0.elapsed_time(e1) / args.steps
    assigned = int(out.assigned(min_score).sum().item())
    a_np = out.assigned(min_score).cpu().numpy()
    idx_filtered = out.idx.cpu().numpy()
    score_filtered = out.score.cpu().numpy()
    # accuracy against the generator's ground truth (reported, not a parity criterion: the scoring
    # the reference configures lets a read with barcode errors sit closer to a neighbouring entry)
    t_np = d["true_idx"]
    pos = a_np & (t_np >= 0)
    accuracy = {"assigned_to_true_barcode": float((idx_filtered[pos] == t_np[pos]).mean()) if pos.any() else None,
                "negatives_assigned": float(a_np[t_np < 0].mean()) if (t_np < 0).any() else None}
    # live work counters of the same batch (counting build of the kernel)
    wl.match_device(bases, meta, nmask, min_score=min_score, out=out, workspace=ws, counted=True)
    torch.cuda.synchronize()
    counters = wl.counters(ws)

    # end to end through the host-buffer C-ABI call (H2D + pack + match + D2H inside the timed
    # region); inputs and outputs live in pinned host memory, as the contract asks
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

    def e2e_time(mode, seqs_, off_, res_, n_steps):
        wl.match_host(seqs_, off_, min_score=min_score, mode=mode, out=res_)
        barrier(ctx)
        t0 = time.perf_counter()
        for _ in range(n_steps):
            wl.match_host(seqs_, off_, min_score=min_score, mode=mode, out=res_)
        barrier(ctx)
        return (time.perf_counter() - t0) / n_steps

    e2e_s = e2e_time(NR_MODE_FILTERED, e2e_seqs, e2e_off, h_out, e2e_steps)
    same = bool(np.array_equal(h_out.idx, idx_filtered))
    # same call with pageable (plain numpy) buffers: staged through the library's pinned slots
    g_out = MatchResult(np.empty(B, np.int32), np.empty(B, np.int8), np.empty(B, np.uint8),
                        np.empty(B, np.uint8), np.empty(B, np.uint8))
    e2e_pageable_s = e2e_time(NR_MODE_FILTERED, h_seqs, h_off, g_out, e2e_steps)

    # NR_MODE_AUTO: every candidate exact at every score (what utils.barcode_align defaults to)
    auto = None
    if not args.no_extras:
        def step_auto():
            b2, m2, n2 = wl.pack_device(d_seqs, d_off)
            wl.match_device(b2, m2, n2, min_score=min_score, mode=NR_MODE_AUTO, out=out, workspace=ws)
        ms_auto = timed_steps(ctx, step_auto, max(2, args.steps // 3), 1)
        tiers_auto = wl.tier_counts(ws)
        sc = out.score.cpu().numpy()
        same_assigned = bool(np.array_equal(out.idx.cpu().numpy()[a_np], idx_filtered[a_np]) and
                             np.array_equal(sc[a_np], score_filtered[a_np]))
        hist = {int(v): int(c) for v, c in zip(*np.unique(sc, return_counts=True))}
        e2e_auto_s = e2e_time(NR_MODE_AUTO, e2e_seqs, e2e_off, h_out, 2)
        e2e_auto_s = all_max(ctx, [e2e_auto_s])[0]
        auto = {"value": world * B / (ms_auto * 1e-3), "unit": "candidates/s", "ms_per_step": ms_auto,
                "e2e_value": world * B / e2e_auto_s,
                "tiers_rank0": tiers_auto, "score_histogram_rank0": hist,
                "agrees_with_filtered_on_assigned": same_assigned,
                "note": "every candidate resolved exactly at every score (full AS histogram of "
                        "_barcode_scores.csv): seed filter -> deep tier (meet in the middle over the "
                        "whole whitelist) -> brute-force DP for what is left"}

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
        dp_mask = torch.from_numpy(a_np[:n_dp]).to(dev)
        dp = {"value": dp_cells / (ms_dp * 1e-3) / 1e9 * world, "kernel": "nr_match_bitsliced_kernel<16> (bit-parallel, 32 entries per thread; round 1's DPX kernel: 9.7e3 GCUPS)",
              "sample": f"first {n_dp} candidates of the batch per GPU, both strands, every whitelist entry",
              "ms": ms_dp, "candidates_per_sec": world * n_dp / (ms_dp * 1e-3),
              "agrees_with_filtered_on_assigned": bool(np.array_equal(dp_out.idx.cpu().numpy()[a_np[:n_dp]],
                                                                      idx_filtered[:n_dp][a_np[:n_dp]])),
              "note": "executed cell updates of the brute-force DP (NR_MODE_EXHAUSTIVE): the like-for-like "
                      "GPU figure against cpu_baseline (the same brute force on the host cores)"}
        del dp_mask

    ms_total, ms_match, e2e_s, e2e_pageable_s = all_max(ctx, [ms_total, ms_match, e2e_s, e2e_pageable_s])
    tot_assigned = all_sum(ctx, [float(assigned)])[0]
    wl_bytes = wl.device_bytes
    ip = int_peak(local_rank, 2000) if rank == 0 else None
    wl.close()
    del d_seqs, d_off, ws, out
    torch.cuda.empty_cache()

    kin = w3m = sls = None
    if not args.no_extras:
        kin = measure_kinnex(args, ctx, args.kinnex_batch, max(3, args.steps // 2), 2)
        torch.cuda.empty_cache()
        w3m = measure_wl3m(args, ctx, args.wl3m_batch, max(2, args.steps // 3), 1)
        torch.cuda.empty_cache()
        sls = measure_slideseq(args, ctx, 1 << 20, max(2, args.steps // 3), 1)
    if rank != 0:
        return

    ms_step = ms_total / args.steps
    value = world * B / (ms_step * 1e-3)
    e2e_value = world * B / e2e_s
    pk, pk_kind = peaks()
    wkey = f"{args.whitelist}-5p"
    ipc = inst_per_candidate(wkey)
    n_wl = len(wl_ascii)
    lanes_peak = ip["alu_ops_per_s"] / 1e9            # G lane-ops/s the ALU pipe can issue
    src_hash = kernel_source_hash()
    stale = None
    if ipc is not None and ipc.get("alu_warp_inst_per_candidate"):
        # ALU-pipe warp instructions per candidate counted by ncu for this kernel on this workload
        # (profiles/): x 32 lanes = issue slots; the ALU share of the active thread instructions =
        # lanes that really carried work
        lane_ops = ipc["alu_warp_inst_per_candidate"] * 32.0
        thr_per_inst = (ipc["thread_inst_per_candidate"] / ipc["warp_inst_per_candidate"]
                        if ipc.get("thread_inst_per_candidate") and ipc.get("warp_inst_per_candidate") else None)
        ops_src = ipc["source"]
        traffic = ipc.get("dram_bytes_per_candidate")
        traffic = traffic * B if traffic is not None else None
        stale = ipc.get("kernel_source_hash") != src_hash
        # the capture's own work counters against this run's: a drift in probes or verifications
        # per candidate means the per-candidate instruction count no longer describes the kernel
        drift = None
        if ipc.get("probes_per_candidate"):
            drift = {"probes": counters["probes"] / B / ipc["probes_per_candidate"],
                     "verifications": counters["verifications"] / B / max(ipc.get("verifications_per_candidate", 0), 1e-9)}
    else:
        lane_ops = (counters["probes"] * 14 + counters["verifications"] * 28) / B
        thr_per_inst, drift = None, None
        ops_src = "estimate from kernel counters (profiles/inst_per_candidate.json missing)"
        traffic = None
    achieved = lane_ops * B / (ms_match * 1e-3) / 1e9
    alg_bytes = n_bytes_in / B + 25 + 8                # ASCII + offsets in, packed record r/w, results out
    roofline = {
        "bound": "alu", "achieved": achieved, "peak": lanes_peak, "unit": "G lane-ops/s",
        "frac": achieved / lanes_peak, "traffic": traffic,
        "frac_issue_slots": achieved / lanes_peak,
        "frac_active_lanes": (achieved / lanes_peak * thr_per_inst / 32.0) if thr_per_inst else None,
        "definition": "ALU-pipe issue slots used by nr_match_filtered_kernel (main pass + N pass): "
                      "ncu-counted ALU warp instructions per candidate x 32 lanes x candidates/s of the "
                      "matcher timed live with CUDA events on its stream, over the LOP3/SHF issue rate "
                      "nr_int_peak measures on the same GPU in the same run (= 148 SM x 64 lanes x SM "
                      "clock); frac_active_lanes weighs the slots by the ncu-counted active threads per "
                      "instruction (SURVEY 8d: 'x active mask')",
        "peak_source": "nr_int_peak, measured in this run (MEASURED_PEAKS.json has no integer peak)",
        "peak_dual_issue": ip["dual_ops_per_s"] / 1e9,
        "alu_lane_ops_per_candidate": lane_ops, "active_threads_per_instruction": thr_per_inst,
        "per_candidate_source": ops_src,
        "per_candidate_source_stale": stale, "kernel_source_hash": src_hash,
        "work_vs_capture": drift,
        "kernel": "nr_match_filtered_kernel", "kernel_ms_per_launch": ms_match,
        "kernel_share_of_step": ms_match / ms_step,
        "hbm": {"algorithmic_bytes_per_candidate": alg_bytes,
                "achieved_gbs": alg_bytes * B / (ms_step * 1e-3) / 1e9,
                "peak_gbs": pk.get("hbm_gbs"), "peak_kind": pk_kind,
                "frac": alg_bytes * B / (ms_step * 1e-3) / 1e9 / pk.get("hbm_gbs")},
    }
    if dp is not None and rank == 0:
        # the bit-parallel kernel against the same ALU-pipe peak.  A thread executes 13.7 ALU-pipe
        # instructions per word of 32 cells (profiles/r2b_bitsliced_ncu.md: 16.5 instructions per
        # cell word issued, 83 % of them on the ALU pipe; the row loop alone is 11.4: 157 LOP3 per
        # 16 cells + 25 per row for the last column), i.e. 13.7 / 32 lane-operations per cell
        dp["alu_inst_per_32_cells"] = 13.7
        dp["frac_of_alu_peak"] = dp["value"] / world * (13.7 / 32.0) / lanes_peak
    line = {
        "metric": "barcode_candidates_per_sec", "value": value, "unit": "candidates/s",
        "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_step,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u32",
        "data": "synthetic",
        "config": {"workload": f"synthetic-ont6pct-5p-flanks-vs-{args.whitelist}",
                   "whitelist": wl_desc, "candidates_per_gpu_per_step": B, "pads": [pad_l, pad_r],
                   "min_score": min_score,
                   "mode": "NR_MODE_FILTERED: exact for everything the reference keeps (AS >= 14); "
                           "auto_mode below is the same batch with every score exact",
                   "error_profile": f"2% sub / 2% ins / 2% del iid, 10% negatives, UMI 12, N with p = {args.p_n} per base",
                   "l2": "inputs larger than L2 (ASCII batch %.0f MB per GPU)" % (n_bytes_in / 1e6),
                   "index_build_s": t_index, "index_device_bytes": int(wl_bytes)},
        "auto_mode": auto,
        "dp_gcups": dp,
        "assigned_per_sec": tot_assigned / (ms_step * 1e-3),
        "assigned_fraction": tot_assigned / (world * B),
        "accuracy_rank0": accuracy,
        "counters_per_candidate": {k: v / B for k, v in counters.items()},
        "tiers_rank0": tiers_filtered,
        "e2e": {"value": e2e_value, "unit": "candidates/s", "h2d_bytes_per_step": int(n_bytes_in),
                "d2h_bytes_per_step": int(8 * B), "steps": e2e_steps,
                "host_buffers": "pinned", "matches_device_path": same,
                "pageable_buffers_value": world * B / e2e_pageable_s},
        "gpu_launches": len(STEP_KERNELS) * args.steps, "kernels_per_step": list(STEP_KERNELS),
        "clocks": clocks, "roofline": roofline,
        "kinnex": kin, "wl3m": w3m, "slideseq": sls,
    }
    if world == 1 and not args.no_cpu_baseline:
        threads = os.cpu_count() or 1
        n_s = min(args.cpu_sample, B)
        v, dt = cpu_oracle_rate(wl_ascii, pad_l, pad_r, h_seqs, h_off, n_s, threads)
        line["cpu_baseline"] = {
            "value": v, "unit": "candidates/s", "cores": threads, "kind": "port",
            "sample": f"first {n_s} candidates of the batch, {dt:.1f} s; oracle/nr_oracle.c brute-force "
                      "scorer (STAR, the reference's implementation of this path, is not installed); "
                      "like for like with dp_gcups.candidates_per_sec"}
        n_f = min(B, 400000)
        vf, dtf, t_idx, fo = cpu_filtered_rate(wl_ascii, pad_l, pad_r, h_seqs, h_off, n_f, threads)
        took = fo["took"] == 1
        agree = bool(np.array_equal(fo["best_idx"][took & a_np[:n_f]], idx_filtered[:n_f][took & a_np[:n_f]]))
        line["cpu_baseline_filtered"] = {
            "value": vf, "unit": "candidates/s", "cores": threads, "kind": "port-filtered",
            "sample": f"first {n_f} candidates of the batch, {dtf:.2f} s; oracle/nr_filter_cpu.cpp = the GPU "
                      f"kernel's header nr_filter_core.h compiled for the host, std::thread over candidates; "
                      f"index build {t_idx:.2f} s",
            "agrees_with_gpu_on_assigned": agree,
            "gpu_over_same_algorithm_on_cpu": value / vf}
    print(json.dumps(line), flush=True)


if __name__ == "__main__":
    main()
