"""Randomised parity campaign: filtered / auto kernels vs the CPU oracle over many seeds and
geometries (tools, not part of the test suite: ~minutes on a GPU box).
  python tools/stress_parity.py [n_rounds]"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
from helpers import compare, mixed_candidates, rs, tie_rich_whitelist
from nanoranger_b200 import NR_MODE_AUTO, NR_MODE_FILTERED, MatchResult, Whitelist, pack_ascii, synth, whitelists
from oracle import oracle as O

rounds = int(sys.argv[1]) if len(sys.argv) > 1 else 60
dev = torch.device("cuda:0")


def run(wl, seqs, mode):
    buf, off = pack_ascii(seqs)
    b = wl.pack_device(torch.from_numpy(buf.copy()).to(dev), torch.from_numpy(off.view(np.int64).copy()).to(dev))
    r = wl.match_device(*b, min_score=14, mode=mode)
    torch.cuda.synchronize()
    return MatchResult(*(t.cpu().numpy() for t in (r.idx, r.score, r.nbest, r.flags, r.umi_q)))


t0 = time.time()
tot = 0
for it in range(rounds):
    rng = np.random.default_rng(10_000 + it)
    pad_l, pad_r = int(rng.integers(0, 41)), int(rng.integers(0, 41))
    qlen = int(rng.integers(24, 65))
    wl_strs = tie_rich_whitelist(rng, int(rng.integers(200, 6000)))
    seqs = mixed_candidates(rng, wl_strs, 2500, pad_l, qlen, with_n=float(rng.choice([0.01, 0.1, 0.5])))
    # second and third N in some of them
    seqs = [s if rng.random() > 0.1 or len(s) < 3 else s[:int(len(s) // 2)] + "N" + s[int(len(s) // 2) + 1:] for s in seqs]
    wlc, _ = O.encode_many(wl_strs, 16)
    cc, cl = O.encode_many(seqs, 64)
    ref = O.match(wlc, pad_l, pad_r, cc, cl)
    wl = Whitelist(wl_strs, pad_l, pad_r)
    tot += compare(ref, run(wl, seqs, NR_MODE_FILTERED), 14, exact_below=False, label=f"round {it} filtered {pad_l}/{pad_r}/{qlen}")
    if it % 2 == 0:
        compare(ref, run(wl, seqs, NR_MODE_AUTO), 14, exact_below=True, label=f"round {it} auto")
    wl.close()
print(f"tie-rich rounds ok: {rounds} rounds, {tot} candidates at AS >= 14, {time.time() - t0:.0f} s", flush=True)
# slide-seq geometry: anchored filter + deep tier + brute force
from test_anchor_emul import slide_candidates, slide_whitelist  # noqa: E402
tot_s = 0
for it in range(max(1, rounds // 4)):
    rng = np.random.default_rng(20_000 + it)
    pad_l, pad_r = int(rng.integers(0, 31)), int(rng.integers(0, 31))
    wl_strs = slide_whitelist(rng, int(rng.integers(300, 4000)))
    seqs = slide_candidates(rng, O, wl_strs, 2000, with_n=float(rng.choice([0.0, 0.05, 0.5])))
    wlc, _ = O.encode_many(wl_strs, 32)
    cc, cl = O.encode_many(seqs, 64)
    ref = O.match(wlc, pad_l, pad_r, cc, cl)
    wl = Whitelist(wl_strs, pad_l, pad_r)
    buf, off = pack_ascii(seqs)
    b = wl.pack_device(torch.from_numpy(buf.copy()).to(dev), torch.from_numpy(off.view(np.int64).copy()).to(dev))
    for mode, eb in ((NR_MODE_FILTERED, False), (NR_MODE_AUTO, True)):
        r = wl.match_device(*b, min_score=30, mode=mode)
        torch.cuda.synchronize()
        res = MatchResult(*(t.cpu().numpy() for t in (r.idx, r.score, r.nbest, r.flags, r.umi_q)))
        tot_s += compare(ref, res, 30, exact_below=eb, label=f"slide-seq round {it} mode {mode} {pad_l}/{pad_r}")
    wl.close()
print(f"slide-seq rounds ok: {max(1, rounds // 4)} rounds, {tot_s} candidate checks at AS >= 30, {time.time() - t0:.0f} s", flush=True)
wl_a = whitelists.load_737k()
wl = Whitelist(wl_a, 30, 40)
wlc = O._CODE[wl_a]
for seed in range(6):
    geo = "5p"
    d = synth.make_candidates(wl_a, 1500, seed=500 + seed, p_sub=0.01 + 0.01 * seed, p_ins=0.03, p_del=0.01 * (seed % 3),
                              p_n=0.005 * seed, frac_negative=0.1 + 0.1 * seed)
    seqs = synth.to_strings(d["seqs"], d["offsets"])
    cc, cl = O.encode_many(seqs, 64)
    ref = O.match(wlc, 30, 40, cc, cl)
    compare(ref, run(wl, seqs, NR_MODE_FILTERED), 14, exact_below=False, label=f"737K seed {seed}")
    compare(ref, run(wl, seqs, NR_MODE_AUTO), 14, exact_below=True, label=f"737K seed {seed} auto")
print(f"737K rounds ok, {time.time() - t0:.0f} s")
