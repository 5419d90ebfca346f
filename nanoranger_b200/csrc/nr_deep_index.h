// nr_deep_index.h -- host-side builder of the prefix/suffix grouping the deep tier works on
// (nr_deep_core.h).  Plain C++ (no CUDA): included by nr_whitelist.cu and by the CPU emulation
// in tests/emul/.
//
// Entries are ordered by their first s core columns; a "prefix group" is a run of entries
// sharing those columns (N columns are part of the identity), a "suffix group" a set of entries
// sharing the last L - s.  The split column s minimises  G_pre(s) * s + G_suf(s) * (L - s),
// the number of automaton column steps a candidate costs.
#pragma once
#include <algorithm>
#include <cstdint>
#include <vector>

struct nr_deep_index_host {
    int L = 0, s = 0;
    uint32_t g_pre = 0, g_suf = 0;
    std::vector<uint32_t> pre_start;    // g_pre + 1: first sorted position of every prefix group
    std::vector<uint32_t> pre_rep;      // g_pre x 4: lo, hi, nm of the group's columns, parent
    std::vector<uint32_t> suf_rep;      // g_suf x 4: lo, hi, nm, parent
    std::vector<uint32_t> ent_suf;      // n: suffix group of the entry at each sorted position
    std::vector<uint32_t> ent_idx;      // n: its index in the caller's whitelist
    // the same entries ordered by suffix group (the deep tier scans whichever side is rarer)
    std::vector<uint32_t> suf_start;    // g_suf + 1
    std::vector<uint32_t> sent_pre;     // n: prefix group of the entry at each suffix-sorted position
    std::vector<uint32_t> sent_idx;     // n: its index in the caller's whitelist
    // one level of sharing inside each half: "mid" groups = distinct first s1 columns (parents of
    // the prefix groups) and distinct last u1 columns (parents of the suffix groups); the automaton
    // runs the shared columns once per mid group.  s1 = 0 / u1 = 0: no sharing on that side.
    int s1 = 0, u1 = 0;
    uint32_t g_pmid = 0, g_smid = 0;
    std::vector<uint32_t> pmid_rep;     // g_pmid x 4: lo, hi, nm, 0
    std::vector<uint32_t> smid_rep;     // g_smid x 4
};

namespace nr_deep_detail {

// 3 bits per column (N << 2 | code), column `first` most significant; 96 bits in two words
struct Key {
    uint64_t a, b;      // a: first 21 columns, b: the other 11
    uint32_t idx;
};

inline Key make_key(uint32_t lo, uint32_t hi, uint32_t nm, int L, bool reversed, uint32_t idx)
{
    Key k{0, 0, idx};
    for (int t = 0; t < L; t++) {
        const int j = reversed ? L - 1 - t : t;
        const uint64_t code = ((j < 16) ? (lo >> (2 * j)) : (hi >> (2 * (j - 16)))) & 3u;
        const uint64_t n = (nm >> j) & 1u;
        const uint64_t sym = n ? 4u : code;
        if (t < 21) k.a |= sym << (3 * (20 - t));
        else k.b |= sym << (3 * (10 - (t - 21)));
    }
    return k;
}

inline bool key_less(const Key &x, const Key &y)
{
    if (x.a != y.a) return x.a < y.a;
    if (x.b != y.b) return x.b < y.b;
    return x.idx < y.idx;
}

// leading columns two keys share (0..32)
inline int key_lcp(const Key &x, const Key &y)
{
    if (x.a != y.a) {
        const int lz = __builtin_clzll(x.a ^ y.a) - 1;    // bit 62 is the top used bit
        return lz / 3;
    }
    if (x.b != y.b) {
        const int lz = __builtin_clzll(x.b ^ y.b) - 31;   // bit 32 is the top used bit
        return 21 + lz / 3;
    }
    return 32;
}

}  // namespace nr_deep_detail

// force_s > 0 fixes the split column (tests)
inline void nr_deep_index_build(const uint32_t *lo, const uint32_t *hi, const uint32_t *nm,
                                uint64_t n, int L, int force_s, nr_deep_index_host &out)
{
    using namespace nr_deep_detail;
    std::vector<Key> fw((size_t)n), bw((size_t)n);
    for (uint64_t e = 0; e < n; e++) {
        const uint32_t h = hi ? hi[e] : 0u, m = nm ? nm[e] : 0u;
        fw[(size_t)e] = make_key(lo[e], h, m, L, false, (uint32_t)e);
        bw[(size_t)e] = make_key(lo[e], h, m, L, true, (uint32_t)e);
    }
    std::sort(fw.begin(), fw.end(), key_less);
    std::sort(bw.begin(), bw.end(), key_less);
    // pairs of neighbours by shared leading columns -> distinct prefixes for every length
    std::vector<uint64_t> hf(34, 0), hb(34, 0);
    for (uint64_t e = 1; e < n; e++) {
        hf[(size_t)std::min(key_lcp(fw[(size_t)e - 1], fw[(size_t)e]), L)]++;
        hb[(size_t)std::min(key_lcp(bw[(size_t)e - 1], bw[(size_t)e]), L)]++;
    }
    auto groups = [&](const std::vector<uint64_t> &h, int len) {
        uint64_t g = 1;
        for (int c = 0; c < len; c++) g += h[(size_t)c];
        return g;
    };
    int s = force_s;
    if (s <= 0 || s >= L) {
        uint64_t best = ~0ull;
        s = L / 2 > 0 ? L / 2 : 1;
        for (int c = 1; c < L; c++) {
            const uint64_t cost = groups(hf, c) * (uint64_t)c + groups(hb, L - c) * (uint64_t)(L - c);
            const int dist = c > L / 2 ? c - L / 2 : L / 2 - c, bdist = s > L / 2 ? s - L / 2 : L / 2 - s;
            if (cost < best || (cost == best && dist < bdist)) { best = cost; s = c; }
        }
        if (L == 1) s = 1;     // degenerate: the prefix is the whole core, the suffix empty
    }
    out.L = L;
    out.s = s;
    // shared columns inside each half: s1 of the prefix's s, u1 of the suffix's L - s
    const uint64_t gpre = groups(hf, s), gsuf = groups(hb, L - s);
    int s1 = 0, u1 = 0;
    {
        uint64_t best = gpre * (uint64_t)s;
        for (int c = 1; c < s; c++) {
            const uint64_t cost = groups(hf, c) * (uint64_t)c + gpre * (uint64_t)(s - c);
            if (cost < best) { best = cost; s1 = c; }
        }
        best = gsuf * (uint64_t)(L - s);
        for (int c = 1; c < L - s; c++) {
            const uint64_t cost = groups(hb, c) * (uint64_t)c + gsuf * (uint64_t)(L - s - c);
            if (cost < best) { best = cost; u1 = c; }
        }
    }
    out.s1 = s1; out.u1 = u1;
    auto push_rep = [&](std::vector<uint32_t> &v, uint32_t i, uint32_t parent) {
        v.push_back(lo[i]); v.push_back(hi ? hi[i] : 0u); v.push_back(nm ? nm[i] : 0u);
        v.push_back(parent);
    };
    // suffix groups (runs of the suffix-sorted order) and their parents
    std::vector<uint32_t> suf_of((size_t)n);
    out.suf_rep.clear(); out.smid_rep.clear(); out.suf_start.clear();
    out.sent_idx.resize((size_t)n);
    uint32_t gs = 0, gsm = 0;
    for (uint64_t e = 0; e < n; e++) {
        const int lcp = e == 0 ? -1 : key_lcp(bw[(size_t)e - 1], bw[(size_t)e]);
        const uint32_t i = bw[(size_t)e].idx;
        if (u1 > 0 && lcp < u1) { push_rep(out.smid_rep, i, 0u); gsm++; }
        if (lcp < L - s) {
            push_rep(out.suf_rep, i, u1 > 0 ? gsm - 1 : 0u);
            out.suf_start.push_back((uint32_t)e);
            gs++;
        }
        suf_of[i] = gs - 1;
        out.sent_idx[(size_t)e] = i;
    }
    out.suf_start.push_back((uint32_t)n);
    out.g_suf = gs; out.g_smid = gsm;
    // prefix groups
    std::vector<uint32_t> pre_of((size_t)n);
    out.pre_start.clear(); out.pre_rep.clear(); out.pmid_rep.clear();
    out.ent_suf.resize((size_t)n); out.ent_idx.resize((size_t)n);
    uint32_t gp = 0, gpm = 0;
    for (uint64_t e = 0; e < n; e++) {
        const int lcp = e == 0 ? -1 : key_lcp(fw[(size_t)e - 1], fw[(size_t)e]);
        const uint32_t i = fw[(size_t)e].idx;
        if (s1 > 0 && lcp < s1) { push_rep(out.pmid_rep, i, 0u); gpm++; }
        if (lcp < s) {
            out.pre_start.push_back((uint32_t)e);
            push_rep(out.pre_rep, i, s1 > 0 ? gpm - 1 : 0u);
            gp++;
        }
        pre_of[i] = gp - 1;
        out.ent_suf[(size_t)e] = suf_of[i];
        out.ent_idx[(size_t)e] = i;
    }
    out.pre_start.push_back((uint32_t)n);
    out.g_pre = gp; out.g_pmid = gpm;
    out.sent_pre.resize((size_t)n);
    for (uint64_t e = 0; e < n; e++) out.sent_pre[(size_t)e] = pre_of[out.sent_idx[(size_t)e]];
}
