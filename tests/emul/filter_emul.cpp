// filter_emul.cpp -- host-side emulation of the filtered matcher (TEST INFRASTRUCTURE).
// Compiled by tests/ with g++ from the same nr_filter_core.h the sm_100a kernel includes, so
// that the probe set and the exact scorer can be checked against the oracle without a GPU.
// It is serial and slow on purpose; nothing in nanoranger_b200/ links or calls it.
#include <algorithm>
#include <cstdint>
#include <cstring>
#include <map>
#include <utility>
#include <vector>

#include "../../nanoranger_b200/csrc/nr_filter_core.h"

namespace {

// same layout as nr_whitelist.cu builds on the device: per dropped quarter j, rows {entry, core}
// sorted by key_j, a bitmap of present keys with, per 32-key word, the number of distinct keys
// below the word, and kstart[rank of a key] = its first row (+ sentinel)
struct Index {
    std::vector<std::pair<uint32_t, uint32_t>> rows[4];  // (key, entry) sorted
    std::vector<uint32_t> bits[4], rank[4], kstart[4];
};

void pack_read(const uint8_t *codes, int m, uint32_t w[4])
{
    w[0] = w[1] = w[2] = w[3] = 0;
    for (int i = 0; i < m; i++) w[i >> 4] |= (uint32_t)(codes[i] & 3) << ((i & 15) * 2);
}

void pad_read(const uint32_t w[4], uint32_t rdp[NR_RDP_WORDS])
{
    memset(rdp, 0, sizeof(uint32_t) * NR_RDP_WORDS);
    for (int k = 0; k < 4; k++) rdp[1 + k] = w[k];
}

void build_index(const uint32_t *wl, int64_t n, Index &ix)
{
    for (int j = 0; j < 4; j++) {
        ix.rows[j].resize((size_t)n);
        for (int64_t e = 0; e < n; e++) ix.rows[j][(size_t)e] = {nr_core_key(wl[e], j), (uint32_t)e};
        std::sort(ix.rows[j].begin(), ix.rows[j].end());
        ix.bits[j].assign((1u << 19) + 1, 0u);
        ix.rank[j].assign((1u << 19) + 1, 0u);
        for (auto &kv : ix.rows[j]) ix.bits[j][kv.first >> 5] |= 1u << (kv.first & 31u);
        std::vector<uint32_t> hs(ix.rows[j].size());
        uint32_t nd = 0;
        for (size_t i = 0; i < ix.rows[j].size(); i++) {
            if (i == 0 || ix.rows[j][i - 1].first != ix.rows[j][i].first) {
                ix.kstart[j].push_back((uint32_t)i);
                nd++;
            }
            hs[i] = nd;
        }
        ix.kstart[j].push_back((uint32_t)ix.rows[j].size());
        size_t r = 0;
        for (uint32_t w = 0; w <= (1u << 19); w++) {
            while (r < ix.rows[j].size() && (uint64_t)ix.rows[j][r].first < ((uint64_t)w << 5)) r++;
            ix.rank[j][w] = r == ix.rows[j].size() ? nd : hs[r] - 1;
        }
    }
}

int g_probe_limit = 0;     // > 0: only the first g_probe_limit main probes, no edge probes

}  // namespace

extern "C" {

// restrict the probe set to a prefix of NR_PROBES (0 = the whole table): used to check that the
// prefixes NR_PROBES_COST0 / NR_PROBES_COST1 are complete for costs 0 / <= 1
void nr_emul_set_probe_limit(int n) { g_probe_limit = n; }

// whole-read exact scorer on one pair: returns cost (0..2, 3 = more), *umi as nr_nfa16
int nr_emul_nfa(const uint8_t *q, int m, const uint8_t *core, int padL, int padR, int *umi)
{
    uint32_t w[4], rdp[NR_RDP_WORDS];
    pack_read(q, m, w);
    pad_read(w, rdp);
    uint32_t c = 0;
    for (int j = 0; j < 16; j++) c |= (uint32_t)(core[j] & 3) << (2 * j);
    return nr_nfa16(rdp, m, c, padL, padR, 0, m, umi);
}

// revcomp helper check
void nr_emul_revcomp(const uint8_t *q, int m, uint8_t *out)
{
    uint32_t w[4], r[4];
    pack_read(q, m, w);
    nr_revcomp4(w, m, r);
    for (int i = 0; i < 64; i++) out[i] = (uint8_t)((r[i >> 4] >> ((i & 15) * 2)) & 3u);
}

// the filtered matcher, serial.  wl: n packed cores.  cand: N x 64 codes (4 = N), clen: N.
// took[i] = 0 when the candidate is left to the exhaustive kernel (contains N, too short).
// score[i] = -128 / idx -1 / nbest 0 when no pair reaches cost <= 2.
// windowed != 0: score hits as the kernel does (nr_verify16: diagonal walk, automaton at the
// read ends); 0: whole-read automaton.
int nr_emul_filtered(const uint32_t *wl, int64_t n, int padL, int padR, const uint8_t *cand,
                     const uint8_t *clen, int64_t N, int windowed, int min_len, int32_t *idx,
                     int8_t *score, int32_t *nbest, uint8_t *strand, int16_t *umi, uint8_t *took,
                     int64_t *counters /* probes, hits, verifies */)
{
    Index ix;
    build_index(wl, n, ix);
    counters[0] = counters[1] = counters[2] = 0;
    for (int64_t c = 0; c < N; c++) {
        int m = clen[c];
        const uint8_t *q = cand + (size_t)c * 64;
        bool has_n = false;
        for (int i = 0; i < m && i < 64; i++) has_n |= q[i] > 3;
        idx[c] = -1; score[c] = -128; nbest[c] = 0; strand[c] = 0; umi[c] = -1;
        if (m > 64 || has_n || m < min_len) { took[c] = 0; continue; }
        took[c] = 1;
        uint32_t w[2][4], rdp[2][NR_RDP_WORDS];
        pack_read(q, m, w[0]);
        nr_revcomp4(w[0], m, w[1]);
        pad_read(w[0], rdp[0]);
        pad_read(w[1], rdp[1]);
        // (entry << 1 | strand) -> (cost, umi)
        std::map<uint32_t, std::pair<int, int>> found;
        int p0 = nr_slot_first(m, padR), p1 = nr_slot_last(m, padL);
        for (int s = 0; s < 2; s++)
            for (int p = p0; p <= p1; p++) {
                uint64_t W = nr_window64(rdp[s], p);
                int nt = (p == -1) ? NR_PROBES_ALL : NR_PROBES_MAIN;
                if (g_probe_limit > 0) nt = std::min(g_probe_limit, NR_PROBES_MAIN);
                for (int t = 0; t < nt; t++) {
                    const nr_probe_t &pr = NR_PROBES[t];
                    if (p + nr_probe_first(pr) < 0 || p + nr_probe_end(pr) > m) continue;
                    uint32_t key = nr_probe_key(W, pr);
                    counters[0]++;
                    auto &rows = ix.rows[pr.drop];
                    uint32_t bw = ix.bits[pr.drop][key >> 5];
                    if (!((bw >> (key & 31u)) & 1u)) continue;
                    size_t kr = ix.rank[pr.drop][key >> 5] +
                                (size_t)nr_popc32(bw & ((1u << (key & 31u)) - 1u));
                    for (size_t r = ix.kstart[pr.drop][kr]; r < ix.kstart[pr.drop][kr + 1]; r++) {
                        auto it = rows.begin() + (long)r;
                        if (it->first != key) return -2;   // index inconsistency
                        counters[1]++;
                        int r0 = windowed ? nr_rows_first(p) : 0;
                        int r1 = windowed ? nr_rows_last(p, m) : m;
                        int u;
                        int cost = windowed
                            ? nr_verify16(rdp[s], m, wl[it->second], padL, padR, p, pr, &u)
                            : nr_nfa16(rdp[s], m, wl[it->second], padL, padR, r0, r1, &u);
                        counters[2]++;
                        if (cost > 2) continue;
                        uint32_t k = (it->second << 1) | (uint32_t)s;
                        auto f = found.find(k);
                        if (f == found.end()) found[k] = {cost, u};
                        else if (cost < f->second.first) f->second = {cost, u};
                        else if (cost == f->second.first) {
                            int a = f->second.second, b = u;
                            // smallest row; -1 (ends inside the core) only if no window saw a row
                            f->second.second = (a < 0) ? b : (b < 0 ? a : std::min(a, b));
                        }
                    }
                }
            }
        int best = 3;
        for (auto &kv : found) best = std::min(best, kv.second.first);
        if (best > 2) continue;
        int cnt = 0; uint32_t bk = 0xFFFFFFFFu; int bu = -1;
        for (auto &kv : found)
            if (kv.second.first == best) {
                cnt++;
                if (kv.first < bk) { bk = kv.first; bu = kv.second.second; }
            }
        idx[c] = (int32_t)(bk >> 1); strand[c] = (uint8_t)(bk & 1u);
        score[c] = (int8_t)(16 - best); nbest[c] = cnt;
        umi[c] = (int16_t)((bk & 1u) ? -1 : bu);
    }
    return 0;
}

// whole-read exact scorer on one pair, read may contain N (code 4)
int nr_emul_nfa_n(const uint8_t *q, int m, const uint8_t *core, int padL, int padR, int *umi)
{
    uint8_t q2[64];
    uint64_t nm = 0;
    for (int i = 0; i < m; i++) { q2[i] = q[i] > 3 ? 0 : q[i]; if (q[i] > 3) nm |= 1ull << i; }
    uint32_t w[4], rdp[NR_RDP_WORDS];
    pack_read(q2, m, w);
    pad_read(w, rdp);
    uint32_t c = 0;
    for (int j = 0; j < 16; j++) c |= (uint32_t)(core[j] & 3) << (2 * j);
    return nr_nfa16n(rdp, nm, m, c, padL, padR, 0, m, umi);
}

// The filtered matcher for reads with one or two N, serial, with the kernel's schedule: rounds
// 0..2, variant v runs probe stage (round - nonzero substitutions of v) on the slots that reach
// its substituted positions, stop as soon as the best cost found is <= the finished round.
// took[i] = 0 for reads this mode does not take (no N, more than two, shorter than min_len).
int nr_emul_filtered_n(const uint32_t *wl, int64_t n, int padL, int padR, const uint8_t *cand,
                       const uint8_t *clen, int64_t N, int min_len, int32_t *idx, int8_t *score,
                       int32_t *nbest, uint8_t *strand, int16_t *umi, uint8_t *took,
                       int64_t *counters /* probes, verifies */)
{
    Index ix;
    build_index(wl, n, ix);
    counters[0] = counters[1] = 0;
    static const int stage_lo[3] = {0, NR_PROBES_COST0, NR_PROBES_COST1};
    static const int stage_hi[3] = {NR_PROBES_COST0, NR_PROBES_COST1, NR_PROBES_MAIN};
    for (int64_t c = 0; c < N; c++) {
        const int m = clen[c];
        const uint8_t *q = cand + (size_t)c * 64;
        idx[c] = -1; score[c] = -128; nbest[c] = 0; strand[c] = 0; umi[c] = -1; took[c] = 0;
        if (m > 64 || m < min_len) continue;
        uint64_t nm = 0;
        uint8_t q2[64];
        for (int i = 0; i < m; i++) { q2[i] = q[i] > 3 ? 0 : q[i]; if (q[i] > 3) nm |= 1ull << i; }
        const int n_n = __builtin_popcountll(nm);
        if (n_n < 1 || n_n > 2) continue;
        took[c] = 1;
        const int n0 = __builtin_ctzll(nm), n1 = n_n == 2 ? 63 - __builtin_clzll(nm) : -1;
        const uint64_t nms[2] = {nm, nr_rev_mask(nm, m)};
        uint32_t w0[4];
        pack_read(q2, m, w0);
        std::map<uint32_t, std::pair<int, int>> found;
        int best = 3;
        const int p0 = nr_slot_first(m, padR), p1 = nr_slot_last(m, padL);
        for (int round = 0; round < 3; round++) {
            if (best < round) break;
            for (int v = 0; v < nr_nvar_count(n_n); v++) {
                const int stage = round - nr_nvar_nonzero(v);
                if (stage < 0) continue;
                uint32_t w[2][4], rdp[2][NR_RDP_WORDS];
                nr_nvar_apply(w0, n0, n1, v, w[0]);
                nr_revcomp4(w[0], m, w[1]);
                pad_read(w[0], rdp[0]);
                pad_read(w[1], rdp[1]);
                for (int s = 0; s < 2; s++)
                    for (int p = p0; p <= p1; p++) {
                        const int n0s = s ? m - 1 - n0 : n0, n1s = n1 < 0 ? -100 : (s ? m - 1 - n1 : n1);
                        if (!nr_nvar_slot_needed(v, p, n0s, n1s)) continue;
                        const uint64_t W = nr_window64(rdp[s], p);
                        for (int pass = 0; pass < 2; pass++) {
                            int t0 = stage_lo[stage], t1 = stage_hi[stage];
                            if (pass == 1) {
                                if (!(stage == 2 && p == -1)) break;
                                t0 = NR_PROBES_MAIN; t1 = NR_PROBES_ALL;
                            }
                            for (int t = t0; t < t1; t++) {
                                const nr_probe_t &pr = NR_PROBES[t];
                                if (p + nr_probe_first(pr) < 0 || p + nr_probe_end(pr) > m) continue;
                                const uint32_t key = nr_probe_key(W, pr);
                                counters[0]++;
                                const uint32_t bw = ix.bits[pr.drop][key >> 5];
                                if (!((bw >> (key & 31u)) & 1u)) continue;
                                const size_t kr = ix.rank[pr.drop][key >> 5] +
                                                  (size_t)nr_popc32(bw & ((1u << (key & 31u)) - 1u));
                                for (size_t r = ix.kstart[pr.drop][kr]; r < ix.kstart[pr.drop][kr + 1]; r++) {
                                    const uint32_t e = ix.rows[pr.drop][r].second;
                                    int u;
                                    const int cost = nr_verify16n(rdp[s], nms[s], m, wl[e], padL, padR, p, pr, &u);
                                    counters[1]++;
                                    if (cost > 2) continue;
                                    if (cost < best) best = cost;
                                    const uint32_t k = (e << 1) | (uint32_t)s;
                                    auto f = found.find(k);
                                    if (f == found.end()) found[k] = {cost, u};
                                    else if (cost < f->second.first) f->second = {cost, u};
                                    else if (cost == f->second.first) {
                                        const int a = f->second.second, b = u;
                                        f->second.second = (a < 0) ? b : (b < 0 ? a : std::min(a, b));
                                    }
                                }
                            }
                        }
                    }
            }
        }
        if (best > 2) continue;
        int cnt = 0; uint32_t bk = 0xFFFFFFFFu; int bu = -1;
        for (auto &kv : found)
            if (kv.second.first == best) {
                cnt++;
                if (kv.first < bk) { bk = kv.first; bu = kv.second.second; }
            }
        idx[c] = (int32_t)(bk >> 1); strand[c] = (uint8_t)(bk & 1u);
        score[c] = (int8_t)(16 - best); nbest[c] = cnt;
        umi[c] = (int16_t)((bk & 1u) ? -1 : bu);
    }
    return 0;
}

}  // extern "C"
