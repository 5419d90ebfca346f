// anchor_emul.cpp -- host-side emulation of the anchored seed filter (TEST INFRASTRUCTURE).
// Compiled by tests/ with g++ from the headers the sm_100a kernel includes (nr_anchor_core.h:
// script table, key generation, linker walk, exact scorer; nr_anchor_index.h: P | K | S split and
// the P table), so the arithmetic is the shipped code; only the warp choreography differs.
#include <algorithm>
#include <cstdint>
#include <cstring>
#include <map>
#include <vector>

#include "../../nanoranger_b200/csrc/nr_anchor_core.h"
#include "../../nanoranger_b200/csrc/nr_anchor_index.h"

namespace {

// 32 read bases from position p on (2 bit/base; positions outside the read: 0), and their validity
void window(const uint8_t *q, int m, int p, uint64_t *V, uint64_t *vm)
{
    uint64_t v = 0, k = 0;
    for (int t = 0; t < 32; t++) {
        const int i = p + t;
        if (i >= 0 && i < m) { v |= (uint64_t)(q[i] & 3) << (2 * t); k |= 1ull << (2 * t); }
    }
    *V = v; *vm = k;
}

}  // namespace

extern "C" {

// cand: N x 64 codes (0..3, 4 = N; reads with more than two N are not taken), clen: N.  took[i] = 0: not taken.
// score -128 / idx -1 when no pair reaches cost <= 2.  info: Lp, Lk, Ls, rows of the P table.
int nr_emul_anchored(const uint32_t *wl_lo, const uint32_t *wl_hi, const uint32_t *wl_nm, int64_t n, int L,
                     int padL, int padR, const uint8_t *cand, const uint8_t *clen, int64_t N, int32_t *idx,
                     int8_t *score, int32_t *nbest, uint8_t *strand, int16_t *umi, uint8_t *took,
                     int32_t *info, int64_t *counters /* keys, nominees */)
{
    nr_anchor_index_host ix;
    nr_anchor_index_build(wl_lo, wl_hi, wl_nm, (uint64_t)n, L, ix);
    if (!ix.ok) return -1;
    if (info) { info[0] = ix.Lp; info[1] = ix.Lk; info[2] = ix.Ls; info[3] = (int32_t)ix.rows.size(); }
    nr_anchor_table tab;
    nr_anchor_build_table(tab);
    counters[0] = counters[1] = 0;
    const int Lp = ix.Lp, Lk = ix.Lk;
    for (int64_t c = 0; c < N; c++) {
        const int m = clen[c];
        idx[c] = -1; score[c] = -128; nbest[c] = 0; strand[c] = 0; umi[c] = -1; took[c] = 0;
        if (m < 1 || m > NR_DEEP_MAXM) continue;
        int n_n = 0;
        for (int i = 0; i < m; i++) n_n += cand[c * 64 + i] > 3;
        if (n_n > 2) continue;
        took[c] = 1;
        uint8_t q[2][64];
        uint64_t nmk[2] = {0, 0};
        nr_deep_rows rows[2];
        for (int st = 0; st < 2; st++) {
            for (int i = 0; i < m; i++) {
                const uint8_t x = cand[c * 64 + (st ? m - 1 - i : i)];
                q[st][i] = x > 3 ? 4 : (st ? (uint8_t)(3 - x) : x);
                if (x > 3) nmk[st] |= 1ull << i;
            }
            nr_deep_rows_from_codes(q[st], m, rows[st]);
        }
        // junctions: (strand, a, min linker cost)
        struct J { int st, a, ck; };
        std::vector<J> js;
        for (int st = 0; st < 2; st++)
            for (int a = Lp - 2; a + Lk - 1 <= m; a++) {
                if (a < 0) continue;
                uint64_t V, vm;
                window(q[st], m, a - 1, &V, &vm);
                // N rows count as matches in the walk
                for (int t = 0; t < 32; t++) {
                    const int i = a - 1 + t;
                    if (i >= 0 && i < m && ((nmk[st] >> i) & 1ull)) vm &= ~(1ull << (2 * t));
                }
                const int fl = nr_anchor_linker(V, vm, ix.link, Lk);
                if (!fl) continue;
                js.push_back({st, a, (fl & 1) ? 0 : ((fl & 2) ? 1 : 2)});
            }
        std::map<uint32_t, int> found;        // (entry << 1 | strand) -> cost
        int best = 3;
        for (int t = 0; t <= 2 && best >= t; t++)
            for (const J &j : js)
                for (int d = 0; j.ck + d <= t; d++) {
                    const int cs = t - j.ck - d, e = j.a - d;
                    uint32_t W0 = 0;
                    int npos[2] = {-1, -1}, nw = 0;               // N positions inside the window
                    for (int k = 0; k < 10; k++) {
                        const int i = e - 10 + k;
                        if (i >= 0 && i < m) {
                            if (q[j.st][i] > 3) npos[nw++] = k;
                            else W0 |= (uint32_t)(q[j.st][i] & 3) << (2 * k);
                        }
                    }
                    const int nvar = nw == 0 ? 1 : (nw == 1 ? 4 : 16);
                    for (int v = 0; v < nvar; v++)
                    for (int s = tab.first[cs]; s < tab.first[cs + 1]; s++) {
                        uint32_t W = W0;
                        if (nw >= 1) W = nr_anchor_subst(W, npos[0], (uint32_t)(v & 3));
                        if (nw >= 2) W = nr_anchor_subst(W, npos[1], (uint32_t)(v >> 2));
                        uint32_t key;
                        if (!nr_anchor_apply(tab.s[s], W, e, &key)) continue;
                        counters[0]++;
                        for (uint32_t r = ix.start[key]; r < ix.start[key + 1]; r++) {
                            const uint32_t en = ix.rows[r];
                            const uint32_t k2 = (en << 1) | (uint32_t)j.st;
                            if (found.count(k2)) continue;
                            counters[1]++;
                            const int cost = nr_anchor_score(rows[j.st], wl_lo[en], wl_hi ? wl_hi[en] : 0u,
                                                             wl_nm ? wl_nm[en] : 0u, L, m, padL, padR);
                            found[k2] = cost;
                            if (cost < best) best = cost;
                        }
                    }
                }
        if (best > 2) continue;
        int cnt = 0;
        uint32_t bk = 0xFFFFFFFFu;
        for (auto &kv : found)
            if (kv.second == best) { cnt++; if (kv.first < bk) bk = kv.first; }
        idx[c] = (int32_t)(bk >> 1); strand[c] = (uint8_t)(bk & 1u);
        score[c] = (int8_t)(L - best); nbest[c] = cnt;
        if (!(bk & 1u)) {
            nr_deep_rows r0 = rows[0];
            r0.edge = 1ull;
            const uint32_t en = bk >> 1;
            umi[c] = (int16_t)nr_deep_umi_row<2>(r0, wl_lo[en], wl_hi ? wl_hi[en] : 0u, wl_nm ? wl_nm[en] : 0u, L,
                                                 m, padL, padR, best);
        }
    }
    return 0;
}

}  // extern "C"
