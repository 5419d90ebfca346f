// deep_emul.cpp -- host-side emulation of the deep matcher tier (TEST INFRASTRUCTURE).
// Compiled by tests/ with g++ from the headers the sm_100a kernel includes
// (nr_deep_core.h: plane automaton + join; nr_deep_index.h: prefix/suffix grouping), so the
// arithmetic is the shipped code; only the block choreography differs.  Serial and slow on
// purpose; nothing in nanoranger_b200/ links or calls it.
#include <cstdint>
#include <cstring>
#include <vector>

#include "../../nanoranger_b200/csrc/nr_deep_core.h"
#include "../../nanoranger_b200/csrc/nr_deep_index.h"

namespace {

template <int K>
void run(const nr_deep_index_host &ix, int padL, int padR, const uint8_t *cand,
         const uint8_t *clen, int64_t N, int32_t *idx, int8_t *score, int32_t *nbest,
         uint8_t *strand, uint8_t *took)
{
    const int L = ix.L, s = ix.s, s1 = ix.s1, u1 = ix.u1;
    // which side scans an entry: prefix groups with minimum <= A scan theirs, suffix groups with
    // minimum <= B scan the entries whose prefix minimum is > A; A + B >= K - 1 covers every
    // pair with prefix minimum + suffix minimum <= K
    const int A = K / 2, B_ = K - 1 - A;
    std::vector<nr_deep_planes<K>> F(ix.g_pre), B(ix.g_suf), FM(ix.g_pmid), BM(ix.g_smid);
    std::vector<int> fmin(ix.g_pre), bmin(ix.g_suf);
    for (int64_t c = 0; c < N; c++) {
        const int m = clen[c];
        took[c] = 0; idx[c] = -1; score[c] = -128; nbest[c] = 0; strand[c] = 0;
        if (m < 1 || m > NR_DEEP_MAXM) continue;
        int best = K + 1;
        int64_t cnt = 0;
        uint32_t key = 0xFFFFFFFFu;
        auto take = [&](uint32_t g, uint32_t h, uint32_t entry, int st) {
            const int bnd = best > K ? K : best;
            if (fmin[g] + bmin[h] > bnd) return;
            const int t = nr_deep_join<K>(F[g].v, B[h].v);
            if (t > bnd) return;
            const uint32_t k = (entry << 1) | (uint32_t)st;
            if (t < best) { best = t; cnt = 1; key = k; }
            else if (t == best) { cnt++; if (k < key) key = k; }
        };
        for (int st = 0; st < 2; st++) {
            uint8_t q[64];
            for (int i = 0; i < m; i++) {
                const uint8_t x = cand[c * 64 + (st ? m - 1 - i : i)];
                q[i] = st ? (x > 3 ? 4 : 3 - x) : x;
            }
            nr_deep_rows rows;
            nr_deep_rows_from_codes(q, m, rows);
            // phase A0: the shared columns, once per mid group
            for (uint32_t g = 0; g < ix.g_pmid; g++) {
                nr_deep_init_fwd<K>(FM[g], m, padL);
                const uint32_t lo = ix.pmid_rep[4 * g], hi = ix.pmid_rep[4 * g + 1], nm = ix.pmid_rep[4 * g + 2];
                for (int j = 0; j < s1; j++)
                    nr_deep_step_fwd<K>(FM[g], rows, nr_core_col(lo, hi, j), (nm >> j) & 1u);
            }
            for (uint32_t h = 0; h < ix.g_smid; h++) {
                nr_deep_init_bwd<K>(BM[h], m, padR, rows.valid);
                const uint32_t lo = ix.smid_rep[4 * h], hi = ix.smid_rep[4 * h + 1], nm = ix.smid_rep[4 * h + 2];
                for (int j = L - 1; j >= L - u1; j--)
                    nr_deep_step_bwd<K>(BM[h], rows, nr_core_col(lo, hi, j), (nm >> j) & 1u);
            }
            // phase A1: the rest of each half
            for (uint32_t g = 0; g < ix.g_pre; g++) {
                if (s1 > 0) F[g] = FM[ix.pre_rep[4 * g + 3]];
                else nr_deep_init_fwd<K>(F[g], m, padL);
                const uint32_t lo = ix.pre_rep[4 * g], hi = ix.pre_rep[4 * g + 1], nm = ix.pre_rep[4 * g + 2];
                for (int j = s1; j < s; j++)
                    nr_deep_step_fwd<K>(F[g], rows, nr_core_col(lo, hi, j), (nm >> j) & 1u);
                fmin[g] = nr_deep_min<K>(F[g]);
            }
            for (uint32_t h = 0; h < ix.g_suf; h++) {
                if (u1 > 0) B[h] = BM[ix.suf_rep[4 * h + 3]];
                else nr_deep_init_bwd<K>(B[h], m, padR, rows.valid);
                const uint32_t lo = ix.suf_rep[4 * h], hi = ix.suf_rep[4 * h + 1], nm = ix.suf_rep[4 * h + 2];
                for (int j = L - 1 - u1; j >= s; j--)
                    nr_deep_step_bwd<K>(B[h], rows, nr_core_col(lo, hi, j), (nm >> j) & 1u);
                bmin[h] = nr_deep_min<K>(B[h]);
            }
            // phase B, two-sided
            for (uint32_t g = 0; g < ix.g_pre; g++) {
                if (fmin[g] > A) continue;
                for (uint32_t p = ix.pre_start[g]; p < ix.pre_start[g + 1]; p++)
                    take(g, ix.ent_suf[p], ix.ent_idx[p], st);
            }
            for (uint32_t h = 0; h < ix.g_suf; h++) {
                if (bmin[h] > B_) continue;
                for (uint32_t p = ix.suf_start[h]; p < ix.suf_start[h + 1]; p++) {
                    const uint32_t g = ix.sent_pre[p];
                    if (fmin[g] <= A) continue;          // scanned from the prefix side
                    take(g, h, ix.sent_idx[p], st);
                }
            }
        }
        if (best <= K) {
            took[c] = 1;
            idx[c] = (int32_t)(key >> 1);
            strand[c] = (uint8_t)(key & 1u);
            score[c] = (int8_t)(L - best);
            nbest[c] = (int32_t)cnt;
        }
    }
}

}  // namespace

extern "C" {

// UMI column of one pair by the plane automaton (cost c of the pair known, c <= 8)
int nr_emul_deep_umi(const uint8_t *q, int m, uint32_t lo, uint32_t hi, uint32_t nm, int L, int padL,
                     int padR, int c)
{
    nr_deep_rows rows;
    nr_deep_rows_from_codes(q, m, rows);
    rows.edge = 1ull;
    return nr_deep_umi_row<8>(rows, lo, hi, nm, L, m, padL, padR, c);
}

// wl_lo / wl_hi / wl_nm: packed cores (hi, nm nullable).  cand: N x 64 codes (4 = N), clen: N.
// took[i] = 0 when no (entry, strand) pair reaches cost <= K (or the read has no / too many
// rows): the kernel hands those to the next tier.
int nr_emul_deep(const uint32_t *wl_lo, const uint32_t *wl_hi, const uint32_t *wl_nm, int64_t n,
                 int L, int padL, int padR, int K, int force_s, const uint8_t *cand,
                 const uint8_t *clen, int64_t N, int32_t *idx, int8_t *score, int32_t *nbest,
                 uint8_t *strand, uint8_t *took, int32_t *info /* s, g_pre, g_suf, s1, u1, g_pmid, g_smid */)
{
    if (L < 2 || L > 32) return -1;
    nr_deep_index_host ix;
    nr_deep_index_build(wl_lo, wl_hi, wl_nm, (uint64_t)n, L, force_s, ix);
    if (info) {
        info[0] = ix.s; info[1] = (int32_t)ix.g_pre; info[2] = (int32_t)ix.g_suf; info[3] = ix.s1;
        info[4] = ix.u1; info[5] = (int32_t)ix.g_pmid; info[6] = (int32_t)ix.g_smid;
    }
    switch (K) {
    case 2: run<2>(ix, padL, padR, cand, clen, N, idx, score, nbest, strand, took); break;
    case 3: run<3>(ix, padL, padR, cand, clen, N, idx, score, nbest, strand, took); break;
    case 5: run<5>(ix, padL, padR, cand, clen, N, idx, score, nbest, strand, took); break;
    case 8: run<8>(ix, padL, padR, cand, clen, N, idx, score, nbest, strand, took); break;
    default: return -2;
    }
    return 0;
}

}  // extern "C"
