// nr_deep_core.h -- arithmetic of the "deep" matcher tier: exact best score of a candidate over
// EVERY whitelist entry for costs up to K, without touching every (entry, read row) cell.
// Shared by the sm_100a kernel (nr_match_deep.cu) and the host emulation the tests compile with
// g++ (tests/emul/deep_emul.cpp).  No CUDA runtime calls in here.
//
// Why.  The reference's STAR call reports every uniquely mapped read whatever its score
// (scripts/barcode_align.sh:22-33: --outFilterScoreMinOverLread 0) and utils.py:698, 728-730
// histograms all of them into `_barcode_scores.csv`; the seed filter (nr_filter_core.h) is only
// complete for cost <= 2.  Reads below that need the exact optimum over the whole whitelist.
//
// Scoring in cost form, cost = L - AS, read rows i = 0..m consumed, core
// columns j = 0..L consumed:
//     T[i][0] = max(0, i - padL)                       read prefix beyond the left pad
//     T[0][j] = j                                      core starts before the read: 1 per column
//     T[i][j] = min(T[i-1][j-1] + w, T[i-1][j] + 1, T[i][j-1] + d(i))
//         w = 0 match, 2 mismatch, 1 when the read base or the core column is N
//         d(i) = 1 for i = 0 and i = m (column hangs over a read end), else 2 (deleted column)
//     cost = min_i T[i][L] + max(0, m - i - padR)      read suffix beyond the right pad
// (this is the recurrence of oracle/nr_oracle.c tier 1 with S = j - T; tests/test_deep_emul.py
// checks the two against each other).
//
// Meet in the middle.  Every alignment path crosses the line between columns s-1 and s at some
// row i, so with the mirror-image table U[i][j] = cheapest way to consume columns j..L-1 and the
// read suffix starting after row i,
//     cost(entry) = min_i T_prefix(entry)[i][s] + U_suffix(entry)[i][s].
// T[.][s] depends only on the entry's first s columns, U[.][s] only on its last L-s: they are
// computed once per DISTINCT prefix / suffix of the whitelist (737K-august-2016: 1 920 distinct
// 8-column prefixes, 1 536 suffixes) instead of once per entry.  Costs are kept as K+1 bit
// planes over the read rows (plane e = rows with cost <= e; a Wu-Manber style automaton with the
// read in the bit lanes, m <= 63), so joining a prefix with a suffix is
//     cost <= t  <=>  OR_{a+b=t} F_a & B_b != 0
// and an entry whose plane minima already exceed the running best is skipped on two byte loads.
#pragma once
#include <stdint.h>

#ifndef NR_HD
#if defined(__CUDACC__)
#define NR_HD __host__ __device__ __forceinline__
#else
#define NR_HD static inline
#endif
#endif

#define NR_DEEP_MAXM 63          // read rows 0..m must fit 64 bit lanes

// per-read (per-strand) row masks; bit i <-> read row i (row i consumes read base i-1)
struct nr_deep_rows {
    uint64_t eq[4];    // rows whose base equals code c (never set for N bases, never bit 0)
    uint64_t nrow;     // rows whose base is N
    uint64_t valid;    // bits 0..m
    uint64_t edge;     // bit 0 | bit m
};

template <int K>
struct nr_deep_planes {
    uint64_t v[K + 1];
};

// codes[0..m-1]: 0..3, 4 = N
NR_HD void nr_deep_rows_from_codes(const uint8_t *codes, int m, nr_deep_rows &r)
{
    r.eq[0] = r.eq[1] = r.eq[2] = r.eq[3] = 0;
    r.nrow = 0;
    for (int i = 0; i < m; i++) {
        const uint64_t bit = 1ull << (i + 1);
        if (codes[i] > 3) r.nrow |= bit;
        else r.eq[codes[i]] |= bit;
    }
    r.valid = (m >= 63) ? ~0ull : ((1ull << (m + 1)) - 1ull);
    r.edge = 1ull | (1ull << m);
}

template <int K>
NR_HD void nr_deep_init_fwd(nr_deep_planes<K> &f, int m, int padL)
{
#pragma unroll
    for (int e = 0; e <= K; e++) {
        int top = padL + e < m ? padL + e : m;              // rows 0..top
        f.v[e] = (top >= 63) ? ~0ull : ((1ull << (top + 1)) - 1ull);
    }
}

template <int K>
NR_HD void nr_deep_init_bwd(nr_deep_planes<K> &b, int m, int padR, uint64_t valid)
{
#pragma unroll
    for (int e = 0; e <= K; e++) {
        int lo = m - padR - e;                              // rows lo..m
        if (lo < 0) lo = 0;
        b.v[e] = valid & ~((1ull << lo) - 1ull);
    }
}

// consume one more core column (code c, col_n: the column is N) on the prefix side.
// NTERM = false drops the N terms: for callers that know neither the read nor the column has N.
template <int K, bool NTERM = true>
NR_HD void nr_deep_step_fwd(nr_deep_planes<K> &x, const nr_deep_rows &r, int c, bool col_n)
{
    const uint64_t eq = (NTERM && col_n) ? 0ull : r.eq[c];
    const uint64_t nr = !NTERM ? 0ull : (col_n ? (r.valid & ~1ull) : r.nrow);
    nr_deep_planes<K> y;
#pragma unroll
    for (int e = 0; e <= K; e++) {
        uint64_t v = (x.v[e] << 1) & eq;                               // match
        if (e >= 1) {
            if (NTERM) v |= (x.v[e - 1] << 1) & nr;                    // N on either side: 1
            v |= x.v[e - 1] & r.edge;                                  // column over a read end: 1
            v |= (y.v[e - 1] << 1);                                    // extra read base: 1
        }
        if (e >= 2) {
            v |= (x.v[e - 2] << 1);                                    // mismatch: 2
            v |= x.v[e - 2];                                           // deleted column: 2
        }
        y.v[e] = v & r.valid;
    }
#pragma unroll
    for (int e = 0; e <= K; e++) x.v[e] = y.v[e];
}

// the mirror image: prepend one core column on the suffix side
template <int K, bool NTERM = true>
NR_HD void nr_deep_step_bwd(nr_deep_planes<K> &x, const nr_deep_rows &r, int c, bool col_n)
{
    const uint64_t eq = ((NTERM && col_n) ? 0ull : r.eq[c]) >> 1;      // bit i: base of row i+1
    const uint64_t nr = !NTERM ? 0ull : ((col_n ? (r.valid & ~1ull) : r.nrow) >> 1);
    nr_deep_planes<K> y;
#pragma unroll
    for (int e = 0; e <= K; e++) {
        uint64_t v = (x.v[e] >> 1) & eq;
        if (e >= 1) {
            if (NTERM) v |= (x.v[e - 1] >> 1) & nr;
            v |= x.v[e - 1] & r.edge;
            v |= (y.v[e - 1] >> 1);
        }
        if (e >= 2) {
            v |= (x.v[e - 2] >> 1);
            v |= x.v[e - 2];
        }
        y.v[e] = v;                                                    // bits > m never appear
    }
#pragma unroll
    for (int e = 0; e <= K; e++) x.v[e] = y.v[e];
}

// smallest e with a non-empty plane, K + 1 if none
template <int K>
NR_HD int nr_deep_min(const nr_deep_planes<K> &x)
{
    int r = K + 1;
#pragma unroll
    for (int e = K; e >= 0; e--)
        if (x.v[e]) r = e;
    return r;
}

// cost of joining a prefix state with a suffix state: smallest t with OR_{a+b=t} F_a & B_b != 0
// (planes are cumulative, so that t is the exact cost when it is <= K), K + 1 if none.  Fully
// unrolled with compile-time plane indices so that f and b stay in registers on the device.
template <int K>
NR_HD int nr_deep_join(const uint64_t *f, const uint64_t *b)
{
    int res = K + 1;
#pragma unroll
    for (int t = K; t >= 0; t--) {
        uint64_t any = 0;
#pragma unroll
        for (int a = 0; a <= t; a++) any |= f[a] & b[t - a];
        if (any) res = t;
    }
    return res;
}

// column j of a packed core (lo: columns 0..15, hi: 16..31, nm: N columns)
NR_HD int nr_core_col(uint32_t lo, uint32_t hi, int j)
{
    return (int)(((j < 16) ? (lo >> (2 * j)) : (hi >> (2 * (j - 16)))) & 3u);
}

// UMI column of one (candidate, entry) pair whose cost c <= K is known: the smallest read row i at
// which an optimal alignment leaves the core, i.e. the smallest i with
//     T0[i][L] + max(0, m - i - padR) == c,
// where T0 is T without the "column hangs over the read END" rule (an alignment that ends inside
// the core never reaches reference column padL+L: utils.py:705-710 then finds no query index).
// Returns -1 when no row qualifies.  rows.edge must be 1 (bit 0 only) for this automaton.
template <int K>
NR_HD int nr_deep_umi_row(const nr_deep_rows &rows0, uint32_t lo, uint32_t hi, uint32_t nm, int L,
                          int m, int padL, int padR, int c)
{
    nr_deep_planes<K> x;
    nr_deep_init_fwd<K>(x, m, padL);
    for (int j = 0; j < L; j++) nr_deep_step_fwd<K>(x, rows0, nr_core_col(lo, hi, j), (nm >> j) & 1u);
    int best = 64;
#pragma unroll
    for (int e = 0; e <= K; e++) {
        if (e > c) continue;
        int from = m - padR - (c - e);                  // rows with suffix excess <= c - e
        if (from < 0) from = 0;
        const uint64_t s = x.v[e] & ~((1ull << from) - 1ull);
        if (s) {
#if defined(__CUDA_ARCH__)
            const int i = __ffsll((long long)s) - 1;
#else
            const int i = __builtin_ctzll(s);
#endif
            if (i < best) best = i;
        }
    }
    return best == 64 ? -1 : best;
}
