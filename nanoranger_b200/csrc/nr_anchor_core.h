// nr_anchor_core.h -- arithmetic of the anchored seed filter for cores with a constant middle
// (slide-seq: 8 barcode columns + the 18-column linker + 6 barcode columns, utils.py:584-601 of
// the reference; threshold AS >= 30, utils.py:638, i.e. cost <= 2 of 32).  Shared by the sm_100a
// kernel (nr_match_anchored.cu) and the host emulation the tests compile with g++
// (tests/emul/anchor_emul.cpp).  No CUDA runtime calls in here.
//
// Core = P (Lp <= 8 columns, differs between entries) + K (Lk columns, the same in every entry)
// + S (Ls columns).  An alignment of cost <= 2 crosses from P into K at some read row a and from K
// into S at some row b:  cost = cP(.., a) + cK(a, b) + cS(b, ..).  Hence
//   1. cK(a, b) <= 2: the linker lies in the read with at most one substitution / one deleted
//      column / two extra bases.  A furthest-reaching-diagonal walk from every start row a finds
//      all (a, b, cK) -- usually exactly one per read;
//   2. cP <= 2 - cK =: r.  EVERY 8-mer that can stand in front of row a at cost <= r is
//      enumerated from the read bases (NR_ANCHOR_SCRIPTS: exact, extra read bases inside P or at
//      the P|K junction, one substitution, one missing column, P hanging over the read start)
//      and looked up in a direct-address table of the entries' P parts (N columns of an entry
//      expanded to the four bases).  Nothing else can be a pair of cost <= 2, whatever S is;
//   3. every nominated (entry, strand) is scored exactly over all 32 columns by the plane
//      automaton of nr_deep_core.h (K = 2), which also yields the UMI column of the winner.
// Scripts are ordered by cost, so a read whose best pair costs t only runs the scripts with
// linker cost + script cost <= t (the stages of the 16-column filter, nr_filter_core.h).
//
// Reads with one or two N.  An N row scores 0 against any column (cost 1, no shift).  The linker
// walk counts N rows as matches (a lower bound on the linker cost: the budget for P only grows);
// the scripts run on the window with each N inside it substituted by the four bases (4 or 16
// windows): a placement of true cost c that aligns the N to a column of P costs c - 1 in the
// window that carries that column's base, so it is nominated at the same stage or earlier; the
// scorer knows the N rows (nr_deep_rows.nrow).
#pragma once
#include <stdint.h>

#include "nr_deep_core.h"

#define NR_ANCHOR_MAXP 8

// A script turns the 10 read bases in front of the junction row e (W = read[e-10 .. e), base k of
// the window at bits 2k) into one candidate 8-mer of P.
//   kind 0  the 8 bases before e, optionally with column `j` substituted by base ^ x   (cost 0 / 2)
//   kind 1  9 bases, interior base r1 removed                                           (cost 1)
//   kind 2  10 bases, interior bases r1 < r2 removed                                    (cost 2)
//   kind 3  7 bases, base x inserted as column j (a column missing from the read)       (cost 2)
//   kind 4  e == 8 - ov: P hangs over the read start by ov columns (unknown, bases x)   (cost ov)
//   kind 5  e == 8: one column over the read start (base x) and interior base r1 of the
//           8 read bases removed                                                         (cost 2)
struct nr_anchor_script {
    uint8_t kind, cost, j, x, r1, r2;
};

#define NR_ANCHOR_NSCRIPTS 192

struct nr_anchor_table {
    nr_anchor_script s[NR_ANCHOR_NSCRIPTS];
    int n;
    int first[4];       // scripts of cost c are s[first[c] .. first[c + 1])
};

// host-side (and constexpr-free) generation of the script table
NR_HD void nr_anchor_build_table(nr_anchor_table &t)
{
    int n = 0;
    t.first[0] = 0;
    t.s[n++] = {0, 0, 0, 0, 0, 0};                                            // exact
    t.first[1] = n;
    for (int r = 1; r <= 7; r++) t.s[n++] = {1, 1, 0, 0, (uint8_t)r, 0};      // one extra base inside P
    for (int x = 0; x < 4; x++) t.s[n++] = {4, 1, 1, (uint8_t)x, 0, 0};       // one column over the start
    t.first[2] = n;
    for (int j = 0; j < 8; j++)
        for (int x = 1; x < 4; x++) t.s[n++] = {0, 2, (uint8_t)j, (uint8_t)x, 0, 0};       // substitution
    for (int j = 0; j < 8; j++)
        for (int x = 0; x < 4; x++) t.s[n++] = {3, 2, (uint8_t)j, (uint8_t)x, 0, 0};       // missing column
    for (int r1 = 1; r1 <= 8; r1++)
        for (int r2 = r1 + 1; r2 <= 8; r2++) t.s[n++] = {2, 2, 0, 0, (uint8_t)r1, (uint8_t)r2};   // two extra bases
    for (int x = 0; x < 16; x++) t.s[n++] = {4, 2, 2, (uint8_t)x, 0, 0};      // two columns over the start
    for (int r = 1; r <= 7; r++)
        for (int x = 0; x < 4; x++) t.s[n++] = {5, 2, 0, (uint8_t)x, (uint8_t)r, 0};
    t.first[3] = n;
    t.n = n;
}

// remove base k (0-based) from a packed string of `len` bases
NR_HD uint32_t nr_anchor_remove(uint32_t v, int k)
{
    const uint32_t low = v & ((1u << (2 * k)) - 1u);
    return low | ((v >> (2 * k + 2)) << (2 * k));
}

// Apply a script.  W: the 10 bases before row e (bases before the read start are garbage: `e`
// tells how many are real).  Returns false when the script needs bases the read does not have
// or does not apply at this e.  key: 8 bases, column k of P at bits 2k.
NR_HD bool nr_anchor_apply(const nr_anchor_script &s, uint32_t W, int e, uint32_t *key)
{
    uint32_t k;
    switch (s.kind) {
    case 0:
        if (e < 8) return false;
        k = (W >> 4) & 0xFFFFu;
        if (s.cost) k ^= (uint32_t)s.x << (2 * s.j);
        break;
    case 1:
        if (e < 9) return false;
        k = nr_anchor_remove((W >> 2) & 0x3FFFFu, s.r1) & 0xFFFFu;
        break;
    case 2:
        if (e < 10) return false;
        k = nr_anchor_remove(nr_anchor_remove(W & 0xFFFFFu, s.r2), s.r1) & 0xFFFFu;
        break;
    case 3: {
        if (e < 7) return false;
        const uint32_t v = (W >> 6) & 0x3FFFu;                       // 7 bases
        const uint32_t low = v & ((1u << (2 * s.j)) - 1u);
        k = low | ((uint32_t)s.x << (2 * s.j)) | ((v >> (2 * s.j)) << (2 * s.j + 2));
        k &= 0xFFFFu;
        break;
    }
    case 4:
        if (e != 8 - (int)s.j) return false;                         // s.j = columns over the start
        k = ((W >> (2 * (2 + s.j))) << (2 * s.j)) & 0xFFFFu;         // the e real bases, shifted up
        k |= (uint32_t)s.x & ((1u << (2 * s.j)) - 1u);
        break;
    default:
        if (e != 8) return false;
        k = nr_anchor_remove((W >> 4) & 0xFFFFu, s.r1) & 0x3FFFu;    // 7 bases = columns 1..7
        k = (k << 2) | (uint32_t)s.x;
        break;
    }
    *key = k;
    return true;
}

// Linker placements starting at read row a: walk of the furthest-reaching diagonals (as
// nr_lv16, nr_filter_core.h) over Lk <= 24 columns.  V: read bases from row a - 1 on (base t of
// V = read[a - 1 + t]), vm: bit 2t set where that base exists.  link: the Lk linker columns.
// Returns reach flags: 1 (cost 0, b = a + Lk), 2 (cost 1, b = a + Lk + 1), 4 (cost 2, b = a + Lk
// + 2), 8 (cost 2 substitution, b = a + Lk), 16 (cost 2 missing column, b = a + Lk - 1).
// Bases outside the read count as matches: callers require a >= 0 and reject b > m.
NR_HD uint64_t nr_anchor_diag(uint64_t V, uint64_t vm, uint64_t link, int s, uint64_t lmask)
{
    const uint64_t w = V >> (2 * (s + 1));
    const uint64_t v = (vm >> (2 * (s + 1))) & 0x5555555555555555ull;
    const uint64_t x = link ^ w;
    return ((x | (x >> 1)) & 0x5555555555555555ull) & v & lmask;
}

NR_HD int nr_anchor_run(uint64_t nz, int j, int Lk)
{
    if (j >= Lk) return 0;
    const uint64_t t = nz >> (2 * j);
    if (!t) return Lk - j;
#if defined(__CUDA_ARCH__)
    return (__ffsll((long long)t) - 1) >> 1;
#else
    return __builtin_ctzll(t) >> 1;
#endif
}

NR_HD int nr_anchor_linker(uint64_t V, uint64_t vm, uint64_t link, int Lk)
{
    const uint64_t lmask = (Lk >= 32) ? ~0ull : ((1ull << (2 * Lk)) - 1ull);
    const uint64_t n0 = nr_anchor_diag(V, vm, link, 0, lmask);
    const int j0 = nr_anchor_run(n0, 0, Lk);
    if (j0 == Lk) return 1;
    int r = 0;
    const int j1 = j0 + nr_anchor_run(nr_anchor_diag(V, vm, link, 1, lmask), j0, Lk);
    if (j1 == Lk) r |= 2;
    else if (j1 + nr_anchor_run(nr_anchor_diag(V, vm, link, 2, lmask), j1, Lk) == Lk) r |= 4;
    if (j0 + 1 + nr_anchor_run(n0, j0 + 1, Lk) == Lk) r |= 8;
    if (j0 + 1 + nr_anchor_run(nr_anchor_diag(V, vm, link, -1, lmask), j0 + 1, Lk) == Lk) r |= 16;
    return r;
}

// cost and end-row shift of reach flag bit c
NR_HD int nr_anchor_flag_cost(int c) { return c == 0 ? 0 : (c == 1 ? 1 : 2); }
NR_HD int nr_anchor_flag_shift(int c) { return c == 0 ? 0 : (c == 1 ? 1 : (c == 2 ? 2 : (c == 3 ? 0 : -1))); }

// exact cost (<= 2, else 3) of one (read strand, entry) pair over all L columns; rows: the
// strand's row masks with the end-overhang rule on (edge = bit 0 | bit m)
NR_HD int nr_anchor_score(const nr_deep_rows &rows, uint32_t lo, uint32_t hi, uint32_t nm, int L, int m,
                          int padL, int padR)
{
    nr_deep_planes<2> x;
    nr_deep_init_fwd<2>(x, m, padL);
    for (int j = 0; j < L; j++) nr_deep_step_fwd<2>(x, rows, nr_core_col(lo, hi, j), (nm >> j) & 1u);
    int best = 3;
#pragma unroll
    for (int e = 0; e <= 2; e++) {
        // rows i with T[i][L] <= e and suffix excess max(0, m - i - padR) <= 2 - e
        for (int t = e; t <= 2; t++) {
            int from = m - padR - (t - e);
            if (from < 0) from = 0;
            if ((x.v[e] & ~((1ull << from) - 1ull)) && t < best) best = t;
        }
    }
    return best;
}

// substitute base `x` at window position k (0..9)
NR_HD uint32_t nr_anchor_subst(uint32_t W, int k, uint32_t x)
{
    return (W & ~(3u << (2 * k))) | (x << (2 * k));
}
