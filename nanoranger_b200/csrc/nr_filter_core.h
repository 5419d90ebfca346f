// nr_filter_core.h -- the arithmetic of the filtered matcher, written once for both the
// sm_100a kernel (nr_match_filtered.cu) and the host-side emulation that tests/ compiles with
// g++ to check the filter's losslessness against the oracle without a GPU
// (tests/emul/filter_emul.cpp).  No CUDA runtime calls in here.
//
// Problem (scripts/barcode_align.sh:18-33 of the reference): a candidate Q
// scores AS = 16 - cost against a 16-column N-free core, where cost is the cheapest way to
// place the core inside Q with
//     substitution 2, core column missing from Q (deletion) 2, extra Q base inside the core
//     (insertion) 1, core column hanging over either end of Q 1 per column,
//     Q prefix longer than padL / suffix longer than padR 1 per excess base.
// The reference keeps AS >= 14 (utils.py:699, 845, 1150, 1479), i.e. cost <= 2.
//
// Filter.  Cut the core into four 4-column quarters.  Any placement of cost <= 2 damages at
// most two quarters, and two only when both damages are cost-1 events (an insertion strictly
// inside a quarter, or a one-column overhang).  Every placement is therefore found by at
// least one of the probes in NR_PROBES: three quarters are read from Q at the stated offsets
// (one of them possibly as a 5-mer with one interior base removed) and looked up in the
// 24-bit "three of four quarters" index of the whitelist (nr_whitelist.cu).  A probe can
// only nominate entries; every nominated (entry, strand) is scored exactly by nr_nfa16().
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define NR_HD __host__ __device__ __forceinline__
#else
#define NR_HD static inline
#endif

// ---- padded packed reads -------------------------------------------------------------------
// rdp[0] = 0, rdp[1..4] = bases 0..63 (2 bit/base, base k at bit 2k), rdp[5..7] = 0.
#define NR_RDP_WORDS 8

NR_HD uint32_t nr_funnel_r(uint32_t lo, uint32_t hi, uint32_t s)  // 0 <= s < 32
{
#if defined(__CUDA_ARCH__)
    return __funnelshift_r(lo, hi, s);
#else
    return s ? ((lo >> s) | (hi << (32u - s))) : lo;
#endif
}

NR_HD int nr_popc32(uint32_t x)
{
#if defined(__CUDA_ARCH__)
    return __popc(x);
#else
    return __builtin_popcount(x);
#endif
}

// 32 bases (64 bit) starting at base position p, -16 <= p <= 64.
NR_HD uint64_t nr_window64(const uint32_t *rdp, int p)
{
    int q = p + 16;
    int w = q >> 4;
    uint32_t s = (uint32_t)(q & 15) * 2u;
    uint32_t a = rdp[w], b = rdp[w + 1], c = (w + 2 < NR_RDP_WORDS) ? rdp[w + 2] : 0u;
    uint32_t lo = nr_funnel_r(a, b, s);
    uint32_t hi = nr_funnel_r(b, c, s);
    return ((uint64_t)hi << 32) | lo;
}

NR_HD int nr_read_base(const uint32_t *rdp, int i)  // 0 <= i < 64
{
    return (int)((rdp[1 + (i >> 4)] >> ((i & 15) * 2)) & 3u);
}

// forward words (bases >= m must be zero) -> reverse complement words
NR_HD void nr_revcomp4(const uint32_t in[4], int m, uint32_t out[4])
{
    uint32_t r[6];
    for (int k = 0; k < 4; k++) {
        uint32_t x = ~in[3 - k];
#if defined(__CUDA_ARCH__)
        x = __brev(x);
#else
        x = ((x >> 1) & 0x55555555u) | ((x & 0x55555555u) << 1);
        x = ((x >> 2) & 0x33333333u) | ((x & 0x33333333u) << 2);
        x = ((x >> 4) & 0x0F0F0F0Fu) | ((x & 0x0F0F0F0Fu) << 4);
        x = ((x >> 8) & 0x00FF00FFu) | ((x & 0x00FF00FFu) << 8);
        x = (x >> 16) | (x << 16);
#endif
        // bit reversal also swapped the two bits of every base: swap them back
        r[k] = ((x & 0x55555555u) << 1) | ((x >> 1) & 0x55555555u);
    }
    r[4] = 0; r[5] = 0;
    int sh = 64 - m;  // leading garbage bases to drop
    int ws = sh >> 4;
    uint32_t bs = (uint32_t)(sh & 15) * 2u;
    for (int k = 0; k < 4; k++) {
        uint32_t a = (k + ws < 6) ? r[k + ws] : 0u;
        uint32_t b = (k + ws + 1 < 6) ? r[k + ws + 1] : 0u;
        uint32_t v = nr_funnel_r(a, b, bs);
        int lo = k * 16;
        if (m <= lo) v = 0;
        else if (m < lo + 16) v &= (1u << ((m - lo) * 2)) - 1u;
        out[k] = v;
    }
}

// ---- probes ----------------------------------------------------------------------------------
// A probe reads the three kept quarters at offsets o[0..2] (bases, relative to the slot
// position p) and drops quarter `drop`.  var = index (0..2) of the kept quarter that is read
// as a 5-mer with interior base `del` (1..3) removed, or -1.
struct nr_probe_t {
    int8_t drop, o0, o1, o2, var, del;
};

#define NR_PROBES_MAIN 34   // run at every slot position
#define NR_PROBES_EDGE 9    // run at p = -1 only (one-column start overhang + interior insertion)
#define NR_PROBES_ALL (NR_PROBES_MAIN + NR_PROBES_EDGE)

#define NR_V3(d, a, b, c, v) {d, a, b, c, v, 1}, {d, a, b, c, v, 2}, {d, a, b, c, v, 3}

// Prefixes of the table are complete for smaller costs (checked against the oracle in
// tests/test_filter_emul.py): the first NR_PROBES_COST0 probe finds every placement of cost 0
// (an undamaged core has its quarters 0..2 at p, p+4, p+8), the first NR_PROBES_COST1 every
// placement of cost <= 1 (one cost-1 event damages at most one quarter or one boundary); the
// variant probes and the edge probes only serve placements with two cost-1 events.
#define NR_PROBES_COST0 1
#define NR_PROBES_COST1 16

static constexpr nr_probe_t NR_PROBES[NR_PROBES_ALL] = {
    // one damaged quarter (or none) plus boundary insertions
    {3, 0, 4, 8, -1, 0},  {0, 4, 9, 13, -1, 0}, {0, 4, 8, 13, -1, 0},
    {0, 4, 8, 12, -1, 0}, {3, 0, 5, 9, -1, 0},  {3, 0, 4, 9, -1, 0},
    {1, 0, 7, 11, -1, 0}, {1, 0, 8, 12, -1, 0}, {1, 0, 9, 13, -1, 0}, {1, 0, 10, 14, -1, 0},
    {1, 0, 9, 14, -1, 0},
    {2, 0, 4, 11, -1, 0}, {2, 0, 4, 12, -1, 0}, {2, 0, 4, 13, -1, 0}, {2, 0, 4, 14, -1, 0},
    {2, 0, 5, 14, -1, 0},
    // interior insertions in two different quarters v < d: drop d, read v as 5-mer minus one
    NR_V3(1, 0, 10, 14, 0), NR_V3(2, 0, 5, 14, 0), NR_V3(3, 0, 5, 9, 0),
    NR_V3(2, 0, 4, 14, 1),  NR_V3(3, 0, 4, 9, 1),  NR_V3(3, 0, 4, 8, 2),
    // p = -1 only: quarter 0 loses its first column to the read start, insertion inside v >= 1
    NR_V3(0, 4, 9, 13, 0),  NR_V3(0, 4, 8, 13, 1), NR_V3(0, 4, 8, 12, 2),
};

// one kept quarter out of the 64-bit window W (base 0 of W = slot position p)
NR_HD uint32_t nr_quarter(uint64_t W, int off, int del)
{
    uint32_t v = (uint32_t)(W >> (2 * off));
    if (del == 0) return v & 0xFFu;
    uint32_t low = v & ((1u << (2 * del)) - 1u);
    uint32_t high = (v >> (2 * del + 2)) << (2 * del);
    return (low | high) & 0xFFu;
}

// extent of the probe in Q: first base used and one past the last base used (relative to p)
NR_HD int nr_probe_first(const nr_probe_t &t) { return t.o0; }
NR_HD int nr_probe_end(const nr_probe_t &t) { return t.o2 + (t.var == 2 ? 5 : 4); }

NR_HD uint32_t nr_probe_key(uint64_t W, const nr_probe_t &t)
{
    uint32_t a = nr_quarter(W, t.o0, t.var == 0 ? t.del : 0);
    uint32_t b = nr_quarter(W, t.o1, t.var == 1 ? t.del : 0);
    uint32_t c = nr_quarter(W, t.o2, t.var == 2 ? t.del : 0);
    return a | (b << 8) | (c << 16);
}

// ---- bitmap layouts ------------------------------------------------------------------------------
// A 32-byte sector of a key bitmap holds the 256 keys that differ in the key's LOW byte.  The
// device keeps every bitmap in three bit orders so that the probes of a slot that differ in one
// kept quarter only fall into one sector whichever quarter that is (L1TEX/L2 work is per sector):
//   layout 0: a | b << 8 | c << 16   (the index order: rank / kstart / ents use it)
//   layout 1: b | a << 8 | c << 16
//   layout 2: c | a << 8 | b << 16
// Table id = layout * 4 + dropped quarter; bit address = id << 24 | permuted key.
#define NR_BM_LAYOUTS 3

NR_HD uint32_t nr_key_layout(uint32_t key, int layout)
{
    const uint32_t a = key & 0xFFu, b = (key >> 8) & 0xFFu, c = key >> 16;
    if (layout == 1) return b | (a << 8) | (c << 16);
    if (layout == 2) return c | (a << 8) | (b << 16);
    return key;
}

// layout a main probe reads: the one in which its family (same offsets, one quarter varying
// between the members) shares a sector
#if defined(__CUDACC__)
__host__ __device__
#endif
constexpr int nr_probe_layout_of(const nr_probe_t &t)
{
    if (t.var == 1) return 1;                                  // b varies between the 3 variants
    if (t.var == 2) return 2;                                  // c varies
    if (t.var == 0) return 0;                                  // a varies
    // plain probes: c-varying groups (same drop, o0, o1)
    if (t.drop == 0 && t.o0 == 4 && t.o1 == 8) return 2;       // (4,8,12) (4,8,13)
    if (t.drop == 3 && t.o0 == 0 && t.o1 == 4) return 2;       // (0,4,8) (0,4,9) + the c variants
    if (t.drop == 2 && t.o0 == 0 && t.o1 == 4) return 2;       // (0,4,11..14)
    if (t.drop == 1 && t.o0 == 0 && t.o1 == 9) return 2;       // (0,9,13) (0,9,14)
    return 0;                                                  // heads of the a-varying families, singles
}

// key of a whitelist core with quarter j removed (must agree with nr_probe_key)
NR_HD uint32_t nr_core_key(uint32_t core, int j)
{
    switch (j) {
    case 0: return core >> 8;
    case 1: return (core & 0xFFu) | ((core >> 8) & 0xFFFF00u);
    case 2: return (core & 0xFFFFu) | ((core >> 8) & 0xFF0000u);
    default: return core & 0xFFFFFFu;
    }
}

// slot positions worth probing for a read of length m: first..last inclusive
NR_HD int nr_slot_first(int m, int padR) { int a = m - padR - 22; return a > -2 ? a : -2; }
NR_HD int nr_slot_last(int m, int padL) { int a = m - 10, b = padL + 3; return a < b ? a : b; }

// rows of Q the exact scorer must cover for a hit at slot position p
NR_HD int nr_rows_first(int p) { return p - 3 > 0 ? p - 3 : 0; }
NR_HD int nr_rows_last(int p, int m) { return p + 24 < m ? p + 24 : m; }

// ---- exact scorer for cost <= 2 -----------------------------------------------------------------
// Three-level shift-and automaton over the core columns, state j (1..16 columns consumed) at
// bit 2(j-1); level k holds the states reachable with cost <= k.  State 0 (still in the left
// pad) is implicit with cost z(i) = max(0, i - padL).  Rows r0..r1 of Q are consumed (r0 = 0
// applies the start-overhang rule, r1 = m the end-overhang rule).  Returns the cost (0..2) of
// the best placement inside those rows, 3 if none; *umi = smallest row at which a best
// placement leaves the core (the query index aligned to reference column padL+16,
// utils.py:705-708), -1 if the best placement ends inside the core.
// NR_NFA16_BODY(BASE_AT, ISN_AT) is the body; BASE_AT(i) yields base i of Q, ISN_AT(i) whether it
// is N: an N row scores 0 against every column (cost 1 on the diagonal instead of 0 / 2).
#define NR_NFA16_BODY(BASE_AT, ISN_AT)                                                        \
    const uint32_t KEEP = 0x15555555u; /* states 1..15 */                                     \
    const uint32_t FIN = 0x40000000u;  /* state 16 */                                         \
    uint32_t R0 = 0, R1, R2;                                                                  \
    if (r0 == 0) { R1 = 1u; R2 = 5u; }                                                        \
    else { R1 = 0u; R2 = (r0 <= padL) ? 1u : 0u; }                                            \
    int best = 3, arg = -1;                                                                   \
    for (int i = r0; i < r1; i++) {                                                           \
        uint32_t c = (uint32_t)(BASE_AT(i));                                                  \
        uint32_t x = core ^ (c * 0x55555555u);                                                \
        const uint32_t nrow = (ISN_AT(i)) ? 0xFFFFFFFFu : 0u;                                 \
        uint32_t M = ~(x | (x >> 1)) & 0x55555555u & ~nrow;                                   \
        int z = i - padL; /* cost of state 0 at row i when positive */                        \
        uint32_t S0 = (R0 << 2) | (z <= 0 ? 1u : 0u);                                         \
        uint32_t S1 = (R1 << 2) | (z <= 1 ? 1u : 0u);                                         \
        uint32_t S2 = (R2 << 2) | (z <= 2 ? 1u : 0u);                                         \
        uint32_t A0 = S0 & M;                                                                 \
        uint32_t A1 = (S1 & M) | (R0 & KEEP) | A0 | (S0 & nrow);                              \
        uint32_t A2 = (S2 & M) | (R1 & KEEP) | S0 | A1 | (S1 & nrow);                         \
        A2 |= (A0 << 2) | (z + 1 <= 0 ? 1u : 0u); /* one deleted column, cost 2 */            \
        R0 = A0; R1 = A1; R2 = A2;                                                            \
        if (R2 & FIN) {                                                                       \
            int k = (R0 & FIN) ? 0 : ((R1 & FIN) ? 1 : 2);                                    \
            int t = m - (i + 1) - padR;                                                       \
            int tot = k + (t > 0 ? t : 0);                                                    \
            if (tot < best) { best = tot; arg = i + 1; }                                      \
        }                                                                                     \
    }                                                                                         \
    if (r1 == m) {                                                                            \
        int tot = 3;                                                                          \
        if (R0 & 0x10000000u) tot = 1; /* state 15 at cost 0 */                               \
        else if ((R1 & 0x10000000u) | (R0 & 0x04000000u)) tot = 2; /* 15 @1 or 14 @0 */       \
        if (tot < best) { best = tot; arg = -1; }                                             \
    }                                                                                         \
    *umi = arg;                                                                               \
    return best;

// any rows, bases fetched from the padded read
NR_HD int nr_nfa16(const uint32_t *rdp, int m, uint32_t core, int padL, int padR, int r0, int r1,
                   int *umi)
{
#define NR_BASE_AT(i) nr_read_base(rdp, (i))
#define NR_ISN_AT(i) 0
    NR_NFA16_BODY(NR_BASE_AT, NR_ISN_AT)
#undef NR_BASE_AT
#undef NR_ISN_AT
}

// at most 32 rows, bases r0.. held in Wn = nr_window64(rdp, r0)
NR_HD int nr_nfa16_w(uint64_t Wn, int m, uint32_t core, int padL, int padR, int r0, int r1,
                     int *umi)
{
#define NR_BASE_AT(i) ((uint32_t)(Wn >> (2 * ((i) - r0))) & 3u)
#define NR_ISN_AT(i) 0
    NR_NFA16_BODY(NR_BASE_AT, NR_ISN_AT)
#undef NR_BASE_AT
#undef NR_ISN_AT
}

// reads with N: nm has bit i set where base i of Q is N (its 2-bit code is then ignored)
NR_HD int nr_nfa16n(const uint32_t *rdp, uint64_t nm, int m, uint32_t core, int padL, int padR,
                    int r0, int r1, int *umi)
{
#define NR_BASE_AT(i) nr_read_base(rdp, (i))
#define NR_ISN_AT(i) ((nm >> (i)) & 1ull)
    NR_NFA16_BODY(NR_BASE_AT, NR_ISN_AT)
#undef NR_BASE_AT
#undef NR_ISN_AT
}

// at most 32 rows; nmw = nm >> r0
NR_HD int nr_nfa16n_w(uint64_t Wn, uint32_t nmw, int m, uint32_t core, int padL, int padR, int r0,
                      int r1, int *umi)
{
#define NR_BASE_AT(i) ((uint32_t)(Wn >> (2 * ((i) - r0))) & 3u)
#define NR_ISN_AT(i) ((nmw >> ((i) - r0)) & 1u)
    NR_NFA16_BODY(NR_BASE_AT, NR_ISN_AT)
#undef NR_BASE_AT
#undef NR_ISN_AT
}

// ---- reads with one or two N ----------------------------------------------------------------
// A read N scores 0 against any core column: cost 1, no shift -- an event the probe table does
// not enumerate (two damaged quarters are only covered for insertions / overhangs).  Instead the
// N is SUBSTITUTED: variant v of the read carries base (v & 3) at the first N and ((v >> 2) & 3)
// at the second.  A placement of true cost c that aligns a set A of the N's to core columns has
// cost c - |A| in the variant that carries those columns' bases (and base 0 at the other N's),
// and the key of the probe that finds a placement only depends on read bases that match the
// core, so:
//   - a variant with z non-zero substitutions is only needed for placements with |A| >= z, and
//     there only the probe stages complete for cost <= 2 - z;
//   - in round r (after which every placement of true cost <= r must be known) variant v runs
//     stage r - z(v);
//   - slots whose probes cannot reach a substituted N read the same keys as the variant with
//     that substitution zeroed, which runs the same or a later stage: they are skipped.
// Nominated (entry, strand) pairs are scored against the ORIGINAL read by nr_nfa16n_w.
NR_HD int nr_nvar_count(int n_n) { return n_n == 1 ? 4 : 16; }
NR_HD int nr_nvar_nonzero(int v) { return ((v & 3) != 0) + ((v >> 2) != 0); }
// read positions a probe at slot p may use: p .. p + 18; one base of margin on both sides
NR_HD bool nr_slot_reaches(int p, int pos) { return pos >= p - 1 && pos <= p + 20; }
NR_HD bool nr_nvar_slot_needed(int v, int p, int n0s, int n1s)   // N positions on the slot's strand
{
    if ((v & 3) != 0 && !nr_slot_reaches(p, n0s)) return false;
    if ((v >> 2) != 0 && !nr_slot_reaches(p, n1s)) return false;
    return true;
}
// the four packed words of variant v (n1 < 0: one N)
NR_HD void nr_nvar_apply(const uint32_t in[4], int n0, int n1, int v, uint32_t out[4])
{
    for (int k = 0; k < 4; k++) out[k] = in[k];
    out[n0 >> 4] = (out[n0 >> 4] & ~(3u << ((n0 & 15) * 2))) | ((uint32_t)(v & 3) << ((n0 & 15) * 2));
    if (n1 >= 0)
        out[n1 >> 4] = (out[n1 >> 4] & ~(3u << ((n1 & 15) * 2))) |
                       ((uint32_t)((v >> 2) & 3) << ((n1 & 15) * 2));
}
// N mask of the reverse-complement strand
NR_HD uint64_t nr_rev_mask(uint64_t nm, int m)
{
    uint64_t r = 0;
    for (int k = 0; k < m; k++)
        if ((nm >> k) & 1ull) r |= 1ull << (m - 1 - k);
    return r;
}

// ---- interior scorer: furthest-reaching diagonals ------------------------------------------------
// For a hit whose probe pins one end of the core (quarter 0 kept: the core starts at read
// position a; quarter 0 dropped: the core ends just before read position e) the placements that
// agree with the probe all start (end) there, and with cost <= 2 they are: no edit, one
// insertion, two insertions, one substitution, one deletion.  Landau-Vishkin style: follow
// diagonal 0 while it matches, then branch.  V holds the read as seen from the pinned end:
// base t of V is sequence index t-1 relative to the pinned column (forward: read[a-1+t];
// backward: read[e-t], used with the base-reversed core), vm has bit 2t set where base t exists.
// Positions outside the read count as matches, so the result is exact when all of
// V[0..18] exists and a lower bound on the cost otherwise (then the automaton decides).
// Returns reach flags: 1 (cost 0, diagonal 0), 2 (cost 1, diagonal +1), 4 (cost 2, diagonal +2),
// 8 (cost 2, diagonal 0), 16 (cost 2, diagonal -1).
NR_HD int nr_ctz32(uint32_t x)
{
#if defined(__CUDA_ARCH__)
    return __ffs((int)x) - 1;
#else
    return __builtin_ctz(x);
#endif
}

NR_HD uint32_t nr_lv_diag(uint64_t V, uint64_t vm, uint32_t core, int s)   // s in -1..2
{
    uint32_t w = (uint32_t)(V >> (2 * (s + 1)));
    uint32_t v = (uint32_t)(vm >> (2 * (s + 1))) & 0x55555555u;
    uint32_t x = core ^ w;
    return ((x | (x >> 1)) & 0x55555555u) & v;      // bit 2j: column j mismatches on diagonal s
}

NR_HD int nr_lv_run(uint32_t nz, int j)   // matches from column j on (0 <= j <= 16)
{
    if (j >= 16) return 0;
    uint32_t t = nz >> (2 * j);
    return t ? (nr_ctz32(t) >> 1) : 16 - j;
}

NR_HD int nr_lv16(uint64_t V, uint64_t vm, uint32_t core)
{
    const uint32_t n0 = nr_lv_diag(V, vm, core, 0);
    const int j0 = nr_lv_run(n0, 0);
    if (j0 == 16) return 1;
    int r = 0;
    const int j1 = j0 + nr_lv_run(nr_lv_diag(V, vm, core, 1), j0);
    if (j1 == 16) r |= 2;
    else if (j1 + nr_lv_run(nr_lv_diag(V, vm, core, 2), j1) == 16) r |= 4;
    if (j0 + 1 + nr_lv_run(n0, j0 + 1) == 16) r |= 8;
    if (j0 + 1 + nr_lv_run(nr_lv_diag(V, vm, core, -1), j0 + 1) == 16) r |= 16;
    return r;
}

// bit 2t set for t in [lo, hi) (0 <= lo, hi <= 32 after clamping)
NR_HD uint64_t nr_valid_mask(int lo, int hi)
{
    if (lo < 0) lo = 0;
    if (hi > 32) hi = 32;
    if (hi <= lo) return 0;
    uint64_t a = hi >= 32 ? ~0ull : ((1ull << (2 * hi)) - 1ull);
    uint64_t b = (1ull << (2 * lo)) - 1ull;
    return (a & ~b) & 0x5555555555555555ull;
}

NR_HD uint64_t nr_rev_bases64(uint64_t x)
{
    uint32_t lo = (uint32_t)x, hi = (uint32_t)(x >> 32);
#if defined(__CUDA_ARCH__)
    lo = __brev(lo); hi = __brev(hi);
#else
    for (int k = 0; k < 2; k++) {
        uint32_t v = k ? hi : lo;
        v = ((v >> 1) & 0x55555555u) | ((v & 0x55555555u) << 1);
        v = ((v >> 2) & 0x33333333u) | ((v & 0x33333333u) << 2);
        v = ((v >> 4) & 0x0F0F0F0Fu) | ((v & 0x0F0F0F0Fu) << 4);
        v = ((v >> 8) & 0x00FF00FFu) | ((v & 0x00FF00FFu) << 8);
        v = (v >> 16) | (v << 16);
        if (k) hi = v; else lo = v;
    }
#endif
    lo = ((lo & 0x55555555u) << 1) | ((lo >> 1) & 0x55555555u);
    hi = ((hi & 0x55555555u) << 1) | ((hi >> 1) & 0x55555555u);
    return ((uint64_t)lo << 32) | hi;        // halves swapped: full 32-base reversal
}

NR_HD uint32_t nr_rev_bases32(uint32_t x) { return (uint32_t)(nr_rev_bases64((uint64_t)x) >> 32); }

// Verification of one nominated (entry, strand) for the probe `t` at slot position p.
// Returns the cost (0..2, 3 = none) and *umi as nr_nfa16 does.  Exact for the placements that
// agree with the probe; where the probe's pinned end is too close to a read end for the
// diagonal walk to be exact, the walk only filters and the automaton scores.
NR_HD int nr_verify16(const uint32_t *rdp, int m, uint32_t core, int padL, int padR, int p,
                      const nr_probe_t &t, int *umi)
{
    const int fwd = t.drop != 0;
    // forward: core column 0 at read position a = p; backward: core column 15 at read
    // position e - 1, base t of V = read[e - t], scored against the base-reversed core
    const int pinned = fwd ? p : p + nr_probe_end(t);
    uint64_t V = nr_window64(rdp, fwd ? pinned - 1 : pinned - 19);
    if (!fwd) V = nr_rev_bases64(V) >> 24;
    const int interior = fwd ? (pinned >= 0 && pinned + 18 <= m) : (pinned >= 19 && pinned <= m);
    // every base of V[0..18] exists when the pinned end is at least one base inside the read on
    // both sides: the mask is then all ones (the usual case; saves building it)
    const int deep = fwd ? (pinned >= 1 && pinned + 18 <= m) : (pinned >= 19 && pinned < m);
    const uint64_t vm = deep ? 0x5555555555555555ull
                             : (fwd ? nr_valid_mask(1 - pinned, m - pinned + 1)
                                    : nr_valid_mask(pinned - m + 1, pinned + 1));
    const int flags = nr_lv16(V, vm, fwd ? core : nr_rev_bases32(core));
    if (!flags) { *umi = -1; return 3; }
    if (!interior)
        return nr_nfa16_w(nr_window64(rdp, nr_rows_first(p)), m, core, padL, padR,
                          nr_rows_first(p), nr_rows_last(p, m), umi);
    // totals: edit cost + read prefix beyond padL + read suffix beyond padR
    int best = 3, arg = -1;
    const int ks[5] = {0, 1, 2, 2, 2}, ss[5] = {0, 1, 2, 0, -1};
    for (int c = 0; c < 5; c++) {
        if (!(flags >> c & 1)) continue;
        int start = fwd ? pinned : pinned - 16 - ss[c];
        int end = fwd ? pinned + 16 + ss[c] : pinned;
        int pre = start - padL, suf = m - end - padR;
        int tot = ks[c] + (pre > 0 ? pre : 0) + (suf > 0 ? suf : 0);
        if (tot < best || (tot == best && end < arg)) { best = tot; arg = end; }
    }
    if (best > 2) { *umi = -1; return 3; }
    *umi = arg;
    return best;
}

// bit t of x -> bit 2t (t < 32)
NR_HD uint64_t nr_spread_even(uint32_t x)
{
    uint64_t v = x;
    v = (v | (v << 16)) & 0x0000FFFF0000FFFFull;
    v = (v | (v << 8)) & 0x00FF00FF00FF00FFull;
    v = (v | (v << 4)) & 0x0F0F0F0F0F0F0F0Full;
    v = (v | (v << 2)) & 0x3333333333333333ull;
    v = (v | (v << 1)) & 0x5555555555555555ull;
    return v;
}

NR_HD uint64_t nr_rev_bits64(uint64_t x)
{
#if defined(__CUDA_ARCH__)
    return __brevll(x);
#else
    x = ((x >> 1) & 0x5555555555555555ull) | ((x & 0x5555555555555555ull) << 1);
    x = ((x >> 2) & 0x3333333333333333ull) | ((x & 0x3333333333333333ull) << 2);
    x = ((x >> 4) & 0x0F0F0F0F0F0F0F0Full) | ((x & 0x0F0F0F0F0F0F0F0Full) << 4);
    x = ((x >> 8) & 0x00FF00FF00FF00FFull) | ((x & 0x00FF00FF00FF00FFull) << 8);
    x = ((x >> 16) & 0x0000FFFF0000FFFFull) | ((x & 0x0000FFFF0000FFFFull) << 16);
    return (x >> 32) | (x << 32);
#endif
}

// Verification of one nominated (entry, strand) for a read with N (nm: N mask of the strand,
// rdp: any substituted variant), in two halves so that the kernel can run the dear half on full
// warps.  nr_prefilter16n: the diagonal walk of nr_verify16 with the N rows counted as matches --
// a lower bound on the true cost, so "nothing within cost 2" is exact.  nr_score16n: the N-aware
// automaton over the rows around the slot, for whatever survives.
NR_HD bool nr_prefilter16n(const uint32_t *rdp, uint64_t nm, int m, uint32_t core, int p,
                           const nr_probe_t &t)
{
    const int fwd = t.drop != 0;
    const int pinned = fwd ? p : p + nr_probe_end(t);
    uint64_t V = nr_window64(rdp, fwd ? pinned - 1 : pinned - 19);
    if (!fwd) V = nr_rev_bases64(V) >> 24;
    uint64_t vm = fwd ? nr_valid_mask(1 - pinned, m - pinned + 1)
                      : nr_valid_mask(pinned - m + 1, pinned + 1);
    // N rows as seen from the pinned end: base t of V is read[pinned - 1 + t] (forward) or
    // read[pinned - t] (backward)
    uint32_t nt;
    if (fwd) nt = (uint32_t)(pinned >= 1 ? nm >> (pinned - 1) : nm << (1 - pinned));
    else nt = (pinned >= 0 && pinned <= 63) ? (uint32_t)(nr_rev_bits64(nm) >> (63 - pinned))
                                            : (uint32_t)(nr_rev_bits64(nm) << (pinned - 63));
    vm &= ~nr_spread_even(nt);
    return nr_lv16(V, vm, fwd ? core : nr_rev_bases32(core)) != 0;
}

NR_HD int nr_score16n(const uint32_t *rdp, uint64_t nm, int m, uint32_t core, int padL, int padR,
                      int p, int *umi)
{
    const int r0 = nr_rows_first(p);
    return nr_nfa16n_w(nr_window64(rdp, r0), (uint32_t)(nm >> r0), m, core, padL, padR, r0,
                       nr_rows_last(p, m), umi);
}

NR_HD int nr_verify16n(const uint32_t *rdp, uint64_t nm, int m, uint32_t core, int padL, int padR,
                       int p, const nr_probe_t &t, int *umi)
{
    if (!nr_prefilter16n(rdp, nm, m, core, p, t)) { *umi = -1; return 3; }
    return nr_score16n(rdp, nm, m, core, padL, padR, p, umi);
}
