// nr_bitslice_core.h -- bit-parallel form of the semi-global DP that defines parity with the
// oracle (scripts/barcode_align.sh:18-33 of the reference), written once for the
// sm_100a kernel (nr_match_exhaustive.cu) and for the host build that tests/ compiles with g++
// (tests/emul/bitslice_emul.cpp).  No CUDA runtime calls in here.
//
// 32 whitelist entries ride in the 32 bit lanes of a word; the read is the same for all of them.
// In cost form (AS = L - cost): D[i][j] = cheapest way to have consumed read rows 1..i and core
// columns 1..j,
//     D[i][j] = min(D[i-1][j-1] + s, D[i-1][j] + 1, D[i][j-1] + 2),
//     s = 0 match / 2 mismatch / 1 when either side is N,
//     D[0][j] = j (columns hanging over the read start), D[i][0] = max(0, i - padL).
// The weights are not unit cost, so Myers' single delta bit does not carry; BitPAl-style the cell
// is kept as two bounded differences of two bits each,
//     a = D[i][j-1] - D[i-1][j-1] + 2  in 0..3   (handed to the right neighbour)
//     b = D[i-1][j] - D[i-1][j-1] + 1  in 0..3   (handed to the row below),
// and with d = D[i][j] - D[i-1][j-1] = min(s, a, b) in 0..2 the cell emits a' = d - b + 3 and
// b' = d - a + 3: a fixed boolean function of five bits per lane, ten LOP3 per 32 cells
// (nr_bs_cell).  Absolute values are only needed where an alignment may end: down the last column
// (read suffix beyond padR costs 1 per base) and along the last row (remaining columns hang over
// the read end at 1 each); there the differences are summed into 7 bit planes and folded into a
// running minimum (nr_bs_add_sext, nr_bs_min).
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define NR_BS_HD __host__ __device__ __forceinline__
#else
#define NR_BS_HD static inline
#endif

#define NR_BS_PLANES 7      // costs stay below 128: D <= L + m <= 96, penalties included

// 32 x 32 bit transpose, bit 0 = column 0: afterwards bit i of r[k] is bit k of the old r[i]
NR_BS_HD void nr_bs_transpose32(uint32_t r[32])
{
#define NR_BS_TSTAGE(J, MASK)                                          \
    _Pragma("unroll") for (int k = 0; k < 32; k++)                     \
        if (!(k & (J))) {                                              \
            const uint32_t t = ((r[k] >> (J)) ^ r[k | (J)]) & (MASK);  \
            r[k] ^= t << (J);                                          \
            r[k | (J)] ^= t;                                           \
        }
    NR_BS_TSTAGE(16, 0x0000FFFFu)
    NR_BS_TSTAGE(8, 0x00FF00FFu)
    NR_BS_TSTAGE(4, 0x0F0F0F0Fu)
    NR_BS_TSTAGE(2, 0x33333333u)
    NR_BS_TSTAGE(1, 0x55555555u)
#undef NR_BS_TSTAGE
}

// One DP cell of 32 entries.  In: (a1 a0) from the left neighbour, (b1 b0) from the row above,
// e = lanes whose column matches the read base, n1 = lanes where the pair scores 0 (N on either
// side; e is clear there).  Out: (a1 a0) for the right neighbour, (b1 b0) for the row below.
//   d == 0  <=>  e | a == 0 | b == 0
//   d == 1  <=>  not that, and (n1 | a == 1 | b == 1)      (a, b >= 1 here, so a == 1 <=> !a1)
//   a' = 3 - b (d = 0), 4 - b (d = 1), 5 - b (d = 2); b' the same with a
NR_BS_HD void nr_bs_cell(uint32_t &a0, uint32_t &a1, uint32_t &b0, uint32_t &b1, uint32_t e, uint32_t n1)
{
    const uint32_t d0 = e | ~(b1 | b0) | ~(a1 | a0);
    const uint32_t u = ~d0 & (n1 | ~a1 | ~b1);
    const uint32_t x = d0 | (u & b0), y = d0 | (u & a0);
    const uint32_t na0 = ~(b0 ^ u), nb0 = ~(a0 ^ u);
    const uint32_t na1 = ~(b1 & x), nb1 = ~(a1 & y);
    a0 = na0; a1 = na1; b0 = nb0; b1 = nb1;
}

NR_BS_HD void nr_bs_set_const(uint32_t P[NR_BS_PLANES], int v)
{
#pragma unroll
    for (int k = 0; k < NR_BS_PLANES; k++) P[k] = ((v >> k) & 1) ? 0xFFFFFFFFu : 0u;
}

// Y += the two's-complement number whose bit 0 is x0, bit 1 is x1 and every higher bit xh
NR_BS_HD void nr_bs_add_sext(uint32_t Y[NR_BS_PLANES], uint32_t x0, uint32_t x1, uint32_t xh)
{
    uint32_t c = Y[0] & x0;
    Y[0] ^= x0;
    uint32_t s = Y[1] ^ x1 ^ c;
    c = (Y[1] & x1) | (c & (Y[1] | x1));
    Y[1] = s;
#pragma unroll
    for (int k = 2; k < NR_BS_PLANES; k++) {
        s = Y[k] ^ xh ^ c;
        c = (Y[k] & xh) | (c & (Y[k] | xh));
        Y[k] = s;
    }
}

// Y += v - 2 (v in 0..3 given as planes v1 v0): bits ...(~v1)(~v1)(v0)
NR_BS_HD void nr_bs_add_m2(uint32_t Y[NR_BS_PLANES], uint32_t v0, uint32_t v1)
{
    nr_bs_add_sext(Y, v0, ~v1, ~v1);
}
// Y += v - 3: 0 -> ...101, 1 -> ...110, 2 -> ...111, 3 -> 0
NR_BS_HD void nr_bs_add_m3(uint32_t Y[NR_BS_PLANES], uint32_t v0, uint32_t v1)
{
    nr_bs_add_sext(Y, ~v0, v1 ^ v0, ~(v1 & v0));
}

// M = min(M, Y) per lane (values are non-negative)
NR_BS_HD void nr_bs_min(uint32_t M[NR_BS_PLANES], const uint32_t Y[NR_BS_PLANES])
{
    uint32_t lt = ~Y[0] & M[0];
#pragma unroll
    for (int k = 1; k < NR_BS_PLANES; k++) lt = (~Y[k] & M[k]) | (~(Y[k] ^ M[k]) & lt);
#pragma unroll
    for (int k = 0; k < NR_BS_PLANES; k++) M[k] = (Y[k] & lt) | (M[k] & ~lt);
}

// Match masks of 16 columns from their transposed planes (r[2c] = low bit of column c's base over
// the 32 entries, r[2c + 1] = high bit): eq[(base * L + col0 + c) * stride] = lanes whose column
// carries `base` and is not N.  nmp (HAS_N) = the N planes, already stored: nmp[col * stride].
template <int L, bool HAS_N>
NR_BS_HD void nr_bs_build_eq16(const uint32_t r[32], int col0, uint32_t *eq, const uint32_t *nmp, int stride)
{
#pragma unroll
    for (int c = 0; c < 16; c++) {
        const int col = col0 + c;
        const uint32_t lo = r[2 * c], hi = r[2 * c + 1];
        const uint32_t keep = HAS_N ? ~nmp[col * stride] : 0xFFFFFFFFu;
        eq[(0 * L + col) * stride] = ~lo & ~hi & keep;
        eq[(1 * L + col) * stride] = lo & ~hi & keep;
        eq[(2 * L + col) * stride] = ~lo & hi & keep;
        eq[(3 * L + col) * stride] = lo & hi & keep;
    }
}

// The whole DP of one strand of the read against 32 entries.  q: byte codes of the strand (4 = N),
// m rows.  M receives the cost planes: AS of entry lane l = L - (bits l of M[6..0]).
template <int L, bool HAS_N>
NR_BS_HD void nr_bs_word_strand(const uint32_t *eq, const uint32_t *nmp, int stride, const uint8_t *q,
                                int m, int padL, int padR, uint32_t M[NR_BS_PLANES])
{
    uint32_t b0[L], b1[L];
#pragma unroll
    for (int j = 0; j < L; j++) { b0[j] = 0u; b1[j] = 0xFFFFFFFFu; }     // D[0][j] - D[0][j-1] = 1
    // Y = D[i][L] + max(0, m - i - padR): the alignment leaves the core after row i
    const int y_init = L + (m - padR > 0 ? m - padR : 0);
    // read entirely in front of the core (column 0 of the last row), L columns hanging over
    const int t_init = L + (m - padL > 0 ? m - padL : 0);
    uint32_t Y[NR_BS_PLANES];
    nr_bs_set_const(Y, y_init);
    nr_bs_set_const(M, y_init < t_init ? y_init : t_init);
#pragma unroll 1
    for (int i = 1; i <= m; i++) {
        const int c = q[i - 1];
        uint32_t a1 = 0xFFFFFFFFu, a0 = i > padL ? 0xFFFFFFFFu : 0u;     // D[i][0] - D[i-1][0] in {0, 1}
        if (c < 4) {
            const uint32_t *pe = eq + c * L * stride;
#pragma unroll
            for (int j = 0; j < L; j++)
                nr_bs_cell(a0, a1, b0[j], b1[j], pe[j * stride], HAS_N ? nmp[j * stride] : 0u);
        } else {
#pragma unroll
            for (int j = 0; j < L; j++) nr_bs_cell(a0, a1, b0[j], b1[j], 0u, 0xFFFFFFFFu);
        }
        // (a1 a0) now is D[i][L] - D[i-1][L] + 2; the suffix penalty drops by one while i <= m - padR
        if (i <= m - padR) nr_bs_add_m3(Y, a0, a1);
        else nr_bs_add_m2(Y, a0, a1);
        nr_bs_min(M, Y);
    }
    // the read ends inside the core at column j (1 <= j < L): D[m][j] + (L - j)
    uint32_t T[NR_BS_PLANES];
    nr_bs_set_const(T, t_init);
#pragma unroll
    for (int j = 1; j < L; j++) {
        nr_bs_add_m2(T, b0[j - 1], b1[j - 1]);      // (D[m][j] - D[m][j-1] + 1) - 2
        nr_bs_min(M, T);
    }
}

// smallest cost among the lanes of `valid` and the lanes attaining it
NR_BS_HD uint32_t nr_bs_lane_min(const uint32_t M[NR_BS_PLANES], uint32_t valid, int *value)
{
    uint32_t cand = valid;
    int v = 0;
#pragma unroll
    for (int k = NR_BS_PLANES - 1; k >= 0; k--) {
        const uint32_t t = cand & ~M[k];
        if (t) cand = t; else v |= 1 << k;
    }
    *value = v;
    return cand;
}
