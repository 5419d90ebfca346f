// bitslice_emul.cpp -- host-side build of the bit-parallel brute-force DP (TEST INFRASTRUCTURE).
// Compiled by tests/ with g++ from the header the sm_100a kernel includes (nr_bitslice_core.h), so
// the cell function, the plane adders, the transpose and the lane minimum are the shipped code;
// only the thread / block choreography differs.  Nothing in nanoranger_b200/ links or calls it.
#include <cstdint>
#include <cstring>
#include <vector>

#include "../../nanoranger_b200/csrc/nr_bitslice_core.h"

namespace {

template <int L, bool HAS_N>
void run(const uint32_t *lo, const uint32_t *hi, const uint32_t *nm, int64_t n, int padL, int padR,
         const uint8_t *qf, const uint8_t *qr, int m, int8_t *as_f, int8_t *as_r, int32_t *best)
{
    int b_score = -1000;
    int64_t b_cnt = 0;
    uint32_t b_key = 0xFFFFFFFFu;
    const int64_t nwords = (n + 31) / 32;
    std::vector<uint32_t> eq(4 * L), nmp(L);
    for (int64_t g = 0; g < nwords; g++) {
        const int64_t base = g * 32;
        uint32_t r[32];
        if (HAS_N) {
            for (int k = 0; k < 32; k++) r[k] = nm[base + k < n ? base + k : n - 1];
            nr_bs_transpose32(r);
            for (int j = 0; j < L; j++) nmp[j] = r[j];
        }
        for (int k = 0; k < 32; k++) r[k] = lo[base + k < n ? base + k : n - 1];
        nr_bs_transpose32(r);
        nr_bs_build_eq16<L, HAS_N>(r, 0, eq.data(), nmp.data(), 1);
        if (L == 32) {
            for (int k = 0; k < 32; k++) r[k] = hi[base + k < n ? base + k : n - 1];
            nr_bs_transpose32(r);
            nr_bs_build_eq16<L, HAS_N>(r, L == 32 ? 16 : 0, eq.data(), nmp.data(), 1);
        }
        const uint32_t valid = n - base >= 32 ? 0xFFFFFFFFu : ((1u << (n - base)) - 1u);
        for (int st = 0; st < 2; st++) {
            uint32_t M[NR_BS_PLANES];
            nr_bs_word_strand<L, HAS_N>(eq.data(), nmp.data(), 1, st ? qr : qf, m, padL, padR, M);
            for (int l = 0; l < 32 && base + l < n; l++) {
                int cost = 0;
                for (int k = 0; k < NR_BS_PLANES; k++) cost |= (int)((M[k] >> l) & 1u) << k;
                (st ? as_r : as_f)[base + l] = (int8_t)(L - cost);
            }
            int v;
            const uint32_t at = nr_bs_lane_min(M, valid, &v);
            const int score = L - v;
            const uint32_t key = (uint32_t)((base + __builtin_ctz(at)) << 1) | (uint32_t)st;
            if (score > b_score) { b_score = score; b_cnt = __builtin_popcount(at); b_key = key; }
            else if (score == b_score) { b_cnt += __builtin_popcount(at); if (key < b_key) b_key = key; }
        }
    }
    best[0] = b_score; best[1] = (int32_t)b_cnt; best[2] = (int32_t)(b_key >> 1); best[3] = (int32_t)(b_key & 1u);
}

}  // namespace

// qf / qr: byte codes (4 = N) of the read and of its reverse complement, m of them.
// as_f / as_r: AS of every entry on either strand; best = {score, pairs attaining it, smallest
// entry among them, its strand}.
extern "C" int nr_emul_bitslice(const uint32_t *lo, const uint32_t *hi, const uint32_t *nm, int64_t n,
                                int L, int has_n, int padL, int padR, const uint8_t *qf,
                                const uint8_t *qr, int m, int8_t *as_f, int8_t *as_r, int32_t *best)
{
    if (n < 1 || m < 0 || m > 64) return -1;
    if (L == 16 && !has_n) run<16, false>(lo, hi, nm, n, padL, padR, qf, qr, m, as_f, as_r, best);
    else if (L == 16) run<16, true>(lo, hi, nm, n, padL, padR, qf, qr, m, as_f, as_r, best);
    else if (L == 32 && !has_n) run<32, false>(lo, hi, nm, n, padL, padR, qf, qr, m, as_f, as_r, best);
    else if (L == 32) run<32, true>(lo, hi, nm, n, padL, padR, qf, qr, m, as_f, as_r, best);
    else return -2;
    return 0;
}

// the cell function against integer arithmetic, over every input: returns the number of
// disagreements (0 expected)
extern "C" int nr_emul_bitslice_cell_check(void)
{
    int bad = 0;
    for (int a = 0; a < 4; a++)
        for (int b = 0; b < 4; b++)
            for (int s = 0; s < 3; s++) {
                uint32_t a0 = (a & 1) ? ~0u : 0u, a1 = (a & 2) ? ~0u : 0u;
                uint32_t b0 = (b & 1) ? ~0u : 0u, b1 = (b & 2) ? ~0u : 0u;
                nr_bs_cell(a0, a1, b0, b1, s == 0 ? ~0u : 0u, s == 1 ? ~0u : 0u);
                int d = s < a ? s : a;
                if (b < d) d = b;
                const int ea = d - b + 3, eb = d - a + 3;
                const int ga = (int)(a0 & 1u) | ((int)(a1 & 1u) << 1), gb = (int)(b0 & 1u) | ((int)(b1 & 1u) << 1);
                if (ea != ga || eb != gb || a0 != (ea & 1 ? ~0u : 0u) || b1 != (eb & 2 ? ~0u : 0u)) bad++;
            }
    // the adders and the minimum on random lanes
    uint64_t x = 88172645463325252ull;
    for (int it = 0; it < 2000; it++) {
        int val[32], add[32], other[32];
        uint32_t Y[NR_BS_PLANES] = {0}, O[NR_BS_PLANES] = {0}, v0 = 0, v1 = 0;
        for (int l = 0; l < 32; l++) {
            x ^= x << 13; x ^= x >> 7; x ^= x << 17;
            val[l] = 3 + (int)(x % 100); add[l] = (int)((x >> 20) & 3); other[l] = (int)((x >> 30) % 120);
            for (int k = 0; k < NR_BS_PLANES; k++) {
                Y[k] |= (uint32_t)((val[l] >> k) & 1) << l;
                O[k] |= (uint32_t)((other[l] >> k) & 1) << l;
            }
            v0 |= (uint32_t)(add[l] & 1) << l; v1 |= (uint32_t)(add[l] >> 1) << l;
        }
        const bool m3 = it & 1;
        if (m3) nr_bs_add_m3(Y, v0, v1); else nr_bs_add_m2(Y, v0, v1);
        nr_bs_min(O, Y);
        int lo = 1000;
        for (int l = 0; l < 32; l++) {
            int y = 0, o = 0;
            for (int k = 0; k < NR_BS_PLANES; k++) { y |= (int)((Y[k] >> l) & 1u) << k; o |= (int)((O[k] >> l) & 1u) << k; }
            const int ey = val[l] + add[l] - (m3 ? 3 : 2);
            const int eo = ey < other[l] ? ey : other[l];
            if (y != ey || o != eo) bad++;
            if (eo < lo) lo = eo;
        }
        int v;
        const uint32_t at = nr_bs_lane_min(O, 0xFFFFFFFFu, &v);
        if (v != lo) bad++;
        for (int l = 0; l < 32; l++) {
            int o = 0;
            for (int k = 0; k < NR_BS_PLANES; k++) o |= (int)((O[k] >> l) & 1u) << k;
            if (((at >> l) & 1u) != (uint32_t)(o == lo)) bad++;
        }
    }
    // the transpose
    uint32_t r[32], t[32];
    for (int k = 0; k < 32; k++) { x ^= x << 13; x ^= x >> 7; x ^= x << 17; r[k] = t[k] = (uint32_t)x; }
    nr_bs_transpose32(t);
    for (int i = 0; i < 32; i++)
        for (int k = 0; k < 32; k++)
            if (((t[k] >> i) & 1u) != ((r[i] >> k) & 1u)) bad++;
    return bad;
}
