// nr_common.cuh -- shared definitions of the B200 barcode matcher (sm_100a).
//
// Scoring (what scripts/barcode_align.sh:18-33 configures STAR with):
//   reference R = N^padL + core + N^padR, query Q consumed end to end, free start/end in R,
//   match +1, mismatch -1, gap -1 per base, N on either side scores 0.
// Packed layouts:
//   candidate  16 B, 2 bit/base (A0 C1 G2 T3), base k at bit 2k of the little-endian record;
//              meta byte = len | 0x80 (contains non-ACGT) or 0xFF (too long); nmask u64
//   whitelist  core columns 0..15 in `lo`, 16..31 in `hi` (same bit order), N columns in `nm`
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/nanoranger_b200.h"

#define NR_CHECK_CUDA(expr)                                                        \
    do {                                                                           \
        cudaError_t _e = (expr);                                                   \
        if (_e != cudaSuccess) {                                                   \
            nr_set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e),   \
                         __FILE__, __LINE__);                                      \
            return NR_ECUDA;                                                       \
        }                                                                          \
    } while (0)

void nr_set_error(const char *fmt, ...);

struct nr_whitelist {
    int device;
    uint64_t n;
    uint32_t L, pad_l, pad_r;
    int has_n;       // some entry contains N
    int has_index;   // quarter-key seed index present (L == 16 && !has_n)
    uint32_t *d_lo;  // n
    uint32_t *d_hi;  // n (L > 16) or nullptr
    uint32_t *d_nm;  // n (has_n) or nullptr
    // seed index: for j in 0..3 the 24-bit key made of the three 4-base quarters other than
    // quarter j.  bm[j][w] = {bitmap word of keys 32w..32w+31, number of index rows with a
    // smaller key}; ents[j] = rows {entry idx, core} sorted by key_j (2^19 + 1 bm words).
    uint2 *d_bm[4];
    uint2 *d_ents[4];
    size_t bytes;
};

#define NR_BM_WORDS (1u << 19)

// ---------------------------------------------------------------------------------------
// device helpers

__device__ __forceinline__ uint32_t nr_lane() { return threadIdx.x & 31u; }

// key of a 16-column core with quarter j (byte j) removed
__device__ __host__ __forceinline__ uint32_t nr_key_drop(uint32_t core, int j)
{
    switch (j) {
    case 0: return core >> 8;
    case 1: return (core & 0xFFu) | ((core >> 8) & 0xFFFF00u);
    case 2: return (core & 0xFFFFu) | ((core >> 8) & 0xFF0000u);
    default: return core & 0xFFFFFFu;
    }
}

// 32-bit window (16 bases) starting at base position p of a padded packed read:
// rd[0] is a zero word, rd[1..4] hold bases 0..63, rd[5..6] are zero; valid for -16 <= p <= 64.
__device__ __forceinline__ uint32_t nr_window(const uint32_t *rd, int p)
{
    int q = p + 16;
    int w = q >> 4;
    uint32_t s = (uint32_t)(q & 15) * 2u;
    return __funnelshift_r(rd[w], rd[w + 1], s);
}

__device__ __forceinline__ int nr_base(const uint32_t *rd, int p)  // 0 <= p < 64
{
    return (int)((rd[1 + (p >> 4)] >> ((p & 15) * 2)) & 3u);
}

// reverse complement of the low m bases of a 4-word packed read (no N): out[0..3]
__device__ __forceinline__ void nr_revcomp_words(const uint32_t in[4], int m, uint32_t out[4])
{
    uint32_t r[6];
#pragma unroll
    for (int k = 0; k < 4; k++) {
        uint32_t x = ~in[3 - k];
        x = __brev(x);
        r[k] = ((x & 0x55555555u) << 1) | ((x >> 1) & 0x55555555u);
    }
    r[4] = 0; r[5] = 0;
    // r = revcomp of all 64 positions; drop the 64-m leading (garbage) bases
    int sh = 64 - m;                // 0..64
    int ws = sh >> 4;
    uint32_t bs = (uint32_t)(sh & 15) * 2u;
#pragma unroll
    for (int k = 0; k < 4; k++) {
        uint32_t a = 0, b = 0;
#pragma unroll
        for (int t = 0; t < 6; t++) {
            if (t == k + ws) a = r[t];
            if (t == k + ws + 1) b = r[t];
        }
        out[k] = __funnelshift_r(a, b, bs);
    }
    // clear bases >= m
#pragma unroll
    for (int k = 0; k < 4; k++) {
        int lo = k * 16;
        if (m <= lo) out[k] = 0;
        else if (m < lo + 16) out[k] &= (1u << ((m - lo) * 2)) - 1u;
    }
}

// Exact score of one (candidate strand, 16-column N-free core) pair: plain DP over the core
// columns with the closed-form pad boundaries (same recurrence as oracle tier 1, restated
// here for the device).  rd = padded packed read (see nr_window), N-free.
// Returns AS; *iend = smallest query row at which an optimal alignment leaves the core
// (= query index aligned to reference column padL+16), or -1.
__device__ __noinline__ int nr_pair_dp16(const uint32_t *rd, int m, uint32_t core, int padL,
                                         int padR, int *iend)
{
    int C[17];
#pragma unroll
    for (int j = 0; j <= 16; j++) C[j] = 0;
    int a_r = -max(0, m - padR), arg = 0;
#pragma unroll 1
    for (int i = 1; i <= m; i++) {
        int q = nr_base(rd, i - 1);
        int diag = C[0];
        C[0] = -max(0, i - padL);
#pragma unroll
        for (int j = 1; j <= 16; j++) {
            int s = (int)((core >> (2 * (j - 1))) & 3u) == q ? 1 : -1;
            int v = max(diag + s, max(C[j], C[j - 1]) - 1);
            diag = C[j];
            C[j] = v;
        }
        int v = C[16] - max(0, m - i - padR);
        if (v > a_r) { a_r = v; arg = i; }
    }
    int a_in = -1000;
#pragma unroll
    for (int j = 1; j < 16; j++) a_in = max(a_in, C[j]);
    int as = max(-max(0, m - padL), max(a_in, a_r));
    *iend = (a_r == as) ? arg : -1;
    return as;
}
