// nr_ex_common.cuh -- pieces shared by the kernels that resolve a candidate against the whole
// whitelist (nr_match_exhaustive.cu, nr_match_deep.cu): the (score, tie count, smallest pair)
// accumulator, the candidate's byte codes, the scalar pair DP that yields the UMI column, and
// the result writer.
#pragma once
#include "nr_common.cuh"

namespace nr_ex {

struct Best {
    int score;       // best AS
    uint32_t cnt;    // pairs attaining it
    uint32_t key;    // smallest (idx << 1 | strand) among them
};

__device__ __forceinline__ void best_add(Best &b, int score, uint32_t key)
{
    if (score > b.score) { b.score = score; b.cnt = 1; b.key = key; }
    else if (score == b.score) { b.cnt++; b.key = min(b.key, key); }
}

__device__ __forceinline__ void best_merge(Best &a, int score, uint32_t cnt, uint32_t key)
{
    if (score > a.score) { a.score = score; a.cnt = cnt; a.key = key; }
    else if (score == a.score) { a.cnt += cnt; a.key = min(a.key, key); }
}

__device__ __forceinline__ Best block_reduce_best(Best b, Best *sh /* >= 32 */)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        int s = __shfl_xor_sync(0xffffffffu, b.score, o);
        uint32_t c = __shfl_xor_sync(0xffffffffu, b.cnt, o);
        uint32_t k = __shfl_xor_sync(0xffffffffu, b.key, o);
        // symmetric merge (both lanes end with the same value)
        if (s > b.score) { b.score = s; b.cnt = c; b.key = k; }
        else if (s == b.score) { b.cnt += c; b.key = min(b.key, k); }
    }
    int warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
    if (nr_lane() == 0) sh[warp] = b;
    __syncthreads();
    Best r = sh[0];
    for (int w = 1; w < nw; w++) best_merge(r, sh[w].score, sh[w].cnt, sh[w].key);
    __syncthreads();
    return r;
}

// byte codes of the candidate (forward and reverse complement), 4 = N
__device__ __forceinline__ void load_codes(const uint4 *bases, const uint64_t *nmask,
                                           uint64_t cand, int m, uint8_t *cf, uint8_t *cr)
{
    const uint32_t *bw = reinterpret_cast<const uint32_t *>(bases + cand);
    uint64_t nm = nmask ? nmask[cand] : 0ull;
    for (int p = threadIdx.x; p < NR_MAX_QUERY; p += blockDim.x) {
        int c = 4;
        if (p < m) {
            c = (int)((bw[p >> 4] >> ((p & 15) * 2)) & 3u);
            if ((nm >> p) & 1ull) c = 4;
        }
        cf[p] = (uint8_t)c;
        if (p < m) cr[m - 1 - p] = (uint8_t)(c > 3 ? 4 : 3 - c);
        else cr[p] = 4;
    }
}

// generic scalar pair DP on byte codes (any L <= 32, N aware): AS and UMI column.
static __device__ int pair_dp_codes(const uint8_t *q, int m, uint32_t lo, uint32_t hi, uint32_t nm,
                             int L, int padL, int padR, int *iend)
{
    int C[NR_MAX_CORE + 1];
    for (int j = 0; j <= L; j++) C[j] = 0;
    int a_r = -max(0, m - padR), arg = 0;
    for (int i = 1; i <= m; i++) {
        int qq = q[i - 1];
        int diag = C[0];
        C[0] = -max(0, i - padL);
        for (int j = 1; j <= L; j++) {
            int col = j - 1;
            int x = (int)(((col < 16 ? lo >> (2 * col) : hi >> (2 * (col - 16)))) & 3u);
            int s = (qq > 3 || ((nm >> col) & 1u)) ? 0 : (x == qq ? 1 : -1);
            int v = max(diag + s, max(C[j], C[j - 1]) - 1);
            diag = C[j];
            C[j] = v;
        }
        int v = C[L] - max(0, m - i - padR);
        if (v > a_r) { a_r = v; arg = i; }
    }
    int a_in = -1000;
    for (int j = 1; j < L; j++) a_in = max(a_in, C[j]);
    int as = max(-max(0, m - padL), max(a_in, a_r));
    *iend = (a_r == as) ? arg : -1;
    return as;
}

__device__ __forceinline__ void write_result(const Best &r, uint64_t cand, int m,
                                             const uint8_t *cf, const uint32_t *lo,
                                             const uint32_t *hi, const uint32_t *nmw, int L,
                                             int padL, int padR, int min_score, int32_t *o_idx,
                                             int8_t *o_score, uint8_t *o_nbest, uint8_t *o_flags,
                                             uint8_t *o_umi)
{
    int32_t idx = (int32_t)(r.key >> 1);
    int strand = (int)(r.key & 1u);
    uint8_t fl = NR_FLAG_EXHAUSTIVE;
    if (r.cnt > 1) fl |= NR_FLAG_TIE;
    if (strand) fl |= NR_FLAG_RC;
    if (r.score < min_score) fl |= NR_FLAG_BELOW;
    int u = -1;
    if (!strand) {
        int as = pair_dp_codes(cf, m, lo[idx], hi ? hi[idx] : 0u, nmw ? nmw[idx] : 0u, L, padL,
                               padR, &u);
        (void)as;
    }
    if (u < 0) fl |= NR_FLAG_NO_UMI;
    o_idx[cand] = idx;
    o_score[cand] = (int8_t)r.score;
    o_nbest[cand] = (uint8_t)min(r.cnt, 255u);
    o_flags[cand] = fl;
    o_umi[cand] = (uint8_t)(u < 0 ? NR_UMI_NONE : u);
}



}  // namespace nr_ex
