// nr_match_exhaustive.cu -- score every (entry, strand) pair of a candidate: exact at every
// score, any core length <= 32, N on either side.  This is the kernel that defines parity
// with the oracle; the filtered kernel falls back to it for the candidates it cannot take.
//
// Replaces scripts/barcode_align.sh:14-41 (STAR EndToEnd, unique mappers only).
//
// DP on the core columns with closed-form pad boundaries, in the
// gap-free form  H'[i][j] = S[i][j] + i + j:
//     H'[i][j] = max(H'[i-1][j-1] + s + 2, H'[i-1][j], H'[i][j-1]),  s+2 in {3 match, 1 mismatch, 2 N}
// so one cell is one VIMNMX + one VIADDMNMX.  The L = 16 / N-free-whitelist kernel keeps two
// entries per thread in the s16x2 halves of every register (DPX), the 4 x 16 score profile
// of the entry pair in registers, and walks the query rows with a warp-uniform base switch.
#include <cstdlib>

#include "nr_common.cuh"
#include "nr_ex_common.cuh"
#include "nr_bitslice_core.h"

namespace {

using namespace nr_ex;

// ---- short lists: split the whitelist over blocks -------------------------------------------
// When fewer candidates than blocks are to be resolved (the usual case for the list the filtered
// kernel leaves: reads with N, very short reads), each candidate's scan of the whitelist is cut
// into S slices handled by S blocks.  Partials go to `part` (one slot per work item); the last
// block to arrive for a candidate (counter in `done`) merges them and writes the result.
// total * S <= gridDim.x, so a block handles at most one work item when S > 1.
#define NR_EX_MAXS 64

struct ExScratch {
    uint32_t *done;   // NR_EX_MAXGRID counters, zero between launches (the merging block resets)
    uint4 *part;      // NR_EX_MAXGRID partials {score, cnt, key, -}
};

__device__ __forceinline__ uint32_t ex_slices(uint64_t total, const ExScratch &sc)
{
    if (!sc.done || total == 0 || total >= gridDim.x || gridDim.x > NR_EX_MAXGRID) return 1u;
    uint32_t S = gridDim.x / (uint32_t)total;
    return S > NR_EX_MAXS ? NR_EX_MAXS : S;
}

// returns true in the block that must write the candidate's result; `r` then holds the merge
__device__ __forceinline__ bool ex_merge(Best &r, uint64_t w, uint64_t it, uint32_t S,
                                         const ExScratch &sc, int *sh_flag)
{
    if (S == 1) return true;
    if (threadIdx.x == 0) {
        sc.part[w] = make_uint4((uint32_t)r.score, r.cnt, r.key, 0u);
        __threadfence();
        uint32_t prev = atomicAdd(&sc.done[it], 1u);
        *sh_flag = (prev == S - 1);
    }
    __syncthreads();
    const bool last = *sh_flag != 0;
    if (last && threadIdx.x == 0) {
        __threadfence();
        Best m; m.score = -1000; m.cnt = 0; m.key = 0xFFFFFFFFu;
        for (uint32_t k = 0; k < S; k++) {
            uint4 v = __ldcg(sc.part + it * S + k);
            best_merge(m, (int)v.x, v.y, v.z);
        }
        r = m;
        sc.done[it] = 0u;
    }
    __syncthreads();
    return last;
}

#define P2(x) ((uint32_t)((x) & 0xFFFF) * 0x10001u)

// ---- L = 16, N-free whitelist: DPX s16x2, two entries per thread ---------------------------
__global__ void __launch_bounds__(256, 2)
nr_match_exhaustive16_kernel(const uint32_t *__restrict__ wl, uint32_t n, int padL, int padR,
                             const uint4 *__restrict__ bases, const uint8_t *__restrict__ meta,
                             const uint64_t *__restrict__ nmask,
                             const uint32_t *__restrict__ list,
                             const uint32_t *__restrict__ list_count, uint64_t n_cand,
                             int min_score, int32_t *__restrict__ o_idx,
                             int8_t *__restrict__ o_score, uint8_t *__restrict__ o_nbest,
                             uint8_t *__restrict__ o_flags, uint8_t *__restrict__ o_umi,
                             ExScratch sc)
{
    __shared__ uint8_t cf[NR_MAX_QUERY], cr[NR_MAX_QUERY];
    __shared__ Best shb[32];
    __shared__ int sh_flag;
    uint64_t total = list ? (uint64_t)*list_count : n_cand;
    uint32_t npairs = (n + 1) >> 1;
    const uint32_t S = ex_slices(total, sc);
    const uint32_t per = (npairs + S - 1) / S;
    for (uint64_t w = blockIdx.x; w < total * S; w += gridDim.x) {
        const uint64_t it = w / S;
        const uint32_t p_lo = (uint32_t)(w % S) * per;
        const uint32_t p_hi = min(npairs, p_lo + per);
        uint64_t cand = list ? (uint64_t)list[it] : it;
        uint8_t mt = meta[cand];
        if (mt == 0xFF) {
            if (threadIdx.x == 0) {
                o_idx[cand] = -1; o_score[cand] = NR_SCORE_BELOW; o_nbest[cand] = 0;
                o_flags[cand] = NR_FLAG_TOO_LONG | NR_FLAG_BELOW | NR_FLAG_NO_UMI;
                o_umi[cand] = NR_UMI_NONE;
            }
            continue;
        }
        int m = mt & 0x7F;
        __syncthreads();
        load_codes(bases, (mt & 0x80) ? nmask : nullptr, cand, m, cf, cr);
        __syncthreads();
        Best b; b.score = -1000; b.cnt = 0; b.key = 0xFFFFFFFFu;
        const uint32_t aleft = P2(-max(0, m - padL));
        for (uint32_t p = p_lo + threadIdx.x; p < p_hi; p += blockDim.x) {
            uint32_t ea = wl[2 * p];
            bool vb = 2 * p + 1 < n;
            uint32_t eb = vb ? wl[2 * p + 1] : ea;
            // score profile: P[q][j] = s16x2 of (s + 2) = 1 + 2 * match
            uint32_t P0[16], P1[16], P2_[16], P3[16];
            {
                uint32_t la = ea & 0x55555555u, ha = (ea >> 1) & 0x55555555u;
                uint32_t lb = eb & 0x55555555u, hb = (eb >> 1) & 0x55555555u;
                uint32_t w0 = ((~la & ~ha) & 0x55555555u) | (((~lb & ~hb) & 0x55555555u) << 1);
                uint32_t w1 = ((la & ~ha)) | ((lb & ~hb) << 1);
                uint32_t w2 = ((~la & ha)) | ((~lb & hb) << 1);
                uint32_t w3 = ((la & ha)) | ((lb & hb) << 1);
#pragma unroll
                for (int j = 0; j < 16; j++) {
                    P0[j] = 0x00010001u + (((((w0 >> (2 * j)) & 3u) * 0x8001u) & 0x00010001u) << 1);
                    P1[j] = 0x00010001u + (((((w1 >> (2 * j)) & 3u) * 0x8001u) & 0x00010001u) << 1);
                    P2_[j] = 0x00010001u + (((((w2 >> (2 * j)) & 3u) * 0x8001u) & 0x00010001u) << 1);
                    P3[j] = 0x00010001u + (((((w3 >> (2 * j)) & 3u) * 0x8001u) & 0x00010001u) << 1);
                }
            }
#pragma unroll 1
            for (int strand = 0; strand < 2; strand++) {
                const uint8_t *q = strand ? cr : cf;
                uint32_t C[17];
#pragma unroll
                for (int j = 0; j <= 16; j++) C[j] = P2(j);
                uint32_t ar = P2(-max(0, m - padR));
#pragma unroll 1
                for (int i = 1; i <= m; i++) {
                    int qq = q[i - 1];
                    uint32_t diag = C[0];
                    C[0] = P2(min(i, padL));
#define NR_ROW(PX)                                                              \
    _Pragma("unroll") for (int j = 1; j <= 16; j++) {                           \
        uint32_t t = C[j];                                                      \
        C[j] = __viaddmax_s16x2(diag, PX, __vmaxs2(C[j], C[j - 1]));            \
        diag = t;                                                               \
    }
                    if (qq == 0) { NR_ROW(P0[j - 1]) }
                    else if (qq == 1) { NR_ROW(P1[j - 1]) }
                    else if (qq == 2) { NR_ROW(P2_[j - 1]) }
                    else if (qq == 3) { NR_ROW(P3[j - 1]) }
                    else { NR_ROW(0x00020002u) }
#undef NR_ROW
                    uint32_t off = P2(i + 16 + max(0, m - i - padR));
                    ar = __vmaxs2(ar, __vsub2(C[16], off));
                }
                uint32_t ain = P2(-1000);
#pragma unroll
                for (int j = 1; j < 16; j++) ain = __vmaxs2(ain, __vsub2(C[j], P2(m + j)));
                uint32_t as2 = __vmaxs2(aleft, __vmaxs2(ain, ar));
                int sa = (int)(int16_t)(as2 & 0xFFFFu);
                int sb = (int)(int16_t)(as2 >> 16);
                best_add(b, sa, ((2 * p) << 1) | (uint32_t)strand);
                if (vb) best_add(b, sb, ((2 * p + 1) << 1) | (uint32_t)strand);
            }
        }
        Best r = block_reduce_best(b, shb);
        if (ex_merge(r, w, it, S, sc, &sh_flag) && threadIdx.x == 0)
            write_result(r, cand, m, cf, wl, nullptr, nullptr, 16, padL, padR, min_score, o_idx,
                         o_score, o_nbest, o_flags, o_umi);
    }
}

// ---- generic: any L <= 32, N columns in the whitelist; DPX s16x2, two entries per thread -----
// The (s + 2) profile of the thread's entry pair -- 4 read bases x L columns, both entries in
// the halves of a word -- lives in shared memory ([base][column][thread]: conflict-free), so a
// cell pair costs one LDS + VIMNMX + VIADDMNMX like in the L = 16 kernel.  A read N scores 0
// against everything: constant term.
#define NR_EXG_THREADS 128

template <int LT>      // LT > 0: core length known at compile time (32: slide-seq); 0: runtime L
__global__ void __launch_bounds__(NR_EXG_THREADS, 3)
nr_match_exhaustive_generic_kernel(const uint32_t *__restrict__ wlo,
                                   const uint32_t *__restrict__ whi,
                                   const uint32_t *__restrict__ wnm, uint32_t n, int L_rt, int padL,
                                   int padR, const uint4 *__restrict__ bases,
                                   const uint8_t *__restrict__ meta,
                                   const uint64_t *__restrict__ nmask,
                                   const uint32_t *__restrict__ list,
                                   const uint32_t *__restrict__ list_count, uint64_t n_cand,
                                   int min_score, int32_t *__restrict__ o_idx,
                                   int8_t *__restrict__ o_score, uint8_t *__restrict__ o_nbest,
                                   uint8_t *__restrict__ o_flags, uint8_t *__restrict__ o_umi,
                                   ExScratch sc)
{
    extern __shared__ uint32_t prof[];          // [4][L][NR_EXG_THREADS]
    const int L = LT ? LT : L_rt;
    __shared__ uint8_t cf[NR_MAX_QUERY], cr[NR_MAX_QUERY];
    __shared__ Best shb[32];
    __shared__ int sh_flag;
    uint64_t total = list ? (uint64_t)*list_count : n_cand;
    const uint32_t npairs = (n + 1) >> 1;
    const uint32_t S = ex_slices(total, sc);
    const uint32_t per = (npairs + S - 1) / S;
    uint32_t *mine = prof + threadIdx.x;
    for (uint64_t w = blockIdx.x; w < total * S; w += gridDim.x) {
        const uint64_t it = w / S;
        const uint32_t p_lo = (uint32_t)(w % S) * per;
        const uint32_t p_hi = min(npairs, p_lo + per);
        uint64_t cand = list ? (uint64_t)list[it] : it;
        uint8_t mt = meta[cand];
        if (mt == 0xFF) {
            if (threadIdx.x == 0) {
                o_idx[cand] = -1; o_score[cand] = NR_SCORE_BELOW; o_nbest[cand] = 0;
                o_flags[cand] = NR_FLAG_TOO_LONG | NR_FLAG_BELOW | NR_FLAG_NO_UMI;
                o_umi[cand] = NR_UMI_NONE;
            }
            continue;
        }
        int m = mt & 0x7F;
        __syncthreads();
        load_codes(bases, (mt & 0x80) ? nmask : nullptr, cand, m, cf, cr);
        __syncthreads();
        Best b; b.score = -1000; b.cnt = 0; b.key = 0xFFFFFFFFu;
        const uint32_t aleft = P2(-max(0, m - padL));
        for (uint32_t p = p_lo + threadIdx.x; p < p_hi; p += blockDim.x) {
            const uint32_t ia = 2 * p;
            const bool vb = ia + 1 < n;
            const uint32_t ib = vb ? ia + 1 : ia;
            const uint32_t la = wlo[ia], lb = wlo[ib];
            const uint32_t ha = whi ? whi[ia] : 0u, hb = whi ? whi[ib] : 0u;
            const uint32_t na = wnm ? wnm[ia] : 0u, nb = wnm ? wnm[ib] : 0u;
            // profile of the pair: term = 3 match, 1 mismatch, 2 where the entry column is N
            for (int col = 0; col < L; col++) {
                const uint32_t ca = ((col < 16 ? la >> (2 * col) : ha >> (2 * (col - 16)))) & 3u;
                const uint32_t cb = ((col < 16 ? lb >> (2 * col) : hb >> (2 * (col - 16)))) & 3u;
                const bool a_n = (na >> col) & 1u, b_n = (nb >> col) & 1u;
#pragma unroll
                for (uint32_t q = 0; q < 4; q++) {
                    const uint32_t ta = a_n ? 2u : (ca == q ? 3u : 1u);
                    const uint32_t tb = b_n ? 2u : (cb == q ? 3u : 1u);
                    mine[(q * (uint32_t)L + (uint32_t)col) * NR_EXG_THREADS] = ta | (tb << 16);
                }
            }
#pragma unroll 1
            for (int strand = 0; strand < 2; strand++) {
                const uint8_t *q = strand ? cr : cf;
                uint32_t C[NR_MAX_CORE + 1];
#pragma unroll
                for (int j = 0; j <= NR_MAX_CORE; j++) C[j] = P2(j);
                uint32_t ar = P2(-max(0, m - padR));
#pragma unroll 1
                for (int i = 1; i <= m; i++) {
                    const int qq = q[i - 1];
                    uint32_t diag = C[0];
                    C[0] = P2(min(i, padL));
                    if (qq < 4) {
                        const uint32_t *pq = mine + (uint32_t)qq * (uint32_t)L * NR_EXG_THREADS;
#pragma unroll
                        for (int j = 1; j <= NR_MAX_CORE; j++) {
                            if (j <= L) {
                                const uint32_t t = C[j];
                                C[j] = __viaddmax_s16x2(diag, pq[(j - 1) * NR_EXG_THREADS],
                                                        __vmaxs2(C[j], C[j - 1]));
                                diag = t;
                            }
                        }
                    } else {
#pragma unroll
                        for (int j = 1; j <= NR_MAX_CORE; j++) {
                            if (j <= L) {
                                const uint32_t t = C[j];
                                C[j] = __viaddmax_s16x2(diag, 0x00020002u, __vmaxs2(C[j], C[j - 1]));
                                diag = t;
                            }
                        }
                    }
                    uint32_t cl = C[0];
#pragma unroll
                    for (int j = 1; j <= NR_MAX_CORE; j++) if (j == L) cl = C[j];
                    const uint32_t off = P2(i + L + max(0, m - i - padR));
                    ar = __vmaxs2(ar, __vsub2(cl, off));
                }
                uint32_t ain = P2(-1000);
#pragma unroll
                for (int j = 1; j < NR_MAX_CORE; j++)
                    if (j < L) ain = __vmaxs2(ain, __vsub2(C[j], P2(m + j)));
                const uint32_t as2 = __vmaxs2(aleft, __vmaxs2(ain, ar));
                const int sa = (int)(int16_t)(as2 & 0xFFFFu);
                const int sb = (int)(int16_t)(as2 >> 16);
                best_add(b, sa, (ia << 1) | (uint32_t)strand);
                if (vb) best_add(b, sb, (ib << 1) | (uint32_t)strand);
            }
        }
        Best r = block_reduce_best(b, shb);
        if (ex_merge(r, w, it, S, sc, &sh_flag) && threadIdx.x == 0)
            write_result(r, cand, m, cf, wlo, whi, wnm, L, padL, padR, min_score, o_idx, o_score,
                         o_nbest, o_flags, o_umi);
    }
}


// ---- bit-parallel: 32 entries per thread in the bit lanes (nr_bitslice_core.h) ----------------
// The weighted form of Myers/Hyyro: every cell carries two 2-bit differences per lane, one cell of
// 32 entries is ten LOP3 (a DPX cell pair is two instructions for TWO entries).  A thread takes a
// word of 32 consecutive entries, transposes their packed columns in registers and leaves the
// four match masks of every column in shared memory ([base][column][thread]: conflict-free);
// a cell is then one LDS + the cell function.  The read's byte codes are block-uniform.  L = 16
// and L = 32 (slide-seq), with or without N columns; other core lengths keep the DPX kernel.
template <int L, bool HAS_N, int T>
__global__ void __launch_bounds__(T, 2)
nr_match_bitsliced_kernel(const uint32_t *__restrict__ wlo, const uint32_t *__restrict__ whi,
                          const uint32_t *__restrict__ wnm, uint32_t n, int padL, int padR,
                          const uint4 *__restrict__ bases, const uint8_t *__restrict__ meta,
                          const uint64_t *__restrict__ nmask, const uint32_t *__restrict__ list,
                          const uint32_t *__restrict__ list_count, uint64_t n_cand, int min_score,
                          int32_t *__restrict__ o_idx, int8_t *__restrict__ o_score,
                          uint8_t *__restrict__ o_nbest, uint8_t *__restrict__ o_flags,
                          uint8_t *__restrict__ o_umi, ExScratch sc)
{
    extern __shared__ uint32_t tabs[];          // eq [4][L][T], then (HAS_N) the N planes [L][T]
    __shared__ uint8_t cf[NR_MAX_QUERY], cr[NR_MAX_QUERY];
    __shared__ Best shb[32];
    __shared__ int sh_flag;
    const uint64_t total = list ? (uint64_t)*list_count : n_cand;
    const uint32_t nwords = (n + 31u) >> 5;
    const uint32_t S = ex_slices(total, sc);
    const uint32_t per = (nwords + S - 1) / S;
    uint32_t *eq = tabs + threadIdx.x;
    uint32_t *nmp = tabs + 4 * L * T + threadIdx.x;
    for (uint64_t w = blockIdx.x; w < total * S; w += gridDim.x) {
        const uint64_t it = w / S;
        const uint32_t g_lo = (uint32_t)(w % S) * per;
        const uint32_t g_hi = min(nwords, g_lo + per);
        const uint64_t cand = list ? (uint64_t)list[it] : it;
        const uint8_t mt = meta[cand];
        if (mt == 0xFF) {
            if (threadIdx.x == 0) {
                o_idx[cand] = -1; o_score[cand] = NR_SCORE_BELOW; o_nbest[cand] = 0;
                o_flags[cand] = NR_FLAG_TOO_LONG | NR_FLAG_BELOW | NR_FLAG_NO_UMI;
                o_umi[cand] = NR_UMI_NONE;
            }
            continue;
        }
        const int m = mt & 0x7F;
        __syncthreads();
        load_codes(bases, (mt & 0x80) ? nmask : nullptr, cand, m, cf, cr);
        __syncthreads();
        Best b; b.score = -1000; b.cnt = 0; b.key = 0xFFFFFFFFu;
        for (uint32_t g = g_lo + threadIdx.x; g < g_hi; g += T) {
            const uint32_t base = g << 5;
            const bool whole = base + 32u <= n;
            uint32_t r[32];
            if (HAS_N) {
#pragma unroll
                for (int k = 0; k < 32; k++) r[k] = wnm[min(base + k, n - 1)];
                nr_bs_transpose32(r);
#pragma unroll
                for (int j = 0; j < L; j++) nmp[j * T] = r[j];
            }
            if (whole) {
                const uint4 *src = reinterpret_cast<const uint4 *>(wlo + base);
#pragma unroll
                for (int k = 0; k < 8; k++) {
                    const uint4 v = __ldg(src + k);
                    r[4 * k] = v.x; r[4 * k + 1] = v.y; r[4 * k + 2] = v.z; r[4 * k + 3] = v.w;
                }
            } else {
#pragma unroll
                for (int k = 0; k < 32; k++) r[k] = wlo[min(base + k, n - 1)];
            }
            nr_bs_transpose32(r);
            nr_bs_build_eq16<L, HAS_N>(r, 0, eq, nmp, T);
            if (L == 32) {
#pragma unroll
                for (int k = 0; k < 32; k++) r[k] = whi[min(base + k, n - 1)];
                nr_bs_transpose32(r);
                nr_bs_build_eq16<L, HAS_N>(r, L == 32 ? 16 : 0, eq, nmp, T);
            }
            const uint32_t valid = whole ? 0xFFFFFFFFu : ((1u << (n - base)) - 1u);
#pragma unroll 1
            for (int strand = 0; strand < 2; strand++) {
                uint32_t M[NR_BS_PLANES];
                nr_bs_word_strand<L, HAS_N>(eq, nmp, T, strand ? cr : cf, m, padL, padR, M);
                int v;
                const uint32_t at = nr_bs_lane_min(M, valid, &v);
                best_merge(b, L - v, (uint32_t)__popc(at),
                           ((base + (uint32_t)(__ffs((int)at) - 1)) << 1) | (uint32_t)strand);
            }
        }
        Best r = block_reduce_best(b, shb);
        if (ex_merge(r, w, it, S, sc, &sh_flag) && threadIdx.x == 0)
            write_result(r, cand, m, cf, wlo, L == 32 ? whi : nullptr, HAS_N ? wnm : nullptr, L, padL,
                         padR, min_score, o_idx, o_score, o_nbest, o_flags, o_umi);
    }
}

template <int L, bool HAS_N, int T>
int launch_bitsliced(const nr_whitelist *wl, unsigned grid, const void *d_bases, const uint8_t *d_meta,
                     const uint64_t *d_nmask, const uint32_t *d_list, const uint32_t *d_list_count,
                     uint64_t n_cand, int min_score, int32_t *d_idx, int8_t *d_score, uint8_t *d_nbest,
                     uint8_t *d_flags, uint8_t *d_umi, const ExScratch &sc, cudaStream_t stream)
{
    const size_t smem = (size_t)(HAS_N ? 5 : 4) * L * T * sizeof(uint32_t);
    NR_CHECK_CUDA(cudaFuncSetAttribute(nr_match_bitsliced_kernel<L, HAS_N, T>,
                                       cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    nr_match_bitsliced_kernel<L, HAS_N, T><<<grid, T, smem, stream>>>(
        wl->d_lo, wl->d_hi, wl->d_nm, (uint32_t)wl->n, (int)wl->pad_l, (int)wl->pad_r,
        (const uint4 *)d_bases, d_meta, d_nmask, d_list, d_list_count, n_cand, min_score, d_idx,
        d_score, d_nbest, d_flags, d_umi, sc);
    NR_CHECK_CUDA(cudaGetLastError());
    return NR_OK;
}

}  // namespace

// host launcher shared with the C API: either all n_cand candidates, or the device-resident
// list (d_list, d_list_count) written by the filtered kernel.
int nr_launch_exhaustive(const nr_whitelist *wl, const void *d_bases, const uint8_t *d_meta,
                         const uint64_t *d_nmask, const uint32_t *d_list,
                         const uint32_t *d_list_count, uint64_t n_cand, int min_score,
                         int32_t *d_idx, int8_t *d_score, uint8_t *d_nbest, uint8_t *d_flags,
                         uint8_t *d_umi, int grid_cap, void *d_scratch, cudaStream_t stream)
{
    if (!d_list && n_cand == 0) return NR_OK;
    // list mode: a fixed grid (the count lives on the device); full mode: one block per
    // candidate, but never fewer blocks than grid_cap so that short batches get split
    uint64_t want = d_list ? (uint64_t)grid_cap : (n_cand > (uint64_t)grid_cap ? n_cand : (uint64_t)grid_cap);
    unsigned grid = (unsigned)(want < 65535ull * 8 ? want : 65535ull * 8);
    if (grid == 0) grid = 1;
    ExScratch sc;
    sc.done = d_scratch ? (uint32_t *)d_scratch : nullptr;
    sc.part = d_scratch ? (uint4 *)((uint8_t *)d_scratch + NR_EX_MAXGRID * sizeof(uint32_t)) : nullptr;
    static const bool force_generic = getenv("NR_FORCE_GENERIC_EXHAUSTIVE") != nullptr;   // experiments only
    // the DPX kernels of round 1 stay selectable for A/B timing (tools/time_exhaustive.py)
    static const bool force_dpx = getenv("NR_EXHAUSTIVE_DPX") != nullptr || force_generic;
    if (!force_dpx && (wl->L == 16 || wl->L == 32)) {
        const bool hn = wl->has_n && wl->d_nm;
        if (wl->L == 16)
            return hn ? launch_bitsliced<16, true, 256>(wl, grid, d_bases, d_meta, d_nmask, d_list, d_list_count,
                                                        n_cand, min_score, d_idx, d_score, d_nbest, d_flags,
                                                        d_umi, sc, stream)
                      : launch_bitsliced<16, false, 256>(wl, grid, d_bases, d_meta, d_nmask, d_list, d_list_count,
                                                         n_cand, min_score, d_idx, d_score, d_nbest, d_flags,
                                                         d_umi, sc, stream);
        return hn ? launch_bitsliced<32, true, 128>(wl, grid, d_bases, d_meta, d_nmask, d_list, d_list_count,
                                                    n_cand, min_score, d_idx, d_score, d_nbest, d_flags, d_umi,
                                                    sc, stream)
                  : launch_bitsliced<32, false, 128>(wl, grid, d_bases, d_meta, d_nmask, d_list, d_list_count,
                                                     n_cand, min_score, d_idx, d_score, d_nbest, d_flags, d_umi,
                                                     sc, stream);
    }
    if (wl->L == 16 && !wl->has_n && !force_generic) {
        nr_match_exhaustive16_kernel<<<grid, 256, 0, stream>>>(
            wl->d_lo, (uint32_t)wl->n, (int)wl->pad_l, (int)wl->pad_r, (const uint4 *)d_bases,
            d_meta, d_nmask, d_list, d_list_count, n_cand, min_score, d_idx, d_score, d_nbest,
            d_flags, d_umi, sc);
    } else {
        const size_t smem = (size_t)4 * wl->L * NR_EXG_THREADS * sizeof(uint32_t);
        if (wl->L == 32) {
            NR_CHECK_CUDA(cudaFuncSetAttribute(nr_match_exhaustive_generic_kernel<32>,
                                               cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            nr_match_exhaustive_generic_kernel<32><<<grid, NR_EXG_THREADS, smem, stream>>>(
                wl->d_lo, wl->d_hi, wl->d_nm, (uint32_t)wl->n, 32, (int)wl->pad_l,
                (int)wl->pad_r, (const uint4 *)d_bases, d_meta, d_nmask, d_list, d_list_count,
                n_cand, min_score, d_idx, d_score, d_nbest, d_flags, d_umi, sc);
        } else if (wl->L == 16) {          // 16 columns with N inside some entry
            NR_CHECK_CUDA(cudaFuncSetAttribute(nr_match_exhaustive_generic_kernel<16>,
                                               cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            nr_match_exhaustive_generic_kernel<16><<<grid, NR_EXG_THREADS, smem, stream>>>(
                wl->d_lo, wl->d_hi, wl->d_nm, (uint32_t)wl->n, 16, (int)wl->pad_l,
                (int)wl->pad_r, (const uint4 *)d_bases, d_meta, d_nmask, d_list, d_list_count,
                n_cand, min_score, d_idx, d_score, d_nbest, d_flags, d_umi, sc);
        } else {
            NR_CHECK_CUDA(cudaFuncSetAttribute(nr_match_exhaustive_generic_kernel<0>,
                                               cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            nr_match_exhaustive_generic_kernel<0><<<grid, NR_EXG_THREADS, smem, stream>>>(
                wl->d_lo, wl->d_hi, wl->d_nm, (uint32_t)wl->n, (int)wl->L, (int)wl->pad_l,
                (int)wl->pad_r, (const uint4 *)d_bases, d_meta, d_nmask, d_list, d_list_count,
                n_cand, min_score, d_idx, d_score, d_nbest, d_flags, d_umi, sc);
        }
    }
    NR_CHECK_CUDA(cudaGetLastError());
    return NR_OK;
}
