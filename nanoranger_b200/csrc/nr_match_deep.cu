// nr_match_deep.cu -- the deep matcher tier: exact best (entry, strand) of a candidate over the
// WHOLE whitelist for costs up to K, by meeting in the middle (nr_deep_core.h).
//
// Replaces scripts/barcode_align.sh:14-41 for the reads the seed filter cannot decide -- best
// score below core_len - 2, more than 32 co-optimal pairs, reads shorter than the filter takes,
// whitelists without a seed index (slide-seq cores) -- so that `_barcode_scores.csv`
// (utils.py:698, 728-730: every uniquely mapped forward read at every AS) is exact at any input
// size.  Per candidate the work is  (G_pre * s + G_suf * (L - s)) automaton column steps plus two
// byte loads per entry, instead of the n * L * m cells of the brute-force DP
// (nr_match_exhaustive.cu): 737K-august-2016 has 1 920 distinct 8-column prefixes and 1 536
// suffixes, i.e. 55 K column steps against 23.6 M DP columns.
//
// Mapping.  One block resolves one candidate at a time (dynamic hand-out), strand after strand:
//   rows     the strand's row masks (which read rows carry base c / N) from two ballots
//   phase A  thread = one prefix or suffix group: K+1 planes of 64 read rows in registers, one
//            automaton step per core column.  A0 runs the columns shared inside each half once
//            per "mid" group (distinct first s1 / last u1 columns), A1 continues every group from
//            its parent's planes.  Planes [e][group] and the group's minimum go to shared memory
//            (global scratch, L2, when the whitelist has too many groups)
//   phase B  two-sided: a pair (prefix group, suffix group) can only reach cost <= K if the
//            prefix minimum is <= K/2 or the suffix minimum is <= K - 1 - K/2, and groups that
//            close to the read are rare (a handful out of thousands).  Warps scan the entries of
//            those groups only -- prefix-sorted order for the hot prefix groups, suffix-sorted
//            order for the hot suffix groups (minus what the prefix side already saw) -- lane =
//            one entry: two byte minima, and only if they can still reach the running best the
//            plane join  OR_{a+b=t} F_a & B_b
//   merge    best cost / pairs attaining it / smallest pair, per thread, then over the block
// The winner's UMI column and flags come from the same scalar pair DP the exhaustive kernel uses.
// Candidates without any pair at cost <= K (or with more than 63 bases) go to the next tier's list.
#include <cstdlib>
#include <cstring>

#include "nr_common.cuh"
#include "nr_deep_core.h"
#include "nr_ex_common.cuh"

#define NR_DEEP_MAXTHREADS 1024
#define NR_DEEP_HCAP 2048          // groups per phase-B window (= capacity of the hot list)
#define NR_DEEP_ITEM 512           // entries per phase-B work item: four batches of NR_DEEP_BATCH
#define NR_DEEP_BATCH 128          // entries per batch (4 per lane, loads issued together)
#define NR_UMI_PENDING 254         // umi_q placeholder between the resolver and the finaliser

struct nr_deep_params {
    // grouping (nr_deep_index.h)
    const uint32_t *pre_start, *suf_start;
    const uint4 *pre_rep, *suf_rep, *pmid_rep, *smid_rep;
    const uint32_t *ent_suf, *ent_idx, *sent_pre, *sent_idx;
    uint32_t g_pre, g_suf, g_pmid, g_smid;
    int L, s, s1, u1, padL, padR;
    // whitelist cores (result writer)
    const uint32_t *lo, *hi, *nm;
    // candidates
    const uint4 *bases;
    const uint8_t *meta;
    const uint64_t *nmask;
    const uint32_t *list;          // nullable: all n_cand candidates
    const uint32_t *list_count;
    uint64_t n_cand;
    int min_score;
    int32_t *o_idx;
    int8_t *o_score;
    uint8_t *o_nbest, *o_flags, *o_umi;
    uint32_t *next_list;           // candidates this tier leaves
    uint32_t *next_count;
    unsigned long long *work_next; // next work item (zeroed with the workspace header)
    unsigned long long *resolved;  // nullable counter
    uint8_t *scratch;              // plane tables when they do not fit shared memory
    size_t scratch_per_block;
};

namespace {

using namespace nr_ex;

template <bool SMEM>
__device__ __forceinline__ uint64_t plane_ld(const uint64_t *p)
{
    if (SMEM) return *p;
    return (uint64_t)__ldcg(reinterpret_cast<const unsigned long long *>(p));
}
template <bool SMEM>
__device__ __forceinline__ void plane_st(uint64_t *p, uint64_t v)
{
    if (SMEM) *p = v;
    else __stcg(reinterpret_cast<unsigned long long *>(p), (unsigned long long)v);
}

// phase A for one strand; NTERM = false when neither the read nor the whitelist has N
template <int K, bool SMEM, bool NTERM>
__device__ __forceinline__ void phase_a(const nr_deep_params &P, const nr_deep_rows &rows, int m,
                                        uint64_t *planes, uint64_t *mid, uint8_t *mins,
                                        int *s_gmin /* [0] prefixes, [1] suffixes */)
{
    const uint32_t tid = threadIdx.x, lane = tid & 31u;
    const uint32_t G = P.g_pre + P.g_suf, GM = P.g_pmid + P.g_smid;
    // A0: shared columns, once per mid group
    for (uint32_t g = tid; g < GM; g += blockDim.x) {
        nr_deep_planes<K> x;
        if (g < P.g_pmid) {
            const uint4 rep = __ldg(P.pmid_rep + g);
            nr_deep_init_fwd<K>(x, m, P.padL);
#pragma unroll 1
            for (int j = 0; j < P.s1; j++)
                nr_deep_step_fwd<K, NTERM>(x, rows, nr_core_col(rep.x, rep.y, j), (rep.z >> j) & 1u);
        } else {
            const uint4 rep = __ldg(P.smid_rep + (g - P.g_pmid));
            nr_deep_init_bwd<K>(x, m, P.padR, rows.valid);
#pragma unroll 1
            for (int j = P.L - 1; j >= P.L - P.u1; j--)
                nr_deep_step_bwd<K, NTERM>(x, rows, nr_core_col(rep.x, rep.y, j), (rep.z >> j) & 1u);
        }
#pragma unroll
        for (int e = 0; e <= K; e++) plane_st<SMEM>(mid + (size_t)e * GM + g, x.v[e]);
    }
    if (GM) __syncthreads();
    // A1: the rest of each half, from the parent's planes
    int my_fmin = K + 1, my_smin = K + 1;
    for (uint32_t g = tid; g < G; g += blockDim.x) {
        nr_deep_planes<K> x;
        if (g < P.g_pre) {
            const uint4 rep = __ldg(P.pre_rep + g);
            if (P.s1 > 0) {
#pragma unroll
                for (int e = 0; e <= K; e++) x.v[e] = plane_ld<SMEM>(mid + (size_t)e * GM + rep.w);
            } else {
                nr_deep_init_fwd<K>(x, m, P.padL);
            }
#pragma unroll 1
            for (int j = P.s1; j < P.s; j++)
                nr_deep_step_fwd<K, NTERM>(x, rows, nr_core_col(rep.x, rep.y, j), (rep.z >> j) & 1u);
        } else {
            const uint4 rep = __ldg(P.suf_rep + (g - P.g_pre));
            if (P.u1 > 0) {
#pragma unroll
                for (int e = 0; e <= K; e++)
                    x.v[e] = plane_ld<SMEM>(mid + (size_t)e * GM + P.g_pmid + rep.w);
            } else {
                nr_deep_init_bwd<K>(x, m, P.padR, rows.valid);
            }
#pragma unroll 1
            for (int j = P.L - 1 - P.u1; j >= P.s; j--)
                nr_deep_step_bwd<K, NTERM>(x, rows, nr_core_col(rep.x, rep.y, j), (rep.z >> j) & 1u);
        }
        const int mn = nr_deep_min<K>(x);
#pragma unroll
        for (int e = 0; e <= K; e++) plane_st<SMEM>(planes + (size_t)e * G + g, x.v[e]);
        mins[g] = (uint8_t)mn;
        if (g < P.g_pre) my_fmin = min(my_fmin, mn); else my_smin = min(my_smin, mn);
    }
    my_fmin = __reduce_min_sync(0xffffffffu, my_fmin);
    my_smin = __reduce_min_sync(0xffffffffu, my_smin);
    if (lane == 0) {
        if (my_fmin <= K) atomicMin(&s_gmin[0], my_fmin);
        if (my_smin <= K) atomicMin(&s_gmin[1], my_smin);
    }
}

template <int K, bool SMEM, bool WLN>
__global__ void __launch_bounds__(NR_DEEP_MAXTHREADS, 1)
nr_match_deep_kernel(const nr_deep_params P)
{
    extern __shared__ __align__(16) uint8_t dyn[];
    __shared__ uint8_t cf[NR_MAX_QUERY], cr[NR_MAX_QUERY];
    __shared__ nr_deep_rows s_rows;
    __shared__ uint32_t s_bal[2][5];
    __shared__ int s_bound, s_gmin[2];
    __shared__ uint32_t s_nhot, s_item_next;
    __shared__ uint32_t s_hot[NR_DEEP_HCAP], s_hpre[NR_DEEP_HCAP + 1];
    __shared__ unsigned long long s_item;
    __shared__ int s_rc[32];
    __shared__ uint32_t s_rn[32], s_rk[32];

    constexpr int A = K / 2, B_ = K - 1 - K / 2;     // which side scans a pair (see header)
    const uint32_t tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
    const uint32_t G = P.g_pre + P.g_suf, GM = P.g_pmid + P.g_smid;
    uint64_t *planes, *mid;
    uint8_t *mins;
    if (SMEM) {
        planes = reinterpret_cast<uint64_t *>(dyn);
        mid = planes + (size_t)G * (K + 1);
        mins = reinterpret_cast<uint8_t *>(mid + (size_t)GM * (K + 1));
    } else {
        planes = reinterpret_cast<uint64_t *>(P.scratch + (size_t)blockIdx.x * P.scratch_per_block);
        mid = planes + (size_t)G * (K + 1);
        mins = dyn;
    }
    const uint64_t total = P.list ? (uint64_t)*P.list_count : P.n_cand;

    for (;;) {
        __syncthreads();
        if (tid == 0) s_item = atomicAdd(P.work_next, 1ull);
        __syncthreads();
        const uint64_t it = s_item;
        if (it >= total) break;
        const uint64_t cand = P.list ? (uint64_t)P.list[it] : it;
        const uint8_t mt = P.meta[cand];
        const int m = mt & 0x7F;
        if (mt == 0xFF || m < 1 || m > NR_DEEP_MAXM) {
            if (tid == 0) P.next_list[atomicAdd(P.next_count, 1u)] = (uint32_t)cand;
            continue;
        }
        load_codes(P.bases, (mt & 0x80) ? P.nmask : nullptr, cand, m, cf, cr);
        if (tid == 0) s_bound = K;
        int bc = K + 1;                 // this thread's best cost / pairs / smallest pair
        uint32_t bn = 0, bk = 0xFFFFFFFFu;

#pragma unroll 1
        for (int st = 0; st < 2; st++) {
            __syncthreads();            // codes visible; previous strand's phase B finished
            if (tid < 64) {
                const int code = (st ? cr : cf)[tid];
                const bool ok = (int)tid < m;
#pragma unroll
                for (int c = 0; c < 4; c++) {
                    const uint32_t b = __ballot_sync(0xffffffffu, ok && code == c);
                    if (lane == 0) s_bal[warp][c] = b;
                }
                const uint32_t bn_ = __ballot_sync(0xffffffffu, ok && code > 3);
                if (lane == 0) s_bal[warp][4] = bn_;
            }
            if (tid == 64) { s_gmin[0] = K + 1; s_gmin[1] = K + 1; }
            __syncthreads();
            if (tid < 5) {
                const uint64_t v = ((((uint64_t)s_bal[1][tid]) << 32) | (uint64_t)s_bal[0][tid]) << 1;
                if (tid < 4) s_rows.eq[tid] = v; else s_rows.nrow = v;
            }
            if (tid == 5) {
                s_rows.valid = (m >= 63) ? ~0ull : ((1ull << (m + 1)) - 1ull);
                s_rows.edge = 1ull | (1ull << m);
            }
            __syncthreads();

            if (!WLN && !(mt & 0x80)) phase_a<K, SMEM, false>(P, s_rows, m, planes, mid, mins, s_gmin);
            else phase_a<K, SMEM, true>(P, s_rows, m, planes, mid, mins, s_gmin);
            __syncthreads();

            // ---- phase B: the entries of the groups close to the read ---------------------------
            // Per window of NR_DEEP_HCAP groups: list the hot groups, cut their entries into items
            // of NR_DEEP_ITEM, hand the items to the warps (a hot group has hundreds of entries and
            // there are only a few dozen of them: one warp per group would leave most warps idle).
            const int gfmin = s_gmin[0], gsmin = s_gmin[1];
            for (uint32_t w0 = 0; w0 < G; w0 += NR_DEEP_HCAP) {
                if (tid == 0) { s_nhot = 0; s_item_next = 0; }
                __syncthreads();
                const uint32_t w1 = min(G, w0 + NR_DEEP_HCAP);
                for (uint32_t g = w0 + tid; g < w1; g += blockDim.x) {
                    const int mn = mins[g];
                    const bool hot = g < P.g_pre ? (mn <= A && mn + gsmin <= K)
                                                 : (mn <= B_ && mn + gfmin <= K);
                    if (hot) s_hot[atomicAdd(&s_nhot, 1u)] = g;
                }
                __syncthreads();
                const uint32_t nh = s_nhot;
                if (warp == 0) {
                    uint32_t running = 0;
                    for (uint32_t base = 0; base < nh; base += 32) {
                        const uint32_t i = base + lane;
                        uint32_t cnt = 0;
                        if (i < nh) {
                            const uint32_t g = s_hot[i];
                            const uint32_t *st_ = g < P.g_pre ? P.pre_start + g : P.suf_start + (g - P.g_pre);
                            cnt = (__ldg(st_ + 1) - __ldg(st_) + NR_DEEP_ITEM - 1) / NR_DEEP_ITEM;
                        }
                        uint32_t incl = cnt;
#pragma unroll
                        for (int o = 1; o < 32; o <<= 1) {
                            const uint32_t v = __shfl_up_sync(0xffffffffu, incl, o);
                            if ((int)lane >= o) incl += v;
                        }
                        if (i < nh) s_hpre[i] = running + incl - cnt;
                        running += __shfl_sync(0xffffffffu, incl, 31);
                    }
                    if (lane == 0) s_hpre[nh] = running;
                }
                __syncthreads();
                const uint32_t n_items = s_hpre[nh];
                for (;;) {
                    uint32_t item = 0;
                    if (lane == 0) item = atomicAdd(&s_item_next, 1u);
                    item = __shfl_sync(0xffffffffu, item, 0);
                    if (item >= n_items) break;
                    uint32_t lo_ = 0, hi_ = nh;                 // largest k with s_hpre[k] <= item
                    while (hi_ - lo_ > 1) {
                        const uint32_t md = (lo_ + hi_) >> 1;
                        if (s_hpre[md] <= item) lo_ = md; else hi_ = md;
                    }
                    const uint32_t gg = s_hot[lo_];
                    const uint32_t off = (item - s_hpre[lo_]) * NR_DEEP_ITEM + lane;
                    const int gmn = mins[gg];
                    uint64_t own[K + 1];
#pragma unroll
                    for (int e = 0; e <= K; e++) own[e] = plane_ld<SMEM>(planes + (size_t)e * G + gg);
                    if (gg < P.g_pre) {
                        if (gmn + gsmin > *(volatile int *)&s_bound) continue;
                        const uint32_t p1 = __ldg(P.pre_start + gg + 1);
                        const uint32_t first = __ldg(P.pre_start + gg) + off;          // this lane's first entry
                        const uint32_t last = min(p1, first - lane + NR_DEEP_ITEM);     // end of the item
#pragma unroll 1
                        for (uint32_t p0 = first; p0 < last; p0 += NR_DEEP_BATCH) {
                            uint32_t h[NR_DEEP_BATCH / 32];
#pragma unroll
                            for (int u = 0; u < NR_DEEP_BATCH / 32; u++)
                                h[u] = p0 + 32 * u < p1 ? __ldg(P.ent_suf + p0 + 32 * u) : 0xFFFFFFFFu;
#pragma unroll
                            for (int u = 0; u < NR_DEEP_BATCH / 32; u++) {
                                if (h[u] == 0xFFFFFFFFu) continue;
                                const int bound = *(volatile int *)&s_bound;
                                if (gmn + (int)mins[P.g_pre + h[u]] > bound) continue;
                                uint64_t b[K + 1];
#pragma unroll
                                for (int e = 0; e <= K; e++)
                                    b[e] = plane_ld<SMEM>(planes + (size_t)e * G + P.g_pre + h[u]);
                                const int t = nr_deep_join<K>(own, b);
                                if (t <= bound) {
                                    const uint32_t key = (__ldg(P.ent_idx + p0 + 32 * u) << 1) | (uint32_t)st;
                                    if (t < bc) { bc = t; bn = 1; bk = key; atomicMin(&s_bound, t); }
                                    else if (t == bc) { bn++; bk = min(bk, key); }
                                }
                            }
                        }
                    } else {
                        const uint32_t hh = gg - P.g_pre;
                        if (gmn + gfmin > *(volatile int *)&s_bound) continue;
                        const uint32_t p1 = __ldg(P.suf_start + hh + 1);
                        const uint32_t first = __ldg(P.suf_start + hh) + off;
                        const uint32_t last = min(p1, first - lane + NR_DEEP_ITEM);
#pragma unroll 1
                        for (uint32_t p0 = first; p0 < last; p0 += NR_DEEP_BATCH) {
                            uint32_t gq[NR_DEEP_BATCH / 32];
#pragma unroll
                            for (int u = 0; u < NR_DEEP_BATCH / 32; u++)
                                gq[u] = p0 + 32 * u < p1 ? __ldg(P.sent_pre + p0 + 32 * u) : 0xFFFFFFFFu;
#pragma unroll
                            for (int u = 0; u < NR_DEEP_BATCH / 32; u++) {
                                if (gq[u] == 0xFFFFFFFFu) continue;
                                const int fm = mins[gq[u]];
                                const int bound = *(volatile int *)&s_bound;
                                if (fm <= A || fm + gmn > bound) continue;   // fm <= A: the prefix side has it
                                uint64_t f[K + 1];
#pragma unroll
                                for (int e = 0; e <= K; e++) f[e] = plane_ld<SMEM>(planes + (size_t)e * G + gq[u]);
                                const int t = nr_deep_join<K>(f, own);
                                if (t <= bound) {
                                    const uint32_t key = (__ldg(P.sent_idx + p0 + 32 * u) << 1) | (uint32_t)st;
                                    if (t < bc) { bc = t; bn = 1; bk = key; atomicMin(&s_bound, t); }
                                    else if (t == bc) { bn++; bk = min(bk, key); }
                                }
                            }
                        }
                    }
                }
                __syncthreads();        // every warp is done with this window's list and counters
            }
        }

        // ---- merge over the block --------------------------------------------------------------
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const int c2 = __shfl_xor_sync(0xffffffffu, bc, o);
            const uint32_t n2 = __shfl_xor_sync(0xffffffffu, bn, o);
            const uint32_t k2 = __shfl_xor_sync(0xffffffffu, bk, o);
            if (c2 < bc) { bc = c2; bn = n2; bk = k2; }
            else if (c2 == bc) { bn += n2; bk = min(bk, k2); }
        }
        if (lane == 0) { s_rc[warp] = bc; s_rn[warp] = bn; s_rk[warp] = bk; }
        __syncthreads();
        if (tid == 0) {
            int c = K + 1;
            uint32_t n = 0, k = 0xFFFFFFFFu;
            for (uint32_t w = 0; w < (blockDim.x >> 5); w++) {
                if (s_rc[w] < c) { c = s_rc[w]; n = s_rn[w]; k = s_rk[w]; }
                else if (s_rc[w] == c) { n += s_rn[w]; k = min(k, s_rk[w]); }
            }
            if (c <= K) {
                // the UMI column of a forward-strand winner is left to nr_deep_finalize_kernel (one
                // thread per candidate): a scalar pair DP here would hold the whole block
                const int score = P.L - c;
                const int strand = (int)(k & 1u);
                uint8_t fl = NR_FLAG_EXHAUSTIVE;
                if (n > 1) fl |= NR_FLAG_TIE;
                if (strand) fl |= NR_FLAG_RC | NR_FLAG_NO_UMI;
                if (score < P.min_score) fl |= NR_FLAG_BELOW;
                P.o_idx[cand] = (int32_t)(k >> 1);
                P.o_score[cand] = (int8_t)score;
                P.o_nbest[cand] = (uint8_t)min(n, 255u);
                P.o_flags[cand] = fl;
                P.o_umi[cand] = strand ? NR_UMI_NONE : NR_UMI_PENDING;
                if (P.resolved) atomicAdd(P.resolved, 1ull);
            } else {
                P.next_list[atomicAdd(P.next_count, 1u)] = (uint32_t)cand;
            }
        }
    }
}

// UMI columns of the candidates the deep tier resolved on the forward strand: thread = candidate,
// the plane automaton over the winner's L columns (nr_deep_umi_row), K = 5 planes cover every cost
// the tier reports.
__global__ void __launch_bounds__(256)
nr_deep_finalize_kernel(const uint4 *__restrict__ bases, const uint8_t *__restrict__ meta,
                        const uint64_t *__restrict__ nmask, const uint32_t *__restrict__ list,
                        const uint32_t *__restrict__ list_count, uint64_t n_cand,
                        const uint32_t *__restrict__ lo, const uint32_t *__restrict__ hi,
                        const uint32_t *__restrict__ nm, int L, int padL, int padR,
                        const int32_t *__restrict__ o_idx, const int8_t *__restrict__ o_score,
                        uint8_t *__restrict__ o_flags, uint8_t *__restrict__ o_umi)
{
    const uint64_t total = list ? (uint64_t)*list_count : n_cand;
    for (uint64_t it = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; it < total;
         it += (uint64_t)gridDim.x * blockDim.x) {
        const uint64_t cand = list ? (uint64_t)list[it] : it;
        if (o_umi[cand] != NR_UMI_PENDING) continue;
        const uint8_t mt = meta[cand];
        const int m = mt & 0x7F;
        const uint4 b4 = __ldg(bases + cand);
        const uint32_t w[4] = {b4.x, b4.y, b4.z, b4.w};
        const uint64_t len_mask = (1ull << m) - 1ull;                 // m <= 63 here
        const uint64_t nmk = (mt & 0x80) ? (nmask[cand] & len_mask) : 0ull;
        nr_deep_rows rows;
#pragma unroll
        for (uint32_t c = 0; c < 4; c++) {
            uint64_t eq = 0;
#pragma unroll
            for (int k = 0; k < 4; k++) {
                const uint32_t x = ~(w[k] ^ (c * 0x55555555u));
                uint32_t t = x & (x >> 1) & 0x55555555u;              // bit 2i: base i equals c
                t = (t | (t >> 1)) & 0x33333333u;
                t = (t | (t >> 2)) & 0x0F0F0F0Fu;
                t = (t | (t >> 4)) & 0x00FF00FFu;
                t = (t | (t >> 8)) & 0x0000FFFFu;
                eq |= (uint64_t)t << (16 * k);
            }
            rows.eq[c] = (eq & len_mask & ~nmk) << 1;
        }
        rows.nrow = nmk << 1;
        rows.valid = (m >= 63) ? ~0ull : ((1ull << (m + 1)) - 1ull);
        rows.edge = 1ull;                                             // no end-overhang rule (T0)
        const int32_t idx = o_idx[cand];
        const int u = nr_deep_umi_row<5>(rows, lo[idx], hi ? hi[idx] : 0u, nm ? nm[idx] : 0u, L, m,
                                         padL, padR, L - (int)o_score[cand]);
        if (u < 0) { o_flags[cand] |= NR_FLAG_NO_UMI; o_umi[cand] = NR_UMI_NONE; }
        else o_umi[cand] = (uint8_t)u;
    }
}

size_t deep_planes_bytes(const nr_whitelist *wl, int K)
{
    return (size_t)(wl->deep_gpre + wl->deep_gsuf + wl->deep_gpmid + wl->deep_gsmid) *
           (size_t)(K + 1) * sizeof(uint64_t);
}

constexpr size_t DEEP_SMEM_MAX = 227 * 1024 - 18 * 1024;    // dynamic budget next to ~17 KB static

// threads per block: every thread gets the same number of phase-A groups (to within one warp).
// `cap` = 512 when two blocks fit an SM (they then hide each other's barriers and serial parts).
unsigned deep_threads(size_t G, size_t cap)
{
    const size_t rounds = (G + cap - 1) / cap;
    size_t t = ((G + rounds - 1) / rounds + 31) & ~(size_t)31;
    if (t < 256) t = 256;
    if (t > cap) t = cap;
    return (unsigned)t;
}

// small: the caller expects few candidates (the leftovers of NR_MODE_FILTERED).  The tier then runs
// with its plane tables in global scratch and 256 threads per block, i.e. with the shared-memory
// footprint of the filtered kernel: a launch that asks for > 100 KB of shared memory makes every
// SM it lands on switch its L1/shared split, which serialises it against the filtered kernels of
// the other chunks in flight (nr_match_host pipelines chunks over several streams).
template <int K, bool WLN>
int launch_kw(const nr_whitelist *wl, nr_deep_params &P, uint8_t *d_scratch, size_t scratch_bytes,
              bool small, cudaStream_t stream)
{
    int sms = 148;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, wl->device);
    const size_t G = (size_t)wl->deep_gpre + wl->deep_gsuf;
    const size_t pb = deep_planes_bytes(wl, K);
    const size_t per = (pb + 255) & ~(size_t)255;
    const bool fits = pb + G <= DEEP_SMEM_MAX;
    const bool have_scratch = d_scratch && scratch_bytes >= per;
    if (fits && !(small && have_scratch)) {
        const size_t smem = pb + G;
        // two resident blocks per SM when their tables fit side by side (static ~17 KB each)
        const bool two = 2 * (smem + 18 * 1024) <= 227 * 1024;
        NR_CHECK_CUDA(cudaFuncSetAttribute(nr_match_deep_kernel<K, true, WLN>,
                                           cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        nr_match_deep_kernel<K, true, WLN><<<two ? 2 * sms : sms, deep_threads(G, two ? 512 : NR_DEEP_MAXTHREADS),
                                             smem, stream>>>(P);
    } else {
        size_t blocks = have_scratch ? scratch_bytes / per : 0;
        if (blocks > (size_t)sms) blocks = (size_t)sms;
        if (blocks == 0 || G > DEEP_SMEM_MAX) {
            nr_set_error("deep tier: no scratch for %zu groups", G);
            return NR_EINVAL;
        }
        P.scratch = d_scratch;
        P.scratch_per_block = per;
        NR_CHECK_CUDA(cudaFuncSetAttribute(nr_match_deep_kernel<K, false, WLN>,
                                           cudaFuncAttributeMaxDynamicSharedMemorySize, (int)G));
        nr_match_deep_kernel<K, false, WLN><<<(unsigned)blocks, fits ? 256u : deep_threads(G, NR_DEEP_MAXTHREADS),
                                              G, stream>>>(P);
    }
    NR_CHECK_CUDA(cudaGetLastError());
    return NR_OK;
}

template <int K>
int launch_k(const nr_whitelist *wl, nr_deep_params &P, uint8_t *d_scratch, size_t scratch_bytes,
             bool small, cudaStream_t stream)
{
    return wl->has_n ? launch_kw<K, true>(wl, P, d_scratch, scratch_bytes, small, stream)
                     : launch_kw<K, false>(wl, P, d_scratch, scratch_bytes, small, stream);
}

}  // namespace

// global scratch one launch of the deep tier may use: the plane tables of one block per SM (the
// only home of the tables when they do not fit shared memory; the small-footprint variant
// otherwise)
size_t nr_deep_scratch_bytes(const nr_whitelist *wl, int K)
{
    if (!wl->has_deep) return 0;
    const size_t pb = deep_planes_bytes(wl, K);
    int sms = 148;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, wl->device);
    return (size_t)sms * ((pb + 255) & ~(size_t)255);
}

// Whether the tier runs on this whitelist: the group minima must fit shared memory, and meeting in
// the middle must share enough to beat the bit-parallel brute-force kernel it stands in front of.
// One pass steps the plane automaton through `cols` group columns per strand (~60 instructions
// each) where the brute force runs n * L entry columns (~18 each, 32 entries per instruction):
// 737K-august-2016 has 11 K group columns against 11.8 M entry columns (and nearly every read a
// pair at cost <= 3), a 17 K-entry slide-seq list 2e5 against 5.7e5 -- there a pass costs as much
// as the brute force and most reads below the threshold fall through both passes anyway.
// NR_DEEP_TIER=always / never overrides the rule (tests, A/B timing).
int nr_deep_usable(const nr_whitelist *wl)
{
    const char *e = getenv("NR_DEEP_TIER");
    if (e && !strcmp(e, "never")) return 0;
    if (!wl->has_deep || ((size_t)wl->deep_gpre + wl->deep_gsuf) > DEEP_SMEM_MAX) return 0;
    if (e && !strcmp(e, "always")) return 1;
    const int L = (int)wl->L, s = wl->deep_s, s1 = wl->deep_s1, u1 = wl->deep_u1;
    const uint64_t cols = (uint64_t)wl->deep_gpmid * (uint64_t)s1 + (uint64_t)wl->deep_gpre * (uint64_t)(s - s1) +
                          (uint64_t)wl->deep_gsmid * (uint64_t)u1 +
                          (uint64_t)wl->deep_gsuf * (uint64_t)(L - s - u1);
    return cols * 8 <= wl->n * (uint64_t)L;
}

// Enqueue the deep tier (K = 3 or 5) on `stream`: candidates of (d_list, d_list_count) -- or all
// n_cand when d_list is null -- are resolved exactly when their best cost is <= K; the others are
// appended to (d_next_list, d_next_count).  *d_work_next must be zero.
int nr_launch_deep(const nr_whitelist *wl, int K, const void *d_bases, const uint8_t *d_meta,
                   const uint64_t *d_nmask, const uint32_t *d_list, const uint32_t *d_list_count,
                   uint64_t n_cand, int min_score, int32_t *d_idx, int8_t *d_score,
                   uint8_t *d_nbest, uint8_t *d_flags, uint8_t *d_umi, uint32_t *d_next_list,
                   uint32_t *d_next_count, unsigned long long *d_work_next,
                   unsigned long long *d_resolved, uint8_t *d_scratch, size_t scratch_bytes,
                   int small, cudaStream_t stream)
{
    if (!nr_deep_usable(wl)) { nr_set_error("deep tier not available for this whitelist"); return NR_EUNSUPPORTED; }
    if (!d_list && n_cand == 0) return NR_OK;
    nr_deep_params P;
    P.pre_start = wl->d_deep_pre_start; P.pre_rep = wl->d_deep_pre_rep; P.suf_rep = wl->d_deep_suf_rep;
    P.ent_suf = wl->d_deep_ent_suf; P.ent_idx = wl->d_deep_ent_idx;
    P.suf_start = wl->d_deep_suf_start; P.sent_pre = wl->d_deep_sent_pre; P.sent_idx = wl->d_deep_sent_idx;
    P.pmid_rep = wl->d_deep_pmid_rep; P.smid_rep = wl->d_deep_smid_rep;
    P.g_pre = wl->deep_gpre; P.g_suf = wl->deep_gsuf; P.g_pmid = wl->deep_gpmid; P.g_smid = wl->deep_gsmid;
    P.L = (int)wl->L; P.s = wl->deep_s; P.s1 = wl->deep_s1; P.u1 = wl->deep_u1;
    P.padL = (int)wl->pad_l; P.padR = (int)wl->pad_r;
    P.lo = wl->d_lo; P.hi = wl->d_hi; P.nm = wl->d_nm;
    P.bases = (const uint4 *)d_bases; P.meta = d_meta; P.nmask = d_nmask;
    P.list = d_list; P.list_count = d_list_count; P.n_cand = n_cand; P.min_score = min_score;
    P.o_idx = d_idx; P.o_score = d_score; P.o_nbest = d_nbest; P.o_flags = d_flags; P.o_umi = d_umi;
    P.next_list = d_next_list; P.next_count = d_next_count; P.work_next = d_work_next;
    P.resolved = d_resolved; P.scratch = nullptr; P.scratch_per_block = 0;
    if (K == 3) return launch_k<3>(wl, P, d_scratch, scratch_bytes, small != 0, stream);
    if (K == 5) return launch_k<5>(wl, P, d_scratch, scratch_bytes, small != 0, stream);
    nr_set_error("deep tier: K must be 3 or 5");
    return NR_EINVAL;
}

// UMI columns for everything the deep tier resolved among (d_list, d_list_count) -- or among all
// n_cand candidates when d_list is null.  Runs after the last deep launch of a call.
int nr_launch_deep_finalize(const nr_whitelist *wl, const void *d_bases, const uint8_t *d_meta,
                            const uint64_t *d_nmask, const uint32_t *d_list,
                            const uint32_t *d_list_count, uint64_t n_cand, const int32_t *d_idx,
                            const int8_t *d_score, uint8_t *d_flags, uint8_t *d_umi,
                            cudaStream_t stream)
{
    if (!d_list && n_cand == 0) return NR_OK;
    int sms = 148;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, wl->device);
    const uint64_t want = d_list ? (uint64_t)sms * 8 : (n_cand + 255) / 256;
    const unsigned grid = (unsigned)(want < (uint64_t)sms * 8 ? (want ? want : 1) : (uint64_t)sms * 8);
    nr_deep_finalize_kernel<<<grid, 256, 0, stream>>>(
        (const uint4 *)d_bases, d_meta, d_nmask, d_list, d_list_count, n_cand, wl->d_lo, wl->d_hi,
        wl->d_nm, (int)wl->L, (int)wl->pad_l, (int)wl->pad_r, d_idx, d_score, d_flags, d_umi);
    NR_CHECK_CUDA(cudaGetLastError());
    return NR_OK;
}
