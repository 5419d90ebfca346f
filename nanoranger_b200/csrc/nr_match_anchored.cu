// nr_match_anchored.cu -- lossless seed filter + exact verification for cores with a constant
// middle: the slide-seq geometry, 8 barcode columns + 18-column linker + 6 barcode columns
// (utils.py:584-601 of the reference; threshold AS >= 30, utils.py:638, i.e. cost <= 2).
//
// Replaces scripts/barcode_align.sh:14-41 for that whitelist in the score range the reference
// keeps.  The arithmetic (linker walk, the scripts that enumerate every 8-mer that can stand in
// front of the linker at cost <= 2, the exact scorer) is in nr_anchor_core.h and is the same code
// the CPU emulation in tests/emul/anchor_emul.cpp checks against the oracle.
//
// Mapping.  One warp owns a tile of 32 consecutive candidates (one coalesced 512 B request brings
// their records) and resolves them one after another:
//   junctions  lane = (strand, read row a): diagonal walk of the linker from row a; the rows where
//              the linker fits at cost <= 2 (usually one) are compacted into a shared-memory list
//   stages     t = 0, 1, 2: for every junction and every number d of extra bases at the P|K
//              junction, lane = one script of cost t - linker cost - d: the script's 8-mer from the
//              read window, one read of the direct-address P table (256 KB, L1/L2 resident), its
//              rows into the warp's queue; a candidate whose best pair already costs <= t stops
//   verify     lane = one queued row: the entry's 32 columns against the strand's row masks with
//              the plane automaton (K = 2, nr_deep_core.h), exact
//   merge      best cost, the distinct (entry, strand) pairs attaining it, smallest entry
// The winner's UMI column comes from the same automaton without the end-overhang rule
// (nr_deep_umi_row).  Reads with one or two N: wildcard linker walk, script windows with the N
// substituted by every base, N-aware scorer (nr_anchor_core.h).  Candidates the filter cannot take
// (more than two N, more than 63 bases, more junctions or co-optimal pairs than the warp's lists
// hold) go to the device list the deep tier resolves.
#include "nr_common.cuh"
#include "nr_anchor_core.h"
#include "nr_filter_core.h"

#define NR_AWARPS 8
#define NR_AQCAP 256            // queued rows per warp
#define NR_AJCAP 32             // junctions per candidate

struct nr_anchor_params {
    const uint32_t *pstart;     // 4^8 + 1: rows of every 8-mer
    const uint32_t *prows;      // entry indices grouped by P key
    const uint32_t *lo, *hi, *nm;
    uint64_t link;
    int L, Lk, padL, padR;
    const uint4 *bases;
    const uint8_t *meta;
    const uint64_t *nmask;
    uint64_t n_cand;
    int min_score;
    int resolve_below;
    int32_t *o_idx;
    int8_t *o_score;
    uint8_t *o_nbest, *o_flags, *o_umi;
    uint32_t *list;
    uint32_t *list_count;
    unsigned long long *tile_next;
    unsigned long long *counters;   // nullable: keys, rows, verifications, passes, listed
};

namespace {

struct WarpSmemA {
    uint4 tile[32];
    uint32_t rdp[2][NR_RDP_WORDS];
    nr_deep_rows rows[2];
    uint64_t nm[2];                 // N rows of the candidate in flight, per strand (bit i = base i)
    uint32_t junc[NR_AJCAP];        // strand | a << 1 | linker cost << 8
    uint32_t queue[NR_AQCAP];       // row of the P table | strand << 31
};

struct AccA {
    int best, nb;
    uint32_t key;                   // per lane: a pair at `best` (lane < nb)
    int overflow, qn;
};

__device__ __forceinline__ uint32_t script_pack(const nr_anchor_script &s)
{
    return (uint32_t)s.kind | ((uint32_t)s.cost << 3) | ((uint32_t)s.j << 5) | ((uint32_t)s.x << 9) |
           ((uint32_t)s.r1 << 13) | ((uint32_t)s.r2 << 17);
}
__device__ __forceinline__ nr_anchor_script script_unpack(uint32_t w)
{
    nr_anchor_script s;
    s.kind = (uint8_t)(w & 7u); s.cost = (uint8_t)((w >> 3) & 3u); s.j = (uint8_t)((w >> 5) & 15u);
    s.x = (uint8_t)((w >> 9) & 15u); s.r1 = (uint8_t)((w >> 13) & 15u); s.r2 = (uint8_t)((w >> 17) & 15u);
    return s;
}

// row masks of one strand from its packed words (bit i <-> read row i); nm: bases that are N
__device__ __forceinline__ void rows_from_words(const uint32_t *w, int m, uint64_t nm, nr_deep_rows &r)
{
    const uint64_t len_mask = ((1ull << m) - 1ull) & ~nm;       // m <= 63
#pragma unroll
    for (uint32_t c = 0; c < 4; c++) {
        uint64_t eq = 0;
#pragma unroll
        for (int k = 0; k < 4; k++) {
            const uint32_t x = ~(w[k] ^ (c * 0x55555555u));
            uint32_t t = x & (x >> 1) & 0x55555555u;
            t = (t | (t >> 1)) & 0x33333333u;
            t = (t | (t >> 2)) & 0x0F0F0F0Fu;
            t = (t | (t >> 4)) & 0x00FF00FFu;
            t = (t | (t >> 8)) & 0x0000FFFFu;
            eq |= (uint64_t)t << (16 * k);
        }
        r.eq[c] = (eq & len_mask) << 1;
    }
    r.nrow = nm << 1;
    r.valid = (m >= 63) ? ~0ull : ((1ull << (m + 1)) - 1ull);
    r.edge = 1ull | (1ull << m);
}

template <bool COUNT>
__device__ __forceinline__ void drain_a(const nr_anchor_params &P, WarpSmemA &sm, AccA &acc, int m,
                                        unsigned long long &c_ver, unsigned long long &c_pass)
{
    const uint32_t lane = nr_lane();
    while (acc.qn > 0) {
        const int cnt = min(32, acc.qn);
        const int base = acc.qn - cnt;
        acc.qn = base;
        int cost = 3;
        uint32_t k = 0;
        if ((int)lane < cnt) {
            const uint32_t it = sm.queue[base + lane];
            const int strand = (int)(it >> 31);
            const uint32_t en = __ldg(P.prows + (it & 0x7FFFFFFFu));
            cost = nr_anchor_score(sm.rows[strand], __ldg(P.lo + en), P.hi ? __ldg(P.hi + en) : 0u,
                                   P.nm ? __ldg(P.nm + en) : 0u, P.L, m, P.padL, P.padR);
            k = (en << 1) | (uint32_t)strand;
            if (COUNT) { c_ver++; c_pass += cost < 3; }
        }
        __syncwarp();
        const int rb = __reduce_min_sync(0xffffffffu, cost);
        if (rb < 3 && rb <= acc.best) {
            if (rb < acc.best) { acc.best = rb; acc.nb = 0; }
            uint32_t contrib = __ballot_sync(0xffffffffu, cost == rb);
            while (contrib) {
                const int src = __ffs(contrib) - 1;
                contrib &= contrib - 1;
                const uint32_t kk = __shfl_sync(0xffffffffu, k, src);
                const uint32_t found = __ballot_sync(0xffffffffu, (int)lane < acc.nb && acc.key == kk);
                if (!found) {
                    if (acc.nb < 32) {
                        if ((int)lane == acc.nb) acc.key = kk;
                        acc.nb++;
                    } else {
                        acc.overflow = 1;
                    }
                }
            }
        }
    }
}

template <bool COUNT>
__global__ void __launch_bounds__(NR_AWARPS * 32, 3)
nr_match_anchored_kernel(const nr_anchor_params P)
{
    __shared__ WarpSmemA smem[NR_AWARPS];
    __shared__ uint32_t s_scripts[NR_ANCHOR_NSCRIPTS];
    __shared__ int s_first[4];
    if (threadIdx.x == 0) {
        nr_anchor_table t;
        nr_anchor_build_table(t);
        for (int i = 0; i < t.n; i++) s_scripts[i] = script_pack(t.s[i]);
        for (int i = 0; i < 4; i++) s_first[i] = t.first[i];
    }
    __syncthreads();
    const uint32_t lane = nr_lane();
    WarpSmemA &sm = smem[threadIdx.x >> 5];
    const uint64_t n_tiles = (P.n_cand + 31) >> 5;
    const int Lk = P.Lk;
    unsigned long long c_keys = 0, c_rows = 0, c_ver = 0, c_pass = 0, c_listed = 0;

    for (;;) {
        unsigned long long t64 = 0;
        if (lane == 0) t64 = atomicAdd(P.tile_next, 1ull);
        const uint64_t tile = ((uint64_t)__shfl_sync(0xffffffffu, (uint32_t)(t64 >> 32), 0) << 32) |
                              (uint64_t)__shfl_sync(0xffffffffu, (uint32_t)t64, 0);
        if (tile >= n_tiles) break;
        uint32_t mt = 0x100u, nm_lo = 0, nm_hi = 0;
        {
            const uint64_t mine = tile * 32 + lane;
            uint4 b = make_uint4(0u, 0u, 0u, 0u);
            if (mine < P.n_cand) {
                b = __ldcs(P.bases + mine); mt = P.meta[mine];
                if (mt != 0xFFu && (mt & 0x80u)) {
                    const uint64_t nm = P.nmask[mine];
                    nm_lo = (uint32_t)nm; nm_hi = (uint32_t)(nm >> 32);
                }
            }
            __syncwarp();
            sm.tile[lane] = b;
            __syncwarp();
        }
        const int in_tile = (int)min((uint64_t)32, P.n_cand - tile * 32);
#pragma unroll 1
        for (int c = 0; c < in_tile; c++) {
            const uint32_t cmt = __shfl_sync(0xffffffffu, mt, c);
            const uint64_t cand = tile * 32 + c;
            if (cmt == 0xFFu) {
                if (lane == 0) {
                    P.o_idx[cand] = -1; P.o_score[cand] = NR_SCORE_BELOW; P.o_nbest[cand] = 0;
                    P.o_flags[cand] = NR_FLAG_TOO_LONG | NR_FLAG_BELOW | NR_FLAG_NO_UMI;
                    P.o_umi[cand] = NR_UMI_NONE;
                }
                continue;
            }
            const int m = (int)(cmt & 0x7Fu);
            uint64_t nm = ((uint64_t)__shfl_sync(0xffffffffu, nm_hi, c) << 32) |
                          (uint64_t)__shfl_sync(0xffffffffu, nm_lo, c);
            if (m >= 1 && m <= NR_DEEP_MAXM) nm &= (1ull << m) - 1ull;
            bool to_list = __popcll(nm) > 2 || m > NR_DEEP_MAXM || m < 1;
            AccA acc;
            acc.best = 3; acc.nb = 0; acc.key = 0; acc.overflow = 0; acc.qn = 0;
            if (!to_list) {
                // both strands, padded, and their row masks
                {
                    const uint4 t4 = sm.tile[c];
                    const uint32_t w4[4] = {t4.x, t4.y, t4.z, t4.w};
                    uint32_t rc[4];
                    nr_revcomp4(w4, m, rc);
                    uint32_t vf = 0, vr = 0;
#pragma unroll
                    for (int k = 0; k < 4; k++)
                        if ((int)lane == k + 1) { vf = w4[k]; vr = rc[k]; }
                    __syncwarp();
                    if (lane < NR_RDP_WORDS) { sm.rdp[0][lane] = vf; sm.rdp[1][lane] = vr; }
                    const uint64_t nmr = nm ? (__brevll(nm) >> (64 - m)) : 0ull;
                    if (lane == 8) { rows_from_words(w4, m, nm, sm.rows[0]); sm.nm[0] = nm; }
                    if (lane == 9) { rows_from_words(rc, m, nmr, sm.rows[1]); sm.nm[1] = nmr; }
                    __syncwarp();
                }
                // junctions: rows a where the linker fits at cost <= 2 (a >= 6: P may hang over the
                // read start by two columns at most)
                int nj = 0;
                const int a0 = NR_ANCHOR_MAXP - 2, nA = m - Lk + 1 - a0 + 1;
                const int nslots = nA > 0 ? 2 * nA : 0;
                for (int base = 0; base < nslots; base += 32) {
                    const int slot = base + (int)lane;
                    const bool ok = slot < nslots;
                    const int strand = slot >= nA ? 1 : 0;
                    const int a = a0 + slot - strand * nA;
                    int ck = 3;
                    if (ok) {
                        const uint64_t V = nr_window64(sm.rdp[strand], a - 1);
                        // N rows count as matches in the walk (a lower bound on the linker cost)
                        const uint64_t vm = nr_valid_mask(1 - a, m - a + 1) &
                                            ~nr_spread_even((uint32_t)(sm.nm[strand] >> (a - 1)));
                        const int fl = nr_anchor_linker(V, vm, P.link, Lk);
                        if (fl) ck = (fl & 1) ? 0 : ((fl & 2) ? 1 : 2);
                    }
                    const uint32_t bal = __ballot_sync(0xffffffffu, ck < 3);
                    if (bal) {
                        const int pos = nj + __popc(bal & ((1u << lane) - 1u));
                        if (ck < 3 && pos < NR_AJCAP)
                            sm.junc[pos] = (uint32_t)strand | ((uint32_t)a << 1) | ((uint32_t)ck << 8);
                        nj += __popc(bal);
                    }
                }
                __syncwarp();
                if (nj > NR_AJCAP) {
                    to_list = true;                  // low-complexity read: leave it to the deep tier
                } else {
#pragma unroll 1
                    for (int t = 0; t <= 2; t++) {
                        if (acc.best < t) break;
#pragma unroll 1
                        for (int ji = 0; ji < nj; ji++) {
                            const uint32_t jw = sm.junc[ji];
                            const int strand = (int)(jw & 1u), a = (int)((jw >> 1) & 127u), ck = (int)(jw >> 8);
#pragma unroll 1
                            for (int d = 0; ck + d <= t; d++) {
                                const int cs = t - ck - d, e = a - d;
                                const uint32_t W0 = (uint32_t)nr_window64(sm.rdp[strand], e - 10) & 0xFFFFFu;
                                // N rows inside the window: every base they can stand for
                                const uint64_t nms = sm.nm[strand];
                                const uint32_t nw = (uint32_t)(e >= 10 ? nms >> (e - 10) : nms << (10 - e)) & 0x3FFu;
                                const int nn = __popc(nw);
                                const int k0 = nn ? __ffs((int)nw) - 1 : 0, k1 = nn > 1 ? 31 - __clz((int)nw) : 0;
                                const int nvar = nn == 0 ? 1 : (nn == 1 ? 4 : 16);
                                const int s0 = s_first[cs], s1 = s_first[cs + 1];
#pragma unroll 1
                                for (int v = 0; v < nvar; v++)
                                for (int sb = s0; sb < s1; sb += 32) {
                                    uint32_t W = W0;
                                    if (nn >= 1) W = nr_anchor_subst(W, k0, (uint32_t)(v & 3));
                                    if (nn >= 2) W = nr_anchor_subst(W, k1, (uint32_t)(v >> 2));
                                    const int si = sb + (int)lane;
                                    uint32_t start = 0, cnt = 0;
                                    if (si < s1) {
                                        uint32_t key;
                                        if (nr_anchor_apply(script_unpack(s_scripts[si]), W, e, &key)) {
                                            start = __ldg(P.pstart + key);
                                            cnt = __ldg(P.pstart + key + 1) - start;
                                            if (COUNT) c_keys++;
                                        }
                                    }
                                    // rows of all lanes into the queue
                                    uint32_t incl = cnt;
#pragma unroll
                                    for (int o = 1; o < 32; o <<= 1) {
                                        const uint32_t v = __shfl_up_sync(0xffffffffu, incl, o);
                                        if ((int)lane >= o) incl += v;
                                    }
                                    const uint32_t total = __shfl_sync(0xffffffffu, incl, 31);
                                    if (total == 0) continue;
                                    if (COUNT) c_rows += cnt;
                                    if (total > NR_AQCAP) { acc.overflow = 1; continue; }
                                    if (acc.qn + (int)total > NR_AQCAP) drain_a<COUNT>(P, sm, acc, m, c_ver, c_pass);
                                    uint32_t pos = (uint32_t)acc.qn + incl - cnt;
                                    for (uint32_t r = 0; r < cnt; r++)
                                        sm.queue[pos++] = (start + r) | ((uint32_t)strand << 31);
                                    acc.qn += (int)total;
                                    __syncwarp();
                                }
                            }
                        }
                        drain_a<COUNT>(P, sm, acc, m, c_ver, c_pass);
                    }
                }
                if (acc.overflow || (acc.best == 3 && P.resolve_below)) to_list = true;
            }
            if (to_list) {
                if (lane == 0) P.list[atomicAdd(P.list_count, 1u)] = (uint32_t)cand;
                if (COUNT) c_listed += lane == 0;
                continue;
            }
            if (acc.best == 3) {
                if (lane == 0) {
                    P.o_idx[cand] = -1; P.o_score[cand] = NR_SCORE_BELOW; P.o_nbest[cand] = 0;
                    P.o_flags[cand] = NR_FLAG_BELOW | NR_FLAG_NO_UMI; P.o_umi[cand] = NR_UMI_NONE;
                }
                continue;
            }
            const uint32_t kmine = (int)lane < acc.nb ? acc.key : 0xFFFFFFFFu;
            const uint32_t kmin = __reduce_min_sync(0xffffffffu, kmine);
            if (lane == 0) {
                const int strand = (int)(kmin & 1u);
                const uint32_t en = kmin >> 1;
                const int score = P.L - acc.best;
                int u = -1;
                if (!strand) {
                    nr_deep_rows r0 = sm.rows[0];
                    r0.edge = 1ull;
                    u = nr_deep_umi_row<2>(r0, __ldg(P.lo + en), P.hi ? __ldg(P.hi + en) : 0u,
                                           P.nm ? __ldg(P.nm + en) : 0u, P.L, m, P.padL, P.padR, acc.best);
                }
                uint32_t fl = 0;
                if (acc.nb > 1) fl |= NR_FLAG_TIE;
                if (strand) fl |= NR_FLAG_RC;
                if (score < P.min_score) fl |= NR_FLAG_BELOW;
                if (u < 0) fl |= NR_FLAG_NO_UMI;
                P.o_idx[cand] = (int32_t)en; P.o_score[cand] = (int8_t)score;
                P.o_nbest[cand] = (uint8_t)acc.nb; P.o_flags[cand] = (uint8_t)fl;
                P.o_umi[cand] = (uint8_t)(u < 0 ? NR_UMI_NONE : u);
            }
        }
    }
    if (COUNT && P.counters) {
        atomicAdd(P.counters + 0, c_keys);
        atomicAdd(P.counters + 1, c_rows);
        atomicAdd(P.counters + 2, c_ver);
        atomicAdd(P.counters + 3, c_pass);
        atomicAdd(P.counters + 4, c_listed);
    }
}

}  // namespace

// Enqueue the anchored matcher on `stream`.  The workspace header (list count, tile counter) must
// have been zeroed on the stream.
int nr_launch_anchored(const nr_whitelist *wl, const void *d_bases, const uint8_t *d_meta,
                       const uint64_t *d_nmask, uint64_t n_cand, int min_score, int resolve_below, int32_t *d_idx,
                       int8_t *d_score, uint8_t *d_nbest, uint8_t *d_flags, uint8_t *d_umi,
                       uint32_t *d_list, uint32_t *d_list_count, unsigned long long *d_tile_next,
                       unsigned long long *d_counters, cudaStream_t stream)
{
    if (n_cand == 0) return NR_OK;
    if (!wl->has_anchor) {
        nr_set_error("anchored matcher needs cores with a constant middle (8 + linker + tail)");
        return NR_EUNSUPPORTED;
    }
    nr_anchor_params P;
    P.pstart = wl->d_anchor_start; P.prows = wl->d_anchor_rows;
    P.lo = wl->d_lo; P.hi = wl->d_hi; P.nm = wl->d_nm;
    P.link = wl->anchor_link; P.L = (int)wl->L; P.Lk = wl->anchor_lk;
    P.padL = (int)wl->pad_l; P.padR = (int)wl->pad_r;
    P.bases = (const uint4 *)d_bases; P.meta = d_meta; P.nmask = d_nmask; P.n_cand = n_cand;
    P.min_score = min_score; P.resolve_below = resolve_below;
    P.o_idx = d_idx; P.o_score = d_score; P.o_nbest = d_nbest; P.o_flags = d_flags; P.o_umi = d_umi;
    P.list = d_list; P.list_count = d_list_count; P.tile_next = d_tile_next; P.counters = d_counters;
    int sms = 148, per_sm = 1;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, wl->device);
    if (d_counters)
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, nr_match_anchored_kernel<true>, NR_AWARPS * 32, 0);
    else
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, nr_match_anchored_kernel<false>, NR_AWARPS * 32, 0);
    if (per_sm < 1) per_sm = 1;
    const uint64_t tiles = (n_cand + 31) / 32;
    const uint64_t want = (tiles + NR_AWARPS - 1) / NR_AWARPS;
    const uint64_t cap = (uint64_t)sms * (uint64_t)per_sm;
    const unsigned grid = (unsigned)(want < cap ? want : cap);
    if (d_counters)
        nr_match_anchored_kernel<true><<<grid, NR_AWARPS * 32, 0, stream>>>(P);
    else
        nr_match_anchored_kernel<false><<<grid, NR_AWARPS * 32, 0, stream>>>(P);
    NR_CHECK_CUDA(cudaGetLastError());
    return NR_OK;
}
