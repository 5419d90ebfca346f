// nr_umi.cu -- UMI collapse: segmented sort by (cell barcode, transcript, UMI) + clustering.
//
// Replaces the per-barcode exact dedup of the reference (utils.py:759-777 = 910-928 = 1212-1230:
// np.unique over the UMI strings of each barcode) and finishes what
// utils.make_count_mtx_3p10XGEX (utils.py:1523-1548) starts.  The reference only ever counts
// exact-distinct UMIs (max_dist = 0, gene ignored: pass gene = 0).  max_dist = 1 adds the
// one-hop directional merge defined in DESIGN.md (checked against oracle/nr_oracle.c's twin):
// inside a (barcode, gene) group the distinct UMIs are walked by (reads desc, umi asc); a UMI
// joins the earliest representative within Hamming distance 1 whose reads >= 2 * reads - 1,
// otherwise it becomes a representative itself.
//
// Pipeline (all on `stream`, no host synchronisation):
//   1  LSD radix sort: by umi (2 * umi_len bits), then stable by (barcode << 32 | gene)
//   2  head flags + scans -> distinct-UMI ids, group ids; run lengths -> reads per distinct UMI
//   3  small groups: one thread / one warp ranks the distinct UMIs, walks them and picks the
//      representatives; large groups: hash sets + parallel rounds over the whole grid (below)
//   4  reads per representative, compaction of the representatives into the group table,
//      scatter of the representative UMI back to input order
#include <cooperative_groups.h>
#include <cub/cub.cuh>

#include "nr_common.cuh"

namespace cg = cooperative_groups;

namespace {

struct UmiWs {
    // n-sized arrays
    uint32_t *umi_a, *umi_b, *idx_a, *idx_b;
    uint64_t *key_a, *key_b;
    uint32_t *s_bc, *s_gene, *s_umi;       // sorted records
    uint32_t *head_u, *head_g;             // flags, then exclusive ids after the scan
    uint32_t *du_id, *grp_id;              // per sorted record
    uint32_t *du_first;                    // per distinct: first sorted position (+ sentinel)
    uint32_t *grp_first;                   // per group: first distinct id (+ sentinel)
    uint32_t *du_rep;                      // per distinct: distinct id of its representative
    uint32_t *rep_reads;                   // per distinct: reads of the cluster it represents
    uint32_t *rep_flag, *rep_pos;          // per distinct
    uint32_t *rep_u, *rep_c;               // work lists of the large-group rounds (distinct ids)
    uint4 *rec;                            // per distinct UMI of a large group: UMI, walk position, reads, group
    uint32_t *htab;                        // 4 slots per distinct UMI: hash set UMI -> distinct id per large group
    uint32_t *totals;                      // [0] n distinct, [1] n groups, [2] n reps, [8..15] list lengths and cursors, [16] key-width flag
    void *cub_tmp;
    size_t cub_bytes;
};

size_t align_up(size_t x) { return (x + 255) & ~(size_t)255; }

size_t carve(uint8_t *base, uint64_t n, UmiWs *w, size_t cub_bytes)
{
    size_t off = 0;
    auto take = [&](size_t bytes) { size_t o = off; off += align_up(bytes); return base ? base + o : nullptr; };
    size_t n4 = (size_t)(n + 1) * 4, n8 = (size_t)(n + 1) * 8;
    uint32_t **a32[] = {&w->umi_a, &w->umi_b, &w->idx_a, &w->idx_b, &w->s_bc, &w->s_gene, &w->s_umi,
                        &w->head_u, &w->head_g, &w->du_id, &w->grp_id, &w->du_first,
                        &w->grp_first, &w->du_rep, &w->rep_reads,
                        &w->rep_flag, &w->rep_pos, &w->rep_u, &w->rep_c};
    for (auto p : a32) *p = (uint32_t *)take(n4);
    w->htab = (uint32_t *)take(4 * n4);
    w->rec = (uint4 *)take(4 * n4);
    w->key_a = (uint64_t *)take(n8);
    w->key_b = (uint64_t *)take(n8);
    w->totals = (uint32_t *)take(256);
    w->cub_tmp = take(cub_bytes);
    w->cub_bytes = cub_bytes;
    return off;
}

size_t cub_temp_bytes(uint64_t n)
{
    // The in/out (non-DoubleBuffer) radix sort keeps one more copy of keys and values for its
    // odd passes (12 B per record for u64 keys + u32 values) next to the per-tile histograms;
    // a closed form keeps this callable without a device.
    return align_up((size_t)n * 16 + (size_t)(n / 64 + 1) * 64 + (8u << 20));
}

// key widths a caller may declare (nr_umi_collapse_device_keyed): a value that does not fit raises
// *bad, which k_emit turns into n_groups = ~0
__device__ __forceinline__ bool fits(uint32_t v, int bits) { return bits >= 32 || (v >> bits) == 0u; }

__global__ void k_init(const uint32_t *umi, uint64_t n, int umi_bits, uint32_t *umi_a, uint32_t *idx_a,
                       uint32_t *bad)
{
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint32_t u = umi[i];
    umi_a[i] = u; idx_a[i] = (uint32_t)i;
    if (!fits(u, umi_bits)) *bad = 1u;
}

__global__ void k_gather_key(const uint32_t *bc, const uint32_t *gene, const uint32_t *idx,
                             uint64_t n, int bc_bits, int gene_bits, uint64_t *key, uint32_t *bad)
{
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint32_t s = idx[i], b = bc[s], g = gene[s];
    key[i] = ((uint64_t)b << gene_bits) | g;
    if (!fits(b, bc_bits) || !fits(g, gene_bits)) *bad = 1u;
}

// (barcode, gene, umi) in ONE key when the declared widths add up to <= 64 bits
__global__ void k_compose(const uint32_t *bc, const uint32_t *gene, const uint32_t *umi, uint64_t n,
                          int bc_bits, int gene_bits, int umi_bits, uint64_t *key, uint32_t *idx,
                          uint32_t *bad)
{
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint32_t b = bc[i], g = gene[i], u = umi[i];
    key[i] = ((((uint64_t)b << gene_bits) | g) << umi_bits) | u;
    idx[i] = (uint32_t)i;
    if (!fits(b, bc_bits) || !fits(g, gene_bits) || !fits(u, umi_bits)) *bad = 1u;
}

// sorted records + head flags.  umi_bits < 0: key = (barcode, gene), the UMI comes through idx;
// else key = (barcode, gene, umi)
__global__ void k_sorted(const uint64_t *key, const uint32_t *umi_in, const uint32_t *idx,
                         uint64_t n, int gene_bits, int umi_bits, uint32_t *s_bc, uint32_t *s_gene,
                         uint32_t *s_umi, uint32_t *head_u, uint32_t *head_g)
{
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint64_t gmask = (1ull << gene_bits) - 1ull;
    uint64_t k = key[i], kp = i > 0 ? key[i - 1] : 0ull;
    uint32_t u, up = 0;
    if (umi_bits < 0) {
        u = umi_in[idx[i]];
        if (i > 0) up = umi_in[idx[i - 1]];
    } else {
        const uint64_t umask = (1ull << umi_bits) - 1ull;
        u = (uint32_t)(k & umask); up = (uint32_t)(kp & umask);
        k >>= umi_bits; kp >>= umi_bits;
    }
    s_bc[i] = (uint32_t)(k >> gene_bits); s_gene[i] = (uint32_t)(k & gmask); s_umi[i] = u;
    const bool hg = i == 0 || kp != k;
    head_u[i] = hg || up != u; head_g[i] = hg;
}

__global__ void k_ids(const uint32_t *scan_u, const uint32_t *scan_g, uint64_t n, uint32_t *du_id,
                      uint32_t *grp_id, uint32_t *du_first, uint32_t *grp_first, uint32_t *totals)
{
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint32_t du = scan_u[i] - 1, g = scan_g[i] - 1;
    bool hu = i == 0 || scan_u[i - 1] != scan_u[i];
    bool hg = i == 0 || scan_g[i - 1] != scan_g[i];
    du_id[i] = du; grp_id[i] = g;
    if (hu) du_first[du] = (uint32_t)i;
    if (hg) grp_first[g] = du;
    if (i == n - 1) {
        totals[0] = du + 1; totals[1] = g + 1;
        du_first[du + 1] = (uint32_t)n;      // sentinels
        grp_first[g + 1] = du + 1;
    }
}

__device__ __forceinline__ int hamming_2bit(uint32_t a, uint32_t b)
{
    uint32_t x = a ^ b;
    x = (x | (x >> 1)) & 0x55555555u;
    return __popc(x);
}

// Walk order of the distinct UMIs of a group: (reads desc, umi asc).  Distinct ids are
// umi-ascending inside a group, so "a is walked before b" <=> reads(a) > reads(b), or the reads are
// equal and id(a) < id(b): no kernel needs the order materialised.
#define NR_UMI_LARGE 96      // groups with more distinct UMIs go to the hash-set rounds:
#define NR_UMI_BLOCK 2048    // up to here one block per group in shared memory, beyond the whole grid

__device__ __forceinline__ bool umi_joins(uint32_t rep_umi, uint32_t rep_cnt, uint32_t u,
                                          uint32_t cnt, int max_dist)
{
    return hamming_2bit(rep_umi, u) <= max_dist && rep_cnt + 1 >= 2 * cnt;
}

// max_dist 0: every distinct UMI is its own representative
__global__ void k_self(const uint32_t *__restrict__ totals, uint32_t *__restrict__ du_rep)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < totals[0]) du_rep[i] = i;
}

// Small (barcode, gene) groups, sequential walk.  Nearly every group of a gene-expression library
// has a handful of distinct UMIs: up to NR_UMI_TINY are ranked and walked by ONE thread in
// registers (all loads of the group in flight together).  Groups of up to NR_UMI_LARGE = 3 x 32 go
// onto a list (they cluster in the deeply sequenced cells: the list spreads them over the grid)
// and are walked by one warp each: the lanes rank the UMIs against each other, a pass through
// shared memory puts them in walk order, lane l then holds walk positions l, 32 + l, 64 + l; a
// step broadcasts the UMI at position r and every lane compares it with the representatives
// among its own positions (all earlier than r by construction); the lowest position that accepts
// it wins.  No global memory access inside the walk.
#define NR_UMI_TINY 8

__global__ void __launch_bounds__(256)
k_cluster_tiny(const uint32_t *__restrict__ s_umi, const uint32_t *__restrict__ du_first,
               const uint32_t *__restrict__ grp_first, const uint32_t *__restrict__ totals,
               int max_dist, uint32_t *__restrict__ du_rep, uint32_t *__restrict__ medium,
               uint32_t *__restrict__ medium_count, uint32_t *__restrict__ big,
               uint32_t *__restrict__ big_count)
{
    const uint32_t n_groups = totals[1];
    const uint32_t nthr = gridDim.x * blockDim.x;
    for (uint32_t g = blockIdx.x * blockDim.x + threadIdx.x; g < n_groups; g += nthr) {
        const uint32_t d0 = grp_first[g], nd = grp_first[g + 1] - d0;
        if (nd > NR_UMI_BLOCK) continue;
        if (nd > NR_UMI_LARGE) {
            big[atomicAdd(big_count, 1u)] = g;
            continue;
        }
        if (nd > NR_UMI_TINY) {
            medium[atomicAdd(medium_count, 1u)] = g;
            continue;
        }
        if (nd == 1) {
            du_rep[d0] = d0;
            continue;
        }
        uint32_t f[NR_UMI_TINY + 1], u[NR_UMI_TINY], c[NR_UMI_TINY], rank[NR_UMI_TINY];
#pragma unroll
        for (int r = 0; r <= NR_UMI_TINY; r++) f[r] = (uint32_t)r <= nd ? du_first[d0 + r] : 0u;
#pragma unroll
        for (int r = 0; r < NR_UMI_TINY; r++) {
            u[r] = (uint32_t)r < nd ? s_umi[f[r]] : 0u;
            c[r] = (uint32_t)r < nd ? f[r + 1] - f[r] : 0u;       // absent: 0 reads, ranked last
        }
#pragma unroll
        for (int r = 0; r < NR_UMI_TINY; r++) {
            uint32_t k = 0;
#pragma unroll
            for (int q = 0; q < NR_UMI_TINY; q++)
                k += (c[q] > c[r] || (c[q] == c[r] && q < r)) ? 1u : 0u;
            rank[r] = k;
        }
        uint32_t reps = 0;
        for (uint32_t step = 0; step < nd; step++) {
            uint32_t ue = 0, ce = 0, e = 0;
#pragma unroll
            for (int r = 0; r < NR_UMI_TINY; r++)
                if (rank[r] == step) { ue = u[r]; ce = c[r]; e = (uint32_t)r; }
            uint32_t best = NR_UMI_TINY, rep = e;
#pragma unroll
            for (int q = 0; q < NR_UMI_TINY; q++)
                if (((reps >> q) & 1u) && rank[q] < best && umi_joins(u[q], c[q], ue, ce, max_dist)) {
                    best = rank[q];
                    rep = (uint32_t)q;
                }
            if (rep == e) reps |= 1u << e;
            du_rep[d0 + e] = d0 + rep;
        }
    }
}

// the walk of one medium group by one warp; NS = 32-lane slots in use (1..3), uniform per warp
template <int NS>
__device__ __forceinline__ void medium_walk(uint32_t gd0, uint32_t gnd, const uint32_t (&u_id)[3],
                                            const uint32_t (&c_id)[3], int max_dist, uint32_t lane,
                                            uint32_t (*s_walk)[NR_UMI_LARGE], uint32_t *__restrict__ du_rep)
{
    constexpr uint32_t FULL = 0xffffffffu, NONE = 0xFFFFFFFFu;
    auto pick = [](const uint32_t (&a)[NS], uint32_t t) {
        uint32_t v = a[0];
#pragma unroll
        for (int q = 1; q < NS; q++) v = t == (uint32_t)q ? a[q] : v;
        return v;
    };
    uint32_t d[NS], u[NS], c[NS], res[NS], repm[NS], rank[NS];
#pragma unroll
    for (int t = 0; t < NS; t++) { u[t] = u_id[t]; c[t] = c_id[t]; rank[t] = 0u; repm[t] = 0u; }
    for (uint32_t j = 0; j < gnd; j++) {
        const uint32_t cj = __shfl_sync(FULL, pick(c, j >> 5), j & 31u);
#pragma unroll
        for (int t = 0; t < NS; t++) rank[t] += (cj > c[t] || (cj == c[t] && j < 32u * t + lane)) ? 1u : 0u;
    }
    __syncwarp();
#pragma unroll
    for (int t = 0; t < NS; t++)
        if (32u * t + lane < gnd) {
            s_walk[0][rank[t]] = gd0 + 32u * t + lane;
            s_walk[1][rank[t]] = u[t];
            s_walk[2][rank[t]] = c[t];
        }
    __syncwarp();
#pragma unroll
    for (int t = 0; t < NS; t++) {
        const uint32_t r = 32u * t + lane;
        d[t] = u[t] = c[t] = res[t] = 0u;
        if (r < gnd) {
            d[t] = s_walk[0][r];
            u[t] = s_walk[1][r];
            c[t] = s_walk[2][r];
        }
    }
    for (uint32_t r = 0; r < gnd; r++) {
        const uint32_t t = r >> 5, src = r & 31u;
        const uint32_t ur = __shfl_sync(FULL, pick(u, t), src);
        const uint32_t cr = __shfl_sync(FULL, pick(c, t), src);
        uint32_t rep = NONE;
#pragma unroll
        for (int q = 0; q < NS; q++) {
            const bool ok = ((repm[q] >> lane) & 1u) && umi_joins(u[q], c[q], ur, cr, max_dist);
            const uint32_t m = __ballot_sync(FULL, ok);
            if (rep == NONE && m) rep = __shfl_sync(FULL, d[q], __ffs(m) - 1);
        }
#pragma unroll
        for (int q = 0; q < NS; q++) {
            if (rep == NONE && t == (uint32_t)q) repm[q] |= 1u << src;
            if (lane == src && t == (uint32_t)q) res[q] = rep == NONE ? d[q] : rep;
        }
    }
#pragma unroll
    for (int t = 0; t < NS; t++)
        if (32u * t + lane < gnd) du_rep[d[t]] = res[t];
}

__global__ void __launch_bounds__(256)
k_cluster_medium(const uint32_t *__restrict__ s_umi, const uint32_t *__restrict__ du_first,
                 const uint32_t *__restrict__ grp_first, int max_dist, uint32_t *__restrict__ du_rep,
                 const uint32_t *__restrict__ medium, const uint32_t *__restrict__ medium_count)
{
    __shared__ uint32_t s_walk[8][3][NR_UMI_LARGE];        // per warp: id, UMI, reads by walk position
    const uint32_t lane = threadIdx.x & 31u, wib = threadIdx.x >> 5;
    const uint32_t n_medium = *medium_count;
    const uint32_t warps = (gridDim.x * blockDim.x) >> 5;
    // by id (= UMI ascending): lane l holds ids gd0 + l, + 32, + 64; the next group's loads are
    // issued before the current group is walked
    auto fetch = [&](uint32_t i, uint32_t &gd0, uint32_t &gnd, uint32_t (&u)[3], uint32_t (&c)[3]) {
        gd0 = gnd = 0u;
#pragma unroll
        for (int t = 0; t < 3; t++) u[t] = c[t] = 0u;
        if (i >= n_medium) return;
        const uint32_t g = medium[i];
        gd0 = grp_first[g];
        gnd = grp_first[g + 1] - gd0;
#pragma unroll
        for (int t = 0; t < 3; t++) {
            const uint32_t e = 32u * t + lane;
            if (e < gnd) {
                const uint32_t f = du_first[gd0 + e];
                c[t] = du_first[gd0 + e + 1] - f;
                u[t] = s_umi[f];
            }
        }
    };
    uint32_t i = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    uint32_t gd0, gnd, u[3], c[3];
    fetch(i, gd0, gnd, u, c);
    while (i < n_medium) {
        uint32_t n_gd0, n_gnd, n_u[3], n_c[3];
        fetch(i + warps, n_gd0, n_gnd, n_u, n_c);
        if (gnd <= 32) medium_walk<1>(gd0, gnd, u, c, max_dist, lane, s_walk[wib], du_rep);
        else if (gnd <= 64) medium_walk<2>(gd0, gnd, u, c, max_dist, lane, s_walk[wib], du_rep);
        else medium_walk<3>(gd0, gnd, u, c, max_dist, lane, s_walk[wib], du_rep);
        i += warps;
        gd0 = n_gd0; gnd = n_gnd;
#pragma unroll
        for (int t = 0; t < 3; t++) { u[t] = n_u[t]; c[t] = n_c[t]; }
    }
}

// Large groups (more than NR_UMI_LARGE distinct UMIs), no sequential walk.  The directional rule -- a
// UMI joins the EARLIEST representative (in walk order) within Hamming distance 1 whose reads >=
// 2 * reads - 1, else it becomes a representative -- only depends on the fate of the UMI's eligible
// neighbours at earlier walk positions, and a UMI of len bases has 3 * len neighbours.  So:
//   init    every distinct UMI of a large group goes into the group's hash set (UMI -> distinct id)
//           and is marked undecided;
//   rounds  one cooperative launch filling the GPU; 8 lanes = one undecided UMI of the work list:
//           look the 3 * len neighbours up; among the eligible ones at earlier positions let r =
//           earliest representative, q = earliest undecided.  r < q: join r (everything before r
//           is known not to be a representative).  No undecided and no representative: become a
//           representative.  Otherwise go onto the next round's work list.  A grid-wide barrier
//           separates the rounds; the earliest undecided position of every group is decided in
//           every round, so the loop ends, after as many rounds as the longest chain of
//           neighbours waiting on each other (a handful to a few dozen).
// Decisions are final once written, so reading a neighbour's state while its own thread decides it
// in the same round is harmless: "undecided" only makes the reader wait one more round.
// A (cell, gene) group of 20 000 distinct UMIs is thereby settled by the whole GPU in a few rounds
// instead of by one block in 160 sequential rounds of 128 -- the straggler of the multi-GPU
// collapse, where the deepest group grows with the number of ranks.  State lives in du_rep: NONE
// undecided, own id = representative, other id = joined.

__device__ __forceinline__ uint32_t umi_hash(uint32_t u, uint32_t bits)
{
    return (u * 0x9E3779B1u) >> (32u - bits);
}

__device__ __forceinline__ uint32_t large_hbits(uint32_t nd)
{
    return 32u - (uint32_t)__clz((int)(2u * nd - 1u));   // smallest power of two >= 2 nd, of the group's 4 nd slots
}

// Groups of NR_UMI_LARGE < nd <= NR_UMI_BLOCK distinct UMIs: the rounds above by ONE block with the
// group in shared memory -- UMI words, reads, state and the hash set (UMI -> index in the group,
// 2 nd slots) -- so a probe is a shared-memory access and a round ends at a __syncthreads().
// Which neighbours are eligible and walked earlier never changes: the first round keeps up to
// NBR of them per UMI (there are ~3 on average), later rounds only re-read their states
// (a UMI with more keeps probing).  State: 0xFFFF undecided, else the index of the representative.
// 256 threads, four blocks per SM; blocks take groups off the list one at a time (*next): group
// sizes differ by 20x.  A group is bound to one SM here, ~0.1 us per UMI: beyond NR_UMI_BLOCK
// UMIs the grid-wide rounds are faster (measured: a 3 500-UMI group took 190 us in one
// 1024-thread block, 50 us on the grid).
constexpr size_t block_smem(int cap, int nbr) { return (size_t)cap * (4 + 4 + 2 + 4 + 2 * nbr + 1); }

template <int CAP, int NBR, int THREADS>
__global__ void __launch_bounds__(THREADS)
k_cluster_block(const uint32_t *__restrict__ s_umi, const uint32_t *__restrict__ du_first,
                const uint32_t *__restrict__ grp_first, int umi_len, uint32_t *__restrict__ du_rep,
                const uint32_t *__restrict__ big, const uint32_t *__restrict__ big_count,
                uint32_t *__restrict__ next)
{
    __shared__ uint32_t s_bi;
    constexpr uint32_t EMPTY = 0xFFFFu, UNPROBED = 0xFFu, OVERFLOW = 0xFEu;
    constexpr unsigned long long LAST = ~0ull;
    extern __shared__ uint32_t s_dyn[];
    uint32_t *su = s_dyn, *sc = s_dyn + CAP;
    unsigned short *sst = (unsigned short *)(s_dyn + 2 * CAP);
    unsigned short *tab = sst + CAP;                      // 2 * CAP slots
    unsigned short *nbr = tab + 2 * CAP;                  // NBR per UMI
    unsigned char *nn = (unsigned char *)(nbr + NBR * CAP);
    volatile unsigned short *vst = sst;
    const uint32_t n_big = *big_count, per = 3u * (uint32_t)umi_len;
    for (;;) {
        __syncthreads();                                   // the previous group is done with the arrays
        if (threadIdx.x == 0) s_bi = atomicAdd(next, 1u);
        __syncthreads();
        const uint32_t bi = s_bi;
        if (bi >= n_big) break;
        const uint32_t g = big[bi];
        const uint32_t d0 = grp_first[g], nd = grp_first[g + 1] - d0;
        const uint32_t hbits = large_hbits(nd), hmask = (1u << hbits) - 1u;
        for (uint32_t i = threadIdx.x; i <= hmask; i += blockDim.x) tab[i] = (unsigned short)EMPTY;
        for (uint32_t i = threadIdx.x; i < nd; i += blockDim.x) {
            const uint32_t f = du_first[d0 + i];
            sc[i] = du_first[d0 + i + 1] - f;
            su[i] = s_umi[f];
            sst[i] = (unsigned short)EMPTY;
            nn[i] = (unsigned char)UNPROBED;
        }
        __syncthreads();
        for (uint32_t i = threadIdx.x; i < nd; i += blockDim.x) {
            uint32_t slot = umi_hash(su[i], hbits);
            while (atomicCAS(&tab[slot], (unsigned short)EMPTY, (unsigned short)i) != EMPTY) slot = (slot + 1u) & hmask;
        }
        __syncthreads();
        for (;;) {
            uint32_t waiting = 0;
            for (uint32_t i = threadIdx.x; i < nd; i += blockDim.x) {
                if (vst[i] != EMPTY) continue;
                const uint32_t u = su[i], c = sc[i];
                // walk-order keys (~reads << 16 | index: smaller = earlier; indices < 2^16)
                const unsigned long long mine = ((unsigned long long)(~c) << 16) | i;
                unsigned long long rbest = LAST, qmin = LAST;
                auto consider = [&](uint32_t e, unsigned long long theirs) {
                    const uint32_t st = vst[e];
                    if (st == e) rbest = min(rbest, theirs);
                    else if (st == EMPTY) qmin = min(qmin, theirs);
                };
                const uint32_t known = nn[i];
                if (known <= NBR) {
                    for (uint32_t j = 0; j < known; j++) {
                        const uint32_t e = nbr[i * NBR + j];
                        consider(e, ((unsigned long long)(~sc[e]) << 16) | e);
                    }
                } else {
                    uint32_t m = 0;
                    for (uint32_t k = 0; k < per; k++) {
                        const uint32_t key = u ^ ((k % 3u + 1u) << (2u * (k / 3u)));
                        uint32_t slot = umi_hash(key, hbits);
                        for (uint32_t e = tab[slot]; e != EMPTY; slot = (slot + 1u) & hmask, e = tab[slot]) {
                            if (su[e] != key) continue;
                            const uint32_t ce = sc[e];
                            const unsigned long long theirs = ((unsigned long long)(~ce) << 16) | e;
                            if (theirs < mine && ce + 1 >= 2 * c) {
                                consider(e, theirs);
                                if (m < NBR) nbr[i * NBR + m] = (unsigned short)e;
                                m++;
                            }
                            break;
                        }
                    }
                    if (known == UNPROBED) nn[i] = (unsigned char)(m <= NBR ? m : OVERFLOW);
                }
                if (rbest < qmin) vst[i] = (unsigned short)(rbest & 0xFFFFu);
                else if (qmin == LAST) vst[i] = (unsigned short)i;
                else waiting++;
            }
            if (__syncthreads_count(waiting != 0) == 0) break;
        }
        for (uint32_t i = threadIdx.x; i < nd; i += blockDim.x) du_rep[d0 + i] = d0 + sst[i];
    }
}

// rec[], per distinct UMI of a large group: x = UMI word, y = reads, z = group id.  One 16-byte
// load tells a prober whether a slot holds the neighbour it looks for, whether that neighbour is
// eligible and whether it is walked earlier.

__global__ void __launch_bounds__(256)
k_large_init(const uint32_t *__restrict__ s_umi, const uint32_t *__restrict__ du_first,
             const uint32_t *__restrict__ grp_id, const uint32_t *__restrict__ grp_first,
             const uint32_t *__restrict__ totals,
             uint32_t *__restrict__ du_rep, uint4 *__restrict__ rec, uint32_t *__restrict__ htab,
             uint32_t *__restrict__ work, uint32_t *__restrict__ work_count)
{
    constexpr uint32_t NONE = 0xFFFFFFFFu;
    const uint32_t d = blockIdx.x * blockDim.x + threadIdx.x;
    if (d >= totals[0]) return;
    const uint32_t f = du_first[d], g = grp_id[f];
    const uint32_t d0 = grp_first[g], nd = grp_first[g + 1] - d0;
    if (nd <= NR_UMI_BLOCK) return;
    const uint32_t u = s_umi[f];
    rec[d] = make_uint4(u, du_first[d + 1] - f, g, 0u);
    du_rep[d] = NONE;
    uint32_t *tab = htab + 4ull * d0;
    const uint32_t hbits = large_hbits(nd), hmask = (1u << hbits) - 1u;
    uint32_t slot = umi_hash(u, hbits);
    while (atomicCAS(&tab[slot], NONE, d) != NONE) slot = (slot + 1u) & hmask;
    work[atomicAdd(work_count, 1u)] = d;
}

// One round for the undecided UMI d, run by NR_UMI_LANES adjacent lanes (`sub` = lane within them;
// the whole warp is converged here, lanes without an item pass valid = false).  Each lane looks up
// every NR_UMI_LANES-th neighbour, all its table loads in flight together, then all the record
// loads of the occupied slots; the few hits read the neighbour's state.  Returns true on lane 0
// of the group when d is now decided.
#define NR_UMI_LANES 8
#define NR_UMI_PROBES 6       // per lane: NR_UMI_LANES * NR_UMI_PROBES >= 3 * 16 neighbours

__device__ __forceinline__ bool large_round(uint32_t d, bool valid, uint32_t sub, uint32_t per,
                                            const uint32_t *__restrict__ grp_first,
                                            uint32_t *__restrict__ du_rep, const uint4 *__restrict__ rec,
                                            const uint32_t *__restrict__ htab)
{
    constexpr uint32_t NONE = 0xFFFFFFFFu;
    uint4 me = make_uint4(0, 0, 0, 0);
    uint32_t d0 = 0, hbits = 1;
    if (valid) {
        me = __ldg(rec + d);
        d0 = __ldg(grp_first + me.z);
        hbits = large_hbits(__ldg(grp_first + me.z + 1) - d0);
    }
    const uint32_t *tab = htab + 4ull * d0;
    const uint32_t hmask = (1u << hbits) - 1u;
    uint32_t key[NR_UMI_PROBES], slot[NR_UMI_PROBES], dj[NR_UMI_PROBES];
    uint4 rj[NR_UMI_PROBES];
#pragma unroll
    for (int j = 0; j < NR_UMI_PROBES; j++) {
        const uint32_t k = sub + NR_UMI_LANES * j;
        key[j] = me.x ^ ((k % 3u + 1u) << (2u * (k / 3u)));
        slot[j] = umi_hash(key[j], hbits);
        dj[j] = (valid && k < per) ? __ldg(tab + slot[j]) : NONE;
    }
#pragma unroll
    for (int j = 0; j < NR_UMI_PROBES; j++)
        if (dj[j] != NONE) rj[j] = __ldg(rec + dj[j]);
    // walk-order keys (~reads << 32 | distinct id; smaller = earlier) of the earliest representative
    // and of the earliest undecided UMI among the eligible neighbours walked before d
    constexpr unsigned long long LAST = ~0ull;
    const unsigned long long mine = ((unsigned long long)(~me.y) << 32) | d;
    unsigned long long rbest = LAST, qmin = LAST;
#pragma unroll
    for (int j = 0; j < NR_UMI_PROBES; j++) {
        uint32_t e = dj[j], sl = slot[j];
        uint4 r = rj[j];
        while (e != NONE) {
            if (r.x == key[j]) {
                const unsigned long long theirs = ((unsigned long long)(~r.y) << 32) | e;
                if (theirs < mine && r.y + 1 >= 2 * me.y) {
                    const uint32_t st = __ldcg(du_rep + e);
                    if (st == e) rbest = min(rbest, theirs);
                    else if (st == NONE) qmin = min(qmin, theirs);
                }
                break;
            }
            sl = (sl + 1u) & hmask;
            e = __ldg(tab + sl);
            if (e != NONE) r = __ldg(rec + e);
        }
    }
#pragma unroll
    for (int o = NR_UMI_LANES / 2; o > 0; o >>= 1) {
        rbest = min(rbest, __shfl_xor_sync(0xffffffffu, rbest, o));
        qmin = min(qmin, __shfl_xor_sync(0xffffffffu, qmin, o));
    }
    if (!valid || sub != 0) return false;
    if (rbest < qmin) { __stcg(du_rep + d, (uint32_t)rbest); return true; }
    if (qmin == LAST) { __stcg(du_rep + d, d); return true; }
    return false;
}

// work_a / work_b: the undecided UMIs of this and the next round; count[3]: their lengths in
// rotation (round r reads count[r % 3], appends to count[(r + 1) % 3], clears count[(r + 2) % 3])
__global__ void __launch_bounds__(256)
k_large_rounds(const uint32_t *__restrict__ grp_first, int umi_len, uint32_t *__restrict__ du_rep,
               const uint4 *__restrict__ rec, const uint32_t *__restrict__ htab, uint32_t *work_a,
               uint32_t *work_b, uint32_t *count)
{
    cg::grid_group grid = cg::this_grid();
    const uint32_t tid = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t item0 = tid / NR_UMI_LANES, sub = tid % NR_UMI_LANES;
    const uint32_t items = gridDim.x * blockDim.x / NR_UMI_LANES;
    const uint32_t per = 3u * (uint32_t)umi_len;
    for (uint32_t r = 0;; r++) {
        const uint32_t n_in = __ldcg(count + r % 3u);
        if (n_in == 0) break;
        if (tid == 0) __stcg(count + (r + 2u) % 3u, 0u);
        const uint32_t *in = (r & 1u) ? work_b : work_a;
        uint32_t *out = (r & 1u) ? work_a : work_b;
        for (uint32_t base = 0; base < n_in; base += items) {          // uniform trip count
            const uint32_t i = base + item0;
            const bool valid = i < n_in;
            const uint32_t d = valid ? __ldcg(in + i) : 0u;
            const bool done = large_round(d, valid, sub, per, grp_first, du_rep, rec, htab);
            if (valid && sub == 0 && !done) __stcg(out + atomicAdd(count + (r + 1u) % 3u, 1u), d);
        }
        grid.sync();
    }
}

__global__ void k_rep_reads(const uint32_t *du_first, const uint32_t *du_rep, const uint32_t *totals,
                            uint32_t *rep_reads, uint32_t *rep_flag)
{
    uint32_t nd = totals[0];
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < nd; i += gridDim.x * blockDim.x) {
        uint32_t r = du_rep[i];
        atomicAdd(&rep_reads[r], du_first[i + 1] - du_first[i]);
        rep_flag[i] = r == i;
    }
}

__global__ void k_emit(const uint32_t *s_bc, const uint32_t *s_gene, const uint32_t *s_umi,
                       const uint32_t *du_first, const uint32_t *rep_flag, const uint32_t *rep_pos,
                       const uint32_t *rep_reads, const uint32_t *totals, uint32_t *g_bc,
                       uint32_t *g_gene, uint32_t *g_umi, uint32_t *g_reads, uint64_t *n_groups)
{
    uint32_t nd = totals[0];
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < nd; i += gridDim.x * blockDim.x) {
        if (rep_flag[i]) {
            uint32_t o = rep_pos[i] - 1, s = du_first[i];
            g_bc[o] = s_bc[s]; g_gene[o] = s_gene[s]; g_umi[o] = s_umi[s]; g_reads[o] = rep_reads[i];
        }
        if (i == nd - 1) *n_groups = totals[16] ? ~0ull : (uint64_t)rep_pos[i];
    }
}

__global__ void k_scatter_rep(const uint32_t *idx, const uint32_t *du_id, const uint32_t *du_rep,
                              const uint32_t *du_first, const uint32_t *s_umi, uint64_t n,
                              uint32_t *rep_umi)
{
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) rep_umi[idx[i]] = s_umi[du_first[du_rep[du_id[i]]]];
}

}  // namespace

extern "C" size_t nr_umi_workspace_bytes(uint64_t n)
{
    UmiWs w;
    return carve(nullptr, n, &w, cub_temp_bytes(n));
}

extern "C" int nr_umi_collapse_device_keyed(const uint32_t *d_bc, const uint32_t *d_gene,
                                            const uint32_t *d_umi, uint64_t n, int umi_len, int max_dist,
                                            int bc_bits, int gene_bits, int umi_bits,
                                            uint32_t *d_rep_umi, uint64_t *d_n_groups, uint32_t *d_g_bc,
                                            uint32_t *d_g_gene, uint32_t *d_g_umi, uint32_t *d_g_reads,
                                            void *d_workspace, size_t workspace_bytes, void *stream)
{
    cudaStream_t st = (cudaStream_t)stream;
    if (!d_n_groups) { nr_set_error("nr_umi_collapse_device: null pointer"); return NR_EINVAL; }
    if (n == 0) {
        NR_CHECK_CUDA(cudaMemsetAsync(d_n_groups, 0, sizeof(uint64_t), st));
        return NR_OK;
    }
    if (!d_bc || !d_gene || !d_umi || !d_rep_umi || !d_g_bc || !d_g_gene || !d_g_umi ||
        !d_g_reads || !d_workspace) {
        nr_set_error("nr_umi_collapse_device: null pointer");
        return NR_EINVAL;
    }
    if (umi_len < 1 || umi_len > 16 || max_dist < 0 || max_dist > 1 || n >= (1ull << 31)) {
        nr_set_error("nr_umi_collapse_device: umi_len 1..16, max_dist 0..1, n < 2^31");
        return NR_EINVAL;
    }
    if (bc_bits < 1 || bc_bits > 32 || gene_bits < 1 || gene_bits > 32 || umi_bits < 2 * umi_len ||
        umi_bits > 32) {
        nr_set_error("nr_umi_collapse_device_keyed: bc_bits, gene_bits 1..32, umi_bits 2 * umi_len..32");
        return NR_EINVAL;
    }
    if (workspace_bytes < nr_umi_workspace_bytes(n)) {
        nr_set_error("nr_umi_collapse_device: workspace too small");
        return NR_EINVAL;
    }
    UmiWs w;
    carve((uint8_t *)d_workspace, n, &w, cub_temp_bytes(n));
    const int T = 256;
    const unsigned nb = (unsigned)((n + T - 1) / T);
    const int N = (int)n;

    // NR_UMI_TRACE=1: per-stage device times on stderr (synchronises; diagnosis only)
    static const bool trace = getenv("NR_UMI_TRACE") != nullptr;
    cudaEvent_t tev[10];
    int ntev = 0;
    const char *tname[10];
    auto mark = [&](const char *stage) {          // `stage` starts here
        if (!trace || ntev >= 10) return;
        tname[ntev] = stage;
        cudaEventCreate(&tev[ntev]);
        cudaEventRecord(tev[ntev++], st);
    };
    mark("sorts");
    uint32_t *bad = w.totals + 16;
    NR_CHECK_CUDA(cudaMemsetAsync(bad, 0, 4, st));
    const int key_bits = bc_bits + gene_bits, all_bits = key_bits + umi_bits;
    const bool one_sort = all_bits <= 64;
    size_t need = 0, need2 = 0, need3 = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, need, w.umi_a, w.umi_b, w.idx_a, w.idx_b, N, 0, 32, st);
    cub::DeviceRadixSort::SortPairs(nullptr, need2, w.key_a, w.key_b, w.idx_b, w.idx_a, N, 0, 64, st);
    cub::DeviceScan::InclusiveSum(nullptr, need3, w.head_u, w.head_u, N, st);
    if (need > w.cub_bytes || need2 > w.cub_bytes || need3 > w.cub_bytes) {
        nr_set_error("nr_umi_collapse_device: CUB needs %zu bytes of temporary storage, have %zu",
                     std::max(need, std::max(need2, need3)), w.cub_bytes);
        return NR_ENOMEM;
    }
    size_t tb = w.cub_bytes;
    if (one_sort) {
        // declared widths fit one 64-bit key: a single radix sort over all_bits bits
        k_compose<<<nb, T, 0, st>>>(d_bc, d_gene, d_umi, n, bc_bits, gene_bits, umi_bits, w.key_a, w.idx_b, bad);
        NR_CHECK_CUDA(cub::DeviceRadixSort::SortPairs(w.cub_tmp, tb, w.key_a, w.key_b, w.idx_b, w.idx_a,
                                                      N, 0, all_bits, st));
    } else {
        // LSD: by umi, then stable by (barcode, gene).  The default width is all 32 bits of the UMI
        // word: a caller may carry escape codes above bit 2 * umi_len (the host path's N-containing
        // UMIs), and equal words must end up adjacent whatever their width
        k_init<<<nb, T, 0, st>>>(d_umi, n, umi_bits, w.umi_a, w.idx_a, bad);
        NR_CHECK_CUDA(cub::DeviceRadixSort::SortPairs(w.cub_tmp, tb, w.umi_a, w.umi_b, w.idx_a, w.idx_b,
                                                      N, 0, umi_bits, st));
        k_gather_key<<<nb, T, 0, st>>>(d_bc, d_gene, w.idx_b, n, bc_bits, gene_bits, w.key_a, bad);
        tb = w.cub_bytes;
        NR_CHECK_CUDA(cub::DeviceRadixSort::SortPairs(w.cub_tmp, tb, w.key_a, w.key_b, w.idx_b, w.idx_a,
                                                      N, 0, key_bits, st));
    }
    mark("ids");
    // sorted order: key_b (barcode, gene), idx_a (source record)
    k_sorted<<<nb, T, 0, st>>>(w.key_b, d_umi, w.idx_a, n, gene_bits, one_sort ? umi_bits : -1, w.s_bc,
                               w.s_gene, w.s_umi, w.head_u, w.head_g);
    tb = w.cub_bytes;
    NR_CHECK_CUDA(cub::DeviceScan::InclusiveSum(w.cub_tmp, tb, w.head_u, w.head_u, N, st));
    tb = w.cub_bytes;
    NR_CHECK_CUDA(cub::DeviceScan::InclusiveSum(w.cub_tmp, tb, w.head_g, w.head_g, N, st));
    k_ids<<<nb, T, 0, st>>>(w.head_u, w.head_g, n, w.du_id, w.grp_id, w.du_first, w.grp_first,
                            w.totals);
    int sms = 148;
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    mark("tiny groups");
    if (max_dist <= 0) {
        k_self<<<nb, T, 0, st>>>(w.totals, w.du_rep);
    } else {
        // totals[8..10]: work-list lengths of the large-group rounds, [11] medium groups, [12] block-kernel groups, [13] their cursor
        uint32_t *count = w.totals + 8, *work_a = w.rep_u, *work_b = w.rep_c;
        NR_CHECK_CUDA(cudaMemsetAsync(count, 0, 32, st));
        k_cluster_tiny<<<sms * 8, 256, 0, st>>>(w.s_umi, w.du_first, w.grp_first, w.totals, max_dist,
                                                w.du_rep, work_b, count + 3, work_a, count + 4);
        mark("medium groups");
        k_cluster_medium<<<sms * 8, 256, 0, st>>>(w.s_umi, w.du_first, w.grp_first, max_dist,
                                                  w.du_rep, work_b, count + 3);
        mark("block groups");
        {
            auto *kb = k_cluster_block<NR_UMI_BLOCK, 6, 256>;
            constexpr size_t smem = block_smem(NR_UMI_BLOCK, 6);
            static const cudaError_t attr = cudaFuncSetAttribute(kb, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            NR_CHECK_CUDA(attr);
            kb<<<sms * 4, 256, smem, st>>>(w.s_umi, w.du_first, w.grp_first, umi_len, w.du_rep, work_a, count + 4, count + 5);
        }
        mark("large: table + list");
        NR_CHECK_CUDA(cudaMemsetAsync(w.htab, 0xFF, (size_t)(n + 1) * 16, st));
        // large groups: hash sets and work list, then the rounds
        k_large_init<<<nb, T, 0, st>>>(w.s_umi, w.du_first, w.grp_id, w.grp_first, w.totals,
                                       w.du_rep, w.rec, w.htab, work_a, count);
        mark("large: rounds");
        int per_sm = 0;
        NR_CHECK_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_large_rounds, 256, 0));
        if (per_sm < 1) {
            nr_set_error("nr_umi_collapse_device: k_large_rounds does not fit an SM");
            return NR_ECUDA;
        }
        const uint32_t *c_gfirst = w.grp_first, *c_tab = w.htab;
        const uint4 *c_rec = w.rec;
        void *args[] = {&c_gfirst, &umi_len, &w.du_rep, &c_rec, &c_tab, &work_a, &work_b, &count};
        NR_CHECK_CUDA(cudaLaunchCooperativeKernel((void *)k_large_rounds, dim3(sms * std::min(per_sm, 2)),
                                                  dim3(256), args, 0, st));
    }
    mark("emit");
    NR_CHECK_CUDA(cudaMemsetAsync(w.rep_reads, 0, (size_t)(n + 1) * 4, st));
    k_rep_reads<<<sms * 8, 256, 0, st>>>(w.du_first, w.du_rep, w.totals, w.rep_reads, w.rep_flag);
    // rep_flag is defined for the first n_distinct entries only; the scan also runs over the
    // tail, whose prefix sums are never read back
    tb = w.cub_bytes;
    NR_CHECK_CUDA(cub::DeviceScan::InclusiveSum(w.cub_tmp, tb, w.rep_flag, w.rep_pos, N, st));
    k_emit<<<sms * 8, 256, 0, st>>>(w.s_bc, w.s_gene, w.s_umi, w.du_first, w.rep_flag, w.rep_pos,
                                    w.rep_reads, w.totals, d_g_bc, d_g_gene, d_g_umi, d_g_reads,
                                    d_n_groups);
    k_scatter_rep<<<nb, T, 0, st>>>(w.idx_a, w.du_id, w.du_rep, w.du_first, w.s_umi, n, d_rep_umi);
    NR_CHECK_CUDA(cudaGetLastError());
    if (trace) {
        const int last = ntev;
        cudaEvent_t end;
        cudaEventCreate(&end);
        cudaEventRecord(end, st);
        cudaEventSynchronize(end);
        for (int i = 0; i < last; i++) {
            float ms = 0.f;
            cudaEventElapsedTime(&ms, tev[i], i + 1 < last ? tev[i + 1] : end);
            fprintf(stderr, "[nr_umi] %-28s %8.1f us\n", tname[i], ms * 1e3f);
            cudaEventDestroy(tev[i]);
        }
        cudaEventDestroy(end);
    }
    return NR_OK;
}

extern "C" int nr_umi_collapse_device(const uint32_t *d_bc, const uint32_t *d_gene,
                                      const uint32_t *d_umi, uint64_t n, int umi_len, int max_dist,
                                      uint32_t *d_rep_umi, uint64_t *d_n_groups, uint32_t *d_g_bc,
                                      uint32_t *d_g_gene, uint32_t *d_g_umi, uint32_t *d_g_reads,
                                      void *d_workspace, size_t workspace_bytes, void *stream)
{
    return nr_umi_collapse_device_keyed(d_bc, d_gene, d_umi, n, umi_len, max_dist, 32, 32, 32, d_rep_umi,
                                        d_n_groups, d_g_bc, d_g_gene, d_g_umi, d_g_reads, d_workspace,
                                        workspace_bytes, stream);
}
