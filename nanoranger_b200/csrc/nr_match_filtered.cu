// nr_match_filtered.cu -- lossless seed filter + exact verification for cost <= 2 (AS >= 14).
//
// Replaces scripts/barcode_align.sh:14-41 (STAR EndToEnd against the N-padded whitelist, unique
// mappers only) for the score range the reference keeps (AS >= 14: utils.py:699, 845, 1150,
// 1479).  The arithmetic (probe set, exact automaton) is in nr_filter_core.h and is the same
// code the CPU emulation in tests/emul/ checks against the oracle.
//
// Mapping.  One warp owns a tile of 32 consecutive candidates: lane l loads candidate l's
// 128-bit packed record (one coalesced 512 B request per warp) and stores its 8 B of results;
// the candidates of the tile are then resolved one after another by the whole warp:
//   probe    lane = (strand, slot position); the probes of a slot are unrolled with compile-time
//            offsets; each is one read of the L2-resident key bitmaps (one base pointer, three
//            bit orders so that probe families share a 32 B sector; 256-bit loads for the
//            "5-mer minus one base" triples); the hit bits of a lane are collected in a mask
//   stages   the probe table is walked in three stages over ALL slots -- probe 0 (complete for
//            cost 0), probes 1..15 (complete for cost <= 1), the two-event probes -- and the
//            queue is drained after each; a candidate whose best pair is already at a cost the
//            finished stages are complete for stops there
//   queue    one ballot (or one scan when a lane holds several hits) places the hits in a
//            per-warp shared-memory queue as (probe, strand, slot)
//   verify   up to 32 queued hits are expanded into the index rows sharing their keys (key rank =
//            rank word + popc of the bitmap word below the key; prefix sum over the row counts);
//            lane = one row: reads {entry, core} and scores it with nr_verify16: a furthest-
//            reaching-diagonal walk from the end of the core the probe pins (exact in the read
//            interior), the 3-level shift-and automaton over the <= 27 read rows around the
//            slot where the core may hang over a read end
//   merge    best cost, the distinct (entry, strand) pairs attaining it (kept one per lane),
//            smallest UMI row per pair
// Candidates the filter cannot take (contain N, shorter than NR_FILTER_MIN_LEN, more than 32
// co-optimal pairs) are appended to a device list that the exhaustive kernel resolves.
#include <atomic>
#include <type_traits>
#include <utility>

#include "nr_common.cuh"
#include "nr_filter_core.h"

#define NR_FILTER_MIN_LEN 24
#define NR_FWARPS 8                        // warps per block
#ifndef NR_FBLOCKS
#define NR_FBLOCKS 4                       // resident blocks per SM the register budget is set for
#endif
#define NR_QCAP 480                        // queue slots per warp
#define NR_VQCAP 64                        // N pass: rows waiting for the N-aware automaton, per warp

struct nr_filter_params {
    const uint32_t *bits[4];      // key bitmap, 2^19 words per dropped quarter, contiguous:
                                  // bits[j] == bits[0] + j * 2^19 (nr_whitelist.cu)
    const uint32_t *rank[4];      // distinct keys below each bitmap word
    const uint2 *ents[4];
    const uint32_t *kstart[4];
    uint32_t n;
    int padL, padR;
    const uint4 *bases;
    const uint8_t *meta;
    uint64_t n_cand;
    int min_score;
    int resolve_below;      // 1: candidates without a pair at cost <= 2 go to the list as well
    int32_t *o_idx;
    int8_t *o_score;
    uint8_t *o_nbest;
    uint8_t *o_flags;
    uint8_t *o_umi;
    uint32_t *list;         // candidates left to the next tier (deep / exhaustive)
    uint32_t *list_count;
    uint32_t *list_n;       // main pass: candidates with N, handed to the N pass (nullable)
    uint32_t *list_n_count;
    const uint64_t *nmask;  // N pass: per-candidate N masks
    unsigned long long *tile_next;  // next tile to hand out (zeroed with the workspace header)
    unsigned long long *counters;   // nullable: probes, hits, verifications, passes, listed
};

namespace {

// probe table as packed words in shared memory: the hits a warp expands carry lane-varying probe
// numbers, and a lane-varying index into __constant__ memory is serialised by the address
// divergence unit (ncu: pipe_adu at 52 % of peak with the table in constant memory)
__device__ __forceinline__ uint32_t probe_pack(const nr_probe_t &t)
{
    return (uint32_t)t.drop | ((uint32_t)t.o0 << 2) | ((uint32_t)t.o1 << 6) | ((uint32_t)t.o2 << 10) |
           ((uint32_t)(t.var + 1) << 14) | ((uint32_t)t.del << 16);
}
__device__ __forceinline__ nr_probe_t probe_unpack(uint32_t w)
{
    nr_probe_t t;
    t.drop = (int8_t)(w & 3u); t.o0 = (int8_t)((w >> 2) & 15u); t.o1 = (int8_t)((w >> 6) & 15u);
    t.o2 = (int8_t)((w >> 10) & 15u); t.var = (int8_t)((int)((w >> 14) & 3u) - 1);
    t.del = (int8_t)((w >> 16) & 3u);
    return t;
}

// the four index tables' base pointers, copied to shared memory for the same reason (the
// dropped quarter of a queued hit is lane-varying; kernel parameters live in constant memory)
struct Tables {
    const uint32_t *bits[4];
    const uint32_t *rank[4];
    const uint2 *ents[4];
    const uint32_t *kstart[4];
};

struct WarpSmem {
    uint4 tile[32];                  // packed records of the warp's 32 candidates
    uint32_t rdp[2][NR_RDP_WORDS];   // candidate in flight: forward / reverse complement, padded
    uint32_t queue[NR_QCAP];         // bitmap hits waiting for verification: probe | strand | slot
};
struct WarpSmemN : WarpSmem {        // the N pass carries more (kept out of the main pass: L1 space)
    uint64_t nm[2];                  // N mask of the candidate in flight, per strand
    uint4 vq[NR_VQCAP];              // rows that passed the wildcard walk: entry, core, strand | slot
};
template <bool NMODE> using WarpSmemT = std::conditional_t<NMODE, WarpSmemN, WarpSmem>;

struct Acc {                 // running answer of the candidate in flight (warp-uniform unless noted)
    int best;                // best cost so far (3 = none)
    int nb;                  // distinct pairs at `best`
    uint32_t key;            // per lane: pair (entry << 1 | strand) held by this lane (lane < nb)
    int umi;                 // per lane: smallest leaving row of that pair, -1 none
    int overflow;            // more than 32 pairs at `best`
    int qn;                  // queued hits
    int vn;                  // N pass: rows waiting in vq
    unsigned long long c_hits, c_ver, c_pass;   // per lane partial counters
};

// the whole probe table in constant memory, for the places that index it dynamically
__device__ __constant__ nr_probe_t c_probes[NR_PROBES_ALL];

__device__ __forceinline__ int merge_umi(int a, int b) { return a < 0 ? b : (b < 0 ? a : min(a, b)); }

// fold one batch of verified rows (per lane: cost, pair, UMI row) into the running answer
__device__ __forceinline__ void merge_batch(Acc &acc, int cost, uint32_t k, int u)
{
    const uint32_t lane = nr_lane();
    const int rb = __reduce_min_sync(0xffffffffu, cost);
    if (rb < 3 && rb <= acc.best) {
        if (rb < acc.best) { acc.best = rb; acc.nb = 0; }
        uint32_t contrib = __ballot_sync(0xffffffffu, cost == rb);
        while (contrib) {
            const int src = __ffs(contrib) - 1;
            contrib &= contrib - 1;
            const uint32_t kk = __shfl_sync(0xffffffffu, k, src);
            const int uu = __shfl_sync(0xffffffffu, u, src);
            const uint32_t found = __ballot_sync(0xffffffffu, (int)lane < acc.nb && acc.key == kk);
            if (found) {
                if ((int)lane == __ffs(found) - 1) acc.umi = merge_umi(acc.umi, uu);
            } else if (acc.nb < 32) {
                if ((int)lane == acc.nb) { acc.key = kk; acc.umi = uu; }
                acc.nb++;
            } else {
                acc.overflow = 1;
            }
        }
    }
}

// N pass: score up to 32 waiting rows with the N-aware automaton, one row per lane.  The automaton
// costs ~800 instructions per row and only a lane or two of a verification batch survive the
// wildcard walk: scoring them where they arise would run it at 2 active lanes (ncu), so survivors
// are parked in vq and scored a full warp at a time.
template <bool COUNT>
__device__ __forceinline__ void flush_vq(const nr_filter_params &P, WarpSmemN &sm, Acc &acc, int m)
{
    const uint32_t lane = nr_lane();
    while (acc.vn > 0) {
        const int cnt = min(32, acc.vn);
        const int base = acc.vn - cnt;
        acc.vn = base;
        int cost = 3, u = -1;
        uint32_t k = 0;
        if ((int)lane < cnt) {
            const uint4 it = sm.vq[base + lane];
            const int strand = (int)(it.z & 1u);
            const int p = (int)(it.z >> 1) - 16;
            cost = nr_score16n(sm.rdp[strand], sm.nm[strand], m, it.y, P.padL, P.padR, p, &u);
            k = (it.x << 1) | (uint32_t)strand;
            if (COUNT) acc.c_pass += cost < 3;
        }
        __syncwarp();
        merge_batch(acc, cost, k, u);
    }
}

// Take up to 32 queued bitmap hits, expand each into the index rows that share its key
// (kstart gives first row and count), and verify the rows 32 at a time, one row per lane.
template <bool COUNT, bool NMODE>
__device__ __forceinline__ void drain(const nr_filter_params &P, WarpSmemT<NMODE> &sm, Acc &acc, int m,
                                      const uint32_t *s_probes, const Tables &T4)
{
    const uint32_t lane = nr_lane();
    const int cnt = min(32, acc.qn);
    const int base = acc.qn - cnt;
    acc.qn = base;
    const bool have = (int)lane < cnt;
    const uint32_t item = have ? sm.queue[base + lane] : 0u;
    __syncwarp();
    // the hit's key again (cheap, and now 32 hits wide), its rank among the distinct keys of the
    // table, and through kstart its rows; probes reaching outside the read nominate nothing
    const nr_probe_t t = probe_unpack(s_probes[(item >> 16) & 63u]);
    const int h_strand = (int)((item >> 24) & 1u);
    const int h_p = (int)(item >> 25) - 16;
    const uint32_t d = (uint32_t)t.drop;
    uint32_t start = 0, rows = 0;
    if (have && h_p + nr_probe_first(t) >= 0 && h_p + nr_probe_end(t) <= m) {
        const uint32_t key = nr_probe_key(nr_window64(sm.rdp[h_strand], h_p), t);
        const uint32_t w = __ldg(T4.bits[0] + (((d << 24) | key) >> 5));
        // rank / kstart / rows are touched once per hit: keep them out of L1 (L2 only) so that the
        // bitmap sectors stay
        const uint32_t kr = __ldcg(T4.rank[d] + (key >> 5)) +
                            (uint32_t)__popc(w & ((1u << (key & 31u)) - 1u));
        const uint32_t *ks = T4.kstart[d] + kr;
        start = __ldcg(ks);
        rows = __ldcg(ks + 1) - start;
        if (COUNT) acc.c_hits++;
    }
    // exclusive prefix of the row counts
    uint32_t incl = rows;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        uint32_t v = __shfl_up_sync(0xffffffffu, incl, o);
        if ((int)lane >= o) incl += v;
    }
    const uint32_t total = __shfl_sync(0xffffffffu, incl, 31);
    const uint32_t excl = incl - rows;
    for (uint32_t b = 0; b < total; b += 32) {
        const uint32_t g = b + lane;
        const bool active = g < total;
        // owner = last lane whose exclusive prefix is <= g (lanes without rows never win: their
        // prefix equals that of the next lane with rows, which is preferred by the search order)
        int lo = 0;
#pragma unroll
        for (int step = 16; step > 0; step >>= 1) {
            uint32_t v = __shfl_sync(0xffffffffu, excl, (lo + step) & 31);
            if (lo + step < cnt && v <= g) lo += step;
        }
        const uint32_t o_start = __shfl_sync(0xffffffffu, start, lo);
        const uint32_t o_excl = __shfl_sync(0xffffffffu, excl, lo);
        const uint32_t o_item = __shfl_sync(0xffffffffu, item, lo);
        int cost = 3, u = -1;
        uint32_t k = 0, vcore = 0, vwhere = 0;
        if (active) {
            const nr_probe_t ot = probe_unpack(s_probes[(o_item >> 16) & 63u]);
            const uint32_t od = (uint32_t)ot.drop;
            const int strand = (int)((o_item >> 24) & 1u);
            const int p = (int)(o_item >> 25) - 16;
            const uint2 e = __ldcg(T4.ents[od] + o_start + (g - o_excl));
            if constexpr (NMODE) {
                // reads with N: diagonal walk with the N rows as wildcards here (sm.rdp holds a
                // substituted variant; the N rows ignore their base), the N-aware automaton later
                cost = nr_prefilter16n(sm.rdp[strand], sm.nm[strand], m, e.y, p, ot) ? 0 : 3;
                vcore = e.y;
                vwhere = (uint32_t)strand | ((uint32_t)(p + 16) << 1);
            } else {
                cost = nr_verify16(sm.rdp[strand], m, e.y, P.padL, P.padR, p, ot, &u);
            }
            k = (e.x << 1) | (uint32_t)strand;
            if (COUNT) { acc.c_ver++; if (!NMODE) acc.c_pass += cost < 3; }
        }
        if constexpr (NMODE) {
            // park the survivors of the wildcard walk; score them a full warp at a time
            const uint32_t bal = __ballot_sync(0xffffffffu, cost < 3);
            if (bal) {
                if (cost < 3)
                    sm.vq[acc.vn + __popc(bal & ((1u << lane) - 1u))] = make_uint4(k >> 1, vcore, vwhere, 0u);
                acc.vn += __popc(bal);
                __syncwarp();
                if (acc.vn > NR_VQCAP - 32) flush_vq<COUNT>(P, sm, acc, m);
            }
        } else {
            merge_batch(acc, cost, k, u);
        }
    }
}

// The three "5-mer minus one base" variants of a probe differ in one kept quarter only; in the
// bitmap layout that puts this quarter in the key's low byte (nr_filter_core.h) they address the
// same 32-byte sector: the first variant brings the whole sector with one 256-bit load and all
// three test their bit in registers.
// (ncu: the kernel is L1TEX-bound -- l1tex throughput 73 %, ALU pipe 58 % -- and the tag stage
// works per sector.)
__host__ __device__ constexpr bool probe_family_first(int T)
{
    return T >= NR_PROBES_COST1 && T < NR_PROBES_MAIN && (T - NR_PROBES_COST1) % 3 == 0;
}
__host__ __device__ constexpr bool probe_family_rest(int T)
{
    return T >= NR_PROBES_COST1 && T < NR_PROBES_MAIN && (T - NR_PROBES_COST1) % 3 != 0;
}
__host__ __device__ constexpr bool probe_family_ok(int T)   // T, T+1, T+2 = del 1, 2, 3 of one probe
{
    return NR_PROBES[T].var >= 0 && NR_PROBES[T].del == 1 && NR_PROBES[T + 1].del == 2 &&
           NR_PROBES[T + 2].del == 3 && NR_PROBES[T].var == NR_PROBES[T + 2].var &&
           NR_PROBES[T].drop == NR_PROBES[T + 2].drop && NR_PROBES[T].o0 == NR_PROBES[T + 2].o0 &&
           NR_PROBES[T].o1 == NR_PROBES[T + 2].o1 && NR_PROBES[T].o2 == NR_PROBES[T + 2].o2;
}
static_assert(NR_PROBES_MAIN - NR_PROBES_COST1 == 18 && probe_family_ok(16) && probe_family_ok(19) &&
              probe_family_ok(22) && probe_family_ok(25) && probe_family_ok(28) && probe_family_ok(31),
              "probe families out of step with NR_PROBES");

struct Sector { uint32_t v[8]; };

__device__ __forceinline__ uint32_t sector_bit(const Sector &s, uint32_t a)   // a = 0..255
{
    const uint32_t lo = (a & 64u) ? ((a & 32u) ? s.v[3] : s.v[2]) : ((a & 32u) ? s.v[1] : s.v[0]);
    const uint32_t hi = (a & 64u) ? ((a & 32u) ? s.v[7] : s.v[6]) : ((a & 32u) ? s.v[5] : s.v[4]);
    return (((a & 128u) ? hi : lo) >> (a & 31u)) & 1u;
}

// one probe of a slot: key from the window, one read of the key bitmap, hit bit into `mask`.
// `bits` is the base of the four contiguous bitmaps; lanes without a slot carry W = 0 and read a
// valid word whose result the caller discards (no predicated loads in the unrolled sequence).
template <int T>
__device__ __forceinline__ void probe_one(const uint32_t *__restrict__ bits, uint64_t W, uint64_t &mask)
{
    constexpr nr_probe_t t = NR_PROBES[T];
    constexpr int layout = nr_probe_layout_of(t);
    constexpr uint32_t table = (uint32_t)(layout * 4 + t.drop) << 24;
    if constexpr (probe_family_rest(T)) {
        return;                                   // tested by the first variant of its family
    } else if constexpr (probe_family_first(T)) {
        const uint32_t key = nr_key_layout(nr_probe_key(W, t), layout) | table;
        constexpr int off = t.var == 0 ? t.o0 : (t.var == 1 ? t.o1 : t.o2);   // the varying quarter
        Sector s;
        const uint32_t *sp = bits + ((key >> 8) << 3);
        asm volatile("ld.global.nc.v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                     : "=r"(s.v[0]), "=r"(s.v[1]), "=r"(s.v[2]), "=r"(s.v[3]), "=r"(s.v[4]),
                       "=r"(s.v[5]), "=r"(s.v[6]), "=r"(s.v[7])
                     : "l"(sp));
        mask |= (uint64_t)sector_bit(s, key & 255u) << T;
        mask |= (uint64_t)sector_bit(s, nr_quarter(W, off, 2)) << (T + 1);
        mask |= (uint64_t)sector_bit(s, nr_quarter(W, off, 3)) << (T + 2);
    } else {
        const uint32_t key = nr_key_layout(nr_probe_key(W, t), layout) | table;
        const uint32_t w = __ldg(bits + (key >> 5));
        mask |= (uint64_t)((w >> (key & 31u)) & 1u) << T;
    }
}

// probes LO .. LO + N - 1 of a slot
template <int LO, int... I>
__device__ __forceinline__ uint64_t probe_range(const uint32_t *__restrict__ bits, uint64_t W,
                                                bool slot_ok, std::integer_sequence<int, I...>)
{
    uint64_t mask = 0;
    (probe_one<LO + I>(bits, W, mask), ...);
    return slot_ok ? mask : 0ull;
}

// Place the hits of one work item (per-lane probe mask at (strand, p)) in the warp's queue,
// draining it as often as needed.
template <bool COUNT, bool NMODE>
__device__ __forceinline__ void enqueue(const nr_filter_params &P, WarpSmemT<NMODE> &sm, Acc &acc, int m,
                                        const uint32_t *s_probes, const Tables &T4, uint64_t mask,
                                        int strand, int p)
{
    const uint32_t lane = nr_lane();
    // an item with more hits than the queue holds (dense key bitmaps: whitelists of millions of
    // entries) is queued in four probe ranges of <= 9 x 32 hits
    const uint32_t where = ((uint32_t)strand << 24) | ((uint32_t)(p + 16) << 25);
    const int mine_all = __popcll(mask);
    // common case: no lane holds more than one hit -> positions from one ballot
    if (!__any_sync(0xffffffffu, mine_all > 1)) {
        const uint32_t b = __ballot_sync(0xffffffffu, mine_all != 0);
        if (b == 0u) return;
        const int total = __popc(b);
        while (acc.qn + total > NR_QCAP) drain<COUNT, NMODE>(P, sm, acc, m, s_probes, T4);
        if (mine_all)
            sm.queue[acc.qn + __popc(b & ((1u << lane) - 1u))] =
                where | ((uint32_t)(__ffsll((long long)mask) - 1) << 16);
        acc.qn += total;
        __syncwarp();
        return;
    }
    const int total_all = __reduce_add_sync(0xffffffffu, mine_all);
    const int nparts = total_all > NR_QCAP ? 4 : 1;
#pragma unroll 1
    for (int part = 0; part < nparts; part++) {
        uint64_t pm = nparts == 1 ? mask : (mask & (0x1FFull << (9 * part)));
        // queue positions of the hits: exclusive scan of the per-lane counts
        const int mine = __popcll(pm);
        int incl = mine;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            int v = __shfl_up_sync(0xffffffffu, incl, o);
            if ((int)lane >= o) incl += v;
        }
        const int total = __shfl_sync(0xffffffffu, incl, 31);
        while (acc.qn + total > NR_QCAP) drain<COUNT, NMODE>(P, sm, acc, m, s_probes, T4);
        int pos = acc.qn + incl - mine;
        // queue item = (probe, strand, slot position); the key, its rank and its rows are worked
        // out in drain(), one hit per lane
        while (pm) {
            const int T = __ffsll((long long)pm) - 1;
            pm &= pm - 1;
            sm.queue[pos++] = where | ((uint32_t)T << 16);
        }
        acc.qn += total;
        __syncwarp();
    }
}

// One probe stage (0: probe 0; 1: probes 1..15; 2: the two-event probes + the edge probes at slot
// -1) over all slots of both strands of the read staged in sm.rdp, drained at the end.
// NMODE: the staged read is substituted variant `v`; slots that cannot reach a substituted N
// position are skipped (nr_filter_core.h).
template <bool COUNT, bool NMODE>
__device__ __forceinline__ void run_stage(const nr_filter_params &P, WarpSmemT<NMODE> &sm, Acc &acc, int m,
                                          const uint32_t *s_probes, const Tables &s_tab,
                                          const uint32_t *__restrict__ bits_all, int stage, int p0,
                                          int nP, bool edge, int v, int n0, int n1,
                                          unsigned long long &c_probes_n)
{
    const uint32_t lane = nr_lane();
    const int nslots = nP > 0 ? 2 * nP : 0;
    const int nchunks = (nslots + 31) >> 5;
#pragma unroll 1
    for (int item = 0; item < nchunks; item++) {
        const int slot = item * 32 + (int)lane;
        bool slot_ok = slot < nslots;
        const int strand = slot >= nP ? 1 : 0;
        const int p = p0 + slot - strand * nP;
        if (NMODE && v != 0)
            slot_ok = slot_ok && nr_nvar_slot_needed(v, p, strand ? m - 1 - n0 : n0,
                                                     n1 < 0 ? -100 : (strand ? m - 1 - n1 : n1));
        if (NMODE && !__any_sync(0xffffffffu, slot_ok)) continue;
        const uint64_t W = slot_ok ? nr_window64(sm.rdp[strand], p) : 0ull;
        uint64_t mask;
        if (stage == 0) {
            mask = probe_range<0>(bits_all, W, slot_ok, std::make_integer_sequence<int, NR_PROBES_COST0>{});
            if (COUNT) c_probes_n += slot_ok ? NR_PROBES_COST0 : 0;
        } else if (stage == 1) {
            mask = probe_range<NR_PROBES_COST0>(bits_all, W, slot_ok, std::make_integer_sequence<int, NR_PROBES_COST1 - NR_PROBES_COST0>{});
            if (COUNT) c_probes_n += slot_ok ? NR_PROBES_COST1 - NR_PROBES_COST0 : 0;
        } else {
            mask = probe_range<NR_PROBES_COST1>(bits_all, W, slot_ok, std::make_integer_sequence<int, NR_PROBES_MAIN - NR_PROBES_COST1>{});
            if (COUNT) c_probes_n += slot_ok ? NR_PROBES_MAIN - NR_PROBES_COST1 : 0;
        }
        enqueue<COUNT, NMODE>(P, sm, acc, m, s_probes, s_tab, mask, strand, p);
    }
    if (stage == 2 && edge) {
        // one-column start overhang + interior insertion (slot -1 only)
        uint64_t mask = 0;
        const int strand = lane >= NR_PROBES_EDGE ? 1 : 0;
        if (lane < 2 * NR_PROBES_EDGE) {
            const int ti = NR_PROBES_MAIN + (int)lane - strand * NR_PROBES_EDGE;
            const uint64_t W = nr_window64(sm.rdp[strand], -1);
            const nr_probe_t t = probe_unpack(s_probes[ti]);
            const uint32_t key = nr_probe_key(W, t);
            const uint32_t w = __ldg(P.bits[0] + (key >> 5));
            mask = (uint64_t)((w >> (key & 31u)) & 1u) << ti;
            if (COUNT) c_probes_n++;
        }
        enqueue<COUNT, NMODE>(P, sm, acc, m, s_probes, s_tab, mask, strand, -1);
    }
    while (acc.qn > 0) drain<COUNT, NMODE>(P, sm, acc, m, s_probes, s_tab);
}

// stage both strands of the packed read w4, padded, in shared memory
__device__ __forceinline__ void stage_read(WarpSmem &sm, const uint32_t w4[4], int m)
{
    const uint32_t lane = nr_lane();
    uint32_t rc[4];
    nr_revcomp4(w4, m, rc);
    uint32_t vf = 0, vr = 0;
#pragma unroll
    for (int k = 0; k < 4; k++)
        if ((int)lane == k + 1) { vf = w4[k]; vr = rc[k]; }
    __syncwarp();
    if (lane < NR_RDP_WORDS) { sm.rdp[0][lane] = vf; sm.rdp[1][lane] = vr; }
    __syncwarp();
}

// NMODE = false: the main pass over all n_cand candidates; reads with N (and at least
// NR_FILTER_MIN_LEN bases) are handed to the N pass through list_n when there is one.
// NMODE = true: the N pass over list_n -- reads with one or two N as substituted variants
// (nr_filter_core.h), reads with more N on to the next tier.
template <bool COUNT, bool NMODE>
__global__ void __launch_bounds__(NR_FWARPS * 32, NR_FBLOCKS)
nr_match_filtered_kernel(const nr_filter_params P)
{
    __shared__ WarpSmemT<NMODE> smem[NR_FWARPS];
    __shared__ uint32_t s_probes[64];
    __shared__ Tables s_tab;
    if (threadIdx.x < 64)
        s_probes[threadIdx.x] = threadIdx.x < NR_PROBES_ALL ? probe_pack(c_probes[threadIdx.x]) : 0u;
    if (threadIdx.x < 4) {
        s_tab.bits[threadIdx.x] = P.bits[threadIdx.x]; s_tab.rank[threadIdx.x] = P.rank[threadIdx.x];
        s_tab.ents[threadIdx.x] = P.ents[threadIdx.x]; s_tab.kstart[threadIdx.x] = P.kstart[threadIdx.x];
    }
    __syncthreads();
    const uint32_t lane = nr_lane();
    const int warp = threadIdx.x >> 5;
    WarpSmemT<NMODE> &sm = smem[warp];
    const uint32_t *__restrict__ bits_all = P.bits[0];
    // candidates per tile: the N pass gets few reads (a few per cent of the batch), each several
    // times dearer than a plain read: small tiles keep its tail short
    constexpr uint32_t TS = NMODE ? 2u : 32u;
    const uint64_t n_items = NMODE ? (uint64_t)*P.list_n_count : P.n_cand;
    const uint64_t n_tiles = (n_items + TS - 1) / TS;
    unsigned long long c_probes_n = 0, c_listed = 0;
    Acc acc;
    acc.c_hits = acc.c_ver = acc.c_pass = 0;

    // tiles are handed out dynamically: their cost varies several-fold (a tile of unassignable
    // candidates runs the whole probe table 32 times), a static stride leaves a long tail
    for (;;) {
        unsigned long long t64 = 0;
        if (lane == 0) t64 = atomicAdd(P.tile_next, 1ull);
        const uint64_t tile = ((uint64_t)__shfl_sync(0xffffffffu, (uint32_t)(t64 >> 32), 0) << 32) |
                              (uint64_t)__shfl_sync(0xffffffffu, (uint32_t)t64, 0);
        if (tile >= n_tiles) break;
        // one coalesced 512 B request brings the tile's records; meta bytes stay in registers
        uint32_t mt = 0x100u;   // no candidate
        uint32_t cidx = 0;      // N pass: the candidate this lane loaded
        uint32_t nm_lo = 0, nm_hi = 0;
        {
            const uint64_t mine = tile * TS + lane;
            uint4 b = make_uint4(0u, 0u, 0u, 0u);
            if (lane < TS && mine < n_items) {
                if (NMODE) {
                    cidx = P.list_n[mine];
                    b = __ldg(P.bases + cidx); mt = P.meta[cidx];
                    const uint64_t nm = P.nmask[cidx];
                    nm_lo = (uint32_t)nm; nm_hi = (uint32_t)(nm >> 32);
                } else {
                    b = __ldcs(P.bases + mine); mt = P.meta[mine];
                }
            }
            __syncwarp();
            sm.tile[lane] = b;
            __syncwarp();
        }
        const int in_tile = (int)min((uint64_t)TS, n_items - tile * TS);

#pragma unroll 1
        for (int c = 0; c < in_tile; c++) {
            const uint32_t cmt = __shfl_sync(0xffffffffu, mt, c);
            const uint64_t cand = NMODE ? (uint64_t)__shfl_sync(0xffffffffu, cidx, c) : tile * TS + c;
            if (cmt == 0xFFu) {          // longer than NR_MAX_QUERY: not scored
                if (lane == 0) {
                    P.o_idx[cand] = -1; P.o_score[cand] = NR_SCORE_BELOW; P.o_nbest[cand] = 0;
                    P.o_flags[cand] = NR_FLAG_TOO_LONG | NR_FLAG_BELOW | NR_FLAG_NO_UMI;
                    P.o_umi[cand] = NR_UMI_NONE;
                }
                continue;
            }
            const int m = (int)(cmt & 0x7Fu);
            bool to_list = m < NR_FILTER_MIN_LEN;
            bool to_nlist = false;
            uint64_t nm = 0;
            int n_n = 0;
            if (NMODE) {
                nm = ((uint64_t)__shfl_sync(0xffffffffu, nm_hi, c) << 32) |
                     (uint64_t)__shfl_sync(0xffffffffu, nm_lo, c);
                n_n = __popcll(nm);
                if (n_n < 1 || n_n > 2) to_list = true;
            } else if (cmt & 0x80u) {
                if (P.list_n && !to_list) to_nlist = true; else to_list = true;
            }
            if (!to_list && !to_nlist) {
                acc.best = 3; acc.nb = 0; acc.key = 0; acc.umi = -1; acc.overflow = 0; acc.qn = 0; acc.vn = 0;
                const int p0 = nr_slot_first(m, P.padR), p1 = nr_slot_last(m, P.padL);
                const int nP = p1 - p0 + 1;
                const bool edge = p0 <= -1 && p1 >= -1;
                const uint4 t4 = sm.tile[c];
                const uint32_t w4[4] = {t4.x, t4.y, t4.z, t4.w};
                if constexpr (!NMODE) {
                    stage_read(sm, w4, m);
                    // Three stages over all slots, each drained before the next starts.  Prefixes of
                    // the probe table are complete for small costs (nr_filter_core.h): after stage 0
                    // (probe 0) every placement of cost 0 is known, after stage 1 (probes 1..15)
                    // every placement of cost <= 1; a candidate whose best pair is already that good
                    // needs nothing more -- only placements at the best cost count.  Stage 2 runs
                    // the variant probes and the edge probes (two cost-1 events).
#pragma unroll 1
                    for (int stage = 0; stage < 3; stage++) {
                        if (acc.best < stage) break;
                        run_stage<COUNT, false>(P, sm, acc, m, s_probes, s_tab, bits_all, stage, p0,
                                                nP, edge, 0, 0, -1, c_probes_n);
                    }
                } else {
                    // rounds: after round r every placement of true cost <= r is known; variant v
                    // (z non-zero substitutions) runs probe stage r - z
                    const int n0 = __ffsll((long long)nm) - 1;
                    const int n1 = n_n == 2 ? 63 - __clzll((long long)nm) : -1;
                    if (lane == 0) { sm.nm[0] = nm; sm.nm[1] = __brevll(nm) >> (64 - m); }
                    const int nvar = nr_nvar_count(n_n);
#pragma unroll 1
                    for (int round = 0; round < 3; round++) {
                        if (acc.best < round) break;
#pragma unroll 1
                        for (int v = 0; v < nvar; v++) {
                            const int stage = round - nr_nvar_nonzero(v);
                            if (stage < 0) continue;
                            uint32_t wv[4];
                            nr_nvar_apply(w4, n0, n1, v, wv);
                            stage_read(sm, wv, m);
                            run_stage<COUNT, true>(P, sm, acc, m, s_probes, s_tab, bits_all, stage, p0,
                                                   nP, edge, v, n0, n1, c_probes_n);
                        }
                        // the rows parked by all variants of the round (any variant's bases serve:
                        // variants differ at the N rows only, and those ignore their base)
                        flush_vq<COUNT>(P, sm, acc, m);
                    }
                }

                if (acc.overflow || (acc.best == 3 && P.resolve_below)) {
                    to_list = true;
                } else if (acc.best == 3) {
                    if (lane == 0) {
                        P.o_idx[cand] = -1; P.o_score[cand] = NR_SCORE_BELOW; P.o_nbest[cand] = 0;
                        P.o_flags[cand] = NR_FLAG_BELOW | NR_FLAG_NO_UMI; P.o_umi[cand] = NR_UMI_NONE;
                    }
                } else {
                    uint32_t kmine = (int)lane < acc.nb ? acc.key : 0xFFFFFFFFu;
                    uint32_t kmin = __reduce_min_sync(0xffffffffu, kmine);
                    uint32_t holder = __ballot_sync(0xffffffffu, kmine == kmin);
                    int u = __shfl_sync(0xffffffffu, acc.umi, __ffs(holder) - 1);
                    if (lane == 0) {
                        int strand = (int)(kmin & 1u);
                        int score = 16 - acc.best;
                        uint32_t fl = 0;
                        if (acc.nb > 1) fl |= NR_FLAG_TIE;
                        if (strand) { fl |= NR_FLAG_RC; u = -1; }
                        if (score < P.min_score) fl |= NR_FLAG_BELOW;
                        if (u < 0) fl |= NR_FLAG_NO_UMI;
                        P.o_idx[cand] = (int32_t)(kmin >> 1); P.o_score[cand] = (int8_t)score;
                        P.o_nbest[cand] = (uint8_t)acc.nb; P.o_flags[cand] = (uint8_t)fl;
                        P.o_umi[cand] = (uint8_t)(u < 0 ? NR_UMI_NONE : u);
                    }
                }
            }
            if (to_nlist) {
                if (lane == 0) P.list_n[atomicAdd(P.list_n_count, 1u)] = (uint32_t)cand;
            } else if (to_list) {
                if (lane == 0) {
                    uint32_t at = atomicAdd(P.list_count, 1u);
                    P.list[at] = (uint32_t)cand;
                }
                if (COUNT) c_listed += lane == 0;
            }
        }
    }
    if (COUNT && P.counters) {
        atomicAdd(P.counters + 0, c_probes_n);
        atomicAdd(P.counters + 1, acc.c_hits);
        atomicAdd(P.counters + 2, acc.c_ver);
        atomicAdd(P.counters + 3, acc.c_pass);
        atomicAdd(P.counters + 4, c_listed);
    }
}

}  // namespace

// devices whose copy of c_probes has been written (bit per device ordinal)
static std::atomic<unsigned long long> g_probes_uploaded{0ull};

// Enqueue the filtered matcher on `stream`: the main pass over all candidates, then (when
// d_list_n is given) the N pass over the reads with N the main pass set aside.  The workspace
// header (list counts, tile counters) must have been zeroed on the stream.
int nr_launch_filtered(const nr_whitelist *wl, const void *d_bases, const uint8_t *d_meta,
                       const uint64_t *d_nmask, uint64_t n_cand, int min_score, int resolve_below,
                       int32_t *d_idx, int8_t *d_score, uint8_t *d_nbest, uint8_t *d_flags,
                       uint8_t *d_umi, uint32_t *d_list, uint32_t *d_list_count,
                       uint32_t *d_list_n, uint32_t *d_list_n_count,
                       unsigned long long *d_tile_next, unsigned long long *d_tile_next_n,
                       unsigned long long *d_counters, int *grid_out, cudaStream_t stream)
{
    if (n_cand == 0) return NR_OK;
    if (!wl->has_index) {
        nr_set_error("filtered matcher needs a 16-column N-free whitelist");
        return NR_EUNSUPPORTED;
    }
    const unsigned long long dev_bit = 1ull << (wl->device & 63);
    if (!(g_probes_uploaded.load(std::memory_order_acquire) & dev_bit)) {
        // a concurrent first call on the same device writes the same bytes: harmless
        NR_CHECK_CUDA(cudaMemcpyToSymbol(c_probes, &NR_PROBES[0],
                                         sizeof(nr_probe_t) * NR_PROBES_ALL));
        g_probes_uploaded.fetch_or(dev_bit, std::memory_order_release);
    }
    nr_filter_params P;
    for (int j = 0; j < 4; j++) {
        P.bits[j] = wl->d_bits[j]; P.rank[j] = wl->d_rank[j];
        P.ents[j] = wl->d_ents[j]; P.kstart[j] = wl->d_kstart[j];
    }
    P.n = (uint32_t)wl->n; P.padL = (int)wl->pad_l; P.padR = (int)wl->pad_r;
    P.bases = (const uint4 *)d_bases; P.meta = d_meta; P.n_cand = n_cand;
    P.min_score = min_score; P.resolve_below = resolve_below;
    P.o_idx = d_idx; P.o_score = d_score; P.o_nbest = d_nbest; P.o_flags = d_flags; P.o_umi = d_umi;
    P.list = d_list; P.list_count = d_list_count; P.counters = d_counters;
    P.list_n = d_list_n; P.list_n_count = d_list_n_count; P.nmask = d_nmask;
    P.tile_next = d_tile_next;
    int sms = 148, per_sm = 1;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, wl->device);
    if (d_counters)
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, nr_match_filtered_kernel<true, false>,
                                                      NR_FWARPS * 32, 0);
    else
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, nr_match_filtered_kernel<false, false>,
                                                      NR_FWARPS * 32, 0);
    if (per_sm < 1) per_sm = 1;
    uint64_t tiles = (n_cand + 31) / 32;
    uint64_t want = (tiles + NR_FWARPS - 1) / NR_FWARPS;
    uint64_t cap = (uint64_t)sms * (uint64_t)per_sm;
    unsigned grid = (unsigned)(want < cap ? want : cap);
    if (grid_out) *grid_out = (int)grid;
    if (d_counters)
        nr_match_filtered_kernel<true, false><<<grid, NR_FWARPS * 32, 0, stream>>>(P);
    else
        nr_match_filtered_kernel<false, false><<<grid, NR_FWARPS * 32, 0, stream>>>(P);
    NR_CHECK_CUDA(cudaGetLastError());
    if (d_list_n) {
        // the N pass: how many reads it gets is only known on the device; a grid sized for one
        // read in eight (far more than any real rate of N) exits at once where there is no tile
        P.tile_next = d_tile_next_n;
        // (each of its reads is a latency chain of several probe stages: the more warps the better)
        uint64_t want_n = (want + 1) / 2;
        unsigned grid_n = (unsigned)(want_n < cap ? (want_n ? want_n : 1) : cap);
        if (d_counters)
            nr_match_filtered_kernel<true, true><<<grid_n, NR_FWARPS * 32, 0, stream>>>(P);
        else
            nr_match_filtered_kernel<false, true><<<grid_n, NR_FWARPS * 32, 0, stream>>>(P);
        NR_CHECK_CUDA(cudaGetLastError());
    }
    return NR_OK;
}
