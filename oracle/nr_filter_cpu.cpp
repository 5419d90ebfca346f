// nr_filter_cpu.cpp -- multi-threaded CPU build of the SAME algorithm the GPU path runs (lossless
// seed filter + exact verification for cost <= 2, reads with one or two N as substituted variants),
// compiled from the very header the sm_100a kernel includes (nr_filter_core.h).
//
// BENCH / TEST INFRASTRUCTURE, like everything under oracle/: bench.py times it as
// `cpu_baseline_filtered` (kind "port-filtered") so that the GPU number can be read against the
// same algorithm on the host cores, next to the brute-force oracle (nr_oracle.c); tests compare
// it with the oracle.  Nothing in nanoranger_b200/ links or calls it.
//
// Replaces, for that comparison, scripts/barcode_align.sh:14-41 (STAR, absent here) restricted to
// the score range the reference keeps (AS >= 14: utils.py:699, 845, 1150, 1479); candidates whose
// best score is lower are reported as score -128 / idx -1, as NR_MODE_FILTERED does.
#include <algorithm>
#include <atomic>
#include <cstdint>
#include <cstring>
#include <thread>
#include <vector>

#include "../nanoranger_b200/csrc/nr_filter_core.h"

namespace {

struct Index {
    int64_t n = 0;
    int padL = 0, padR = 0;
    std::vector<uint32_t> bits[4], rank[4], kstart[4];
    std::vector<uint32_t> ent[4], core[4];           // rows sorted by key: entry, packed core
};

void build(const uint32_t *wl, int64_t n, Index &ix)
{
    ix.n = n;
    std::vector<std::pair<uint32_t, uint32_t>> rows((size_t)n);
    for (int j = 0; j < 4; j++) {
        for (int64_t e = 0; e < n; e++) rows[(size_t)e] = {nr_core_key(wl[e], j), (uint32_t)e};
        std::sort(rows.begin(), rows.end());
        ix.bits[j].assign((1u << 19) + 1, 0u);
        ix.rank[j].assign((1u << 19) + 1, 0u);
        ix.ent[j].resize((size_t)n);
        ix.core[j].resize((size_t)n);
        ix.kstart[j].clear();
        std::vector<uint32_t> hs((size_t)n);
        uint32_t nd = 0;
        for (size_t i = 0; i < rows.size(); i++) {
            ix.bits[j][rows[i].first >> 5] |= 1u << (rows[i].first & 31u);
            if (i == 0 || rows[i - 1].first != rows[i].first) { ix.kstart[j].push_back((uint32_t)i); nd++; }
            hs[i] = nd;
            ix.ent[j][i] = rows[i].second;
            ix.core[j][i] = wl[rows[i].second];
        }
        ix.kstart[j].push_back((uint32_t)n);
        size_t r = 0;
        for (uint32_t w = 0; w <= (1u << 19); w++) {
            while (r < rows.size() && (uint64_t)rows[r].first < ((uint64_t)w << 5)) r++;
            ix.rank[j][w] = r == rows.size() ? nd : hs[r] - 1;
        }
    }
}

struct Found { uint32_t key; int umi; };

struct Acc {
    int best = 3;
    std::vector<Found> pairs;          // distinct (entry << 1 | strand) at `best`
    void add(int cost, uint32_t key, int umi)
    {
        if (cost > best) return;
        if (cost < best) { best = cost; pairs.clear(); }
        for (auto &f : pairs)
            if (f.key == key) { f.umi = f.umi < 0 ? umi : (umi < 0 ? f.umi : std::min(f.umi, umi)); return; }
        pairs.push_back({key, umi});
    }
};

// one probe stage over the slots of both strands of the read staged in rdp
void run_stage(const Index &ix, const uint32_t *wl, const uint32_t rdp[2][NR_RDP_WORDS], int m,
               int stage, Acc &acc, bool nmode, const uint64_t nms[2], int v, int n0, int n1)
{
    static const int lo[3] = {0, NR_PROBES_COST0, NR_PROBES_COST1};
    static const int hi[3] = {NR_PROBES_COST0, NR_PROBES_COST1, NR_PROBES_MAIN};
    const int p0 = nr_slot_first(m, ix.padR), p1 = nr_slot_last(m, ix.padL);
    for (int s = 0; s < 2; s++)
        for (int p = p0; p <= p1; p++) {
            if (nmode && v != 0 &&
                !nr_nvar_slot_needed(v, p, s ? m - 1 - n0 : n0, n1 < 0 ? -100 : (s ? m - 1 - n1 : n1)))
                continue;
            const uint64_t W = nr_window64(rdp[s], p);
            for (int pass = 0; pass < 2; pass++) {
                int t0 = lo[stage], t1 = hi[stage];
                if (pass == 1) {
                    if (!(stage == 2 && p == -1)) break;
                    t0 = NR_PROBES_MAIN; t1 = NR_PROBES_ALL;
                }
                for (int t = t0; t < t1; t++) {
                    const nr_probe_t &pr = NR_PROBES[t];
                    if (p + nr_probe_first(pr) < 0 || p + nr_probe_end(pr) > m) continue;
                    const uint32_t key = nr_probe_key(W, pr);
                    const int d = pr.drop;
                    const uint32_t bw = ix.bits[d][key >> 5];
                    if (!((bw >> (key & 31u)) & 1u)) continue;
                    const size_t kr = ix.rank[d][key >> 5] + (size_t)nr_popc32(bw & ((1u << (key & 31u)) - 1u));
                    for (size_t r = ix.kstart[d][kr]; r < ix.kstart[d][kr + 1]; r++) {
                        int u;
                        const int cost = nmode
                            ? nr_verify16n(rdp[s], nms[s], m, ix.core[d][r], ix.padL, ix.padR, p, pr, &u)
                            : nr_verify16(rdp[s], m, ix.core[d][r], ix.padL, ix.padR, p, pr, &u);
                        if (cost <= 2) acc.add(cost, (ix.ent[d][r] << 1) | (uint32_t)s, u);
                    }
                }
            }
        }
    (void)wl;
}

void stage_read(const uint32_t w[4], int m, uint32_t rdp[2][NR_RDP_WORDS])
{
    uint32_t rc[4];
    nr_revcomp4(w, m, rc);
    memset(rdp, 0, sizeof(uint32_t) * 2 * NR_RDP_WORDS);
    for (int k = 0; k < 4; k++) { rdp[0][1 + k] = w[k]; rdp[1][1 + k] = rc[k]; }
}

void one(const Index &ix, const uint32_t *wl, const uint8_t *q, int m, int min_len, int32_t *idx,
         int8_t *score, int32_t *nbest, uint8_t *strand, int16_t *umi, uint8_t *took)
{
    *idx = -1; *score = -128; *nbest = 0; *strand = 0; *umi = -1; *took = 0;
    if (m > 64 || m < min_len) return;
    uint64_t nm = 0;
    uint32_t w0[4] = {0, 0, 0, 0};
    for (int i = 0; i < m; i++) {
        if (q[i] > 3) nm |= 1ull << i;
        else w0[i >> 4] |= (uint32_t)q[i] << ((i & 15) * 2);
    }
    const int n_n = __builtin_popcountll(nm);
    if (n_n > 2) return;
    *took = 1;
    Acc acc;
    uint32_t rdp[2][NR_RDP_WORDS];
    if (n_n == 0) {
        stage_read(w0, m, rdp);
        for (int stage = 0; stage < 3 && acc.best >= stage; stage++)
            run_stage(ix, wl, rdp, m, stage, acc, false, nullptr, 0, 0, -1);
    } else {
        const int n0 = __builtin_ctzll(nm), n1 = n_n == 2 ? 63 - __builtin_clzll(nm) : -1;
        const uint64_t nms[2] = {nm, nr_rev_mask(nm, m)};
        for (int round = 0; round < 3 && acc.best >= round; round++)
            for (int v = 0; v < nr_nvar_count(n_n); v++) {
                const int stage = round - nr_nvar_nonzero(v);
                if (stage < 0) continue;
                uint32_t wv[4];
                nr_nvar_apply(w0, n0, n1, v, wv);
                stage_read(wv, m, rdp);
                run_stage(ix, wl, rdp, m, stage, acc, true, nms, v, n0, n1);
            }
    }
    if (acc.best > 2) return;
    uint32_t bk = 0xFFFFFFFFu;
    int bu = -1;
    for (auto &f : acc.pairs)
        if (f.key < bk) { bk = f.key; bu = f.umi; }
    *idx = (int32_t)(bk >> 1); *strand = (uint8_t)(bk & 1u);
    *score = (int8_t)(16 - acc.best); *nbest = (int32_t)acc.pairs.size();
    *umi = (int16_t)((bk & 1u) ? -1 : bu);
}

}  // namespace

extern "C" {

struct nr_cpu_filter {
    Index ix;
    std::vector<uint32_t> wl;
};

// wl: n packed 16-column cores (2 bit/base, column k at bit 2k)
nr_cpu_filter *nr_cpu_filter_create(const uint32_t *wl, int64_t n, int padL, int padR)
{
    nr_cpu_filter *h = new nr_cpu_filter();
    h->wl.assign(wl, wl + n);
    h->ix.padL = padL; h->ix.padR = padR;
    build(h->wl.data(), n, h->ix);
    return h;
}

void nr_cpu_filter_destroy(nr_cpu_filter *h) { delete h; }

// cand: N x 64 codes (0..3, 4 = N), clen: N.  took[i] = 0 for candidates the filter does not take
// (shorter than min_len, more than two N): the GPU path hands those to its deep tier.
int nr_cpu_filter_match(const nr_cpu_filter *h, const uint8_t *cand, const uint8_t *clen, int64_t N,
                        int min_len, int threads, int32_t *idx, int8_t *score, int32_t *nbest,
                        uint8_t *strand, int16_t *umi, uint8_t *took)
{
    if (!h || threads < 1) return -1;
    std::atomic<int64_t> next{0};
    auto work = [&]() {
        for (;;) {
            const int64_t c0 = next.fetch_add(256);
            if (c0 >= N) break;
            const int64_t c1 = std::min(N, c0 + 256);
            for (int64_t c = c0; c < c1; c++)
                one(h->ix, h->wl.data(), cand + (size_t)c * 64, clen[c], min_len, idx + c, score + c,
                    nbest + c, strand + c, umi + c, took + c);
        }
    };
    std::vector<std::thread> th;
    for (int t = 1; t < threads; t++) th.emplace_back(work);
    work();
    for (auto &t : th) t.join();
    return 0;
}

}  // extern "C"
