// nr_anchor_index.h -- host-side builder of the anchored filter's index (nr_anchor_core.h).
// Plain C++: included by nr_whitelist.cu and by the CPU emulation in tests/emul/.
//
// Finds the split P | K | S of the cores: K = the longest run of columns that carry the same
// (non-N) base in every entry, P / S = what is left of it on either side.  The filter applies when
// |P| = 8 (the scripts of nr_anchor_core.h enumerate 8-column neighbourhoods; direct-address table
// of 4^8 keys) and 12 <= |K| <= 28 (the linker walk keeps 32 read bases in one word): the slide-seq
// geometry.
// Table: start[key] .. start[key + 1] delimit the rows (entry indices) whose P part equals `key`;
// an entry with N in P appears under each base it stands for.
#pragma once
#include <algorithm>
#ifndef NR_ANCHOR_LP
#define NR_ANCHOR_LP 8
#endif
#include <cstdint>
#include <vector>

struct nr_anchor_index_host {
    int ok = 0;
    int Lp = 0, Lk = 0, Ls = 0;
    uint64_t link = 0;                   // K, column k at bits 2k
    std::vector<uint32_t> start;         // 4^Lp + 1
    std::vector<uint32_t> rows;          // entry indices, grouped by key
};

inline void nr_anchor_index_build(const uint32_t *lo, const uint32_t *hi, const uint32_t *nm, uint64_t n,
                                  int L, nr_anchor_index_host &out)
{
    out.ok = 0;
    if (n == 0 || L < 14) return;
    auto col = [&](uint64_t e, int j) {
        return (int)(((j < 16) ? (lo[e] >> (2 * j)) : ((hi ? hi[e] : 0u) >> (2 * (j - 16)))) & 3u);
    };
    auto isn = [&](uint64_t e, int j) { return nm ? (int)((nm[e] >> j) & 1u) : 0; };
    // constant columns
    std::vector<char> konst((size_t)L, 1);
    for (int j = 0; j < L; j++) {
        if (isn(0, j)) { konst[(size_t)j] = 0; continue; }
        const int c0 = col(0, j);
        for (uint64_t e = 1; e < n && konst[(size_t)j]; e++)
            if (isn(e, j) || col(e, j) != c0) konst[(size_t)j] = 0;
    }
    int best_a = 0, best_len = 0;
    for (int j = 0; j < L;) {
        if (!konst[(size_t)j]) { j++; continue; }
        int k = j;
        while (k < L && konst[(size_t)k]) k++;
        if (k - j > best_len) { best_len = k - j; best_a = j; }
        j = k;
    }
    const int Lp = best_a, Lk = best_len, Ls = L - Lp - Lk;
    if (Lp != NR_ANCHOR_LP || Lk < 12 || Lk > 28 || Ls < 0) return;
    out.Lp = Lp; out.Lk = Lk; out.Ls = Ls;
    out.link = 0;
    for (int j = 0; j < Lk; j++) out.link |= (uint64_t)col(0, Lp + j) << (2 * j);
    const uint32_t nkeys = 1u << (2 * Lp);
    std::vector<std::pair<uint32_t, uint32_t>> kv;          // (key, entry)
    kv.reserve((size_t)n + n / 4);
    for (uint64_t e = 0; e < n; e++) {
        uint32_t base = 0, nmask = 0;
        for (int j = 0; j < Lp; j++) {
            if (isn(e, j)) nmask |= 1u << j;
            else base |= (uint32_t)col(e, j) << (2 * j);
        }
        // expand the N columns of P to the four bases
        std::vector<uint32_t> keys{base};
        for (int j = 0; j < Lp; j++)
            if ((nmask >> j) & 1u) {
                std::vector<uint32_t> nx;
                for (uint32_t k : keys)
                    for (uint32_t x = 0; x < 4; x++) nx.push_back(k | (x << (2 * j)));
                keys.swap(nx);
            }
        for (uint32_t k : keys) kv.push_back({k, (uint32_t)e});
    }
    std::sort(kv.begin(), kv.end());
    out.start.assign((size_t)nkeys + 1, 0u);
    out.rows.resize(kv.size());
    for (size_t i = 0; i < kv.size(); i++) { out.start[kv[i].first + 1]++; out.rows[i] = kv[i].second; }
    for (uint32_t k = 0; k < nkeys; k++) out.start[k + 1] += out.start[k];
    out.ok = 1;
}
