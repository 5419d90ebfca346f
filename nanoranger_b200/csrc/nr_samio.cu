// nr_samio.cu -- host-side writer of `{sample}_matching.sam` (no device code in here).
//
// Replaces the output stage of scripts/barcode_align.sh:14-41 (STAR --outSAMtype SAM,
// --outSAMmode NoQS, unique mappers only, unmapped reads absent): one record per candidate whose
// best score is reached by exactly one (entry, strand) pair, with the fields
// utils.process_matching_* consume (QNAME FLAG RNAME POS CIGAR SEQ AS:i; utils.py:697-708).
// Same bytes as nanoranger_b200/samio.py:write_sam + utils.match_records produce (tests compare
// the two); this one formats ~1e7 records/s instead of ~2e5.
#include <algorithm>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include "nr_common.cuh"

namespace {

struct Out {
    FILE *f;
    std::vector<char> buf;
    size_t n = 0;
    explicit Out(FILE *fp) : f(fp), buf(8u << 20) {}
    void need(size_t k) { if (n + k > buf.size()) flush(); }
    void flush() { if (n) { fwrite(buf.data(), 1, n, f); n = 0; } }
    void put(const char *p, size_t k)
    {
        if (k > buf.size() / 2) { flush(); fwrite(p, 1, k, f); return; }
        need(k); memcpy(buf.data() + n, p, k); n += k;
    }
    void put(const char *z) { put(z, strlen(z)); }
    void put_uint(uint64_t v)
    {
        char t[24]; int k = 0;
        do { t[k++] = (char)('0' + v % 10); v /= 10; } while (v);
        need((size_t)k);
        while (k) buf[n++] = t[--k];
    }
    void put_int(int64_t v) { if (v < 0) { put("-", 1); put_uint((uint64_t)(-v)); } else put_uint((uint64_t)v); }
};

inline char comp(char c)
{
    switch (c) {
    case 'A': return 'T'; case 'C': return 'G'; case 'G': return 'C'; case 'T': return 'A';
    case 'a': return 't'; case 'c': return 'g'; case 'g': return 'c'; case 't': return 'a';
    default: return c;       // N, n and anything else map to themselves
    }
}

}  // namespace

extern "C" int nr_sam_write(const char *path, int header_full, const char *names,
                            const uint64_t *name_off, const char *seqs, const uint64_t *seq_off,
                            uint64_t n, const int32_t *idx, const int8_t *score,
                            const uint8_t *nbest, const uint8_t *flags, const uint8_t *umi_q,
                            const char *ref_names, const uint64_t *ref_off, uint64_t n_ref,
                            uint32_t pad_l, uint32_t core_len, uint32_t pad_r, uint64_t *n_written)
{
    if (!path || !name_off || !seq_off || !ref_off || !n_written || (n && (!names || !seqs || !idx ||
        !score || !nbest || !flags || !umi_q)) || (n_ref && !ref_names)) {
        nr_set_error("nr_sam_write: null pointer");
        return NR_EINVAL;
    }
    // records kept: unique best pair, not too long (STAR: --outFilterMultimapNmax 1,
    // --outSAMunmapped None)
    std::vector<uint64_t> keep;
    keep.reserve((size_t)n);
    for (uint64_t i = 0; i < n; i++)
        if (nbest[i] == 1 && !(flags[i] & NR_FLAG_TOO_LONG)) {
            if (idx[i] < 0 || (uint64_t)idx[i] >= n_ref) {
                nr_set_error("nr_sam_write: candidate %llu refers to entry %d of %llu",
                             (unsigned long long)i, idx[i], (unsigned long long)n_ref);
                return NR_EINVAL;
            }
            keep.push_back(i);
        }
    FILE *fp = fopen(path, "wb");
    if (!fp) { nr_set_error("nr_sam_write: cannot open %s", path); return NR_EINVAL; }
    Out o(fp);
    const uint64_t ref_len = (uint64_t)pad_l + core_len + pad_r;
    o.put("@HD\tVN:1.4\n");
    auto sq = [&](uint64_t r) {
        o.put("@SQ\tSN:");
        o.put(ref_names + ref_off[r], (size_t)(ref_off[r + 1] - ref_off[r]));
        o.put("\tLN:");
        o.put_uint(ref_len);
        o.put("\n");
    };
    if (header_full) {
        for (uint64_t r = 0; r < n_ref; r++) sq(r);
    } else {
        std::vector<int32_t> used;
        used.reserve(keep.size());
        for (uint64_t i : keep) used.push_back(idx[i]);
        std::sort(used.begin(), used.end());
        used.erase(std::unique(used.begin(), used.end()), used.end());
        for (int32_t r : used) sq((uint64_t)r);
    }
    o.put("@PG\tID:nanoranger_b200\tPN:nanoranger_b200\n");
    o.put("@CO\tuser command line: nanoranger_b200.utils.barcode_align\n");
    std::string rc;
    for (uint64_t i : keep) {
        const char *s = seqs + seq_off[i];
        const int64_t m = (int64_t)(seq_off[i + 1] - seq_off[i]);
        const bool rev = (flags[i] & NR_FLAG_RC) != 0;
        const int64_t u = (umi_q[i] == NR_UMI_NONE || rev) ? -1 : (int64_t)umi_q[i];
        o.put(names + name_off[i], (size_t)(name_off[i + 1] - name_off[i]));
        o.put(rev ? "\t16\t" : "\t0\t");
        o.put(ref_names + ref_off[idx[i]], (size_t)(ref_off[idx[i] + 1] - ref_off[idx[i]]));
        o.put("\t");
        // anchored placement: reference column pad_l + core_len pairs with read base u (the read
        // ends one column before it when no optimal alignment reaches that column)
        if (m == 0) {
            o.put("1\t255\t*");
        } else {
            const int64_t umi_col = (int64_t)pad_l + core_len;
            const int64_t start = umi_col - (u < 0 ? m : u);
            const int64_t lead = std::max<int64_t>(0, -start);
            const int64_t trail = std::max<int64_t>(0, start + m - (int64_t)ref_len);
            const int64_t mid = m - lead - trail;
            if (mid <= 0) {
                o.put("1\t255\t");
                o.put_int(m); o.put("I");
            } else {
                o.put_int(std::max<int64_t>(start, 0) + 1);
                o.put("\t255\t");
                if (lead) { o.put_int(lead); o.put("I"); }
                o.put_int(mid); o.put("M");
                if (trail) { o.put_int(trail); o.put("I"); }
            }
        }
        o.put("\t*\t0\t0\t");
        if (rev) {
            rc.resize((size_t)m);
            for (int64_t k = 0; k < m; k++) rc[(size_t)k] = comp(s[m - 1 - k]);
            o.put(rc.data(), (size_t)m);
        } else {
            o.put(s, (size_t)m);
        }
        o.put("\t*\tNH:i:1\tHI:i:1\tAS:i:");
        o.put_int((int64_t)score[i]);
        o.put("\n");
    }
    o.flush();
    const bool bad = ferror(fp) != 0;
    fclose(fp);
    if (bad) { nr_set_error("nr_sam_write: write to %s failed", path); return NR_EINVAL; }
    *n_written = (uint64_t)keep.size();
    return NR_OK;
}
