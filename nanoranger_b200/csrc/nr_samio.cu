// nr_samio.cu -- host-side writer of `{sample}_matching.sam` (no device code in here).
//
// Replaces the output stage of scripts/barcode_align.sh:14-41 (STAR --outSAMtype SAM,
// --outSAMmode NoQS, unique mappers only, unmapped reads absent): one record per candidate whose
// best score is reached by exactly one (entry, strand) pair, with the fields
// utils.process_matching_* consume (QNAME FLAG RNAME POS CIGAR SEQ AS:i; utils.py:697-708).
// Same bytes as nanoranger_b200/samio.py:write_sam + utils.match_records produce (tests compare
// the two); this one formats ~1e7 records/s instead of ~2e5.
#include <algorithm>
#include <atomic>
#include <cstdio>
#include <cstring>
#include <string>
#include <thread>
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

inline int code_of(char c)
{
    switch (c) {
    case 'A': case 'a': return 0; case 'C': case 'c': return 1;
    case 'G': case 'g': return 2; case 'T': case 't': return 3;
    default: return 4;
    }
}

// One record's alignment against its padded reference N^padL + core + N^padR, by traceback of
// the score DP the kernels evaluate (scripts/barcode_align.sh:18-33): query
// global, reference local, match +1, mismatch -1, gap -1 per base, N on either side 0.  Among
// the optimal alignments the one that leaves the core at the smallest read row is traced (that
// row is the kernel's umi_q); when no optimal alignment reaches reference column padL+L, the best
// one that ends inside the core or left of it.  Ties inside the matrix: diagonal before insertion
// before deletion.  a.leave = that row, -1 if none (the caller checks it against umi_q).
struct Aln {
    int64_t pos = 1;                  // 1-based leftmost reference position
    std::string cigar, md;
    int nm = 0;                       // mismatches (both bases ACGT and different)
    int score = 0;
    int leave = -1;
};

void align_record(const char *q, int m, const uint32_t lo, const uint32_t hi, const uint32_t nmk,
                  int L, int padL, int padR, Aln &a)
{
    int S[NR_MAX_QUERY + 1][NR_MAX_CORE + 1];
    uint8_t qc[NR_MAX_QUERY];
    int8_t cc[NR_MAX_CORE];
    for (int i = 0; i < m; i++) qc[i] = (uint8_t)code_of(q[i]);
    for (int j = 0; j < L; j++)
        cc[j] = ((nmk >> j) & 1u) ? 4 : (int8_t)(((j < 16) ? (lo >> (2 * j)) : (hi >> (2 * (j - 16)))) & 3u);
    auto sc = [&](int i, int j) { return (qc[i - 1] > 3 || cc[j - 1] > 3) ? 0 : (qc[i - 1] == cc[j - 1] ? 1 : -1); };
    for (int j = 0; j <= L; j++) S[0][j] = 0;
    for (int i = 1; i <= m; i++) {
        S[i][0] = -std::max(0, i - padL);
        for (int j = 1; j <= L; j++)
            S[i][j] = std::max(S[i - 1][j - 1] + sc(i, j), std::max(S[i - 1][j], S[i][j - 1]) - 1);
    }
    // where the alignment leaves the scored columns (the rule of the kernels / the oracle: through
    // column L at the smallest row if that is optimal, else inside the core, else left of it)
    int a_r = -std::max(0, m - padR), arg = 0;
    for (int i = 1; i <= m; i++) {
        const int v = S[i][L] - std::max(0, m - i - padR);
        if (v > a_r) { a_r = v; arg = i; }
    }
    int a_in = -1000, jin = 0;
    for (int j = 1; j < L; j++)
        if (S[m][j] > a_in) { a_in = S[m][j]; jin = j; }
    const int a_left = -std::max(0, m - padL);
    const int as = std::max(a_left, std::max(a_in, a_r));
    int ei, ej;
    if (a_r == as) { ei = arg; ej = L; a.leave = arg; }
    else if (a_in == as) { ei = m; ej = jin; a.leave = -1; }
    else { ei = m; ej = 0; a.leave = -1; }
    a.score = as;
    // operations, last to first: 'M' diagonal, 'I' read base without a column, 'D' column without
    // a read base; pads are appended below
    std::string ops;
    int i = ei, j = ej;
    while (i > 0 && j > 0) {
        if (S[i][j] == S[i - 1][j - 1] + sc(i, j)) { ops.push_back('M'); i--; j--; }
        else if (S[i][j] == S[i - 1][j] - 1) { ops.push_back('I'); i--; }
        else { ops.push_back('D'); j--; }
    }
    // i read bases left of the core: the last min(i, padL) of them pair with pad columns, the rest
    // are insertions before the first aligned column; j > 0 only when the alignment starts inside
    // the core (i == 0)
    const int lead_m = std::min(i, padL), lead_i = i - lead_m;
    a.pos = (int64_t)padL + j - lead_m + 1;
    std::string full;
    full.append((size_t)lead_i, 'I');
    full.append((size_t)lead_m, 'P');               // P: M against a pad column
    full.append(ops.rbegin(), ops.rend());
    if (ej == L) {
        const int rest = m - ei, tail_m = std::min(rest, padR);
        full.append((size_t)tail_m, 'P');
        full.append((size_t)(rest - tail_m), 'I');
    }
    // CIGAR, nM, MD
    a.cigar.clear(); a.md.clear(); a.nm = 0;
    int run = 0, qi = 0, cj = j, mdrun = 0;
    char prev = 0;
    bool in_del = false;
    auto flush = [&](char op) {
        if (run) { a.cigar += std::to_string(run); a.cigar.push_back(op == 'P' ? 'M' : op); }
        run = 0;
    };
    for (char op : full) {
        const char cop = op == 'P' ? 'M' : op;
        if (prev && cop != (prev == 'P' ? 'M' : prev)) flush(prev);
        run++;
        prev = op;
        if (op == 'P') { mdrun++; qi++; in_del = false; }
        else if (op == 'M') {
            const int qb = qc[qi], rb = cc[cj];
            if (qb > 3 || rb > 3 || qb == rb) mdrun++;
            else { a.nm++; a.md += std::to_string(mdrun); a.md.push_back("ACGT"[rb]); mdrun = 0; }
            qi++; cj++; in_del = false;
        } else if (op == 'I') { qi++; }
        else {                                      // D
            if (!in_del) { a.md += std::to_string(mdrun); a.md.push_back('^'); mdrun = 0; in_del = true; }
            a.md.push_back(cc[cj] > 3 ? 'N' : "ACGT"[cc[cj]]);
            cj++;
        }
    }
    flush(prev);
    a.md += std::to_string(mdrun);
    if (a.cigar.empty()) a.cigar = "*";
}

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

// Same file contract, with REAL alignments: POS / CIGAR from a traceback of every kept record and
// the attributes scripts/barcode_align.sh:21 asks STAR for, in that order: AS:i nM:i MD:Z.
extern "C" int nr_sam_write_aligned(const nr_whitelist_t *wl, const char *path, int header_full,
                                    const char *names, const uint64_t *name_off, const char *seqs,
                                    const uint64_t *seq_off, uint64_t n, const int32_t *idx,
                                    const int8_t *score, const uint8_t *nbest, const uint8_t *flags,
                                    const uint8_t *umi_q, const char *ref_names,
                                    const uint64_t *ref_off, uint64_t n_ref, int threads,
                                    uint64_t *n_written)
{
    if (!wl || !path || !name_off || !seq_off || !ref_off || !n_written || (n && (!names || !seqs ||
        !idx || !score || !nbest || !flags || !umi_q)) || (n_ref && !ref_names)) {
        nr_set_error("nr_sam_write_aligned: null pointer");
        return NR_EINVAL;
    }
    if (n_ref != wl->n) {
        nr_set_error("nr_sam_write_aligned: %llu reference names for %llu whitelist entries",
                     (unsigned long long)n_ref, (unsigned long long)wl->n);
        return NR_EINVAL;
    }
    std::vector<uint64_t> keep;
    keep.reserve((size_t)n);
    for (uint64_t i = 0; i < n; i++)
        if (nbest[i] == 1 && !(flags[i] & NR_FLAG_TOO_LONG)) {
            if (idx[i] < 0 || (uint64_t)idx[i] >= n_ref) {
                nr_set_error("nr_sam_write_aligned: candidate %llu refers to entry %d of %llu",
                             (unsigned long long)i, idx[i], (unsigned long long)n_ref);
                return NR_EINVAL;
            }
            keep.push_back(i);
        }
    FILE *fp = fopen(path, "wb");
    if (!fp) { nr_set_error("nr_sam_write_aligned: cannot open %s", path); return NR_EINVAL; }
    Out o(fp);
    const int L = (int)wl->L, padL = (int)wl->pad_l, padR = (int)wl->pad_r;
    const uint64_t ref_len = (uint64_t)padL + L + padR;
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
    o.put("@CO\tuser command line: nanoranger_b200.utils.barcode_align (alignments by traceback)\n");
    o.flush();
    // records: formatted by `threads` workers into their own buffers, written in order
    unsigned nt = threads > 0 ? (unsigned)threads : std::max(1u, std::thread::hardware_concurrency());
    nt = (unsigned)std::min<size_t>(nt, std::max<size_t>(1, keep.size() / 4096));
    std::vector<std::string> bufs(nt);
    std::atomic<uint64_t> mismatch{0};
    auto work = [&](unsigned t) {
        const size_t a0 = keep.size() * t / nt, a1 = keep.size() * (t + 1) / nt;
        std::string &b = bufs[t];
        b.reserve((a1 - a0) * 200);
        std::string rc;
        Aln al;
        for (size_t k = a0; k < a1; k++) {
            const uint64_t i = keep[k];
            const char *s = seqs + seq_off[i];
            const int m = (int)(seq_off[i + 1] - seq_off[i]);
            const bool rev = (flags[i] & NR_FLAG_RC) != 0;
            const char *q = s;
            if (rev) {
                rc.resize((size_t)m);
                for (int x = 0; x < m; x++) rc[(size_t)x] = comp(s[m - 1 - x]);
                q = rc.data();
            }
            b.append(names + name_off[i], (size_t)(name_off[i + 1] - name_off[i]));
            b += rev ? "\t16\t" : "\t0\t";
            b.append(ref_names + ref_off[idx[i]], (size_t)(ref_off[idx[i] + 1] - ref_off[idx[i]]));
            if (m == 0) {
                b += "\t1\t255\t*\t*\t0\t0\t*\t*\tAS:i:0\tnM:i:0\tMD:Z:0\n";
                continue;
            }
            const int e = idx[i];
            align_record(q, m, wl->h_lo[e], wl->h_hi ? wl->h_hi[e] : 0u, wl->h_nm ? wl->h_nm[e] : 0u, L,
                         padL, padR, al);
            // the traceback re-derives what the kernels reported: score, and for forward records
            // the row at which the alignment leaves the core
            if (al.score != (int)score[i] ||
                (!rev && al.leave != (umi_q[i] == NR_UMI_NONE ? -1 : (int)umi_q[i])))
                mismatch.fetch_add(1);
            b.push_back('\t'); b += std::to_string(al.pos); b += "\t255\t"; b += al.cigar;
            b += "\t*\t0\t0\t"; b.append(q, (size_t)m);
            b += "\t*\tAS:i:"; b += std::to_string((int)score[i]);
            b += "\tnM:i:"; b += std::to_string(al.nm);
            b += "\tMD:Z:"; b += al.md;
            b.push_back('\n');
        }
    };
    std::vector<std::thread> th;
    for (unsigned t = 1; t < nt; t++) th.emplace_back(work, t);
    work(0);
    for (auto &t : th) t.join();
    for (auto &b : bufs) fwrite(b.data(), 1, b.size(), fp);
    const bool bad = ferror(fp) != 0;
    fclose(fp);
    if (bad) { nr_set_error("nr_sam_write_aligned: write to %s failed", path); return NR_EINVAL; }
    if (mismatch.load()) {
        nr_set_error("nr_sam_write_aligned: %llu tracebacks disagree with the matcher's score / UMI column",
                     (unsigned long long)mismatch.load());
        return NR_EINVAL;
    }
    *n_written = (uint64_t)keep.size();
    return NR_OK;
}
