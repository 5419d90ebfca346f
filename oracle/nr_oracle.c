/*
 * nr_oracle.c -- CPU oracle for the nanoranger barcode-match + UMI-dedup path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under nanoranger_b200/ may import, link or
 * execute this file; only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs use it, and only as the checker or the
 * timed CPU baseline.
 *
 * PARITY UNPINNED at the aligner boundary: the arithmetic of this path lives in
 * STAR (external C++ binary, only hint of a version: "#module load star/2.7.9a",
 * /root/reference/scripts/barcode_align.sh:12), which is neither vendored in the
 * reference nor installed here, and the reference holds no tests, golden SAMs or
 * expected outputs.  This file therefore restates the *scoring the reference
 * configures STAR with* and searches it exhaustively:
 *
 *   scripts/barcode_align.sh:18-33   --alignEndsType EndToEnd, match +1 / mismatch -1
 *                                    (STAR built-ins), --scoreDelOpen 0 --scoreDelBase -1
 *                                    --scoreInsOpen 0 --scoreInsBase -1,
 *                                    --scoreGenomicLengthLog2scale 0, N scores 0,
 *                                    --outFilterMultimapNmax 1 --outFilterMultimapScoreRange 0
 *   utils.py:604-622, 584-601,       reference = "N"*padL + core + "N"*padR
 *            1116-1132, 1451-1453
 *   utils.py:699-718, 845-868,       accept AS >= thr and flag == 0; UMI = query slice at
 *            638-651, 1150-1170,      the query index aligned to reference column padL+L
 *            1479-1504
 *
 * Tiers (each validated against the one above it in tests/test_oracle.py):
 *   tier 0  nr_oracle_as_padded   literal DP over the padded reference
 *   tier 1  nr_oracle_pair        core-only DP with closed-form pad boundaries + UMI column
 *   tier 2  nr_oracle_match       exhaustive scan of all entries x both strands, int8 lanes
 *                                 across entries (auto-vectorised), pthreads over candidates
 *
 * Base codes: A=0 C=1 G=2 T=3 N=4 (anything else is treated as N).
 */
#include <pthread.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define NR_MAXQ 64
#define NR_MAXL 32
#define NR_LANES 64

static inline int imax(int a, int b) { return a > b ? a : b; }

static inline int sc(int q, int r) { return (q > 3 || r > 3) ? 0 : (q == r ? 1 : -1); }

int nr_oracle_code(int ch)
{
    switch (ch) {
    case 'A': case 'a': return 0;
    case 'C': case 'c': return 1;
    case 'G': case 'g': return 2;
    case 'T': case 't': return 3;
    default: return 4;
    }
}

/* tier 0: SURVEY Appendix C verbatim.  q, r are code arrays; r is the *padded* reference.
 *   S[0][j] = 0, S[i][0] = -i, S[i][j] = max(diag + s, up - 1, left - 1), AS = max_j S[m][j] */
int nr_oracle_as_padded(const uint8_t *q, int m, const uint8_t *r, int rl)
{
    int *prev = (int *)calloc((size_t)(rl + 1), sizeof(int));
    int *cur = (int *)calloc((size_t)(rl + 1), sizeof(int));
    for (int j = 0; j <= rl; j++) prev[j] = 0;
    for (int i = 1; i <= m; i++) {
        cur[0] = -i;
        for (int j = 1; j <= rl; j++) {
            int d = prev[j - 1] + sc(q[i - 1], r[j - 1]);
            int u = prev[j] - 1;
            int l = cur[j - 1] - 1;
            cur[j] = imax(d, imax(u, l));
        }
        int *t = prev; prev = cur; cur = t;
    }
    int best = prev[0];
    for (int j = 1; j <= rl; j++) best = imax(best, prev[j]);
    free(prev); free(cur);
    return best;
}

/* tier 1: one (candidate, entry) pair on the core columns only.
 * Left pad closed form:  S[i][padL] = -max(0, i - padL)   (all-N columns score 0).
 * Right pad closed form: from node (i, padL+L) the m-i remaining query bases ride padR free
 *                        diagonals, the excess pays 1 each: tail(i) = -max(0, m-i-padR).
 * AS = max( -max(0,m-padL),  max_{1<=j<L} C_j[m],  max_i C_L[i] + tail(i) ).
 * *umi_q = smallest i with C_L[i] + tail(i) == AS (the query index aligned to reference
 * column padL+L, utils.py:705-708), or -1 when no optimal alignment reaches that column. */
int nr_oracle_pair(const uint8_t *q, int m, const uint8_t *core, int L, int padL, int padR,
                   int *umi_q)
{
    int col[NR_MAXQ + 1], nxt[NR_MAXQ + 1];
    int a_in = -1000;
    for (int i = 0; i <= m; i++) col[i] = -imax(0, i - padL);
    for (int j = 1; j <= L; j++) {
        nxt[0] = 0;
        for (int i = 1; i <= m; i++) {
            int d = col[i - 1] + sc(q[i - 1], core[j - 1]);
            int del = col[i] - 1;
            int ins = nxt[i - 1] - 1;
            nxt[i] = imax(d, imax(del, ins));
        }
        memcpy(col, nxt, sizeof(int) * (size_t)(m + 1));
        if (j < L) a_in = imax(a_in, col[m]);
    }
    int a_right = -1000, arg = -1;
    for (int i = 0; i <= m; i++) {
        int v = col[i] - imax(0, m - i - padR);
        if (v > a_right) { a_right = v; arg = i; }
    }
    int as = imax(-imax(0, m - padL), imax(a_in, a_right));
    if (umi_q) *umi_q = (a_right == as) ? arg : -1;
    return as;
}

/* ------------------------------------------------------------------------------------ */
/* tier 2: exhaustive scan                                                               */

typedef struct {
    const uint8_t *wl_t;   /* transposed whitelist codes: wl_t[j * n_pad + e], n_pad % 64 == 0 */
    int64_t n, n_pad;
    int L, padL, padR;
    const uint8_t *cand;   /* N x 64 codes */
    const uint8_t *clen;   /* N lengths */
    int64_t N;
    int64_t lo, hi;        /* candidate range of this worker */
    int32_t *best_idx; int8_t *best_score; int32_t *n_best; uint8_t *strand; int16_t *umi_q;
    const uint8_t *wl;     /* row-major codes wl[e * L + j] for the tier-1 UMI pass */
} job_t;

/* Score profile of 64 consecutive entries: prof[q][j][lane] = s(q, core_j) as int8. */
typedef struct { int8_t p[5][NR_MAXL][NR_LANES]; } prof_t;

static void build_profile(const uint8_t *wl_t, int64_t n_pad, int64_t e0, int L, prof_t *pf)
{
    for (int q = 0; q < 5; q++)
        for (int j = 0; j < L; j++) {
            const uint8_t *x = wl_t + (size_t)j * (size_t)n_pad + (size_t)e0;
            for (int l = 0; l < NR_LANES; l++)
                pf->p[q][j][l] = (int8_t)((q > 3 || x[l] > 3) ? 0 : (x[l] == q ? 1 : -1));
        }
}

/* AS of one candidate strand against 64 consecutive entries (int8 lanes across entries). */
__attribute__((target_clones("arch=x86-64-v4", "default")))
static void score_block(const prof_t *pf, int L, int padL, int padR, const uint8_t *q, int m,
                        int8_t *out)
{
    int8_t C[NR_MAXL + 1][NR_LANES] __attribute__((aligned(64)));
    int8_t diag[NR_LANES] __attribute__((aligned(64)));
    int8_t a_in[NR_LANES] __attribute__((aligned(64)));
    int8_t a_r[NR_LANES] __attribute__((aligned(64)));
    int t0 = -imax(0, m - padR);
    for (int l = 0; l < NR_LANES; l++) { a_in[l] = -100; a_r[l] = (int8_t)t0; }
    for (int j = 0; j <= L; j++)
        for (int l = 0; l < NR_LANES; l++) C[j][l] = 0;          /* row 0: free start in R */
    for (int i = 1; i <= m; i++) {
        int qi = q[i - 1] > 3 ? 4 : q[i - 1];
        int8_t c0_old = (int8_t)(-imax(0, i - 1 - padL));
        int8_t c0_new = (int8_t)(-imax(0, i - padL));
        for (int l = 0; l < NR_LANES; l++) { diag[l] = c0_old; C[0][l] = c0_new; }
        for (int j = 1; j <= L; j++) {
            const int8_t *sp = pf->p[qi][j - 1];
            int8_t *cj = C[j];
            const int8_t *cl = C[j - 1];
            for (int l = 0; l < NR_LANES; l++) {
                int8_t d = (int8_t)(diag[l] + sp[l]);
                int8_t u = (int8_t)(cj[l] - 1);
                int8_t lf = (int8_t)(cl[l] - 1);
                diag[l] = cj[l];
                int8_t v = d > u ? d : u;
                cj[l] = v > lf ? v : lf;
            }
        }
        int8_t tail = (int8_t)imax(0, m - i - padR);
        const int8_t *cL = C[L];
        for (int l = 0; l < NR_LANES; l++) {
            int8_t v = (int8_t)(cL[l] - tail);
            a_r[l] = a_r[l] > v ? a_r[l] : v;
        }
    }
    for (int j = 1; j < L; j++)
        for (int l = 0; l < NR_LANES; l++) a_in[l] = a_in[l] > C[j][l] ? a_in[l] : C[j][l];
    int8_t a_left = (int8_t)(-imax(0, m - padL));
    for (int l = 0; l < NR_LANES; l++) {
        int8_t v = a_r[l] > a_in[l] ? a_r[l] : a_in[l];
        out[l] = v > a_left ? v : a_left;
    }
}

static void revcomp(const uint8_t *q, int m, uint8_t *rc)
{
    for (int i = 0; i < m; i++) {
        int c = q[m - 1 - i];
        rc[i] = (uint8_t)(c > 3 ? 4 : 3 - c);
    }
}

static inline int clamp_len(int m) { return m > NR_MAXQ ? NR_MAXQ : m; }

/* Entry blocks outermost so each 64-entry score profile is built once per worker; the
 * running (best, count, argmin, strand) of every candidate of the worker lives in the
 * output arrays.  (idx, strand) = smallest entry index among the best pairs, forward
 * preferred when both strands of that entry tie. */
static void *worker(void *arg)
{
    job_t *jb = (job_t *)arg;
    int8_t out[NR_LANES];
    int64_t nc = jb->hi - jb->lo;
    if (nc <= 0) return NULL;
    uint8_t *rcs = (uint8_t *)malloc((size_t)nc * NR_MAXQ);
    int64_t *cnt = (int64_t *)calloc((size_t)nc, sizeof(int64_t));
    prof_t *pf = (prof_t *)aligned_alloc(64, sizeof(prof_t));
    for (int64_t c = jb->lo; c < jb->hi; c++) {
        revcomp(jb->cand + (size_t)c * NR_MAXQ, clamp_len(jb->clen[c]),
                rcs + (size_t)(c - jb->lo) * NR_MAXQ);
        jb->best_score[c] = -127; jb->best_idx[c] = -1; jb->strand[c] = 0;
    }
    for (int64_t e0 = 0; e0 < jb->n_pad; e0 += NR_LANES) {
        build_profile(jb->wl_t, jb->n_pad, e0, jb->L, pf);
        int64_t lim = jb->n - e0 < NR_LANES ? jb->n - e0 : NR_LANES;
        for (int64_t c = jb->lo; c < jb->hi; c++) {
            int m = clamp_len(jb->clen[c]);
            int best = jb->best_score[c]; int64_t idx = jb->best_idx[c]; int st = jb->strand[c];
            int64_t k = cnt[c - jb->lo];
            for (int s = 0; s < 2; s++) {
                const uint8_t *qq = s ? rcs + (size_t)(c - jb->lo) * NR_MAXQ
                                      : jb->cand + (size_t)c * NR_MAXQ;
                score_block(pf, jb->L, jb->padL, jb->padR, qq, m, out);
                for (int l = 0; l < lim; l++) {
                    int v = out[l];
                    if (v > best) { best = v; k = 1; idx = e0 + l; st = s; }
                    else if (v == best) {
                        k++;
                        if (e0 + l < idx) { idx = e0 + l; st = s; }
                    }
                }
            }
            jb->best_score[c] = (int8_t)best; jb->best_idx[c] = (int32_t)idx;
            jb->strand[c] = (uint8_t)st; cnt[c - jb->lo] = k;
        }
    }
    for (int64_t c = jb->lo; c < jb->hi; c++) {
        int64_t k = cnt[c - jb->lo];
        jb->n_best[c] = (int32_t)(k > 0x7fffffff ? 0x7fffffff : k);
        int u = -1;
        if (jb->best_idx[c] >= 0 && jb->strand[c] == 0)
            (void)nr_oracle_pair(jb->cand + (size_t)c * NR_MAXQ, clamp_len(jb->clen[c]),
                                 jb->wl + (size_t)jb->best_idx[c] * (size_t)jb->L, jb->L,
                                 jb->padL, jb->padR, &u);
        jb->umi_q[c] = (int16_t)u;
    }
    free(rcs); free(cnt); free(pf);
    return NULL;
}

/* wl: n x L codes (row-major).  cand: N x 64 codes, clen: N lengths (0..64).
 * Outputs, per candidate:
 *   best_score  max AS over all (entry, strand) pairs
 *   n_best      number of (entry, strand) pairs attaining it (STAR multimapper count)
 *   best_idx    smallest entry index among those pairs
 *   strand      0 if (best_idx, forward) attains best_score, else 1
 *   umi_q       query index aligned to reference column padL+L for (best_idx, forward);
 *               -1 if strand == 1 or no optimal alignment reaches that column
 * Returns 0, or -1 on bad arguments. */
int nr_oracle_match(const uint8_t *wl, int64_t n, int L, int padL, int padR,
                    const uint8_t *cand, const uint8_t *clen, int64_t N, int threads,
                    int32_t *best_idx, int8_t *best_score, int32_t *n_best, uint8_t *strand,
                    int16_t *umi_q)
{
    if (L < 1 || L > NR_MAXL || n < 1 || threads < 1) return -1;
    int64_t n_pad = (n + NR_LANES - 1) / NR_LANES * NR_LANES;
    uint8_t *wl_t = (uint8_t *)aligned_alloc(64, (size_t)n_pad * (size_t)L);
    if (!wl_t) return -1;
    memset(wl_t, 4, (size_t)n_pad * (size_t)L);
    for (int64_t e = 0; e < n; e++)
        for (int j = 0; j < L; j++) wl_t[(size_t)j * (size_t)n_pad + (size_t)e] = wl[e * L + j];
    if (threads > 256) threads = 256;
    if ((int64_t)threads > N) threads = N > 0 ? (int)N : 1;
    pthread_t th[256]; job_t jb[256];
    for (int t = 0; t < threads; t++) {
        jb[t] = (job_t){wl_t, n, n_pad, L, padL, padR, cand, clen, N,
                        N * t / threads, N * (t + 1) / threads,
                        best_idx, best_score, n_best, strand, umi_q, wl};
        pthread_create(&th[t], NULL, worker, &jb[t]);
    }
    for (int t = 0; t < threads; t++) pthread_join(th[t], NULL);
    free(wl_t);
    return 0;
}

/* Scores of one candidate strand against every entry (for targeted tests). */
int nr_oracle_scores(const uint8_t *wl, int64_t n, int L, int padL, int padR,
                     const uint8_t *q, int m, int8_t *scores)
{
    if (L < 1 || L > NR_MAXL || n < 1) return -1;
    int64_t n_pad = (n + NR_LANES - 1) / NR_LANES * NR_LANES;
    uint8_t *wl_t = (uint8_t *)aligned_alloc(64, (size_t)n_pad * (size_t)L);
    if (!wl_t) return -1;
    memset(wl_t, 4, (size_t)n_pad * (size_t)L);
    for (int64_t e = 0; e < n; e++)
        for (int j = 0; j < L; j++) wl_t[(size_t)j * (size_t)n_pad + (size_t)e] = wl[e * L + j];
    int8_t out[NR_LANES];
    prof_t *pf = (prof_t *)aligned_alloc(64, sizeof(prof_t));
    m = clamp_len(m);
    for (int64_t e0 = 0; e0 < n_pad; e0 += NR_LANES) {
        build_profile(wl_t, n_pad, e0, L, pf);
        score_block(pf, L, padL, padR, q, m, out);
        int64_t lim = n - e0 < NR_LANES ? n - e0 : NR_LANES;
        memcpy(scores + e0, out, (size_t)lim);
    }
    free(pf);
    free(wl_t);
    return 0;
}

/* ------------------------------------------------------------------------------------ */
/* UMI dedup oracle: per-barcode exact unique-UMI counts (utils.py:759-773) and the      */
/* directional edit<=1 clustering twin of the CUDA UMI kernel (definition in DESIGN.md). */

typedef struct { uint32_t bc, gene, umi, src; } rec_t;

static int rec_cmp(const void *a, const void *b)
{
    const rec_t *x = (const rec_t *)a, *y = (const rec_t *)b;
    if (x->bc != y->bc) return x->bc < y->bc ? -1 : 1;
    if (x->gene != y->gene) return x->gene < y->gene ? -1 : 1;
    if (x->umi != y->umi) return x->umi < y->umi ? -1 : 1;
    if (x->src != y->src) return x->src < y->src ? -1 : 1;
    return 0;
}

static int hamming2bit(uint32_t a, uint32_t b)
{
    uint32_t x = a ^ b;
    x = (x | (x >> 1)) & 0x55555555u;
    return __builtin_popcount(x);
}

/* Records (bc, gene, umi) with 2-bit packed UMIs (umi_len <= 16).
 * Within each (bc, gene) group: distinct UMIs are ordered by (count desc, umi asc); walking
 * that order, a UMI joins the first earlier *cluster representative* at Hamming distance <= 1
 * whose count is >= 2*count-1 (UMI-tools "directional" rule restricted to one hop), else it
 * founds a new cluster.  cluster_umi[i] = representative UMI of record i.
 * Returns the number of clusters, or -1. */
int64_t nr_oracle_umi_cluster(const uint32_t *bc, const uint32_t *gene, const uint32_t *umi,
                              int64_t n, int max_dist, uint32_t *cluster_umi)
{
    if (n == 0) return 0;
    rec_t *r = (rec_t *)malloc(sizeof(rec_t) * (size_t)n);
    if (!r) return -1;
    for (int64_t i = 0; i < n; i++) r[i] = (rec_t){bc[i], gene[i], umi[i], (uint32_t)i};
    qsort(r, (size_t)n, sizeof(rec_t), rec_cmp);
    int64_t clusters = 0;
    int64_t g0 = 0;
    while (g0 < n) {
        int64_t g1 = g0;
        while (g1 < n && r[g1].bc == r[g0].bc && r[g1].gene == r[g0].gene) g1++;
        /* distinct UMIs of the group */
        int64_t nd = 0;
        for (int64_t i = g0; i < g1; i++) if (i == g0 || r[i].umi != r[i - 1].umi) nd++;
        uint32_t *du = (uint32_t *)malloc(sizeof(uint32_t) * (size_t)nd);
        int64_t *dc = (int64_t *)malloc(sizeof(int64_t) * (size_t)nd);
        int64_t *ord = (int64_t *)malloc(sizeof(int64_t) * (size_t)nd);
        uint32_t *rep = (uint32_t *)malloc(sizeof(uint32_t) * (size_t)nd);
        int64_t k = -1;
        for (int64_t i = g0; i < g1; i++) {
            if (i == g0 || r[i].umi != r[i - 1].umi) { k++; du[k] = r[i].umi; dc[k] = 0; }
            dc[k]++;
        }
        for (int64_t i = 0; i < nd; i++) ord[i] = i;
        /* insertion sort by (count desc, umi asc); du is already ascending */
        for (int64_t i = 1; i < nd; i++) {
            int64_t v = ord[i], j = i - 1;
            while (j >= 0 && dc[ord[j]] < dc[v]) { ord[j + 1] = ord[j]; j--; }
            ord[j + 1] = v;
        }
        int64_t nrep = 0;
        int64_t *repi = (int64_t *)malloc(sizeof(int64_t) * (size_t)nd);
        for (int64_t o = 0; o < nd; o++) {
            int64_t d = ord[o];
            int64_t found = -1;
            if (max_dist > 0)
                for (int64_t p = 0; p < nrep; p++) {
                    int64_t e = repi[p];
                    if (hamming2bit(du[e], du[d]) <= max_dist && dc[e] >= 2 * dc[d] - 1) {
                        found = e; break;
                    }
                }
            if (found < 0) { repi[nrep++] = d; rep[d] = du[d]; }
            else rep[d] = du[found];
        }
        clusters += nrep;
        k = -1;
        for (int64_t i = g0; i < g1; i++) {
            if (i == g0 || r[i].umi != r[i - 1].umi) k++;
            cluster_umi[r[i].src] = rep[k];
        }
        free(du); free(dc); free(ord); free(rep); free(repi);
        g0 = g1;
    }
    free(r);
    return clusters;
}

/* ---- adapter-motif search (candidate extraction) ---------------------
 * Restates what the reference asks of edlib (third-party, pinned edlib==1.3.9 in
 * /root/reference/requirements.txt, not vendored; PARITY UNPINNED -- edlib is not installed here):
 *   edlib.align(query, target, "HW", "locations", k[, additionalEqualities])
 * at /root/reference/utils.py:134, 271, 345, 437, 1051, 1367.  Published definition (Sosic &
 * Sikic, Bioinformatics 2017; edlib.h): HW = infix mode, unit-cost edit distance of the whole
 * query against the best substring of the target; editDistance = -1 when it exceeds k; end
 * locations = all target positions (ascending) at which an optimal alignment ends; the start
 * location given for an end location is the smallest start of an optimal alignment ending there
 * (edlib.cpp: reversed SHW alignment, "taking last location as start").  With the reference's
 * ad_seq (utils.py:15) N equals A, C, G, T on either side; otherwise bytes compare exactly.
 * Plain O(m*n) DP, no bit tricks.  out: ed, nloc, first (start,end), last (start,end). */
static int hw_equal(uint8_t p, uint8_t c, int wild)
{
    if (p == c) return 1;
    if (!wild) return 0;
    int pb = p == 'A' || p == 'C' || p == 'G' || p == 'T';
    int cb = c == 'A' || c == 'C' || c == 'G' || c == 'T';
    return (p == 'N' && cb) || (c == 'N' && pb);
}

static int hw_smallest_start(const uint8_t *q, int m, const uint8_t *t, int e, int best, int wild)
{
    /* D[i][r]: first i chars of reversed q vs first r chars of t[e], t[e-1], ...; whole query,
     * prefix of the reversed target (SHW): D[0][r] = r */
    int lim = e + 1 < m + best ? e + 1 : m + best;
    int *prev = (int *)malloc(sizeof(int) * (size_t)(m + 1));
    int *cur = (int *)malloc(sizeof(int) * (size_t)(m + 1));
    for (int i = 0; i <= m; i++) prev[i] = i;
    int far = 0;
    for (int r = 1; r <= lim; r++) {
        cur[0] = r;
        uint8_t c = t[e - (r - 1)];
        for (int i = 1; i <= m; i++) {
            int s = prev[i - 1] + (hw_equal(q[m - i], c, wild) ? 0 : 1);
            int a = prev[i] + 1, b = cur[i - 1] + 1;
            if (a < s) s = a;
            if (b < s) s = b;
            cur[i] = s;
        }
        if (cur[m] == best) far = r - 1;
        int *tmp = prev; prev = cur; cur = tmp;
    }
    free(prev); free(cur);
    return e - far;
}

int nr_oracle_hw_search(const uint8_t *q, int m, const uint8_t *t, int n, int k, int wild,
                        int *out6)
{
    int *col = (int *)malloc(sizeof(int) * (size_t)(m + 1));
    for (int i = 0; i <= m; i++) col[i] = i;
    int best = m + 1, first = -1, last = -1, nloc = 0;
    for (int j = 0; j < n; j++) {
        int diag = col[0];              /* D[0][j] = 0: free start in the target */
        col[0] = 0;
        for (int i = 1; i <= m; i++) {
            int s = diag + (hw_equal(q[i - 1], t[j], wild) ? 0 : 1);
            int a = col[i] + 1, b = col[i - 1] + 1;
            diag = col[i];
            if (a < s) s = a;
            if (b < s) s = b;
            col[i] = s;
        }
        if (col[m] < best) { best = col[m]; first = last = j; nloc = 1; }
        else if (col[m] == best) { last = j; nloc++; }
    }
    free(col);
    if (n == 0 || best > k) {
        out6[0] = -1; out6[1] = 0; out6[2] = out6[3] = out6[4] = out6[5] = -1;
        return 0;
    }
    out6[0] = best; out6[1] = nloc;
    out6[2] = hw_smallest_start(q, m, t, first, best, wild); out6[3] = first;
    out6[4] = hw_smallest_start(q, m, t, last, best, wild); out6[5] = last;
    return 0;
}
