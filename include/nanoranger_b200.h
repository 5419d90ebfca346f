/*
 * nanoranger_b200.h -- C ABI of the B200 barcode matcher / UMI collapser.
 *
 * The reference (mehdiborji/nanoranger) has no FFI: the hot path is two bash scripts that
 * exec STAR between two Python functions.  Each entry point below names the reference
 * interface it replaces.  INTEGRATION.md shows the ctypes binding a maintainer would add to
 * utils.py / pipeline.py.
 *
 * Conventions
 *   - plain pointers and sizes; no torch / C++ types cross this boundary
 *   - "_device" entry points take device pointers, enqueue on `stream` (a cudaStream_t passed
 *     as void*) and never synchronise; "_host" entry points take host pointers, do their own
 *     H2D/D2H copies and return after the results are in the host buffers
 *   - every function returns 0 on success or a negative NR_E* code; nr_last_error() gives the
 *     message of the calling thread's last failure
 *   - a handle is immutable after creation and may be used from several streams
 */
#ifndef NANORANGER_B200_H
#define NANORANGER_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define NR_MAX_QUERY 64       /* longest candidate (nt) a packed record holds               */
#define NR_MAX_CORE 32        /* longest scored core: 16 (10x) or 8+18+6 = 32 (slide-seq)   */

/* error codes */
#define NR_OK 0
#define NR_EINVAL (-1)
#define NR_ECUDA (-2)
#define NR_ENOMEM (-3)
#define NR_EUNSUPPORTED (-4)

/* per-candidate flag bits (uint8) */
#define NR_FLAG_TIE 0x01        /* more than one (entry, strand) pair attains the best score  */
#define NR_FLAG_RC 0x02         /* the reported pair is on the reverse-complement strand      */
#define NR_FLAG_BELOW 0x04      /* best score < min_score (filtered mode: score not resolved)  */
#define NR_FLAG_NO_UMI 0x08     /* no optimal alignment reaches reference column padL+L        */
#define NR_FLAG_TOO_LONG 0x10   /* candidate longer than NR_MAX_QUERY: not scored               */
#define NR_FLAG_EXHAUSTIVE 0x20 /* resolved by the exhaustive kernel (informational)            */

#define NR_SCORE_BELOW (-128)   /* score value reported together with NR_FLAG_BELOW            */
#define NR_UMI_NONE 255

/* matcher modes */
#define NR_MODE_AUTO 0        /* exact at every score: seed filter where the whitelist geometry
                                 allows it, deep tier (meet in the middle) for what it leaves
                                 where the whitelist shares enough columns for it to pay,
                                 brute-force DP for the rest                                     */
#define NR_MODE_EXHAUSTIVE 1  /* brute-force DP over every (entry, strand) pair (bit-parallel,
                                 32 entries per thread, for 16- and 32-column cores)             */
#define NR_MODE_FILTERED 2    /* lossless seed filter + exact verification; exact for
                                 score >= min_score; requires a seed index
                                 (nr_whitelist_has_index) and min_score >= L - 2                 */

typedef struct nr_whitelist nr_whitelist_t;

const char *nr_last_error(void);
const char *nr_version(void);

/* ---- whitelist ("genome") -----------------------------------------------------------------
 * Replaces scripts/barcode_ref.sh:11-18 (STAR --runMode genomeGenerate over the padded FASTA
 * written by utils.write_bc_*, utils.py:584-622, 1116-1132, 1412-1458).
 * cores: n * core_len ASCII bytes (ACGTN, row-major, no separators) -- the FASTA sequences
 * without their N pads; pad_l / pad_r are the pad lengths.  Builds the packed whitelist, the
 * grouping the deep tier works on, and a seed index on `device`: the quarter-key index when
 * core_len == 16 and no entry contains N (10x lists), the anchored index when the cores are
 * 8 columns + a constant run of 12..28 columns + a tail (slide-seq: utils.py:584-601). */
int nr_whitelist_create(const char *cores, uint64_t n, uint32_t core_len, uint32_t pad_l,
                        uint32_t pad_r, int device, nr_whitelist_t **out);
void nr_whitelist_destroy(nr_whitelist_t *wl);
uint64_t nr_whitelist_size(const nr_whitelist_t *wl);
int nr_whitelist_has_index(const nr_whitelist_t *wl);
uint64_t nr_whitelist_device_bytes(const nr_whitelist_t *wl);

/* ---- candidate packing ---------------------------------------------------------------------
 * Replaces STAR's read loading (--readFilesIn/--readFilesCommand zcat,
 * scripts/barcode_align.sh:16,35).  seqs: concatenated ASCII candidate sequences;
 * offsets[i]..offsets[i+1] delimit candidate i (n + 1 offsets).  Outputs, all device:
 *   bases  n x 16 bytes   2 bit/base, base k at bit 2k of the little-endian 128-bit record
 *   meta   n bytes        length (0..64) | 0x80 if the candidate contains a non-ACGT byte;
 *                         0xFF = longer than NR_MAX_QUERY
 *   nmask  n x 8 bytes    bit k set where base k is not ACGT */
int nr_pack_device(const uint8_t *d_seqs, const uint64_t *d_offsets, uint64_t n,
                   void *d_bases, uint8_t *d_meta, uint64_t *d_nmask, void *stream);

/* ---- matcher --------------------------------------------------------------------------------
 * Replaces scripts/barcode_align.sh:14-41 (STAR EndToEnd alignment of every candidate against
 * the N-padded whitelist, unique mappers only) plus the per-record geometry that
 * utils.process_matching_* derives from the SAM (AS tag, flag, RNAME, query index aligned to
 * reference column padL+L: utils.py:697-708, 843-856, 1148-1159, 1477-1497, 636-649).
 * Per candidate i:
 *   idx[i]    entry index of the best pair (smallest index among ties), -1 if none
 *   score[i]  best AS over all entries and both strands (NR_SCORE_BELOW with NR_FLAG_BELOW)
 *   nbest[i]  number of (entry, strand) pairs attaining it, saturated at 255
 *   flags[i]  NR_FLAG_* bits
 *   umi_q[i]  query index aligned to reference column padL+L, NR_UMI_NONE if undefined
 * A candidate is "assigned" (would appear in STAR's SAM with flag 0 and pass the reference's
 * threshold) iff nbest == 1 && !(flags & (NR_FLAG_RC|NR_FLAG_BELOW|NR_FLAG_TOO_LONG)) &&
 * score >= min_score. */
int nr_match_device(const nr_whitelist_t *wl, const void *d_bases, const uint8_t *d_meta,
                    const uint64_t *d_nmask, uint64_t n, int min_score, int mode,
                    int32_t *d_idx, int8_t *d_score, uint8_t *d_nbest, uint8_t *d_flags,
                    uint8_t *d_umi_q, void *d_workspace, size_t workspace_bytes, void *stream);
size_t nr_match_workspace_bytes(const nr_whitelist_t *wl, uint64_t n, int mode);

/* Host-buffer form of the same call: H2D of the ASCII batch, pack, match, D2H of the five
 * result arrays, all on an internal stream with pinned staging; returns when the host
 * arrays are filled.  This is the call the reference-side stub binds. */
int nr_match_host(const nr_whitelist_t *wl, const char *seqs, const uint64_t *offsets,
                  uint64_t n, int min_score, int mode, int32_t *idx, int8_t *score,
                  uint8_t *nbest, uint8_t *flags, uint8_t *umi_q);

/* Pinned host memory for callers that want nr_match_host's staging copies to be the only
 * copies (optional; any host pointer works). */
void *nr_host_alloc(size_t bytes);
void nr_host_free(void *p);

/* ---- UMI collapse ---------------------------------------------------------------------------
 * Replaces the per-barcode exact UMI dedup of utils.py:759-777 (= 910-928, 1212-1230) and
 * finishes what utils.make_count_mtx_3p10XGEX (utils.py:1523-1548) starts: records
 * (barcode idx, gene id, 2-bit packed UMI) are sorted by (barcode, gene, umi); identical
 * records collapse (max_dist = 0: the reference's np.unique), and with max_dist = 1 distinct
 * UMIs of a (barcode, gene) group are merged by the directional one-hop rule.
 * UMI words are compared on all 32 bits (a caller may use codes above bit 2 * umi_len as escape
 * values for UMIs it cannot pack; with max_dist = 1 only bits below 2 * umi_len take part in the
 * neighbour search).
 * Outputs (device, caller-allocated, n entries each unless noted):
 *   rep_umi[i]     representative UMI of the cluster record i belongs to
 *   n_groups[0]    number of (barcode, gene, cluster) groups
 *   g_bc/g_gene/g_umi/g_reads  the groups in sorted order with their read counts (>= n slots) */
int nr_umi_collapse_device(const uint32_t *d_bc, const uint32_t *d_gene, const uint32_t *d_umi,
                           uint64_t n, int umi_len, int max_dist, uint32_t *d_rep_umi,
                           uint64_t *d_n_groups, uint32_t *d_g_bc, uint32_t *d_g_gene,
                           uint32_t *d_g_umi, uint32_t *d_g_reads, void *d_workspace,
                           size_t workspace_bytes, void *stream);
/* The same with declared key widths: every barcode idx < 2^bc_bits, gene id < 2^gene_bits, UMI
 * word < 2^umi_bits (umi_bits >= 2 * umi_len; 32 keeps the escape codes).  The sorts then run over
 * those bits only -- one radix sort of a 64-bit (barcode, gene, umi) key when bc_bits + gene_bits +
 * umi_bits <= 64 (a 3M-barcode whitelist, 65 536 genes, 12-nt UMIs: 62 bits), else two.  A record
 * that does not fit its declared width is not sorted wrongly in silence: n_groups[0] = ~0 and
 * the other outputs are undefined.  nr_umi_collapse_device = widths 32, 32, 32. */
int nr_umi_collapse_device_keyed(const uint32_t *d_bc, const uint32_t *d_gene, const uint32_t *d_umi,
                                 uint64_t n, int umi_len, int max_dist, int bc_bits, int gene_bits,
                                 int umi_bits, uint32_t *d_rep_umi, uint64_t *d_n_groups,
                                 uint32_t *d_g_bc, uint32_t *d_g_gene, uint32_t *d_g_umi,
                                 uint32_t *d_g_reads, void *d_workspace, size_t workspace_bytes,
                                 void *stream);
size_t nr_umi_workspace_bytes(uint64_t n);

/* ---- matcher results -> UMI records (device-resident pipelines) ------------------------------
 * Replaces the per-record loop of utils.process_matching_* (utils.py:697-718, 843-868,
 * 1148-1170, 1477-1504): a candidate becomes a record iff it is assigned (see nr_match_device)
 * and SEQ[umi_q : umi_q + umi_len] exists; record = (barcode idx, gene id, 2-bit packed UMI,
 * source candidate).  d_gene (nullable) gives the gene/transcript id of every candidate.
 * Records keep candidate order.  d_stats[0] = records written, [1] = assigned candidates dropped
 * for a short/missing UMI (the reference's "short UMI reads", utils.py:715-716), [2] = dropped
 * because the UMI contains a non-ACGT base (not representable in 2 bit/base). */
int nr_umi_records_device(const void *d_bases, const uint8_t *d_meta, const uint64_t *d_nmask,
                          const int32_t *d_idx, const int8_t *d_score, const uint8_t *d_nbest,
                          const uint8_t *d_flags, const uint8_t *d_umi_q, const uint32_t *d_gene,
                          uint64_t n, int min_score, int umi_len, uint32_t *d_rec_bc,
                          uint32_t *d_rec_gene, uint32_t *d_rec_umi, uint32_t *d_rec_src,
                          uint64_t *d_stats, void *d_workspace, size_t workspace_bytes,
                          void *stream);
size_t nr_umi_records_workspace_bytes(uint64_t n);

/* Multi-GPU UMI collapse (no reference counterpart: the reference is single-process).  Orders
 * the records by owner rank = hash(barcode idx) % world into 16-byte rows (bc, gene, umi, src)
 * and fills d_counts[world], the send counts of the one variable-count all-to-all that puts
 * every record of a barcode on one GPU.  d_cursor_scratch: world x u64.  nr_umi_unzip_device
 * splits received rows back into the arrays nr_umi_collapse_device takes (d_src nullable). */
int nr_umi_partition_device(const uint32_t *d_bc, const uint32_t *d_gene, const uint32_t *d_umi,
                            const uint32_t *d_src, uint64_t n, int world, void *d_out_records,
                            uint64_t *d_counts, uint64_t *d_cursor_scratch, void *stream);
int nr_umi_unzip_device(const void *d_records, uint64_t n, uint32_t *d_bc, uint32_t *d_gene,
                        uint32_t *d_umi, uint32_t *d_src, void *stream);

/* ---- adapter-motif search (candidate extraction) ---------------------------------------------
 * Replaces edlib.align(const, window, "HW", "locations", k[, ad_seq]) in the reference's
 * extractors (utils.py:134, 271, 345, 437, 1051, 1367): infix unit-cost edit distance of
 * `pattern` (m <= 64) against every window.  text/offsets: concatenated ASCII windows,
 * n + 1 offsets.  wildcard_n != 0: N equals A/C/G/T on either side (utils.py:15).  Per window:
 *   ed[i]        edit distance, -1 if > k (requires k < m)
 *   first[2i..]  (start, end) of the first optimal end location, end inclusive as in edlib;
 *                start = smallest start of an optimal alignment ending there
 *   last[2i..]   the same for the last optimal end location (`ed["locations"][-1]`)
 *   nloc[i]      number of optimal end locations */
int nr_hw_search_device(const uint8_t *d_text, const uint64_t *d_offsets, uint64_t n,
                        const char *pattern, int m, int k, int wildcard_n, int8_t *d_ed,
                        int32_t *d_first, int32_t *d_last, int32_t *d_nloc, void *stream);
int nr_hw_search_host(const char *text, const uint64_t *offsets, uint64_t n, const char *pattern,
                      int m, int k, int wildcard_n, int8_t *ed, int32_t *first, int32_t *last,
                      int32_t *nloc, int device);

/* ---- SAM output (host only) -------------------------------------------------------------------
 * Replaces the output stage of scripts/barcode_align.sh:14-41: writes `path` with one record per
 * candidate whose best score is reached by exactly one (entry, strand) pair (STAR:
 * --outFilterMultimapNmax 1, --outSAMunmapped None, --outSAMmode NoQS), carrying the fields
 * utils.process_matching_* read: QNAME FLAG RNAME POS 255 CIGAR * 0 0 SEQ * NH:i:1 HI:i:1 AS:i.
 * This is the FAST form (opt-in): POS/CIGAR are anchored -- an ungapped M run placed so that
 * reference column pad_l + core_len pairs with read base umi_q -- which is all the reference's
 * own parser needs; nr_sam_write_aligned below writes real alignments and nM / MD.
 * names / seqs / ref_names: concatenated bytes with n + 1 (n_ref + 1) offsets.
 * header_full != 0: one @SQ per whitelist entry (STAR's layout), else only the entries used. */
int nr_sam_write(const char *path, int header_full, const char *names, const uint64_t *name_off,
                 const char *seqs, const uint64_t *seq_off, uint64_t n, const int32_t *idx,
                 const int8_t *score, const uint8_t *nbest, const uint8_t *flags,
                 const uint8_t *umi_q, const char *ref_names, const uint64_t *ref_off,
                 uint64_t n_ref, uint32_t pad_l, uint32_t core_len, uint32_t pad_r,
                 uint64_t *n_written);

/* The same file with REAL alignments: every kept record is traced back on the host (threads
 * workers, 0 = all cores) against N^padL + core + N^padR of its entry; POS / CIGAR are that
 * alignment (M, I, D; EndToEnd: no clipping), and the attributes are the ones
 * scripts/barcode_align.sh:21 asks STAR for, in its order: AS:i nM:i MD:Z (nM = mismatches where
 * both bases are ACGT; N on either side counts as a match in MD).  Among co-optimal alignments
 * the one leaving the core at the smallest read row is reported, i.e. the pysam aligned_pairs
 * lookup of utils.py:705-708 at reference column padL+L returns umi_q.  The traceback is checked
 * against the matcher's score and umi_q; a disagreement fails the call. */
int nr_sam_write_aligned(const nr_whitelist_t *wl, const char *path, int header_full,
                         const char *names, const uint64_t *name_off, const char *seqs,
                         const uint64_t *seq_off, uint64_t n, const int32_t *idx,
                         const int8_t *score, const uint8_t *nbest, const uint8_t *flags,
                         const uint8_t *umi_q, const char *ref_names, const uint64_t *ref_off,
                         uint64_t n_ref, int threads, uint64_t *n_written);

/* ---- measurement support -------------------------------------------------------------------
 * INT-pipe roofline denominator: runs a dependent LOP3/IADD3 chain on
 * every SM for `iters` iterations and returns executed integer thread-ops per second. */
int nr_int_peak(int device, int iters, double *ops_per_s, double *ms);
/* same chains with the add half placed on the FMA pipe (LOP3 + IMAD.IADD): integer issue rate
 * with both pipes busy */
int nr_int_peak_dual(int device, int iters, double *ops_per_s, double *ms);

/* NR_MODE_FILTERED with the counting build of the kernel: fills the counters below. */
int nr_match_device_counted(const nr_whitelist_t *wl, const void *d_bases, const uint8_t *d_meta,
                            const uint64_t *d_nmask, uint64_t n, int min_score, int32_t *d_idx,
                            int8_t *d_score, uint8_t *d_nbest, uint8_t *d_flags,
                            uint8_t *d_umi_q, void *d_workspace, size_t workspace_bytes,
                            void *stream);

/* Counters of the last nr_match_device_counted launch on this workspace:
 * c[0] index probes, c[1] bitmap hits, c[2] exact verifications, c[3] verifications that found
 * cost <= 2, c[4] candidates sent to the exhaustive kernel. */
int nr_match_counters(const void *d_workspace, uint64_t *c5, void *stream);

/* Where the candidates of the last nr_match_device / nr_match_device_counted call on this
 * workspace were resolved: t[0] left by the seed filter (contain N, shorter than 24 nt, more than
 * 32 co-optimal pairs, and in NR_MODE_AUTO every read below the threshold) or, on whitelists
 * without a seed index, 0; t[1] resolved by the deep tier at cost <= 3, t[2] at cost <= 5
 * (meet-in-the-middle over the whole whitelist, nr_match_deep.cu); t[3] left to the brute-force
 * DP kernel (all of t[0] on whitelists where the deep tier does not pay: nr_match_deep.cu,
 * nr_deep_usable).  These are the reads whose scores fill the low tail of `_barcode_scores.csv`
 * (utils.py:698, 728-730). */
int nr_match_tier_counts(const void *d_workspace, uint64_t *t4, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* NANORANGER_B200_H */
