// nr_common.cuh -- shared definitions of the B200 barcode matcher (sm_100a).
//
// Scoring (what scripts/barcode_align.sh:18-33 configures STAR with):
//   reference R = N^padL + core + N^padR, query Q consumed end to end, free start/end in R,
//   match +1, mismatch -1, gap -1 per base, N on either side scores 0.
// Packed layouts:
//   candidate  16 B, 2 bit/base (A0 C1 G2 T3), base k at bit 2k of the little-endian record;
//              meta byte = len | 0x80 (contains non-ACGT) or 0xFF (too long); nmask u64
//   whitelist  core columns 0..15 in `lo`, 16..31 in `hi` (same bit order), N columns in `nm`
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/nanoranger_b200.h"

#define NR_CHECK_CUDA(expr)                                                        \
    do {                                                                           \
        cudaError_t _e = (expr);                                                   \
        if (_e != cudaSuccess) {                                                   \
            nr_set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e),   \
                         __FILE__, __LINE__);                                      \
            return NR_ECUDA;                                                       \
        }                                                                          \
    } while (0)

void nr_set_error(const char *fmt, ...);

struct nr_whitelist {
    int device;
    uint64_t n;
    uint32_t L, pad_l, pad_r;
    int has_n;       // some entry contains N
    int has_index;   // quarter-key seed index present (L == 16 && !has_n)
    uint32_t *d_lo;  // n
    uint32_t *d_hi;  // n (L > 16) or nullptr
    uint32_t *d_nm;  // n (has_n) or nullptr
    uint32_t *h_lo, *h_hi, *h_nm;   // host copies (hi / nm nullable): the SAM writer's tracebacks
    // seed index: for j in 0..3 the 24-bit key made of the three 4-base quarters other than
    // quarter j.  bits[j][w] = bitmap word of keys 32w..32w+31, rank[j][w] = number of distinct
    // keys below 32w (2^19 + 1 words each); kstart[j][r] = first row of the r-th distinct key
    // (+ sentinel n); ents[j] = rows {entry idx, core} sorted by key_j.
    uint32_t *d_bits[4];
    uint32_t *d_rank[4];
    uint2 *d_ents[4];
    uint32_t *d_kstart[4];
    // anchored seed filter (nr_anchor_core.h): cores = 8 columns + constant linker + tail (slide-seq)
    int has_anchor;
    int anchor_lk;                // linker columns
    uint64_t anchor_link;         // the linker, column k at bits 2k
    uint32_t *d_anchor_start;     // 4^8 + 1: rows of every 8-mer in front of the linker
    uint32_t *d_anchor_rows;      // entry indices grouped by that 8-mer (N columns expanded)
    // deep tier (nr_deep_core.h, nr_deep_index.h): entries grouped by their first deep_s columns
    // (prefix groups, runs of the sorted order) and by their last L - deep_s (suffix groups)
    int has_deep;
    int deep_s, deep_s1, deep_u1;     // split column; columns shared inside each half (mid groups)
    uint32_t deep_gpre, deep_gsuf, deep_gpmid, deep_gsmid;
    uint32_t *d_deep_pre_start;   // deep_gpre + 1
    uint4 *d_deep_pre_rep;        // deep_gpre: {lo, hi, nm, 0} of the group's columns
    uint4 *d_deep_suf_rep;        // deep_gsuf
    uint32_t *d_deep_ent_suf;     // n: suffix group per sorted position
    uint32_t *d_deep_ent_idx;     // n: entry index per sorted position
    uint32_t *d_deep_suf_start;   // deep_gsuf + 1: the same entries in suffix-group order
    uint32_t *d_deep_sent_pre;    // n: prefix group per suffix-sorted position
    uint32_t *d_deep_sent_idx;    // n: entry index per suffix-sorted position
    uint4 *d_deep_pmid_rep;       // deep_gpmid: mid groups of the prefix side (pre_rep.w = parent)
    uint4 *d_deep_smid_rep;       // deep_gsmid
    size_t bytes;
    void *host_ctx;  // lazily created staging state of nr_match_host (nr_match_api.cu)
};

void nr_host_ctx_destroy(void *ctx);

#define NR_BM_WORDS (1u << 19)

// exhaustive kernel's split scratch: NR_EX_MAXGRID u32 counters, then NR_EX_MAXGRID uint4 partials
#define NR_EX_MAXGRID 2048
#define NR_EX_SCRATCH_BYTES (NR_EX_MAXGRID * 4 + NR_EX_MAXGRID * 16)

// ---------------------------------------------------------------------------------------
// device helpers

__device__ __forceinline__ uint32_t nr_lane() { return threadIdx.x & 31u; }
