// nr_records.cu -- matcher results -> UMI records, and their partition by owner rank.
//
// Device-resident form of what utils.process_matching_* does per SAM record in the reference
// (utils.py:697-718 and twins at :843-868, :1148-1170, :1477-1504): accept a candidate iff its
// best score is reached by exactly one (entry, strand) pair on the forward strand with
// AS >= thr (`AS>=14 and flag==0`), take umi = SEQ[q : q+umi_len] at the query index q aligned
// to reference column padL+L, drop it when shorter than umi_len.  UMIs containing a non-ACGT
// base cannot be 2-bit packed; they are dropped here and counted (the host layer in utils.py
// keeps them, as the reference does).
//
// nr_umi_partition_device orders the records by owner rank = hash(barcode) % world so that ONE
// variable-count all-to-all (NCCL over NVLink) puts every record of a barcode on one GPU
// (SURVEY.md section 8e); the hash is the one nanoranger_b200/umi.py:owner_rank computes.
#include <cub/cub.cuh>

#include "nr_common.cuh"

namespace {

constexpr int REC_T = 256;     // threads per block
constexpr int REC_I = 4;       // candidates per thread
constexpr int REC_B = REC_T * REC_I;

struct MatchView {
    const uint4 *bases;
    const uint8_t *meta;
    const uint64_t *nmask;
    const int32_t *idx;
    const int8_t *score;
    const uint8_t *nbest, *flags, *umi_q;
};

// 0 = not assigned, 1 = record, 2 = assigned but UMI short / missing, 3 = UMI contains N
__device__ __forceinline__ int extract(const MatchView &v, uint64_t i, int min_score, int umi_len,
                                       uint32_t *bc, uint32_t *umi)
{
    const uint32_t fl = v.flags[i];
    if (v.nbest[i] != 1 || (fl & (NR_FLAG_RC | NR_FLAG_BELOW | NR_FLAG_TOO_LONG)) ||
        (int)v.score[i] < min_score)
        return 0;
    const uint32_t q = v.umi_q[i];
    const uint32_t m = v.meta[i] & 0x7Fu;
    if (q == NR_UMI_NONE || q + (uint32_t)umi_len > m) return 2;
    const uint64_t span = (umi_len >= 32 ? ~0ull : ((1ull << umi_len) - 1ull)) << q;
    if ((v.meta[i] & 0x80u) && (v.nmask[i] & span)) return 3;
    const uint32_t *w = reinterpret_cast<const uint32_t *>(v.bases + i);
    const uint32_t k = q >> 4, s = (q & 15u) * 2u;
    const uint32_t lo = w[k], hi = k + 1 < 4 ? w[k + 1] : 0u;
    uint32_t u = __funnelshift_r(lo, hi, s);
    if (umi_len < 16) u &= (1u << (2 * umi_len)) - 1u;
    *bc = (uint32_t)v.idx[i];
    *umi = u;
    return 1;
}

__global__ void __launch_bounds__(REC_T)
k_rec_count(MatchView v, uint64_t n, int min_score, int umi_len, uint32_t *block_cnt,
            unsigned long long *stats)
{
    __shared__ uint32_t sh[4];
    if (threadIdx.x < 4) sh[threadIdx.x] = 0;
    __syncthreads();
    uint32_t c[4] = {0, 0, 0, 0};
    const uint64_t base = (uint64_t)blockIdx.x * REC_B;
#pragma unroll
    for (int k = 0; k < REC_I; k++) {
        uint64_t i = base + (uint64_t)k * REC_T + threadIdx.x;
        if (i < n) { uint32_t b, u; c[extract(v, i, min_score, umi_len, &b, &u)]++; }
    }
#pragma unroll
    for (int t = 1; t < 4; t++) {
        uint32_t s = __reduce_add_sync(0xffffffffu, c[t]);
        if ((threadIdx.x & 31) == 0 && s) atomicAdd(&sh[t], s);
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        block_cnt[blockIdx.x] = sh[1];
        if (sh[2]) atomicAdd(stats + 1, (unsigned long long)sh[2]);
        if (sh[3]) atomicAdd(stats + 2, (unsigned long long)sh[3]);
    }
}

// block_off = exclusive scan of block_cnt; records keep candidate order
__global__ void __launch_bounds__(REC_T)
k_rec_write(MatchView v, const uint32_t *gene, uint64_t n, int min_score, int umi_len,
            const uint32_t *block_off, const uint32_t *block_cnt, uint32_t n_blocks,
            uint32_t *r_bc, uint32_t *r_gene, uint32_t *r_umi, uint32_t *r_src,
            unsigned long long *stats)
{
    typedef cub::BlockScan<uint32_t, REC_T> Scan;
    __shared__ typename Scan::TempStorage tmp;
    const uint64_t base = (uint64_t)blockIdx.x * REC_B;
    // item order inside the block: thread t owns candidates base + t*REC_I .. +REC_I-1
    uint32_t bc[REC_I], um[REC_I], ok[REC_I], mine = 0;
#pragma unroll
    for (int k = 0; k < REC_I; k++) {
        uint64_t i = base + (uint64_t)threadIdx.x * REC_I + k;
        ok[k] = 0;
        if (i < n) ok[k] = extract(v, i, min_score, umi_len, &bc[k], &um[k]) == 1;
        mine += ok[k];
    }
    uint32_t excl;
    Scan(tmp).ExclusiveSum(mine, excl);
    uint32_t pos = block_off[blockIdx.x] + excl;
#pragma unroll
    for (int k = 0; k < REC_I; k++) {
        if (ok[k]) {
            uint64_t i = base + (uint64_t)threadIdx.x * REC_I + k;
            r_bc[pos] = bc[k];
            r_gene[pos] = gene ? gene[i] : 0u;
            r_umi[pos] = um[k];
            if (r_src) r_src[pos] = (uint32_t)i;
            pos++;
        }
    }
    if (blockIdx.x == n_blocks - 1 && threadIdx.x == 0)
        stats[0] = (unsigned long long)block_off[blockIdx.x] + block_cnt[blockIdx.x];
}

__device__ __forceinline__ uint32_t owner_of(uint32_t bc, uint32_t world)
{
    // nanoranger_b200/umi.py:owner_rank
    return (uint32_t)(((uint64_t)bc * 0x9E3779B97F4A7C15ull >> 40) % world);
}

constexpr int PART_MAXW = 256;

__global__ void __launch_bounds__(REC_T)
k_part_hist(const uint32_t *bc, uint64_t n, uint32_t world, unsigned long long *counts)
{
    __shared__ uint32_t h[PART_MAXW];
    for (uint32_t t = threadIdx.x; t < world; t += REC_T) h[t] = 0;
    __syncthreads();
    for (uint64_t i = (uint64_t)blockIdx.x * REC_T + threadIdx.x; i < n;
         i += (uint64_t)gridDim.x * REC_T)
        atomicAdd(&h[owner_of(bc[i], world)], 1u);
    __syncthreads();
    for (uint32_t t = threadIdx.x; t < world; t += REC_T)
        if (h[t]) atomicAdd(counts + t, (unsigned long long)h[t]);
}

__global__ void k_part_cursor(const unsigned long long *counts, uint32_t world,
                              unsigned long long *cursor)
{
    if (threadIdx.x == 0 && blockIdx.x == 0) {
        unsigned long long a = 0;
        for (uint32_t t = 0; t < world; t++) { cursor[t] = a; a += counts[t]; }
    }
}

// each block reserves one range per owner, then places its records inside the ranges
__global__ void __launch_bounds__(REC_T)
k_part_scatter(const uint32_t *bc, const uint32_t *gene, const uint32_t *umi, const uint32_t *src,
               uint64_t n, uint32_t world, unsigned long long *cursor, uint4 *out)
{
    __shared__ uint32_t h[PART_MAXW];
    __shared__ unsigned long long basep[PART_MAXW];
    const uint64_t base = (uint64_t)blockIdx.x * REC_B;
    for (uint32_t t = threadIdx.x; t < world; t += REC_T) h[t] = 0;
    __syncthreads();
    uint32_t own[REC_I];
#pragma unroll
    for (int k = 0; k < REC_I; k++) {
        uint64_t i = base + (uint64_t)k * REC_T + threadIdx.x;
        own[k] = 0xFFFFFFFFu;
        if (i < n) { own[k] = owner_of(bc[i], world); atomicAdd(&h[own[k]], 1u); }
    }
    __syncthreads();
    for (uint32_t t = threadIdx.x; t < world; t += REC_T) {
        basep[t] = h[t] ? atomicAdd(cursor + t, (unsigned long long)h[t]) : 0ull;
        h[t] = 0;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < REC_I; k++) {
        uint64_t i = base + (uint64_t)k * REC_T + threadIdx.x;
        if (own[k] != 0xFFFFFFFFu) {
            unsigned long long p = basep[own[k]] + atomicAdd(&h[own[k]], 1u);
            out[p] = make_uint4(bc[i], gene ? gene[i] : 0u, umi[i], src ? src[i] : (uint32_t)i);
        }
    }
}

__global__ void k_unzip(const uint4 *rec, uint64_t n, uint32_t *bc, uint32_t *gene, uint32_t *umi,
                        uint32_t *src)
{
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) {
        uint4 r = rec[i];
        bc[i] = r.x; gene[i] = r.y; umi[i] = r.z;
        if (src) src[i] = r.w;
    }
}

size_t rec_cub_bytes(uint64_t nblocks) { return ((nblocks / 64 + 1) * 64 + (1u << 20) + 255) & ~(size_t)255; }

}  // namespace

extern "C" size_t nr_umi_records_workspace_bytes(uint64_t n)
{
    uint64_t nb = (n + REC_B - 1) / REC_B + 1;
    return 256 + 2 * ((nb * 4 + 255) & ~(size_t)255) + rec_cub_bytes(nb);
}

extern "C" int nr_umi_records_device(const void *d_bases, const uint8_t *d_meta,
                                     const uint64_t *d_nmask, const int32_t *d_idx,
                                     const int8_t *d_score, const uint8_t *d_nbest,
                                     const uint8_t *d_flags, const uint8_t *d_umi_q,
                                     const uint32_t *d_gene, uint64_t n, int min_score, int umi_len,
                                     uint32_t *d_rec_bc, uint32_t *d_rec_gene, uint32_t *d_rec_umi,
                                     uint32_t *d_rec_src, uint64_t *d_stats, void *d_workspace,
                                     size_t workspace_bytes, void *stream)
{
    cudaStream_t st = (cudaStream_t)stream;
    if (!d_stats) { nr_set_error("nr_umi_records_device: null pointer"); return NR_EINVAL; }
    NR_CHECK_CUDA(cudaMemsetAsync(d_stats, 0, 3 * sizeof(uint64_t), st));
    if (n == 0) return NR_OK;
    if (!d_bases || !d_meta || !d_nmask || !d_idx || !d_score || !d_nbest || !d_flags ||
        !d_umi_q || !d_rec_bc || !d_rec_gene || !d_rec_umi || !d_workspace) {
        nr_set_error("nr_umi_records_device: null pointer");
        return NR_EINVAL;
    }
    if (umi_len < 1 || umi_len > 16 || n >= (1ull << 32)) {
        nr_set_error("nr_umi_records_device: umi_len 1..16, n < 2^32");
        return NR_EINVAL;
    }
    if (workspace_bytes < nr_umi_records_workspace_bytes(n)) {
        nr_set_error("nr_umi_records_device: workspace too small");
        return NR_EINVAL;
    }
    const uint32_t nb = (uint32_t)((n + REC_B - 1) / REC_B);
    const size_t arr = ((size_t)(nb + 1) * 4 + 255) & ~(size_t)255;
    uint8_t *ws = (uint8_t *)d_workspace;
    uint32_t *cnt = (uint32_t *)(ws + 256);
    uint32_t *off = (uint32_t *)(ws + 256 + arr);
    void *tmp = ws + 256 + 2 * arr;
    size_t tmp_bytes = rec_cub_bytes(nb + 1);
    MatchView v{(const uint4 *)d_bases, d_meta, d_nmask, d_idx, d_score, d_nbest, d_flags, d_umi_q};
    unsigned long long *stats = (unsigned long long *)d_stats;
    k_rec_count<<<nb, REC_T, 0, st>>>(v, n, min_score, umi_len, cnt, stats);
    size_t need = 0;
    cub::DeviceScan::ExclusiveSum(nullptr, need, cnt, off, (int)nb, st);
    if (need > tmp_bytes) { nr_set_error("nr_umi_records_device: scan storage"); return NR_ENOMEM; }
    NR_CHECK_CUDA(cub::DeviceScan::ExclusiveSum(tmp, tmp_bytes, cnt, off, (int)nb, st));
    k_rec_write<<<nb, REC_T, 0, st>>>(v, d_gene, n, min_score, umi_len, off, cnt, nb, d_rec_bc,
                                      d_rec_gene, d_rec_umi, d_rec_src, stats);
    NR_CHECK_CUDA(cudaGetLastError());
    return NR_OK;
}

extern "C" int nr_umi_partition_device(const uint32_t *d_bc, const uint32_t *d_gene,
                                       const uint32_t *d_umi, const uint32_t *d_src, uint64_t n,
                                       int world, void *d_out_records, uint64_t *d_counts,
                                       uint64_t *d_cursor_scratch, void *stream)
{
    cudaStream_t st = (cudaStream_t)stream;
    if (world < 1 || world > PART_MAXW) {
        nr_set_error("nr_umi_partition_device: world must be 1..%d", PART_MAXW);
        return NR_EINVAL;
    }
    if (!d_counts || !d_cursor_scratch) { nr_set_error("nr_umi_partition_device: null pointer"); return NR_EINVAL; }
    NR_CHECK_CUDA(cudaMemsetAsync(d_counts, 0, (size_t)world * sizeof(uint64_t), st));
    if (n == 0) return NR_OK;
    if (!d_bc || !d_umi || !d_out_records) { nr_set_error("nr_umi_partition_device: null pointer"); return NR_EINVAL; }
    int sms = 148, dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    uint64_t want = (n + REC_T - 1) / REC_T;
    unsigned hb = (unsigned)(want < (uint64_t)sms * 8 ? want : (uint64_t)sms * 8);
    k_part_hist<<<hb, REC_T, 0, st>>>(d_bc, n, (uint32_t)world, (unsigned long long *)d_counts);
    k_part_cursor<<<1, 32, 0, st>>>((const unsigned long long *)d_counts, (uint32_t)world,
                                    (unsigned long long *)d_cursor_scratch);
    unsigned nb = (unsigned)((n + REC_B - 1) / REC_B);
    k_part_scatter<<<nb, REC_T, 0, st>>>(d_bc, d_gene, d_umi, d_src, n, (uint32_t)world,
                                         (unsigned long long *)d_cursor_scratch,
                                         (uint4 *)d_out_records);
    NR_CHECK_CUDA(cudaGetLastError());
    return NR_OK;
}

extern "C" int nr_umi_unzip_device(const void *d_records, uint64_t n, uint32_t *d_bc,
                                   uint32_t *d_gene, uint32_t *d_umi, uint32_t *d_src, void *stream)
{
    if (n == 0) return NR_OK;
    if (!d_records || !d_bc || !d_gene || !d_umi) { nr_set_error("nr_umi_unzip_device: null pointer"); return NR_EINVAL; }
    k_unzip<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
        (const uint4 *)d_records, n, d_bc, d_gene, d_umi, d_src);
    NR_CHECK_CUDA(cudaGetLastError());
    return NR_OK;
}
