// nr_whitelist.cu -- whitelist handle: 2-bit packing, N mask and the quarter-key seed index.
// Replaces scripts/barcode_ref.sh:11-18 (STAR genomeGenerate) for the matcher kernels.
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>
#include <vector>

#include <thrust/device_ptr.h>
#include <thrust/execution_policy.h>
#include <thrust/scan.h>
#include <thrust/sort.h>

#include "nr_common.cuh"
#include "nr_filter_core.h"
#include "nr_deep_index.h"
#include "nr_anchor_index.h"

static thread_local char g_err[512] = "";

void nr_set_error(const char *fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

extern "C" const char *nr_last_error(void) { return g_err; }
extern "C" const char *nr_version(void) { return "nanoranger_b200 0.1 (sm_100a)"; }

// ---------------------------------------------------------------------------------------

__global__ void nr_index_keys_kernel(const uint32_t *__restrict__ lo, uint32_t n, int j,
                                     uint32_t *__restrict__ keys, uint32_t *__restrict__ vals)
{
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) {
        keys[i] = nr_core_key(lo[i], j);
        vals[i] = i;
    }
}

// after the (key, idx) rows are sorted by key: bitmap bits, {idx, core} rows, head flags
__global__ void nr_index_fill_kernel(const uint32_t *__restrict__ lo, uint32_t n,
                                     const uint32_t *__restrict__ keys,
                                     const uint32_t *__restrict__ vals, uint32_t *__restrict__ bits,
                                     uint2 *__restrict__ ents, uint32_t *__restrict__ heads)
{
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) {
        uint32_t k = keys[i];
        atomicOr(&bits[k >> 5], 1u << (k & 31u));
        // the same key in the other two bit orders (nr_filter_core.h: bitmap layouts)
        for (int l = 1; l < NR_BM_LAYOUTS; l++) {
            uint32_t kl = nr_key_layout(k, l);
            atomicOr(&bits[(size_t)l * 4 * NR_BM_WORDS + (kl >> 5)], 1u << (kl & 31u));
        }
        uint32_t e = vals[i];
        ents[i] = make_uint2(e, lo[e]);
        heads[i] = (i == 0 || keys[i - 1] != k) ? 1u : 0u;
    }
}

// hs = inclusive scan of the head flags: kstart[rank of a key] = its first row
__global__ void nr_index_kstart_kernel(const uint32_t *__restrict__ keys,
                                       const uint32_t *__restrict__ hs, uint32_t n,
                                       uint32_t *__restrict__ kstart)
{
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) {
        if (i == 0 || keys[i - 1] != keys[i]) kstart[hs[i] - 1] = i;
        if (i == n - 1) kstart[hs[i]] = n;
    }
}

// rank[w] = number of distinct keys < 32 w  (w in 0..2^19 inclusive)
__global__ void nr_index_rank_kernel(const uint32_t *__restrict__ keys,
                                     const uint32_t *__restrict__ hs, uint32_t n,
                                     uint32_t *__restrict__ rank)
{
    uint32_t w = blockIdx.x * blockDim.x + threadIdx.x;
    if (w > NR_BM_WORDS) return;
    uint64_t target = (uint64_t)w << 5;
    uint32_t a = 0, b = n;
    while (a < b) {
        uint32_t mid = (a + b) >> 1;
        if ((uint64_t)keys[mid] < target) a = mid + 1; else b = mid;
    }
    rank[w] = (a == n) ? hs[n - 1] : hs[a] - 1;
}

static int code_of(char c)
{
    switch (c) {
    case 'A': case 'a': return 0;
    case 'C': case 'c': return 1;
    case 'G': case 'g': return 2;
    case 'T': case 't': return 3;
    default: return 4;
    }
}

extern "C" int nr_whitelist_create(const char *cores, uint64_t n, uint32_t core_len,
                                   uint32_t pad_l, uint32_t pad_r, int device,
                                   nr_whitelist_t **out)
{
    if (!cores || !out || n == 0 || n >= (1ull << 31) || core_len == 0 ||
        core_len > NR_MAX_CORE || pad_l > 127 || pad_r > 127) {
        nr_set_error("nr_whitelist_create: bad arguments (n=%llu core_len=%u pads=%u/%u)",
                     (unsigned long long)n, core_len, pad_l, pad_r);
        return NR_EINVAL;
    }
    int prev_dev = -1;
    cudaGetDevice(&prev_dev);
    NR_CHECK_CUDA(cudaSetDevice(device));
    struct Restore { int d; ~Restore() { if (d >= 0) cudaSetDevice(d); } } restore{prev_dev};
    std::vector<uint32_t> lo(n), hi, nm;
    if (core_len > 16) hi.assign(n, 0);
    bool has_n = false;
    std::vector<uint32_t> nmv(n, 0);
    for (uint64_t e = 0; e < n; e++) {
        uint32_t l = 0, h = 0, m = 0;
        for (uint32_t j = 0; j < core_len; j++) {
            int c = code_of(cores[e * core_len + j]);
            if (c > 3) { m |= 1u << j; c = 0; }
            if (j < 16) l |= (uint32_t)c << (2 * j);
            else h |= (uint32_t)c << (2 * (j - 16));
        }
        lo[e] = l;
        if (core_len > 16) hi[e] = h;
        nmv[e] = m;
        has_n |= (m != 0);
    }
    nr_whitelist *w = new (std::nothrow) nr_whitelist();
    if (!w) { nr_set_error("out of host memory"); return NR_ENOMEM; }
    memset(w, 0, sizeof(*w));
    w->device = device; w->n = n; w->L = core_len; w->pad_l = pad_l; w->pad_r = pad_r;
    w->has_n = has_n ? 1 : 0;
    size_t nb = n * sizeof(uint32_t);
    auto fail = [&](int rc) { nr_whitelist_destroy(w); return rc; };
    if (cudaMalloc(&w->d_lo, nb) != cudaSuccess) { nr_set_error("cudaMalloc whitelist"); return fail(NR_ENOMEM); }
    w->bytes += nb;
    cudaMemcpy(w->d_lo, lo.data(), nb, cudaMemcpyHostToDevice);
    if (core_len > 16) {
        if (cudaMalloc(&w->d_hi, nb) != cudaSuccess) { nr_set_error("cudaMalloc whitelist"); return fail(NR_ENOMEM); }
        w->bytes += nb;
        cudaMemcpy(w->d_hi, hi.data(), nb, cudaMemcpyHostToDevice);
    }
    if (has_n) {
        if (cudaMalloc(&w->d_nm, nb) != cudaSuccess) { nr_set_error("cudaMalloc whitelist"); return fail(NR_ENOMEM); }
        w->bytes += nb;
        cudaMemcpy(w->d_nm, nmv.data(), nb, cudaMemcpyHostToDevice);
    }
    if (core_len == 16 && !has_n) {
        uint32_t *d_keys = nullptr, *d_vals = nullptr, *d_heads = nullptr;
        if (cudaMalloc(&d_keys, nb) != cudaSuccess || cudaMalloc(&d_vals, nb) != cudaSuccess ||
            cudaMalloc(&d_heads, nb) != cudaSuccess) {
            cudaFree(d_keys); cudaFree(d_vals);
            nr_set_error("cudaMalloc index scratch");
            return fail(NR_ENOMEM);
        }
        size_t bmb = (size_t)(NR_BM_WORDS + 1) * sizeof(uint32_t);
        uint32_t nn = (uint32_t)n;
        uint32_t blocks = (nn + 255) / 256;
        // the key bitmaps (4 dropped quarters x NR_BM_LAYOUTS bit orders) are one allocation, 2^19
        // words apart: a probe addresses bit (table id << 24 | key) from a single base pointer
        uint32_t *bits_all = nullptr;
        if (cudaMalloc(&bits_all, (size_t)4 * NR_BM_LAYOUTS * NR_BM_WORDS * sizeof(uint32_t) + 256) != cudaSuccess) {
            cudaFree(d_keys); cudaFree(d_vals); cudaFree(d_heads);
            nr_set_error("cudaMalloc seed index");
            return fail(NR_ENOMEM);
        }
        cudaMemset(bits_all, 0, (size_t)4 * NR_BM_LAYOUTS * NR_BM_WORDS * sizeof(uint32_t) + 256);
        for (int j = 0; j < 4; j++) w->d_bits[j] = bits_all + (size_t)j * NR_BM_WORDS;
        for (int j = 0; j < 4; j++) {
            if (cudaMalloc(&w->d_rank[j], bmb) != cudaSuccess ||
                cudaMalloc(&w->d_ents[j], n * sizeof(uint2)) != cudaSuccess ||
                cudaMalloc(&w->d_kstart[j], (n + 1) * sizeof(uint32_t)) != cudaSuccess) {
                cudaFree(d_keys); cudaFree(d_vals); cudaFree(d_heads);
                nr_set_error("cudaMalloc seed index");
                return fail(NR_ENOMEM);
            }
            w->bytes += (size_t)(1 + NR_BM_LAYOUTS) * bmb + n * sizeof(uint2) + (n + 1) * sizeof(uint32_t);
            nr_index_keys_kernel<<<blocks, 256>>>(w->d_lo, nn, j, d_keys, d_vals);
            thrust::stable_sort_by_key(thrust::device, thrust::device_pointer_cast(d_keys),
                                       thrust::device_pointer_cast(d_keys) + n,
                                       thrust::device_pointer_cast(d_vals));
            nr_index_fill_kernel<<<blocks, 256>>>(w->d_lo, nn, d_keys, d_vals, w->d_bits[j],
                                                  w->d_ents[j], d_heads);
            thrust::inclusive_scan(thrust::device, thrust::device_pointer_cast(d_heads),
                                   thrust::device_pointer_cast(d_heads) + n,
                                   thrust::device_pointer_cast(d_heads));
            nr_index_kstart_kernel<<<blocks, 256>>>(d_keys, d_heads, nn, w->d_kstart[j]);
            nr_index_rank_kernel<<<(NR_BM_WORDS + 1 + 255) / 256, 256>>>(d_keys, d_heads, nn,
                                                                         w->d_rank[j]);
        }
        cudaError_t e = cudaDeviceSynchronize();
        cudaFree(d_keys); cudaFree(d_vals); cudaFree(d_heads);
        if (e != cudaSuccess) {
            nr_set_error("seed index build failed: %s", cudaGetErrorString(e));
            return fail(NR_ECUDA);
        }
        w->has_index = 1;
    }
    {
        // host copies for nr_sam_write_aligned (tracebacks of the kept records run on the host)
        w->h_lo = (uint32_t *)malloc(nb);
        if (core_len > 16) w->h_hi = (uint32_t *)malloc(nb);
        if (has_n) w->h_nm = (uint32_t *)malloc(nb);
        if (!w->h_lo || (core_len > 16 && !w->h_hi) || (has_n && !w->h_nm)) {
            nr_set_error("out of host memory");
            return fail(NR_ENOMEM);
        }
        memcpy(w->h_lo, lo.data(), nb);
        if (w->h_hi) memcpy(w->h_hi, hi.data(), nb);
        if (w->h_nm) memcpy(w->h_nm, nmv.data(), nb);
    }
    if (!w->has_index) {
        // anchored seed filter: cores with a constant middle (slide-seq: 8 + linker 18 + 6)
        nr_anchor_index_host ax;
        nr_anchor_index_build(lo.data(), core_len > 16 ? hi.data() : nullptr, has_n ? nmv.data() : nullptr,
                              n, (int)core_len, ax);
        if (ax.ok) {
            const size_t b_st = ax.start.size() * 4, b_rw = ax.rows.size() * 4;
            if (cudaMalloc(&w->d_anchor_start, b_st) != cudaSuccess ||
                cudaMalloc(&w->d_anchor_rows, b_rw ? b_rw : 4) != cudaSuccess) {
                nr_set_error("cudaMalloc anchored index");
                return fail(NR_ENOMEM);
            }
            w->bytes += b_st + b_rw;
            cudaMemcpy(w->d_anchor_start, ax.start.data(), b_st, cudaMemcpyHostToDevice);
            if (b_rw) cudaMemcpy(w->d_anchor_rows, ax.rows.data(), b_rw, cudaMemcpyHostToDevice);
            w->anchor_lk = ax.Lk; w->anchor_link = ax.link;
            w->has_anchor = 1;
        }
    }
    if (core_len >= 2) {
        // deep tier: prefix / suffix grouping (host: two sorts of n keys)
        nr_deep_index_host ix;
        nr_deep_index_build(lo.data(), core_len > 16 ? hi.data() : nullptr,
                            has_n ? nmv.data() : nullptr, n, (int)core_len, 0, ix);
        w->deep_s = ix.s; w->deep_s1 = ix.s1; w->deep_u1 = ix.u1;
        w->deep_gpre = ix.g_pre; w->deep_gsuf = ix.g_suf;
        w->deep_gpmid = ix.g_pmid; w->deep_gsmid = ix.g_smid;
        const size_t b_ps = (size_t)(ix.g_pre + 1) * 4, b_pr = (size_t)ix.g_pre * 16,
                     b_ss = (size_t)(ix.g_suf + 1) * 4, b_sr = (size_t)ix.g_suf * 16,
                     b_pm = (size_t)ix.g_pmid * 16 + 16, b_sm = (size_t)ix.g_smid * 16 + 16;
        if (cudaMalloc(&w->d_deep_pre_start, b_ps) != cudaSuccess ||
            cudaMalloc(&w->d_deep_pre_rep, b_pr) != cudaSuccess ||
            cudaMalloc(&w->d_deep_suf_start, b_ss) != cudaSuccess ||
            cudaMalloc(&w->d_deep_suf_rep, b_sr) != cudaSuccess ||
            cudaMalloc(&w->d_deep_pmid_rep, b_pm) != cudaSuccess ||
            cudaMalloc(&w->d_deep_smid_rep, b_sm) != cudaSuccess ||
            cudaMalloc(&w->d_deep_ent_suf, nb) != cudaSuccess ||
            cudaMalloc(&w->d_deep_ent_idx, nb) != cudaSuccess ||
            cudaMalloc(&w->d_deep_sent_pre, nb) != cudaSuccess ||
            cudaMalloc(&w->d_deep_sent_idx, nb) != cudaSuccess) {
            nr_set_error("cudaMalloc deep index");
            return fail(NR_ENOMEM);
        }
        w->bytes += b_ps + b_pr + b_ss + b_sr + b_pm + b_sm + 4 * nb;
        cudaMemcpy(w->d_deep_pre_start, ix.pre_start.data(), b_ps, cudaMemcpyHostToDevice);
        cudaMemcpy(w->d_deep_pre_rep, ix.pre_rep.data(), b_pr, cudaMemcpyHostToDevice);
        cudaMemcpy(w->d_deep_suf_start, ix.suf_start.data(), b_ss, cudaMemcpyHostToDevice);
        cudaMemcpy(w->d_deep_suf_rep, ix.suf_rep.data(), b_sr, cudaMemcpyHostToDevice);
        if (ix.g_pmid)
            cudaMemcpy(w->d_deep_pmid_rep, ix.pmid_rep.data(), (size_t)ix.g_pmid * 16, cudaMemcpyHostToDevice);
        if (ix.g_smid)
            cudaMemcpy(w->d_deep_smid_rep, ix.smid_rep.data(), (size_t)ix.g_smid * 16, cudaMemcpyHostToDevice);
        cudaMemcpy(w->d_deep_ent_suf, ix.ent_suf.data(), nb, cudaMemcpyHostToDevice);
        cudaMemcpy(w->d_deep_ent_idx, ix.ent_idx.data(), nb, cudaMemcpyHostToDevice);
        cudaMemcpy(w->d_deep_sent_pre, ix.sent_pre.data(), nb, cudaMemcpyHostToDevice);
        cudaMemcpy(w->d_deep_sent_idx, ix.sent_idx.data(), nb, cudaMemcpyHostToDevice);
        w->has_deep = 1;
    }
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) {
        nr_set_error("whitelist upload failed: %s", cudaGetErrorString(e));
        return fail(NR_ECUDA);
    }
    *out = w;
    return NR_OK;
}

extern "C" void nr_whitelist_destroy(nr_whitelist_t *w)
{
    if (!w) return;
    int prev = -1;
    cudaGetDevice(&prev);
    cudaSetDevice(w->device);
    nr_host_ctx_destroy(w->host_ctx);
    cudaFree(w->d_lo); cudaFree(w->d_hi); cudaFree(w->d_nm);
    free(w->h_lo); free(w->h_hi); free(w->h_nm);
    cudaFree(w->d_bits[0]);
    cudaFree(w->d_anchor_start); cudaFree(w->d_anchor_rows);
    cudaFree(w->d_deep_pre_start); cudaFree(w->d_deep_pre_rep); cudaFree(w->d_deep_suf_rep);
    cudaFree(w->d_deep_ent_suf); cudaFree(w->d_deep_ent_idx);
    cudaFree(w->d_deep_suf_start); cudaFree(w->d_deep_sent_pre); cudaFree(w->d_deep_sent_idx);
    cudaFree(w->d_deep_pmid_rep); cudaFree(w->d_deep_smid_rep);
    for (int j = 0; j < 4; j++) { cudaFree(w->d_rank[j]); cudaFree(w->d_ents[j]); cudaFree(w->d_kstart[j]); }
    delete w;
    if (prev >= 0) cudaSetDevice(prev);
}

extern "C" uint64_t nr_whitelist_size(const nr_whitelist_t *w) { return w ? w->n : 0; }
extern "C" int nr_whitelist_has_index(const nr_whitelist_t *w) { return w ? (w->has_index || w->has_anchor) : 0; }
extern "C" uint64_t nr_whitelist_device_bytes(const nr_whitelist_t *w) { return w ? w->bytes : 0; }
