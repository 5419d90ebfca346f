// nr_match_api.cu -- C ABI of the matcher: device-pointer entry point, host-buffer entry point
// (chunked, double-buffered H2D -> pack -> match -> D2H) and the workspace/counter helpers.
// Replaces scripts/barcode_align.sh:14-41 of the reference (see include/nanoranger_b200.h).
#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <new>
#include <thread>
#include <vector>

#include "nr_common.cuh"

int nr_launch_exhaustive(const nr_whitelist *wl, const void *d_bases, const uint8_t *d_meta,
                         const uint64_t *d_nmask, const uint32_t *d_list,
                         const uint32_t *d_list_count, uint64_t n_cand, int min_score,
                         int32_t *d_idx, int8_t *d_score, uint8_t *d_nbest, uint8_t *d_flags,
                         uint8_t *d_umi, int grid_cap, void *d_scratch, cudaStream_t stream);
int nr_launch_filtered(const nr_whitelist *wl, const void *d_bases, const uint8_t *d_meta,
                       const uint64_t *d_nmask, uint64_t n_cand, int min_score, int resolve_below,
                       int32_t *d_idx, int8_t *d_score, uint8_t *d_nbest, uint8_t *d_flags,
                       uint8_t *d_umi, uint32_t *d_list, uint32_t *d_list_count,
                       uint32_t *d_list_n, uint32_t *d_list_n_count,
                       unsigned long long *d_tile_next, unsigned long long *d_tile_next_n,
                       unsigned long long *d_counters, int *grid_out, cudaStream_t stream);

int nr_launch_anchored(const nr_whitelist *wl, const void *d_bases, const uint8_t *d_meta,
                       const uint64_t *d_nmask, uint64_t n_cand, int min_score, int resolve_below, int32_t *d_idx,
                       int8_t *d_score, uint8_t *d_nbest, uint8_t *d_flags, uint8_t *d_umi,
                       uint32_t *d_list, uint32_t *d_list_count, unsigned long long *d_tile_next,
                       unsigned long long *d_counters, cudaStream_t stream);
int nr_launch_deep(const nr_whitelist *wl, int K, const void *d_bases, const uint8_t *d_meta,
                   const uint64_t *d_nmask, const uint32_t *d_list, const uint32_t *d_list_count,
                   uint64_t n_cand, int min_score, int32_t *d_idx, int8_t *d_score,
                   uint8_t *d_nbest, uint8_t *d_flags, uint8_t *d_umi, uint32_t *d_next_list,
                   uint32_t *d_next_count, unsigned long long *d_work_next,
                   unsigned long long *d_resolved, uint8_t *d_scratch, size_t scratch_bytes,
                   int small, cudaStream_t stream);
int nr_launch_deep_finalize(const nr_whitelist *wl, const void *d_bases, const uint8_t *d_meta,
                            const uint64_t *d_nmask, const uint32_t *d_list,
                            const uint32_t *d_list_count, uint64_t n_cand, const int32_t *d_idx,
                            const int8_t *d_score, uint8_t *d_flags, uint8_t *d_umi,
                            cudaStream_t stream);
size_t nr_deep_scratch_bytes(const nr_whitelist *wl, int K);
int nr_deep_usable(const nr_whitelist *wl);

// workspace layout (header zeroed before every call):
//   [0,64)    eight u64 counters: probes, hits, verifications, passes, listed (counting build of
//             the filtered kernel), candidates resolved by the deep tier at K = 3 and at K = 5
//   [64,68)   count of list A: candidates the filtered kernel leaves (-> deep tier, K = 3)
//   [68,72)   count of list B: left by the deep tier at K = 3 (-> K = 5)
//   [72,80)   next tile of the filtered kernel
//   [80,84)   count of list C: left by the deep tier at K = 5 (-> exhaustive DP kernel)
//   [84,88)   count of list N: reads with N the filter's main pass hands to its N pass
//   [88,96)   next work item of the deep tier at K = 3, [96,104) at K = 5
//   [104,112) next tile of the filtered kernel's N pass
//   [128, 128 + 8 KB)  the exhaustive kernel's arrival counters, then its partials
//   [NR_WS_HEADER, +4n) list A, then list B, list C, list N (4n each), then the deep tier's scratch
#define NR_WS_ZERO (128 + NR_EX_MAXGRID * 4)
#define NR_WS_HEADER 65536
static_assert(128 + NR_EX_SCRATCH_BYTES <= NR_WS_HEADER, "workspace header too small");

namespace {
struct DeviceGuard {
    int prev = -1;
    bool switched = false;
    explicit DeviceGuard(int dev)
    {
        if (cudaGetDevice(&prev) == cudaSuccess && prev != dev) {
            cudaSetDevice(dev);
            switched = true;
        }
    }
    ~DeviceGuard() { if (switched) cudaSetDevice(prev); }
};

size_t lists_bytes(uint64_t n) { return (((size_t)n * sizeof(uint32_t) * 4) + 255) & ~(size_t)255; }

size_t deep_scratch(const nr_whitelist *wl)
{
    const size_t a = nr_deep_scratch_bytes(wl, 3), b = nr_deep_scratch_bytes(wl, 5);
    return a > b ? a : b;
}

// list A (or every candidate when from_all) -> deep tier K = 3 -> K = 5 -> exhaustive DP kernel,
// all on the stream, counts on the device
int resolve_chain(const nr_whitelist *wl, const void *d_bases, const uint8_t *d_meta,
                  const uint64_t *d_nmask, uint64_t n, int min_score, bool from_all, bool few,
                  int32_t *d_idx, int8_t *d_score, uint8_t *d_nbest, uint8_t *d_flags,
                  uint8_t *d_umi_q, uint8_t *ws, size_t ws_bytes, int sms, cudaStream_t st)
{
    uint32_t *listA = (uint32_t *)(ws + NR_WS_HEADER), *listB = listA + n, *listC = listB + n;
    uint32_t *cntA = (uint32_t *)(ws + 64), *cntB = (uint32_t *)(ws + 68), *cntC = (uint32_t *)(ws + 80);
    unsigned long long *ctr = (unsigned long long *)ws;
    const uint32_t *ex_list = listA, *ex_cnt = cntA;
    if (nr_deep_usable(wl)) {
        uint8_t *scratch = ws + NR_WS_HEADER + lists_bytes(n);
        const size_t sb = ws_bytes - NR_WS_HEADER - lists_bytes(n);
        int rc = nr_launch_deep(wl, 3, d_bases, d_meta, d_nmask, from_all ? nullptr : listA, cntA, n,
                                min_score, d_idx, d_score, d_nbest, d_flags, d_umi_q, listB, cntB,
                                (unsigned long long *)(ws + 88), ctr + 5, scratch, sb, few, st);
        if (rc != NR_OK) return rc;
        rc = nr_launch_deep(wl, 5, d_bases, d_meta, d_nmask, listB, cntB, n, min_score, d_idx,
                            d_score, d_nbest, d_flags, d_umi_q, listC, cntC,
                            (unsigned long long *)(ws + 96), ctr + 6, scratch, sb, few, st);
        if (rc != NR_OK) return rc;
        ex_list = listC; ex_cnt = cntC;
    } else if (from_all) {
        ex_list = nullptr; ex_cnt = nullptr;
    } else {
        // no deep tier on this whitelist: everything the filter left is the brute-force kernel's
        // (list C's count is what nr_match_tier_counts reports for it)
        NR_CHECK_CUDA(cudaMemcpyAsync(cntC, cntA, sizeof(uint32_t), cudaMemcpyDeviceToDevice, st));
    }
    int rc = nr_launch_exhaustive(wl, d_bases, d_meta, d_nmask, ex_list, ex_cnt, n, min_score, d_idx,
                                  d_score, d_nbest, d_flags, d_umi_q, sms * 2, ws + 128, st);
    if (rc != NR_OK || !nr_deep_usable(wl)) return rc;
    // UMI columns of what the deep tier resolved.  LAST: the finaliser walks the tier's input list
    // and picks the candidates whose umi_q is the pending mark, so every candidate of that list
    // must have been written by now (the brute-force kernel never writes the mark; before it ran,
    // its candidates would still hold whatever the caller's buffer contained)
    return nr_launch_deep_finalize(wl, d_bases, d_meta, d_nmask, from_all ? nullptr : listA, cntA, n,
                                   d_idx, d_score, d_flags, d_umi_q, st);
}
}  // namespace

extern "C" size_t nr_match_workspace_bytes(const nr_whitelist_t *wl, uint64_t n, int mode)
{
    (void)mode;
    return NR_WS_HEADER + lists_bytes(n) + (wl ? deep_scratch(wl) : 0);
}

static int resolve_mode(const nr_whitelist *wl, int min_score, int mode)
{
    bool can_filter = (wl->has_index || wl->has_anchor) && min_score >= (int)wl->L - 2;
    if (mode == NR_MODE_AUTO) return can_filter ? NR_MODE_AUTO : NR_MODE_EXHAUSTIVE;
    if (mode == NR_MODE_FILTERED && !can_filter) return -1;
    return mode;
}

extern "C" int nr_match_device(const nr_whitelist_t *wl, const void *d_bases,
                               const uint8_t *d_meta, const uint64_t *d_nmask, uint64_t n,
                               int min_score, int mode, int32_t *d_idx, int8_t *d_score,
                               uint8_t *d_nbest, uint8_t *d_flags, uint8_t *d_umi_q,
                               void *d_workspace, size_t workspace_bytes, void *stream)
{
    if (!wl) { nr_set_error("nr_match_device: null whitelist"); return NR_EINVAL; }
    if (n == 0) return NR_OK;
    if (!d_bases || !d_meta || !d_nmask || !d_idx || !d_score || !d_nbest || !d_flags || !d_umi_q) {
        nr_set_error("nr_match_device: null pointer");
        return NR_EINVAL;
    }
    if (n >= (1ull << 32)) { nr_set_error("nr_match_device: n must be < 2^32 per call"); return NR_EINVAL; }
    if (mode != NR_MODE_AUTO && mode != NR_MODE_EXHAUSTIVE && mode != NR_MODE_FILTERED) {
        nr_set_error("nr_match_device: unknown mode %d", mode);
        return NR_EINVAL;
    }
    int eff = resolve_mode(wl, min_score, mode);
    if (eff < 0) {
        nr_set_error("NR_MODE_FILTERED needs a whitelist with a seed index (16-column N-free cores, or "
                     "8 + linker + tail cores) and min_score >= %d",
                     (int)wl->L - 2);
        return NR_EUNSUPPORTED;
    }
    DeviceGuard guard(wl->device);
    cudaStream_t st = (cudaStream_t)stream;
    int sms = 148;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, wl->device);
    const bool ws_ok = d_workspace && workspace_bytes >= nr_match_workspace_bytes(wl, n, mode);
    if (eff == NR_MODE_EXHAUSTIVE && (mode == NR_MODE_EXHAUSTIVE || !nr_deep_usable(wl) || !ws_ok)) {
        // the brute-force DP over every (entry, strand) pair.  The workspace is optional here:
        // with one, batches smaller than the grid are split over the whitelist
        void *scratch = nullptr;
        if (d_workspace && workspace_bytes >= NR_WS_HEADER) {
            NR_CHECK_CUDA(cudaMemsetAsync(d_workspace, 0, NR_WS_ZERO, st));
            scratch = (uint8_t *)d_workspace + 128;
        }
        return nr_launch_exhaustive(wl, d_bases, d_meta, d_nmask, nullptr, nullptr, n, min_score,
                                    d_idx, d_score, d_nbest, d_flags, d_umi_q, sms * 2, scratch, st);
    }
    if (!ws_ok) {
        nr_set_error("nr_match_device: workspace too small (%zu < %zu)", workspace_bytes,
                     nr_match_workspace_bytes(wl, n, mode));
        return NR_EINVAL;
    }
    uint8_t *ws = (uint8_t *)d_workspace;
    NR_CHECK_CUDA(cudaMemsetAsync(ws, 0, NR_WS_ZERO, st));
    if (eff == NR_MODE_EXHAUSTIVE)
        // AUTO on a whitelist without a seed index (slide-seq cores, N columns): every candidate
        // through the deep tier, the brute-force kernel only for what it leaves
        return resolve_chain(wl, d_bases, d_meta, d_nmask, n, min_score, true, false, d_idx, d_score,
                             d_nbest, d_flags, d_umi_q, ws, workspace_bytes, sms, st);
    uint32_t *d_count = (uint32_t *)(ws + 64);
    uint32_t *d_list = (uint32_t *)(ws + NR_WS_HEADER);
    int grid = 0;
    if (wl->has_anchor) {
        int rc = nr_launch_anchored(wl, d_bases, d_meta, d_nmask, n, min_score, eff == NR_MODE_AUTO ? 1 : 0, d_idx,
                                    d_score, d_nbest, d_flags, d_umi_q, d_list, d_count,
                                    (unsigned long long *)(ws + 72), nullptr, st);
        if (rc != NR_OK) return rc;
        return resolve_chain(wl, d_bases, d_meta, d_nmask, n, min_score, false, eff != NR_MODE_AUTO, d_idx,
                             d_score, d_nbest, d_flags, d_umi_q, ws, workspace_bytes, sms, st);
    }
    int rc = nr_launch_filtered(wl, d_bases, d_meta, d_nmask, n, min_score,
                                eff == NR_MODE_AUTO ? 1 : 0, d_idx, d_score, d_nbest, d_flags,
                                d_umi_q, d_list, d_count, d_list + 3 * n, (uint32_t *)(ws + 84),
                                (unsigned long long *)(ws + 72), (unsigned long long *)(ws + 104),
                                nullptr, &grid, st);
    if (rc != NR_OK) return rc;
    // candidates the filter left (more than two N, short, > 32 co-optimal pairs; in AUTO also
    // everything below cost 2) are resolved exactly, counts read on the device
    return resolve_chain(wl, d_bases, d_meta, d_nmask, n, min_score, false, eff != NR_MODE_AUTO, d_idx, d_score,
                         d_nbest, d_flags, d_umi_q, ws, workspace_bytes, sms, st);
}

// Debug/bench variant: same as NR_MODE_FILTERED but with the counting kernel; fills the five
// counters at the head of the workspace.
extern "C" int nr_match_device_counted(const nr_whitelist_t *wl, const void *d_bases,
                                       const uint8_t *d_meta, const uint64_t *d_nmask, uint64_t n,
                                       int min_score, int32_t *d_idx, int8_t *d_score,
                                       uint8_t *d_nbest, uint8_t *d_flags, uint8_t *d_umi_q,
                                       void *d_workspace, size_t workspace_bytes, void *stream)
{
    if (!wl || n == 0 || !d_workspace || workspace_bytes < nr_match_workspace_bytes(wl, n, 0) ||
        resolve_mode(wl, min_score, NR_MODE_FILTERED) < 0) {
        nr_set_error("nr_match_device_counted: bad arguments");
        return NR_EINVAL;
    }
    DeviceGuard guard(wl->device);
    cudaStream_t st = (cudaStream_t)stream;
    uint8_t *ws = (uint8_t *)d_workspace;
    NR_CHECK_CUDA(cudaMemsetAsync(ws, 0, NR_WS_ZERO, st));
    uint32_t *listA = (uint32_t *)(ws + NR_WS_HEADER);
    int sms = 148;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, wl->device);
    if (wl->has_anchor) {
        int rc = nr_launch_anchored(wl, d_bases, d_meta, d_nmask, n, min_score, 0, d_idx, d_score, d_nbest, d_flags,
                                    d_umi_q, listA, (uint32_t *)(ws + 64), (unsigned long long *)(ws + 72),
                                    (unsigned long long *)ws, st);
        if (rc != NR_OK) return rc;
        return resolve_chain(wl, d_bases, d_meta, d_nmask, n, min_score, false, true, d_idx, d_score,
                             d_nbest, d_flags, d_umi_q, ws, workspace_bytes, sms, st);
    }
    int rc = nr_launch_filtered(wl, d_bases, d_meta, d_nmask, n, min_score, 0, d_idx, d_score,
                                d_nbest, d_flags, d_umi_q, listA, (uint32_t *)(ws + 64),
                                listA + 3 * n, (uint32_t *)(ws + 84),
                                (unsigned long long *)(ws + 72), (unsigned long long *)(ws + 104),
                                (unsigned long long *)ws, nullptr, st);
    if (rc != NR_OK) return rc;
    return resolve_chain(wl, d_bases, d_meta, d_nmask, n, min_score, false, true, d_idx, d_score,
                         d_nbest, d_flags, d_umi_q, ws, workspace_bytes, sms, st);
}

extern "C" int nr_match_counters(const void *d_workspace, uint64_t *c5, void *stream)
{
    if (!d_workspace || !c5) { nr_set_error("nr_match_counters: null pointer"); return NR_EINVAL; }
    uint64_t h[9];
    NR_CHECK_CUDA(cudaMemcpyAsync(h, d_workspace, sizeof(h), cudaMemcpyDeviceToHost,
                                  (cudaStream_t)stream));
    NR_CHECK_CUDA(cudaStreamSynchronize((cudaStream_t)stream));
    for (int i = 0; i < 4; i++) c5[i] = h[i];
    c5[4] = (uint32_t)(h[8] & 0xFFFFFFFFu);   // list count (candidates sent to the exhaustive kernel)
    return NR_OK;
}

extern "C" int nr_match_tier_counts(const void *d_workspace, uint64_t *t4, void *stream)
{
    if (!d_workspace || !t4) { nr_set_error("nr_match_tier_counts: null pointer"); return NR_EINVAL; }
    uint64_t h[11];
    NR_CHECK_CUDA(cudaMemcpyAsync(h, d_workspace, sizeof(h), cudaMemcpyDeviceToHost,
                                  (cudaStream_t)stream));
    NR_CHECK_CUDA(cudaStreamSynchronize((cudaStream_t)stream));
    t4[0] = (uint32_t)(h[8] & 0xFFFFFFFFu);     // list A count (offset 64)
    t4[1] = h[5];
    t4[2] = h[6];
    t4[3] = (uint32_t)(h[10] & 0xFFFFFFFFu);    // list C count (offset 80)
    return NR_OK;
}

// ---- pinned host memory for callers that want truly asynchronous copies ----------------------
extern "C" void *nr_host_alloc(size_t bytes)
{
    void *p = nullptr;
    if (cudaHostAlloc(&p, bytes ? bytes : 1, cudaHostAllocDefault) != cudaSuccess) {
        nr_set_error("cudaHostAlloc(%zu) failed", bytes);
        return nullptr;
    }
    return p;
}
extern "C" void nr_host_free(void *p) { if (p) cudaFreeHost(p); }

// ---- host-buffer entry point ---------------------------------------------------------------------
// Chunks of CHUNK_CAND candidates rotate through NSLOT slots, each with its own stream, so the
// H2D copy of chunk k+1 and the D2H copy of chunk k-1 overlap the kernels of chunk k.  Caller
// buffers that are pinned (nr_host_alloc, cudaHostAlloc, cudaHostRegister) are copied from / to
// directly; pageable buffers go through the slot's pinned staging area.
namespace {

// candidates per chunk: 2^20 by default (NR_HOST_CHUNK_LOG2 = 16..22 overrides it, for experiments:
// 2^18 1.67e8, 2^19 1.87e8, 2^20 1.95e8, 2^21 1.95e8 candidates/s end to end on the bench workload --
// every chunk pays the N pass and four near-empty tail launches)
uint64_t chunk_cand()
{
    const char *e = getenv("NR_HOST_CHUNK_LOG2");
    long v = e ? atol(e) : 20;
    if (v < 16 || v > 22) v = 20;
    return 1ull << v;
}
const uint64_t CHUNK_CAND = chunk_cand();
const uint64_t CHUNK_BYTES = CHUNK_CAND * 64;             // sequence bytes per chunk
constexpr int NSLOT = 4;

struct Slot {
    cudaStream_t st = nullptr;
    cudaEvent_t done = nullptr;
    uint8_t *h_in = nullptr;        // pinned: sequence bytes, then offsets
    uint8_t *h_out = nullptr;       // pinned: idx | score | nbest | flags | umi
    uint8_t *d_seqs = nullptr;
    uint64_t *d_off = nullptr;
    uint8_t *d_bases = nullptr, *d_meta = nullptr;
    uint64_t *d_nmask = nullptr;
    int32_t *d_idx = nullptr;
    int8_t *d_score = nullptr;
    uint8_t *d_nbest = nullptr, *d_flags = nullptr, *d_umi = nullptr;
    uint8_t *d_ws = nullptr;
    size_t ws_bytes = 0;
    uint64_t c0 = 0, cn = 0;        // chunk in flight
    bool busy = false, staged_out = false;
};

struct HostCtx {
    std::mutex mu;
    Slot slot[NSLOT];
    bool ready = false;
};

// Staging copy between the caller's pageable arrays and the pinned slots, spread over a few
// host threads (one thread moves ~10 GB/s, PCIe 5 x16 takes ~50 GB/s).
void staged_copy(void *dst, const void *src, size_t n)
{
    constexpr size_t MIN_PER_THREAD = 4u << 20;
    // host threads this process may use for staging: its share of the cores when several ranks
    // run on one box (torchrun exports LOCAL_WORLD_SIZE), never more than 6
    static const size_t budget = [] {
        unsigned hw = std::thread::hardware_concurrency();
        size_t share = hw ? hw : 1;
        const char *lws = getenv("LOCAL_WORLD_SIZE");
        const long ranks = lws ? atol(lws) : 1;
        if (ranks > 1) share = std::max<size_t>(1, share / (size_t)ranks);
        return std::min<size_t>(share, 6);
    }();
    size_t want = n / MIN_PER_THREAD;
    size_t nt = std::min<size_t>(want, budget);
    if (nt <= 1) { memcpy(dst, src, n); return; }
    std::vector<std::thread> th;
    size_t per = (n + nt - 1) / nt;
    per = (per + 4095) & ~(size_t)4095;
    for (size_t k = 1; k < nt; k++) {
        size_t a = k * per;
        if (a >= n) break;
        size_t len = std::min(per, n - a);
        th.emplace_back([=] { memcpy((char *)dst + a, (const char *)src + a, len); });
    }
    memcpy(dst, src, std::min(per, n));
    for (auto &t : th) t.join();
}

bool is_pinned(const void *p)
{
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess) { cudaGetLastError(); return false; }
    return a.type == cudaMemoryTypeHost;
}

int slot_init(Slot &s, const nr_whitelist *wl)
{
    NR_CHECK_CUDA(cudaStreamCreateWithFlags(&s.st, cudaStreamNonBlocking));
    NR_CHECK_CUDA(cudaEventCreateWithFlags(&s.done, cudaEventDisableTiming));
    size_t in_bytes = CHUNK_BYTES + (CHUNK_CAND + 1) * sizeof(uint64_t);
    NR_CHECK_CUDA(cudaHostAlloc((void **)&s.h_in, in_bytes, cudaHostAllocDefault));
    NR_CHECK_CUDA(cudaHostAlloc((void **)&s.h_out, CHUNK_CAND * 8, cudaHostAllocDefault));
    NR_CHECK_CUDA(cudaMalloc((void **)&s.d_seqs, CHUNK_BYTES));
    NR_CHECK_CUDA(cudaMalloc((void **)&s.d_off, (CHUNK_CAND + 1) * sizeof(uint64_t)));
    NR_CHECK_CUDA(cudaMalloc((void **)&s.d_bases, CHUNK_CAND * 16));
    NR_CHECK_CUDA(cudaMalloc((void **)&s.d_meta, CHUNK_CAND));
    NR_CHECK_CUDA(cudaMalloc((void **)&s.d_nmask, CHUNK_CAND * 8));
    NR_CHECK_CUDA(cudaMalloc((void **)&s.d_idx, CHUNK_CAND * 8));   // idx|score|nbest|flags|umi
    s.d_score = (int8_t *)(s.d_idx + CHUNK_CAND);
    s.d_nbest = (uint8_t *)s.d_score + CHUNK_CAND;
    s.d_flags = s.d_nbest + CHUNK_CAND;
    s.d_umi = s.d_flags + CHUNK_CAND;
    s.ws_bytes = nr_match_workspace_bytes(wl, CHUNK_CAND, NR_MODE_AUTO);
    NR_CHECK_CUDA(cudaMalloc((void **)&s.d_ws, s.ws_bytes));
    return NR_OK;
}

void slot_free(Slot &s)
{
    if (s.st) cudaStreamDestroy(s.st);
    if (s.done) cudaEventDestroy(s.done);
    cudaFreeHost(s.h_in); cudaFreeHost(s.h_out);
    cudaFree(s.d_seqs); cudaFree(s.d_off); cudaFree(s.d_bases); cudaFree(s.d_meta);
    cudaFree(s.d_nmask); cudaFree(s.d_idx); cudaFree(s.d_ws);
    s = Slot();
}

// wait for the slot's chunk; copy it out of the staging area when the caller's arrays are pageable
int slot_collect(Slot &s, int32_t *idx, int8_t *score, uint8_t *nbest, uint8_t *flags,
                 uint8_t *umi_q)
{
    if (!s.busy) return NR_OK;
    NR_CHECK_CUDA(cudaEventSynchronize(s.done));
    if (s.staged_out) {
        const uint8_t *o = s.h_out;
        staged_copy(idx + s.c0, o, s.cn * 4); o += s.cn * 4;
        memcpy(score + s.c0, o, s.cn); o += s.cn;
        memcpy(nbest + s.c0, o, s.cn); o += s.cn;
        memcpy(flags + s.c0, o, s.cn); o += s.cn;
        memcpy(umi_q + s.c0, o, s.cn);
    }
    s.busy = false;
    return NR_OK;
}

}  // namespace

void nr_host_ctx_destroy(void *p)
{
    HostCtx *c = (HostCtx *)p;
    if (!c) return;
    for (int k = 0; k < NSLOT; k++) slot_free(c->slot[k]);
    delete c;
}

extern "C" int nr_match_host(const nr_whitelist_t *wlc, const char *seqs, const uint64_t *offsets,
                             uint64_t n, int min_score, int mode, int32_t *idx, int8_t *score,
                             uint8_t *nbest, uint8_t *flags, uint8_t *umi_q)
{
    nr_whitelist *wl = const_cast<nr_whitelist *>(wlc);
    if (!wl) { nr_set_error("nr_match_host: null whitelist"); return NR_EINVAL; }
    if (n == 0) return NR_OK;
    if (!seqs || !offsets || !idx || !score || !nbest || !flags || !umi_q) {
        nr_set_error("nr_match_host: null pointer");
        return NR_EINVAL;
    }
    if (resolve_mode(wl, min_score, mode) < 0 ||
        (mode != NR_MODE_AUTO && mode != NR_MODE_EXHAUSTIVE && mode != NR_MODE_FILTERED)) {
        nr_set_error("nr_match_host: mode %d not usable with this whitelist / min_score", mode);
        return NR_EUNSUPPORTED;
    }
    DeviceGuard guard(wl->device);
    static std::mutex create_mu;
    {
        std::lock_guard<std::mutex> g(create_mu);
        if (!wl->host_ctx) {
            HostCtx *c = new (std::nothrow) HostCtx();
            if (!c) { nr_set_error("out of host memory"); return NR_ENOMEM; }
            wl->host_ctx = c;
        }
    }
    HostCtx *ctx = (HostCtx *)wl->host_ctx;
    std::lock_guard<std::mutex> g(ctx->mu);
    if (!ctx->ready) {
        for (int k = 0; k < NSLOT; k++) {
            int rc = slot_init(ctx->slot[k], wl);
            if (rc != NR_OK) {
                for (int j = 0; j < NSLOT; j++) slot_free(ctx->slot[j]);
                return rc;
            }
        }
        ctx->ready = true;
    }
    const bool in_pinned = is_pinned(seqs) && is_pinned(offsets);
    const bool out_pinned = is_pinned(idx) && is_pinned(score) && is_pinned(nbest) &&
                            is_pinned(flags) && is_pinned(umi_q);
    uint64_t c0 = 0;
    int k = 0, n_chunk = 0;
    int rc = NR_OK;
    while (c0 < n) {
        // chunk = as many candidates as fit both limits; the first chunks are smaller so that the
        // GPU starts working before much of the input has crossed PCIe
        const uint64_t want = n_chunk == 0 ? CHUNK_CAND / 8 : (n_chunk == 1 ? CHUNK_CAND / 4 :
                              (n_chunk == 2 ? CHUNK_CAND / 2 : CHUNK_CAND));
        n_chunk++;
        uint64_t c1 = std::min(n, c0 + want);
        if (offsets[c1] - offsets[c0] > CHUNK_BYTES) {
            const uint64_t *hi = std::upper_bound(offsets + c0, offsets + c1 + 1,
                                                  offsets[c0] + CHUNK_BYTES);
            c1 = (uint64_t)(hi - offsets) - 1;
            if (c1 <= c0) {
                nr_set_error("nr_match_host: candidate %llu is longer than %llu bytes",
                             (unsigned long long)c0, (unsigned long long)CHUNK_BYTES);
                rc = NR_EINVAL;
                break;
            }
        }
        Slot &s = ctx->slot[k];
        if ((rc = slot_collect(s, idx, score, nbest, flags, umi_q)) != NR_OK) break;
        const uint64_t cn = c1 - c0, b0 = offsets[c0], nb = offsets[c1] - b0;
        const void *src_seq = seqs + b0;
        const void *src_off = offsets + c0;
        if (!in_pinned) {
            staged_copy(s.h_in, seqs + b0, nb);
            memcpy(s.h_in + CHUNK_BYTES, offsets + c0, (cn + 1) * sizeof(uint64_t));
            src_seq = s.h_in;
            src_off = s.h_in + CHUNK_BYTES;
        }
        cudaError_t e;
        e = cudaMemcpyAsync(s.d_seqs, src_seq, nb, cudaMemcpyHostToDevice, s.st);
        if (e == cudaSuccess)
            e = cudaMemcpyAsync(s.d_off, src_off, (cn + 1) * sizeof(uint64_t),
                                cudaMemcpyHostToDevice, s.st);
        if (e != cudaSuccess) { nr_set_error("H2D failed: %s", cudaGetErrorString(e)); rc = NR_ECUDA; break; }
        // offsets are absolute: bias the sequence pointer instead of rewriting them
        rc = nr_pack_device(s.d_seqs - b0, s.d_off, cn, s.d_bases, s.d_meta, s.d_nmask, s.st);
        if (rc != NR_OK) break;
        rc = nr_match_device(wl, s.d_bases, s.d_meta, s.d_nmask, cn, min_score, mode, s.d_idx,
                             s.d_score, s.d_nbest, s.d_flags, s.d_umi, s.d_ws, s.ws_bytes, s.st);
        if (rc != NR_OK) break;
        uint8_t *o = s.h_out;
        void *dst[5] = {o, o + cn * 4, o + cn * 5, o + cn * 6, o + cn * 7};
        if (out_pinned) {
            dst[0] = idx + c0; dst[1] = score + c0; dst[2] = nbest + c0; dst[3] = flags + c0;
            dst[4] = umi_q + c0;
        }
        e = cudaMemcpyAsync(dst[0], s.d_idx, cn * 4, cudaMemcpyDeviceToHost, s.st);
        if (e == cudaSuccess) e = cudaMemcpyAsync(dst[1], s.d_score, cn, cudaMemcpyDeviceToHost, s.st);
        if (e == cudaSuccess) e = cudaMemcpyAsync(dst[2], s.d_nbest, cn, cudaMemcpyDeviceToHost, s.st);
        if (e == cudaSuccess) e = cudaMemcpyAsync(dst[3], s.d_flags, cn, cudaMemcpyDeviceToHost, s.st);
        if (e == cudaSuccess) e = cudaMemcpyAsync(dst[4], s.d_umi, cn, cudaMemcpyDeviceToHost, s.st);
        if (e == cudaSuccess) e = cudaEventRecord(s.done, s.st);
        if (e != cudaSuccess) { nr_set_error("D2H failed: %s", cudaGetErrorString(e)); rc = NR_ECUDA; break; }
        s.c0 = c0; s.cn = cn; s.busy = true; s.staged_out = !out_pinned;
        c0 = c1;
        k = (k + 1) % NSLOT;
    }
    for (int j = 0; j < NSLOT; j++) {
        int r2 = slot_collect(ctx->slot[(k + j) % NSLOT], idx, score, nbest, flags, umi_q);
        if (rc == NR_OK) rc = r2;
    }
    if (rc != NR_OK) {
        cudaDeviceSynchronize();
        for (int j = 0; j < NSLOT; j++) ctx->slot[j].busy = false;
    }
    return rc;
}
