// nr_hwsearch.cu -- adapter-motif search in read windows: bit-parallel (Myers/Hyyro) infix edit
// distance, one window per thread.
//
// Replaces the edlib calls of the reference's candidate extractors:
//   edlib.align(const, window, "HW", "locations", k[, ad_seq])
//   utils.py:134, 271 (k = 6, N wildcard, last location), :345 (k = 5, first), :437 (k = 2, no
//   wildcard, first), :1051 (k = 2, first), :1367 (k = 3, first).
// edlib (Sosic & Sikic 2017; the reference pins edlib 1.3.9 in requirements.txt, not vendored)
// defines, for mode HW ("infix"): editDistance = min over all substrings of the target of the
// unit-cost edit distance to the query, -1 if > k; end locations = every target position where an
// optimal alignment ends (ascending); the start reported for an end location is the SMALLEST
// start of an optimal alignment ending there (edlib aligns the reversed query against the
// reversed target prefix in SHW mode and takes the last position).  additionalEqualities
// [("N","A"),("N","T"),("N","G"),("N","C")] (utils.py:15) make N equal to every base, on
// either side.  Bytes are otherwise compared exactly (edlib is case-sensitive).
//
// Pattern length m <= 64 (the reference's motifs are 18..53 nt): one 64-bit word per column.
#include "nr_common.cuh"

namespace {

struct HwParams {
    uint8_t pat[64];
    int m, k, wildcard;
};

__device__ __forceinline__ bool hw_eq(uint8_t p, uint32_t c, int wildcard)
{
    if (p == c) return true;
    if (!wildcard) return false;
    const bool pb = p == 'A' || p == 'C' || p == 'G' || p == 'T';
    const bool cb = c == 'A' || c == 'C' || c == 'G' || c == 'T';
    return (p == 'N' && cb) || (c == 'N' && pb);
}

// one Myers column: updates Pv/Mv, returns the change of the bottom cell (+1, 0, -1)
// carry_in = 0: free start in the text (HW); 1: the first row grows by one per column (SHW)
__device__ __forceinline__ int myers_step(uint64_t Eq, uint64_t &Pv, uint64_t &Mv, uint64_t top,
                                          uint64_t carry_in)
{
    const uint64_t Xv = Eq | Mv;
    const uint64_t Xh = (((Eq & Pv) + Pv) ^ Pv) | Eq;
    uint64_t Ph = Mv | ~(Xh | Pv);
    uint64_t Mh = Pv & Xh;
    const int d = (int)((Ph & top) != 0) - (int)((Mh & top) != 0);
    Ph = (Ph << 1) | carry_in;
    Mh <<= 1;
    Pv = Mh | ~(Xv | Ph);
    Mv = Ph & Xv;
    return d;
}

__global__ void __launch_bounds__(128)
nr_hw_search_kernel(const uint8_t *__restrict__ text, const uint64_t *__restrict__ offsets,
                    uint64_t n, HwParams P, int8_t *__restrict__ o_ed, int32_t *__restrict__ o_first,
                    int32_t *__restrict__ o_last, int32_t *__restrict__ o_nloc)
{
    __shared__ uint64_t peq[256], rpeq[256];
    for (uint32_t c = threadIdx.x; c < 256; c += blockDim.x) {
        uint64_t a = 0, b = 0;
        for (int i = 0; i < P.m; i++) {
            if (hw_eq(P.pat[i], c, P.wildcard)) a |= 1ull << i;
            if (hw_eq(P.pat[P.m - 1 - i], c, P.wildcard)) b |= 1ull << i;
        }
        peq[c] = a; rpeq[c] = b;
    }
    __syncthreads();
    // grid-stride over the windows: the two tables above are built once per block, not per 128
    // windows
    for (uint64_t w = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; w < n;
         w += (uint64_t)gridDim.x * blockDim.x) {
    const uint8_t *t = text + offsets[w];
    const int64_t len = (int64_t)(offsets[w + 1] - offsets[w]);
    const int m = P.m;
    const uint64_t top = 1ull << (m - 1);
    const uint64_t ones = m == 64 ? ~0ull : ((1ull << m) - 1ull);
    uint64_t Pv = ones, Mv = 0;
    int score = m, best = m + 1;
    int64_t first = -1, last = -1;
    int nloc = 0;
    for (int64_t j = 0; j < len; j++) {
        score += myers_step(peq[t[j]], Pv, Mv, top, 0ull);
        if (score < best) { best = score; first = j; last = j; nloc = 1; }
        else if (score == best) { last = j; nloc++; }
    }
    if (len == 0 || best > P.k) {
        o_ed[w] = -1; o_nloc[w] = 0;
        o_first[2 * w] = -1; o_first[2 * w + 1] = -1; o_last[2 * w] = -1; o_last[2 * w + 1] = -1;
        continue;
    }
    // smallest start of an optimal alignment ending at e: reversed pattern against the reversed
    // prefix t[e], t[e-1], ...; bottom-row value == best at the farthest column
    int64_t starts[2];
    const int64_t ends[2] = {first, last};
#pragma unroll 1
    for (int which = 0; which < 2; which++) {
        const int64_t e = ends[which];
        if (which == 1 && last == first) { starts[1] = starts[0]; break; }
        uint64_t pv = ones, mv = 0;
        int sc = m;
        int64_t far = 0;
        const int64_t lim = min((int64_t)(m + best), e + 1);
        for (int64_t r = 0; r < lim; r++) {
            sc += myers_step(rpeq[t[e - r]], pv, mv, top, 1ull);
            if (sc == best) far = r;
        }
        starts[which] = e - far;
    }
    o_ed[w] = (int8_t)best; o_nloc[w] = nloc;
    o_first[2 * w] = (int32_t)starts[0]; o_first[2 * w + 1] = (int32_t)first;
    o_last[2 * w] = (int32_t)starts[1]; o_last[2 * w + 1] = (int32_t)last;
    }
}

int check_args(const char *pattern, int m, int k)
{
    if (!pattern || m < 1 || m > 64) { nr_set_error("nr_hw_search: pattern length must be 1..64"); return NR_EINVAL; }
    if (k < 0 || k >= m || k > 127) { nr_set_error("nr_hw_search: need 0 <= k < pattern length"); return NR_EINVAL; }
    return NR_OK;
}

}  // namespace

extern "C" int nr_hw_search_device(const uint8_t *d_text, const uint64_t *d_offsets, uint64_t n,
                                   const char *pattern, int m, int k, int wildcard_n,
                                   int8_t *d_ed, int32_t *d_first, int32_t *d_last,
                                   int32_t *d_nloc, void *stream)
{
    int rc = check_args(pattern, m, k);
    if (rc != NR_OK) return rc;
    if (n == 0) return NR_OK;
    if (!d_text || !d_offsets || !d_ed || !d_first || !d_last || !d_nloc) {
        nr_set_error("nr_hw_search_device: null pointer");
        return NR_EINVAL;
    }
    HwParams P;
    for (int i = 0; i < 64; i++) P.pat[i] = i < m ? (uint8_t)pattern[i] : 0;
    P.m = m; P.k = k; P.wildcard = wildcard_n ? 1 : 0;
    int sms = 148, dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const uint64_t want = (n + 127) / 128, cap = (uint64_t)sms * 16;
    const uint64_t blocks = want < cap ? want : cap;
    nr_hw_search_kernel<<<(unsigned)blocks, 128, 0, (cudaStream_t)stream>>>(
        d_text, d_offsets, n, P, d_ed, d_first, d_last, d_nloc);
    NR_CHECK_CUDA(cudaGetLastError());
    return NR_OK;
}

extern "C" int nr_hw_search_host(const char *text, const uint64_t *offsets, uint64_t n,
                                 const char *pattern, int m, int k, int wildcard_n, int8_t *ed,
                                 int32_t *first, int32_t *last, int32_t *nloc, int device)
{
    int rc = check_args(pattern, m, k);
    if (rc != NR_OK) return rc;
    if (n == 0) return NR_OK;
    if (!text || !offsets || !ed || !first || !last || !nloc) {
        nr_set_error("nr_hw_search_host: null pointer");
        return NR_EINVAL;
    }
    int prev = -1;
    cudaGetDevice(&prev);
    NR_CHECK_CUDA(cudaSetDevice(device));
    const uint64_t b0 = offsets[0], nb = offsets[n] - b0;
    uint8_t *d_text = nullptr, *d_out = nullptr;
    uint64_t *d_off = nullptr;
    cudaStream_t st = nullptr;
    rc = NR_ECUDA;
    do {
        if (cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking) != cudaSuccess) break;
        if (cudaMalloc((void **)&d_text, nb ? nb : 1) != cudaSuccess) break;
        if (cudaMalloc((void **)&d_off, (n + 1) * sizeof(uint64_t)) != cudaSuccess) break;
        if (cudaMalloc((void **)&d_out, n * 24) != cudaSuccess) break;      // first|last|nloc|ed
        if (cudaMemcpyAsync(d_text, text + b0, nb, cudaMemcpyHostToDevice, st) != cudaSuccess) break;
        if (cudaMemcpyAsync(d_off, offsets, (n + 1) * sizeof(uint64_t), cudaMemcpyHostToDevice, st) != cudaSuccess) break;
        int32_t *d_first = (int32_t *)d_out, *d_last = d_first + 2 * n, *d_nloc = d_last + 2 * n;
        int8_t *d_ed = (int8_t *)(d_nloc + n);
        // offsets are absolute: bias the text pointer
        rc = nr_hw_search_device(d_text - b0, d_off, n, pattern, m, k, wildcard_n, d_ed, d_first,
                                 d_last, d_nloc, st);
        if (rc != NR_OK) break;
        rc = NR_ECUDA;
        if (cudaMemcpyAsync(first, d_first, n * 8, cudaMemcpyDeviceToHost, st) != cudaSuccess) break;
        if (cudaMemcpyAsync(last, d_last, n * 8, cudaMemcpyDeviceToHost, st) != cudaSuccess) break;
        if (cudaMemcpyAsync(nloc, d_nloc, n * 4, cudaMemcpyDeviceToHost, st) != cudaSuccess) break;
        if (cudaMemcpyAsync(ed, d_ed, n, cudaMemcpyDeviceToHost, st) != cudaSuccess) break;
        if (cudaStreamSynchronize(st) != cudaSuccess) break;
        rc = NR_OK;
    } while (0);
    if (rc == NR_ECUDA) nr_set_error("nr_hw_search_host: %s", cudaGetErrorString(cudaGetLastError()));
    cudaFree(d_text); cudaFree(d_off); cudaFree(d_out);
    if (st) cudaStreamDestroy(st);
    if (prev >= 0) cudaSetDevice(prev);
    return rc;
}
