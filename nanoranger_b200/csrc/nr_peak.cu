// nr_peak.cu -- integer ALU-pipe roofline denominator.
// MEASURED_PEAKS.json holds HBM and bf16 tensor peaks only; the matcher is bound by the
// LOP3/IADD3/SHF (ALU) pipe, so the denominator is measured here: every thread runs eight
// independent dependent chains of (LOP3, IADD3) pairs, enough warps per SM to saturate issue.
#include "nr_common.cuh"

#define NR_PEAK_CHAINS 8
#define NR_PEAK_UNROLL 16

// MODE 0: LOP3 + SHF per step (both on the ALU pipe) -> ALU-pipe peak
// MODE 1: LOP3 + add per step (ptxas places the add on the FMA pipe as IMAD.IADD) -> the
//         integer issue rate with both pipes busy
template <int MODE>
__global__ void __launch_bounds__(256) nr_int_peak_kernel(uint32_t *out, int iters, uint32_t seed)
{
    uint32_t a[NR_PEAK_CHAINS], b = seed ^ threadIdx.x, c = seed * 2654435761u + blockIdx.x;
#pragma unroll
    for (int k = 0; k < NR_PEAK_CHAINS; k++) a[k] = seed + k * 0x9E3779B9u + threadIdx.x;
#pragma unroll 1
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int u = 0; u < NR_PEAK_UNROLL; u++) {
#pragma unroll
            for (int k = 0; k < NR_PEAK_CHAINS; k++) {
                // one LOP3 (xor-and-or of three registers) and one IADD3 per step
                a[k] = ((a[k] ^ b) | (a[k] & c));
                if (MODE == 0) a[k] = __funnelshift_l(a[k], c, 7);
                else a[k] = a[k] + b + c;
            }
        }
        b += 0x01000193u;
    }
    uint32_t r = 0;
#pragma unroll
    for (int k = 0; k < NR_PEAK_CHAINS; k++) r ^= a[k];
    if (r == 0x12345678u) out[0] = r;   // keeps the chains alive; practically never true
}

static int int_peak_mode(int device, int iters, int mode, double *ops_per_s, double *ms)
{
    if (iters < 1 || !ops_per_s) { nr_set_error("nr_int_peak: bad arguments"); return NR_EINVAL; }
    int prev = -1;
    cudaGetDevice(&prev);
    NR_CHECK_CUDA(cudaSetDevice(device));
    int sms = 0;
    NR_CHECK_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device));
    uint32_t *d_out = nullptr;
    NR_CHECK_CUDA(cudaMalloc(&d_out, 64));
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    int grid = sms * 8;
    if (mode == 0) nr_int_peak_kernel<0><<<grid, 256>>>(d_out, iters / 4 + 1, 1u);   // warm-up
    else nr_int_peak_kernel<1><<<grid, 256>>>(d_out, iters / 4 + 1, 1u);
    float best = 1e30f;
    for (int rep = 0; rep < 3; rep++) {
        cudaEventRecord(e0);
        if (mode == 0) nr_int_peak_kernel<0><<<grid, 256>>>(d_out, iters, 7u + rep);
        else nr_int_peak_kernel<1><<<grid, 256>>>(d_out, iters, 7u + rep);
        cudaEventRecord(e1);
        cudaError_t e = cudaEventSynchronize(e1);
        if (e != cudaSuccess) {
            nr_set_error("nr_int_peak: %s", cudaGetErrorString(e));
            cudaFree(d_out);
            if (prev >= 0) cudaSetDevice(prev);
            return NR_ECUDA;
        }
        float t = 0;
        cudaEventElapsedTime(&t, e0, e1);
        if (t < best) best = t;
    }
    double ops = (double)grid * 256.0 * (double)iters * NR_PEAK_UNROLL * NR_PEAK_CHAINS * 2.0;
    *ops_per_s = ops / (best * 1e-3);
    if (ms) *ms = best;
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    cudaFree(d_out);
    if (prev >= 0) cudaSetDevice(prev);
    return NR_OK;
}

// ALU-pipe peak (LOP3 + SHF chains): thread-ops per second
extern "C" int nr_int_peak(int device, int iters, double *ops_per_s, double *ms)
{
    return int_peak_mode(device, iters, 0, ops_per_s, ms);
}

// integer issue peak with the add half on the FMA pipe (LOP3 + IMAD.IADD chains)
extern "C" int nr_int_peak_dual(int device, int iters, double *ops_per_s, double *ms)
{
    return int_peak_mode(device, iters, 1, ops_per_s, ms);
}
