// nr_pack.cu -- ASCII candidate batch -> 2-bit packed records on the device.
// Replaces STAR's FASTA read loading (scripts/barcode_align.sh:16,35) for the matcher.
#include "nr_common.cuh"

// One thread per candidate; a warp's 32 candidates are contiguous in `seqs`, so the byte
// loads of neighbouring lanes share 128 B lines (L1 absorbs the overlap).
__global__ void __launch_bounds__(256)
nr_pack_kernel(const uint8_t *__restrict__ seqs, const uint64_t *__restrict__ offsets,
               uint64_t n, uint4 *__restrict__ bases, uint8_t *__restrict__ meta,
               uint64_t *__restrict__ nmask)
{
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint64_t b = offsets[i], e = offsets[i + 1];
    uint64_t len = e - b;
    uint32_t w[4] = {0, 0, 0, 0};
    uint64_t nm = 0;
    if (len > NR_MAX_QUERY) {
        bases[i] = make_uint4(0, 0, 0, 0);
        meta[i] = 0xFF;
        nmask[i] = 0;
        return;
    }
    const uint8_t *s = seqs + b;
#pragma unroll
    for (int k = 0; k < 4; k++) {
        uint32_t acc = 0;
#pragma unroll
        for (int t = 0; t < 16; t++) {
            int p = k * 16 + t;
            if ((uint64_t)p < len) {
                uint32_t c = s[p];
                // ASCII trick: A 0x41, C 0x43, G 0x47, T 0x54 (case bit 0x20 ignored)
                uint32_t u = c & 0xDFu;
                uint32_t code = (u >> 1) & 3u;          // A->0 C->1 G->3 T->2
                code ^= code >> 1;                       // A->0 C->1 G->2 T->3
                bool ok = (u == 'A') | (u == 'C') | (u == 'G') | (u == 'T');
                if (!ok) { nm |= 1ull << p; code = 0; }
                acc |= code << (2 * t);
            }
        }
        w[k] = acc;
    }
    bases[i] = make_uint4(w[0], w[1], w[2], w[3]);
    meta[i] = (uint8_t)(len | (nm ? 0x80u : 0u));
    nmask[i] = nm;
}

extern "C" int nr_pack_device(const uint8_t *d_seqs, const uint64_t *d_offsets, uint64_t n,
                              void *d_bases, uint8_t *d_meta, uint64_t *d_nmask, void *stream)
{
    if (n == 0) return NR_OK;
    if (!d_seqs || !d_offsets || !d_bases || !d_meta || !d_nmask) {
        nr_set_error("nr_pack_device: null pointer");
        return NR_EINVAL;
    }
    uint64_t blocks = (n + 255) / 256;
    nr_pack_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(
        d_seqs, d_offsets, n, (uint4 *)d_bases, d_meta, d_nmask);
    NR_CHECK_CUDA(cudaGetLastError());
    return NR_OK;
}
