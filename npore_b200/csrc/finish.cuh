// finish.cuh -- per-item epilogue on device:
//   gather_kernel       concatenates the chunks' op strings (aln.pyx:742 `full_aln += aln[::-1]`)
//   standardize_kernel  src/bam.pyx:65-78: X,= -> M; push_indels_left(D, ref); push_inss_thru_dels;
//                       push_indels_left(I, seq); push_inss_thru_dels (src/cig.pyx:102-192; the reference's
//                       `while True` body runs exactly once because old_cig aliases int_cig); 'ID' -> 'M'
//   rle_kernel          src/cig.pyx:13-38 collapse_cigar as BAM-style words (len<<4 | op)
// The standardisation passes are sequential and data dependent per item: one thread per item.
#pragma once
#include "common.cuh"

#define FIN_THREADS 128

struct FinishArgs {
    const ItemDesc *items; int n_items;
    const ChunkOut *chunk_out;
    const uint8_t *scratch;       // right-aligned chunk pieces
    uint8_t *ops;                 // final per-item op strings (item region at out_off)
    int32_t *item_len;            // ops per item
    int32_t *item_status;
    const uint8_t *ref_codes, *seq_codes;
    uint32_t *rle; int32_t *rle_len;
};

__global__ void __launch_bounds__(FIN_THREADS) gather_kernel(const FinishArgs a)
{
    __shared__ int s_status;
    const int it = blockIdx.x;
    if (it >= a.n_items) return;
    const ItemDesc &I = a.items[it];
    if (threadIdx.x == 0) s_status = I.status;
    __syncthreads();
    uint8_t *dst = a.ops + I.out_off;
    const uint8_t *src = a.scratch + I.out_off;
    int off = 0;
    for (int k = 0; k < I.n_chunks; k++) {
        const ChunkOut co = a.chunk_out[I.chunk_first + k];
        for (int t = threadIdx.x; t < co.len; t += FIN_THREADS) dst[off + t] = src[co.start + t];
        if (threadIdx.x == 0 && co.status && !s_status) s_status = co.status;
        off += co.len;
    }
    __syncthreads();
    if (threadIdx.x == 0) { a.item_len[it] = off; a.item_status[it] = s_status; }
}

// cig.pyx:102-159 on an op array holding only M(0) / I(1) / D(2): every op the run is pushed through is an M,
// so the rotation of cig.pyx:141-149 reduces to rewriting [cp-k, cp+len) as len push_ops followed by k M's.
__device__ __forceinline__ void dev_push_indels_left(uint8_t *cig, int n, const uint8_t *__restrict__ seq, uint8_t push_op)
{
    int sp = 0, cp = 0;
    while (cp < n) {
        const uint8_t op = cig[cp];
        if (op != push_op) { cp++; if (op == 0) sp++; continue; }
        int len = 1;
        while (cp + len < n && cig[cp + len] == push_op) len++;
        int k = 0;
        while (cp - k > 0 && sp - k > 0 && seq[sp - k - 1] == seq[sp - k - 1 + len] && cig[cp - k - 1] == 0) k++;
        if (k) {
            for (int t = 0; t < len; t++) cig[cp - k + t] = push_op;
            for (int t = 0; t < k; t++) cig[cp - k + len + t] = 0;
        }
        cp += len; sp += len;
    }
}

// cig.pyx:164-192
__device__ __forceinline__ void dev_push_inss_thru_dels(uint8_t *cig, int n)
{
    for (int i = 0; i + 1 < n; i++) {
        if (cig[i] == 2 && cig[i + 1] == 1) {
            int di = i - 1;
            while (di >= 0 && cig[di] == 2) di--;
            const int nd = i - di;
            int ii = i + 1;
            while (ii < n && cig[ii] == 1) ii++;
            const int ni = ii - i - 1;
            for (int t = 0; t < ni; t++) cig[di + 1 + t] = 1;
            for (int t = 0; t < nd; t++) cig[di + 1 + ni + t] = 2;
        }
    }
}

__global__ void __launch_bounds__(FIN_THREADS) standardize_kernel(const FinishArgs a)
{
    const int it = blockIdx.x * FIN_THREADS + threadIdx.x;
    if (it >= a.n_items) return;
    const ItemDesc &I = a.items[it];
    uint8_t *c = a.ops + I.out_off;
    const int n = a.item_len[it];
    const uint8_t *ref = a.ref_codes + I.ref_start, *seq = a.seq_codes + I.seq_start;
    for (int k = 0; k < n; k++) { const uint8_t ch = c[k]; c[k] = ch == 'I' ? 1 : ch == 'D' ? 2 : 0; }
    dev_push_indels_left(c, n, ref, 2);
    dev_push_inss_thru_dels(c, n);
    dev_push_indels_left(c, n, seq, 1);
    dev_push_inss_thru_dels(c, n);
    int m = 0;
    for (int k = 0; k < n; k++) {
        if (c[k] == 1 && k + 1 < n && c[k + 1] == 2) { c[m++] = 'M'; k++; }
        else { const uint8_t v = c[k]; c[m++] = v == 0 ? 'M' : v == 1 ? 'I' : 'D'; }
    }
    a.item_len[it] = m;
}

__device__ __forceinline__ uint32_t op_code(uint8_t ch)
{
    return ch == 'M' ? 0u : ch == 'I' ? 1u : ch == 'D' ? 2u : ch == '=' ? 7u : 8u;   // cfg.py:28-32
}

__global__ void __launch_bounds__(FIN_THREADS) rle_kernel(const FinishArgs a)
{
    const int it = blockIdx.x * FIN_THREADS + threadIdx.x;
    if (it >= a.n_items) return;
    const ItemDesc &I = a.items[it];
    const uint8_t *c = a.ops + I.out_off;
    uint32_t *w = a.rle + I.out_off;
    const int n = a.item_len[it];
    int m = 0, k = 0;
    while (k < n) {
        const uint8_t ch = c[k];
        int e = k + 1;
        while (e < n && c[e] == ch) e++;
        int cnt = e - k;
        while (cnt > 0) { const int part = min(cnt, (1 << 28) - 1); w[m++] = ((uint32_t)part << 4) | op_code(ch); cnt -= part; }
        k = e;
    }
    a.rle_len[it] = m;
}
