// finish.cuh -- per-item epilogue on device:
//   gather_kernel        concatenates the chunks' op strings (aln.pyx:742 `full_aln += aln[::-1]`)
//   rle_kernel           src/cig.pyx:13-38 collapse_cigar as BAM-style words (len<<4 | op); with `to_m` set X and = are
//                        first mapped to M (bam.pyx:65)
//   standardize_kernel   src/bam.pyx:65-78 in the RUN-LENGTH domain: push_indels_left(D, ref); push_inss_thru_dels;
//                        push_indels_left(I, seq); push_inss_thru_dels (src/cig.pyx:102-192; the reference's `while True`
//                        body runs exactly once because old_cig aliases int_cig); 'ID' -> 'M'.  Each pass is one
//                        sequential sweep over the item's groups with a stack-like output (O(#groups), ~1000 per 10 kb
//                        read, instead of the reference's 4 sweeps over every op), one warp (lane 0) per item.
//   expand_kernel        run-length words -> one char per op (the expanded 'MID' string realign_hap returns)
//   scan / pack kernels  exclusive prefix of the per-item output sizes and a dense copy, so that the D2H transfer moves
//                        exactly the bytes the caller gets.
#pragma once
#include "common.cuh"

#define FIN_THREADS 128

struct FinishArgs {
    const ItemDesc *items; int n_items;
    const ChunkOut *chunk_out;
    const uint8_t *scratch;       // right-aligned chunk pieces
    uint8_t *ops;                 // per-item op strings (item region at out_off)
    int32_t *item_len;            // ops per item
    int32_t *item_status;
    const uint8_t *ref_codes, *seq_codes;
    uint32_t *rleA, *rleB;        // per-item group buffers (item region at out_off), ping-pong
    int32_t *rle_len;
    int32_t *rle_which;           // 0: result in rleA, 1: in rleB
    int to_m;
    // packing
    int64_t *ops_off, *rle_off;   // [n+1] exclusive prefixes (device)
    uint8_t *pack_ops; uint32_t *pack_rle;
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

__device__ __forceinline__ uint32_t op_code(uint8_t ch, int to_m)
{
    if (ch == 'I') return 1u;
    if (ch == 'D') return 2u;
    if (to_m || ch == 'M') return 0u;
    return ch == '=' ? 7u : 8u;                                   // cfg.py:28-32
}

// expanded chars -> run-length words, one CTA per item: boundary flags, block scan, starts, lengths
__global__ void __launch_bounds__(FIN_THREADS) rle_kernel(const FinishArgs a)
{
    __shared__ int s_warp[FIN_THREADS / 32];
    __shared__ int s_carry;
    const int it = blockIdx.x;
    if (it >= a.n_items) return;
    const ItemDesc &I = a.items[it];
    const uint8_t *c = a.ops + I.out_off;
    uint32_t *w = a.rleA + I.out_off, *startp = a.rleB + I.out_off;    // rleB: start position of group g (scratch)
    const int n = a.item_len[it];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    if (threadIdx.x == 0) s_carry = 0;
    __syncthreads();
    for (int base = 0; base < n; base += FIN_THREADS) {
        const int k = base + threadIdx.x;
        int flag = 0;
        if (k < n) flag = (k == 0) || (op_code(c[k], a.to_m) != op_code(c[k - 1], a.to_m));
        int x = flag;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const int y = __shfl_up_sync(NP_FULL, x, o); if (lane >= o) x += y; }
        if (lane == 31) s_warp[wid] = x;
        __syncthreads();
        int pre = s_carry;
        for (int q = 0; q < wid; q++) pre += s_warp[q];
        if (flag) startp[pre + x - 1] = (uint32_t)k;
        __syncthreads();
        if (threadIdx.x == FIN_THREADS - 1) s_carry = pre + x;
        __syncthreads();
    }
    const int m = s_carry;
    for (int g = threadIdx.x; g < m; g += FIN_THREADS) {
        const uint32_t st = startp[g], en = (g + 1 < m) ? startp[g + 1] : (uint32_t)n;
        w[g] = ((en - st) << 4) | op_code(c[st], a.to_m);
    }
    if (threadIdx.x == 0) { a.rle_len[it] = m; a.rle_which[it] = 0; }
}

// ---- stack-like output list of run-length groups (merges equal neighbours, drops empty groups)
struct GroupStack {
    uint32_t *g; int m;
    __device__ __forceinline__ void push(uint32_t op, uint32_t len)
    {
        if (!len) return;
        if (m > 0 && (g[m - 1] & 15u) == op) g[m - 1] += len << 4;
        else g[m++] = (len << 4) | op;
    }
};

// cig.pyx:102-159 on groups.  sp counts M ops and push_op runs passed (= position in `seq`, which is the reference for
// push_op == D and the read for push_op == I).  A run of push_op of length L preceded by an M group moves left over k of
// those M's, k = number of consecutive t with seq[sp-t-1] == seq[sp-t-1+L].
__device__ __forceinline__ int rle_push_indels_left(const uint32_t *in, int m, uint32_t *out, const uint8_t *__restrict__ seq, uint32_t push_op)
{
    GroupStack st{out, 0};
    int sp = 0;
    for (int g = 0; g < m; g++) {
        const uint32_t op = in[g] & 15u, len = in[g] >> 4;
        if (op != push_op) { st.push(op, len); if (op == 0u) sp += (int)len; continue; }
        int k = 0;
        if (st.m > 0 && (st.g[st.m - 1] & 15u) == 0u) {
            const int lim = min((int)(st.g[st.m - 1] >> 4), sp);
            while (k < lim && seq[sp - k - 1] == seq[sp - k - 1 + (int)len]) k++;
            if (k) { st.g[st.m - 1] -= (uint32_t)k << 4; if ((st.g[st.m - 1] >> 4) == 0u) st.m--; }
        }
        st.push(push_op, len);
        st.push(0u, (uint32_t)k);
        sp += (int)len;
    }
    return st.m;
}

// cig.pyx:164-192 on groups: an I run that follows a D run is moved in front of it
__device__ __forceinline__ int rle_push_inss_thru_dels(const uint32_t *in, int m, uint32_t *out)
{
    GroupStack st{out, 0};
    for (int g = 0; g < m; g++) {
        const uint32_t op = in[g] & 15u, len = in[g] >> 4;
        if (op == 1u && st.m > 0 && (st.g[st.m - 1] & 15u) == 2u) {
            const uint32_t d = st.g[st.m - 1] >> 4;
            st.m--;
            st.push(1u, len);
            st.push(2u, d);
        } else st.push(op, len);
    }
    return st.m;
}

// bam.pyx:78  .replace('ID','M') on groups: the last I of a run and the first D of the run that follows become one M
__device__ __forceinline__ int rle_id_to_m(const uint32_t *in, int m, uint32_t *out)
{
    GroupStack st{out, 0};
    for (int g = 0; g < m; g++) {
        const uint32_t op = in[g] & 15u, len = in[g] >> 4;
        if (op == 2u && st.m > 0 && (st.g[st.m - 1] & 15u) == 1u) {
            st.g[st.m - 1] -= 1u << 4;
            if ((st.g[st.m - 1] >> 4) == 0u) st.m--;
            st.push(0u, 1u);
            st.push(2u, len - 1u);
        } else st.push(op, len);
    }
    return st.m;
}

// one item per WARP, lane 0 does the sweeps: 32 different sequential sweeps inside one warp would serialise
__global__ void __launch_bounds__(FIN_THREADS) standardize_kernel(const FinishArgs a)
{
    const int it = blockIdx.x * (FIN_THREADS / 32) + (threadIdx.x >> 5);
    if (it >= a.n_items || (threadIdx.x & 31) != 0) return;
    const ItemDesc &I = a.items[it];
    uint32_t *A = a.rleA + I.out_off, *B = a.rleB + I.out_off;
    const uint8_t *ref = a.ref_codes + I.ref_start, *seq = a.seq_codes + I.seq_start;
    int m = a.rle_len[it];
    m = rle_push_indels_left(A, m, B, ref, 2u);
    m = rle_push_inss_thru_dels(B, m, A);
    m = rle_push_indels_left(A, m, B, seq, 1u);
    m = rle_push_inss_thru_dels(B, m, A);
    m = rle_id_to_m(A, m, B);
    int tot = 0;
    for (int g = 0; g < m; g++) tot += (int)(B[g] >> 4);
    a.rle_len[it] = m; a.rle_which[it] = 1; a.item_len[it] = tot;
}

// run-length words -> chars, one CTA per item
__global__ void __launch_bounds__(FIN_THREADS) expand_kernel(const FinishArgs a)
{
    __shared__ int s_warp[FIN_THREADS / 32];
    __shared__ int s_carry;
    const int it = blockIdx.x;
    if (it >= a.n_items) return;
    const ItemDesc &I = a.items[it];
    const uint32_t *w = (a.rle_which[it] ? a.rleB : a.rleA) + I.out_off;
    uint8_t *c = a.ops + I.out_off;
    const int m = a.rle_len[it];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    if (threadIdx.x == 0) s_carry = 0;
    __syncthreads();
    for (int base = 0; base < m; base += FIN_THREADS) {
        const int g = base + threadIdx.x;
        const uint32_t word = g < m ? w[g] : 0u;
        const int len = (int)(word >> 4);
        int x = len;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const int y = __shfl_up_sync(NP_FULL, x, o); if (lane >= o) x += y; }
        if (lane == 31) s_warp[wid] = x;
        __syncthreads();
        int pre = s_carry;
        for (int q = 0; q < wid; q++) pre += s_warp[q];
        const int start = pre + x - len;
        const uint8_t ch = "MIDNSHP=XB"[word & 15u];
        for (int t = 0; t < len; t++) c[start + t] = ch;
        __syncthreads();
        if (threadIdx.x == FIN_THREADS - 1) s_carry = pre + x;
        __syncthreads();
    }
}

// exclusive prefix sums of item_len / rle_len over items (single CTA; n is small next to the DP)
__global__ void __launch_bounds__(1024) scan_kernel(const FinishArgs a)
{
    __shared__ long long s_w[32][2];
    __shared__ long long s_carry[2];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    if (threadIdx.x == 0) { s_carry[0] = s_carry[1] = 0; }
    __syncthreads();
    for (int base = 0; base < a.n_items; base += 1024) {
        const int it = base + threadIdx.x;
        long long v0 = it < a.n_items ? a.item_len[it] : 0, v1 = it < a.n_items ? a.rle_len[it] : 0;
        long long x0 = v0, x1 = v1;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const long long y0 = __shfl_up_sync(NP_FULL, x0, o), y1 = __shfl_up_sync(NP_FULL, x1, o);
            if (lane >= o) { x0 += y0; x1 += y1; }
        }
        if (lane == 31) { s_w[wid][0] = x0; s_w[wid][1] = x1; }
        __syncthreads();
        long long p0 = s_carry[0], p1 = s_carry[1];
        for (int q = 0; q < wid; q++) { p0 += s_w[q][0]; p1 += s_w[q][1]; }
        if (it < a.n_items) { a.ops_off[it] = p0 + x0 - v0; a.rle_off[it] = p1 + x1 - v1; }
        __syncthreads();
        if (threadIdx.x == 1023) { s_carry[0] = p0 + x0; s_carry[1] = p1 + x1; }
        __syncthreads();
    }
    if (threadIdx.x == 0) { a.ops_off[a.n_items] = s_carry[0]; a.rle_off[a.n_items] = s_carry[1]; }
}

__global__ void __launch_bounds__(FIN_THREADS) pack_kernel(const FinishArgs a, int want_ops, int want_rle)
{
    const int it = blockIdx.x;
    if (it >= a.n_items) return;
    const ItemDesc &I = a.items[it];
    if (want_ops) {
        const uint8_t *src = a.ops + I.out_off; uint8_t *dst = a.pack_ops + a.ops_off[it];
        const int n = a.item_len[it];
        for (int t = threadIdx.x; t < n; t += FIN_THREADS) dst[t] = src[t];
    }
    if (want_rle) {
        const uint32_t *src = (a.rle_which[it] ? a.rleB : a.rleA) + I.out_off; uint32_t *dst = a.pack_rle + a.rle_off[it];
        const int m = a.rle_len[it];
        for (int t = threadIdx.x; t < m; t += FIN_THREADS) dst[t] = src[t];
    }
}
