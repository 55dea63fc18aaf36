// annotate.cuh -- get_np_info on device (reference: src/aln.pyx:179-251), per chunk slice (aln.pyx:453-456),
// emitting the packed per-position records the forward kernel consumes.  Dataflow validated on the CPU by
// oracle/pull_model.c:pm_np_raw / relay_col / relay_row (same statements).
//
// The reference scans positions sequentially with "don't overwrite" rules (aln.pyx:238-249).  Equivalent
// parallel form, for n = 1..max_n in order (the `longest` test of period n reads final L of periods < n):
//   e[q]   = (q+n < len && s[q]==s[q+n])
//   runs   = maximal stretches of e==true plus their terminating false position
//   chains = positions of one run congruent mod n.  One walker per chain head visits its chain in ascending
//            order carrying (first active writer, last active writer with l > max_l) -- that is all the
//            sequential overwrite rules can depend on (incl. the max_l clamp quirk).
// raw[p][n-1] = L | 0x80 if L_IDX == 0.
//
// Records (one per slice position), pre-decoded for the forward kernel (forward.cuh):
//   relaid[j] (uint2): .x = raw[j-n][n] for n=1..4 (one byte each); .y = raw[j-5][5] | raw[j-6][6]<<8 |
//                      LENmask<<22 (bit n-1: raw[j][n] has L!=0 && L_IDX==0).  "Relaid" = the record of column j
//                      holds what the SHR gather of a cell in column j needs from columns j-1..j-6.  Only the
//                      rare generic path of the forward kernel reads it.
//   colrec[2j], colrec[2j+1] (2 x uint4): {S0.A, S0.B, S0.C, S1.A}, {S1.B, S1.C, Z, LEN}
//     S0 / S1 = the first two SHR candidate descriptors of column j, period n descending:
//        A [31:19] (byte offset of the source pair in the forward kernel's history ring) >> 3: ring row (-n mod 8), position of
//                  slot (j-n) mod NC in the row, pair {MAT.VAL, -} if the source column starts the tract (L_IDX==0) else
//                  {carried SHR run-start value, runs}      [18:16] n      [15:0] score-table row (n-1)*(max_l+1) + L
//        B ceil(65536 / n)         C 0 if the source column starts the tract, else 0xffff0000
//        "no candidate": A = the index of the +INF table row, B = C = 0
//     Z   = [0] generic path (more than two SHR candidates or more than one LEN-eligible period; then S0/S1/LEN are empty)
//           [1] k-mer contains N  [2:4] base ref[j-1]  [8:19] 2-bit codes of the LEN unit ref[j..j+n) in the TOP 2n bits of the field
//           (aligned with the last n codes of rowrec's 6-mer; zero below)  [20:31] the mask of those 2n bits (field position + 12)
//     LEN = descriptor of the single LEN-eligible period at j: [0:2] n  [3:12] table row  [20:25] one-hot period mask
//           aligned with rowrec's "tract present" bits (bit 19+n)
//   rowrec[i] (uint32): [1] k-mer contains N  [5:7] base seq[i-1]  [8:19] 2-bit k-mer seq[i-6..i-1] (the LEN unit
//                       seq[i-n..i) is its last n codes)
//                       [20:25] tract present at i-n (bit 19+n)  [26:31] tract start at i-n (bit 25+n, L_IDX==0)
#pragma once
#include "common.cuh"

#define ANN_THREADS 256

struct AnnotateArgs {
    const ChunkDesc *chunks;
    const ChunkSlot *slots;        // indexed like `order`
    const int32_t *order;          // chunk ids of this sub-batch
    int n;                         // chunks in this sub-batch
    const ItemDesc *items;
    const uint8_t *ref_codes, *seq_codes;
    uint8_t *raw_ref, *raw_seq;    // 8 B per entry
    uint4 *colrec; uint2 *relaid; uint32_t *rowrec;
    int max_n, max_l, nc, inf_row;      // inf_row = max_n * (max_l + 1)
    int e6_stride;                      // words per period in the dynamic shared window (0: none)
    int cpl;                            // cells per lane of the forward instantiation that will read the records (ring position of a slot)
};

// np_info of one slice into raw (and optionally the reference's int32 [len][2][max_n] array).
// Per period n: (1) the equality bits e[q] = (q+n < len && s[q]==s[q+n]) of every 32-position window go to shared
// memory; (2) every position finds, with bit operations on those words, how many equalities immediately precede it
// (fewer than n => it heads a phase chain) and where its run ends; (3) chain heads write their chain.  Chains of one
// period write disjoint bytes, so two barriers per period suffice (the `longest` test reads bytes of smaller periods).
#define ANN_MAX_WORDS 2048          // 65536 positions (max_b_rows <= 65000)
// e6 (optional): dynamic shared memory for the equality words of ALL periods, NP_MAXN planes of `e6_stride` words (>= nwords + 2).
// With it the slice is read once (7 bytes per position instead of 2 per position and period) and every period costs one barrier
// instead of two; without it (long stand-alone sequences, callers without the dynamic window) the words are recomputed per period.
__device__ void annotate_slice(const uint8_t *__restrict__ s, int len, int max_n, int max_l, uint8_t *raw, int32_t *full_out,
                               uint32_t *ebits_long = nullptr, uint32_t *e6 = nullptr, int e6_stride = 0)
{
    __shared__ uint32_t s_e_smem[ANN_MAX_WORDS + 2];
    const int tid = threadIdx.x, lane = tid & 31;
    for (int q = tid; q < len; q += ANN_THREADS) reinterpret_cast<uint2 *>(raw)[q] = make_uint2(0u, 0u);
    if (full_out) for (int q = tid; q < len * 2 * max_n; q += ANN_THREADS) full_out[q] = 0;
    const int nwords = (len + 31) >> 5;
    const bool all_periods = e6 != nullptr && !ebits_long && e6_stride >= nwords + 2;
    if (all_periods) {
        for (int base = (tid >> 5) << 5; base < nwords * 32; base += ANN_THREADS) {
            const int q = base + lane;
            uint32_t c[NP_MAXN + 1];
#pragma unroll
            for (int t = 0; t <= NP_MAXN; t++) c[t] = (q + t < len) ? (uint32_t)s[q + t] : 0xffu + (uint32_t)t;      // past the end: equal to nothing
#pragma unroll
            for (int n = 1; n <= NP_MAXN; n++) {
                const uint32_t bits = __ballot_sync(NP_FULL, c[0] == c[n]);
                if (lane == 0) e6[(n - 1) * e6_stride + (base >> 5)] = bits;
            }
        }
        if (tid < NP_MAXN) { e6[tid * e6_stride + nwords] = 0u; e6[tid * e6_stride + nwords + 1] = 0u; }
    }
    __syncthreads();
    for (int n = 1; n <= max_n; n++) {
        uint32_t *s_e = ebits_long ? ebits_long : all_periods ? e6 + (n - 1) * e6_stride : s_e_smem;      // long stand-alone sequences: global memory
        if (!all_periods) {
            for (int base = (tid >> 5) << 5; base < nwords * 32; base += ANN_THREADS) {
                const int q = base + lane;
                const bool e = (q + n < len) && (s[q] == s[q + n]);
                const uint32_t bits = __ballot_sync(NP_FULL, e);
                if (lane == 0) s_e[base >> 5] = bits;
            }
            if (tid == 0) s_e[nwords] = 0u;
            __syncthreads();
        }
        for (int h = tid; h < len; h += ANN_THREADS) {
            const int w = h >> 5, b = h & 31;
            const uint32_t cur = s_e[w];
            {   // a chain head needs 2n equalities from h on (3 copies).  Positions of this word where such a stretch starts,
                // from the 64-bit window cur | next (warp-uniform: most words have none and the warp moves on)
                const uint64_t x = ((uint64_t)s_e[w + 1] << 32) | cur;
                uint64_t y = x & (x >> 1);                                   // 2 ones
                if (n >= 2) {
                    const uint64_t y2 = y & (y >> 2);                        // 4 ones
                    if (n == 2) y = y2;
                    else if (n == 3) y = y2 & (y >> 4);                      // 6
                    else {
                        const uint64_t y4 = y2 & (y2 >> 4);                  // 8
                        y = n == 4 ? y4 : n == 5 ? (y4 & (y >> 8)) : (y4 & (y2 >> 8));      // 10 / 12
                    }
                }
                if (!(((uint32_t)y >> b) & 1u)) continue;
            }
            // equalities immediately before h (only the first n matter)
            const uint64_t below = (((uint64_t)cur << 32) | (uint64_t)(w ? s_e[w - 1] : 0u)) << (32 - b);     // bit 63 = e[h-1]
            const int t = __clzll(~below | 1ull);
            if (t >= n) continue;                             // not among the first n positions of its run
            // first position >= h with e[] false
            int end;
            {
                uint32_t inv = ~cur >> b;
                if (b) inv &= (1u << (32 - b)) - 1u;
                if (inv) end = h + __ffs(inv) - 1;
                else {
                    int ww = w + 1;
                    while (s_e[ww] == 0xffffffffu) ww++;      // s_e[nwords] == 0 terminates
                    end = (ww << 5) + __ffs(~s_e[ww]) - 1;
                }
            }
            if (end - h < 2 * n) continue;                    // fewer than 3 copies from the chain head on: nothing to write
            // head of a chain with l >= 3: is the head itself an active writer (aln.pyx:237-243)?
            const int l0 = (end - h) / n + 1;
            bool act0 = s[h] != 0;
            if (act0) {
                const uint2 rw = reinterpret_cast<const uint2 *>(raw)[h];
                for (int n2 = 1; n2 < n; n2++) {
                    const uint32_t bb = ((n2 <= 4 ? rw.x >> (8 * (n2 - 1)) : rw.y >> (8 * (n2 - 5))) & 0x7fu);
                    if (l0 * n <= (int)bb * n2) act0 = false;
                }
            }
            if (act0 && l0 <= max_l) {
                // common case: the head is the first writer and nothing is clamped -> L = l0, L_IDX = k for member k
                for (int k = 0, p = h; k < l0; k++, p += n) {
                    raw[(size_t)p * 8 + n - 1] = (uint8_t)(l0 | (k == 0 ? 0x80 : 0));
                    if (full_out) {
                        full_out[((size_t)p * 2 + 0) * max_n + n - 1] = l0;
                        full_out[((size_t)p * 2 + 1) * max_n + n - 1] = k;
                    }
                }
                continue;
            }
            // general walk: first active writer, last active writer with l > max_l (the clamp quirk)
            int first = -1, lfirst = 0, zlast = -1;
            for (int p = h; p <= end; p += n) {
                const int m = end - p;
                int l = m / n; if (l > 0) l++;
                if (l <= 2 && first < 0) break;               // l only decreases along the chain
                bool act = s[p] != 0 && l > 2;
                if (act) {
                    const uint2 rw = reinterpret_cast<const uint2 *>(raw)[p];
                    for (int n2 = 1; n2 < n; n2++) {
                        const uint32_t bb = ((n2 <= 4 ? rw.x >> (8 * (n2 - 1)) : rw.y >> (8 * (n2 - 5))) & 0x7fu);
                        if (l * n <= (int)bb * n2) act = false;
                    }
                }
                if (act) {
                    if (first < 0) { first = p; lfirst = l; }
                    if (l > max_l) zlast = p;
                }
                if (first >= 0) {
                    int L, X;
                    if (zlast >= 0) { L = max_l; X = (p - zlast) / n; }
                    else { L = min(lfirst, max_l); X = (p - first) / n; }
                    raw[(size_t)p * 8 + n - 1] = (uint8_t)(L | (X == 0 ? 0x80 : 0));
                    if (full_out) {
                        full_out[((size_t)p * 2 + 0) * max_n + n - 1] = L;
                        full_out[((size_t)p * 2 + 1) * max_n + n - 1] = X;
                    }
                }
            }
        }
        __syncthreads();
    }
}

__device__ __forceinline__ uint32_t raw_byte(const uint8_t *raw, int len, int p, int n)
{
    return (p >= 0 && p < len) ? (uint32_t)raw[(size_t)p * 8 + n - 1] : 0u;
}
// the 8 bytes (periods 1..6) of one position in one load; byte n-1 of the pair = raw[p][n]
__device__ __forceinline__ uint2 raw_row(const uint8_t *raw, int len, int p)
{
    return (p >= 0 && p < len) ? reinterpret_cast<const uint2 *>(raw)[p] : make_uint2(0u, 0u);
}
__device__ __forceinline__ uint32_t row_byte(const uint2 v, int n) { return ((n <= 4 ? v.x >> (8 * (n - 1)) : v.y >> (8 * (n - 5))) & 0xffu); }

__device__ __forceinline__ uint32_t kmer2_of(const uint8_t *s, int len, int p, uint32_t &hasN)
{
    uint32_t km = 0;
#pragma unroll
    for (int t = 0; t < 6; t++) {
        const int q = p + t;
        uint32_t b = 0;
        if (q >= 0 && q < len) { b = s[q]; if (b == 0 || b > 4) hasN = 1; }
        km |= ((b - 1u) & 3u) << (2 * t);
    }
    return km;
}

__global__ void __launch_bounds__(ANN_THREADS) annotate_kernel(AnnotateArgs a)
{
    const int ci = blockIdx.x >> 1, side = blockIdx.x & 1;
    if (ci >= a.n) return;
    const ChunkDesc c = a.chunks[a.order[ci]];
    const ChunkSlot sl = a.slots[ci];
    if (!c.valid) return;
    const ItemDesc &I = a.items[c.item];
    extern __shared__ uint32_t ann_e6[];                 // NP_MAXN planes of a.e6_stride equality words
    if (side == 0) {
        const int len = c.rlen;
        const uint8_t *s = a.ref_codes + I.ref_start + min(c.c0, I.ref_len);
        uint8_t *raw = a.raw_ref + sl.col_off * 8;
        annotate_slice(s, len, a.max_n, a.max_l, raw, nullptr, nullptr, a.e6_stride ? ann_e6 : nullptr, a.e6_stride);
        uint4 *out = a.colrec + 2 * sl.col_off;
        uint2 *rel = a.relaid + sl.col_off;
        const int NC = a.nc, CPL = a.cpl, PER = NC / CPL;      // slot s sits at ring position (s % CPL) * PER + s / CPL (forward.cuh)
        for (int j = threadIdx.x; j < sl.col_cap; j += ANN_THREADS) {
            uint2 v = make_uint2(0u, 0u);
            // "no candidate": the all-INF table row; the source pair it reads is irrelevant, but it must be one nobody writes during the
            // step -- position 0 of the PREVIOUS anti-diagonal's ring row (a read of the current row would race with its owner's store)
            const uint32_t empty = (((uint32_t)(NP_RING - 1) * (uint32_t)(NC * 16)) << 16) | (uint32_t)a.inf_row;
            uint32_t sA[2] = {empty, empty}, sB[2] = {0u, 0u}, sC[2] = {0u, 0u}, z = 0u, lenw = 0u;
            if (j < len + 8) {
                uint32_t lenm = 0, nshr = 0, nlen = 0;
                uint2 rows[NP_MAXN + 1];
#pragma unroll
                for (int t = 0; t <= NP_MAXN; t++) rows[t] = raw_row(raw, len, j - t);
#pragma unroll
                for (int n = NP_MAXN; n >= 1; n--) {
                    const uint32_t b = row_byte(rows[n], n);
                    if (n <= 4) v.x |= b << (8 * (n - 1)); else v.y |= b << (8 * (n - 5));
                    const uint32_t L = b & 0x7fu;
                    if (L) {
                        // byte offset inside one ring [NP_RING rows][NC positions][16 B]; pair 0 = {MAT.VAL, -}, pair 1 = {SHR run-start value, runs}
                        const uint32_t ss = (uint32_t)(j - n) & (uint32_t)(NC - 1);
                        const uint32_t F = (uint32_t)((-n) & (NP_RING - 1)) * (uint32_t)(NC * 16) + ((ss % CPL) * (uint32_t)PER + ss / CPL) * 16u + ((b & 0x80u) ? 0u : 8u);
                        if (nshr < 2) {
                            sA[nshr] = ((F | (uint32_t)n) << 16) | (uint32_t)((n - 1) * (a.max_l + 1) + (int)L);
                            sB[nshr] = (65536u + (uint32_t)n - 1u) / (uint32_t)n;
                            sC[nshr] = (b & 0x80u) ? 0u : 0xffff0000u;
                        }
                        nshr++;
                    }
                    const uint32_t o = row_byte(rows[0], n);
                    if ((o & 0x7fu) && (o & 0x80u)) {
                        lenm |= 1u << (n - 1);
                        const uint32_t Lo = o & 0x7fu;
                        lenw = (uint32_t)n | ((uint32_t)((n - 1) * (a.max_l + 1) + (int)Lo) << 3) | (1u << (19 + n));
                        nlen++;
                    }
                }
                v.y |= lenm << 22;
                const uint32_t base = (j >= 1 && j - 1 < len) ? s[j - 1] : 0u;
                uint32_t hasN = 0;
                const uint32_t km = kmer2_of(s, len, j, hasN);
                const bool more = nshr > 2 || nlen > 1;
                if (more) { sA[0] = sA[1] = empty; sB[0] = sB[1] = sC[0] = sC[1] = 0u; lenw = 0u; }
                uint32_t kf = km, kmask = 0u;
                if (lenw) {      // the LEN unit ref[j..j+n) moved to the top 2n bits of the k-mer field, where the row's unit seq[i-n..i) sits
                    const int n = (int)(lenw & 7u);
                    kmask = ((1u << (2 * n)) - 1u) << (12 - 2 * n);
                    kf = (km << (12 - 2 * n)) & kmask;
                }
                z = (more ? 1u : 0u) | (hasN << 1) | ((base & 7u) << 2) | (kf << 8) | (kmask << 20);
            }
            out[2 * j] = make_uint4(sA[0], sB[0], sC[0], sA[1]);
            out[2 * j + 1] = make_uint4(sB[1], sC[1], z, lenw);
            rel[j] = v;
        }
    } else {
        const int len = c.slen;
        const uint8_t *s = a.seq_codes + I.seq_start + min(c.r0, I.seq_len);
        uint8_t *raw = a.raw_seq + sl.row_off * 8;
        annotate_slice(s, len, a.max_n, a.max_l, raw, nullptr, nullptr, a.e6_stride ? ann_e6 : nullptr, a.e6_stride);
        uint32_t *out = a.rowrec + sl.row_off;
        for (int i = threadIdx.x; i < sl.row_cap; i += ANN_THREADS) {
            uint32_t v = 0;
            if (i < len + 8) {
#pragma unroll
                for (int n = 1; n <= NP_MAXN; n++) {
                    const uint32_t b = row_byte(raw_row(raw, len, i - n), n);
                    if (b & 0x7fu) v |= 1u << (19 + n);
                    if (b & 0x80u) v |= 1u << (25 + n);
                }
                const uint32_t base = (i >= 1 && i - 1 < len) ? s[i - 1] : 0u;
                uint32_t hasN = 0;
                const uint32_t km = kmer2_of(s, len, i - 6, hasN);      // the 6-mer ending at seq[i-1]
                v |= hasN << 1 | (base & 7u) << 5 | km << 8;
            }
            out[i] = v;
        }
    }
}

// src/aln.pyx:179-251 as a stand-alone device entry (npore_get_np_info): one CTA, one sequence.
__global__ void __launch_bounds__(ANN_THREADS)
np_info_kernel(const uint8_t *codes, const int64_t *off, int max_n, int max_l, uint8_t *raw, int32_t *out, uint32_t *ebits)
{
    // one CTA per sequence of the batch; sequences longer than the shared-memory window use global equality words
    const int64_t b = off[blockIdx.x];
    const int len = (int)(off[blockIdx.x + 1] - b);
    annotate_slice(codes + b, len, max_n, max_l, raw + b * 8, out + b * 2 * max_n,
                   len > ANN_MAX_WORDS * 32 ? ebits + (b >> 5) + 2 * blockIdx.x : nullptr);
}
