// forward.cuh -- the align() recurrence (reference: src/aln.pyx:465-667) as an anti-diagonal wavefront.
//
// One warp owns one chunk at a time (persistent warps pull chunk indices from a global counter, largest chunks
// first).  Lane layout is COLUMN-STATIONARY: physical slot s = (column index j) mod NC, NC = 32*CPL >= W = 2r+1,
// lane l holds slots l*CPL .. l*CPL+CPL-1 in registers (CPL = 2 at the default r = 30).  The band of anti-diagonal d
// is the window of columns [jlo, jlo+NC) with jlo = (#D ops so far) - r; b_col = j - jlo.  Consequences:
//   * a cell's row advances by one every anti-diagonal, so the reference's coordinate transforms (aln.pyx:485-492)
//     become FIXED neighbour relations: top (i-1,j) = the slot's own previous value, left (i,j-1) = the previous
//     value of slot s-1, diag (i-1,j-1) = the `left` value fetched one step earlier.  Per anti-diagonal: one
//     register rename + 3 warp shuffles (lane-1, with wrap), independent of the op string -- no divergent paths.
//   * everything that depends on the reference column (np-info of aln.pyx:510-515, pre-decoded by annotate.cuh into
//     SHR/LEN candidate descriptors, the base, a 2-bit 6-mer) stays put in 4 registers per slot for as long as the
//     column is in the band; when a 'D' op turns a column into the b_col==0 EDGE cell its record is dead and the
//     slot loads the record of column j+NC in place (one lane, one 16-byte load, ~2r anti-diagonals before first use).
//   * read-side context (aln.pyx:516-521; 4 bytes per row) moves one slot to the right every step (1 shuffle);
//     an 'I' op inserts the next row at b_col == 0.
//   * LEN / SHR (aln.pyx:596-667 scatter) in gather form, period n descending with strict '<' (= the reference's
//     processing order; larger n wins ties).  The source cell n anti-diagonals back is read from an 8-deep
//     shared-memory ring indexed by physical slot: SHR reads slot s-n (descriptor field), LEN its own slot.  The
//     ring holds MAT.VAL, both run lengths, and the value at each run's start ("BASE"), which replaces the
//     lookback of aln.pyx:623-629, 657-663.
//   * MAT's packed 16-bit record (TYP, RUN, two INDEL 'extended' bits; common.cuh) -- all that traceback reads
//     (aln.pyx:683-685) -- is streamed to HBM, one
//     coalesced 64*CPL-byte row per anti-diagonal, in slot order (traceback indexes it by j mod NC).
//   * two instantiations of the step: GENERIC (chunk head/tail: first row/column values of aln.pyx:525-528,547-550,
//     cells outside the chunk, aln.pyx:497-499) and STEADY (every cell 1 <= b_col <= 2r-1 is an interior cell).
// Arithmetic: fp32 add and strict compare only, tie-break order of aln.pyx:585-592; compiled with --fmad=false.
// The algorithmic form (gather + carried BASE, relaid np-info) is validated on the CPU by oracle/pull_model.c.
#pragma once
#include "common.cuh"
#include <type_traits>

#ifndef FWD_WARPS
#define FWD_WARPS 4
#endif
#ifndef FWD_WARPS_WIDE
#define FWD_WARPS_WIDE 6   // wide bands (CPL >= 4): a ring is 16-32 KB per warp, so fewer, larger CTAs waste less of the
#endif                     // shared memory on the alignment slack (6+1 rings of 32 KB = 224 KB fill one SM at CPL = 8)
static __host__ __device__ constexpr int fwd_warps(int cpl) { return cpl >= 4 ? FWD_WARPS_WIDE : FWD_WARPS; }

#ifndef FWD_ALIGNED
#define FWD_ALIGNED 1   // 1: rings aligned to their size (one ring of slack per CTA), cell address = offset | base
#endif
#if FWD_ALIGNED
#define FWD_ADDR(off, base) ((off) | (base))
#else
#define FWD_ADDR(off, base) ((off) + (base))
#endif

struct ForwardArgs {
    const ChunkDesc *chunks;
    const ChunkSlot *slots;       // indexed like `order`
    const int32_t *order;
    int n;
    int *counter;                 // work queue head
    const ItemDesc *items;
    const uint32_t *bits;
    const uint8_t *ref_codes, *seq_codes;
    const uint4 *colrec;
    const uint2 *relaid;
    const uint32_t *rowrec;
    uint16_t *tb;
    const float *np_tab;          // [np_n][np_dim][np_dim]
    const float *sub_tab;         // [5][5]  indexed [seq_base][ref_base]
    ChunkOut *out;                // indexed by chunk id
    // round-robin time slicing (FWD_RR): run queue + per-chunk saved state
    int *rr_q; int rr_mask; int *rr_ctl;      // ctl[0] head, ctl[1] tail, ctl[2] finished chunks
    uint32_t *rr_state; int rr_slice;
    AlignParams P;
};

#ifndef FWD_RR
#define FWD_RR 1
#endif
#ifndef FWD_SPIN_NS
#define FWD_SPIN_NS 256
#endif
// words of saved per-lane state per cell: Mv1 Iv1 Dv1 dgv Mr1 dgr cc(4) rw
#define FWD_RR_WORDS 11
#define FWD_RR_HDR 32      // uint32 words: d, Id, Dd, hist
static __host__ __device__ inline size_t fwd_rr_state_words(int cpl) { return (size_t)FWD_RR_HDR + (size_t)FWD_RR_WORDS * cpl * 32 + (size_t)32 * cpl * 32; }

__global__ void rr_init_kernel(int *q, int cap, int n, int *ctl)
{
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < cap; i += gridDim.x * blockDim.x) q[i] = i < n ? i : -1;
    if (blockIdx.x == 0 && threadIdx.x == 0) { ctl[0] = 0; ctl[1] = n; ctl[2] = 0; ctl[8] = 0; }
}

// #I among the last n ops (n = 1..6) for every 6-bit op history, packed 4 bits per n at nibble n
__constant__ uint32_t c_sipack[64];

static void fwd_init_constants()
{
    uint32_t h[64];
    for (int x = 0; x < 64; x++) {
        uint32_t v = 0;
        for (int n = 1; n <= 6; n++) v |= (uint32_t)__builtin_popcount(x & ((1 << n) - 1)) << (4 * n);
        h[x] = v;
    }
    cudaMemcpyToSymbol(c_sipack, h, sizeof(h));
}

// run-queue words are polled with L2-only loads: a volatile (system-scope) load from a spinning warp costs every other
// warp of the SM its L1 contents (measured: one waiting warp made a lone chunk 3x slower)
__device__ __forceinline__ int ld_cg_poll(const int *p)
{ int v; asm volatile("ld.global.cg.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory"); return v; }

// ---- explicit shared-memory access by 32-bit byte address (keeps address arithmetic to what is written here)
__device__ __forceinline__ float lds_f(uint32_t a) { float v; asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(a)); return v; }
template <int OFF> __device__ __forceinline__ uint32_t lds_u_off(uint32_t a)
{ uint32_t v; asm volatile("ld.shared.u32 %0, [%1+%2];" : "=r"(v) : "r"(a), "n"(OFF)); return v; }
__device__ __forceinline__ uint2 lds_u2(uint32_t a)
{ uint2 v; asm volatile("ld.shared.v2.u32 {%0,%1}, [%2];" : "=r"(v.x), "=r"(v.y) : "r"(a)); return v; }
template <int OFF> __device__ __forceinline__ void sts_f_off(uint32_t a, float v)
{ asm volatile("st.shared.f32 [%0+%1], %2;" :: "r"(a), "n"(OFF), "f"(v) : "memory"); }
template <int OFF> __device__ __forceinline__ void sts_u_off(uint32_t a, uint32_t v)
{ asm volatile("st.shared.u32 [%0+%1], %2;" :: "r"(a), "n"(OFF), "r"(v) : "memory"); }
template <int OFF> __device__ __forceinline__ void sts_f2_off(uint32_t a, float v0, float v1)
{ asm volatile("st.shared.v2.f32 [%0+%1], {%2,%3};" :: "r"(a), "n"(OFF), "f"(v0), "f"(v1) : "memory"); }
template <int OFF> __device__ __forceinline__ void sts_u2_off(uint32_t a, uint32_t v0, uint32_t v1)
{ asm volatile("st.shared.v2.u32 [%0+%1], {%2,%3};" :: "r"(a), "n"(OFF), "r"(v0), "r"(v1) : "memory"); }

// History ring of one warp: float ring[NP_RING][4][NC]   (arrays: 0 MAT.VAL, 1 SHR run-start value, 2 LEN run-start
// value, 3 LEN.RUN | SHR.RUN<<16), NC*128 bytes; a cell is addressed as (offset & mask) + base.
// One SHR candidate from a pre-decoded descriptor (aln.pyx:642-667 in gather form; annotate.cuh for the fields).
// One SHR candidate (aln.pyx:642-667 in gather form) from a pre-decoded descriptor (annotate.cuh).  Straight-line code:
// a zero descriptor reads valid dummy locations and is rejected by `pred`.  np2 is the score table re-laid with one
// guard column: np2[row][c] = np_scores[row][c-1], np2[row][0] = 100.0 (np_score()'s "ref_l + indel < 0" answer,
// aln.pyx:262-263), row stride T2 = T+1, so the clamp of aln.pyx:269-272 and the guard are one DPX instruction.
template <int NC>
__device__ __forceinline__ void shr_eval(uint32_t D, bool pred, uint32_t dsh, uint32_t wbase, uint32_t lutbase, int bc, uint32_t sip,
                                         const float *__restrict__ np2, int T2, int cl1, float &Sv, int &Sr, float &Sb)
{
    constexpr bool TROW = NC <= 128;
    const uint32_t f = (D >> (TROW ? 18 : 17)) + dsh;
    const uint32_t a = FWD_ADDR(f & (uint32_t)(NC * 128 - 4), wbase);
    const float base = lds_f(a);
    const uint32_t rr = lds_u_off<NC * 8>(a);                         // array + 2: run word when base is array 1
    const uint32_t n = D & 7u, n4 = n << 2;
    const bool start = (D & ((uint32_t)NC << (TROW ? 20 : 19))) == 0u;  // array bit of the descriptor offset field
    const int run0 = start ? 0 : (int)(rr >> 16);
    // (NC <= 128: an empty descriptor points at the all-INF table row and needs no predicate)
    const bool ok = (TROW || pred) && (bc > (int)((sip >> n4) & 7u)) && (start || run0 > 0);
    const int L = (int)((D >> 3) & 0x7fu);
    uint32_t idx;
    if (TROW) {
        const uint32_t magic = lds_u_off<0>(lutbase + n4 * 2u);
        const int q = (int)__umulhi((uint32_t)run0 << 1, magic);
        idx = ((D >> 10) & 0x3ffu) * (uint32_t)T2 + (uint32_t)__vimin_s32_relu(L - q, cl1);
    } else {
        const uint2 lut = lds_u2(lutbase + n4 * 2u);                  // {ceil(2^31/n), (n-1)*T}
        const int q = (int)__umulhi((uint32_t)run0 << 1, lut.x);
        idx = (lut.y + (uint32_t)min(L, cl1 - 1)) * (uint32_t)T2 + (uint32_t)__vimin_s32_relu(L - q, cl1);
    }
    const float cand = base + __ldg(np2 + idx);
    const bool better = ok && cand < Sv;
    Sv = better ? cand : Sv; Sr = better ? __viaddmin_s32(run0, (int)n, NP_RUN_SAT) : Sr; Sb = better ? base : Sb;
}

#ifndef FWD_MINB
#define FWD_MINB 1      // >1: min resident CTAs hint (measured: the hint makes ptxas schedule worse at equal registers)
#endif
#ifndef FWD_MAXREG
#define FWD_MAXREG 0
#endif
template <int CPL>
#if FWD_MAXREG
__global__ void __launch_bounds__(fwd_warps(CPL) * 32) __maxnreg__(CPL <= 2 ? FWD_MAXREG : 255) forward_kernel(const ForwardArgs a)
#elif FWD_MINB > 1
__global__ void __launch_bounds__(fwd_warps(CPL) * 32, (CPL <= 2 ? FWD_MINB : CPL <= 4 ? 12 : 8) / fwd_warps(CPL)) forward_kernel(const ForwardArgs a)
#else
__global__ void __launch_bounds__(fwd_warps(CPL) * 32) forward_kernel(const ForwardArgs a)
#endif
{
    constexpr int NC = 32 * CPL;
    constexpr int TBS = CPL <= 1 ? 1 : CPL <= 2 ? 2 : CPL <= 4 ? 4 : 8;
    constexpr uint32_t RING_BYTES = NC * 128;                    // per warp
    extern __shared__ float smem[];
    __shared__ float s_sub[64];
    __shared__ uint2 s_lut[8];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int t = threadIdx.x; t < 64; t += fwd_warps(CPL) * 32) {
        const int sb = t >> 3, rb = t & 7;
        s_sub[t] = (sb < 5 && rb < 5) ? a.sub_tab[sb * 5 + rb] : 0.f;
    }
    if (threadIdx.x < 8) {
        const uint32_t n = threadIdx.x;
        s_lut[n] = make_uint2(n >= 1 ? (uint32_t)((0x80000000ull + n - 1) / n) : 0u, n >= 1 ? (n - 1) * a.P.np_dim : 0u);
    }
    __syncthreads();
#if FWD_ALIGNED
    const uint32_t smem_base = ((uint32_t)__cvta_generic_to_shared(smem) + RING_BYTES - 1u) & ~(RING_BYTES - 1u);
    if (smem_base - (uint32_t)__cvta_generic_to_shared(smem) > RING_BYTES - 1024u) __trap();      // slack assumed by launch_forward
#else
    const uint32_t smem_base = (uint32_t)__cvta_generic_to_shared(smem);
#endif
    const uint32_t wbase = smem_base + (uint32_t)warp * RING_BYTES;
    const uint32_t subbase = (uint32_t)__cvta_generic_to_shared(s_sub);
    const uint32_t lutbase = (uint32_t)__cvta_generic_to_shared(s_lut);
    const uint32_t myslot4 = (uint32_t)(lane * CPL) * 4u;

    const int r = a.P.r, T2 = a.P.np_dim + 1, cl1 = a.P.np_clamp + 1;
    const float gopen = a.P.gap_open, gext = a.P.gap_ext;
    const uint32_t nmask = ((1u << a.P.max_n) - 1u) << 20;      // rowrec "present" bits live at [20:25]
    const float *__restrict__ np = a.np_tab;          // re-laid table with guard column (api.cu)
    const int src_lane = (lane + 31) & 31;
    constexpr uint32_t empty_f = (uint32_t)((NP_RING - 1) * NC * 16) >> 2;
    const uint32_t empty_desc = NC <= 128 ? (((uint32_t)(a.P.np_rows) << 10) | (empty_f << 20)) : (empty_f << 19);   // annotate.cuh: "no candidate"

    for (;;) {
        int idx = 0;
#if FWD_RR
        {   // pop the next runnable chunk (FIFO); wait for a push if the queue is momentarily empty
            int pos = 0;
            if (lane == 0) {
                pos = atomicAdd(a.rr_ctl, 1);
                int *qp = a.rr_q + (pos & a.rr_mask);
                int v;
                unsigned ns = FWD_SPIN_NS;
                while ((v = ld_cg_poll(qp)) < 0) {
                    if (ld_cg_poll(a.rr_ctl + 2) >= a.n) { v = -2; break; }
#ifdef FWD_SPIN_DEBUG
                    atomicAdd(a.rr_ctl + 8, 1);
#endif
                    __nanosleep(ns);
                    if (ns < 4096u) ns <<= 1;
                }
                if (v >= 0) __stcg(qp, -1);
                idx = v;
            }
            idx = __shfl_sync(NP_FULL, idx, 0);
            if (idx < 0) break;
        }
#else
        if (lane == 0) idx = atomicAdd(a.counter, 1);
        idx = __shfl_sync(NP_FULL, idx, 0);
        if (idx >= a.n) break;
#endif
        const int cid = a.order[idx];
        const ChunkDesc c = a.chunks[cid];
        if (!c.valid) {
            if (lane == 0) {
                a.out[cid].score = 0.f;
#if FWD_RR
                atomicAdd(a.rr_ctl + 2, 1);
#endif
            }
            continue;
        }
        const ChunkSlot sl = a.slots[idx];
        const ItemDesc &I = a.items[c.item];
        const uint32_t *__restrict__ bits = a.bits + I.bit_word_off;
        const uint4 *__restrict__ col = a.colrec + sl.col_off;
        const uint2 *__restrict__ rel = a.relaid + sl.col_off;
        const uint32_t *__restrict__ row = a.rowrec + sl.row_off;
        const uint8_t *__restrict__ refs = a.ref_codes + I.ref_start + min(c.c0, I.ref_len);
        const uint8_t *__restrict__ seqs = a.seq_codes + I.seq_start + min(c.r0, I.seq_len);
        uint16_t *tbp = a.tb + (size_t)sl.tb_off * (32 * TBS) + lane * TBS;
        const int B = c.B, imax = c.imax, jmax = c.jmax;

        // ---- per-slot state.  At d = 0: jlo = -r, slot s holds column j = -r + ((s + r) mod NC), row i = -j.
        float Mv1[CPL], Iv1[CPL], Dv1[CPL], dgv[CPL];   // previous anti-diagonal: MAT/INS/DEL values; diag MAT value
        int Mr1[CPL], dgr[CPL];                           // matrun (RUN if TYP==MAT else 0) of the cell; of the diag cell
        uint4 cc[CPL]; uint32_t rw[CPL]; int bc[CPL];
#pragma unroll
        for (int k = 0; k < CPL; k++) {
            Mv1[k] = Iv1[k] = Dv1[k] = dgv[k] = 0.f; Mr1[k] = dgr[k] = 0;
            const int s = lane * CPL + k;
            bc[k] = (s + r) & (NC - 1);
            const int j0 = bc[k] - r;
            rw[k] = (j0 <= 0) ? row[-j0] : 0u;
            cc[k] = (j0 > 0) ? col[j0] : (j0 == -r ? col[NC - r] : make_uint4(empty_desc, empty_desc, 0u, 0u));
            if (j0 == 0) cc[k] = col[0];
        }
        int d0 = 0; uint32_t hist = 0; int Id = 0, Dd = 0;
#if FWD_RR
        uint32_t *st = a.rr_state + (size_t)idx * fwd_rr_state_words(CPL);
        d0 = (int)__ldcg(st);
        if (d0 > 0) {     // resume: scalars, per-lane registers, history ring
            Id = (int)__ldcg(st + 1); Dd = (int)__ldcg(st + 2); hist = __ldcg(st + 3);
            const uint32_t *sp = st + FWD_RR_HDR + lane;
#pragma unroll
            for (int k = 0; k < CPL; k++) {
                const uint32_t *q = sp + (size_t)k * FWD_RR_WORDS * 32;
                Mv1[k] = __uint_as_float(__ldcg(q)); Iv1[k] = __uint_as_float(__ldcg(q + 32)); Dv1[k] = __uint_as_float(__ldcg(q + 64));
                dgv[k] = __uint_as_float(__ldcg(q + 96)); Mr1[k] = (int)__ldcg(q + 128); dgr[k] = (int)__ldcg(q + 160);
                cc[k] = make_uint4(__ldcg(q + 192), __ldcg(q + 224), __ldcg(q + 256), __ldcg(q + 288));
                rw[k] = __ldcg(q + 320);
                bc[k] = (lane * CPL + k + r - Dd) & (NC - 1);
            }
            const uint4 *rp = reinterpret_cast<const uint4 *>(st + FWD_RR_HDR + (size_t)FWD_RR_WORDS * CPL * 32);
#pragma unroll
            for (int t = 0; t < NC / 4; t++) {
                const uint4 v = __ldcg(rp + t * 32 + lane);
                asm volatile("st.shared.v4.u32 [%0], {%1,%2,%3,%4};" :: "r"(wbase + (uint32_t)(t * 32 + lane) * 16u), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
            }
            __syncwarp();
        }
#endif
        uint32_t nrow = row[Id + r + 1];                     // next row to enter the band (at b_col == 0)
        const int g00 = c.brk + (d0 > 0 ? d0 - 1 : 0);
        int wbase_w = g00 >> 5; uint32_t wbuf = bits[wbase_w + lane];
        uint32_t cw = __shfl_sync(NP_FULL, wbuf, 0);
        float infd = (float)(100 * (d0 > 0 ? d0 - 1 : 0));   // 100*d, exact in fp32 (d < 2^16); advanced at the top of a step
        // steady state = every cell with 1 <= b_col <= 2r-1 is an interior cell with i >= 2 and j >= 2
        const int idLo = r + 1, idSpan = imax - 2 * r, ddSpan = jmax - 2 * r;
#if FWD_RR
        int dEnd = min(B, d0 + a.rr_slice);
#else
        const int dEnd = B;
#endif

        for (int d = d0; ; d++) {
            if (d >= dEnd) {
                if (dEnd >= B) break;
#if FWD_RR
                // slice over: hand the chunk back only if another chunk is waiting for a warp (tail - head > 0)
                int queued = 0;
                if (lane == 0) queued = ld_cg_poll(a.rr_ctl + 1) - ld_cg_poll(a.rr_ctl);
                if (__shfl_sync(NP_FULL, queued, 0) > 0) break;
                dEnd = min(B, dEnd + a.rr_slice);
#endif
            }
            float lMv[CPL], lDv[CPL]; int lMr[CPL];
            if (d > 0) {
                const int g = c.brk + d - 1;                 // op that leads to this anti-diagonal
                if ((g & 31) == 0 && d > 1) {
                    int wi = (g >> 5) - wbase_w;
                    if (wi >= 32) { wbase_w += 32; wbuf = bits[wbase_w + lane]; wi -= 32; }
                    cw = __shfl_sync(NP_FULL, wbuf, wi);
                }
                const uint32_t o = (cw >> (g & 31)) & 1u;
                hist = ((hist << 1) | o) & 0x3fu;
                infd += 100.f;
                const float a0 = __shfl_sync(NP_FULL, Mv1[CPL - 1], src_lane);
                const float a1 = __shfl_sync(NP_FULL, Dv1[CPL - 1], src_lane);
                const int a2 = __shfl_sync(NP_FULL, Mr1[CPL - 1], src_lane);
                const uint32_t a3 = __shfl_sync(NP_FULL, rw[CPL - 1], src_lane);
#pragma unroll
                for (int k = CPL - 1; k >= 0; k--) {
                    lMv[k] = k ? Mv1[k > 0 ? k - 1 : 0] : a0;
                    lDv[k] = k ? Dv1[k > 0 ? k - 1 : 0] : a1;
                    lMr[k] = k ? Mr1[k > 0 ? k - 1 : 0] : a2;
                    rw[k] = k ? rw[k > 0 ? k - 1 : 0] : a3;
                }
                if (o) {
                    Id++;
#pragma unroll
                    for (int k = 0; k < CPL; k++) if (bc[k] == 0) rw[k] = nrow;
                    nrow = row[Id + r + 1];
                } else {
                    Dd++;
#pragma unroll
                    for (int k = 0; k < CPL; k++) {
                        bc[k] = (bc[k] - 1) & (NC - 1);
                        if (bc[k] == 0) cc[k] = col[Dd - r + NC];      // this column is now the EDGE cell: its record is dead
                    }
                }
            } else {
#pragma unroll
                for (int k = 0; k < CPL; k++) { lMv[k] = lDv[k] = 0.f; lMr[k] = 0; }
            }
            const uint32_t sip = c_sipack[hist];
            const float edgev = infd + 100.f;
            const uint32_t dsh = (uint32_t)d * (uint32_t)(NC * 16);
            const bool steady = (unsigned)(Id - idLo) <= (unsigned)idSpan && (unsigned)(Dd - idLo) <= (unsigned)ddSpan && idSpan >= 0 && ddSpan >= 0;
            // The cell body is instantiated twice: STEADY (every cell 1 <= b_col <= 2r-1 is interior with i,j >= 2: constant
            // bounds, no first-row/column code) and generic (chunk head / tail).
            auto cell_body = [&](auto steady_tag) {
            constexpr bool STEADY = decltype(steady_tag)::value;
            int lo = 1, hi = 2 * r - 1;
            if (!STEADY) {      // interior-cell bounds on b_col for this anti-diagonal (aln.pyx:497-507)
                lo = max(1, max(Id + r - imax, r - Dd)); hi = min(2 * r - 1, min(Id + r, jmax + r - Dd));
                if (hi < lo) { lo = 1; hi = 0; }
            }
            const unsigned span = (unsigned)(hi - lo);

            bool in[CPL];
            float Sv[CPL], Sb[CPL], Lv[CPL], Lb[CPL]; int Sr[CPL], Lr[CPL];
            bool p0[CPL], p1[CPL], pg[CPL], pl[CPL];
            uint32_t any1 = 0u, anyg = 0u, anyl = 0u;      // warp votes on plain ORs of the raw words (slightly conservative)
#pragma unroll
            for (int k = 0; k < CPL; k++) {
                in[k] = STEADY ? ((unsigned)(bc[k] - 1) <= (unsigned)(2 * r - 2)) : ((unsigned)(bc[k] - lo) <= span && hi >= lo);
                Sv[k] = infd; Lv[k] = infd; Sb[k] = 0.f; Lb[k] = 0.f; Sr[k] = 0; Lr[k] = 0;
                p0[k] = in[k] && cc[k].x != empty_desc;          // (only consulted by the NC = 256 instantiation)
                p1[k] = in[k] && cc[k].y != empty_desc;
                pg[k] = in[k] && (cc[k].z & 1u);
                const uint32_t lw = rw[k] & cc[k].w & nmask;               // one-hot LEN period vs "tract present" bits
                pl[k] = in[k] && lw != 0u;
                any1 |= cc[k].y ^ empty_desc; anyg |= cc[k].z; anyl |= lw;
            }
            // ---- SHR gather: descriptor 0 (largest period; some lane almost always has one), then descriptor 1
#pragma unroll
            for (int k = 0; k < CPL; k++) shr_eval<NC>(cc[k].x, p0[k], dsh, wbase, lutbase, bc[k], sip, np, T2, cl1, Sv[k], Sr[k], Sb[k]);
            if (__any_sync(NP_FULL, any1 != 0u)) {
#pragma unroll
                for (int k = 0; k < CPL; k++) shr_eval<NC>(cc[k].y, p1[k], dsh, wbase, lutbase, bc[k], sip, np, T2, cl1, Sv[k], Sr[k], Sb[k]);
            }
            // ---- LEN gather (aln.pyx:602-633): single eligible period, 2-bit k-mer unit compare in registers
            if (__any_sync(NP_FULL, anyl != 0u)) {
#pragma unroll
                for (int k = 0; k < CPL; k++) {
                    if (pl[k]) {
                        const uint32_t D = cc[k].w;
                        if ((cc[k].z | (rw[k] >> 3)) & 2u) {
                            pg[k] = true; anyg |= 1u;              // an N inside a k-mer: byte-wise compare on the generic path
                        } else {
                            const int n = (int)(D & 7u);
                            const uint32_t n4 = (uint32_t)n << 2;
                            // match() of aln.pyx:606-607: seq[i-n .. i) against ref[j .. j+n).  The row record carries the 6-mer
                            // that ENDS at seq[i-1] (its last n codes are the read-side unit), the column record the 6-mer that
                            // starts at ref[j]
                            const bool eq = (((((rw[k] >> (12 - 2 * n)) ^ cc[k].z) >> 8) << (32 - 2 * n)) == 0u);
                            const bool start = ((rw[k] >> (25 + n)) & 1u) != 0u;
                            const uint32_t f = (uint32_t)((-n) & (NP_RING - 1)) * (uint32_t)(NC * 16) + dsh + myslot4 + (uint32_t)(k * 4) + (start ? 0u : (uint32_t)(NC * 8));
                            const uint32_t ad = FWD_ADDR(f & (uint32_t)(NC * 128 - 4), wbase);
                            const float base = lds_f(ad);
                            const uint32_t rr = lds_u_off<NC * 4>(ad);                 // array 2 + 1 = run word
                            const uint32_t magic = lds_u_off<0>(lutbase + n4 * 2u);
                            const int run0 = start ? 0 : (int)(rr & 0xffffu);
                            const bool ok = eq && (bc[k] + n - (int)((sip >> n4) & 7u) <= 2 * r - 1) && (start || run0 > 0);
                            const int q = (int)__umulhi((uint32_t)run0 << 1, magic);
                            const int L = (int)((D >> 3) & 0x7fu);
                            const float cand = base + __ldg(np + (((D >> 10) & 0x3ffu) * (uint32_t)T2 + (uint32_t)min(L + q + 2, cl1)));
                            if (ok && cand < Lv[k]) { Lv[k] = cand; Lr[k] = min(run0 + n, NP_RUN_SAT); Lb[k] = base; }
                        }
                    }
                }
            }
            // ---- generic path (rare): all periods from the relaid byte record in global memory
            if (__any_sync(NP_FULL, (anyg & 1u) != 0u)) {
#pragma unroll
                for (int k = 0; k < CPL; k++) {
                    if (pg[k]) {
                        const int i = Id + r - bc[k], j = Dd - r + bc[k];
                        const uint2 rb = rel[j];
                        if (cc[k].z & 1u) {
                            for (int n = a.P.max_n; n >= 1; n--) {      // SHR, every period
                                const uint32_t byte = ((n <= 4 ? rb.x : rb.y) >> ((8 * (n - 1)) & 31)) & 0xffu;
                                const uint32_t L = byte & 0x7fu;
                                if (!L) continue;
                                const uint32_t F = (uint32_t)((-n) & (NP_RING - 1)) * (NC * 16) + ((byte & 0x80u) ? 0u : (uint32_t)(NC * 4)) +
                                                   (uint32_t)((j - n) & (NC - 1)) * 4u;
                                const uint32_t trow = (uint32_t)((n - 1) * (T2 - 1) + min((int)L, cl1 - 1));
                                const uint32_t D = (uint32_t)n | (L << 3) | (NC <= 128 ? (trow << 10) | ((F >> 2) << 20) : ((F >> 2) << 19));
                                shr_eval<NC>(D, true, dsh, wbase, lutbase, bc[k], sip, np, T2, cl1, Sv[k], Sr[k], Sb[k]);
                            }
                        }
                        uint32_t lm = (rb.y >> 22) & ((rw[k] & nmask) >> 20) & 0x3fu;   // LEN, every eligible period
                        while (lm) {
                            const int n = 32 - __clz(lm);
                            lm &= ~(1u << (n - 1));
                            const int sI = (int)((sip >> (4 * n)) & 7u);
                            if (bc[k] + n - sI > 2 * r - 1) continue;
                            const int si = i - n;
                            bool eq = true;
                            for (int t = 0; t < n; t++) eq = eq && (seqs[si + t] == refs[j + t]);
                            if (!eq) continue;
                            const uint2 cjn = rel[j + n];
                            const int L = (int)(((n <= 4 ? cjn.x : cjn.y) >> ((8 * (n - 1)) & 31)) & 0x7fu);
                            const bool start = ((rw[k] >> (25 + n)) & 1u) != 0u;
                            const uint32_t f = (uint32_t)((-n) & (NP_RING - 1)) * (NC * 16) + dsh + myslot4 + (uint32_t)(k * 4) + (start ? 0u : (uint32_t)(NC * 8));
                            const uint32_t ad = FWD_ADDR(f & (uint32_t)(NC * 128 - 4), wbase);
                            const float base = lds_f(ad);
                            const int run0 = start ? 0 : (int)(lds_u_off<NC * 4>(ad) & 0xffffu);
                            if (!start && run0 <= 0) continue;
                            const int call = L + run0 / n + 1;
                            const float cand = base + __ldg(np + ((n - 1) * (T2 - 1) + min(L, cl1 - 1)) * T2 + min(call + 1, cl1));
                            if (cand < Lv[k]) { Lv[k] = cand; Lr[k] = min(run0 + n, NP_RUN_SAT); Lb[k] = base; }
                        }
                    }
                }
            }

            // ---- INS / DEL / MAT (aln.pyx:525-592)
            uint32_t recs[CPL];
            float Mv[CPL], Iv[CPL], Dv[CPL]; int Mr[CPL];
#pragma unroll
            for (int k = 0; k < CPL; k++) {
                // INS from top = own previous value; DEL from left.  Only the "extended" bits are kept (see common.cuh)
                const float iv1 = Mv1[k] + gopen, iv2 = Iv1[k] + gext;
                Iv[k] = fminf(iv1, iv2);                                             // equal on ties: either operand is the value
                uint32_t ieb = __float_as_uint(iv2 - iv1);                           // sign bit set <=> iv2 < iv1 (aln.pyx:536)
                const float dv1 = lMv[k] + gopen, dv2 = lDv[k] + gext;
                Dv[k] = fminf(dv1, dv2);
                uint32_t deb = __float_as_uint(dv2 - dv1);                           // aln.pyx:558
                uint32_t pk = (uint32_t)min(dgr[k] + 1, NP_RUN_SAT);                 // typ MAT = 0
                float best = dgv[k] + lds_f(subbase + (rw[k] & 0xe0u) + (cc[k].z & 0x1cu));
                if (!STEADY) {
                    const int i = Id + r - bc[k], j = Dd - r + bc[k];
                    if (i <= 1) ieb = 0u;                                            // aln.pyx:537-538 (run restarts), :525-528
                    if (j <= 1) deb = 0u;                                            // aln.pyx:559-560, :547-550
                    if (i == 0) Iv[k] = (float)(100 * (j + 1));
                    if (j == 0) Dv[k] = (float)(100 * (i + 1));
                    if (!(i > 0 && j > 0)) { best = Dv[k] + 100.f; pk = 0u; }
                }
                if (Iv[k] < best) { best = Iv[k]; pk = (uint32_t)T_INS << NP_REC_TYP; }
                if (Lv[k] < best) { best = Lv[k]; pk = ((uint32_t)T_LEN << NP_REC_TYP) + (uint32_t)Lr[k]; }
                if (Dv[k] < best) { best = Dv[k]; pk = (uint32_t)T_DEL << NP_REC_TYP; }
                if (Sv[k] < best) { best = Sv[k]; pk = ((uint32_t)T_SHR << NP_REC_TYP) + (uint32_t)Sr[k]; }
                // EDGE (b_col 0 / 2r): every state INF*(b_row+1), TYP MAT, RUN 0 (aln.pyx:502-507).  Cells outside the chunk
                // (aln.pyx:497-499) are never read by interior cells; they get the same harmless value.
                Mv[k] = in[k] ? best : edgev;
                Iv[k] = in[k] ? Iv[k] : edgev;
                Dv[k] = in[k] ? Dv[k] : edgev;
                Mr[k] = (in[k] && pk < (1u << NP_REC_TYP)) ? (int)pk : 0;
                recs[k] = in[k] ? (pk + (ieb >> 31) * NP_REC_IE + (deb >> 31) * NP_REC_DE) : 0u;   // IMADs: FMA pipe
            }

            // ---- history ring [ring][array][slot] + traceback row (slot order)
            {
                const uint32_t ad = FWD_ADDR((dsh & (uint32_t)(NC * 128 - 1)) + myslot4, wbase);
                if (CPL % 2 == 0) {
#pragma unroll
                    for (int k = 0; k < CPL; k += 2) {
                        const uint32_t adk = ad + (uint32_t)(k * 4);
                        const int k1 = k + 1 < CPL ? k + 1 : k;
                        sts_f2_off<0>(adk, Mv[k], Mv[k1]); sts_f2_off<NC * 4>(adk, Sb[k], Sb[k1]); sts_f2_off<NC * 8>(adk, Lb[k], Lb[k1]);
                        sts_u2_off<NC * 12>(adk, (uint32_t)Lr[k] | ((uint32_t)Sr[k] << 16), (uint32_t)Lr[k1] | ((uint32_t)Sr[k1] << 16));
                    }
                } else {
#pragma unroll
                    for (int k = 0; k < CPL; k++) {
                        const uint32_t adk = ad + (uint32_t)(k * 4);
                        sts_f_off<0>(adk, Mv[k]); sts_f_off<NC * 4>(adk, Sb[k]); sts_f_off<NC * 8>(adk, Lb[k]);
                        sts_u_off<NC * 12>(adk, (uint32_t)Lr[k] | ((uint32_t)Sr[k] << 16));
                    }
                }
                uint16_t *rowp = tbp + (size_t)d * (32 * TBS);
                if (CPL == 2) __stcs(reinterpret_cast<unsigned int *>(rowp), recs[0] | (recs[CPL - 1] << 16));    // streamed once, read once by the traceback
                else {
#pragma unroll
                    for (int k = 0; k < CPL; k++) __stcs(reinterpret_cast<unsigned short *>(rowp) + k, (unsigned short)recs[k]);
                }
            }
            __syncwarp();
#pragma unroll
            for (int k = 0; k < CPL; k++) {
                dgv[k] = lMv[k]; dgr[k] = lMr[k];                 // next step's diagonal neighbour = this step's left
                Mv1[k] = Mv[k]; Iv1[k] = Iv[k]; Dv1[k] = Dv[k]; Mr1[k] = Mr[k];
            }
            };   // cell_body
            if (steady) cell_body(std::true_type{}); else cell_body(std::false_type{});
        }
#if FWD_RR
        if (dEnd < B) {     // slice over: save the chunk's state and hand it back to the run queue
            uint32_t *sp = st + FWD_RR_HDR + lane;
#pragma unroll
            for (int k = 0; k < CPL; k++) {
                uint32_t *q = sp + (size_t)k * FWD_RR_WORDS * 32;
                __stcg(q, __float_as_uint(Mv1[k])); __stcg(q + 32, __float_as_uint(Iv1[k])); __stcg(q + 64, __float_as_uint(Dv1[k]));
                __stcg(q + 96, __float_as_uint(dgv[k])); __stcg(q + 128, (uint32_t)Mr1[k]); __stcg(q + 160, (uint32_t)dgr[k]);
                __stcg(q + 192, cc[k].x); __stcg(q + 224, cc[k].y); __stcg(q + 256, cc[k].z); __stcg(q + 288, cc[k].w);
                __stcg(q + 320, rw[k]);
            }
            uint4 *rp = reinterpret_cast<uint4 *>(st + FWD_RR_HDR + (size_t)FWD_RR_WORDS * CPL * 32);
#pragma unroll
            for (int t = 0; t < NC / 4; t++) {
                uint4 v;
                asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(wbase + (uint32_t)(t * 32 + lane) * 16u));
                __stcg(rp + t * 32 + lane, v);
            }
            if (lane == 0) { __stcg(st, (uint32_t)dEnd); __stcg(st + 1, (uint32_t)Id); __stcg(st + 2, (uint32_t)Dd); __stcg(st + 3, hist); }
            __threadfence();
            __syncwarp();
            if (lane == 0) {
                const int p2 = atomicAdd(a.rr_ctl + 1, 1);
                __stcg(a.rr_q + (p2 & a.rr_mask), idx);
            }
            continue;
        }
#endif
        // chunk score = MAT value of the end cell (b_col == r on the last anti-diagonal)
#pragma unroll
        for (int k = 0; k < CPL; k++)
            if (bc[k] == r) a.out[cid].score = Mv1[k];
        __syncwarp();
#if FWD_RR
        if (lane == 0) { __threadfence(); atomicAdd(a.rr_ctl + 2, 1); }
#endif
    }
}
