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
//   * MAT's packed (TYP:3, RUN:13) record -- all that traceback reads (aln.pyx:683-685) -- is streamed to HBM, one
//     coalesced 64*CPL-byte row per anti-diagonal, in slot order (traceback indexes it by j mod NC).
//   * two instantiations of the step: GENERIC (chunk head/tail: first row/column values of aln.pyx:525-528,547-550,
//     cells outside the chunk, aln.pyx:497-499) and STEADY (every cell 1 <= b_col <= 2r-1 is an interior cell).
// Arithmetic: fp32 add and strict compare only, tie-break order of aln.pyx:585-592; compiled with --fmad=false.
// The algorithmic form (gather + carried BASE, relaid np-info) is validated on the CPU by oracle/pull_model.c.
#pragma once
#include "common.cuh"

#define FWD_WARPS 4

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
    OverflowRec *ovf; int *ovf_count; int ovf_cap;
    AlignParams P;
};

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

// One SHR candidate from a pre-decoded descriptor (aln.pyx:642-667 in gather form).
template <int NC>
__device__ __forceinline__ void shr_eval(uint32_t D, bool pred, int d, int bc, uint32_t sip, const float *rgM,
                                         const uint32_t *rgR, const uint32_t *s_magic, const float *__restrict__ np, int T, int cl,
                                         float &Sv, int &Sr, float &Sb)
{
    if (pred) {
        const int n = (int)(D & 7u);
        const bool start = (D & 8u) != 0u;
        const int at = ((d - n) & (NP_RING - 1)) * NC + (int)((D >> 11) & 0xffu);
        const float base = rgM[at + (start ? 0 : NP_RING * NC)];          // rgS follows rgM
        const int run0 = start ? 0 : (int)(rgR[at] >> 16);
        const bool ok = (bc > (int)((sip >> (4 * n)) & 7u)) && (start || run0 > 0);
        const int q = (int)__umulhi((uint32_t)run0 << 1, s_magic[n]);
        const int call = (int)((D >> 4) & 0x7fu) - q - 1;
        float sc = 100.f;
        if (call >= 0) sc = __ldg(np + (int)(D >> 19) * T + min(call, cl));
        const float cand = base + sc;
        if (ok && cand < Sv) { Sv = cand; Sr = run0 + n; Sb = base; }
    }
}

template <int CPL>
__global__ void __launch_bounds__(FWD_WARPS * 32) forward_kernel(const ForwardArgs a)
{
    constexpr int NC = 32 * CPL;
    constexpr int TBS = CPL <= 1 ? 1 : CPL <= 2 ? 2 : CPL <= 4 ? 4 : 8;
    extern __shared__ float smem[];
    __shared__ float s_sub[64];
    __shared__ uint32_t s_magic[8];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x < 64) {
        const int sb = threadIdx.x >> 3, rb = threadIdx.x & 7;
        s_sub[threadIdx.x] = (sb < 5 && rb < 5) ? a.sub_tab[sb * 5 + rb] : 0.f;
    }
    if (threadIdx.x < 8) s_magic[threadIdx.x] = threadIdx.x >= 1 ? (uint32_t)((0x80000000ull + threadIdx.x - 1) / threadIdx.x) : 0u;
    __syncthreads();

    float *rgM = smem + (size_t)warp * 4 * NP_RING * NC;       // MAT.VAL            [ring][slot]
    float *rgS = rgM + NP_RING * NC;                             // SHR run-start value (must follow rgM)
    float *rgL = rgS + NP_RING * NC;                             // LEN run-start value
    uint32_t *rgR = reinterpret_cast<uint32_t *>(rgL + NP_RING * NC);   // LEN.RUN | SHR.RUN << 16

    const int r = a.P.r, T = a.P.np_dim, cl = a.P.np_clamp;
    const float gopen = a.P.gap_open, gext = a.P.gap_ext;
    const uint32_t nmask = (1u << a.P.max_n) - 1u;
    const float *__restrict__ np = a.np_tab;
    const int src_lane = (lane + 31) & 31;

    for (;;) {
        int idx = 0;
        if (lane == 0) idx = atomicAdd(a.counter, 1);
        idx = __shfl_sync(NP_FULL, idx, 0);
        if (idx >= a.n) break;
        const int cid = a.order[idx];
        const ChunkDesc c = a.chunks[cid];
        if (!c.valid) { if (lane == 0) a.out[cid].score = 0.f; continue; }
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
        int Ir1[CPL], DM1[CPL], dgr[CPL];                 // INS.RUN; DEL.RUN<<13 | matrun; diag matrun
        uint4 cc[CPL]; uint32_t rw[CPL]; int bc[CPL];
#pragma unroll
        for (int k = 0; k < CPL; k++) {
            Mv1[k] = Iv1[k] = Dv1[k] = dgv[k] = 0.f; Ir1[k] = DM1[k] = dgr[k] = 0;
            const int s = lane * CPL + k;
            bc[k] = (s + r) & (NC - 1);
            const int j0 = bc[k] - r;
            rw[k] = (j0 <= 0) ? row[-j0] : 0u;
            cc[k] = (j0 > 0) ? col[j0] : (j0 == -r ? col[NC - r] : make_uint4(0u, 0u, 0u, 0u));
            if (j0 == 0) cc[k] = col[0];
        }
        uint32_t nrow = row[r + 1];                          // next row to enter the band (at b_col == 0)
        int wbase = c.brk >> 5; uint32_t wbuf = bits[wbase + lane];
        uint32_t cw = __shfl_sync(NP_FULL, wbuf, 0);
        uint32_t hist = 0; int Id = 0, Dd = 0;

        for (int d = 0; d < B; d++) {
            float lMv[CPL], lDv[CPL]; int lDM[CPL];
            if (d > 0) {
                const int g = c.brk + d - 1;                 // op that leads to this anti-diagonal
                if ((g & 31) == 0 && d > 1) {
                    int wi = (g >> 5) - wbase;
                    if (wi >= 32) { wbase += 32; wbuf = bits[wbase + lane]; wi -= 32; }
                    cw = __shfl_sync(NP_FULL, wbuf, wi);
                }
                const uint32_t o = (cw >> (g & 31)) & 1u;
                hist = ((hist << 1) | o) & 0x3fu;
                const float a0 = __shfl_sync(NP_FULL, Mv1[CPL - 1], src_lane);
                const float a1 = __shfl_sync(NP_FULL, Dv1[CPL - 1], src_lane);
                const int a2 = __shfl_sync(NP_FULL, DM1[CPL - 1], src_lane);
                const uint32_t a3 = __shfl_sync(NP_FULL, rw[CPL - 1], src_lane);
#pragma unroll
                for (int k = CPL - 1; k >= 0; k--) {
                    lMv[k] = k ? Mv1[k > 0 ? k - 1 : 0] : a0;
                    lDv[k] = k ? Dv1[k > 0 ? k - 1 : 0] : a1;
                    lDM[k] = k ? DM1[k > 0 ? k - 1 : 0] : a2;
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
                for (int k = 0; k < CPL; k++) { lMv[k] = lDv[k] = 0.f; lDM[k] = 0; }
            }
            const int jlo = Dd - r;
            const uint32_t sip = c_sipack[hist];
            const float infd = (float)(100 * d), edgev = (float)(100 * (d + 1));
            // interior-cell bounds on b_col for this anti-diagonal (aln.pyx:497-507)
            const int lo = max(1, max(Id + r - imax, r - Dd)), hi = min(2 * r - 1, min(Id + r, jmax + r - Dd));
            const bool steady = (Id > r) && (Dd > r) && (Id <= imax - r + 1) && (Dd <= jmax - r + 1);

            bool in[CPL];
            float Sv[CPL], Sb[CPL], Lv[CPL], Lb[CPL]; int Sr[CPL], Lr[CPL];
            bool p0[CPL], p1[CPL], pg[CPL], pl[CPL];
            bool any0 = false, any1 = false, anyg = false, anyl = false;
#pragma unroll
            for (int k = 0; k < CPL; k++) {
                in[k] = (unsigned)(bc[k] - lo) <= (unsigned)(hi - lo) && hi >= lo;
                Sv[k] = infd; Lv[k] = infd; Sb[k] = 0.f; Lb[k] = 0.f; Sr[k] = 0; Lr[k] = 0;
                const bool more = (cc[k].z & 24u) != 0u;          // >2 SHR candidates or >1 LEN-eligible period: generic path
                p0[k] = in[k] && !more && cc[k].x != 0u;
                p1[k] = in[k] && !more && cc[k].y != 0u;
                pg[k] = in[k] && more;
                const uint32_t ln = cc[k].w & 7u;
                pl[k] = in[k] && !more && ln != 0u && (((rw[k] << 1) >> ln) & 1u);
                any0 |= p0[k]; any1 |= p1[k]; anyg |= pg[k]; anyl |= pl[k];
            }
            // ---- SHR gather: descriptor 0 (largest period), then descriptor 1
            if (__any_sync(NP_FULL, any0)) {
#pragma unroll
                for (int k = 0; k < CPL; k++) shr_eval<NC>(cc[k].x, p0[k], d, bc[k], sip, rgM, rgR, s_magic, np, T, cl, Sv[k], Sr[k], Sb[k]);
            }
            if (__any_sync(NP_FULL, any1)) {
#pragma unroll
                for (int k = 0; k < CPL; k++) shr_eval<NC>(cc[k].y, p1[k], d, bc[k], sip, rgM, rgR, s_magic, np, T, cl, Sv[k], Sr[k], Sb[k]);
            }
            // ---- LEN gather (aln.pyx:602-633): single eligible period, 2-bit k-mer unit compare in registers
            if (__any_sync(NP_FULL, anyl)) {
#pragma unroll
                for (int k = 0; k < CPL; k++) {
                    if (pl[k]) {
                        const uint32_t D = cc[k].w;
                        const int n = (int)(D & 7u);
                        if (((cc[k].z >> 5) | (rw[k] >> 15)) & 1u) {
                            pg[k] = true; anyg = true;             // an N inside a k-mer: byte-wise compare on the generic path
                        } else {
                            const bool eq = ((((cc[k].z >> 6) ^ (rw[k] >> 16)) & ((1u << (2 * n)) - 1u)) == 0u);
                            const int sI = (int)((sip >> (4 * n)) & 7u);
                            const bool start = ((rw[k] >> (6 + n - 1)) & 1u) != 0u;
                            const int at = ((d - n) & (NP_RING - 1)) * NC + lane * CPL + k;
                            const float base = start ? rgM[at] : rgL[at];
                            const int run0 = start ? 0 : (int)(rgR[at] & 0xffffu);
                            const bool ok = eq && (bc[k] + n - sI <= 2 * r - 1) && (start || run0 > 0);
                            const int q = (int)__umulhi((uint32_t)run0 << 1, s_magic[n]);
                            const int call = (int)((D >> 4) & 0x7fu) + q + 1;
                            const float cand = base + __ldg(np + (int)(D >> 19) * T + min(call, cl));
                            if (ok && cand < Lv[k]) { Lv[k] = cand; Lr[k] = run0 + n; Lb[k] = base; }
                        }
                    }
                }
            }
            // ---- generic path (rare): all periods from the relaid byte record in global memory
            if (__any_sync(NP_FULL, anyg)) {
#pragma unroll
                for (int k = 0; k < CPL; k++) {
                    if (pg[k]) {
                        const int i = Id + r - bc[k], j = jlo + bc[k];
                        const uint2 rb = rel[j];
                        const bool more = (cc[k].z & 24u) != 0u;
                        if (more) {
                            for (int n = a.P.max_n; n >= 1; n--) {      // SHR, every period
                                const uint32_t byte = ((n <= 4 ? rb.x : rb.y) >> ((8 * (n - 1)) & 31)) & 0xffu;
                                const int L = (int)(byte & 0x7fu);
                                if (!L) continue;
                                const uint32_t D = (uint32_t)n | ((byte >> 7) << 3) | ((uint32_t)L << 4) |
                                                   ((uint32_t)((j - n) & (NC - 1)) << 11) | ((uint32_t)((n - 1) * T + min(L, cl)) << 19);
                                shr_eval<NC>(D, true, d, bc[k], sip, rgM, rgR, s_magic, np, T, cl, Sv[k], Sr[k], Sb[k]);
                            }
                        }
                        uint32_t lm = (rb.y >> 22) & rw[k] & 0x3fu & nmask;   // LEN, every eligible period
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
                            const int at = ((d - n) & (NP_RING - 1)) * NC + lane * CPL + k;
                            float base; int run0 = 0;
                            if ((rw[k] >> (6 + n - 1)) & 1u) base = rgM[at];
                            else { run0 = (int)(rgR[at] & 0xffffu); base = rgL[at]; if (run0 <= 0) continue; }
                            const int call = L + run0 / n + 1;
                            const float cand = base + __ldg(np + ((n - 1) * T + min(L, cl)) * T + min(call, cl));
                            if (cand < Lv[k]) { Lv[k] = cand; Lr[k] = run0 + n; Lb[k] = base; }
                        }
                    }
                }
            }

            // ---- INS / DEL / MAT (aln.pyx:525-592)
            uint32_t recs[CPL];
            float Mv[CPL], Iv[CPL], Dv[CPL]; int Ir[CPL], DM[CPL];
#pragma unroll
            for (int k = 0; k < CPL; k++) {
                // INS from top = own previous value; DEL from left
                const float iv1 = Mv1[k] + gopen, iv2 = Iv1[k] + gext;
                const bool ie = iv2 < iv1;
                Iv[k] = ie ? iv2 : iv1;
                Ir[k] = ie ? Ir1[k] + 1 : 1;
                const float dv1 = lMv[k] + gopen, dv2 = lDv[k] + gext;
                const bool de = dv2 < dv1;
                Dv[k] = de ? dv2 : dv1;
                int Dr = de ? (lDM[k] >> 13) + 1 : 1;
                int run = min(dgr[k] + 1, NP_RUN_SAT);
                float best = dgv[k] + s_sub[((rw[k] >> 12) & 7u) * 8 + (cc[k].z & 7u)];
                if (!steady) {
                    const int i = Id + r - bc[k], j = jlo + bc[k];
                    if (ie && i == 1) Ir[k] = 1;
                    if (de && j == 1) Dr = 1;
                    if (i == 0) { Iv[k] = (float)(100 * (j + 1)); Ir[k] = j; }
                    if (j == 0) { Dv[k] = (float)(100 * (i + 1)); Dr = i; }
                    if (!(i > 0 && j > 0)) { best = Dv[k] + 100.f; run = 0; }
                }
                uint32_t pk = (uint32_t)run << 3;                                  // typ MAT = 0
                if (Iv[k] < best) { best = Iv[k]; pk = ((uint32_t)Ir[k] << 3) | T_INS; }
                if (Lv[k] < best) { best = Lv[k]; pk = ((uint32_t)Lr[k] << 3) | T_LEN; }
                if (Dv[k] < best) { best = Dv[k]; pk = ((uint32_t)Dr << 3) | T_DEL; }
                if (Sv[k] < best) { best = Sv[k]; pk = ((uint32_t)Sr[k] << 3) | T_SHR; }
                if (pk >= (((uint32_t)NP_RUN_SAT << 3) | 1u) && in[k]) {          // non-MAT run beyond the 13-bit field
                    if ((pk & 7u) != 0u) {
                        const int pos = atomicAdd(a.ovf_count, 1);
                        if (pos < a.ovf_cap) { OverflowRec ov; ov.chunk = cid; ov.d = d; ov.bc = lane * CPL + k; ov.run = (int)(pk >> 3); a.ovf[pos] = ov; }
                        pk = ((uint32_t)NP_RUN_SAT << 3) | (pk & 7u);
                    }
                }
                // EDGE (b_col 0 / 2r): every state INF*(b_row+1), TYP MAT, RUN 0 (aln.pyx:502-507).  Cells outside the chunk
                // (aln.pyx:497-499) are never read by interior cells; they get the same harmless value.
                Mv[k] = in[k] ? best : edgev;
                Iv[k] = in[k] ? Iv[k] : edgev;
                Dv[k] = in[k] ? Dv[k] : edgev;
                Ir[k] = in[k] ? Ir[k] : 0;
                pk = in[k] ? pk : 0u;
                recs[k] = pk;
                const int mr = (pk & 7u) ? 0 : (int)(pk >> 3);
                DM[k] = ((in[k] ? Dr : 0) << 13) | mr;
                if (!in[k]) { Sb[k] = Lb[k] = 0.f; Sr[k] = Lr[k] = 0; }
            }

            // ---- history ring + traceback row (slot order)
            {
                const int at = (d & (NP_RING - 1)) * NC + lane * CPL;
#pragma unroll
                for (int k = 0; k < CPL; k++) {
                    rgM[at + k] = Mv[k]; rgS[at + k] = Sb[k]; rgL[at + k] = Lb[k];
                    rgR[at + k] = (uint32_t)Lr[k] | ((uint32_t)Sr[k] << 16);
                }
                uint16_t *rowp = tbp + (size_t)d * (32 * TBS);
                if (CPL == 2) *reinterpret_cast<uint32_t *>(rowp) = recs[0] | (recs[CPL - 1] << 16);
                else {
#pragma unroll
                    for (int k = 0; k < CPL; k++) rowp[k] = (uint16_t)recs[k];
                }
            }
            __syncwarp();
#pragma unroll
            for (int k = 0; k < CPL; k++) {
                dgv[k] = lMv[k]; dgr[k] = lDM[k] & 8191;          // next step's diagonal neighbour = this step's left
                Mv1[k] = Mv[k]; Iv1[k] = Iv[k]; Dv1[k] = Dv[k]; Ir1[k] = Ir[k]; DM1[k] = DM[k];
            }
        }
        // chunk score = MAT value of the end cell (b_col == r on the last anti-diagonal)
#pragma unroll
        for (int k = 0; k < CPL; k++)
            if (bc[k] == r) a.out[cid].score = Mv1[k];
        __syncwarp();
    }
}
