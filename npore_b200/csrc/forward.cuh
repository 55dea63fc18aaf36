// forward.cuh -- the align() recurrence (reference: src/aln.pyx:465-667) as an anti-diagonal wavefront.
//
// One warp owns one chunk at a time (persistent warps pull chunk indices from a global counter, largest
// chunks first).  The band of W = 2r+1 cells of one anti-diagonal is laid out BLOCKED over the warp:
// lane l holds cells b_col = l*CPL .. l*CPL+CPL-1 in registers (CPL = ceil(W/32): 2 at the default r=30).
// Per anti-diagonal d (b_row of aln.pyx:481):
//   * neighbours: the coordinate transforms of aln.pyx:485-492 collapse to shifts that are uniform over the
//     diagonal: op(d)=='I' -> top = same b_col, left = b_col-1; op(d)=='D' -> left = same, top = b_col+1;
//     diag = the neighbour fetched one step earlier when op(d)==op(d-1), else the own cell.  In the blocked
//     layout a +-1 shift is a register rename plus ONE warp shuffle per quantity (3 per diagonal).
//   * per-column / per-row context (np-info bytes, base codes) rides along in registers and is shifted the
//     same way; new records enter at the band edge from a lane-distributed, double-buffered prefetch.
//   * LEN / SHR (aln.pyx:596-667 scatter) in gather form, n descending, strict '<': the candidate of period n
//     reads the source cell n anti-diagonals back from an 8-deep shared-memory ring holding MAT.VAL, the run
//     lengths and the value at the run's start ("BASE", which replaces the lookback of aln.pyx:623-629,657-663).
//   * MAT's packed (TYP:3, RUN:13) record -- the only thing traceback reads (aln.pyx:683-685) -- is streamed to
//     HBM, one fully coalesced 64*CPL-byte row per anti-diagonal.
// Arithmetic: fp32 add and strict compare only (no FMA contraction possible), tie-break order of aln.pyx:585-592.
// The statement-by-statement CPU model of this kernel is oracle/pull_model.c:pm_align.
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
    const uint2 *colrec;
    const uint32_t *rowrec;
    uint16_t *tb;
    const float *np_tab;          // [np_n][np_dim][np_dim]
    const float *sub_tab;         // [5][5]  indexed [seq_base][ref_base]
    ChunkOut *out;                // indexed by chunk id
    OverflowRec *ovf; int *ovf_count; int ovf_cap;
    AlignParams P;
};

template <int CPL>
__global__ void __launch_bounds__(FWD_WARPS * 32) forward_kernel(const ForwardArgs a)
{
    constexpr int NC = 32 * CPL;                 // physical cells per anti-diagonal (>= W)
    constexpr int TBS = CPL <= 1 ? 1 : CPL <= 2 ? 2 : CPL <= 4 ? 4 : 8;
    extern __shared__ float smem[];
    __shared__ float s_sub[64];
    __shared__ uint32_t s_magic[8];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x < 64) {
        const int sb = threadIdx.x >> 3, rb = threadIdx.x & 7;
        s_sub[threadIdx.x] = (sb < 5 && rb < 5) ? a.sub_tab[sb * 5 + rb] : 0.f;
    }
    if (threadIdx.x < 8) s_magic[threadIdx.x] = threadIdx.x >= 2 ? (uint32_t)(0x100000000ull / threadIdx.x) + 1u : 0u;
    __syncthreads();

    float *rgM = smem + (size_t)warp * 4 * NP_RING * NC;
    float *rgS = rgM + NP_RING * NC;
    float *rgL = rgS + NP_RING * NC;
    uint32_t *rgR = reinterpret_cast<uint32_t *>(rgL + NP_RING * NC);

    const int r = a.P.r, W = a.P.W, T = a.P.np_dim, cl = a.P.np_clamp;
    const float gopen = a.P.gap_open, gext = a.P.gap_ext;
    const uint32_t nmask = (1u << a.P.max_n) - 1u;
    const float *__restrict__ np = a.np_tab;

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
        const uint2 *__restrict__ col = a.colrec + sl.col_off;
        const uint32_t *__restrict__ row = a.rowrec + sl.row_off;
        const uint8_t *__restrict__ refs = a.ref_codes + I.ref_start + min(c.c0, I.ref_len);
        const uint8_t *__restrict__ seqs = a.seq_codes + I.seq_start + min(c.r0, I.seq_len);
        uint16_t *tbp = a.tb + (size_t)sl.tb_off * (32 * TBS) + lane * TBS;
        const int B = c.B, imax = c.imax, jmax = c.jmax;

        // ---- carried per-cell state (previous anti-diagonal) and contexts
        float Mv1[CPL], Iv1[CPL], Dv1[CPL], Dgv[CPL];
        int Mr1[CPL], Ir1[CPL], Dr1[CPL], Dgr[CPL];
        uint2 cc[CPL]; uint32_t rw[CPL];
#pragma unroll
        for (int k = 0; k < CPL; k++) {
            Mv1[k] = Iv1[k] = Dv1[k] = Dgv[k] = 0.f; Mr1[k] = Ir1[k] = Dr1[k] = Dgr[k] = 0;
            const int bc = lane * CPL + k, i0 = r - bc, j0 = bc - r;
            rw[k] = i0 >= 0 ? row[i0] : 0u;
            cc[k] = j0 >= 0 ? col[j0] : make_uint2(0u, 0u);
        }
        // band-edge prefetch streams (lane-distributed, double-buffered)
        int cnext = NC - r, cbase = cnext; uint2 cbufA = col[cbase + lane], cbufB = col[cbase + 32 + lane];
        int rnext = r + 1, rbase = rnext; uint32_t rbufA = row[rbase + lane], rbufB = row[rbase + 32 + lane];
        // op bit stream
        int wbase = c.brk >> 5; uint32_t wbuf = bits[wbase + lane];
        uint32_t cw = __shfl_sync(NP_FULL, wbuf, 0);
        uint32_t o = 0, on = (cw >> (c.brk & 31)) & 1u;      // on = op of step d=1
        uint32_t hist = 0; int Id = 0;

        for (int d = 0; d < B; d++) {
            float tMv[CPL], tIv[CPL], lMv[CPL], lDv[CPL], sMv[CPL];
            int tIr[CPL], lDr[CPL], sMr[CPL];
            if (d > 0) {
                o = on;
                {   // look-ahead op bit (op index brk + d), needed for the diag selection of the next step
                    const int g = c.brk + d;
                    if ((g & 31) == 0) {
                        int wi = (g >> 5) - wbase;
                        if (wi >= 32) { wbase += 32; wbuf = bits[wbase + lane]; wi -= 32; }
                        cw = __shfl_sync(NP_FULL, wbuf, wi);
                    }
                    on = (cw >> (g & 31)) & 1u;
                }
                hist = ((hist << 1) | o) & 0xffu;
                Id += (int)o;
                if (o) {
                    // 'I': rows advance. top = same b_col, left = b_col-1
                    const float a0 = __shfl_up_sync(NP_FULL, Mv1[CPL - 1], 1);
                    const float a1 = __shfl_up_sync(NP_FULL, Dv1[CPL - 1], 1);
                    const int a2 = __shfl_up_sync(NP_FULL, (Dr1[CPL - 1] << 13) | Mr1[CPL - 1], 1);
                    const uint32_t a3 = __shfl_up_sync(NP_FULL, rw[CPL - 1], 1);
                    const uint32_t nr = __shfl_sync(NP_FULL, rbufA, rnext - rbase);
#pragma unroll
                    for (int k = CPL - 1; k >= 0; k--) {
                        tMv[k] = Mv1[k]; tIv[k] = Iv1[k]; tIr[k] = Ir1[k];
                        lMv[k] = k ? Mv1[k - 1] : a0; lDv[k] = k ? Dv1[k - 1] : a1;
                        lDr[k] = k ? Dr1[k - 1] : (a2 >> 13); sMr[k] = k ? Mr1[k - 1] : (a2 & 8191);
                        sMv[k] = lMv[k];
                        rw[k] = k ? rw[k - 1] : a3;
                    }
                    if (lane == 0) rw[0] = nr;
                    rnext++;
                    if (rnext - rbase == 32) { rbufA = rbufB; rbase += 32; rbufB = row[rbase + 32 + lane]; }
                } else {
                    // 'D': columns advance. left = same b_col, top = b_col+1
                    const float a0 = __shfl_down_sync(NP_FULL, Mv1[0], 1);
                    const float a1 = __shfl_down_sync(NP_FULL, Iv1[0], 1);
                    const int a2 = __shfl_down_sync(NP_FULL, (Ir1[0] << 13) | Mr1[0], 1);
                    const uint32_t a3 = __shfl_down_sync(NP_FULL, cc[0].x, 1);
                    const uint32_t a4 = __shfl_down_sync(NP_FULL, cc[0].y, 1);
                    const uint32_t n0 = __shfl_sync(NP_FULL, cbufA.x, cnext - cbase);
                    const uint32_t n1 = __shfl_sync(NP_FULL, cbufA.y, cnext - cbase);
#pragma unroll
                    for (int k = 0; k < CPL; k++) {
                        lMv[k] = Mv1[k]; lDv[k] = Dv1[k]; lDr[k] = Dr1[k];
                        tMv[k] = (k < CPL - 1) ? Mv1[k + 1 < CPL ? k + 1 : k] : a0;
                        tIv[k] = (k < CPL - 1) ? Iv1[k + 1 < CPL ? k + 1 : k] : a1;
                        tIr[k] = (k < CPL - 1) ? Ir1[k + 1 < CPL ? k + 1 : k] : (a2 >> 13);
                        sMr[k] = (k < CPL - 1) ? Mr1[k + 1 < CPL ? k + 1 : k] : (a2 & 8191);
                        sMv[k] = tMv[k];
                        cc[k] = (k < CPL - 1) ? cc[k + 1 < CPL ? k + 1 : k] : make_uint2(a3, a4);
                    }
                    if (lane == 31) cc[CPL - 1] = make_uint2(n0, n1);
                    cnext++;
                    if (cnext - cbase == 32) { cbufA = cbufB; cbase += 32; cbufB = col[cbase + 32 + lane]; }
                }
            } else {
#pragma unroll
                for (int k = 0; k < CPL; k++) { tMv[k] = tIv[k] = lMv[k] = lDv[k] = sMv[k] = 0.f; tIr[k] = lDr[k] = sMr[k] = 0; }
            }

            // ---- cell classes (aln.pyx:497-507)
            const int Dd = d - Id;
            const float infd = (float)(100 * d);
            bool in[CPL]; int ci[CPL], cj[CPL];
            uint32_t sm[CPL], lm[CPL];
            float Sv[CPL], Sb[CPL], Lv[CPL], Lb[CPL]; int Sr[CPL], Lr[CPL];
            uint32_t anyS = 0, anyL = 0;
#pragma unroll
            for (int k = 0; k < CPL; k++) {
                const int bc = lane * CPL + k;
                ci[k] = Id + r - bc; cj[k] = Dd - r + bc;
                const bool outc = !(bc < W && ci[k] >= 0 && cj[k] >= 0 && ci[k] <= imax && cj[k] <= jmax);
                in[k] = !outc && bc != 0 && bc != 2 * r;
                sm[k] = in[k] ? ((cc[k].y >> 16) & 0x3fu & nmask) : 0u;
                lm[k] = in[k] ? ((cc[k].y >> 22) & rw[k] & 0x3fu & nmask) : 0u;
                anyS |= sm[k]; anyL |= lm[k];
                Sv[k] = infd; Lv[k] = infd; Sb[k] = 0.f; Lb[k] = 0.f; Sr[k] = 0; Lr[k] = 0;
            }

            // ---- SHR gather (aln.pyx:642-667), n descending
            while (__any_sync(NP_FULL, anyS != 0u)) {
                anyS = 0;
#pragma unroll
                for (int k = 0; k < CPL; k++) {
                    if (sm[k]) {
                        const int n = 32 - __clz(sm[k]);
                        sm[k] &= ~(1u << (n - 1));
                        const uint32_t byte = ((n <= 4 ? cc[k].x : cc[k].y) >> ((8 * (n - 1)) & 31)) & 0xffu;
                        const int L = (int)(byte & 0x7fu);
                        const int sb = lane * CPL + k - __popc(hist & ((1u << n) - 1u));
                        if (sb >= 1) {
                            const int at = ((d - n) & (NP_RING - 1)) * NC + sb;
                            float base; int run0 = 0; bool ok = true;
                            if (byte & 0x80u) base = rgM[at];
                            else { run0 = (int)(rgR[at] >> 16); base = rgS[at]; ok = run0 > 0; }
                            if (ok) {
                                const int q = (n == 1) ? run0 : (int)__umulhi((uint32_t)run0, s_magic[n]);
                                const int call = L - q - 1;
                                float sc = 100.f;
                                if (call >= 0) sc = __ldg(np + ((size_t)(n - 1) * T + min(L, cl)) * T + min(call, cl));
                                const float cand = base + sc;
                                if (cand < Sv[k]) { Sv[k] = cand; Sr[k] = run0 + n; Sb[k] = base; }
                            }
                        }
                    }
                    anyS |= sm[k];
                }
            }
            // ---- LEN gather (aln.pyx:602-633), n descending; rare
            if (__any_sync(NP_FULL, anyL != 0u)) {
#pragma unroll
                for (int k = 0; k < CPL; k++) {
                    while (lm[k]) {
                        const int n = 32 - __clz(lm[k]);
                        lm[k] &= ~(1u << (n - 1));
                        const int sb = lane * CPL + k + n - __popc(hist & ((1u << n) - 1u));
                        if (sb > 2 * r - 1) continue;
                        const int si = ci[k] - n, j = cj[k];
                        bool eq = true;
                        for (int t = 0; t < n; t++) eq = eq && (seqs[si + t] == refs[j + t]);
                        if (!eq) continue;
                        const uint2 cjn = col[j + n];
                        const int L = (int)(((n <= 4 ? cjn.x : cjn.y) >> ((8 * (n - 1)) & 31)) & 0x7fu);
                        const int at = ((d - n) & (NP_RING - 1)) * NC + sb;
                        float base; int run0 = 0;
                        if ((rw[k] >> (8 + n - 1)) & 1u) base = rgM[at];
                        else { run0 = (int)(rgR[at] & 0xffffu); base = rgL[at]; if (run0 <= 0) continue; }
                        const int call = L + run0 / n + 1;
                        const float sc = __ldg(np + ((size_t)(n - 1) * T + min(L, cl)) * T + min(call, cl));
                        const float cand = base + sc;
                        if (cand < Lv[k]) { Lv[k] = cand; Lr[k] = run0 + n; Lb[k] = base; }
                    }
                }
            }

            // ---- INS / DEL / MAT (aln.pyx:525-592)
            uint32_t recs[CPL];
            float Mv[CPL], Iv[CPL], Dv[CPL]; int Mr[CPL], Ir[CPL], Dr[CPL];
#pragma unroll
            for (int k = 0; k < CPL; k++) {
                const int i = ci[k], j = cj[k];
                if (i == 0) { Iv[k] = (float)(100 * (j + 1)); Ir[k] = j; }
                else {
                    const float v1 = tMv[k] + gopen, v2 = tIv[k] + gext;
                    if (v2 < v1) { Iv[k] = v2; Ir[k] = (i == 1) ? 1 : tIr[k] + 1; } else { Iv[k] = v1; Ir[k] = 1; }
                }
                if (j == 0) { Dv[k] = (float)(100 * (i + 1)); Dr[k] = i; }
                else {
                    const float v1 = lMv[k] + gopen, v2 = lDv[k] + gext;
                    if (v2 < v1) { Dv[k] = v2; Dr[k] = (j == 1) ? 1 : lDr[k] + 1; } else { Dv[k] = v1; Dr[k] = 1; }
                }
                float best; int typ = T_MAT, run = 0;
                if (i > 0 && j > 0) {
                    run = min(Dgr[k] + 1, NP_RUN_SAT);
                    best = Dgv[k] + s_sub[((rw[k] >> 16) & 7u) * 8 + (cc[k].y >> 28)];
                } else best = Dv[k] + 100.f;
                if (Iv[k] < best) { best = Iv[k]; typ = T_INS; run = Ir[k]; }
                if (Lv[k] < best) { best = Lv[k]; typ = T_LEN; run = Lr[k]; }
                if (Dv[k] < best) { best = Dv[k]; typ = T_DEL; run = Dr[k]; }
                if (Sv[k] < best) { best = Sv[k]; typ = T_SHR; run = Sr[k]; }
                Mv[k] = best; Mr[k] = (typ == T_MAT) ? run : 0;
                if (in[k] && typ != T_MAT && run >= NP_RUN_SAT) {
                    const int pos = atomicAdd(a.ovf_count, 1);
                    if (pos < a.ovf_cap) { OverflowRec ov; ov.chunk = cid; ov.d = d; ov.bc = lane * CPL + k; ov.run = run; a.ovf[pos] = ov; }
                    run = NP_RUN_SAT;
                }
                recs[k] = (uint32_t)typ | ((uint32_t)run << 3);
                if (!in[k]) {
                    // EDGE: every state = INF*(b_row+1), TYP=MAT, RUN=0 (aln.pyx:502-507); OUT: untouched zeros (aln.pyx:497-499)
                    const int bc = lane * CPL + k;
                    const bool edge = (bc == 0 || bc == 2 * r) && i >= 0 && j >= 0 && i <= imax && j <= jmax;
                    const float v = edge ? (float)(100 * (d + 1)) : 0.f;
                    Mv[k] = Iv[k] = Dv[k] = v; Mr[k] = Ir[k] = Dr[k] = 0; recs[k] = 0u;
                    Sb[k] = Lb[k] = 0.f; Sr[k] = Lr[k] = 0;
                }
            }

            // ---- history ring + traceback row
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

            // ---- rotate: next step's diagonal neighbour is MAT[d-1] shifted like this step iff op(d+1)==op(d)
#pragma unroll
            for (int k = 0; k < CPL; k++) {
                const bool same = (d > 0) && (on == o);
                Dgv[k] = same ? sMv[k] : Mv1[k];
                Dgr[k] = same ? sMr[k] : Mr1[k];
                Mv1[k] = Mv[k]; Iv1[k] = Iv[k]; Dv1[k] = Dv[k]; Mr1[k] = Mr[k]; Ir1[k] = Ir[k]; Dr1[k] = Dr[k];
            }
        }
        // chunk score = MAT value of the end cell (b_col == r on the last anti-diagonal)
#pragma unroll
        for (int k = 0; k < CPL; k++)
            if (lane * CPL + k == r) a.out[cid].score = Mv1[k];
        __syncwarp();
    }
}
