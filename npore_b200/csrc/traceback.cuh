// traceback.cuh -- align() traceback (reference: src/aln.pyx:671-742) over the packed (TYP:3,RUN:13) records
// the forward kernel streamed to HBM.  Separate kernel, one walker thread per chunk: pointer chasing that
// jumps RUN cells at a time (MAT-typed runs are diagonal steps re-evaluated as '='/'X' by base compare,
// aln.pyx:727-736).  Ops are emitted right-aligned into the chunk's slice of the item's op scratch
// (a chunk covering anti-diagonals [brk, nxt] emits at most nxt-brk ops) and compacted by finish.cuh.
// Anomalies map to the reference's three checks + unknown type (aln.pyx:689-716, 737-739) as status 1..4;
// the partial op string is kept, as there.
#pragma once
#include "common.cuh"

#define TB_THREADS 64

struct TracebackArgs {
    const ChunkDesc *chunks;
    const ChunkSlot *slots;
    const int32_t *order;
    int n;
    const ItemDesc *items;
    const uint32_t *bits, *cum;
    const uint8_t *ref_codes, *seq_codes;
    const uint16_t *tb;
    uint8_t *ops;                 // op scratch, item regions at ItemDesc::out_off
    ChunkOut *out;
    const OverflowRec *ovf; const int *ovf_count; int ovf_cap;
    int r, W, cpl, tbs;
};

__global__ void __launch_bounds__(TB_THREADS) traceback_kernel(const TracebackArgs a)
{
    const int idx = blockIdx.x * TB_THREADS + threadIdx.x;
    if (idx >= a.n) return;
    const int cid = a.order[idx];
    const ChunkDesc c = a.chunks[cid];
    if (!c.valid) { ChunkOut o; o.score = 0.f; o.status = 0; o.start = 0; o.len = 0; a.out[cid] = o; return; }
    const ChunkSlot sl = a.slots[idx];
    const ItemDesc &I = a.items[c.item];
    const uint8_t *refs = a.ref_codes + I.ref_start + min(c.c0, I.ref_len);
    const uint8_t *seqs = a.seq_codes + I.seq_start + min(c.r0, I.seq_len);
    const uint16_t *tb = a.tb + (size_t)sl.tb_off * (32 * a.tbs);
    uint8_t *region = a.ops + I.out_off + c.brk;
    const int cap = c.B - 1;
    int pos = cap, i = c.imax, j = c.jmax, status = 0;
    while (i > 0 || j > 0) {
        if (i < 0) { status = 1; break; }
        if (j < 0) { status = 2; break; }
        // records are stored per anti-diagonal in slot order: slot = column index mod NC (forward.cuh)
        const int d = i + j, bc = j & (32 * a.cpl - 1);
        uint32_t rec = 0;
        if (d < c.B) rec = tb[(size_t)d * (32 * a.tbs) + (bc / a.cpl) * a.tbs + (bc % a.cpl)];
        const int typ = (int)(rec & 7u);
        int run = (int)(rec >> 3);
        if (typ != T_MAT && run == NP_RUN_SAT) {
            const int m = min(*a.ovf_count, a.ovf_cap);
            for (int t = 0; t < m; t++)
                if (a.ovf[t].chunk == cid && a.ovf[t].d == d && a.ovf[t].bc == bc) { run = a.ovf[t].run; break; }
        }
        if (run < 1) { status = 3; break; }
        if (typ == T_INS || typ == T_LEN) {
            for (int t = 0; t < run && pos > 0; t++) region[--pos] = 'I';
            i -= run;
        } else if (typ == T_DEL || typ == T_SHR) {
            for (int t = 0; t < run && pos > 0; t++) region[--pos] = 'D';
            j -= run;
        } else if (typ == T_MAT) {
            for (int t = 0; t < run; t++) {
                i--; j--;
                if (i < 0 || j < 0) break;
                if (pos > 0) region[--pos] = (refs[j] == seqs[i]) ? '=' : 'X';
            }
        } else { status = 4; break; }
    }
    ChunkOut o;
    o.score = a.out[cid].score;
    o.status = status; o.start = c.brk + pos; o.len = cap - pos;
    a.out[cid] = o;
}
