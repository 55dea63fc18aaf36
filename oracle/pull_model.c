/*
 * pull_model.c -- CPU model of the *GPU kernels' dataflow*.  TEST INFRASTRUCTURE ONLY (design validation).
 *
 * npore_b200/csrc/ computes align() (reference: /root/reference/src/aln.pyx:379-787) in a form that differs
 * from the reference's scatter loops:
 *   - one anti-diagonal at a time, every band cell of the diagonal computed from earlier diagonals only,
 *   - LEN/SHR in gather ("pull") form, n descending, with the run's start value carried along (BASE)
 *     instead of looked up (aln.pyx:623-629, 657-663),
 *   - get_np_info (aln.pyx:179-251) as position-parallel run scans + one walker per phase chain,
 *   - per-position packed records: colrec (8 B, ref side) and rowrec (4 B, read side), "relaid" so that the
 *     record of column j holds what the SHR gather of cell (.,j) needs from columns j-1..j-6.
 * This file executes exactly that dataflow sequentially so it can be checked against oracle/npore_oracle.c
 * (and thereby the reference) on the CPU, before/independently of the CUDA transcription.  It pins the ALGORITHMIC
 * form the kernels in npore_b200/csrc/{annotate,forward,traceback}.cuh use (gather + carried BASE, relaid np-info,
 * chain-walker np_info, packed traceback record); the kernels' lane layout, descriptor encodings and scheduling have
 * since moved on and are checked by the GPU parity tests.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define MAXN 6
enum { T_MAT = 0, T_INS = 1, T_LEN = 2, T_DEL = 3, T_SHR = 4 };

/* ---------------------------------------------------------------- annotate (np_info, parallel form) */
/* raw[p*8 + (n-1)] = L | (X==0 ? 0x80 : 0)   (L in 0..127) ; also returns X for checking when xs != NULL */
void pm_np_raw(const uint8_t *s, int len, int max_n, int max_l, uint8_t *raw, int32_t *xs)
{
    memset(raw, 0, (size_t)(len > 0 ? len : 1) * 8);
    if (xs) memset(xs, 0, (size_t)(len > 0 ? len : 1) * 8 * sizeof(int32_t));
    int32_t *nf = (int32_t *)malloc(sizeof(int32_t) * (size_t)(len + 1));   /* next false position >= q */
    int32_t *lf = (int32_t *)malloc(sizeof(int32_t) * (size_t)(len + 1));   /* last false position  <  q, or -1 */
    for (int n = 1; n <= max_n; n++) {
        /* e[q] = (q+n < len && s[q]==s[q+n]) */
        int nxt = len;                      /* virtual false at len (never reached: e[len-n..len-1] are false) */
        for (int q = len - 1; q >= 0; q--) {
            int e = (q + n < len) && (s[q] == s[q + n]);
            if (!e) nxt = q;
            nf[q] = nxt;
        }
        int last = -1;
        for (int q = 0; q < len; q++) {
            lf[q] = last;
            int e = (q + n < len) && (s[q] == s[q + n]);
            if (!e) last = q;
        }
        /* one walker per chain head */
        for (int h = 0; h < len; h++) {
            int rs = lf[h] + 1;
            if (h - rs >= n) continue;              /* not a chain head */
            int end = nf[rs];                       /* terminal (false) position of this run; chain members <= end */
            int first = -1, lfirst = 0, zlast = -1;
            for (int p = h; p <= end; p += n) {
                int m = end - p;
                int l = m / n; if (l > 0) l++;
                int act = s[p] != 0 && l > 2;
                if (act) {
                    for (int n2 = 1; n2 < n; n2++)
                        if (l * n <= (raw[(size_t)p * 8 + n2 - 1] & 0x7f) * n2) act = 0;
                }
                if (act) {
                    if (first < 0) { first = p; lfirst = l; }
                    if (l > max_l) zlast = p;
                }
                if (first >= 0) {
                    int L, Xv;
                    if (zlast >= 0) { L = max_l; Xv = (p - zlast) / n; }
                    else { L = lfirst < max_l ? lfirst : max_l; Xv = (p - first) / n; }
                    raw[(size_t)p * 8 + n - 1] = (uint8_t)(L | (Xv == 0 ? 0x80 : 0));
                    if (xs) xs[(size_t)p * 8 + n - 1] = Xv;
                }
            }
        }
    }
    free(nf); free(lf);
}

/* colrec[j], j in [0, rlen+8): bytes0-5 = raw[j-n][n]; byte6 = LEN-eligible mask of raw[j]; byte7 = base s[j-1] */
static void relay_col(const uint8_t *s, int len, const uint8_t *raw, uint64_t *colrec)
{
    for (int j = 0; j < len + 8; j++) {
        uint64_t v = 0;
        for (int n = 1; n <= MAXN; n++)
            if (j - n >= 0 && j - n < len) v |= (uint64_t)raw[(size_t)(j - n) * 8 + n - 1] << (8 * (n - 1));
        if (j < len)
            for (int n = 1; n <= MAXN; n++) {
                uint8_t b = raw[(size_t)j * 8 + n - 1];
                if ((b & 0x7f) && (b & 0x80)) v |= (uint64_t)1 << (48 + n - 1);
            }
        if (j >= 1 && j - 1 < len) v |= (uint64_t)s[j - 1] << 56;
        colrec[j] = v;
    }
}
/* rowrec[i]: bits0-5 = (Lq[i-n][n] != 0); bits8-13 = (Xq[i-n][n]==0); bits16-18 = base s[i-1] */
static void relay_row(const uint8_t *s, int len, const uint8_t *raw, uint32_t *rowrec)
{
    for (int i = 0; i < len + 8; i++) {
        uint32_t v = 0;
        for (int n = 1; n <= MAXN; n++)
            if (i - n >= 0 && i - n < len) {
                uint8_t b = raw[(size_t)(i - n) * 8 + n - 1];
                if (b & 0x7f) v |= 1u << (n - 1);
                if (b & 0x80) v |= 1u << (8 + n - 1);
            }
        if (i >= 1 && i - 1 < len) v |= (uint32_t)s[i - 1] << 16;
        rowrec[i] = v;
    }
}

static inline float np_lookup(const float *np, int T, int clampv, int n, int L, int call)
{
    if (call < 0) return 100.0f;                /* L > 0 guaranteed by the candidate masks */
    int a = L < clampv ? L : clampv, b = call < clampv ? call : clampv;
    return np[((size_t)(n - 1) * T + a) * T + b];
}

/* ---------------------------------------------------------------- whole align(), chunk by chunk */
int64_t pm_align(const uint8_t *full_ref, int Lr, const uint8_t *full_seq, int Ls,
                 const char *cigar, int64_t cig_len, const float *sub, const float *np, int np_dim,
                 int max_n, int max_l, float gap_open, float gap_ext, int max_b_rows, int r,
                 char *out, int64_t out_cap, float *scores, int scores_cap, int *n_scores, int *status)
{
    /* op bit string: 1 = I.  X,=,M -> D,I */
    int64_t P = 0;
    uint8_t *opI = (uint8_t *)malloc((size_t)cig_len * 2 + 2);
    for (int64_t k = 0; k < cig_len; k++) {
        char c = cigar[k];
        if (c == 'I') opI[P++] = 1; else if (c == 'D') opI[P++] = 0; else { opI[P++] = 0; opI[P++] = 1; }
    }
    int32_t *inss = (int32_t *)calloc((size_t)P + 2, sizeof(int32_t));
    for (int64_t k = 0; k < P; k++) inss[k + 1] = inss[k] + opI[k];
    const int total = Ls + Lr;                       /* == P for a consistent CIGAR */
    const int step = max_b_rows - 1;
    const int nchunks = total > 0 ? (total + step - 1) / step : 0;
    const int W = 2 * r + 1, clampv = max_l - 1;
    int64_t out_len = 0; int st = 0, nsc = 0;
    char *rev = (char *)malloc((size_t)total + 8);

    for (int ci = 0; ci < nchunks; ci++) {
        int brk = ci * step, nxt = (ci + 1 < nchunks) ? (ci + 1) * step : total;
        if (ci > 0 && opI[brk] && !opI[brk - 1]) brk--;
        if (ci + 1 < nchunks && opI[nxt] && !opI[nxt - 1]) nxt--;
        const int B = nxt - brk + 1;
        const int r0 = inss[brk], c0 = brk - r0, r1 = inss[nxt], c1 = nxt - r1;
        const int imax = r1 - r0, jmax = c1 - c0;
        int rlen = (c1 + 1 < Lr ? c1 + 1 : Lr) - c0; if (rlen < 0) rlen = 0;
        int slen = (r1 + 1 < Ls ? r1 + 1 : Ls) - r0; if (slen < 0) slen = 0;
        const uint8_t *ref = full_ref + (c0 < Lr ? c0 : Lr), *seq = full_seq + (r0 < Ls ? r0 : Ls);

        uint8_t *rawr = (uint8_t *)malloc((size_t)(rlen + 1) * 8), *raws = (uint8_t *)malloc((size_t)(slen + 1) * 8);
        pm_np_raw(ref, rlen, max_n, max_l, rawr, NULL);
        pm_np_raw(seq, slen, max_n, max_l, raws, NULL);
        uint64_t *colrec = (uint64_t *)malloc(sizeof(uint64_t) * (size_t)(rlen + 8));
        uint32_t *rowrec = (uint32_t *)malloc(sizeof(uint32_t) * (size_t)(slen + 8));
        relay_col(ref, rlen, rawr, colrec);
        relay_row(seq, slen, raws, rowrec);

        uint16_t *tb = (uint16_t *)calloc((size_t)B * W, sizeof(uint16_t));
        /* previous-diagonal state (index = bc) and 8-deep rings */
        float *Mv1 = calloc(W, 4), *Iv1 = calloc(W, 4), *Dv1 = calloc(W, 4), *Mv2 = calloc(W, 4);
        int *Mr1 = calloc(W, 4), *Ir1 = calloc(W, 4), *Dr1 = calloc(W, 4), *Mr2 = calloc(W, 4);
        float *nMv = calloc(W, 4), *nIv = calloc(W, 4), *nDv = calloc(W, 4);
        int *nMr = calloc(W, 4), *nIr = calloc(W, 4), *nDr = calloc(W, 4);
        float (*rgM)[512] = calloc(8, sizeof(*rgM)), (*rgS)[512] = calloc(8, sizeof(*rgS)), (*rgL)[512] = calloc(8, sizeof(*rgL));
        int (*rgLr)[512] = calloc(8, sizeof(*rgLr)), (*rgSr)[512] = calloc(8, sizeof(*rgSr));
        float endscore = 0.f;

        for (int d = 0; d < B; d++) {
            const int Id = inss[brk + d] - r0, Dd = d - Id;
            const int o1 = d >= 1 ? opI[brk + d - 1] : 0, o2 = d >= 2 ? opI[brk + d - 2] : 0;
            int sI[MAXN + 1], sD[MAXN + 1];
            for (int n = 1; n <= MAXN; n++) { sI[n] = d >= n ? (inss[brk + d] - inss[brk + d - n]) : 0; sD[n] = n - sI[n]; }
            for (int bc = 0; bc < W; bc++) {
                const int i = Id + r - bc, j = Dd - r + bc;
                float Mv = 0.f, Iv = 0.f, Dv = 0.f, Lv, Sv, Lb = 0.f, Sb = 0.f; int Mr = 0, Ir = 0, Dr = 0, Lrn = 0, Srn = 0;
                uint16_t rec = 0;
                if (i < 0 || j < 0 || i > imax || j > jmax) { /* OUT: zeros */ }
                else if (bc == 0 || bc == 2 * r) { Mv = Iv = Dv = (float)(100 * (d + 1)); }
                else {
                    const uint64_t cr = colrec[j]; const uint32_t rr = rowrec[i];
                    /* neighbours: uniform shifts */
                    const int tbc = bc + (o1 ? 0 : 1), lbc = bc - (o1 ? 1 : 0), dbc = bc + 1 - (o1 + o2);
                    /* LEN gather */
                    Lv = (float)(100 * d);
                    {
                        uint32_t mask = (uint32_t)((cr >> 48) & 0x3f) & (rr & 0x3f);
                        for (int n = max_n; n >= 1; n--) {
                            if (!((mask >> (n - 1)) & 1)) continue;
                            if (d < n) continue;
                            int sb = bc + sD[n]; if (sb > 2 * r - 1) continue;
                            int si = i - n, eq = 1;
                            for (int t = 0; t < n; t++) if (seq[si + t] != ref[j + t]) eq = 0;
                            if (!eq) continue;
                            int L = (int)((colrec[j + n] >> (8 * (n - 1))) & 0x7f);
                            int slot = (d - n) & 7; float base; int run0;
                            if ((rr >> (8 + n - 1)) & 1) { base = rgM[slot][sb]; run0 = 0; }
                            else { run0 = rgLr[slot][sb]; base = rgL[slot][sb]; if (run0 <= 0) continue; }
                            float cand = base + np_lookup(np, np_dim, clampv, n, L, L + run0 / n + 1);
                            if (cand < Lv) { Lv = cand; Lrn = run0 + n; Lb = base; }
                        }
                    }
                    /* SHR gather */
                    Sv = (float)(100 * d);
                    for (int n = max_n; n >= 1; n--) {
                        int byte = (int)((cr >> (8 * (n - 1))) & 0xff), L = byte & 0x7f;
                        if (!L || d < n) continue;
                        int sb = bc - sI[n]; if (sb < 1) continue;
                        int slot = (d - n) & 7; float base; int run0;
                        if (byte & 0x80) { base = rgM[slot][sb]; run0 = 0; }
                        else { run0 = rgSr[slot][sb]; base = rgS[slot][sb]; if (run0 <= 0) continue; }
                        float cand = base + np_lookup(np, np_dim, clampv, n, L, L - run0 / n - 1);
                        if (cand < Sv) { Sv = cand; Srn = run0 + n; Sb = base; }
                    }
                    /* INS / DEL */
                    if (i == 0) { Iv = (float)(100 * (j + 1)); Ir = j; }
                    else {
                        float v1 = Mv1[tbc] + gap_open, v2 = Iv1[tbc] + gap_ext;
                        if (v2 < v1) { Iv = v2; Ir = (i == 1) ? 1 : Ir1[tbc] + 1; } else { Iv = v1; Ir = 1; }
                    }
                    if (j == 0) { Dv = (float)(100 * (i + 1)); Dr = i; }
                    else {
                        float v1 = Mv1[lbc] + gap_open, v2 = Dv1[lbc] + gap_ext;
                        if (v2 < v1) { Dv = v2; Dr = (j == 1) ? 1 : Dr1[lbc] + 1; } else { Dv = v1; Dr = 1; }
                    }
                    /* MAT */
                    float best; int typ = T_MAT, run = 0;
                    if (i > 0 && j > 0) {
                        run = Mr2[dbc] + 1; if (run > 8191) run = 8191;
                        best = Mv2[dbc] + sub[((rr >> 16) & 7) * 5 + (int)((cr >> 56) & 7)];
                    } else best = Dv + 100.f;
                    if (Iv < best) { best = Iv; typ = T_INS; run = Ir; }
                    if (Lv < best) { best = Lv; typ = T_LEN; run = Lrn; }
                    if (Dv < best) { best = Dv; typ = T_DEL; run = Dr; }
                    if (Sv < best) { best = Sv; typ = T_SHR; run = Srn; }
                    Mv = best; Mr = (typ == T_MAT) ? run : 0;
                    rec = (uint16_t)(typ | (run << 3));           /* run < 8192 asserted by the model's test sizes */
                }
                nMv[bc] = Mv; nIv[bc] = Iv; nDv[bc] = Dv; nMr[bc] = Mr; nIr[bc] = Ir; nDr[bc] = Dr;
                rgM[d & 7][bc] = Mv; rgS[d & 7][bc] = Sb; rgL[d & 7][bc] = Lb; rgLr[d & 7][bc] = Lrn; rgSr[d & 7][bc] = Srn;
                tb[(size_t)d * W + bc] = rec;
                if (d == B - 1 && bc == r) endscore = Mv;
            }
            memcpy(Mv2, Mv1, W * 4); memcpy(Mr2, Mr1, W * 4);
            memcpy(Mv1, nMv, W * 4); memcpy(Iv1, nIv, W * 4); memcpy(Dv1, nDv, W * 4);
            memcpy(Mr1, nMr, W * 4); memcpy(Ir1, nIr, W * 4); memcpy(Dr1, nDr, W * 4);
        }
        if (scores && nsc < scores_cap) scores[nsc] = endscore;
        nsc++;

        /* traceback over the packed records */
        int i = imax, j = jmax, bad = 0; int64_t nrev = 0;
        while (i > 0 || j > 0) {
            if (i < 0) { bad = 1; break; }
            if (j < 0) { bad = 2; break; }
            int d = i + j, bc = (inss[brk + d] - r0) + r - i;
            uint16_t rec = (bc >= 0 && bc < W) ? tb[(size_t)d * W + bc] : 0;
            int typ = rec & 7, run = rec >> 3;
            if (run < 1) { bad = 3; break; }
            if (typ == T_INS || typ == T_LEN) { for (int t = 0; t < run; t++) rev[nrev++] = 'I'; i -= run; }
            else if (typ == T_DEL || typ == T_SHR) { for (int t = 0; t < run; t++) rev[nrev++] = 'D'; j -= run; }
            else if (typ == T_MAT) { for (int t = 0; t < run; t++) { i--; j--; rev[nrev++] = (ref[j] == seq[i]) ? '=' : 'X'; } }
            else { bad = 4; break; }
        }
        if (bad && !st) st = bad;
        if (out_len + nrev > out_cap) { out_len = -1; ci = nchunks; }
        else { for (int64_t k = 0; k < nrev; k++) out[out_len + k] = rev[nrev - 1 - k]; out_len += nrev; }

        free(rawr); free(raws); free(colrec); free(rowrec); free(tb);
        free(Mv1); free(Iv1); free(Dv1); free(Mv2); free(Mr1); free(Ir1); free(Dr1); free(Mr2);
        free(nMv); free(nIv); free(nDv); free(nMr); free(nIr); free(nDr);
        free(rgM); free(rgS); free(rgL); free(rgLr); free(rgSr);
    }
    free(opI); free(inss); free(rev);
    if (n_scores) *n_scores = nsc;
    if (status) *status = st;
    return out_len;
}

/* =====================================================================================================================
 * pm2_align -- the round-2 dataflow of npore_b200/csrc/forward.cuh ("INF ring" form), modelled with PHYSICAL ring slots
 * (slot = column mod NC, NC = 32*CPL >= W) so that slot aliasing is part of the model:
 *   - every slot of every anti-diagonal writes the ring: IN cells their values, all others (EDGE, cells outside the
 *     chunk, slots beyond the band) +INF for MAT.VAL and for both carried run-start values;
 *   - a cell whose SHR / LEN state was not set by any candidate stores run-start value +INF;
 *   - candidates are then evaluated WITHOUT the source checks of aln.pyx:609-612, 620-622, 645-647, 655-656 (source must be an
 *     interior cell of the chunk, run > 0, look-back inside the band): an invalid source yields cand = +INF, never < the
 *     state's initial 100*d;
 *   - exception ("risky" anti-diagonals): when NC - W < 4, a window of the last n <= max_n ops with >= NC-W+3 equal ops makes
 *     slot (j-n) mod NC alias a live cell of ANOTHER column; those anti-diagonals keep the explicit band check;
 *   - run/n by a 16-bit reciprocal on the run held in the upper half word, index clamped to 127; score tables re-laid as
 *     tabS[(n-1)*(max_l+1)+L][q] = np_score(n, L, -(q+1)), tabL[...][q] = np_score(n, L, q+1), q = 0..127.
 * form bit 0: 1 = never use the lean path (explicit checks everywhere; must give the same answers).
 * ===================================================================================================================== */
#include <math.h>
#define PM2_Q 128
static float *pm2_table(const float *np, int np_dim, int max_n, int max_l, int sign)
{
    const int rows = max_n * (max_l + 1), clampv = max_l - 1;
    float *t = (float *)malloc(sizeof(float) * (size_t)(rows + 1) * PM2_Q);
    for (int n = 1; n <= max_n; n++)
        for (int L = 0; L <= max_l; L++)
            for (int q = 0; q < PM2_Q; q++)
                t[((size_t)(n - 1) * (max_l + 1) + L) * PM2_Q + q] = np_lookup(np, np_dim, clampv, n, L, L + sign * (q + 1));
    for (int q = 0; q < PM2_Q; q++) t[(size_t)rows * PM2_Q + q] = INFINITY;
    return t;
}

int64_t pm2_align(const uint8_t *full_ref, int Lr, const uint8_t *full_seq, int Ls,
                  const char *cigar, int64_t cig_len, const float *sub, const float *np, int np_dim,
                  int max_n, int max_l, float gap_open, float gap_ext, int max_b_rows, int r, int form,
                  char *out, int64_t out_cap, float *scores, int scores_cap, int *n_scores, int *status, int64_t *n_risky)
{
    int64_t P = 0;
    uint8_t *opI = (uint8_t *)malloc((size_t)cig_len * 2 + 2);
    for (int64_t k = 0; k < cig_len; k++) {
        char c = cigar[k];
        if (c == 'I') opI[P++] = 1; else if (c == 'D') opI[P++] = 0; else { opI[P++] = 0; opI[P++] = 1; }
    }
    int32_t *inss = (int32_t *)calloc((size_t)P + 2, sizeof(int32_t));
    for (int64_t k = 0; k < P; k++) inss[k + 1] = inss[k] + opI[k];
    const int total = Ls + Lr;
    const int step = max_b_rows - 1;
    const int nchunks = total > 0 ? (total + step - 1) / step : 0;
    const int W = 2 * r + 1;
    const int NC = W <= 32 ? 32 : W <= 64 ? 64 : W <= 128 ? 128 : 256;
    const int spare = NC - W;
    float *tabS = pm2_table(np, np_dim, max_n, max_l, -1), *tabL = pm2_table(np, np_dim, max_n, max_l, +1);
    uint32_t M16[MAXN + 1];
    for (int n = 1; n <= MAXN; n++) M16[n] = (65536u + (uint32_t)n - 1u) / (uint32_t)n;
    int64_t out_len = 0, risky_cnt = 0; int st = 0, nsc = 0;
    char *rev = (char *)malloc((size_t)total + 8);

    for (int ci = 0; ci < nchunks; ci++) {
        int brk = ci * step, nxt = (ci + 1 < nchunks) ? (ci + 1) * step : total;
        if (ci > 0 && opI[brk] && !opI[brk - 1]) brk--;
        if (ci + 1 < nchunks && opI[nxt] && !opI[nxt - 1]) nxt--;
        const int B = nxt - brk + 1;
        const int r0 = inss[brk], c0 = brk - r0, r1 = inss[nxt], c1 = nxt - r1;
        const int imax = r1 - r0, jmax = c1 - c0;
        int rlen = (c1 + 1 < Lr ? c1 + 1 : Lr) - c0; if (rlen < 0) rlen = 0;
        int slen = (r1 + 1 < Ls ? r1 + 1 : Ls) - r0; if (slen < 0) slen = 0;
        const uint8_t *ref = full_ref + (c0 < Lr ? c0 : Lr), *seq = full_seq + (r0 < Ls ? r0 : Ls);
        uint8_t *rawr = (uint8_t *)malloc((size_t)(rlen + 1) * 8), *raws = (uint8_t *)malloc((size_t)(slen + 1) * 8);
        pm_np_raw(ref, rlen, max_n, max_l, rawr, NULL);
        pm_np_raw(seq, slen, max_n, max_l, raws, NULL);
        uint64_t *colrec = (uint64_t *)malloc(sizeof(uint64_t) * (size_t)(rlen + 8));
        uint32_t *rowrec = (uint32_t *)malloc(sizeof(uint32_t) * (size_t)(slen + 8));
        relay_col(ref, rlen, rawr, colrec);
        relay_row(seq, slen, raws, rowrec);

        uint16_t *tb = (uint16_t *)calloc((size_t)B * W, sizeof(uint16_t));
        float *Mv1 = calloc(W, 4), *Iv1 = calloc(W, 4), *Dv1 = calloc(W, 4), *Mv2 = calloc(W, 4);
        int *Mr1 = calloc(W, 4), *Ir1 = calloc(W, 4), *Dr1 = calloc(W, 4), *Mr2 = calloc(W, 4);
        float *nMv = calloc(W, 4), *nIv = calloc(W, 4), *nDv = calloc(W, 4);
        int *nMr = calloc(W, 4), *nIr = calloc(W, 4), *nDr = calloc(W, 4);
        /* physical rings [8][NC]: MAT value (INF unless IN), SHR / LEN run-start values, runs */
        float (*rgM)[256] = malloc(8 * sizeof(*rgM)), (*rgS)[256] = malloc(8 * sizeof(*rgS)), (*rgL)[256] = malloc(8 * sizeof(*rgL));
        int (*rgLr)[256] = calloc(8, sizeof(*rgLr)), (*rgSr)[256] = calloc(8, sizeof(*rgSr));
        for (int a = 0; a < 8; a++) for (int s = 0; s < 256; s++) rgM[a][s] = rgS[a][s] = rgL[a][s] = INFINITY;
        float endscore = 0.f;

        for (int d = 0; d < B; d++) {
            const int Id = inss[brk + d] - r0, Dd = d - Id;
            const int o1 = d >= 1 ? opI[brk + d - 1] : 0, o2 = d >= 2 ? opI[brk + d - 2] : 0;
            int sI[MAXN + 1], sD[MAXN + 1];
            for (int n = 1; n <= MAXN; n++) { sI[n] = d >= n ? (inss[brk + d] - inss[brk + d - n]) : 0; sD[n] = n - sI[n]; }
            /* risky: some window n <= max_n (fully inside the chunk or not -- the kernel tests the raw op history) holds
               >= spare+3 equal ops */
            int risky = (form & 1);
            for (int n = 1; n <= max_n; n++) {
                int nI = 0;
                for (int t = 1; t <= n; t++) nI += (brk + d - t >= 0) ? opI[brk + d - t] : 0;
                if (nI >= spare + 3 || n - nI >= spare + 3) risky = 1;
            }
            if (form & 2) risky = 0;        /* (test only: shows that the rule is needed) */
            risky_cnt += risky && !(form & 1);
            const int jlo = Dd - r;
            for (int pb = 0; pb < NC; pb++) {        /* pb = displayed b_col of the physical slot */
                const int bc = pb, j = jlo + bc, i = Id + r - bc;
                const int slot = ((j % NC) + NC) % NC;
                float Mv = 0.f, Iv = 0.f, Dv = 0.f, Lv, Sv, Lb = INFINITY, Sb = INFINITY; int Mr = 0, Ir = 0, Dr = 0, Lrn = 0, Srn = 0;
                uint16_t rec = 0;
                int in = 0;
                if (bc > 2 * r || i < 0 || j < 0 || i > imax || j > jmax) { }
                else if (bc == 0 || bc == 2 * r) { Mv = Iv = Dv = (float)(100 * (d + 1)); }
                else {
                    in = 1;
                    const uint64_t cr = colrec[j]; const uint32_t rr = rowrec[i];
                    const int tbc = bc + (o1 ? 0 : 1), lbc = bc - (o1 ? 1 : 0), dbc = bc + 1 - (o1 + o2);
                    Lv = (float)(100 * d);
                    {
                        uint32_t mask = (uint32_t)((cr >> 48) & 0x3f) & (rr & 0x3f);
                        for (int n = max_n; n >= 1; n--) {
                            if (!((mask >> (n - 1)) & 1)) continue;
                            if (risky) { if (d < n) continue; if (bc + sD[n] > 2 * r - 1) continue; }
                            int si = i - n, eq = 1;
                            for (int t = 0; t < n; t++) if (seq[si + t] != ref[j + t]) eq = 0;
                            if (!eq) continue;
                            int L = (int)((colrec[j + n] >> (8 * (n - 1))) & 0x7f);
                            int row = (d - n) & 7; float base; uint32_t xs;
                            if ((rr >> (8 + n - 1)) & 1) { base = rgM[row][slot]; xs = 0; }
                            else { xs = (uint32_t)rgLr[row][slot] << 16; base = rgL[row][slot]; }
                            uint32_t q = (uint32_t)(((uint64_t)xs * M16[n]) >> 32); if (q > PM2_Q - 1) q = PM2_Q - 1;
                            float cand = base + tabL[((size_t)(n - 1) * (max_l + 1) + L) * PM2_Q + q];
                            if (cand < Lv) { Lv = cand; Lrn = (int)(xs >> 16) + n; Lb = base; }
                        }
                    }
                    Sv = (float)(100 * d);
                    for (int n = max_n; n >= 1; n--) {
                        int byte = (int)((cr >> (8 * (n - 1))) & 0xff), L = byte & 0x7f;
                        if (!L) continue;
                        if (risky) { if (d < n) continue; if (bc - sI[n] < 1) continue; }
                        int row = (d - n) & 7, sslot = (((j - n) % NC) + NC) % NC; float base; uint32_t xs;
                        if (byte & 0x80) { base = rgM[row][sslot]; xs = 0; }
                        else { xs = (uint32_t)rgSr[row][sslot] << 16; base = rgS[row][sslot]; }
                        uint32_t q = (uint32_t)(((uint64_t)xs * M16[n]) >> 32); if (q > PM2_Q - 1) q = PM2_Q - 1;
                        float cand = base + tabS[((size_t)(n - 1) * (max_l + 1) + L) * PM2_Q + q];
                        if (cand < Sv) { Sv = cand; Srn = (int)(xs >> 16) + n; Sb = base; }
                    }
                    if (i == 0) { Iv = (float)(100 * (j + 1)); Ir = j; }
                    else {
                        float v1 = Mv1[tbc] + gap_open, v2 = Iv1[tbc] + gap_ext;
                        if (v2 < v1) { Iv = v2; Ir = (i == 1) ? 1 : Ir1[tbc] + 1; } else { Iv = v1; Ir = 1; }
                    }
                    if (j == 0) { Dv = (float)(100 * (i + 1)); Dr = i; }
                    else {
                        float v1 = Mv1[lbc] + gap_open, v2 = Dv1[lbc] + gap_ext;
                        if (v2 < v1) { Dv = v2; Dr = (j == 1) ? 1 : Dr1[lbc] + 1; } else { Dv = v1; Dr = 1; }
                    }
                    float best; int typ = T_MAT, run = 0;
                    if (i > 0 && j > 0) {
                        run = Mr2[dbc] + 1; if (run > 8191) run = 8191;
                        best = Mv2[dbc] + sub[((rr >> 16) & 7) * 5 + (int)((cr >> 56) & 7)];
                    } else best = Dv + 100.f;
                    if (Iv < best) { best = Iv; typ = T_INS; run = Ir; }
                    if (Lv < best) { best = Lv; typ = T_LEN; run = Lrn; }
                    if (Dv < best) { best = Dv; typ = T_DEL; run = Dr; }
                    if (Sv < best) { best = Sv; typ = T_SHR; run = Srn; }
                    Mv = best; Mr = (typ == T_MAT) ? run : 0;
                    rec = (uint16_t)(typ | (run << 3));
                }
                if (bc < W) {
                    nMv[bc] = Mv; nIv[bc] = Iv; nDv[bc] = Dv; nMr[bc] = Mr; nIr[bc] = Ir; nDr[bc] = Dr;
                    tb[(size_t)d * W + bc] = rec;
                    if (d == B - 1 && bc == r) endscore = Mv;
                }
                rgM[d & 7][slot] = in ? Mv : INFINITY; rgS[d & 7][slot] = in ? Sb : INFINITY; rgL[d & 7][slot] = in ? Lb : INFINITY;
                rgLr[d & 7][slot] = Lrn; rgSr[d & 7][slot] = Srn;
            }
            memcpy(Mv2, Mv1, W * 4); memcpy(Mr2, Mr1, W * 4);
            memcpy(Mv1, nMv, W * 4); memcpy(Iv1, nIv, W * 4); memcpy(Dv1, nDv, W * 4);
            memcpy(Mr1, nMr, W * 4); memcpy(Ir1, nIr, W * 4); memcpy(Dr1, nDr, W * 4);
        }
        if (scores && nsc < scores_cap) scores[nsc] = endscore;
        nsc++;

        int i = imax, j = jmax, bad = 0; int64_t nrev = 0;
        while (i > 0 || j > 0) {
            if (i < 0) { bad = 1; break; }
            if (j < 0) { bad = 2; break; }
            int d = i + j, bc = (inss[brk + d] - r0) + r - i;
            uint16_t rec = (bc >= 0 && bc < W) ? tb[(size_t)d * W + bc] : 0;
            int typ = rec & 7, run = rec >> 3;
            if (run < 1) { bad = 3; break; }
            if (typ == T_INS || typ == T_LEN) { for (int t = 0; t < run; t++) rev[nrev++] = 'I'; i -= run; }
            else if (typ == T_DEL || typ == T_SHR) { for (int t = 0; t < run; t++) rev[nrev++] = 'D'; j -= run; }
            else if (typ == T_MAT) { for (int t = 0; t < run; t++) { i--; j--; rev[nrev++] = (ref[j] == seq[i]) ? '=' : 'X'; } }
            else { bad = 4; break; }
        }
        if (bad && !st) st = bad;
        if (out_len + nrev > out_cap) { out_len = -1; ci = nchunks; }
        else { for (int64_t k = 0; k < nrev; k++) out[out_len + k] = rev[nrev - 1 - k]; out_len += nrev; }

        free(rawr); free(raws); free(colrec); free(rowrec); free(tb);
        free(Mv1); free(Iv1); free(Dv1); free(Mv2); free(Mr1); free(Ir1); free(Dr1); free(Mr2);
        free(nMv); free(nIv); free(nDv); free(nMr); free(nIr); free(nDr);
        free(rgM); free(rgS); free(rgL); free(rgLr); free(rgSr);
    }
    free(opI); free(inss); free(rev); free(tabS); free(tabL);
    if (n_scores) *n_scores = nsc;
    if (status) *status = st;
    if (n_risky) *n_risky = risky_cnt;
    return out_len;
}
