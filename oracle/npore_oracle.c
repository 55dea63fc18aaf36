/*
 * npore_oracle.c -- CPU restatement of nPoRe's realignment hot path.  TEST INFRASTRUCTURE ONLY.
 *
 * This file is the parity checker for the CUDA path in npore_b200/csrc/.  It is never linked
 * into, imported by, or called from the product package; only tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline leg use it (through oracle/oracle.py).
 *
 * It restates, in plain C, the algorithm of the reference (TimD1/nPoRe, paths relative to
 * /root/reference/):
 *   npo_get_np_info        <- src/aln.pyx:179-251   get_np_info()
 *   np_score_              <- src/aln.pyx:257-274   np_score()  (called with max_l in the max_n slot)
 *   npo_plan / prefix sums <- src/aln.pyx:279-311 get_inss/get_dels, :344-358 get_breaks
 *   npo_align              <- src/aln.pyx:379-787   align()  (scatter / "push" form, like the reference)
 *   npo_push_indels_left   <- src/cig.pyx:102-159
 *   npo_push_inss_thru_dels<- src/cig.pyx:164-192
 *   npo_standardize        <- src/bam.pyx:65-78 (== :105-118)
 *   npo_collapse           <- src/cig.pyx:13-38
 * Parity pin: tests/test_oracle.py checks this file against the compiled, unmodified reference
 * (oracle/_ref, built by oracle/build_ref.py) on seeded fuzz cases, and against the committed
 * golden vectors in tests/golden/ (generated from the reference by tests/golden/make_golden.py),
 * which include the reference's own test/data/npore_realigned.sam.
 *
 * The GPU kernels use a different formulation (gather / "pull" form with carried run bases);
 * keeping the oracle in the reference's scatter form makes the two independent.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

enum { T_MAT = 0, T_INS = 1, T_LEN = 2, T_DEL = 3, T_SHR = 4, N_TYP = 5 };
#define INF_ 100

typedef struct { float val; int32_t typ; int32_t run; } cell_t;

/* ------------------------------------------------------------------ get_np_info */
/* out layout: [len][2][max_n] int32 (L plane then L_IDX plane per position), zero-initialised here. */
int npo_get_np_info(const uint8_t *s, int len, int max_n, int max_l, int32_t *out)
{
    memset(out, 0, (size_t)len * 2 * max_n * sizeof(int32_t));
#define NP_L(p, n)  out[((size_t)(p) * 2 + 0) * max_n + ((n) - 1)]
#define NP_X(p, n)  out[((size_t)(p) * 2 + 1) * max_n + ((n) - 1)]
    for (int p = 0; p < len; p++) {
        if (!s[p]) continue;                       /* 'N' never starts a tract */
        for (int n = 1; n <= max_n; n++) {
            int l = 0, q = p;
            while (q + n < len && s[q] == s[q + n]) {
                q++;
                if ((q - p) % n == 0) l++;
            }
            if (l) l++;
            if (l > 2) {
                int longest = 1;
                for (int n2 = 1; n2 < n; n2++)
                    if (l * n <= NP_L(p, n2) * n2) longest = 0;
                for (int k = 0; k < l; k++) {
                    int pos = p + k * n;
                    if (longest && l > NP_L(pos, n)) {
                        NP_L(pos, n) = l < max_l ? l : max_l;
                        NP_X(pos, n) = k;
                    }
                }
            }
        }
    }
    return 0;
}

/* ------------------------------------------------------------------ np_score */
static float np_score_(int n, int ref_l, int indel, const float *np, int np_dim, int clampv)
{
    if (ref_l <= 0) return 100.0f;
    if (ref_l + indel < 0) return 100.0f;
    if (n < 1 || n > clampv) return 100.0f;
    int call = ref_l + indel;
    if (ref_l > clampv - 1) ref_l = clampv - 1;
    if (call > clampv - 1) call = clampv - 1;
    return np[((size_t)(n - 1) * np_dim + ref_l) * np_dim + call];
}

/* ------------------------------------------------------------------ planning helpers */
/* ops: 'D'/'I' string produced from the input cigar (every X,=,M -> "DI"). Returns its length. */
static int64_t to_di(const char *cig, int64_t n, char *ops)
{
    int64_t p = 0;
    for (int64_t k = 0; k < n; k++) {
        char c = cig[k];
        if (c == 'X' || c == '=' || c == 'M') { ops[p++] = 'D'; ops[p++] = 'I'; }
        else ops[p++] = c;                         /* 'I' or 'D' (anything else is out of contract) */
    }
    return p;
}

/* Number of breakpoints for (Ls,Lr,max_b_rows); fills breaks (capacity >= return value) if non-NULL.
 * inss/dels are prefix counts over the DI string, length array_size. */
static int plan_breaks(int array_size, int chunk, const int32_t *inss, const int32_t *dels, int32_t *breaks)
{
    int a = array_size - 1, b = chunk - 1;
    int nb = 1 + (a + b - 1) / b;
    if (a <= 0) nb = 1;
    if (!breaks) return nb;
    for (int i = 0; i < nb - 1; i++) {
        breaks[i] = i * b;
        if (i > 0 && inss[breaks[i] + 1] == inss[breaks[i]] + 1 && dels[breaks[i]] == dels[breaks[i] - 1] + 1)
            breaks[i] -= 1;
    }
    breaks[nb - 1] = array_size - 1;
    return nb;
}

/* Exposed for host-logic tests: returns number of chunks and fills breaks[] (caller capacity cap). */
int npo_plan(const char *cig, int64_t cig_len, int Ls, int Lr, int max_b_rows, int32_t *breaks, int cap)
{
    char *ops = (char *)malloc((size_t)cig_len * 2 + 2);
    int64_t P = to_di(cig, cig_len, ops);
    int array_size = Ls + Lr + 1;
    int32_t *inss = (int32_t *)calloc((size_t)P + 2, sizeof(int32_t));
    int32_t *dels = (int32_t *)calloc((size_t)P + 2, sizeof(int32_t));
    for (int64_t k = 0; k < P; k++) {
        inss[k + 1] = inss[k] + (ops[k] == 'I');
        dels[k + 1] = dels[k] + (ops[k] == 'D');
    }
    int nb = plan_breaks(array_size, max_b_rows, inss, dels, NULL);
    int rc = -1;
    if (nb <= cap) { plan_breaks(array_size, max_b_rows, inss, dels, breaks); rc = nb; }
    free(ops); free(inss); free(dels);
    return rc;
}

/* ------------------------------------------------------------------ align */
typedef struct {
    const int32_t *inss, *dels; int r, brk;
} xf_t;
static inline int bcol_of(const xf_t *x, int ar, int ac) { return x->inss[ar + ac] - ar + x->r; }

static int slices_match(const uint8_t *a, int alen, const uint8_t *b, int blen)
{
    if (alen != blen) return 0;
    for (int i = 0; i < alen; i++) if (a[i] != b[i]) return 0;
    return 1;
}
static inline int clip(int lo, int hi, int len) /* length of python slice [lo:hi] on array of len */
{
    if (lo > len) lo = len;
    if (hi > len) hi = len;
    return hi > lo ? hi - lo : 0;
}

/*
 * Returns number of ops written to out (expanded CIGAR over {=,X,I,D}), or -1 on capacity/alloc error.
 * scores[k] = MAT value at the end cell of chunk k (the reference never returns it; see oracle/build_ref.py
 * for the patched reference copy that does).  status: 0 ok, 1 row<0, 2 col<0, 3 run<1, 4 unknown type
 * (first anomaly met during traceback; the CIGAR is then partial, as in aln.pyx:689-716,737-739).
 */
int64_t npo_align(const uint8_t *full_ref, int Lr, const uint8_t *full_seq, int Ls,
                  const char *cigar, int64_t cig_len,
                  const float *sub /*5x5*/, const float *np, int np_dim,
                  int max_n, int max_l, float gap_open, float gap_ext, int max_b_rows, int r,
                  char *out, int64_t out_cap, float *scores, int scores_cap, int *n_scores, int *status)
{
    char *ops = (char *)malloc((size_t)cig_len * 2 + 2);
    int64_t P = to_di(cigar, cig_len, ops);
    int array_size = Ls + Lr + 1;
    int32_t *inss = (int32_t *)calloc((size_t)P + 2, sizeof(int32_t));
    int32_t *dels = (int32_t *)calloc((size_t)P + 2, sizeof(int32_t));
    for (int64_t k = 0; k < P; k++) {
        inss[k + 1] = inss[k] + (ops[k] == 'I');
        dels[k + 1] = dels[k] + (ops[k] == 'D');
    }
    int nb = plan_breaks(array_size, max_b_rows, inss, dels, NULL);
    int32_t *breaks = (int32_t *)malloc(sizeof(int32_t) * (size_t)nb);
    plan_breaks(array_size, max_b_rows, inss, dels, breaks);

    const int W = 2 * r + 1;
    const int a_rows = Ls + 1, a_cols = Lr + 1;
    int64_t out_len = 0;
    int st = 0, nsc = 0;
    cell_t *M = NULL; size_t M_cap = 0;
    int32_t *info_ref = NULL, *info_seq = NULL; size_t ir_cap = 0, is_cap = 0;
    char *rev = (char *)malloc((size_t)array_size + 8);
    int32_t zeros[64] = {0};
    xf_t X = { inss, dels, r, 0 };

    for (int ci = 0; ci + 1 < nb; ci++) {
        const int brk = breaks[ci], nxt = breaks[ci + 1];
        const int B = nxt - brk + 1;
        const int r0 = inss[brk], c0 = dels[brk], r1 = inss[nxt], c1 = dels[nxt];
        size_t need = (size_t)N_TYP * B * W;
        if (need > M_cap) { free(M); M = (cell_t *)malloc(need * sizeof(cell_t)); M_cap = need; if (!M) { out_len = -1; break; } }
        memset(M, 0, need * sizeof(cell_t));
#define CELL(t, br, bc) M[((size_t)(t) * B + (br)) * W + (bc)]

        const int rlen = clip(c0, c1 + 1, Lr), slen = clip(r0, r1 + 1, Ls);
        const uint8_t *ref = full_ref + (c0 < Lr ? c0 : Lr), *seq = full_seq + (r0 < Ls ? r0 : Ls);
        size_t nr = (size_t)(rlen + 1) * 2 * max_n, ns = (size_t)(slen + 1) * 2 * max_n;
        if (nr > ir_cap) { free(info_ref); info_ref = (int32_t *)malloc(nr * sizeof(int32_t)); ir_cap = nr; }
        if (ns > is_cap) { free(info_seq); info_seq = (int32_t *)malloc(ns * sizeof(int32_t)); is_cap = ns; }
        npo_get_np_info(ref, rlen, max_n, max_l, info_ref);
        npo_get_np_info(seq, slen, max_n, max_l, info_seq);

        /* pass 1 (aln.pyx:465-478): LEN/SHR of in-chunk, non-edge cells start at INF*(local diagonal) */
        for (int br = 0; br < B; br++)
            for (int bc = 1; bc < 2 * r; bc++) {
                int ar = inss[br + brk] + r - bc, ac = dels[br + brk] - r + bc;
                if (ar < r0 || ac < c0 || ar > r1 || ac > c1) continue;
                float v = (float)(INF_ * (ar - r0 + ac - c0));
                CELL(T_LEN, br, bc).val = v; CELL(T_SHR, br, bc).val = v;
            }

        /* pass 2 (aln.pyx:481-667) */
        for (int br = 0; br < B; br++) {
            const int g = br + brk;
            for (int bc = 0; bc < W; bc++) {
                const int ar = inss[g] + r - bc, ac = dels[g] - r + bc;
                if (ar < r0 || ac < c0 || ar > r1 || ac > c1) continue;
                if (bc == 0 || bc == 2 * r) {
                    for (int t = 0; t < N_TYP; t++) { cell_t *c = &CELL(t, br, bc); c->val = (float)(INF_ * (br + 1)); c->typ = T_MAT; c->run = 0; }
                    continue;
                }
                const int ri = ac - c0 - 1, si = ar - r0 - 1;     /* last consumed ref/seq index in the slices */
                const int32_t *l, *lx, *ls, *lsx;
                if (ac >= a_cols - 1) { l = zeros; lx = zeros; }
                else { l = info_ref + (size_t)(ri + 1) * 2 * max_n; lx = l + max_n; }
                if (ar >= a_rows - 1) { ls = zeros; lsx = zeros; }
                else { ls = info_seq + (size_t)(si + 1) * 2 * max_n; lsx = ls + max_n; }

                cell_t *cI = &CELL(T_INS, br, bc), *cD = &CELL(T_DEL, br, bc), *cM = &CELL(T_MAT, br, bc);
                cell_t *cL = &CELL(T_LEN, br, bc), *cS = &CELL(T_SHR, br, bc);

                /* INS */
                if (ar == r0) { cI->val = (float)(INF_ * (ac - c0 + 1)); cI->typ = T_DEL; cI->run = ac - c0; }
                else {
                    int tr = br - 1, tc = bcol_of(&X, ar - 1, ac);
                    float v1 = CELL(T_MAT, tr, tc).val + gap_open;
                    cI->val = v1; cI->typ = T_INS; cI->run = 1;
                    float v2 = CELL(T_INS, tr, tc).val + gap_ext;
                    if (v2 < v1) { cI->val = v2; cI->run = (ar == r0 + 1) ? 1 : CELL(T_INS, tr, tc).run + 1; }
                }
                /* DEL */
                if (ac == c0) { cD->val = (float)(INF_ * (ar - r0 + 1)); cD->typ = T_INS; cD->run = ar - r0; }
                else {
                    int lr_ = br - 1, lc = bcol_of(&X, ar, ac - 1);
                    float v1 = CELL(T_MAT, lr_, lc).val + gap_open;
                    cD->val = v1; cD->typ = T_DEL; cD->run = 1;
                    float v2 = CELL(T_DEL, lr_, lc).val + gap_ext;
                    if (v2 < v1) { cD->val = v2; cD->run = (ac == c0 + 1) ? 1 : CELL(T_DEL, lr_, lc).run + 1; }
                }
                /* MAT */
                float best; int run = 0;
                if (ar > r0 && ac > c0) {
                    int dr = br - 2, dc = bcol_of(&X, ar - 1, ac - 1);
                    const cell_t *dg = &CELL(T_MAT, dr, dc);
                    run = (dg->typ == T_MAT) ? dg->run + 1 : 1;
                    best = dg->val + sub[seq[si] * 5 + ref[ri]];
                    cM->val = best; cM->typ = T_MAT; cM->run = run;
                } else best = cD->val + (float)INF_;
                for (int t = 1; t < N_TYP; t++) {
                    const cell_t *c = &CELL(t, br, bc);
                    if (c->val < best) { best = c->val; cM->val = c->val; cM->typ = t; cM->run = c->run; }
                }
                /* LEN: first-row override, then scatter to (ar+n, ac) */
                if (ar == r0) { cL->val = (float)(INF_ * (ac - c0)); cL->typ = T_DEL; cL->run = ac - c0; }
                for (int n = 1; n <= max_n; n++) {
                    if (l[n - 1] == 0 || ls[n - 1] == 0 || lx[n - 1] != 0) continue;
                    if (!slices_match(seq + (si + 1 < slen ? si + 1 : slen), clip(si + 1, si + 1 + n, slen),
                                      ref + (ri + 1 < rlen ? ri + 1 : rlen), clip(ri + 1, ri + 1 + n, rlen))) continue;
                    if (ar + n > r1) continue;
                    int tr = br + n, tc = bcol_of(&X, ar + n, ac);
                    if (tc <= 0) continue;
                    cell_t *tg = &CELL(T_LEN, tr, tc);
                    if (lsx[n - 1] == 0) {
                        float v = cM->val + np_score_(n, l[n - 1], 1, np, np_dim, max_l);
                        if (v < tg->val) { tg->val = v; tg->typ = T_LEN; tg->run = n; }
                    } else {
                        int rn = cL->run;
                        if (rn > 0 && ar - rn >= r0) {
                            int uc = bcol_of(&X, ar - rn, ac);
                            if (uc < 2 * r) {
                                float v = CELL(T_MAT, br - rn, uc).val + np_score_(n, l[n - 1], rn / n + 1, np, np_dim, max_l);
                                if (v < tg->val) { tg->val = v; tg->typ = T_LEN; tg->run = rn + n; }
                            }
                        }
                    }
                }
                /* SHR: first-col override, then scatter to (ar, ac+n) */
                if (ac == c0) { cS->val = (float)(INF_ * (ar - r0)); cS->typ = T_INS; cS->run = ar - r0; }
                for (int n = 1; n <= max_n; n++) {
                    if (l[n - 1] == 0) continue;
                    if (ac + n > c1) continue;
                    int tr = br + n, tc = bcol_of(&X, ar, ac + n);
                    if (tc >= 2 * r) continue;
                    cell_t *tg = &CELL(T_SHR, tr, tc);
                    if (lx[n - 1] == 0) {
                        float v = cM->val + np_score_(n, l[n - 1], -1, np, np_dim, max_l);
                        if (v < tg->val) { tg->val = v; tg->typ = T_SHR; tg->run = n; }
                    } else {
                        int rn = cS->run;
                        if (rn > 0 && ac - rn >= c0) {
                            int uc = bcol_of(&X, ar, ac - rn);
                            if (uc > 0) {
                                float v = CELL(T_MAT, br - rn, uc).val + np_score_(n, l[n - 1], (-rn) / n - 1, np, np_dim, max_l);
                                if (v < tg->val) { tg->val = v; tg->typ = T_SHR; tg->run = rn + n; }
                            }
                        }
                    }
                }
            }
        }

        /* traceback (aln.pyx:671-742) */
        int ar = r1, ac = c1;
        if (nsc < scores_cap && scores) scores[nsc] = CELL(T_MAT, B - 1, bcol_of(&X, ar, ac)).val;
        nsc++;
        int64_t nrev = 0; int bad = 0;
        while (ar > r0 || ac > c0) {
            if (ar < 0) { bad = 1; break; }
            if (ac < 0) { bad = 2; break; }
            int br = ar + ac - brk, bc = bcol_of(&X, ar, ac);
            int typ = 0, run = 0;
            if (br >= 0 && br < B && bc >= 0 && bc < W) { typ = CELL(T_MAT, br, bc).typ; run = CELL(T_MAT, br, bc).run; }
            if (run < 1) { bad = 3; break; }
            if (typ == T_LEN || typ == T_INS) { for (int i = 0; i < run; i++) rev[nrev++] = 'I'; ar -= run; }
            else if (typ == T_SHR || typ == T_DEL) { for (int i = 0; i < run; i++) rev[nrev++] = 'D'; ac -= run; }
            else if (typ == T_MAT) {
                for (int i = 0; i < run; i++) { ar--; ac--; rev[nrev++] = (ref[ac - c0] == seq[ar - r0]) ? '=' : 'X'; }
            } else { bad = 4; break; }
        }
        if (bad && !st) st = bad;
        if (out_len + nrev > out_cap) { out_len = -1; break; }
        for (int64_t k = 0; k < nrev; k++) out[out_len + k] = rev[nrev - 1 - k];
        out_len += nrev;
    }
    free(M); free(info_ref); free(info_seq); free(rev); free(ops); free(inss); free(dels); free(breaks);
    if (n_scores) *n_scores = nsc;
    if (status) *status = st;
    return out_len;
}

/* ------------------------------------------------------------------ CIGAR standardisation */
/* op codes follow the reference's pysam encoding (cfg.py:28-32): M=0 I=1 D=2 ==7 X=8. */
void npo_push_indels_left(uint8_t *cig, int64_t n, const uint8_t *seq, uint8_t push_op)
{
    int64_t sp = 0, cp = 0;
    uint8_t *tmp = NULL; int64_t tmp_cap = 0;
    while (cp < n) {
        uint8_t op = cig[cp];
        if (op != push_op) {
            cp++;
            if (op == 0 || op == 8 || op == 7) sp++;
            continue;
        }
        int64_t len = 1;
        while (cp + len < n && cig[cp + len] == push_op) len++;
        int64_t k = 0;
        while (cp - k > 0 && sp - k > 0 && seq[sp - k - 1] == seq[sp - k - 1 + len] &&
               (cig[cp - k - 1] == 7 || cig[cp - k - 1] == 0)) k++;
        if (k) {
            if (k > tmp_cap) { free(tmp); tmp = (uint8_t *)malloc((size_t)k); tmp_cap = k; }
            memcpy(tmp, cig + cp - k, (size_t)k);
            memmove(cig + cp - k, cig + cp, (size_t)len);
            memcpy(cig + cp - k + len, tmp, (size_t)k);
        }
        cp += len; sp += len;
    }
    free(tmp);
}

void npo_push_inss_thru_dels(uint8_t *cig, int64_t n)
{
    for (int64_t i = 0; i + 1 < n; i++) {
        if (cig[i] == 2 && cig[i + 1] == 1) {
            int64_t di = i - 1;
            while (di >= 0 && cig[di] == 2) di--;
            int64_t nd = i - di;
            int64_t ii = i + 1;
            while (ii < n && cig[ii] == 1) ii++;
            int64_t ni = ii - i - 1;
            for (int64_t j = 0; j < ni; j++) cig[di + 1 + j] = 1;
            for (int64_t j = 0; j < nd; j++) cig[di + 1 + ni + j] = 2;
        }
    }
}

/* bam.pyx:65-78: in = expanded CIGAR chars over {=,X,I,D,M}; out = expanded standardised chars over {M,I,D}.
 * The reference's `while True` body runs exactly once (old_cig aliases int_cig).  Returns output length. */
int64_t npo_standardize(const char *in, int64_t n, const uint8_t *ref, const uint8_t *seq, char *out)
{
    uint8_t *c = (uint8_t *)malloc((size_t)n + 1);
    for (int64_t k = 0; k < n; k++) c[k] = (in[k] == 'I') ? 1 : (in[k] == 'D') ? 2 : 0;
    npo_push_indels_left(c, n, ref, 2);
    npo_push_inss_thru_dels(c, n);
    npo_push_indels_left(c, n, seq, 1);
    npo_push_inss_thru_dels(c, n);
    int64_t m = 0;
    for (int64_t k = 0; k < n; k++) {
        if (c[k] == 1 && k + 1 < n && c[k + 1] == 2) { out[m++] = 'M'; k++; }   /* 'ID' -> 'M', non-overlapping */
        else out[m++] = "MID"[c[k]];
    }
    free(c);
    return m;
}

/* cig.pyx:13-38: run-length collapse to text. out capacity must be >= 11*groups+1. Returns strlen. */
int64_t npo_collapse(const char *ops, int64_t n, char *out)
{
    int64_t m = 0, k = 0;
    while (k < n) {
        int64_t j = k;
        while (j < n && ops[j] == ops[k]) j++;
        int64_t cnt = j - k; char buf[24]; int t = 0;
        while (cnt) { buf[t++] = (char)('0' + cnt % 10); cnt /= 10; }
        while (t) out[m++] = buf[--t];
        out[m++] = ops[k];
        k = j;
    }
    out[m] = 0;
    return m;
}
