/* TEST INFRASTRUCTURE ONLY -- see nbldpc_oracle.h.
 *
 * Plain-C restatement of the reference decode path.  Every routine names the reference lines it
 * follows (/root/reference/...).  The order of floating-point operations is kept (including the
 * reference's add-then-subtract "undo" in its recursive configuration searches and the x87
 * long double accumulators of the QSPA box-plus), so results match the compiled reference
 * bit for bit on the same machine.  Compile with -ffp-contract=off and without -ffast-math.
 */
#include "nbldpc_oracle.h"
#include <float.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>

struct orc_code {
    int q, p, N, M, maxdv, maxdc;
    int *mul, *inv;
    int *vdeg, *cdeg;
    int *vlink, *vcoef, *vdc; /* [N][maxdv] */
    int *clink, *ccoef, *cdv; /* [M][maxdc] */
};

static double g_tems_leaves;
double orc_last_tems_leaves(void) { return g_tems_leaves; }

/* ---------------------------------------------------------------- GF(q) (GF.cpp:27-47) */
void orc_gf_tables(int q, int poly, int *mul, int *inv) {
    for (int a = 0; a < q; ++a)
        for (int b = 0; b < q; ++b) {
            int r = 0, x = a;
            for (int y = b; y; y >>= 1) {
                if (y & 1) r ^= x;
                x <<= 1;
                if (x & q) x ^= poly;
            }
            mul[a * q + b] = r;
        }
    inv[0] = 0;
    for (int a = 1; a < q; ++a)
        for (int b = 1; b < q; ++b)
            if (mul[a * q + b] == 1) inv[a] = b;
}
#define MUL(c, a, b) ((c)->mul[(a) * (c)->q + (b)])
#define INV(c, a) ((c)->inv[(a)])

/* ------------------------------------------------- graph (NBLDPC.cpp:147-263) */
orc_code *orc_code_create(int q, int poly, int N, int M, const int *var_deg, const int *chk_deg,
                          const int *var_link, const int *var_coef, int var_stride,
                          const int *chk_link, const int *chk_coef, int chk_stride) {
    orc_code *c = calloc(1, sizeof *c);
    c->q = q; c->N = N; c->M = M;
    for (c->p = 0; (1 << c->p) < q; ++c->p) ;
    c->mul = malloc(sizeof(int) * q * q);
    c->inv = malloc(sizeof(int) * q);
    orc_gf_tables(q, poly, c->mul, c->inv);
    for (int i = 0; i < N; ++i) if (var_deg[i] > c->maxdv) c->maxdv = var_deg[i];
    for (int i = 0; i < M; ++i) if (chk_deg[i] > c->maxdc) c->maxdc = chk_deg[i];
    c->vdeg = malloc(sizeof(int) * N); memcpy(c->vdeg, var_deg, sizeof(int) * N);
    c->cdeg = malloc(sizeof(int) * M); memcpy(c->cdeg, chk_deg, sizeof(int) * M);
    c->vlink = calloc(N * c->maxdv, sizeof(int)); c->vcoef = calloc(N * c->maxdv, sizeof(int));
    c->vdc = calloc(N * c->maxdv, sizeof(int));
    c->clink = calloc(M * c->maxdc, sizeof(int)); c->ccoef = calloc(M * c->maxdc, sizeof(int));
    c->cdv = calloc(M * c->maxdc, sizeof(int));
    for (int i = 0; i < N; ++i)
        for (int d = 0; d < var_deg[i]; ++d) {
            c->vlink[i * c->maxdv + d] = var_link[i * var_stride + d];
            c->vcoef[i * c->maxdv + d] = var_coef[i * var_stride + d];
        }
    for (int r = 0; r < M; ++r)
        for (int d = 0; d < chk_deg[r]; ++d) {
            c->clink[r * c->maxdc + d] = chk_link[r * chk_stride + d];
            c->ccoef[r * c->maxdc + d] = chk_coef[r * chk_stride + d];
        }
    /* cross indices: last match wins, as in the reference loops (NBLDPC.cpp:236-263) */
    for (int i = 0; i < N; ++i)
        for (int dv = 0; dv < c->vdeg[i]; ++dv) {
            int r = c->vlink[i * c->maxdv + dv];
            for (int dc = 0; dc < c->cdeg[r]; ++dc)
                if (c->clink[r * c->maxdc + dc] == i) c->vdc[i * c->maxdv + dv] = dc;
        }
    for (int r = 0; r < M; ++r)
        for (int dc = 0; dc < c->cdeg[r]; ++dc) {
            int i = c->clink[r * c->maxdc + dc];
            for (int dv = 0; dv < c->vdeg[i]; ++dv)
                if (c->vlink[i * c->maxdv + dv] == r) c->cdv[r * c->maxdc + dc] = dv;
        }
    return c;
}

void orc_code_destroy(orc_code *c) {
    if (!c) return;
    free(c->mul); free(c->inv); free(c->vdeg); free(c->cdeg);
    free(c->vlink); free(c->vcoef); free(c->vdc); free(c->clink); free(c->ccoef); free(c->cdv);
    free(c);
}

/* ------------------------------------------------- soft demapper (Comm.cpp:340-407) */
static int is_punct(const int *ps, int n, int s) {
    for (int i = 0; i < n; ++i) if (ps[i] == s) return 1;
    return 0;
}

void orc_demap_bpsk(int q, int N, const double *y_re, double sigma, const int *punct_sym, int n_punct, double *L_ch) {
    int p = 0; while ((1 << p) < q) ++p;
    double *lb = malloc(sizeof(double) * N * p);
    int pi = 0; /* y_re holds only the transmitted (un-punctured) bits, Comm.cpp:346-358 */
    for (int b = 0; b < N * p; ++b) {
        if (is_punct(punct_sym, n_punct, b / p)) { ++pi; lb[b] = 0; }
        else lb[b] = -2 * y_re[b - pi] / (sigma * sigma);
    }
    for (int s = 0; s < N; ++s)          /* Comm.cpp:360-373: bits added in ascending bit order */
        for (int k = 1; k < q; ++k) {
            double a = 0;
            for (int b = 0; b < p; ++b) if (k & (1 << b)) a += lb[s * p + b];
            L_ch[s * (q - 1) + k - 1] = a;
        }
    free(lb);
}

void orc_demap_qam(int q, int N, const double *y, const double *cs, double sigma,
                   const int *punct_sym, int n_punct, double *L_ch) {
    int pi = 0;
    for (int s = 0; s < N; ++s) {        /* Comm.cpp:378-398 */
        if (is_punct(punct_sym, n_punct, s)) {
            ++pi;
            for (int k = 1; k < q; ++k) L_ch[s * (q - 1) + k - 1] = 0;
        } else {
            double yr = y[2 * (s - pi)], yi = y[2 * (s - pi) + 1];
            for (int k = 1; k < q; ++k)
                L_ch[s * (q - 1) + k - 1] =
                    ((2 * yr - cs[0] - cs[2 * k]) * (cs[2 * k] - cs[0]) +
                     (2 * yi - cs[1] - cs[2 * k + 1]) * (cs[2 * k + 1] - cs[1])) / (2 * sigma * sigma);
        }
    }
}

/* ------------------------------------------------- shared helpers */
/* DecideLLRVector, NBLDPC.cpp:1542-1562: strict '>' scan from 0 => ties go to the lowest symbol */
static int decide(const double *L, int q) {
    double mx = 0; int a = 0;
    for (int k = 0; k < q - 1; ++k) if (L[k] > mx) { mx = L[k]; a = k + 1; }
    return mx <= 0 ? 0 : a;
}

typedef struct {
    const orc_code *c;
    double *v2c;  /* [N][maxdv][q-1] */
    double *c2v;  /* [M][maxdc][q-1] */
    double *tmp;  /* [q-1] */
} state;

#define V2C(s, col, d) ((s)->v2c + ((size_t)((col) * (s)->c->maxdv + (d))) * ((s)->c->q - 1))
#define C2V(s, row, d) ((s)->c2v + ((size_t)((row) * (s)->c->maxdc + (d))) * ((s)->c->q - 1))

/* steps A,B of every Decoding_* (NBLDPC.cpp:676-715 / 808-846 / 977-1026) */
static int posterior_and_syndrome(state *s, const double *L_ch, double *L_post, int *out) {
    const orc_code *c = s->c; int q1 = c->q - 1;
    for (int col = 0; col < c->N; ++col) {
        double *P = L_post + (size_t)col * q1;
        memcpy(P, L_ch + (size_t)col * q1, sizeof(double) * q1);
        for (int d = 0; d < c->vdeg[col]; ++d) {
            const double *m = C2V(s, c->vlink[col * c->maxdv + d], c->vdc[col * c->maxdv + d]);
            for (int k = 0; k < q1; ++k) P[k] = P[k] + m[k];
        }
        out[col] = decide(P, c->q);
    }
    for (int row = 0; row < c->M; ++row) {
        int syn = 0;
        for (int d = 0; d < c->cdeg[row]; ++d)
            syn ^= MUL(c, c->ccoef[row * c->maxdc + d], out[c->clink[row * c->maxdc + d]]);
        if (syn) return 0;
    }
    return 1;
}

/* step C (NBLDPC.cpp:718-744 / 848-857 / 1029-1052); w_old = 0 means "no damping" (EMS) */
static void var_to_check(state *s, const double *L_post, double w_old, double w_new) {
    const orc_code *c = s->c; int q1 = c->q - 1;
    for (int col = 0; col < c->N; ++col)
        for (int dv = 0; dv < c->vdeg[col]; ++dv) {
            double *v = V2C(s, col, dv);
            const double *m = C2V(s, c->vlink[col * c->maxdv + dv], c->vdc[col * c->maxdv + dv]);
            const double *P = L_post + (size_t)col * q1;
            int before = decide(v, c->q);
            memcpy(s->tmp, v, sizeof(double) * q1);
            for (int k = 0; k < q1; ++k) v[k] = P[k] - m[k];
            if (w_old != 0 && decide(v, c->q) != before)
                for (int k = 0; k < q1; ++k) v[k] = w_old * s->tmp[k] + w_new * v[k];
        }
}

static double soft_threshold(double x, double factor, double offset) {
    x = x / factor;                       /* NBLDPC.cpp:903-915, 1114-1126 */
    if (x < -1 * offset) return x + offset;
    if (x > offset) return x - offset;
    return 0;
}

/* ------------------------------------------------- QSPA check node (NBLDPC.cpp:1565-1712) */
typedef struct { state *s; double *sigma, *rho, *llrv; } bp_ctx;

static void box_plus(bp_ctx *b, double *L, const double *L1, const double *L2, int A1, int A2) {
    const orc_code *c = b->s->c; int q = c->q;
    double *out = b->llrv;
    if (A1 == 0) {
        for (int a = 1; a < q; ++a) out[a - 1] = L2[MUL(c, INV(c, A2), a) - 1];
    } else if (A2 == 0) {
        for (int a = 1; a < q; ++a) out[a - 1] = L1[MUL(c, INV(c, A1), a) - 1];
    } else {
        long double den = 0;              /* NBLDPC.cpp:1645: x87 extended accumulators */
        for (int x = 0; x < q; ++x) {
            int v1 = x, v2 = MUL(c, INV(c, A2), MUL(c, A1, x));
            if (v1 != 0 && v2 != 0) {
                double t = L1[v1 - 1] + L2[v2 - 1];
                if (den > t) den = den + logl(1 + expl(-1 * (den - t)));
                else den = t + logl(1 + expl(-1 * (t - den)));
            }
        }
        for (int a = 1; a < q; ++a) {
            long double num;
            int v1 = MUL(c, INV(c, A1), a), v2 = MUL(c, INV(c, A2), a);
            if (L1[v1 - 1] > L2[v2 - 1]) num = L1[v1 - 1] + log(1 + exp(-1 * (L1[v1 - 1] - L2[v2 - 1])));
            else num = L2[v2 - 1] + log(1 + exp(-1 * (L2[v2 - 1] - L1[v1 - 1])));
            for (int x = 0; x < q; ++x) {
                v1 = x; v2 = MUL(c, INV(c, A2), a ^ MUL(c, x, A1));
                if (v1 != 0 && v2 != 0) {
                    double t = L1[v1 - 1] + L2[v2 - 1];
                    if (num > t) num = num + logl(1 + expl(-1 * (num - t)));
                    else num = t + logl(1 + expl(-1 * (t - num)));
                }
            }
            out[a - 1] = (double)(num - den);
        }
    }
    memcpy(L, out, sizeof(double) * (q - 1));
}

static void chain_back(bp_ctx *b, int row, int l) {       /* L_Back, NBLDPC.cpp:1565-1591 */
    const orc_code *c = b->s->c;
    if (l < 0) return;
    if (l > 0) chain_back(b, row, l - 1);
    box_plus(b, b->sigma, b->sigma, V2C(b->s, c->clink[row * c->maxdc + l], c->cdv[row * c->maxdc + l]),
             l == 0 ? 0 : 1, c->ccoef[row * c->maxdc + l]);
}
static void chain_forward(bp_ctx *b, int row, int l) {    /* L_Forward, NBLDPC.cpp:1593-1619 */
    const orc_code *c = b->s->c;
    if (l >= c->cdeg[row]) return;
    if (l < c->cdeg[row] - 1) chain_forward(b, row, l + 1);
    box_plus(b, b->rho, b->rho, V2C(b->s, c->clink[row * c->maxdc + l], c->cdv[row * c->maxdc + l]),
             l == c->cdeg[row] - 1 ? 0 : 1, c->ccoef[row * c->maxdc + l]);
}

static void cn_qspa(state *s, bp_ctx *b) {                /* NBLDPC.cpp:747-767 */
    const orc_code *c = s->c; int q1 = c->q - 1;
    for (int row = 0; row < c->M; ++row)
        for (int d = 0; d < c->cdeg[row]; ++d) {
            memset(b->sigma, 0, sizeof(double) * q1);
            memset(b->rho, 0, sizeof(double) * q1);
            chain_back(b, row, d - 1);
            chain_forward(b, row, d + 1);
            int A1 = INV(c, c->ccoef[row * c->maxdc + d]), A2 = A1;
            if (d == 0) A1 = 0; else if (d == c->cdeg[row] - 1) A2 = 0;
            box_plus(b, C2V(s, row, d), b->sigma, b->rho, A1, A2);
        }
}

/* ------------------------------------------------- EMS check node (NBLDPC.cpp:859-918, 1715-1786) */
typedef struct {
    state *s; int row, nm, nc, except, end;
    double **srt; int **sym; double *best;
    int sum_sym, diff; double sum_llr;
} ems_ctx;

/* SortLLRVector: insertion from the tail, '>=' moves the newcomer ahead of equals */
static void ems_sort(double *Ls, int *Es, const double *L, int q1) {
    Ls[0] = 0; Es[0] = 0;
    for (int n = 0; n < q1; ++n) {
        double cand = L[n];
        Ls[n + 1] = cand; Es[n + 1] = n + 1;
        for (int i = n; i >= 0; --i) {
            if (cand >= Ls[i]) { Ls[i + 1] = Ls[i]; Ls[i] = cand; Es[i + 1] = Es[i]; Es[i] = n + 1; }
            else break;
        }
    }
}

/* Analysis switch (NOT the reference's arithmetic): with it, a configuration's sum is formed afresh, left to right, instead of in
   the reference's one running sum that is added to on the way down and subtracted from on the way back.  The running sum keeps a
   rounding residue that later configurations inherit; where two sums are equal in exact arithmetic that residue decides between
   them.  tools/fuzz_parity.py uses the switch to show that a rare EMS difference between the CUDA path (fresh sums) and the
   reference is exactly this residue and nothing else. */
static int g_ems_fresh = 0;
void orc_set_ems_fresh_sums(int on) { g_ems_fresh = on; }

/* ConstructConf: depth-first over the inputs, running sums updated and then undone in place */
static void ems_conf(ems_ctx *e, int lim, int begin) {
    const orc_code *c = e->s->c;
    if (begin > e->end) {
        if (e->sum_llr > e->best[e->sum_sym]) e->best[e->sum_sym] = e->sum_llr;
        return;
    }
    if (begin == e->except) { ems_conf(e, lim, begin + 1); return; }
    int h = c->ccoef[e->row * c->maxdc + begin];
    for (int k = 0; k < lim; ++k) {
        int g = MUL(c, e->sym[begin][k], h);
        const double before = e->sum_llr;
        e->sum_sym ^= g;
        e->sum_llr = e->sum_llr + e->srt[begin][k];
        e->diff += (k != 0);
        int go = e->diff <= e->nc;
        if (go) ems_conf(e, lim, begin + 1);
        e->sum_sym ^= g;
        if (g_ems_fresh) e->sum_llr = before; else
        e->sum_llr = e->sum_llr - e->srt[begin][k];
        e->diff -= (k != 0);
        if (!go) break;
    }
}

static void cn_ems(state *s, ems_ctx *e, double factor, double offset) {
    const orc_code *c = s->c; int q = c->q;
    for (int row = 0; row < c->M; ++row) {
        int dcn = c->cdeg[row];
        for (int d = 0; d < dcn; ++d)
            ems_sort(e->srt[d], e->sym[d], V2C(s, c->clink[row * c->maxdc + d], c->cdv[row * c->maxdc + d]), q - 1);
        for (int d = 0; d < dcn; ++d) {
            for (int k = 0; k < q; ++k) e->best[k] = -DBL_MAX;
            e->row = row; e->except = d; e->end = dcn - 1;
            int keep_nc = e->nc;
            e->sum_sym = 0; e->sum_llr = 0; e->diff = 0; e->nc = 1;
            ems_conf(e, q, 0);                                   /* conf(q,1)   NBLDPC.cpp:893-894 */
            e->sum_sym = 0; e->sum_llr = 0; e->diff = 0; e->nc = keep_nc;
            ems_conf(e, e->nm, 0);                               /* conf(nm,nc) NBLDPC.cpp:896-897 */
            double *o = C2V(s, row, d);
            int h = c->ccoef[row * c->maxdc + d];
            for (int k = 1; k < q; ++k)
                o[k - 1] = soft_threshold(e->best[MUL(c, k, h)] - e->best[0], factor, offset);
        }
    }
}

/* ------------------------------------------------- T-EMS check node (NBLDPC.cpp:1055-1130, 1789-1944) */
typedef struct {
    state *s; int nr, nc, end;
    double **dU; int *beta; int syn;
    int **ord; unsigned char **allowed;   /* TEMS_Min, TEMS_inConf_Nr: [q][dc] */
    double *dW; int **eta; int *cand; unsigned char *used;
    double *T; unsigned char *hit;
    int sum_sym, diff; double sum;
} tems_ctx;

static void tems_conf(tems_ctx *t, int begin) {            /* TEMS_ConstructConf, NBLDPC.cpp:1892-1944 */
    int q = t->s->c->q;
    if (begin > t->end) {
        g_tems_leaves += 1;
        if (t->sum < t->dW[t->sum_sym]) {
            t->dW[t->sum_sym] = t->sum;
            for (int d = 0; d <= t->end; ++d) t->eta[t->sum_sym][d] = t->cand[d];
        }
        return;
    }
    for (int e = 0; e < q; ++e) {
        if (!t->allowed[e][begin] || t->used[e]) continue;
        t->diff += (e != 0);
        if (t->diff <= t->nc) {
            t->used[e] = (e != 0);
            t->sum_sym ^= e;
            t->sum += t->dU[begin][e];
            t->cand[begin] = e;
            tems_conf(t, begin + 1);
            t->sum_sym ^= e;
            t->sum -= t->dU[begin][e];
            t->used[e] = 0;
            t->diff -= (e != 0);
        } else {
            t->diff -= (e != 0);
            break;
        }
    }
}

static void cn_tems(state *s, tems_ctx *t, double factor, double offset) {
    const orc_code *c = s->c; int q = c->q;
    for (int row = 0; row < c->M; ++row) {
        int dcn = c->cdeg[row];
        /* TEMS_Get_Beta (1789-1812) */
        t->syn = 0;
        for (int d = 0; d < dcn; ++d) {
            const double *v = V2C(s, c->clink[row * c->maxdc + d], c->cdv[row * c->maxdc + d]);
            int h = c->ccoef[row * c->maxdc + d];
            double mx = 0; int me = 0;
            for (int k = 1; k < q; ++k) if (v[k - 1] > mx) { mx = v[k - 1]; me = MUL(c, k, h); }
            t->beta[d] = me; t->syn ^= me;
        }
        /* TEMS_Get_deltaU (1814-1834) */
        for (int d = 0; d < dcn; ++d) {
            const double *v = V2C(s, c->clink[row * c->maxdc + d], c->cdv[row * c->maxdc + d]);
            int hi = INV(c, c->ccoef[row * c->maxdc + d]);
            int bp = MUL(c, hi, t->beta[d]);
            double mx = bp == 0 ? 0 : v[bp - 1];
            t->dU[d][t->beta[d]] = mx - 0;
            for (int x = 1; x < q; ++x) t->dU[d][x ^ t->beta[d]] = mx - v[MUL(c, hi, x) - 1];
        }
        /* TEMS_Get_Min (1836-1890): per eta, stable insertion order of the columns by dU */
        for (int e = 0; e < q; ++e) {
            for (int d = 0; d < dcn; ++d) t->ord[e][d] = d;
            for (int d = 1; d < dcn; ++d)
                for (int k = d; k >= 1; --k) {
                    if (t->dU[t->ord[e][k]][e] < t->dU[t->ord[e][k - 1]][e]) {
                        int tmp = t->ord[e][k]; t->ord[e][k] = t->ord[e][k - 1]; t->ord[e][k - 1] = tmp;
                    } else break;
                }
        }
        for (int d = 0; d < dcn; ++d) t->allowed[0][d] = 1;
        for (int e = 1; e < q; ++e) {
            for (int d = 0; d < dcn; ++d) t->allowed[e][d] = 0;
            for (int i = 0; i < t->nr && i < dcn; ++i) t->allowed[e][t->ord[e][i]] = 1;
        }
        /* configuration search (1064-1072) */
        for (int e = 0; e < q; ++e) { t->dW[e] = DBL_MAX; t->used[e] = 0; }
        t->sum_sym = 0; t->diff = 0; t->sum = 0.0; t->end = dcn - 1;
        tems_conf(t, 0);
        /* outputs (1075-1129) */
        for (int d = 0; d < dcn; ++d) {
            for (int e = 0; e < q; ++e) { t->T[e] = DBL_MAX; t->hit[e] = 0; }
            for (int e = 0; e < q; ++e) {
                int dev = t->eta[e][d], tgt = e ^ dev;
                double m = t->dW[e] - t->dU[d][dev];
                if (t->T[tgt] > m) { t->T[tgt] = m; t->hit[tgt] = 1; }
            }
            for (int e = 0; e < q; ++e)
                if (!t->hit[e]) t->T[e] = (d == t->ord[e][0]) ? t->dU[t->ord[e][1]][e] : t->dU[t->ord[e][0]][e];
            int hi = INV(c, c->ccoef[row * c->maxdc + d]);
            int bs = t->syn ^ t->beta[d];
            double L0 = -1.0 * t->T[bs];
            double *o = C2V(s, row, d);
            for (int e = 0; e < q; ++e)
                if (e != bs) {
                    int k = MUL(c, hi, e ^ bs);
                    o[k - 1] = soft_threshold(-1.0 * t->T[e] - L0, factor, offset);
                }
        }
    }
}


/* ------------------------------------------------- BS-TEMS check node (Decoding_BS_TEMS, NBLDPC.cpp:1145-1396) */
typedef struct {
    tems_ctx *t; int nm, nc;
    int *bq, *bcol; double *bL;           /* basicSet_q / basicSet_col / basicSet_L */
    unsigned char *col_sel;
    int *order; unsigned char *avail; int *R;
} bs_ctx;

/* Construct_BasicSet, NBLDPC.cpp:2023-2101.  std::sort there; any comparison sort gives the same order when no two
 * LLV tie (continuous LLRs); ties other than that are implementation-defined in the reference itself. */
static const double *g_sort_key;
static int cmp_order(const void *a, const void *b) {
    double x = g_sort_key[*(const int *)a], y = g_sort_key[*(const int *)b];
    if (x < y) return -1;
    if (x > y) return 1;
    return *(const int *)a - *(const int *)b;
}
static void bs_construct(bs_ctx *b, double *LLV) {
    tems_ctx *t = b->t; int q = t->s->c->q;
    for (int i = 0; i < q; ++i) { LLV[i] = t->dU[t->ord[i][0]][i]; b->order[i] = i; b->avail[i] = 1; }
    b->avail[0] = 0;
    g_sort_key = LLV;
    qsort(b->order, q, sizeof(int), cmp_order);
    int bit_len = 0; while ((1 << bit_len) < q) ++bit_len;
    if (b->nm > bit_len) {
        for (int i = 0; i < q - 1; ++i) { int e = b->order[i + 1]; b->bq[i] = e; b->bcol[i] = t->ord[e][0]; b->bL[i] = LLV[e]; }
    } else {
        int e0 = b->order[1];
        b->bq[0] = e0; b->bcol[0] = t->ord[e0][0]; b->bL[0] = LLV[e0];
        int rn = 0; b->R[rn++] = e0; b->avail[e0] = 0;
        int R_num = 1, n = 1;
        for (int i = 2; i < q; ++i) {
            if (R_num >= q - 1) break;
            int tq = b->order[i];
            if (b->avail[tq]) {
                b->bq[n] = tq; b->bcol[n] = t->ord[tq][0]; b->bL[n] = LLV[tq]; ++n;
                for (int j = 0; j < R_num; ++j) { int sq = tq ^ b->R[j]; b->R[rn++] = sq; b->avail[sq] = 0; }
                b->R[rn++] = tq; b->avail[tq] = 0;
                R_num = 2 * R_num + 1;
            }
        }
    }
}
/* BasicSet_TEMS_ConstructConf, NBLDPC.cpp:2102-2147: include-first recursion over the first nm basic elements */
static void bs_conf(bs_ctx *b, int begin, int end, int maxdc) {
    tems_ctx *t = b->t;
    if (begin > end) {
        if (t->sum < t->dW[t->sum_sym]) {
            t->dW[t->sum_sym] = t->sum;
            for (int d = 0; d < maxdc; ++d) t->eta[t->sum_sym][d] = t->cand[d];
        }
        return;
    }
    int col = b->bcol[begin];
    if (!b->col_sel[col]) {
        t->diff += 1;
        if (t->diff <= b->nc) {
            b->col_sel[col] = 1;
            t->sum_sym ^= b->bq[begin];
            t->sum += b->bL[begin];
            t->cand[col] = b->bq[begin];
            bs_conf(b, begin + 1, end, maxdc);
            t->sum_sym ^= b->bq[begin];
            t->sum -= b->bL[begin];
            b->col_sel[col] = 0;
            t->diff -= 1;
            t->cand[col] = 0;
            bs_conf(b, begin + 1, end, maxdc);
        } else {
            t->diff -= 1;
            t->cand[col] = 0;
            bs_conf(b, begin + 1, end, maxdc);
        }
    } else bs_conf(b, begin + 1, end, maxdc);
}

static void cn_bstems(state *s, bs_ctx *b, double *LLV, double factor, double offset) {
    tems_ctx *t = b->t;
    const orc_code *c = s->c; int q = c->q;
    for (int row = 0; row < c->M; ++row) {
        int dcn = c->cdeg[row];
        t->syn = 0;                                           /* TEMS_Get_Beta */
        for (int d = 0; d < dcn; ++d) {
            const double *v = V2C(s, c->clink[row * c->maxdc + d], c->cdv[row * c->maxdc + d]);
            int h = c->ccoef[row * c->maxdc + d];
            double mx = 0; int me = 0;
            for (int k = 1; k < q; ++k) if (v[k - 1] > mx) { mx = v[k - 1]; me = MUL(c, k, h); }
            t->beta[d] = me; t->syn ^= me;
        }
        for (int d = 0; d < dcn; ++d) {                       /* TEMS_Get_deltaU */
            const double *v = V2C(s, c->clink[row * c->maxdc + d], c->cdv[row * c->maxdc + d]);
            int hi = INV(c, c->ccoef[row * c->maxdc + d]);
            int bp = MUL(c, hi, t->beta[d]);
            double mx = bp == 0 ? 0 : v[bp - 1];
            t->dU[d][t->beta[d]] = mx - 0;
            for (int x = 1; x < q; ++x) t->dU[d][x ^ t->beta[d]] = mx - v[MUL(c, hi, x) - 1];
        }
        for (int e = 0; e < q; ++e) {                         /* TEMS_Get_Min: column order; TEMS_Nr is 0 in this mode */
            for (int d = 0; d < dcn; ++d) t->ord[e][d] = d;
            for (int d = 1; d < dcn; ++d)
                for (int k = d; k >= 1; --k) {
                    if (t->dU[t->ord[e][k]][e] < t->dU[t->ord[e][k - 1]][e]) {
                        int tmp = t->ord[e][k]; t->ord[e][k] = t->ord[e][k - 1]; t->ord[e][k - 1] = tmp;
                    } else break;
                }
        }
        for (int d = 0; d < dcn; ++d) t->allowed[0][d] = 1;
        for (int e = 1; e < q; ++e) for (int d = 0; d < dcn; ++d) t->allowed[e][d] = 0;
        for (int e = 0; e < q; ++e) { t->dW[e] = DBL_MAX; t->used[e] = 0; }
        for (int d = 0; d < c->maxdc; ++d) b->col_sel[d] = 0;
        t->sum_sym = 0; t->diff = 0; t->sum = 0.0; t->end = dcn - 1; t->nc = 1;
        tems_conf(t, 0);                                      /* TEMS_ConstructConf(1, 1, ...): only the all-zero configuration */
        bs_construct(b, LLV);
        t->sum_sym = 0; t->diff = 0; t->sum = 0.0;
        bs_conf(b, 0, b->nm - 1, c->maxdc);
        for (int d = 0; d < dcn; ++d) {                       /* outputs, NBLDPC.cpp:1330-1384 */
            for (int e = 0; e < q; ++e) { t->T[e] = DBL_MAX; t->hit[e] = 0; }
            for (int e = 0; e < q; ++e) {
                int dev = t->eta[e][d], tgt = e ^ dev;
                double m = t->dW[e] - t->dU[d][dev];
                if (t->T[tgt] > m) { t->T[tgt] = m; t->hit[tgt] = 1; }
            }
            for (int e = 0; e < q; ++e)
                if (!t->hit[e]) t->T[e] = (d == t->ord[e][0]) ? t->dU[t->ord[e][1]][e] : t->dU[t->ord[e][0]][e];
            int hi = INV(c, c->ccoef[row * c->maxdc + d]);
            int bs = t->syn ^ t->beta[d];
            double L0 = -1.0 * t->T[bs];
            double *o = C2V(s, row, d);
            for (int e = 0; e < q; ++e)
                if (e != bs) {
                    int k = MUL(c, hi, e ^ bs);
                    o[k - 1] = soft_threshold(-1.0 * t->T[e] - L0, factor, offset);
                }
        }
    }
}

/* ------------------------------------------------- the decoders */
static void *alloc2(int rows, int cols, size_t esz, void ***index) {
    char *blk = calloc((size_t)rows * cols, esz);
    void **ix = malloc(sizeof(void *) * rows);
    for (int r = 0; r < rows; ++r) ix[r] = blk + (size_t)r * cols * esz;
    *index = ix;
    return blk;
}

/* ------------------------------------------------- bit-level OSD post-processing (OSD.h; set-up NBLDPC.cpp:380-457) */
struct orc_osd {
    const orc_code *c;
    int nbit, mbit, kbit, crc_len, crc_rate, R;
    int *H0;                              /* [R][nbit]: CRC rows on top of the binary image of H (OSD.h:44-59) */
    int *H;                               /* [R][nbit] work copy */
    int *order;                           /* ReliableSeri_BIT */
    int *info, *tmp, *best;               /* infor_record_bit, OSD_temp_out_bit, near_optimal_bit */
    double *lbit;                         /* L_ch_bit */
    int min_distance;                     /* an int in the reference (NBLDPC.h:127): the best distance is stored TRUNCATED */
};

/* G_GaussEliminate_bit, OSD.h:257-313 (rows ^= instead of swaps; a dependent column is rotated to the end) */
static void gauss_g(int *Mx, int cols, int rows, int *order) {
    for (int row = 0; row < rows; ++row) {
        int col = order[row];
        if (Mx[row * cols + col] == 0) {
            int exchanged = 0;
            for (int up = row + 1; up < rows; ++up)
                if (Mx[up * cols + col] != 0) {
                    for (int i = 0; i < cols; ++i) Mx[row * cols + i] ^= Mx[up * cols + i];
                    exchanged = 1;
                    break;
                }
            if (!exchanged) {
                int flag = row;
                row--;
                for (int i = flag; i < cols - 1; ++i) { int t = order[i]; order[i] = order[i + 1]; order[i + 1] = t; }
            }
        }
        for (int up = row + 1; up < rows; ++up)
            if (up >= 0 && Mx[up * cols + col] != 0 && row >= 0)
                for (int i = 0; i < cols; ++i) Mx[up * cols + i] ^= Mx[row * cols + i];
    }
    for (int row = rows - 1; row > 0; --row)
        for (int up = row - 1; up >= 0; --up)
            if (Mx[up * cols + order[row]] == 1)
                for (int i = 0; i < cols; ++i) Mx[up * cols + i] ^= Mx[row * cols + i];
}

/* H_GaussEliminate_bit, OSD.h:326-384: pivots from the bottom row / the LEAST reliable position upwards; a column with
 * no pivot left is rotated to the front of the order */
static void gauss_h(int *Mx, int cols, int rows, int *order) {
    for (int row = rows - 1; row >= 0; --row) {
        int nt = row + cols - rows;
        int col = order[nt];
        if (Mx[row * cols + col] == 0) {
            int exchanged = 0;
            for (int up = row - 1; up >= 0; --up)
                if (Mx[up * cols + col] != 0) {
                    for (int i = 0; i < cols; ++i) Mx[row * cols + i] ^= Mx[up * cols + i];
                    exchanged = 1;
                    break;
                }
            if (!exchanged) {
                row++;
                for (int i = nt - 1; i >= 0; --i) { int t = order[i]; order[i] = order[i + 1]; order[i + 1] = t; }
            }
        }
        if (row < rows)                   /* (after row++ the column is zero in every row above: nothing to clear) */
            for (int up = row - 1; up >= 0; --up)
                if (Mx[up * cols + col] != 0)
                    for (int i = 0; i < cols; ++i) Mx[up * cols + i] ^= Mx[row * cols + i];
    }
    for (int i = 0; i < rows; ++i)
        for (int j = i + 1; j < rows; ++j) {
            int pos = order[cols - rows + i];
            if (Mx[j * cols + pos] == 1)
                for (int k = 0; k < cols; ++k) Mx[j * cols + k] ^= Mx[i * cols + k];
        }
}

orc_osd *orc_osd_create(const orc_code *c, int crc_len, int crc_rate, int fix_matrepr) {
    if (crc_len != 8 && crc_len != 16 && crc_len != 24) return NULL;      /* the reference hangs otherwise (OSD.h:472-499) */
    orc_osd *o = calloc(1, sizeof *o);
    int P = c->p;
    o->c = c; o->nbit = c->N * P; o->mbit = c->M * P; o->kbit = o->nbit - o->mbit;
    o->crc_len = crc_len; o->crc_rate = crc_rate; o->R = crc_rate + o->mbit;
    if (crc_rate < 0 || crc_rate > crc_len || o->kbit <= crc_len) { free(o); return NULL; }
    int nb = o->nbit, kb = o->kbit;
    o->H0 = calloc((size_t)o->R * nb, sizeof(int)); o->H = calloc((size_t)o->R * nb, sizeof(int));
    o->order = calloc(nb, sizeof(int)); o->info = calloc(nb, sizeof(int)); o->tmp = calloc(nb, sizeof(int));
    o->best = calloc(nb, sizeof(int)); o->lbit = calloc(nb, sizeof(double));
    /* H_origin_bit (NBLDPC.cpp:416-430): block (chk, var) = the matrix representation of the coefficient,
     * ValueMatric[l][k] = bit k of g * x^l.  The reference's Mat.Repr reader stops one record early (GF.cpp:137), so
     * the block of alpha^(q-2) = 1/alpha stays all-zero there. */
    int dead = fix_matrepr ? -1 : c->inv[2];
    for (int i = 0; i < c->N; ++i)
        for (int j = 0; j < c->vdeg[i]; ++j) {
            int chk = c->vlink[i * c->maxdv + j], g = c->vcoef[i * c->maxdv + j];
            for (int k = 0; k < P; ++k)
                for (int l = 0; l < P; ++l)
                    o->H0[(size_t)(crc_rate + P * chk + k) * nb + P * i + l] = g == dead ? 0 : (MUL(c, g, 1 << l) >> k) & 1;
        }
    /* CRCMatrixGen (OSD.h:472-510) */
    int gr = kb - crc_len;
    int *G = calloc((size_t)gr * kb, sizeof(int)), *seri = calloc(kb, sizeof(int));
    static const int t8[] = {0, 1, 4, 5, 7, 8}, t16[] = {0, 4, 11, 16}, t24[] = {0, 1, 18, 19, 23, 24};
    const int *taps = crc_len == 8 ? t8 : crc_len == 16 ? t16 : t24;
    int ntap = crc_len == 16 ? 4 : 6;
    for (int i = 0; i < gr; ++i) for (int t = 0; t < ntap; ++t) G[(size_t)i * kb + i + taps[t]] = 1;
    for (int i = 0; i < kb; ++i) seri[i] = i;
    gauss_g(G, kb, gr, seri);
    for (int i = 0; i < crc_rate; ++i) {
        o->H0[(size_t)i * nb + i + kb - crc_rate] = 1;
        for (int j = 0; j < gr; ++j) o->H0[(size_t)i * nb + j] = G[(size_t)j * kb + kb - crc_rate + i];
    }
    free(G); free(seri);
    return o;
}

void orc_osd_destroy(orc_osd *o) {
    if (!o) return;
    free(o->H0); free(o->H); free(o->order); free(o->info); free(o->tmp); free(o->best); free(o->lbit); free(o);
}

static void osd_encode(orc_osd *o) {                       /* OSD_Encode_bit, OSD.h:386-401 */
    int nb = o->nbit, dep = nb - o->R;
    memcpy(o->tmp, o->info, sizeof(int) * nb);
    for (int i = 0; i < o->R; ++i) {
        int t = 0;
        for (int j = 0; j < dep; ++j) t ^= o->H[(size_t)i * nb + o->order[j]] * o->tmp[o->order[j]];
        o->tmp[o->order[dep + i]] = t;
    }
}
static void osd_distance(orc_osd *o) {                     /* compute_min_distance_bit, OSD.h:403-420 */
    /* The function recomputes the bit length as int(CodeLen * log(GFq) / log(2)) -- multiplication first -- which
     * rounds down to N p - 1 for some (N, q), e.g. 96 * log(64) / log(2) = 575.9999999999999: the last bit is then
     * neither counted in the distance nor copied into the best candidate. */
    int len = (int)(o->c->N * log((double)o->c->q) / log(2.0));
    double d = 0;
    for (int i = 0; i < len; ++i)
        if ((o->lbit[i] < 0 && o->tmp[i] == 1) || (o->lbit[i] > 0 && o->tmp[i] == 0)) d = d + fabs(o->lbit[i]);
    if (d < o->min_distance) { memcpy(o->best, o->tmp, sizeof(int) * len); o->min_distance = (int)d; }
}

typedef struct { double key; int idx; } osd_key;
static int osd_key_cmp(const void *a, const void *b) {      /* descending reliability; equal keys: the reference's
                                                               std::sort leaves their order unspecified, here by index */
    const osd_key *x = a, *y = b;
    if (x->key > y->key) return -1;
    if (x->key < y->key) return 1;
    return x->idx - y->idx;
}

/* Decoding_OSD_bit, OSD.h:5-199.  sym: in = the iterative decoder's decisions (ignored when flag != 0),
 * out = the OSD decision.  sym_sum = LLR_Sym_sum [N][q-1] (ignored when flag != 0: the channel LLRs are used). */
void orc_osd_run(orc_osd *o, int order, int flag, const double *L_ch, const double *sym_sum, int *sym) {
    const orc_code *c = o->c; int P = c->p, q1 = c->q - 1, nb = o->nbit, kb = o->kbit;
    const double *rel = flag ? L_ch : sym_sum;
    for (int i = 0; i < c->N; ++i) {
        int rec = flag ? decide(L_ch + (size_t)i * q1, c->q) : sym[i];
        for (int j = 0; j < P; ++j) {
            o->info[i * P + j] = (rec >> j) & 1;
            o->best[i * P + j] = (rec >> j) & 1;
            o->lbit[i * P + j] = L_ch[(size_t)i * q1 + (1 << j) - 1];
        }
    }
    o->min_distance = 1000000;
    osd_key *keys = malloc(sizeof(osd_key) * nb);           /* OSD_permute, OSD.h:202-231 */
    for (int j = 0; j < c->N; ++j)
        for (int k = 0; k < P; ++k) { keys[j * P + k].key = fabs(rel[(size_t)j * q1 + (1 << k) - 1]); keys[j * P + k].idx = j * P + k; }
    qsort(keys, nb, sizeof(osd_key), osd_key_cmp);
    for (int i = 0; i < nb; ++i) o->order[i] = keys[i].idx;
    free(keys);
    memcpy(o->H, o->H0, sizeof(int) * (size_t)o->R * nb);
    gauss_h(o->H, nb, o->R, o->order);
    osd_encode(o); osd_distance(o);
    if (order >= 1)
        for (int i = 0; i < kb; ++i) {
            o->info[o->order[i]] ^= 1;
            osd_encode(o); osd_distance(o);
            o->info[o->order[i]] ^= 1;
        }
    if (order >= 2)
        for (int i = 0; i < kb; ++i) {
            o->info[o->order[i]] ^= 1;
            for (int j = i + 1; j < kb; ++j) {
                o->info[o->order[j]] ^= 1;
                osd_encode(o); osd_distance(o);
                o->info[o->order[j]] ^= 1;
            }
            o->info[o->order[i]] ^= 1;
        }
    if (order >= 3)
        for (int i = 0; i < kb; ++i) {
            o->info[o->order[i]] ^= 1;
            for (int j = i + 1; j < kb; ++j) {
                o->info[o->order[j]] ^= 1;
                for (int k = j + 1; k < kb; ++k) {
                    o->info[o->order[k]] ^= 1;
                    osd_encode(o); osd_distance(o);
                    o->info[o->order[k]] ^= 1;
                }
                o->info[o->order[j]] ^= 1;
            }
            o->info[o->order[i]] ^= 1;
        }
    for (int i = 0; i < c->N; ++i) {
        int v = 0;
        for (int j = 0; j < P; ++j) v = 2 * v + o->best[i * P + P - 1 - j];
        sym[i] = v;
    }
}

/* The ordering the last orc_osd_run ended with (ReliableSeri_BIT after the elimination) and its best distance */
void orc_osd_last(const orc_osd *o, int *order, int *min_distance) {
    if (order) memcpy(order, o->order, sizeof(int) * o->nbit);
    if (min_distance) *min_distance = o->min_distance;
}

int orc_decode_osd(const orc_code *c, int method, int max_iter, int p0, int p1, double factor, double offset,
                   const double *L_ch, int *out, double *L_post, int *iters,
                   orc_osd *osd, int osd_order, double osd_alpha, int osd_flag) {
    int q = c->q, q1 = q - 1;
    state s; s.c = c;
    s.v2c = malloc(sizeof(double) * c->N * c->maxdv * q1);
    s.c2v = calloc((size_t)c->M * c->maxdc * q1, sizeof(double));
    s.tmp = malloc(sizeof(double) * q1);
    for (int col = 0; col < c->N; ++col)                 /* init, NBLDPC.cpp:649-667 / 781-800 / 934-969 */
        for (int d = 0; d < c->vdeg[col]; ++d) memcpy(V2C(&s, col, d), L_ch + (size_t)col * q1, sizeof(double) * q1);

    bp_ctx b = {&s, 0, 0, 0};
    ems_ctx e; memset(&e, 0, sizeof e);
    tems_ctx t; memset(&t, 0, sizeof t);
    void *blk[8] = {0}; int nblk = 0;
    if (method == 1) {
        b.sigma = calloc(q1, sizeof(double)); b.rho = calloc(q1, sizeof(double)); b.llrv = calloc(q1, sizeof(double));
    } else if (method == 2) {
        e.s = &s; e.nm = p0; e.nc = p1;
        blk[nblk++] = alloc2(c->maxdc, q, sizeof(double), (void ***)&e.srt);
        blk[nblk++] = alloc2(c->maxdc, q, sizeof(int), (void ***)&e.sym);
        e.best = malloc(sizeof(double) * q);
    } else {
        t.s = &s; t.nr = method == 7 ? 0 : p0; t.nc = p1;
        blk[nblk++] = alloc2(c->maxdc, q, sizeof(double), (void ***)&t.dU);
        blk[nblk++] = alloc2(q, c->maxdc, sizeof(int), (void ***)&t.ord);
        blk[nblk++] = alloc2(q, c->maxdc, 1, (void ***)&t.allowed);
        blk[nblk++] = alloc2(q, c->maxdc, sizeof(int), (void ***)&t.eta);
        t.beta = calloc(c->maxdc, sizeof(int)); t.cand = calloc(c->maxdc, sizeof(int));
        t.dW = malloc(sizeof(double) * q); t.T = malloc(sizeof(double) * q);
        t.used = calloc(q, 1); t.hit = calloc(q, 1);
        g_tems_leaves = 0;
    }
    bs_ctx bs; memset(&bs, 0, sizeof bs);
    double *LLV = 0;
    if (method == 7) {                                   /* BS-TEMS: p0 = basic_set_Nm, p1 = basic_set_Nc */
        bs.t = &t; bs.nm = p0; bs.nc = p1;
        bs.bq = calloc(q, sizeof(int)); bs.bcol = calloc(q, sizeof(int)); bs.bL = calloc(q, sizeof(double));
        bs.col_sel = calloc(c->maxdc, 1); bs.order = calloc(q, sizeof(int)); bs.avail = calloc(q, 1); bs.R = calloc(2 * q, sizeof(int));
        LLV = calloc(q, sizeof(double));
    }

    /* LLR_Sym_sum: every decoder accumulates alpha * sum + L_post at each tentative decision (NBLDPC.cpp:687, 820, 989, 1205) */
    double *sym_sum = osd ? calloc((size_t)c->N * q1, sizeof(double)) : NULL;
    int ok = 0, it = 0;
    while (it++ < max_iter) {
        ok = posterior_and_syndrome(&s, L_ch, L_post, out);
        if (sym_sum)
            for (size_t i = 0; i < (size_t)c->N * q1; ++i) sym_sum[i] = osd_alpha * sym_sum[i] + L_post[i];
        if (ok) break;
        if (method == 1) { var_to_check(&s, L_post, 0.5, 0.5); cn_qspa(&s, &b); }
        else if (method == 2) { var_to_check(&s, L_post, 0, 1); cn_ems(&s, &e, factor, offset); }
        else if (method == 7) { var_to_check(&s, L_post, 0.25, 0.75); cn_bstems(&s, &bs, LLV, factor, offset); }
        else { var_to_check(&s, L_post, 0.25, 0.75); cn_tems(&s, &t, factor, offset); }
    }
    *iters = ok ? it : max_iter;
    if (osd && osd_order >= 0 && !ok) orc_osd_run(osd, osd_order, osd_flag, L_ch, sym_sum, out);   /* NBLDPC.cpp:770, 921, 1135 */
    free(sym_sum);

    free(s.v2c); free(s.c2v); free(s.tmp);
    free(b.sigma); free(b.rho); free(b.llrv);
    free(e.srt); free(e.sym); free(e.best);
    free(t.dU); free(t.ord); free(t.allowed); free(t.eta); free(t.beta); free(t.cand);
    free(t.dW); free(t.T); free(t.used); free(t.hit);
    free(bs.bq); free(bs.bcol); free(bs.bL); free(bs.col_sel); free(bs.order); free(bs.avail); free(bs.R); free(LLV);
    for (int i = 0; i < nblk; ++i) free(blk[i]);
    return ok;
}

int orc_decode(const orc_code *c, int method, int max_iter, int p0, int p1, double factor, double offset,
               const double *L_ch, int *out, double *L_post, int *iters) {
    return orc_decode_osd(c, method, max_iter, p0, p1, factor, offset, L_ch, out, L_post, iters, NULL, -1, 1.0, 0);
}
