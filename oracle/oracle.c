/* oracle/oracle.c -- CPU restatement of the prrn_aln DP hot path.  TEST INFRASTRUCTURE ONLY
 * (see oracle.h for the rules).  Parity PINNED against the unmodified reference via tests/golden.
 *
 * Written from the algorithm, not from the reference's code shape: row-major (m, n) sweeps with an
 * explicit band test and rolling rows.  The reference scans anti-diagonals in place over diagonal
 * index r = n - m (fwd2d1.cc:136-160); both visit the same cells with the same arithmetic per cell,
 * so results are identical (every operation is a single IEEE add/sub/max in the VTYPE).
 */
#include "oracle.h"
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <float.h>
#include <limits.h>

#define ORC_MIN(a, b) ((a) < (b) ? (a) : (b))
#define ORC_MAX(a, b) ((a) > (b) ? (a) : (b))

/* cmn.h:103-104 */
#define ORC_NEG_INT (INT_MIN / 8 * 7)

void orc_stripe(const orc_seq *a, const orc_seq *b, int sh, orc_window *w)
{   /* aln2.cc:156-174 */
    if (sh < 0) {
        int shorter = ORC_MIN(a->right - a->left, b->right - b->left);
        sh = -sh * shorter / 100;
    }
    w->up = b->right - a->right;
    w->lw = b->left - a->left;
    if (w->up < w->lw) { int t = w->up; w->up = w->lw; w->lw = t; }
    w->up += sh;
    w->lw -= sh;
    int p;
    if ((p = b->right - a->left) < w->up) w->up = p;
    if ((p = b->left - a->right) > w->lw) w->lw = p;
    w->width = w->up - w->lw + 3;
}

int64_t orc_band_cells(const orc_seq *a, const orc_seq *b, int sh)
{   /* rows m in [a.left, a.right); columns max(m+lw, b.left) .. min(m+up+1, b.right)  (fwd2c.h:364-374) */
    orc_window w;
    orc_stripe(a, b, sh, &w);
    int64_t cells = 0;
    for (int m = a->left; m < a->right; ++m) {
        int n0 = ORC_MAX(m + w.lw, b->left), n9 = ORC_MIN(m + w.up + 1, b->right);
        if (n9 > n0) cells += n9 - n0;
    }
    return cells;
}

/* ---- score-only affine fill, instantiated for float and double VTYPE -------------------------- */
#define DEFINE_SCORE_D(VT, SUFFIX, NEVSEL_V)                                                        \
static double aln_score_d_##SUFFIX(const orc_seq *a, const orc_seq *b, const double *mtx, int dim,  \
                                   const orc_params *p)                                             \
{                                                                                                   \
    /* fwd2d1.cc:62-63: uu, vv in VTYPE from float alprm products */                                \
    const VT uu = (VT)((float)p->u * (float)p->scale);                                              \
    const VT vv = (VT)((float)p->v * (float)p->scale);                                              \
    const float tgapf = (float)p->tgapf;                                                            \
    orc_window w;                                                                                   \
    orc_stripe(a, b, p->sh, &w);                                                                    \
    const int lw = w.lw, up = w.up, W = w.width;                                                    \
    const int al = a->left, ar = a->right, bl = b->left, br = b->right;                             \
    const int r0 = bl - al;                                                                         \
    /* fin[r - lw + 1]: what the reference's in-place hh[r] holds when forwardD ends: the last      \
       cell computed on diagonal r, or the boundary value if the diagonal has no cell */            \
    VT *fin = (VT *)malloc(sizeof(VT) * (size_t)W);                                                 \
    VT *FIN = fin - lw + 1;                                                                         \
    /* boundary row/column in diagonal coordinates (fwd2d1.cc:67-87) */                             \
    for (int r = lw; r <= up; ++r) FIN[r] = 0;                                                      \
    if (!a->exgl) {                                                                                 \
        FIN[r0] = 0;                                                                                \
        float ltg = al ? 1.f : tgapf;                                                               \
        VT gp = (VT)(-vv * ltg), ge = (VT)(-uu * ltg);                                              \
        for (int r = r0 + 1; r <= up; ++r) FIN[r] = gp += ge;                                       \
    }                                                                                               \
    FIN[up + 1] = (VT)ORC_NEG_INT;                                                                  \
    if (!b->exgl) {                                                                                 \
        float ltg = bl ? 1.f : tgapf;                                                               \
        VT gp = (VT)(-vv * ltg), ge = (VT)(-uu * ltg);                                              \
        for (int r = r0 - 1; r >= lw; --r) FIN[r] = gp += ge;                                       \
    }                                                                                               \
    FIN[lw - 1] = (VT)ORC_NEG_INT;                                                                  \
    /* rolling rows over columns: Hp/Gp = row m-1, Hc/Gc = row m; slot j = n - bl + 1, slot 0 is  \
       the boundary column bl-1.  BND(r) = boundary value on diagonal r, sentinel outside the band */\
    const int NB = br - bl + 2;                                                                     \
    VT *buf = (VT *)malloc(sizeof(VT) * 4 * (size_t)NB);                                            \
    VT *Hp = buf, *Gp = buf + NB, *Hc = buf + 2 * NB, *Gc = buf + 3 * NB;                           \
    VT *bnd = (VT *)malloc(sizeof(VT) * (size_t)W);                                                 \
    memcpy(bnd, fin, sizeof(VT) * (size_t)W);                                                       \
    VT *BND = bnd - lw + 1;                                                                         \
    for (int n = bl - 1; n < br; ++n) {            /* boundary row al-1 */                          \
        int r = n - (al - 1);                                                                       \
        Hp[n - bl + 1] = (r >= lw && r <= up) ? BND[r] : (VT)ORC_NEG_INT;                           \
        Gp[n - bl + 1] = NEVSEL_V;                                                                  \
    }                                                                                               \
    for (int m = al; m < ar; ++m) {                                                                 \
        const int n0 = ORC_MAX(m + lw, bl), n9 = ORC_MIN(m + up + 1, br);                           \
        const double *srow = mtx + (size_t)a->res[m] * dim;                                         \
        {   /* boundary column cell H(m, bl-1) on diagonal bl-1-m */                                \
            int r = bl - 1 - m;                                                                     \
            Hc[0] = (r >= lw && r <= up) ? BND[r] : (VT)ORC_NEG_INT;                                \
        }                                                                                           \
        /* left neighbour of the first cell: boundary column if n0 == bl, else the hh[lw-1] sentinel */\
        VT hleft = (n0 == bl) ? Hc[0] : (VT)ORC_NEG_INT, fleft = NEVSEL_V;                          \
        for (int n = n0; n < n9; ++n) {                                                             \
            const int j = n - bl + 1;                                                               \
            /* cell above is out of band when r + 1 == up + 1: the hh[up+1] sentinel (fwd2d1.cc:76) */\
            const int inb = (n - m + 1 <= up);                                                      \
            VT habove = inb ? Hp[j] : (VT)ORC_NEG_INT;                                              \
            VT gabove = inb ? Gp[j] : NEVSEL_V;                                                     \
            VT f = ORC_MAX(hleft - vv, fleft) - uu;              /* fwd2d1.cc:147 */                \
            VT g = ORC_MAX(habove - vv, gabove) - uu;            /* fwd2d1.cc:148 */                \
            VT h = Hp[j - 1] + (VT)srow[b->res[n]];              /* fwd2d1.cc:149 */                \
            h = ORC_MAX(ORC_MAX(h, f), g);                       /* fwd2d1.cc:150 */                \
            Hc[j] = h; Gc[j] = g;                                                                   \
            hleft = h; fleft = f;                                                                   \
            FIN[n - m] = h;                                                                         \
        }                                                                                           \
        VT *t = Hp; Hp = Hc; Hc = t; t = Gp; Gp = Gc; Gc = t;                                       \
    }                                                                                               \
    free(bnd);                                                                                      \
    /* lastD (fwd2d1.cc:97-134): discounted / free trailing gaps */                                 \
    const int r9 = br - ar;                                                                         \
    float rtg = b->exgr ? 0.f : tgapf;                                                              \
    if (br == b->len && rtg < 1) {                                                                  \
        int dm = 0, rw = up + 1, rf = br - al;                                                      \
        if (rf < rw) rw = rf;                                                                       \
        for (int r = rw - 1; r >= r9; --r) {                                                        \
            ++dm;                                                                                   \
            VT gpn = dm == 1 ? vv + uu : uu;                                                        \
            FIN[r + 1] += (VT)(gpn * rtg);                                                          \
            if (FIN[r] < FIN[r + 1]) FIN[r] = FIN[r + 1]; else dm = 0;                              \
        }                                                                                           \
    }                                                                                               \
    rtg = a->exgr ? 0.f : tgapf;                                                                    \
    if (ar == a->len && rtg < 1) {                                                                  \
        int dn = 0, rw = lw, rf = bl - ar + 1;                                                      \
        if (rf > rw) rw = rf;                                                                       \
        for (int r = rw + 1; r <= r9; ++r) {                                                        \
            ++dn;                                                                                   \
            VT gpn = dn == 1 ? vv + uu : uu;                                                        \
            FIN[r - 1] += (VT)(gpn * rtg);                                                          \
            if (FIN[r] < FIN[r - 1]) FIN[r] = FIN[r - 1]; else dn = 0;                              \
        }                                                                                           \
    }                                                                                               \
    double res = (double)FIN[r9];                                                                   \
    free(buf); free(fin);                                                                           \
    return res;                                                                                     \
}

DEFINE_SCORE_D(float, f32, (-(FLT_MAX / 16 * 7)))
DEFINE_SCORE_D(double, f64, (-(DBL_MAX / 16 * 7)))

double orc_aln_score_d(const orc_seq *a, const orc_seq *b, const double *mtx, int dim,
                       const orc_params *p)
{
    return p->vtype ? aln_score_d_f64(a, b, mtx, dim, p) : aln_score_d_f32(a, b, mtx, dim, p);
}

double orc_self_score(const orc_seq *a, const double *mtx, int dim, const orc_params *p)
{   /* aln2.cc:54-64 */
    if (p->vtype) {
        double s = 0;
        for (int i = a->left; i < a->right; ++i) s += mtx[(size_t)a->res[i] * dim + a->res[i]];
        return s;
    }
    float s = 0;
    for (int i = a->left; i < a->right; ++i) s += (float)mtx[(size_t)a->res[i] * dim + a->res[i]];
    return s;
}

double orc_score2dist(double scr, int la, int lb, double self_a, double self_b, const orc_params *p)
{   /* phyl.cc:230 denome; aln2.cc:332-333; phyl.cc:249 */
    int dlen = abs(la - lb);
    if (p->vtype) {
        double denome = sqrt(self_a * self_b);
        double s = scr + (float)p->u * dlen / 2;
        double dst = 1. - s / denome;
        return 100. * dst;
    } else {
        float denome = sqrtf((float)self_a * (float)self_b);
        float s = (float)scr;
        s += (float)p->u * dlen / 2;
        float dst = (float)(1. - s / denome);
        return (float)(100. * dst);
    }
}

void orc_calcdist(const orc_seq *seqs, int nn, const double *mtx, int dim, const orc_params *p,
                  double *dist, double *raw_scores)
{   /* phyl.cc:318-342 (DynScr): selfscr per sequence, dpscore per pair in elem(i,j) order */
    double *self = (double *)malloc(sizeof(double) * (size_t)nn);
    for (int i = 0; i < nn; ++i) self[i] = orc_self_score(seqs + i, mtx, dim, p);
    for (int j = 1; j < nn; ++j)
        for (int i = 0; i < j; ++i) {
            size_t k = (size_t)j * (j - 1) / 2 + i;
            double scr = orc_aln_score_d(seqs + i, seqs + j, mtx, dim, p);
            if (raw_scores) raw_scores[k] = scr;
            dist[k] = orc_score2dist(scr, seqs[i].right - seqs[i].left, seqs[j].right - seqs[j].left,
                                     self[i], self[j], p);
        }
    free(self);
}
