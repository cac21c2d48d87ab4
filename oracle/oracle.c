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


/* ---- Smith-Waterman-Gotoh score: Fwd2d ctor + swgforwardD (fwd2d1.cc:57-90, 162-189) ----------- */
#define DEFINE_SCORE_SWG(VT, SUFFIX, NEVSEL_V)                                                      \
static double aln_score_swg_##SUFFIX(const orc_seq *a, const orc_seq *b, const double *mtx, int dim,\
                                     const orc_params *p)                                           \
{                                                                                                   \
    const VT uu = (VT)((float)p->u * (float)p->scale);                                              \
    const VT vv = (VT)((float)p->v * (float)p->scale);                                              \
    const float tgapf = (float)p->tgapf;                                                            \
    orc_window w;                                                                                   \
    orc_stripe(a, b, p->sh, &w);                                                                    \
    const int lw = w.lw, up = w.up, W = w.width;                                                    \
    const int al = a->left, ar = a->right, bl = b->left, br = b->right;                             \
    const int r0 = bl - al;                                                                         \
    VT *bnd = (VT *)malloc(sizeof(VT) * (size_t)W);                                                 \
    VT *BND = bnd - lw + 1;                                                                         \
    for (int r = lw; r <= up; ++r) BND[r] = 0;                                                      \
    if (!a->exgl) {                                                                                 \
        float ltg = al ? 1.f : tgapf;                                                               \
        VT gp = (VT)(-vv * ltg), ge = (VT)(-uu * ltg);                                              \
        for (int r = r0 + 1; r <= up; ++r) BND[r] = gp += ge;                                       \
    }                                                                                               \
    if (!b->exgl) {                                                                                 \
        float ltg = bl ? 1.f : tgapf;                                                               \
        VT gp = (VT)(-vv * ltg), ge = (VT)(-uu * ltg);                                              \
        for (int r = r0 - 1; r >= lw; --r) BND[r] = gp += ge;                                       \
    }                                                                                               \
    const int NB = br - bl + 2;                                                                     \
    VT *buf = (VT *)malloc(sizeof(VT) * 4 * (size_t)NB);                                            \
    VT *Hp = buf, *Gp = buf + NB, *Hc = buf + 2 * NB, *Gc = buf + 3 * NB;                           \
    for (int n = bl - 1; n < br; ++n) {                                                             \
        int r = n - (al - 1);                                                                       \
        Hp[n - bl + 1] = (r >= lw && r <= up) ? BND[r] : (VT)ORC_NEG_INT;                           \
        Gp[n - bl + 1] = NEVSEL_V;                                                                  \
    }                                                                                               \
    VT maxh = NEVSEL_V;                                                                             \
    const VT vzero = 0;                                                                             \
    for (int m = al; m < ar; ++m) {                                                                 \
        const int n0 = ORC_MAX(m + lw, bl), n9 = ORC_MIN(m + up + 1, br);                           \
        const double *srow = mtx + (size_t)a->res[m] * dim;                                         \
        { int r = bl - 1 - m; Hc[0] = (r >= lw && r <= up) ? BND[r] : (VT)ORC_NEG_INT; }            \
        VT hleft = (n0 == bl) ? Hc[0] : (VT)ORC_NEG_INT, fleft = NEVSEL_V;                          \
        for (int n = n0; n < n9; ++n) {                                                             \
            const int j = n - bl + 1;                                                               \
            const int inb = (n - m + 1 <= up);                                                      \
            VT habove = inb ? Hp[j] : (VT)ORC_NEG_INT;                                              \
            VT gabove = inb ? Gp[j] : NEVSEL_V;                                                     \
            VT f = ORC_MAX(hleft - vv, fleft) - uu;              /* fwd2d1.cc:176 */                \
            VT g = ORC_MAX(habove - vv, gabove) - uu;            /* :177 */                         \
            VT h = Hp[j - 1] + (VT)srow[b->res[n]];              /* :178 */                         \
            h = ORC_MAX(ORC_MAX(ORC_MAX(h, f), g), vzero);       /* :179 */                         \
            maxh = ORC_MAX(maxh, h);                             /* :180 */                         \
            Hc[j] = h; Gc[j] = g;                                                                   \
            hleft = h; fleft = f;                                                                   \
        }                                                                                           \
        VT *t = Hp; Hp = Hc; Hc = t; t = Gp; Gp = Gc; Gc = t;                                       \
    }                                                                                               \
    free(bnd); free(buf);                                                                           \
    return (double)maxh;                                                                            \
}

DEFINE_SCORE_SWG(float, f32, (-(FLT_MAX / 16 * 7)))
DEFINE_SCORE_SWG(double, f64, (-(DBL_MAX / 16 * 7)))

/* ---- semi-global score with end points: Fwd2d_vd ctor + forwardD(ends) + lastD(ends)
 *      (fwd2d1.cc:212-322).  Every value carries the diagonal on which its path left the boundary.
 *      Quirks kept: the origin record is vclear'ed, so its r is 0 and not r0 (:226); a's exgl is
 *      ignored while b's is honoured (:227,234); gg subtracts (uu + vv) in one step (:309). -------- */
#define DEFINE_SCORE_VD(VT, SUFFIX)                                                                 \
typedef struct { VT val; int r; } o_vd_##SUFFIX;                                                    \
static double aln_score_vd_##SUFFIX(const orc_seq *a, const orc_seq *b, const double *mtx, int dim, \
                                    const orc_params *p, int *ends)                                 \
{                                                                                                   \
    typedef o_vd_##SUFFIX VD;                                                                       \
    const VT uu = (VT)((float)p->u * (float)p->scale);                                              \
    const VT vv = (VT)((float)p->v * (float)p->scale);                                              \
    const float tgapf = (float)p->tgapf;                                                            \
    orc_window w;                                                                                   \
    orc_stripe(a, b, p->sh, &w);                                                                    \
    const int lw = w.lw, up = w.up, W = w.width;                                                    \
    const int al = a->left, ar = a->right, bl = b->left, br = b->right;                             \
    const int r0 = bl - al;                                                                         \
    const VD black = {(VT)ORC_NEG_INT, 0};                                                          \
    VD *fin = (VD *)malloc(sizeof(VD) * (size_t)W);                                                 \
    VD *FIN = fin - lw + 1;                                                                         \
    for (int r = lw - 1; r <= up + 1; ++r) FIN[r] = black;                                          \
    FIN[r0].val = 0; FIN[r0].r = 0;                                                                 \
    {                                                                                               \
        float ltg = al ? 1.f : tgapf;                                                               \
        VT gp = (VT)(-vv * ltg), ge = (VT)(-uu * ltg);                                              \
        for (int r = r0 + 1; r <= up; ++r) { FIN[r].val = gp += ge; FIN[r].r = r; }                 \
        ltg = bl ? 1.f : (b->exgl ? 0.f : tgapf);                                                   \
        gp = (VT)(-vv * ltg); ge = (VT)(-uu * ltg);                                                 \
        for (int r = r0 - 1; r >= lw; --r) { FIN[r].val = gp += ge; FIN[r].r = r; }                 \
    }                                                                                               \
    VD *bnd = (VD *)malloc(sizeof(VD) * (size_t)W);                                                 \
    memcpy(bnd, fin, sizeof(VD) * (size_t)W);                                                       \
    VD *BND = bnd - lw + 1;                                                                         \
    const int NB = br - bl + 2;                                                                     \
    VD *buf = (VD *)malloc(sizeof(VD) * 4 * (size_t)NB);                                            \
    VD *Hp = buf, *Gp = buf + NB, *Hc = buf + 2 * NB, *Gc = buf + 3 * NB;                           \
    for (int n = bl - 1; n < br; ++n) {                                                             \
        int r = n - (al - 1);                                                                       \
        Hp[n - bl + 1] = (r >= lw && r <= up) ? BND[r] : black;                                     \
        Gp[n - bl + 1] = black;                                                                     \
    }                                                                                               \
    for (int m = al; m < ar; ++m) {                                                                 \
        const int n0 = ORC_MAX(m + lw, bl), n9 = ORC_MIN(m + up + 1, br);                           \
        const double *srow = mtx + (size_t)a->res[m] * dim;                                         \
        { int r = bl - 1 - m; Hc[0] = (r >= lw && r <= up) ? BND[r] : black; }                      \
        VD hleft = (n0 == bl) ? Hc[0] : black, fleft = black;                                       \
        for (int n = n0; n < n9; ++n) {                                                             \
            const int j = n - bl + 1;                                                               \
            const int inb = (n - m + 1 <= up);                                                      \
            VD habove = inb ? Hp[j] : black, gabove = inb ? Gp[j] : black;                          \
            VD ng, eg, f, g, h;                                                                     \
            ng.val = hleft.val - vv - uu; ng.r = hleft.r;        /* fwd2d1.cc:305 */                \
            eg.val = fleft.val - uu; eg.r = fleft.r;             /* :306 */                         \
            f = ng.val > eg.val ? ng : eg;                       /* :307 */                         \
            ng = habove; ng.val -= (uu + vv);                    /* :308 */                         \
            eg = gabove; eg.val -= uu;                           /* :309 */                         \
            g = ng.val > eg.val ? ng : eg;                       /* :310 */                         \
            h = Hp[j - 1]; h.val += (VT)srow[b->res[n]];         /* :311 */                         \
            ng = f.val > g.val ? f : g;                          /* :312 */                         \
            h = h.val > ng.val ? h : ng;                         /* :313 */                         \
            Hc[j] = h; Gc[j] = g;                                                                   \
            hleft = h; fleft = f;                                                                   \
            FIN[n - m] = h;                                                                         \
        }                                                                                           \
        VD *t = Hp; Hp = Hc; Hc = t; t = Gp; Gp = Gc; Gc = t;                                       \
    }                                                                                               \
    free(bnd); free(buf);                                                                           \
    /* lastD(ends), fwd2d1.cc:271-312 */                                                            \
    const int r9 = br - ar;                                                                         \
    float rtg = b->exgr ? 0.f : tgapf;                                                              \
    int dm = 0, dn = 0;                                                                             \
    if (br == b->len && rtg < 1) {                                                                  \
        int rw = up + 1, rf = br - al;                                                              \
        if (rf < rw) rw = rf;                                                                       \
        for (int r = rw - 1; r >= r9; --r) {                                                        \
            ++dm;                                                                                   \
            VT gpn = dm == 1 ? vv + uu : uu;                                                        \
            FIN[r + 1].val += (VT)(gpn * rtg);                                                      \
            if (FIN[r].val < FIN[r + 1].val) FIN[r] = FIN[r + 1]; else dm = 0;                      \
        }                                                                                           \
    }                                                                                               \
    rtg = a->exgr ? 0.f : tgapf;                                                                    \
    if (ar == a->len && rtg < 1) {                                                                  \
        int rw = lw, rf = bl - ar + 1;                                                              \
        if (rf > rw) rw = rf;                                                                       \
        for (int r = rw + 1; r <= r9; ++r) {                                                        \
            ++dn;                                                                                   \
            VT gpn = dn == 1 ? vv + uu : uu;                                                        \
            FIN[r - 1].val += (VT)(gpn * rtg);                                                      \
            if (FIN[r].val < FIN[r - 1].val) FIN[r] = FIN[r - 1]; else dn = 0;                      \
        }                                                                                           \
    }                                                                                               \
    ends[0] = FIN[r9].r - r0;                                                                       \
    ends[1] = dn ? dn : -dm;                                                                        \
    double res = (double)FIN[r9].val;                                                               \
    free(fin);                                                                                      \
    return res;                                                                                     \
}

DEFINE_SCORE_VD(float, f32)
DEFINE_SCORE_VD(double, f64)

double orc_aln_score_d(const orc_seq *a, const orc_seq *b, const double *mtx, int dim,
                       const orc_params *p)
{
    return p->vtype ? aln_score_d_f64(a, b, mtx, dim, p) : aln_score_d_f32(a, b, mtx, dim, p);
}

double orc_aln_score_full(const orc_seq *a, const orc_seq *b, const double *mtx, int dim,
                          const orc_params *p, int *ends)
{   /* alnScoreD dispatch, fwd2d1.cc:324-337 */
    if (p->lcl & 16) return p->vtype ? aln_score_swg_f64(a, b, mtx, dim, p) : aln_score_swg_f32(a, b, mtx, dim, p);
    if (ends) return p->vtype ? aln_score_vd_f64(a, b, mtx, dim, p, ends) : aln_score_vd_f32(a, b, mtx, dim, p, ends);
    return orc_aln_score_d(a, b, mtx, dim, p);
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

/* alnscore2dist(), algmode.lcl branch (aln2.cc:296-320): semi-global score with end points, the
 * denominator from the self scores of the aligned sub-windows, dlen from the trimmed lengths. */
double orc_score2dist_lcl(const orc_seq *a0, const orc_seq *b0, const double *mtx, int dim,
                          const orc_params *p, double *raw, int *ends_out)
{
    orc_seq a = *a0, b = *b0;
    a.exgl = (p->lcl & 1) != 0; a.exgr = (p->lcl & 2) != 0;      /* exg_seq, seq.cc:858-863 */
    b.exgl = (p->lcl & 4) != 0; b.exgr = (p->lcl & 8) != 0;
    int ends[2] = {0, 0};
    double scr = orc_aln_score_full(&a, &b, mtx, dim, p, ends);
    if (raw) *raw = scr;
    if (ends_out) { ends_out[0] = ends[0]; ends_out[1] = ends[1]; }
    int al = a.left, bl = b.left, ar = a.right, br = b.right;
    if (ends[0] > 0) bl += ends[0]; else if (ends[0] < 0) al -= ends[0];
    if (ends[1] > 0) br -= ends[1]; else if (ends[1] < 0) ar += ends[1];
    orc_seq at = a, bt = b;
    at.left = al; at.right = ar; bt.left = bl; bt.right = br;
    const double sa = orc_self_score(&at, mtx, dim, p), sb = orc_self_score(&bt, mtx, dim, p);
    const int dlen = abs(ar - al - br + bl);
    if (p->vtype) {
        double denome = sqrt(sa * sb);
        double s = scr + (float)p->u * dlen / 2;
        return 100. * (1. - s / denome);
    } else {
        float denome = (float)sqrt((float)sa * (float)sb);
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
            if (p->lcl) {
                double scr;
                dist[k] = orc_score2dist_lcl(seqs + i, seqs + j, mtx, dim, p, &scr, 0);
                if (raw_scores) raw_scores[k] = scr;
                continue;
            }
            double scr = orc_aln_score_d(seqs + i, seqs + j, mtx, dim, p);
            if (raw_scores) raw_scores[k] = scr;
            dist[k] = orc_score2dist(scr, seqs[i].right - seqs[i].left, seqs[j].right - seqs[j].left,
                                     self[i], self[j], p);
        }
    free(self);
}

/* ================================================================================================
 * Pairwise alignment with path: Fwd2c<DPunit>::forwardB for two single sequences (NGP).
 * Rows are swept in (m, n) order with rolling rows of records {val, dir, ptr}; the record store is
 * a plain growable array (the reference's Vmf is a linked list of 7,680-record blocks).
 * ================================================================================================ */
enum { O_DEAD = 0, O_DIAG = 2, O_NEWD = 3, O_VERT = 4, O_HORI = 8, O_NEWV = 12, O_NEWH = 13 }; /* aln.h:47-52 */
static int o_isdiag(int d) { d &= 15; return d == 2 || d == 3; }                     /* aln.h:60-62 */
static int o_isvert(int d) { d &= 15; return d == 4 || d == 5 || d == 6 || d == 7 || d == 12; }
static int o_ishori(int d) { d &= 15; return d == 8 || d == 9 || d == 10 || d == 11 || d == 13; }

typedef struct { int32_t m, n; long p; } o_rec;
typedef struct { o_rec *v; long n, cap; } o_store;
static long o_add(o_store *s, int m, int n, long p)
{
    if (s->n == s->cap) { s->cap = s->cap ? 2 * s->cap : 1024; s->v = (o_rec *)realloc(s->v, sizeof(o_rec) * (size_t)s->cap); }
    s->v[s->n].m = m; s->v[s->n].n = n; s->v[s->n].p = p;
    return s->n++;
}

#define DEFINE_ALIGN_NGP(VT, SUFFIX, NEVSEL_V)                                                       \
typedef struct { VT val; int dir; long ptr; } o_unit_##SUFFIX;                                      \
static int align_ngp_##SUFFIX(const orc_seq *a, const orc_seq *b, const double *mtx, int dim,       \
                              const orc_params *p, double *score, orc_skl *out, int cap)            \
{                                                                                                   \
    typedef o_unit_##SUFFIX U;                                                                      \
    /* PwdB::PwdB (aln2.cc:97-117), PwdM::resetuab (maln2.cc:227-243), single sequences: Vab = scale */\
    const float fu = (float)p->u, fv = (float)p->v, fu1 = (float)p->u1, fsc = (float)p->scale;      \
    const VT Vab = (VT)(fsc * 1 * 1);                                                               \
    const VT BasicGOP = (VT)(-fv * Vab), BasicGEP = (VT)(-fu * Vab), LongGEP = (VT)(-fu1 * Vab);    \
    const VT diffu = LongGEP - BasicGEP;                                                            \
    const VT LongGOP = BasicGOP - diffu * p->k1;                                                    \
    const int Noll = p->ls < 2 ? 2 : (p->ls > 3 ? 3 : p->ls);                                       \
    const int codonk1 = p->ls == 3 ? p->k1 : (INT_MAX / 8 * 7); /* LARGEN */                        \
    /* Fwd2c ctor (fwd2c.h:85-86): FTYPE ratios; FTYPE == VT in both builds */                      \
    const VT u2divu1 = BasicGEP < 0 ? (VT)LongGEP / BasicGEP : 0;                                   \
    const VT v2divv1 = BasicGOP < 0 ? (VT)LongGOP / BasicGOP : 0;                                   \
    const VT Basic_GOP = (VT)(-fsc * fv);          /* maln2.cc:232; vgop(x) = Basic_GOP * x */      \
    const VT unp = (VT)(1 * 1 * -fu);              /* unp1 (maln.h:185) with cfq = efq = 1 */       \
    orc_window w;                                                                                   \
    orc_stripe(a, b, p->sh, &w);                                                                    \
    const int lw = w.lw, up = w.up;                                                                 \
    const int al = a->left, ar = a->right, bl = b->left, br = b->right;                             \
    const U black = {NEVSEL_V, 0, 0};                                                               \
    o_store st = {0, 0, 0};                                                                         \
    o_add(&st, 0, 0, 0);                           /* skip 0-th record (fwd2c.h:361) */             \
    /* rows over columns n in [bl-1, br): slot j = n - bl + 1.  Hp/Gp/G2p = row m-1 */              \
    const int NB = br - bl + 2;                                                                     \
    U *buf = (U *)malloc(sizeof(U) * 6 * (size_t)NB);                                               \
    U *Hp = buf, *Gp = buf + NB, *G2p = buf + 2 * NB, *Hc = buf + 3 * NB, *Gc = buf + 4 * NB, *G2c = buf + 5 * NB;\
    for (int j = 0; j < NB; ++j) Hp[j] = Gp[j] = G2p[j] = Hc[j] = Gc[j] = G2c[j] = black;            \
    /* initB (fwd2c.h:138-176): origin + boundary row */                                            \
    Hp[0].val = 0; Hp[0].dir = O_DIAG; Hp[0].ptr = o_add(&st, al, bl, 0);                            \
    {                                                                                               \
        int rr = br - al; if (up < rr) rr = up;                                                     \
        int r0 = bl - al;                                                                           \
        for (int r = r0 + 1, k = 1; r <= rr; ++r, ++k) {                                            \
            const U *src = &Hp[k - 1];                                                              \
            VT gnp = o_ishori(src->dir) ? (VT)0 : (VT)(Basic_GOP * 1);  /* gapopen(h-1,..,-1) */     \
            gnp = (k < codonk1) ? gnp + unp : (VT)(v2divv1 * gnp + u2divu1 * unp);                  \
            Hp[k].dir = o_isvert(src->dir) ? O_NEWH : O_HORI;                                       \
            Hp[k].val = src->val + gnp; Hp[k].ptr = src->ptr;                                       \
        }                                                                                           \
    }                                                                                               \
    /* boundary column values H(m, bl-1), m = al .. : computed incrementally below */               \
    U colprev = Hp[0];                                                                              \
    int colk = 0;                                                                                   \
    const int rr_col = (bl - ar > lw) ? bl - ar : lw;   /* lowest diagonal the column reaches */     \
    for (int m = al; m < ar; ++m) {                                                                 \
        const int n0 = ORC_MAX(m + lw, bl), n9 = ORC_MIN(m + up + 1, br);                           \
        const double *srow = mtx + (size_t)a->res[m] * dim;                                         \
        /* boundary column cell of this row: diagonal r = bl-1-m, exists while r >= rr_col */        \
        {                                                                                           \
            int r = bl - 1 - m;                                                                     \
            if (r >= rr_col) {                                                                      \
                ++colk;                                                                             \
                VT gnp = o_isvert(colprev.dir) ? (VT)0 : (VT)(Basic_GOP * 1);                       \
                gnp = (colk < codonk1) ? gnp + unp : (VT)(v2divv1 * gnp + u2divu1 * unp);           \
                U c; c.dir = o_ishori(colprev.dir) ? O_NEWV : O_VERT;                               \
                c.val = colprev.val + gnp; c.ptr = colprev.ptr;                                     \
                colprev = c; Hc[0] = c;                                                             \
            } else Hc[0] = black;                                                                   \
        }                                                                                           \
        U f1 = black, f2 = black;                                                                   \
        for (int n = n0; n < n9; ++n) {                                                             \
            const int j = n - bl + 1;                                                               \
            const int above_inband = (n - m + 1 <= up);  /* H[r+1]/G[r+1] sentinel otherwise */      \
            const U habove = above_inband ? Hp[j] : black, gabove = above_inband ? Gp[j] : black;   \
            const U g2above = above_inband ? G2p[j] : black;                                        \
            const U hleft = (n - 1 >= n0 || n - 1 == bl - 1) ? Hc[j - 1] : black;                   \
            /* diagonal (fwd2c.h:395-398) */                                                        \
            U h = Hp[j - 1];                                                                        \
            h.dir = o_isdiag(h.dir) ? O_DIAG : O_NEWD;                                              \
            h.val = h.val + ((VT)srow[b->res[n]] + (VT)0);                                          \
            U g = black, g2 = black;                                                                \
            const U *mx = &g;                                                                       \
            VT gop = 0, gnp;                                                                        \
            if (m > al) {                                                                           \
                /* vertical (fwd2c.h:401-409) */                                                    \
                gnp = o_isvert(gabove.dir) ? (VT)0 : Basic_GOP;                                     \
                gop = o_isvert(habove.dir) ? (VT)0 : Basic_GOP;                                     \
                if (!o_isvert(habove.dir) && habove.val + gop > gabove.val + gnp) {                 \
                    g.dir = o_ishori(habove.dir) ? O_NEWV : O_VERT; g.val = habove.val + gop; g.ptr = habove.ptr;\
                } else {                                                                            \
                    g.dir = o_ishori(gabove.dir) ? O_NEWV : O_VERT; g.val = gabove.val + gnp; g.ptr = gabove.ptr;\
                }                                                                                   \
                g.val += unp;                                                                       \
                if (Noll == 3) {                   /* vertical2 (fwd2c.h:411-420) */                \
                    gnp = (VT)(v2divv1 * (o_isvert(g2above.dir) ? (VT)0 : Basic_GOP));              \
                    gop = (VT)(v2divv1 * gop);                                                      \
                    if (!o_isvert(habove.dir) && habove.val + gop > g2above.val + gnp) {            \
                        g2.dir = o_ishori(habove.dir) ? O_NEWV : O_VERT; g2.val = habove.val + gop; g2.ptr = habove.ptr;\
                    } else {                                                                        \
                        g2.dir = o_ishori(g2above.dir) ? O_NEWV : O_VERT; g2.val = g2above.val + gnp; g2.ptr = g2above.ptr;\
                    }                                                                               \
                    g2.val += (VT)(u2divu1 * unp);                                                  \
                    if (g2.val > mx->val) mx = &g2;                                                 \
                }                                                                                   \
            } else {                                                                                \
                /* first row: the g rows keep whatever the buffer holds = black (fwd2c.cc:39) */    \
                g = black; g2 = black;                                                              \
            }                                                                                       \
            if (n > bl) {                                                                           \
                /* horizontal (fwd2c.h:422-431) */                                                  \
                gnp = o_ishori(f1.dir) ? (VT)0 : Basic_GOP;                                         \
                gop = o_ishori(hleft.dir) ? (VT)0 : Basic_GOP;                                      \
                if (!o_ishori(hleft.dir) && hleft.val + gop > f1.val + gnp) {                       \
                    U t; t.dir = o_isvert(hleft.dir) ? O_NEWH : O_HORI; t.val = hleft.val + gop; t.ptr = hleft.ptr; f1 = t;\
                } else {                                                                            \
                    f1.dir = o_isvert(f1.dir) ? O_NEWH : O_HORI; f1.val = f1.val + gnp;             \
                }                                                                                   \
                f1.val += unp;                                                                      \
                if (f1.val >= mx->val) mx = &f1;                                                    \
                if (Noll == 3) {                   /* horizontal2 (fwd2c.h:433-442) */              \
                    gnp = (VT)(v2divv1 * (o_ishori(f2.dir) ? (VT)0 : Basic_GOP));                   \
                    gop = (VT)(v2divv1 * gop);                                                      \
                    if (!o_ishori(hleft.dir) && hleft.val + gop > f2.val + gnp) {                   \
                        U t; t.dir = o_isvert(hleft.dir) ? O_NEWH : O_HORI; t.val = hleft.val + gop; t.ptr = hleft.ptr; f2 = t;\
                    } else {                                                                        \
                        f2.dir = o_isvert(f2.dir) ? O_NEWH : O_HORI; f2.val = f2.val + gnp;         \
                    }                                                                               \
                    f2.val += (VT)(u2divu1 * unp);                                                  \
                    if (f2.val >= mx->val) mx = &f2;                                                \
                }                                                                                   \
            }                                                                                       \
            if (mx->val > h.val) h = *mx;          /* fwd2c.h:453 */                                \
            if (h.dir == O_NEWD || h.dir == O_NEWV || h.dir == O_NEWH)                              \
                h.ptr = o_add(&st, m, n, h.ptr);   /* fwd2c.h:465-467 */                            \
            Hc[j] = h; Gc[j] = g; G2c[j] = g2;                                                      \
        }                                                                                           \
        U *t;                                                                                       \
        t = Hp; Hp = Hc; Hc = t; t = Gp; Gp = Gc; Gc = t; t = G2p; G2p = G2c; G2c = t;              \
    }                                                                                               \
    /* result cell H[b.right - a.right] = H(ar-1, br-1) (fwd2c.h:475-481) */                        \
    U last = Hp[br - 1 - bl + 1];                                                                   \
    long pp = o_add(&st, ar, br, last.ptr);                                                         \
    *score = (double)last.val;                                                                      \
    /* Vmf::traceback (vmf.cc:103-119) */                                                           \
    int cnt = 0, ok = 1;                                                                            \
    for (long q = pp;; q = st.v[q].p) {                                                             \
        if (cnt + 1 >= cap) { ok = 0; break; }                                                      \
        out[++cnt].m = st.v[q].m; out[cnt].n = st.v[q].n;                                           \
        if (!st.v[q].p) break;                                                                      \
    }                                                                                               \
    out[0].m = 0; out[0].n = cnt;                                                                   \
    free(buf); free(st.v);                                                                          \
    return ok ? cnt : -1;                                                                           \
}

DEFINE_ALIGN_NGP(float, f32, (-(FLT_MAX / 16 * 7)))
DEFINE_ALIGN_NGP(double, f64, (-(DBL_MAX / 16 * 7)))

int orc_align_ngp(const orc_seq *a, const orc_seq *b, const double *mtx, int dim, const orc_params *p,
                  double *score, orc_skl *out, int cap)
{
    return p->vtype ? align_ngp_f64(a, b, mtx, dim, p, score, out, cap)
                    : align_ngp_f32(a, b, mtx, dim, p, score, out, cap);
}

static int o_sklcmp(const void *x, const void *y)
{
    const orc_skl *a = (const orc_skl *)x, *b = (const orc_skl *)y;
    int d = a->m - b->m;
    return d ? d : a->n - b->n;
}

int orc_stdskl(const orc_skl *skl, orc_skl *out)
{   /* gaps.cc:139-175 */
    int num = skl[0].n;
    if (num < 2) { for (int i = 0; i <= num; ++i) out[i] = skl[i]; return num; }
    orc_skl *org = (orc_skl *)malloc(sizeof(orc_skl) * (size_t)num);
    memcpy(org, skl + 1, sizeof(orc_skl) * (size_t)num);
    qsort(org, (size_t)num, sizeof(orc_skl), o_sklcmp);
    orc_skl *wrk = out + 1;
    int pr = 2, prv = 0;
    for (int i = 1; i < num; ++i) {
        int dm = org[i].m - org[prv].m, dn = org[i].n - org[prv].n;
        if (!dm && !dn) continue;
        if (dm < 0 || dn < 0) continue;
        int dd = dm < dn ? dm : dn;
        int df = dn - dm;
        if (df) df = df > 0 ? 1 : -1;
        if (dd && df) {
            if (pr) *wrk++ = org[prv];
            wrk->m = org[prv].m + dd; wrk->n = org[prv].n + dd; ++wrk;
        } else if (df != pr || !dm)
            *wrk++ = org[prv];
        pr = df;
        prv = i;
    }
    *wrk++ = org[prv];
    int cnt = (int)(wrk - out - 1);
    out[0].n = cnt; out[0].m = skl[0].m;
    free(org);
    return cnt;
}
