/* oracle/oracle_grp.c -- CPU restatement of the GROUP-TO-GROUP banded fill with path:
 * alignC<recd_t> = Fwd2c<recd_t>::Fwd2c + initB + forwardB + Vmf::traceback for recd_t = DPunit (groups
 * without internal gaps), DPunit_hf (one gap profile) and DPunit_pf (two gap profiles).
 * TEST INFRASTRUCTURE ONLY (see oracle.h for the rules).  Parity PINNED against the unmodified
 * reference (tests/golden/galign_*.json, frozen by tools/make_golden.py through ref_driver galign).
 *
 * Reference: src/fwd2c.h:81-100 (ctor), :138-176 (initB), :359-482 (forwardB); src/fwd2c.cc:32-102
 * (DPunit), :152-198 (DPunit_hf), :202-251 (DPunit_pf); src/gfreq.cc:507-605 (newgap / newdelta /
 * incdelta / copydelta); src/maln.h:185-187 (unp1), :262-312 (newgap1/2/3); src/maln2.cc:534-623,
 * 1230-1296 (sim2 kernels); src/dpunit.cc (reset / copy); src/vmf.cc:103-119 (traceback).
 *
 * Inputs are what the reference's own staging (PwdM::PwdM -> selAlnMode -> convseq / mkthick / Gfq)
 * presents to the DP through mSeqItr, column by column (orc_group).  Written row-major with rolling
 * rows of records; the reference keeps the same records in place, indexed by diagonal.
 */
#include "oracle.h"
#include <float.h>
#include <limits.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>

#define G_MIN(a, b) ((a) < (b) ? (a) : (b))
#define G_MAX(a, b) ((a) > (b) ? (a) : (b))

enum { G_DIAG = 2, G_NEWD = 3, G_VERT = 4, G_HORI = 8, G_NEWV = 12, G_NEWH = 13 };     /* aln.h:47-52 */
static int g_isdiag(int d) { d &= 15; return d == 2 || d == 3; }                       /* aln.h:60-62 */
static int g_isvert(int d) { d &= 15; return (d >= 4 && d <= 7) || d == 12; }
static int g_ishori(int d) { d &= 15; return (d >= 8 && d <= 11) || d == 13; }

typedef struct { int glen, nins; } g_idelta;            /* IDELTA, gfreq.h:26 */
static int g_neodelta(const g_idelta *d) { return d->glen < INT_MAX; }
static int g_neogfq(const orc_gfreq *g) { return g->glen >= 0; }
static const orc_gfreq g_endlist = {-1, 0, 0};

static void g_cleardelta(g_idelta *d)
{   /* gfreq.cc:556-560 */
    d[0].glen = 0; d[0].nins = 0;
    d[1].glen = INT_MAX; d[1].nins = 0;
}
static void g_copydelta(g_idelta *dst, const g_idelta *src)
{   /* gfreq.cc:548-554 */
    do { *dst++ = *src; } while (g_neodelta(++src));
    *dst = *src;
}
static void g_newdelta(g_idelta *dlt, const orc_gfreq *df, const g_idelta *dln, int n)
{   /* gfreq.cc:567-584: filter the dynamic state through a column's static gap state */
    g_idelta *dst = dlt;
    g_idelta tmp = {0, 0};
    for (; g_neogfq(df); ++df) {
        if (df->glen >= dln->glen) {
            while (df->glen >= dln[1].glen) ++dln;
            if (dln->nins > tmp.nins) {
                int dlnins = dln->nins;
                *dst++ = tmp;
                tmp.nins = dlnins;
                tmp.glen = df->glen + n;
            }
        }
    }
    *dst++ = tmp;
    dst->glen = INT_MAX; dst->nins = 0;
}
static void g_incdelta(g_idelta *dlt, const g_idelta *dln, int n)
{   /* gfreq.cc:595-602 */
    do { *dlt = *dln; (dlt++)->nins += n; } while (g_neodelta(++dln));
    *dlt = *dln;
}
static int g_gaplensd(const orc_gfreq *gf, const g_idelta *dl)
{   /* gfreq.h:67-71 */
    while (gf->glen >= dl[1].glen) ++dl;
    return gf->glen + dl->nins;
}

static const orc_gfreq *g_list(const orc_group *g, const int32_t *offs, int ix)
{ return (offs && offs[ix] >= 0) ? g->gpool + offs[ix] : &g_endlist; }

/* everything below is instantiated for float and double VTYPE */
#define DEFINE_GROUP_ALIGN(VT, SUFFIX, NEVSEL_V)                                                     \
typedef struct { VT val; int dir; long ptr; int glb; g_idelta *dla, *dlb; int *gla, *glv;             \
                 int lwr, upr, mlb, nlb;    /* SwgDPunit (dpunit.h:53-63): the box of the local path */\
} g_unit_##SUFFIX;                                                                                   \
                                                                                                     \
static VT g_newgap4_##SUFFIX(const orc_gfreq *cf, const g_idelta *dlc, const orc_gfreq *df, const g_idelta *dld)\
{   /* gfreq.cc:507-521 */                                                                           \
    VT g = 0;                                                                                        \
    for (; g_neogfq(df); ++df) {                                                                     \
        int j = g_gaplensd(df, dld);                                                                 \
        for (; g_neogfq(cf); ++cf) {                                                                 \
            int i = g_gaplensd(cf, dlc);                                                             \
            if (i >= j) break;                                                                       \
        }                                                                                            \
        if (!g_neogfq(cf)) break;                                                                    \
        g += (VT)cf->freq * (VT)df->freq;                                                            \
    }                                                                                                \
    return g;                                                                                        \
}                                                                                                    \
static VT g_newgap_cj_##SUFFIX(const orc_gfreq *cf, const g_idelta *dlc, int j)                      \
{   /* gfreq.cc:523-531 */                                                                           \
    for (; g_neogfq(cf); ++cf)                                                                       \
        if (g_gaplensd(cf, dlc) >= j) return (VT)cf->freq;                                           \
    return 0;                                                                                        \
}                                                                                                    \
static VT g_newgap_di_##SUFFIX(const orc_gfreq *df, int i, const g_idelta *dld)                      \
{   /* gfreq.cc:533-544 */                                                                           \
    VT g = 0;                                                                                        \
    for (; g_neogfq(df); ++df) {                                                                     \
        while (df->glen >= dld[1].glen) ++dld;                                                       \
        if (i < df->glen + dld->nins) break;                                                         \
        g += (VT)df->freq;                                                                           \
    }                                                                                                \
    return g;                                                                                        \
}                                                                                                    \
/* PwdM::newgap1(acf, dla, glb), maln.h:288-291 (+ newgapc :270-273) */                              \
static VT g_newgap1_##SUFFIX(VT wgop, const orc_gfreq *acf, const g_idelta *dla, int glb)            \
{                                                                                                    \
    if (!g_neogfq(acf)) return 0;                                                                    \
    if (g_neogfq(acf + 1)) return wgop * g_newgap_cj_##SUFFIX(acf, dla, glb);                        \
    return (dla->nins + acf->glen >= glb) ? (VT)(wgop * (VT)acf->freq) : 0;                          \
}                                                                                                    \
/* PwdM::newgap2(adf, glb, dla), maln.h:297-300 (+ newgapd :278-281) */                              \
static VT g_newgap2_##SUFFIX(VT wgop, const orc_gfreq *adf, int glb, const g_idelta *dla)            \
{                                                                                                    \
    if (!g_neogfq(adf)) return 0;                                                                    \
    if (g_neogfq(adf + 1)) return wgop * g_newgap_di_##SUFFIX(adf, glb, dla);                        \
    return (glb >= dla->nins + adf->glen) ? (VT)(wgop * (VT)adf->freq) : 0;                          \
}                                                                                                    \
                                                                                                     \
typedef struct {                                                                                     \
    const orc_group *a, *b;                                                                          \
    const orc_gparams *p;                                                                            \
    const double *mtx; int dim;                                                                      \
    int mode;              /* 0 DPunit, 1 DPunit_hf, 2 DPunit_pf */                                  \
    VT wgop, bgop;         /* Weighted_GOP, Basic_GOP */                                             \
} g_ctx_##SUFFIX;                                                                                    \
                                                                                                     \
/* sim2 (maln2.cc:534-623, 1230-1296; maln.h:160-168): ia / ib index the staged columns */          \
static VT g_sim2_##SUFFIX(const g_ctx_##SUFFIX *c, int ia, int ib)                                   \
{                                                                                                    \
    const orc_group *a = c->a, *b = c->b;                                                            \
    const uint8_t *ra = a->res + (size_t)ia * a->many, *rb = b->res + (size_t)ib * b->many;          \
    const double *va = a->vss ? a->vss + (size_t)ia * a->nelm : 0;                                   \
    const double *vb = b->vss ? b->vss + (size_t)ib * b->nelm : 0;                                   \
    const double *M = c->mtx; const int dim = c->dim;                                                \
    VT s = 0;                                                                                        \
    switch (3 * c->p->a_mode + c->p->b_mode) {                                                       \
    case 0: return (VT)M[ra[0] * dim + rb[0]];                                                       \
    case 1:                                                                                          \
        if (b->weight) { for (int j = 0; j < b->many; ++j) s += (VT)M[ra[0] * dim + rb[j]] * (VT)b->weight[j]; }\
        else for (int j = 0; j < b->many; ++j) s += (VT)M[ra[0] * dim + rb[j]];                      \
        return s;                                                                                    \
    case 2: return (VT)vb[b->felm + ra[0]];                                                          \
    case 3:                                                                                          \
        if (a->weight) { for (int i = 0; i < a->many; ++i) s += (VT)M[rb[0] * dim + ra[i]] * (VT)a->weight[i]; }\
        else for (int i = 0; i < a->many; ++i) s += (VT)M[rb[0] * dim + ra[i]];                      \
        return s;                                                                                    \
    case 4:                                                                                          \
        if (a->weight && b->weight) {                                                                \
            for (int i = 0; i < a->many; ++i) {                                                      \
                VT st = 0;                                                                           \
                for (int j = 0; j < b->many; ++j) st += (VT)M[ra[i] * dim + rb[j]] * (VT)b->weight[j];\
                s += st * (VT)a->weight[i];                                                          \
            }                                                                                        \
        } else for (int i = 0; i < a->many; ++i) for (int j = 0; j < b->many; ++j) s += (VT)M[ra[i] * dim + rb[j]];\
        return s;                                                                                    \
    case 5:                                                                                          \
        if (a->weight) { for (int i = 0; i < a->many; ++i) s += (VT)vb[b->felm + ra[i]] * (VT)a->weight[i]; }\
        else for (int i = 0; i < a->many; ++i) s += (VT)vb[b->felm + ra[i]];                         \
        return s;                                                                                    \
    case 6: return (VT)va[a->felm + rb[0]];                                                          \
    case 7:                                                                                          \
        if (b->weight) { for (int j = 0; j < b->many; ++j) s += (VT)va[a->felm + rb[j]] * (VT)b->weight[j]; }\
        else for (int j = 0; j < b->many; ++j) s += (VT)va[a->felm + rb[j]];                         \
        return s;                                                                                    \
    default:                                                                                         \
        if (c->p->dxd) {        /* sim33_n: compact nucleotide frequency vector (maln2.cc:615-623) */\
            static const int decompact[6] = {0, 1, 2, 3, 5, 9};    /* nil, gap, A, C, G, T (mseq.h:38) */\
            for (int k = 0; k < b->felm; ++k) s += (VT)va[a->felm + decompact[k]] * (VT)vb[k];       \
        } else for (int k = 0; k < b->felm; ++k) s += (VT)va[a->felm + k] * (VT)vb[k];               \
        return s;                                                                                    \
    }                                                                                                \
}                                                                                                    \
/* unp1 (maln.h:185-187): unpa(asi, bsi) = a.cfq * b.efq * -u ; unpb(bsi, asi) = b.cfq * a.efq * -u */\
static VT g_unp_##SUFFIX(const g_ctx_##SUFFIX *c, const orc_group *x, int ix, const orc_group *y, int iy)\
{ return (VT)x->cfq[ix] * (VT)y->efq[iy] * -(float)c->p->u; }                                        \
                                                                                                     \
/* gapopen (fwd2c.cc:52-91, 152-161, 202-212) */                                                     \
static VT g_gapopen_##SUFFIX(const g_ctx_##SUFFIX *c, const g_unit_##SUFFIX *r, int ia, int ib, int d3)\
{                                                                                                    \
    const orc_group *a = c->a, *b = c->b;                                                            \
    if (c->mode == 3) {         /* DPunit_nv: crg2 = crg11 / crg12 / crg21 / crg22 (i and w forms are the\
                                   many x many weighted form with unit weights; maln2.cc:881-1024,1454-1614);\
                                   no nil ends: gapdensity = IsGap, postgapdensity = 1 (mseq.h:150-160) */\
        const uint8_t *ra = a->res + (size_t)ia * a->many, *rb = b->res + (size_t)ib * b->many;      \
        const int *gla = r->gla, *glb = r->glv;                                                      \
        VT g = 0;                                                                                    \
        if (d3 == 0) {                                                                               \
            for (int i = 0; i < a->many; ++i) {                                                      \
                VT s = 0;                                                                            \
                if (ra[i] > 1) {                                                                     \
                    for (int j = 0; j < b->many; ++j)                                                \
                        if (rb[j] == 1 && gla[i] >= glb[j]) s += (VT)(b->weight ? b->weight[j] : 1) * (VT)1;\
                } else if (ra[i] == 1) {                                                             \
                    for (int j = 0; j < b->many; ++j)                                                \
                        if (rb[j] > 1 && glb[j] >= gla[i]) s += (VT)(b->weight ? b->weight[j] : 1) * (VT)1;\
                }                                                                                    \
                g += s * (VT)(a->weight ? a->weight[i] : 1);                                         \
            }                                                                                        \
        } else if (d3 > 0) {                                                                         \
            for (int i = 0; i < a->many; ++i) {                                                      \
                if (ra[i] > 1) {                                                                     \
                    VT s = 0;                                                                        \
                    for (int j = 0; j < b->many; ++j)                                                \
                        if (gla[i] >= glb[j]) s += (VT)(b->weight ? b->weight[j] : 1) * (VT)1;        \
                    g += s * (VT)(a->weight ? a->weight[i] : 1);                                     \
                }                                                                                    \
            }                                                                                        \
        } else {                                                                                     \
            for (int j = 0; j < b->many; ++j) {                                                      \
                if (rb[j] > 1) {                                                                     \
                    VT s = 0;                                                                        \
                    for (int i = 0; i < a->many; ++i)                                                \
                        if (glb[j] >= gla[i]) s += (VT)(a->weight ? a->weight[i] : 1) * (VT)1;        \
                    g += s * (VT)(b->weight ? b->weight[j] : 1);                                     \
                }                                                                                    \
            }                                                                                        \
        }                                                                                            \
        return (VT)(g * c->bgop);                                                                    \
    }                                                                                                \
    if (c->mode == 0) {         /* no di-thickness (only under -Q): thickness products */            \
        VT axb = 0;                                                                                  \
        if (d3 > 0) { if (!g_isvert(r->dir)) axb = (VT)a->cfq[ia] * (VT)b->efq[ib]; }                \
        else if (d3 < 0) { if (!g_ishori(r->dir)) axb = (VT)b->cfq[ib] * (VT)a->efq[ia]; }           \
        else return 0;                                                                               \
        return c->bgop * axb;                                                                        \
    }                                                                                                \
    if (c->mode == 1) {                                                                              \
        if (d3 == 0) return g_newgap2_##SUFFIX(c->wgop, g_list(a, a->tfq, ia), r->glb, r->dla);      \
        if (d3 > 0) return g_newgap1_##SUFFIX(c->wgop, g_list(a, a->sfq, ia), r->dla, r->glb);       \
        return g_newgap2_##SUFFIX(c->wgop, g_list(a, a->rfq, ia), r->glb, r->dla);                   \
    }                                                                                                \
    if (d3 == 0)                                                                                     \
        return g_newgap4_##SUFFIX(g_list(a, a->sfq, ia), r->dla, g_list(b, b->tfq, ib), r->dlb) * c->bgop\
             + g_newgap4_##SUFFIX(g_list(b, b->sfq, ib), r->dlb, g_list(a, a->tfq, ia), r->dla) * c->bgop;\
    if (d3 > 0) return g_newgap4_##SUFFIX(g_list(a, a->sfq, ia), r->dla, g_list(b, b->rfq, ib), r->dlb) * c->bgop;\
    return g_newgap4_##SUFFIX(g_list(b, b->sfq, ib), r->dlb, g_list(a, a->rfq, ia), r->dla) * c->bgop;\
}                                                                                                    \
/* update (fwd2c.cc:93-102, 163-182, 214-233); dst may alias src */                                  \
static void g_update_##SUFFIX(const g_ctx_##SUFFIX *c, g_unit_##SUFFIX *dst, const g_unit_##SUFFIX *src,\
                              int ia, int ib, VT gpn, int d3)                                        \
{                                                                                                    \
    const orc_group *a = c->a, *b = c->b;                                                            \
    int dir;                                                                                         \
    if (d3 > 0) dir = g_ishori(src->dir) ? G_NEWV : G_VERT;                                          \
    else if (d3 < 0) dir = g_isvert(src->dir) ? G_NEWH : G_HORI;                                     \
    else dir = g_isdiag(src->dir) ? G_DIAG : G_NEWD;                                                 \
    if (c->mode == 1) {                                                                              \
        if (d3 == 0) { g_newdelta(dst->dla, g_list(a, a->tfq, ia), src->dla, 1); dst->glb = 0; }     \
        else if (d3 > 0) { g_newdelta(dst->dla, g_list(a, a->tfq, ia), src->dla, 1); dst->glb = src->glb + 1; }\
        else { g_incdelta(dst->dla, src->dla, 1); dst->glb = 0; }                                    \
    } else if (c->mode == 2) {                                                                       \
        if (d3 == 0) { g_newdelta(dst->dla, g_list(a, a->tfq, ia), src->dla, 1); g_newdelta(dst->dlb, g_list(b, b->tfq, ib), src->dlb, 1); }\
        else if (d3 > 0) { g_newdelta(dst->dla, g_list(a, a->tfq, ia), src->dla, 1); g_incdelta(dst->dlb, src->dlb, 1); }\
        else { g_newdelta(dst->dlb, g_list(b, b->tfq, ib), src->dlb, 1); g_incdelta(dst->dla, src->dla, 1); }\
    }                                                                                                \
    if (c->mode == 3) {         /* elongap (mgaps.cc:442-451) on both run-length vectors */          \
        const uint8_t *ra = a->res + (size_t)ia * a->many, *rb = b->res + (size_t)ib * b->many;      \
        for (int i = 0; i < a->many; ++i) dst->gla[i] = (d3 >= 0) ? (ra[i] <= 1 ? src->gla[i] + 1 : 0) : src->gla[i] + 1;\
        for (int j = 0; j < b->many; ++j) dst->glv[j] = (d3 <= 0) ? (rb[j] <= 1 ? src->glv[j] + 1 : 0) : src->glv[j] + 1;\
    }                                                                                                \
    dst->dir = dir;                                                                                  \
    dst->val = src->val + gpn;                                                                       \
    dst->ptr = src->ptr;                                                                             \
}                                                                                                    \
static void g_reset_##SUFFIX(const g_ctx_##SUFFIX *c, g_unit_##SUFFIX *r)                            \
{   /* reset<recd_t> (dpunit.cc) */                                                                  \
    r->val = NEVSEL_V; r->dir = 0; r->ptr = 0; r->glb = 0;                                           \
    if (c->mode == 1 || c->mode == 2) g_cleardelta(r->dla);                                          \
    if (c->mode == 2) g_cleardelta(r->dlb);                                                          \
    if (c->mode == 3) { memset(r->gla, 0, sizeof(int) * (size_t)c->a->many); memset(r->glv, 0, sizeof(int) * (size_t)c->b->many); }\
}                                                                                                    \
static void g_copy_##SUFFIX(const g_ctx_##SUFFIX *c, g_unit_##SUFFIX *d, const g_unit_##SUFFIX *s)   \
{   /* copy<recd_t> (dpunit.cc) */                                                                   \
    if (d == s) return;                                                                              \
    d->val = s->val; d->dir = s->dir; d->ptr = s->ptr; d->glb = s->glb;                              \
    if (c->mode == 1 || c->mode == 2) g_copydelta(d->dla, s->dla);                                   \
    if (c->mode == 2) g_copydelta(d->dlb, s->dlb);                                                   \
    if (c->mode == 3) { memcpy(d->gla, s->gla, sizeof(int) * (size_t)c->a->many); memcpy(d->glv, s->glv, sizeof(int) * (size_t)c->b->many); }\
}                                                                                                    \
                                                                                                     \
static int g_align_##SUFFIX(const orc_group *a, const orc_group *b, const double *mtx, int dim,      \
                            const orc_gparams *p, double *score, orc_skl *out, int cap, int64_t *cells,\
                            long *rr /* HomScoreC: Fwd2c without Vmf, fwd2c.h:663-668 */)           \
{                                                                                                    \
    typedef g_unit_##SUFFIX U;                                                                       \
    g_ctx_##SUFFIX C;                                                                                \
    C.a = a; C.b = b; C.p = p; C.mtx = mtx; C.dim = dim;                                             \
    /* alnmode 1 (NGP_ALN): the rectangle variant forwardA + initA (fwd2c.h:111-135,231-356), groups without gap profile */\
    const int rect = p->alnmode == 1;                                                               \
    const int am = rect ? 6 : p->alnmode;                                                           \
    C.mode = am == 6 ? 0 : (am == 9 ? 2 : (am == 10 ? 3 : 1));           /* NGP / HLF, RHF / GPF / NTV */\
    C.wgop = (VT)p->Weighted_GOP; C.bgop = (VT)p->Basic_GOP;                                         \
    const VT BasicGOP = (VT)p->BasicGOP, BasicGEP = (VT)p->BasicGEP, LongGOP = (VT)p->LongGOP, LongGEP = (VT)p->LongGEP;\
    const VT u2divu1 = BasicGEP < 0 ? (VT)LongGEP / BasicGEP : 0;       /* fwd2c.h:85-86 */          \
    const VT v2divv1 = BasicGOP < 0 ? (VT)LongGOP / BasicGOP : 0;                                    \
    const int Noll = p->Noll, codonk1 = p->codonk1;                                                  \
    const int al = a->left, ar = a->right, bl = b->left, br = b->right;                              \
    orc_seq sa = {0, a->len, al, ar, 0, 0}, sb = {0, b->len, bl, br, 0, 0};                          \
    orc_window w;                                                                                    \
    orc_stripe(&sa, &sb, p->sh, &w);                                                                 \
    const int lw = rect ? bl - ar : w.lw, up = rect ? br - al : w.up;    /* rectangle: every cell */  \
    const int capa = (a->hetero > 0 ? a->hetero : 0) + 3, capb = (b->hetero > 0 ? b->hetero : 0) + 3;\
    /* staged column index of sequence position x: x - (left - 1) */                                 \
    const int A0 = al - 1, B0 = bl - 1;                                                              \
    const int NB = br - bl + 2;                                                                      \
    const int nrec = 6 * NB + 8;                                                                     \
    U *buf = (U *)malloc(sizeof(U) * (size_t)nrec);                                                  \
    g_idelta *pool = (g_idelta *)malloc(sizeof(g_idelta) * (size_t)nrec * (size_t)(capa + capb));    \
    int *glpool = (int *)calloc((size_t)nrec * (size_t)(a->many + b->many), sizeof(int));            \
    for (int i = 0; i < nrec; ++i) {                                                                 \
        buf[i].gla = glpool + (size_t)i * (a->many + b->many); buf[i].glv = buf[i].gla + a->many;    \
        buf[i].dla = pool + (size_t)i * (capa + capb); buf[i].dlb = buf[i].dla + capa;               \
        g_cleardelta(buf[i].dla); g_cleardelta(buf[i].dlb);                                          \
        buf[i].val = NEVSEL_V; buf[i].dir = 0; buf[i].ptr = 0; buf[i].glb = 0;                       \
    }                                                                                                \
    U *Hp = buf, *Gp = buf + NB, *G2p = buf + 2 * NB, *Hc = buf + 3 * NB, *Gc = buf + 4 * NB, *G2c = buf + 5 * NB;\
    U *f1 = buf + 6 * NB, *f2 = f1 + 1, *black = f1 + 2, *colprev = f1 + 3, *dg = f1 + 4, *g = f1 + 5, *g2 = f1 + 6;\
    g_reset_##SUFFIX(&C, black);                                                                     \
    o_store st = {0, 0, 0};                                                                          \
    o_add(&st, 0, 0, 0);                           /* skip 0-th record (fwd2c.h:361) */              \
    int64_t ncell = 0;                                                                               \
    /* initB (fwd2c.h:138-176): origin, then the boundary row with asi at a.left - 1 */              \
    Hp[0].val = 0; Hp[0].dir = rect ? 0 : G_DIAG /* initA only clears the origin (:116) */; Hp[0].ptr = rr ? (long)(bl - al) : o_add(&st, al, bl, 0);  /* fwd2c.h:144 */\
    {                                                                                                \
        int rr = br - al; if (up < rr) rr = up;                                                      \
        const int r0 = bl - al;                                                                      \
        for (int r = r0 + 1, k = 1; r <= rr; ++r, ++k) {                                             \
            const int ia = al - 1 - A0, ib = (bl + k - 1) - B0;                                      \
            VT pub = g_unp_##SUFFIX(&C, b, ib, a, ia);                                               \
            VT gnp = g_gapopen_##SUFFIX(&C, &Hp[k - 1], ia, ib, -1);                                 \
            /* initB tests the column count after its increment (:156-160), initA before (:125-131) */ \
            gnp = ((rect ? k - 1 : k) < codonk1) ? gnp + pub : (VT)(v2divv1 * gnp + u2divu1 * pub);  \
            g_update_##SUFFIX(&C, &Hp[k], &Hp[k - 1], ia, ib, gnp, -1);                              \
        }                                                                                            \
    }                                                                                                \
    g_copy_##SUFFIX(&C, colprev, &Hp[0]);                                                            \
    int colk = 0;                                                                                    \
    const int rr_col = (bl - ar > lw) ? bl - ar : lw;                                                \
    for (int m = al; m < ar; ++m) {                                                                  \
        const int n0 = G_MAX(m + lw, bl), n9 = G_MIN(m + up + 1, br);                                \
        const int ia = m - A0;                                                                       \
        /* boundary column cell of this row (initB second loop), bsi at b.left - 1 */                \
        {                                                                                            \
            int r = bl - 1 - m;                                                                      \
            if (rect) {                                                                              \
                /* forwardA computes the boundary cell of a row in place (:245-249): no long-gap switch, and bsi    \
                   is where the previous row left it -- position 0 before the first row (`mSeqItr bsi(b, 0)`,     \
                   :240), position b.right afterwards */                                             \
                const int ib = (m == al ? 0 : br) - B0;                                              \
                VT gnp = g_gapopen_##SUFFIX(&C, colprev, ia, ib, 1) + g_unp_##SUFFIX(&C, a, ia, b, ib);\
                g_update_##SUFFIX(&C, &Hc[0], colprev, ia, ib, gnp, 1);                              \
                g_copy_##SUFFIX(&C, colprev, &Hc[0]);                                                \
            } else if (r >= rr_col) {                                                                \
                ++colk;                                                                              \
                const int ib = bl - 1 - B0;                                                          \
                VT pua = g_unp_##SUFFIX(&C, a, ia, b, ib);                                           \
                VT gnp = g_gapopen_##SUFFIX(&C, colprev, ia, ib, 1);                                 \
                gnp = (colk < codonk1) ? gnp + pua : (VT)(v2divv1 * gnp + u2divu1 * pua);            \
                g_update_##SUFFIX(&C, &Hc[0], colprev, ia, ib, gnp, 1);                              \
                g_copy_##SUFFIX(&C, colprev, &Hc[0]);                                                \
            } else g_copy_##SUFFIX(&C, &Hc[0], black);                                               \
        }                                                                                            \
        /* pua is evaluated once per row, with bsi at the row's first column (fwd2c.h:377) */        \
        VT pua = n0 < br ? g_unp_##SUFFIX(&C, a, ia, b, n0 - B0) : 0;                                \
        g_reset_##SUFFIX(&C, f1);                                                                    \
        g_reset_##SUFFIX(&C, f2);                                                                    \
        for (int n = n0; n < n9; ++n) {                                                              \
            const int j = n - bl + 1, ib = n - B0;                                                   \
            ++ncell;                                                                                 \
            const int above_inband = (n - m + 1 <= up);                                              \
            const U *habove = above_inband ? &Hp[j] : black, *gabove = above_inband ? &Gp[j] : black;\
            const U *g2above = above_inband ? &G2p[j] : black;                                       \
            const U *hleft = (n - 1 >= n0 || n - 1 == bl - 1) ? &Hc[j - 1] : black;                  \
            /* diagonal (fwd2c.h:395-398) */                                                         \
            VT dab = g_sim2_##SUFFIX(&C, ia, ib);                                                    \
            VT gop = g_gapopen_##SUFFIX(&C, &Hp[j - 1], ia, ib, 0);                                  \
            g_update_##SUFFIX(&C, dg, &Hp[j - 1], ia, ib, dab + gop, 0);                             \
            VT gnp;                                                                                  \
            const U *mx;                                                                             \
            if (m > al || rect) {   /* vertical (fwd2c.h:401-409; forwardA has no first-row skip, :263-271) */\
                if (a->nils || rect) pua = g_unp_##SUFFIX(&C, a, ia, b, ib);                         \
                gnp = g_gapopen_##SUFFIX(&C, gabove, ia, ib, 1);                                     \
                gop = g_gapopen_##SUFFIX(&C, habove, ia, ib, 1);                                     \
                if (!g_isvert(habove->dir) && (habove->val + gop > gabove->val + gnp))               \
                    g_update_##SUFFIX(&C, g, habove, ia, ib, gop, 1);                                \
                else g_update_##SUFFIX(&C, g, gabove, ia, ib, gnp, 1);                               \
                g->val += pua;                                                                       \
                mx = g;                                                                              \
                if (Noll == 3) {    /* vertical2 (fwd2c.h:411-420) */                                \
                    gnp = (VT)(v2divv1 * g_gapopen_##SUFFIX(&C, g2above, ia, ib, 1));                \
                    gop = rect ? (VT)(v2divv1 + gop) : (VT)(v2divv1 * gop);     /* forwardA: `+`, fwd2c.h:276 */\
                    if (!g_isvert(habove->dir) && (habove->val + gop > g2above->val + gnp))          \
                        g_update_##SUFFIX(&C, g2, habove, ia, ib, gop, 1);                           \
                    else g_update_##SUFFIX(&C, g2, g2above, ia, ib, gnp, 1);                         \
                    g2->val += (VT)(u2divu1 * pua);                                                  \
                    if (g2->val > mx->val) mx = g2;                                                  \
                }                                                                                    \
            } else {            /* first row: g keeps the untouched buffer record */                 \
                g_copy_##SUFFIX(&C, g, black); g_copy_##SUFFIX(&C, g2, black);                       \
                mx = g;                                                                              \
            }                                                                                        \
            if (n > bl || rect) {   /* horizontal (fwd2c.h:422-431; forwardA: no first-column skip, :286-294) */\
                VT pub = g_unp_##SUFFIX(&C, b, ib, a, ia);                                           \
                gnp = g_gapopen_##SUFFIX(&C, f1, ia, ib, -1);                                        \
                gop = g_gapopen_##SUFFIX(&C, hleft, ia, ib, -1);                                     \
                if (!g_ishori(hleft->dir) && (hleft->val + gop > f1->val + gnp))                     \
                    g_update_##SUFFIX(&C, f1, hleft, ia, ib, gop, -1);                               \
                else g_update_##SUFFIX(&C, f1, f1, ia, ib, gnp, -1);                                 \
                f1->val += pub;                                                                      \
                if (f1->val >= mx->val) mx = f1;                                                     \
                if (Noll == 3) {    /* horizontal2 (fwd2c.h:433-442) */                              \
                    gnp = (VT)(v2divv1 * g_gapopen_##SUFFIX(&C, f2, ia, ib, -1));                    \
                    gop = (VT)(v2divv1 * gop);                                                       \
                    if (!g_ishori(hleft->dir) && (hleft->val + gop > f2->val + gnp))                 \
                        g_update_##SUFFIX(&C, f2, hleft, ia, ib, gop, -1);                           \
                    else g_update_##SUFFIX(&C, f2, f2, ia, ib, gnp, -1);                             \
                    f2->val += (VT)(u2divu1 * pub);                                                  \
                    if (f2->val >= mx->val) mx = f2;                                                 \
                }                                                                                    \
            }                                                                                        \
            if (mx->val > dg->val) g_copy_##SUFFIX(&C, &Hc[j], mx);     /* fwd2c.h:453 */            \
            else g_copy_##SUFFIX(&C, &Hc[j], dg);                                                    \
            if (rr) { if (m == al) Hc[j].ptr = n - m; }                 /* fwd2c.h:468-469 */        \
            else if (Hc[j].dir == G_NEWD || Hc[j].dir == G_NEWV || Hc[j].dir == G_NEWH)              \
                Hc[j].ptr = o_add(&st, m, n, Hc[j].ptr);                /* fwd2c.h:465-467 */        \
            g_copy_##SUFFIX(&C, &Gc[j], g);                                                          \
            g_copy_##SUFFIX(&C, &G2c[j], g2);                                                        \
        }                                                                                            \
        U *t;                                                                                        \
        t = Hp; Hp = Hc; Hc = t; t = Gp; Gp = Gc; Gc = t; t = G2p; G2p = G2c; G2c = t;               \
    }                                                                                                \
    /* result cell H[b.right - a.right] (fwd2c.h:475-481) */                                         \
    const U *last = &Hp[br - 1 - bl + 1];                                                            \
    *score = (double)last->val;                                                                      \
    if (cells) *cells = ncell;                                                                       \
    if (rr) {                                      /* pp[] of forwardB (fwd2c.h:476-479) */          \
        rr[0] = last->ptr; rr[1] = (long)(bl - al) + (br - ar);                                      \
        free(buf); free(pool); free(glpool); free(st.v);                                             \
        return 0;                                                                                    \
    }                                                                                                \
    long pp = o_add(&st, ar, br, last->ptr);                                                         \
    int cnt = 0, ok = 1;                                                                             \
    for (long q = pp;; q = st.v[q].p) {            /* Vmf::traceback (vmf.cc:103-119) */             \
        if (cnt + 1 >= cap) { ok = 0; break; }                                                       \
        out[++cnt].m = st.v[q].m; out[cnt].n = st.v[q].n;                                            \
        if (!st.v[q].p) break;                                                                       \
    }                                                                                                \
    out[0].m = 0; out[0].n = cnt;                                                                    \
    free(buf); free(pool); free(glpool); free(st.v);                                                 \
    return ok ? cnt : -1;                                                                            \
}                                                                                                    \
/* ---- Smith-Waterman: Fwd2c<SwgDPunit*>::initC + forwardC (fwd2c.h:178-207,483-659) for algmode.mlt <= 1, i.e.            \
 * without secondary colonies: the best local score and its box (colony 0).  Records carry (lwr, upr, mlb, nlb);            \
 * SwgDPunit's own gapopen / update (fwd2c.cc:257-296) for groups without gap profile, the banded rules plus the box        \
 * for the others (:298-428); a negative cell is cleared together with the horizontal states (and G2, not G: :601-605). */  \
static void g_swg_blank_##SUFFIX(const g_ctx_##SUFFIX *c, g_unit_##SUFFIX *r, VT v)                                         \
{   /* clear / reset<SwgDPunit*> (dpunit.cc): blank_swgdpunit / black_swgdpunit */                                          \
    g_reset_##SUFFIX(c, r);                                                                                                 \
    r->val = v; r->lwr = INT_MAX / 8 * 7; r->upr = INT_MIN / 8 * 7; r->mlb = r->nlb = 0;                                    \
}                                                                                                                           \
static void g_swg_copy_##SUFFIX(const g_ctx_##SUFFIX *c, g_unit_##SUFFIX *d, const g_unit_##SUFFIX *s)                      \
{                                                                                                                           \
    if (d == s) return;                                                                                                     \
    g_copy_##SUFFIX(c, d, s);                                                                                               \
    d->lwr = s->lwr; d->upr = s->upr; d->mlb = s->mlb; d->nlb = s->nlb;                                                     \
}                                                                                                                           \
static VT g_swg_gapopen_##SUFFIX(const g_ctx_##SUFFIX *c, const g_unit_##SUFFIX *r, int ia, int ib, int d3)                 \
{                                                                                                                           \
    if (c->mode == 0) return (g_isdiag(r->dir) && d3) ? (VT)c->p->BasicGOP : 0;      /* fwd2c.cc:268-271 */                 \
    return g_gapopen_##SUFFIX(c, r, ia, ib, d3);                                                                            \
}                                                                                                                           \
static void g_swg_update_##SUFFIX(const g_ctx_##SUFFIX *c, g_unit_##SUFFIX *dst, const g_unit_##SUFFIX *src,                \
                                  int ia, int ib, VT gpn, int d3, int r)                                                    \
{   /* swg_gdpunit_update (fwd2c.cc:273-289) + the list part of the record type; dst may alias src */                       \
    const int sdir = src->dir, lwr = src->lwr, upr = src->upr, mlb = src->mlb, nlb = src->nlb;                              \
    g_update_##SUFFIX(c, dst, src, ia, ib, gpn, d3);                                                                        \
    dst->lwr = lwr; dst->upr = upr; dst->mlb = mlb; dst->nlb = nlb;                                                         \
    if (d3 > 0) { dst->dir = G_VERT; if (r < dst->lwr) dst->lwr = r; }                                                      \
    else if (d3 < 0) { dst->dir = G_HORI; if (r > dst->upr) dst->upr = r; }                                                 \
    else dst->dir = g_isdiag(sdir) ? G_DIAG : G_NEWD;                                                                       \
}                                                                                                                           \
static int g_swg_##SUFFIX(const orc_group *a, const orc_group *b, const double *mtx, int dim,                               \
                          const orc_gparams *p, double *val, int *box, int64_t *cells)                                      \
{                                                                                                                           \
    typedef g_unit_##SUFFIX U;                                                                                              \
    g_ctx_##SUFFIX C;                                                                                                       \
    C.a = a; C.b = b; C.p = p; C.mtx = mtx; C.dim = dim;                                                                    \
    const int am = p->alnmode;                                                                                              \
    C.mode = am == 6 ? 0 : (am == 9 ? 2 : (am == 10 ? 3 : 1));                                                              \
    C.wgop = (VT)p->Weighted_GOP; C.bgop = (VT)p->Basic_GOP;                                                                \
    const VT BasicGOP = (VT)p->BasicGOP, BasicGEP = (VT)p->BasicGEP, LongGOP = (VT)p->LongGOP, LongGEP = (VT)p->LongGEP;    \
    const VT u2divu1 = BasicGEP < 0 ? (VT)LongGEP / BasicGEP : 0;                                                           \
    const VT v2divv1 = BasicGOP < 0 ? (VT)LongGOP / BasicGOP : 0;                                                           \
    const int Noll = p->Noll;                                                                                               \
    const int al = a->left, ar = a->right, bl = b->left, br = b->right;                                                     \
    orc_seq sa = {0, a->len, al, ar, 0, 0}, sb = {0, b->len, bl, br, 0, 0};                                                 \
    orc_window w;                                                                                                           \
    orc_stripe(&sa, &sb, p->sh, &w);                                                                                        \
    const int lw = w.lw, up = w.up;                                                                                         \
    const int capa = (a->hetero > 0 ? a->hetero : 0) + 3, capb = (b->hetero > 0 ? b->hetero : 0) + 3;                       \
    const int A0 = al - 1, B0 = bl - 1;                                                                                     \
    const int NB = br - bl + 2;                                                                                             \
    const int nrec = 6 * NB + 8;                                                                                            \
    U *buf = (U *)malloc(sizeof(U) * (size_t)nrec);                                                                         \
    g_idelta *pool = (g_idelta *)malloc(sizeof(g_idelta) * (size_t)nrec * (size_t)(capa + capb));                           \
    int *glpool = (int *)calloc((size_t)nrec * (size_t)(a->many + b->many), sizeof(int));                                   \
    for (int i = 0; i < nrec; ++i) {                                                                                        \
        buf[i].gla = glpool + (size_t)i * (a->many + b->many); buf[i].glv = buf[i].gla + a->many;                           \
        buf[i].dla = pool + (size_t)i * (capa + capb); buf[i].dlb = buf[i].dla + capa;                                      \
        g_swg_blank_##SUFFIX(&C, &buf[i], NEVSEL_V);                                                                        \
    }                                                                                                                       \
    U *Hp = buf, *Gp = buf + NB, *G2p = buf + 2 * NB, *Hc = buf + 3 * NB, *Gc = buf + 4 * NB, *G2c = buf + 5 * NB;          \
    U *f1 = buf + 6 * NB, *f2 = f1 + 1, *black = f1 + 2, *dg = f1 + 4, *g = f1 + 5, *g2 = f1 + 6;                           \
    int64_t ncell = 0;                                                                                                      \
    VT c0val = 0;                                   /* colony 0, zeroed by Colonies::Colonies (aln2.cc:418-425) */          \
    int c0[6] = {0, 0, 0, 0, 0, 0};                 /* mlb nlb mrb nrb lwr upr */                                           \
    (void)BasicGOP;                                                                                                         \
    /* initC (:178-207): the records in front of the first row / first column, on the diagonal of the cell that             \
       takes them as its diagonal predecessor */                                                                            \
    {                                                                                                                       \
        int rr = br - al; if (up < rr) rr = up;                                                                             \
        for (int r = bl - al, n = bl; r <= rr && n - bl < NB; ++r, ++n) {                                                   \
            U *h = &Hp[n - bl];                                                                                             \
            g_swg_blank_##SUFFIX(&C, h, 0); h->upr = h->lwr = r; h->mlb = al; h->nlb = n;                                   \
        }                                                                                                                   \
    }                                                                                                                       \
    const int rr_col = (bl - ar > lw) ? bl - ar : lw;                                                                       \
    for (int m = al; m < ar; ++m) {                                                                                         \
        const int n0 = G_MAX(m + lw, bl), n9 = G_MIN(m + up + 1, br);                                                       \
        const int ia = m - A0;                                                                                              \
        if (m > al) {                               /* H(m - 1, bl - 1): initC's second loop, r = bl - m */                 \
            if (bl - m >= rr_col) { g_swg_blank_##SUFFIX(&C, &Hp[0], 0); Hp[0].upr = Hp[0].lwr = bl - m; Hp[0].mlb = m; Hp[0].nlb = bl; } \
            else g_swg_blank_##SUFFIX(&C, &Hp[0], NEVSEL_V);                                                                \
        }                                                                                                                   \
        g_swg_blank_##SUFFIX(&C, &Hc[0], NEVSEL_V);                                                                         \
        g_swg_blank_##SUFFIX(&C, f1, NEVSEL_V);                                                                             \
        g_swg_blank_##SUFFIX(&C, f2, NEVSEL_V);                                                                             \
        for (int n = n0; n < n9; ++n) {                                                                                     \
            const int j = n - bl + 1, ib = n - B0, r = n - m;                                                               \
            ++ncell;                                                                                                        \
            const int above_inband = (r + 1 <= up);                                                                         \
            const U *hdiag = &Hp[j - 1];                                                                                    \
            const U *habove = above_inband ? &Hp[j] : black, *gabove = above_inband ? &Gp[j] : black;                       \
            const U *g2above = above_inband ? &G2p[j] : black;                                                              \
            const U *hleft = (n - 1 >= n0) ? &Hc[j - 1] : black;                                                            \
            const VT diag = hdiag->val;                                                                                     \
            VT dab = g_sim2_##SUFFIX(&C, ia, ib);                                                                           \
            VT gop = g_swg_gapopen_##SUFFIX(&C, hdiag, ia, ib, 0);                                                          \
            g_swg_update_##SUFFIX(&C, dg, hdiag, ia, ib, dab + gop, 0, r);                                                  \
            VT gnp;                                                                                                         \
            const U *mx = g;                                                                                                \
            if (m > al) {                           /* vertical (:531-541) */                                               \
                VT pua = g_unp_##SUFFIX(&C, a, ia, b, ib);                                                                  \
                gnp = g_swg_gapopen_##SUFFIX(&C, gabove, ia, ib, 1);                                                        \
                gop = g_swg_gapopen_##SUFFIX(&C, habove, ia, ib, 1);                                                        \
                if (!g_isvert(habove->dir) && (habove->val + gop > gabove->val + gnp))                                      \
                    g_swg_update_##SUFFIX(&C, g, habove, ia, ib, gop, 1, r);                                                \
                else g_swg_update_##SUFFIX(&C, g, gabove, ia, ib, gnp, 1, r);                                               \
                g->val += pua;                                                                                              \
                if (Noll == 3) {                    /* vertical2 (:543-552) */                                              \
                    gnp = (VT)(v2divv1 * g_swg_gapopen_##SUFFIX(&C, g2above, ia, ib, 1));                                   \
                    gop = (VT)(v2divv1 * gop);                                                                              \
                    if (!g_isvert(habove->dir) && (habove->val + gop > g2above->val + gnp))                                 \
                        g_swg_update_##SUFFIX(&C, g2, habove, ia, ib, gop, 1, r);                                           \
                    else g_swg_update_##SUFFIX(&C, g2, g2above, ia, ib, gnp, 1, r);                                         \
                    g2->val += (VT)(u2divu1 * pua);                                                                         \
                    if (g2->val > mx->val) mx = g2;                                                                         \
                }                                                                                                           \
            } else {                                /* first row: the untouched G / G2 records */                           \
                g_swg_blank_##SUFFIX(&C, g, NEVSEL_V); g_swg_blank_##SUFFIX(&C, g2, NEVSEL_V);                              \
            }                                                                                                               \
            if (n > bl) {                           /* horizontal (:555-564) */                                             \
                VT pub = g_unp_##SUFFIX(&C, b, ib, a, ia);                                                                  \
                gnp = g_swg_gapopen_##SUFFIX(&C, f1, ia, ib, -1);                                                           \
                gop = g_swg_gapopen_##SUFFIX(&C, hleft, ia, ib, -1);                                                        \
                if (!g_ishori(hleft->dir) && (hleft->val + gop > f1->val + gnp))                                            \
                    g_swg_update_##SUFFIX(&C, f1, hleft, ia, ib, gop, -1, r);                                               \
                else g_swg_update_##SUFFIX(&C, f1, f1, ia, ib, gnp, -1, r);                                                 \
                f1->val += pub;                                                                                             \
                if (f1->val >= mx->val) mx = f1;                                                                            \
                if (Noll == 3) {                    /* horizontal2 (:566-575) */                                            \
                    gnp = (VT)(v2divv1 * g_swg_gapopen_##SUFFIX(&C, f2, ia, ib, -1));                                       \
                    gop = (VT)(v2divv1 * gop);                                                                              \
                    if (!g_ishori(hleft->dir) && (hleft->val + gop > f2->val + gnp))                                        \
                        g_swg_update_##SUFFIX(&C, f2, hleft, ia, ib, gop, -1, r);                                           \
                    else g_swg_update_##SUFFIX(&C, f2, f2, ia, ib, gnp, -1, r);                                             \
                    f2->val += (VT)(u2divu1 * pub);                                                                         \
                    if (f2->val >= mx->val) mx = f2;                                                                        \
                }                                                                                                           \
            }                                                                                                               \
            U *h = &Hc[j];                                                                                                  \
            if (mx->val > dg->val) {                /* non-diagonal (:586-589) */                                           \
                g_swg_copy_##SUFFIX(&C, h, mx);                                                                             \
                if (h->lwr > r) h->lwr = r;                                                                                 \
                if (h->upr < r) h->upr = r;                                                                                 \
            } else {                                                                                                        \
                g_swg_copy_##SUFFIX(&C, h, dg);                                                                             \
                if (h->val > diag) {                                                                                        \
                    if (diag == 0) { h->upr = h->lwr = r; h->mlb = m; h->nlb = n; }     /* new colony (:591-595) */         \
                    if (h->val > c0val) {           /* max local score (:596-604) */                                        \
                        c0val = h->val;                                                                                     \
                        c0[0] = h->mlb; c0[1] = h->nlb; c0[2] = m + 1; c0[3] = n + 1; c0[4] = h->lwr; c0[5] = h->upr;       \
                    }                                                                                                       \
                }                                                                                                           \
            }                                                                                                               \
            if (h->val < 0) {                       /* reset to blank (:606-610): h, f1 (twice), f2 and g2 -- not g */      \
                g_swg_blank_##SUFFIX(&C, h, 0); g_swg_blank_##SUFFIX(&C, f1, 0);                                            \
                if (Noll == 3) { g_swg_blank_##SUFFIX(&C, f2, 0); g_swg_blank_##SUFFIX(&C, g2, 0); }                        \
            }                                                                                                               \
            g_swg_copy_##SUFFIX(&C, &Gc[j], g);                                                                             \
            g_swg_copy_##SUFFIX(&C, &G2c[j], g2);                                                                           \
        }                                                                                                                   \
        U *t;                                                                                                               \
        t = Hp; Hp = Hc; Hc = t; t = Gp; Gp = Gc; Gc = t; t = G2p; G2p = G2c; G2c = t;                                      \
    }                                                                                                                       \
    *val = (double)c0val;                                                                                                   \
    for (int k = 0; k < 6; ++k) box[k] = c0[k];                                                                             \
    if (cells) *cells = ncell;                                                                                              \
    free(buf); free(pool); free(glpool);                                                                                    \
    return 0;                                                                                                               \
}

typedef struct { int32_t m, n; long p; } o_rec;
typedef struct { o_rec *v; long n, cap; } o_store;
static long o_add(o_store *s, int m, int n, long p)
{
    if (s->n == s->cap) { s->cap = s->cap ? 2 * s->cap : 1024; s->v = (o_rec *)realloc(s->v, sizeof(o_rec) * (size_t)s->cap); }
    s->v[s->n].m = m; s->v[s->n].n = n; s->v[s->n].p = p;
    return s->n++;
}

/* ================================================================================================
 * Aln2b1: alignB_ng / HomScoreB_ng for two single sequences (reference src/fwd2b1.cc: initB_ng :64-98,
 * forwardB_ng :145-279, lastB_ng :100-143, trcbkalignB_ng :1025-1051, globalB_ng :1286-1315), global
 * mode with full terminal gap penalties (tgapf == 1, no exgr): the path prrn5's DynAln distances take
 * (src/adjmat.cc:84).  Differs from Fwd2c<DPunit>: a gap opens on >= (:192,213), G displaces the
 * diagonal on > and F on >= (:198,219), gap states keep the record of the cell they opened from and
 * only cells where a diagonal run resumes (NEWD) append a path record.
 * ================================================================================================ */
#define DEFINE_ALIGN_B1(VT, SUFFIX, NEVSEL_V)                                                        \
typedef struct { VT val; long ptr; int dir; } b1_unit_##SUFFIX;                                      \
static int align_b1_##SUFFIX(const orc_seq *a, const orc_seq *b, const double *mtx, int dim,         \
                             const orc_params *p, double *score, orc_skl *out, int cap)              \
{                                                                                                    \
    typedef b1_unit_##SUFFIX U;                                                                      \
    const float fu = (float)p->u, fv = (float)p->v, fu1 = (float)p->u1, fsc = (float)p->scale;       \
    const VT Vab = (VT)(fsc * 1 * 1);                      /* PwdB::PwdB, aln2.cc:97-117 */           \
    const VT BasicGOP = (VT)(-fv * Vab), BasicGEP = (VT)(-fu * Vab), LongGEP = (VT)(-fu1 * Vab);     \
    const VT diffu = LongGEP - BasicGEP;                                                             \
    const VT LongGOP = BasicGOP - diffu * p->k1;                                                     \
    const int Noll = p->ls < 2 ? 2 : (p->ls > 3 ? 3 : p->ls);                                        \
    const int codonk1 = p->ls == 3 ? p->k1 : (INT_MAX / 8 * 7);                                      \
    if (p->lcl & 16) return -2;                               /* fwdswgB_ng is another function */   \
    orc_window w;                                                                                    \
    orc_stripe(a, b, p->sh, &w);                                                                     \
    const int lw = w.lw, up = w.up;                                                                  \
    const int al = a->left, ar = a->right, bl = b->left, br = b->right;                              \
    const U black = {NEVSEL_V, 0, 0};                                                                \
    o_store st = {0, 0, 0};                                                                          \
    o_add(&st, 0, 0, 0);                                                                             \
    const long origin = o_add(&st, al, bl, 0);                                                       \
    const int NB = br - bl + 2;                                                                      \
    U *buf = (U *)malloc(sizeof(U) * 6 * (size_t)NB);                                                \
    U *Hp = buf, *Gp = buf + NB, *G2p = buf + 2 * NB, *Hc = buf + 3 * NB, *Gc = buf + 4 * NB, *G2c = buf + 5 * NB;\
    for (int j = 0; j < 6 * NB; ++j) buf[j] = black;                                                 \
    /* initB_ng: origin, boundary row (:71-84) */                                                    \
    Hp[0].val = 0; Hp[0].dir = G_NEWD; Hp[0].ptr = origin;                                           \
    {                                                                                                \
        const float ltg = al ? 1.f : (a->exgl ? 0.f : (float)p->tgapf);                              \
        int rr = br - al; if (up < rr) rr = up;                                                      \
        const int r0 = bl - al;                                                                      \
        for (int r = r0 + 1, i = 1; r <= rr; ++r, ++i) {                                             \
            VT gpn = (i == 1) ? ((1 > codonk1) ? LongGOP + 1 * LongGEP : BasicGOP + 1 * BasicGEP)    \
                              : ((i > codonk1) ? LongGEP : BasicGEP);                                \
            Hp[i].dir = G_HORI; Hp[i].ptr = origin;                                                  \
            Hp[i].val = Hp[i - 1].val + (VT)(gpn * ltg);                                             \
        }                                                                                            \
    }                                                                                                \
    const U topb = Hp[NB - 2];                                  /* boundary cell above the last column */\
    U leftb = black;                                            /* boundary cell left of the last row */\
    U *lastC = (U *)malloc(sizeof(U) * (size_t)(ar - al + 1));  /* cells of the last column, per row */  \
    for (int m = 0; m <= ar - al; ++m) lastC[m] = black;                                             \
    U colprev = Hp[0];                                                                               \
    int colk = 0;                                                                                    \
    const float ltgb = bl ? 1.f : (b->exgl ? 0.f : (float)p->tgapf);                                 \
    const int rr_col = (bl - ar > lw) ? bl - ar : lw;                                                \
    for (int m = al; m < ar; ++m) {                                                                  \
        const int n0 = G_MAX(m + lw, bl), n9 = G_MIN(m + up + 1, br);                                \
        const double *srow = mtx + (size_t)a->res[m] * dim;                                          \
        {   /* boundary column (:86-97) */                                                           \
            int r = bl - 1 - m;                                                                      \
            if (r >= rr_col) {                                                                       \
                ++colk;                                                                              \
                VT gpn = (colk == 1) ? ((1 > codonk1) ? LongGOP + 1 * LongGEP : BasicGOP + 1 * BasicGEP)\
                                     : ((colk > codonk1) ? LongGEP : BasicGEP);                      \
                U c; c.dir = G_VERT; c.ptr = origin; c.val = colprev.val + (VT)(gpn * ltgb);         \
                colprev = c; Hc[0] = c;                                                              \
            } else Hc[0] = black;                                                                    \
        }                                                                                            \
        U f1 = black, f2 = black;                                                                    \
        for (int n = n0; n < n9; ++n) {                                                              \
            const int j = n - bl + 1;                                                                \
            const int above_inband = (n - m + 1 <= up);                                              \
            const U habove = above_inband ? Hp[j] : black;                                           \
            const U gabove = (above_inband && m > al) ? Gp[j] : black;                               \
            const U g2above = (above_inband && m > al) ? G2p[j] : black;                             \
            const U hleft = (n - 1 >= n0 || n - 1 == bl - 1) ? Hc[j - 1] : black;                    \
            U h = Hp[j - 1], g, g2 = black;                                                          \
            int which = 0;     /* 0 diag, 1 g, 2 g2, 3 f1, 4 f2 */                                   \
            h.val += (VT)srow[b->res[n]];                      /* :181-183 */                        \
            h.dir = g_isdiag(Hp[j - 1].dir) ? G_DIAG : G_NEWD;                                       \
            VT mxv = h.val;                                                                          \
            VT x = habove.val + BasicGOP;                       /* vertical :186-193 */              \
            if (x >= gabove.val) { g.val = x; g.ptr = habove.ptr; g.dir = G_VERT; } else g = gabove; \
            g.val += BasicGEP;                                                                       \
            if (g.val > mxv) { which = 1; mxv = g.val; }                                             \
            if (Noll == 3) {                                    /* vertical2 :196-205 */             \
                x = habove.val + LongGOP;                                                            \
                if (x >= g2above.val) { g2.val = x; g2.ptr = habove.ptr; g2.dir = G_VERT; } else g2 = g2above;\
                g2.val += LongGEP;                                                                   \
                if (g2.val > mxv) { which = 2; mxv = g2.val; }                                       \
            }                                                                                        \
            x = hleft.val + BasicGOP;                           /* horizontal :207-214 */            \
            if (x >= f1.val) { f1.val = x; f1.ptr = hleft.ptr; f1.dir = G_HORI; }                    \
            f1.val += BasicGEP;                                                                      \
            if (f1.val >= mxv) { which = 3; mxv = f1.val; }                                          \
            if (Noll == 3) {                                    /* horizontal2 :217-226 */           \
                x = hleft.val + LongGOP;                                                             \
                if (x >= f2.val) { f2.val = x; f2.ptr = hleft.ptr; f2.dir = 9 /* HORL */; }          \
                f2.val += LongGEP;                                                                   \
                if (f2.val >= mxv) { which = 4; mxv = f2.val; }                                      \
            }                                                                                        \
            if (which == 1) h = g; else if (which == 2) h = g2; else if (which == 3) h = f1; else if (which == 4) h = f2;\
            if (h.dir == G_NEWD || h.dir == G_NEWV || h.dir == G_NEWH) h.ptr = o_add(&st, m, n, h.ptr);\
            Hc[j] = h; Gc[j] = g; G2c[j] = g2;                                                       \
            if (n == br - 1) lastC[m - al] = h;                                                      \
        }                                                                                            \
        if (m == ar - 1) leftb = Hc[0];                                                              \
        U *t;                                                                                        \
        t = Hp; Hp = Hc; Hc = t; t = Gp; Gp = Gc; Gc = t; t = G2p; G2p = G2c; G2c = t;               \
    }                                                                                                \
    /* lastB_ng (:100-143): trailing gaps at true sequence ends cost rtgapf times the penalty; the last       \
       column is relaxed downwards, then the last row rightwards, in place as the reference does */   \
    U *lastR = Hp + 1;                                          /* lastR[n - bl] = cell (ar-1, n) */  \
    int dm = 0, dn = 0;                                                                              \
    {                                                                                                \
        const float rtg = b->exgr ? 0.f : (float)p->tgapf;                                           \
        if (br == b->len && rtg < 1) {                                                               \
            const int rw = G_MIN(up, br - al);                                                       \
            U top = topb;                                                                            \
            for (int m = br - rw; m <= ar - 1; ++m) {            /* diagonals rw-1 .. br-ar */         \
                U *gq = (m == al) ? &top : &lastC[m - 1 - al];                                       \
                U *hq = &lastC[m - al];                                                              \
                ++dm;                                                                                \
                const VT gpn = !g_isvert(gq->dir) ? ((1 > codonk1) ? LongGOP + 1 * LongGEP : BasicGOP + 1 * BasicGEP)\
                                                  : ((dm > codonk1) ? LongGEP : BasicGEP);           \
                gq->val += (VT)(gpn * rtg);                                                          \
                if (gq->val > hq->val) { *hq = *gq; hq->dir = G_VERT; } else dm = 0;                 \
            }                                                                                        \
            lastR[br - 1 - bl] = lastC[ar - 1 - al];            /* the corner cell is shared */       \
        }                                                                                            \
    }                                                                                                \
    {                                                                                                \
        const float rtg = a->exgr ? 0.f : (float)p->tgapf;                                           \
        if (ar == a->len && rtg < 1) {                                                               \
            const int rw = G_MAX(lw, bl - ar);                                                       \
            U lft = leftb;                                                                           \
            for (int n = rw + ar; n <= br - 1; ++n) {            /* diagonals rw+1 .. br-ar */         \
                U *fq = (n == bl) ? &lft : &lastR[n - 1 - bl];                                       \
                U *hq = &lastR[n - bl];                                                              \
                ++dn;                                                                                \
                const VT gpn = !g_ishori(fq->dir) ? ((1 > codonk1) ? LongGOP + 1 * LongGEP : BasicGOP + 1 * BasicGEP)\
                                                  : ((dn > codonk1) ? LongGEP : BasicGEP);           \
                fq->val += (VT)(gpn * rtg);                                                          \
                if (fq->val > hq->val) { *hq = *fq; hq->dir = G_VERT; } else dn = 0;                 \
            }                                                                                        \
        }                                                                                            \
    }                                                                                                \
    U last = lastR[br - 1 - bl];                                                                     \
    if (dn || dm) { if (dn) dm = 0; last.ptr = o_add(&st, ar - dm, br - dn, last.ptr); }             \
    free(lastC);                                                                                     \
    long pp = o_add(&st, ar, br, last.ptr);                                                          \
    *score = (double)last.val;                                                                       \
    int cnt = 0, ok = 1;                                                                             \
    for (long q = pp;; q = st.v[q].p) {                         /* Vmf::traceback */                 \
        if (cnt + 2 >= cap) { ok = 0; break; }                                                       \
        out[++cnt].m = st.v[q].m; out[cnt].n = st.v[q].n;                                            \
        if (!st.v[q].p) break;                                                                       \
    }                                                                                                \
    if (ok && (out[cnt].m != al || out[cnt].n != bl)) { out[++cnt].m = al; out[cnt].n = bl; }        /* :1040-1044 */\
    out[0].m = 1; out[0].n = cnt;                                                                    \
    free(buf); free(st.v);                                                                           \
    return ok ? cnt : -1;                                                                            \
}

DEFINE_ALIGN_B1(float, f32, (-(FLT_MAX / 16 * 7)))
DEFINE_ALIGN_B1(double, f64, (-(DBL_MAX / 16 * 7)))

int orc_align_b1(const orc_seq *a, const orc_seq *b, const double *mtx, int dim, const orc_params *p,
                 double *score, orc_skl *out, int cap)
{
    return p->vtype ? align_b1_f64(a, b, mtx, dim, p, score, out, cap) : align_b1_f32(a, b, mtx, dim, p, score, out, cap);
}

DEFINE_GROUP_ALIGN(float, f32, (-(FLT_MAX / 16 * 7)))
DEFINE_GROUP_ALIGN(double, f64, (-(DBL_MAX / 16 * 7)))

int orc_align_groups(const orc_group *a, const orc_group *b, const double *mtx, int dim, const orc_gparams *p,
                     double *score, orc_skl *out, int cap, int64_t *cells)
{
    return p->vtype ? g_align_f64(a, b, mtx, dim, p, score, out, cap, cells, 0)
                    : g_align_f32(a, b, mtx, dim, p, score, out, cap, cells, 0);
}

/* swg1stC<SwgDPunit | _hf | _pf | _nv> (fwd2c.h:697-701) for algmode.mlt <= 1: colony 0 = best local score + box */
int orc_swg_groups(const orc_group *a, const orc_group *b, const double *mtx, int dim, const orc_gparams *p,
                   double *val, int *box, int64_t *cells)
{
    return p->vtype ? g_swg_f64(a, b, mtx, dim, p, val, box, cells) : g_swg_f32(a, b, mtx, dim, p, val, box, cells);
}

int orc_homscore_groups(const orc_group *a, const orc_group *b, const double *mtx, int dim, const orc_gparams *p,
                        double *score, long rr[2])
{
    return p->vtype ? g_align_f64(a, b, mtx, dim, p, score, 0, 0, 0, rr)
                    : g_align_f32(a, b, mtx, dim, p, score, 0, 0, 0, rr);
}
