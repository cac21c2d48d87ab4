/* oracle/oracle.h -- CPU restatement of the prrn_aln DP hot path.  TEST INFRASTRUCTURE ONLY.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load
 * liboracle.so; the product (libprrn_gpu.so) never does and has no CPU fallback.
 *
 * Parity status: PINNED.  Every function here is checked (tests/test_oracle_vs_reference.py, golden
 * vectors under tests/golden/ produced by tools/make_golden.py from the unmodified reference built
 * by oracle/Makefile) against the reference's own outputs: alnScoreD scores, calcdist distance
 * vectors, align2 scores + stdskl corner lists.
 *
 * The restatement is written in plain row-major (m, n) order with an explicit band test; the
 * reference scans anti-diagonals in place (fwd2d1.cc:136-160) or rows in diagonal coordinates
 * (fwd2c.h:359-482).  Each function cites the reference lines it follows.
 */
#ifndef PRRN_ORACLE_H
#define PRRN_ORACLE_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* A view of one sequence as the reference's Seq presents it to the DP (seq.h:253-438):
 * residue codes res[0..len), operated window [left, right), end-gap flags inex.exgl / inex.exgr. */
typedef struct {
    const uint8_t *res;
    int32_t len, left, right;
    int32_t exgl, exgr;
} orc_seq;

/* Scalars of ALPRM (seq.h:27-28) + algmode.lcl that the path reads. */
typedef struct {
    double u, v, scale, tgapf;  /* alprm.u, .v, .scale, .tgapf   (float in the reference)          */
    double u1;                  /* alprm.u1   long-gap extension                                     */
    int32_t k1, ls, sh;         /* alprm.k1, .ls, .sh                                                */
    int32_t lcl;                /* algmode.lcl                                                       */
    int32_t vtype;              /* 0: float VTYPE (aln build), 1: double VTYPE (prrn build)          */
} orc_params;

typedef struct { int32_t lw, up, width; } orc_window;

/* stripe(): reference src/aln2.cc:156-174 */
void orc_stripe(const orc_seq *a, const orc_seq *b, int sh, orc_window *w);

/* number of (m,n) cells the reference loops visit for this pair (SURVEY.md 8(d) cell definition;
 * loop bounds fwd2d1.cc:137-146 == fwd2c.h:364-374) */
int64_t orc_band_cells(const orc_seq *a, const orc_seq *b, int sh);

/* alnScoreD(), global / semi-global branch without `ends` (Fwd2d ctor + forwardD + lastD):
 * reference src/fwd2d1.cc:57-90, 136-160, 97-134, 324-337.  mtx is dim x dim row-major in the
 * VTYPE of p->vtype (passed as double; values are rounded to float when vtype == 0). */
double orc_aln_score_d(const orc_seq *a, const orc_seq *b, const double *mtx, int dim,
                       const orc_params *p);

/* alnScoreD() with its full dispatch (fwd2d1.cc:324-337): algmode.lcl & 16 -> swgforwardD
 * (:162-189); ends != NULL -> Fwd2d_vd (:212-322), ends[0] = diagonal offset of the path's start,
 * ends[1] = trailing overhang (dn or -dm); else the plain (semi-)global score above. */
double orc_aln_score_full(const orc_seq *a, const orc_seq *b, const double *mtx, int dim,
                          const orc_params *p, int *ends);

/* selfAlnScr(): reference src/aln2.cc:54-64 (many == 1) */
double orc_self_score(const orc_seq *a, const double *mtx, int dim, const orc_params *p);

/* alnscore2dist() global branch + dpscore()'s denominator and x100:
 * reference src/aln2.cc:289-334 (else-branch :321-333), src/phyl.cc:221-251.
 * self_a/self_b are selfscr() values (phyl.cc:253-261, sumwt == 1 for single sequences). */
double orc_score2dist(double scr, int la, int lb, double self_a, double self_b, const orc_params *p);

/* alnscore2dist(), algmode.lcl != 0 branch (aln2.cc:296-320) x 100 (phyl.cc:249): exg_seq flags from
 * lcl bits, alnScoreD with ends, denominator from the trimmed windows' self scores.  Undefined in the
 * reference for lcl & 16 (ends is read uninitialised); do not call it so. */
double orc_score2dist_lcl(const orc_seq *a, const orc_seq *b, const double *mtx, int dim,
                          const orc_params *p, double *raw, int *ends_out);

/* calcdist(seqs, nn, DynScr) for single sequences: reference src/phyl.cc:318-342.
 * dist[elem(i,j)], elem(i,j) = j(j-1)/2 + i for i < j (cmn.h:115); a = seq i, b = seq j.
 * raw_scores (optional) receives alnScoreD per pair in the same order. */
void orc_calcdist(const orc_seq *seqs, int nn, const double *mtx, int dim, const orc_params *p,
                  double *dist, double *raw_scores);

/* ---- pairwise alignment with path (NGP: two single, ungapped sequences) ------------------------ */
typedef struct { int32_t m, n; } orc_skl;

/* alignC<DPunit>() = Fwd2c<DPunit>::Fwd2c + initB + forwardB + Vmf::traceback for two single
 * sequences without internal gaps and without nil ends (global mode, tgapf == 1, thickness == 1):
 * reference src/fwd2c.h:81-100,138-176,359-482,670-677; src/fwd2c.cc:32-102; src/vmf.cc:103-119.
 * Affine (ls < 3) or two-piece (ls == 3, src/fwd2c.h:411-442).  Writes the corner list in Vmf
 * back-walk order into out[0..cap) exactly as alignC returns it (out[0].n = count) and returns the
 * number of corners, or -1 if cap is too small.  *score receives the DP score. */
int orc_align_ngp(const orc_seq *a, const orc_seq *b, const double *mtx, int dim, const orc_params *p,
                  double *score, orc_skl *out, int cap);

/* stdskl(): reference src/gaps.cc:139-175.  skl[0].n = count; normalises in place into out
 * (capacity 2*count+2); returns the new count. */
int orc_stdskl(const orc_skl *skl, orc_skl *out);

/* ---- group-to-group alignment with path (oracle_grp.c) ----------------------------------------- */
typedef struct { int32_t glen; double freq; int32_t nres; } orc_gfreq;     /* GFREQ, gfreq.h:25 */

/* One group (mSeq) as Fwd2c reads it through mSeqItr after PwdM's staging, for the columns
 * left-1 .. right-1 (npos = right - left + 1 entries, entry x <-> sequence position left-1+x). */
typedef struct {
    int32_t many, len, left, right;
    int32_t nelm, felm, hetero, nils;       /* mSeq::nelm / felm, Gfq::hetero, inex.nils */
    const double *cfq, *dfq, *efq;          /* SeqThk per column (mseq.h:70-74) */
    const uint8_t *res;                     /* [npos][many] residue codes */
    const double *vss;                      /* [npos][nelm] frequency + profile vector, or NULL */
    const double *weight;                   /* [many] sequence weights, or NULL */
    const orc_gfreq *gpool;                 /* gap-profile lists, each terminated by glen < 0 */
    const int32_t *sfq, *tfq, *rfq;         /* [npos] offsets into gpool (-1: none), or NULL */
} orc_group;

typedef struct {
    int32_t alnmode;            /* ALN_MODE (aln.h:71-76): 6 NGP_ALB, 7 HLF_ALB, 8 RHF_ALB, 9 GPF_ALB */
    int32_t a_mode, b_mode;     /* 0 single, 1 group of residues, 2 profile (maln2.cc:283-284) */
    int32_t Noll, codonk1, sh;
    int32_t vtype, dxd;         /* VTYPE flavour; DvsP == DxD (sim33_n) */
    double u;                   /* alnprm.u */
    double Weighted_GOP, Basic_GOP;                 /* PwdM::resetuab (maln2.cc:227-243) */
    double BasicGOP, BasicGEP, LongGOP, LongGEP;    /* PwdB::PwdB (aln2.cc:97-117) */
} orc_gparams;

/* alignC<DPunit | DPunit_hf | DPunit_pf>(seqs, pwd, &scr): corner list in Vmf back-walk order
 * (out[0].n = count), *score = DP score, *cells = cells visited.  Returns the count or -1. */
int orc_align_groups(const orc_group *a, const orc_group *b, const double *mtx, int dim, const orc_gparams *p,
                     double *score, orc_skl *out, int cap, int64_t *cells);
/* HomScoreC<recd_t>(seqs, pwd, rr) (reference src/fwd2c.h:663-668): the same fill without Vmf; rr = pp[] of forwardB
 * (:468-469, :476-479). */
int orc_swg_groups(const orc_group *a, const orc_group *b, const double *mtx, int dim, const orc_gparams *p,
                   double *val, int *box /* mlb nlb mrb nrb lwr upr */, int64_t *cells);
int orc_homscore_groups(const orc_group *a, const orc_group *b, const double *mtx, int dim, const orc_gparams *p,
                        double *score, long rr[2]);


/* Aln2b1: trcbkalignB_ng inside globalB_ng (reference src/fwd2b1.cc:64-279,1025-1051,1286-1315) for two
 * single sequences, global mode, tgapf == 1: the corner records before stdskl (out[0].n = count) and
 * the forwardB_ng score (== HomScoreB_ng, :1317).  Returns the count, -1 overflow, -2 unsupported mode. */
int orc_align_b1(const orc_seq *a, const orc_seq *b, const double *mtx, int dim, const orc_params *p,
                 double *score, orc_skl *out, int cap);

#ifdef __cplusplus
}
#endif
#endif
