/* include/prrn_gpu.h -- C ABI of libprrn_gpu.so: the B200 (sm_100a) implementation of the
 * prrn_aln dynamic-programming forward fill.
 *
 * Plain pointers and sizes only; no C++ or torch types cross this line.  Every entry point names
 * the reference interface it stands behind (file:line into ogotoh/prrn_aln `src/`).  The reference
 * has no FFI of its own: these are the calls a maintainer binds from the C++ drivers (see
 * INTEGRATION.md for the shim that re-points alnScoreD / calcdist / alignC at this library).
 *
 * There is NO CPU fallback: every compute call fails with PG_ERR_NO_DEVICE / PG_ERR_CUDA when the
 * CUDA device or kernels are unavailable, and with PG_ERR_UNSUPPORTED for modes not built yet.
 */
#ifndef PRRN_GPU_H
#define PRRN_GPU_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PG_OK               0
#define PG_ERR_NO_DEVICE    1   /* no CUDA device / driver                                          */
#define PG_ERR_CUDA         2   /* a CUDA runtime call or kernel failed (pg_last_error has text)   */
#define PG_ERR_ARG          3   /* invalid argument                                                */
#define PG_ERR_UNSUPPORTED  4   /* mode exists in the reference but is not implemented here yet    */
#define PG_ERR_RANGE        5   /* scores would overflow the integer kernels' range                */

typedef struct pg_context pg_context;

/* Mirror of the reference's ALPRM (src/seq.h:27-28), field for field. */
typedef struct {
    float u, v, u0, u1, v0, tgapf, thr, scale, maxsp, gamma;
    int32_t k1, ls, sh, mtx_no;
} pg_alprm;

/* The scalars the path reads besides ALPRM. */
typedef struct {
    pg_alprm alprm;
    int32_t lcl;        /* algmode.lcl (src/clib.h:37-55): 0 global; bits 1,2,4,8 free ends; 16 SWG   */
    int32_t vtype;      /* VTYPE of the caller's build: 0 float (aln), 1 double (prrn, -DDVAL=1)      */
} pg_params;

/* A set of sequences as the reference's Seq objects present them to the DP (src/seq.h:253-438):
 * residue codes (1 byte/residue, src/cmn.h:110-112), concatenated; sequence i occupies
 * res[offs[i] .. offs[i]+lens[i]).  left/right = operated window (NULL: 0 / len);
 * exg[i] bit0 = inex.exgl, bit1 = inex.exgr (NULL: 0). */
typedef struct {
    const uint8_t *res;
    const int64_t *offs;
    const int32_t *lens;
    const int32_t *left;
    const int32_t *right;
    const uint8_t *exg;
    int32_t nseq;
} pg_seqs;

/* ---- context ------------------------------------------------------------------------------- */
int  pg_create(int device, pg_context **out);
void pg_destroy(pg_context *ctx);
const char *pg_last_error(const pg_context *ctx);   /* valid with ctx == NULL after pg_create failed */
const char *pg_version(void);

/* ---- per-call level: stands behind
 *   VTYPE alnScoreD(const Seq* seqs[2], const Simmtx* sm, int* ends)          src/fwd2d1.cc:324-337
 * for a batch of (a, b) pairs.  mtx = Simmtx::mtx flattened dim x dim in the caller's VTYPE
 * (float when vtype == 0, double when 1).  out_scores[p] receives the score in that VTYPE.
 * out_ends (nullable, 2 ints per pair) is the `ends` output of the semi-global variant. */
int pg_score_pairs(pg_context *ctx, const pg_seqs *seqs, const int32_t *a_idx, const int32_t *b_idx,
                   int64_t npairs, const pg_params *prm, const void *mtx, int32_t dim,
                   void *out_scores, int32_t *out_ends);

/* ---- edge-list level: stands behind the DynScr branch of
 *   void AdjacentMat::spaln_job(Seq* sqs[], SrchBlk* bks, AdjMatThQueue* q)   src/adjmat.cc:119-156
 * i.e. `dist = alnscore2dist(sqs, bks->pwd); dist *= 100.` (src/aln2.cc:289-334) for the candidate pairs
 * (query a, database sequence b) the k-mer search returns -- the sparse counterpart of calcdist: the same
 * fill kernels and the same distance epilogue (self scores, |length difference|, algmode.lcl branch with
 * exg_seq and end points) over an explicit pair list instead of the condensed triangle.  out_dist[p] in
 * FTYPE (float when vtype == 0, double when 1), bit-identical to pg_calcdist's entry for the same (a, b).
 * The threshold test (`< alprm.thr`) and fillmat stay with the caller. */
int pg_dist_pairs(pg_context *ctx, const pg_seqs *seqs, const int32_t *a_idx, const int32_t *b_idx,
                  int64_t npairs, const pg_params *prm, const void *mtx, int32_t dim, void *out_dist);

/* ---- per-call level: stands behind
 *   template<class recd_t> SKL* alignC(mSeq* seqs[2], PwdM* pwd, VTYPE* scr, ...)   src/fwd2c.h:670-677
 * for recd_t = DPunit (alnmode NGP_ALB: two single sequences, no internal gaps), i.e. the
 * Fwd2c<DPunit> ctor + initB + forwardB + Vmf::traceback chain (src/fwd2c.h:81-176,359-482,
 * src/vmf.cc:103-119) that align2 (src/maln2.cc:1888-1910) reaches, for a batch of (a, b) pairs.
 * out_scores[p] = *scr in the caller's VTYPE.  The corner lists come back concatenated:
 * (*out_pts)[(*out_offs)[p] .. (*out_offs)[p+1]) are skl[1..n] of pair p in Vmf back-walk order,
 * exactly what alignC returns (the caller runs stdskl, src/gaps.cc:139, as align2 does).
 * Both arrays are allocated by the library and released with pg_free. */
typedef struct { int32_t m, n; } pg_skl;        /* SKL of src/cmn.h:124 */
int pg_align_pairs(pg_context *ctx, const pg_seqs *seqs, const int32_t *a_idx, const int32_t *b_idx,
                   int64_t npairs, const pg_params *prm, const void *mtx, int32_t dim,
                   void *out_scores, int64_t **out_offs, pg_skl **out_pts);
void pg_free(void *p);

/* ---- per-call level: stands behind
 *   SKL*  alignB_ng(const Seq* seqs[2], const PwdB* pwd, VTYPE* scr)                src/fwd2b1.cc:1347
 *   VTYPE HomScoreB_ng(const Seq* seqs[2], const PwdB* pwd, long rr[])              src/fwd2b1.cc:1317
 * (Aln2b1::initB_ng :64, forwardB_ng :145, lastB_ng :100, trcbkalignB_ng :1025, globalB_ng :1286) for a
 * batch of pairs: prrn5's DynAln distances (src/adjmat.cc:84).  Global and semi-global: alprm.tgapf and the
 * inex.exgl / exgr flags of each sequence (seqs->exg, or prm->lcl & 15 applied as aln does: bits 1, 2 to a,
 * 4, 8 to b) scale the leading gaps in initB_ng and drive the relaxation of lastB_ng.  prm->lcl & 16
 * (fwdswgB_ng) returns PG_ERR_UNSUPPORTED.  DP matrices of any size are traced directly (no linear-space
 * recursion), affine or two-piece.  Same outputs as pg_align_pairs: the score (== HomScoreB_ng's) and the
 * corner records before stdskl, which the caller runs as globalB_ng does. */
int pg_align_pairs_ng(pg_context *ctx, const pg_seqs *seqs, const int32_t *a_idx, const int32_t *b_idx,
                      int64_t npairs, const pg_params *prm, const void *mtx, int32_t dim,
                      void *out_scores, int64_t **out_offs, pg_skl **out_pts);

/* ---- per-call level, groups: stands behind
 *   template<class recd_t> SKL* alignC(mSeq* seqs[2], PwdM* pwd, VTYPE* scr, ...)   src/fwd2c.h:670-677
 * for recd_t = DPunit (NGP_ALB, groups without internal gaps), DPunit_hf (HLF_ALB / RHF_ALB),
 * DPunit_pf (GPF_ALB) and DPunit_nv (NTV_ALB), as align2 dispatches it (src/maln2.cc:1899-1910), affine or two-piece, for a
 * BATCH of independent (a, b) pairs -- e.g. the candidate partitions of Prrn::best_of_n
 * (src/prrn5.cc:594-631).  alnmode NGP_ALN (algmode.bnd = 0 on groups without gap profile: align2 calls
 * alignC<DPunit>(seqs, pwd, scr, true), src/maln2.cc:1906) runs the rectangle form, forwardA + initA (src/fwd2c.h:
 * 111-135,231-356): every cell of the window, no band; it needs b.left = 0 and b's per-column arrays ONE COLUMN
 * LONGER (npos + 1 entries: forwardA reads b's thickness at position b.right, :240-249).  The rectangle forms with
 * gap profiles (HLF_ALN / RHF_ALN / GPF_ALN / NTV_ALN) return PG_ERR_UNSUPPORTED: forwardA copies its records by struct
 * assignment (`*hdiag = *h`, :124,247), which for the record types with list pointers aliases the gap states of different
 * cells -- the reference's result there is not a function of the inputs alone.  So does the rectangle HomScoreC.
 * Precondition as in the reference: PwdM pwd(seqs) already ran (sequences
 * swapped if pwd->swp, mkthick / Gfq / convseq done).  A pg_group is what Fwd2c reads from one mSeq
 * through mSeqItr for the columns left-1 .. right-1 (npos = right - left + 1 entries):
 *   cfq, efq   SeqThk::cfq / efq per column                         src/mseq.h:70-74
 *   vec        this side's factor of sim2: S(m,n) = vec_a[m] . vec_b[n] over kdim residue codes; every
 *              sim11 .. sim33 variant (src/maln2.cc:534-623,1230-1296) is such a product (profile part
 *              of vss x frequency vector; INTEGRATION.md gives the rule per a_mode / b_mode)
 *   glen, gfreq, sfq / tfq / rfq   the column's GFREQ lists (Gfq::sfrq / tfrq / rfrq, src/gfreq.h:44-55)
 *              pooled: list of column x starts at glen[sfq[x]], ends at the first glen < 0; -1 = empty
 * Arithmetic is double for either VTYPE flavour; scores agree with the reference within 1e-5 relative
 * and corner lists are identical except at near-ties (BASELINE north_star).  Output as pg_align_pairs:
 * (*out_pts)[(*out_offs)[p] .. (*out_offs)[p+1]) = skl[1..n] in Vmf back-walk order; release with
 * pg_free. */
typedef struct {
    int32_t many, len, left, right;
    int32_t hetero;             /* Gfq::hetero, -1 without gap profile                               */
    int32_t nils;               /* inex.nils                                                          */
    const double *cfq, *efq;
    const double *vec;          /* [npos][kdim]                                                       */
    const int32_t *glen;
    const double *gfreq;
    int32_t npool;
    const int32_t *sfq, *tfq, *rfq;     /* [npos] or NULL                                             */
    /* NTV_ALB only (DPunit_nv: small groups of raw residues, no profile; crg11 .. crg22w, src/maln2.cc:
     * 881-1024,1454-1614): per column, bit i set = member i holds a gap (IsGap, src/seq.h:219); member
     * weights (1.0 each when the mSeq has none).  many <= 32, no nil ends.  NULL otherwise.             */
    const uint32_t *gapmask;            /* [npos]                                                     */
    const double *weight;               /* [many]                                                     */
} pg_group;

typedef struct {
    int32_t alnmode;            /* ALN_MODE, src/aln.h:71-76: 1 NGP_ALN, 6 NGP_ALB, 7 HLF_ALB, 8 RHF_ALB, 9 GPF_ALB, 10 NTV_ALB
                                   (100 is internal: the Aln2b1 recurrence behind pg_align_pairs_ng)   */
    int32_t Noll, codonk1;      /* PwdB::Noll, PwdB::codonk1 (src/aln2.cc:100,117)                     */
    int32_t sh;                 /* pwd->alnprm.sh                                                      */
    int32_t kdim;
    double u;                   /* pwd->alnprm.u                                                       */
    double Weighted_GOP, Basic_GOP;                 /* PwdM::resetuab, src/maln2.cc:227-243           */
    double BasicGOP, BasicGEP, LongGOP, LongGEP;    /* PwdB::PwdB, src/aln2.cc:103-108                */
} pg_gparams;

int pg_align_groups(pg_context *ctx, const pg_group *a, const pg_group *b, const pg_gparams *prm, int64_t npairs,
                    double *out_scores, int64_t **out_offs, pg_skl **out_pts);
/* ---- per-call level, groups, score only: stands behind
 *   template<class recd_t> VTYPE HomScoreC(mSeq* seqs[2], PwdM* pwd, long rr[2], ...)   src/fwd2c.h:663-668
 * as HomScore dispatches it (src/maln2.cc:1853-1857), i.e. Fwd2c without Vmf + forwardB(rr).  Same inputs as
 * pg_align_groups.  out_rr (nullable, 2 per pair) receives pp[] of forwardB (src/fwd2c.h:476-479): rr[0] = the
 * diagonal n - m of the last first-row cell on the optimal path (b.left - a.left if it never enters the first
 * row; :468-469), rr[1] = the end diagonal. */
int pg_score_groups(pg_context *ctx, const pg_group *a, const pg_group *b, const pg_gparams *prm, int64_t npairs,
                    double *out_scores, int64_t *out_rr);
/* ---- per-call level, groups, Smith-Waterman: stands behind
 *   template<class recd_t> Colonies* swg1stC(mSeq* seqs[2], PwdM* pwd, WINDOW* pwdw = 0)   src/fwd2c.h:697-701
 * for recd_t = SwgDPunit | SwgDPunit_hf | SwgDPunit_pf | SwgDPunit_nv as swg1st dispatches it (src/maln2.cc:1975-2010;
 * callers: aln.cc:287-311 under algmode.lcl & 16), i.e. Fwd2c::initC + forwardC (src/fwd2c.h:178-207,483-659), for
 * algmode.mlt <= 1: no secondary colonies, the result is colony 0 = the best local score and the box of its path.
 * Same inputs as pg_align_groups (alnmode 6 .. 10).  out_val[p] = COLONY::val, out_box[6 p ..] = mlb nlb mrb nrb lwr upr
 * (src/aln.h:150-160).  The second pass (swg2ndC: align2 inside the box) is pg_align_groups on that window.
 * algmode.mlt > 1 (several colonies, with their sequential garbage collection) is not built. */
int pg_local_groups(pg_context *ctx, const pg_group *a, const pg_group *b, const pg_gparams *prm, int64_t npairs,
                    double *out_val, int32_t *out_box);
/* DP cells the reference visits for one group pair (band from stripe(), src/aln2.cc:156-174). */
int64_t pg_group_cells(const pg_group *a, const pg_group *b, int32_t sh);

/* ---- batch level: stands behind
 *   FTYPE* calcdist(mSeq** sbuf, int nn, DistCal realign = DynScr)            src/phyl.cc:318-342
 * (selfscr :253-261, dpscore :221-251, alnscore2dist src/aln2.cc:289-334) for single sequences.
 * Fills out_dist[k - k_begin] for the condensed indices k in [k_begin, k_end), k = elem(i,j) =
 * j(j-1)/2 + i, i < j (src/cmn.h:115), a = sequence i, b = sequence j; values 100*(1 - ...)
 * in FTYPE (float when vtype == 0, double when 1).  [k_begin, k_end) is the shard of one rank. */
int pg_calcdist(pg_context *ctx, const pg_seqs *seqs, const pg_params *prm, const void *mtx,
                int32_t dim, int64_t k_begin, int64_t k_end, void *out_dist);

/* ---- device-resident variants (inputs already in HBM; used for sharded multi-GPU runs) ------ */
typedef struct pg_dev_seqs pg_dev_seqs;
int  pg_seqs_upload(pg_context *ctx, const pg_seqs *seqs, pg_dev_seqs **out);
void pg_seqs_free(pg_context *ctx, pg_dev_seqs *d);
/* d_out_dist: DEVICE pointer to (k_end - k_begin) FTYPE values; stream: cudaStream_t or NULL.
 * Asynchronous with respect to the host; *n_launches (nullable) receives the number of kernels.  The context's
 * workspace is shared by all calls: the next call on this context -- on any stream -- is ordered on the device after
 * the work queued here, so calls never overlap; keep `stream` alive until it has been synchronized. */
int pg_calcdist_dev(pg_context *ctx, pg_dev_seqs *seqs, const pg_params *prm, const void *mtx,
                    int32_t dim, int64_t k_begin, int64_t k_end, void *d_out_dist, void *stream,
                    int32_t *n_launches);

/* ---- measurement helpers ------------------------------------------------------------------- */
/* Number of DP cells the reference's loops visit for the pairs k in [k_begin, k_end)
 * (rows m in window, columns in the stripe() band; src/aln2.cc:156-174). */
int64_t pg_calcdist_cells(const pg_seqs *seqs, const pg_params *prm, int64_t k_begin, int64_t k_end);
/* Host-only: the schedule of the packed score kernel for a condensed range -- number of work items,
 * number of (query pair, subject) slots, and per-k coverage counts (each must be 1).  Test aid. */
int pg_debug_packed_plan(const pg_seqs *seqs, int64_t k_begin, int64_t k_end, int32_t grid_blocks,
                         int64_t *nitems, int64_t *nslots, uint8_t *cover);
/* Device time (CUDA events on the library's stream) of the fill kernel(s) of the last pg_align_groups
 * call, in milliseconds; -1 if none.  Measurement aid for bench scripts. */
double pg_last_kernel_ms(pg_context *ctx);
/* Register-only DPX micro-benchmark: measured issue rate of __viaddmax_s32 / __vimax3_s32 chains,
 * in 1e9 thread-instructions per second for the whole device; the roofline denominator. */
int pg_dpx_peak(pg_context *ctx, double *gops_s32, double *gops_s16x2);

#ifdef __cplusplus
}
#endif
#endif
