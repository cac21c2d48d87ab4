// shim/shim_fwd2d1.cc -- the reference-side binding of INTEGRATION.md section 1, made real:
// a drop-in replacement of the reference's src/fwd2d1.o.  It defines the one symbol that object
// exports,
//     VTYPE alnScoreD(const Seq* seqs[], const Simmtx* sm, int* ends)        (src/fwd2d1.cc:324-337)
// and forwards it to libprrn_gpu.so (include/prrn_gpu.h: pg_score_pairs).  Linked IN FRONT OF the
// reference's own archive, every caller inside the unmodified reference -- alnscore2dist
// (src/aln2.cc:299,328), dpscore / calcdist (src/phyl.cc:221-342), AdjacentMat::spaln_job
// (src/adjmat.cc:133) -- reaches the CUDA kernels without a single source change.
//
// Compiled against the reference's headers where they lie (oracle/Makefile: target ref_gpu); no
// reference source is copied.  There is no CPU fallback: a library error is fatal(), the reference's
// own error convention.
#include "seq.h"
#include "aln.h"
#include "prrn_gpu.h"
#include "shim_ctx.h"

#include <string.h>
#include <chrono>

// PRRN_GPU_STATS=1: calls and seconds, printed to stderr at exit
struct PgScoreStats {
	long	n; double t; bool on; std::mutex mu;
	PgScoreStats() : n(0), t(0), on(getenv("PRRN_GPU_STATS") != 0) {}
	~PgScoreStats() {if (on) fprintf(stderr, "prrn_gpu alnScoreD: %ld calls, %.2f s in the library (%.1f us per call)\n", n, t, n? 1e6 * t / n: 0.);}
};
static PgScoreStats	pg_sstats;
static double	pg_snow() {return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();}

static void pg_fill_params(pg_params* p)
{
	p->alprm.u = alprm.u;   p->alprm.v = alprm.v;   p->alprm.u0 = alprm.u0; p->alprm.u1 = alprm.u1;
	p->alprm.v0 = alprm.v0; p->alprm.tgapf = alprm.tgapf; p->alprm.thr = alprm.thr;
	p->alprm.scale = alprm.scale; p->alprm.maxsp = alprm.maxsp; p->alprm.gamma = alprm.gamma;
	p->alprm.k1 = alprm.k1; p->alprm.ls = alprm.ls; p->alprm.sh = alprm.sh; p->alprm.mtx_no = alprm.mtx_no;
	p->lcl = algmode.lcl;
	p->vtype = sizeof(VTYPE) == sizeof(double);
}

VTYPE alnScoreD(const Seq* seqs[], const Simmtx* sm, int* ends)
{
	if (!sm) sm = getSimmtx(0);
const	Seq*	a = seqs[0];
const	Seq*	b = seqs[1];
	if (a->many != 1 || b->many != 1) fatal("prrn_gpu alnScoreD: single sequences only\n");
	// many == 1: at(0) is the contiguous residue array (seq.h)
	uint8_t*	res = new uint8_t[a->len + b->len + 1];
	memcpy(res, ((Seq*) a)->at(0), a->len);
	memcpy(res + a->len, ((Seq*) b)->at(0), b->len);
	int64_t	offs[2] = {0, a->len};
	int32_t	lens[2] = {a->len, b->len};
	int32_t	left[2] = {a->left, b->left};
	int32_t	right[2] = {a->right, b->right};
	uint8_t	exg[2] = {uint8_t((a->inex.exgl? 1: 0) | (a->inex.exgr? 2: 0)),
			  uint8_t((b->inex.exgl? 1: 0) | (b->inex.exgr? 2: 0))};
	pg_seqs	S = {res, offs, lens, left, right, exg, 2};
	pg_params	P;
	pg_fill_params(&P);
	int32_t	ia = 0, ib = 1;
	VTYPE	scr = 0;
	// Simmtx::mtx is dim row pointers into one contiguous dim*dim block; flatten defensively
	VTYPE*	flat = new VTYPE[sm->dim * sm->dim];
	for (int i = 0; i < sm->dim; ++i)
	    for (int j = 0; j < sm->dim; ++j) flat[i * sm->dim + j] = sm->mtx[i][j];
	const double	t0 = pg_sstats.on? pg_snow(): 0;
	PgLease	ctx;		// the reference calls this concurrently from pthread workers: one pooled context per call in flight
	int	rc = pg_score_pairs(ctx, &S, &ia, &ib, 1, &P, flat, sm->dim, &scr,
		    (algmode.lcl & 16)? 0: ends);
	delete[] flat;
	delete[] res;
	if (rc != PG_OK) fatal("prrn_gpu alnScoreD: %s\n", pg_last_error(ctx));
	if (pg_sstats.on) {std::lock_guard<std::mutex> lk(pg_sstats.mu); ++pg_sstats.n; pg_sstats.t += pg_snow() - t0;}
	return (scr);
}
