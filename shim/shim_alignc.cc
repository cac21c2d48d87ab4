// shim/shim_alignc.cc -- the reference-side binding of INTEGRATION.md section 3, made real.
//
// The reference instantiates
//     template<class recd_t> SKL*  alignC(mSeq* seqs[], PwdM*, VTYPE*, bool, WINDOW*)       (src/fwd2c.h:670-677)
//     template<class recd_t> VTYPE HomScoreC(mSeq* seqs[], PwdM*, long rr[], bool, WINDOW*) (src/fwd2c.h:663-668)
// inside maln2.o as WEAK out-of-line functions; align2 / HomScore (src/maln2.cc:1837-1910) and prrn5's refinement
// reach the DP through them.  This file provides STRONG explicit specialisations for recd_t = DPunit, DPunit_hf,
// DPunit_pf, DPunit_nv, so that once it is linked with the unmodified reference objects every one of those calls
// lands here: the groups are staged exactly as Fwd2c would read them (mSeqItr over left-1 .. right-1 after PwdM's
// own convseq / mkthick / Gfq) and the banded fill runs in libprrn_gpu.so:
//     groups, profiles, weighted members      pg_align_groups / pg_score_groups   (kernels K4 + K3)
//     two plain single sequences (NGP_ALB)    pg_align_pairs                      (kernel K2, exact integers on DPX
//                                             when matrix and penalties are integral)
//
// BATCHING (Prrn::best_of_n, src/prrn5.cc:594-631).  With `prrn5 -tB` the reference evaluates B candidate
// partitions of one MSA state on B pthread workers (thread_onecycle, :565-592); each worker reaches alignC on its
// own.  The shim turns those concurrent calls into ONE library call without touching prrn5.cc: a worker parks its
// staged pair in a rendezvous; when every live worker thread of the program is parked there (the shim counts the
// workers by interposing pthread_create for start routines that live in the executable) the last arrival launches
// the whole batch, and every worker returns with its own result.  Threads that never reach the DP (empty
// partitions, nogap_skl) simply end and are no longer waited for; a time-out (PRRN_GPU_BATCH_WAIT_US, default
// 2,000) covers worker pools whose idle members block elsewhere (CalcServer).  PRRN_GPU_BATCH=0 turns it off.
// Results do not depend on how calls were grouped: every alignment of a batch is computed independently.
//
// The rectangle form (algmode.bnd = 0) is taken for NGP_ALN: alignC<DPunit>(seqs, pwd, scr, true), forwardA + initA.
// Smith-Waterman (algmode.lcl & 16): swg1st (first pass, Fwd2c::forwardC) runs pg_local_groups for algmode.mlt <= 1.
// Calls the CUDA path does not take (rectangle with gap profiles, caller-supplied window, secondary colonies, quick mode on
// gapped groups, naive groups with nil
// ends or more than 32 members) are fatal() unless PRRN_GPU_ALLOW_REF=1 (shim_ctx.h); then they run the reference's
// own Fwd2c -- the reference's code, not a port; nothing here re-implements the DP on the CPU.
#ifndef _GNU_SOURCE
#define _GNU_SOURCE
#endif
#include "aln.h"
#include "mseq.h"
#include "maln.h"
#include "mgaps.h"
#include "gfreq.h"
#include "vmf.h"
#include "fwd2c.h"
#include "fwd2h.h"
#include "fwd2s.h"
#include "prrn_gpu.h"
#include "shim_ctx.h"

#include <dlfcn.h>
#include <pthread.h>

#include <math.h>

#include <chrono>
#include <condition_variable>
#include <cstdlib>
#include <string>
#include <vector>

// PRRN_GPU_STATS=1: calls / seconds per route, printed to stderr at exit (where does a prrn run spend its time?)
struct PgStats {
	long	n_gpu, n_k2, n_score, n_local, n_ref, n_batches, max_batch; double t_stage, t_gpu, t_ref, kernel_ms; long cells;
	bool	on;
	std::mutex	mu;
	PgStats() : n_gpu(0), n_k2(0), n_score(0), n_local(0), n_ref(0), n_batches(0), max_batch(0), t_stage(0), t_gpu(0), t_ref(0),
	    kernel_ms(0), cells(0), on(getenv("PRRN_GPU_STATS") != 0) {}
	~PgStats() {
	    if (on && n_local) fprintf(stderr, "prrn_gpu swg1st: %ld Smith-Waterman first passes on the GPU\n", n_local);
	    if (on) fprintf(stderr, "prrn_gpu alignC: %ld calls on the GPU (%ld of them score-only, %ld pairs of single sequences "
		"on K2) in %ld library calls (mean batch %.2f, largest %ld; staging %.2f s, library %.2f s of which kernels "
		"%.2f s, %.3g cells), %d contexts, %ld calls left on the reference's Fwd2c (%.2f s)\n",
		n_gpu, n_score, n_k2, n_batches, n_batches? (double) n_gpu / n_batches: 0., max_batch, t_stage, t_gpu,
		kernel_ms * 1e-3, (double) cells, pg_ctx_created(), n_ref, t_ref);
	}
};
static PgStats	pg_stats;
static double	pg_now() {return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();}

struct PgSide {
	std::vector<double>	cfq, efq, vec, gfreq, weight;
	std::vector<int32_t>	glen, sfq, tfq, rfq;
	std::vector<uint32_t>	gapmask;
	int32_t	pool(GFREQ** pp) {
	    if (!pp || !*pp) return (-1);
	    int32_t	at = (int32_t) glen.size();
	    for (const GFREQ* g = *pp; neogfq(g); ++g) {glen.push_back(g->glen); gfreq.push_back(g->freq);}
	    glen.push_back(-1); gfreq.push_back(0);
	    return (at);
	}
};

static const int pg_decompact[6] = {nil_code, gap_code, A, C, G, T};	// src/mseq.h:38

// one vector per column such that sim2(m, n) = vec_a[m] . vec_b[n]	(src/maln2.cc:534-623, 1230-1296)
// extra: columns staged beyond right - 1 (the rectangle form reads b at position right, src/fwd2c.h:240-249)
static void pg_stage(mSeq* sd, PgSide& S, const PwdM* pwd, bool is_a, const Simmtx* sm, int extra = 0)
{
	const int	npos = sd->right - sd->left + 1 + extra, K = sm->dim;
	const int	my_mode = is_a? pwd->a_mode: pwd->b_mode;
	const int	ot_mode = is_a? pwd->b_mode: pwd->a_mode;
	const bool	dxd = pwd->DvsP == 0;
	S.cfq.resize(npos); S.efq.resize(npos); S.vec.assign((size_t) npos * K, 0.);
	S.sfq.assign(npos, -1); S.tfq.assign(npos, -1); S.rfq.assign(npos, -1);
	const bool	naive = pwd->alnmode == NTV_ALB && sd->many <= 32;	// DPunit_nv: IsGap bits + member weights
	if (naive) {
	    S.gapmask.assign(npos, 0);
	    S.weight.assign(sd->many, 1.);
	    if (sd->weight) for (int i = 0; i < sd->many; ++i) S.weight[i] = sd->weight[i];
	}
	for (int x = 0, p = sd->left - 1; x < npos; ++x, ++p) {
	    mSeqItr	it(sd, p);
	    S.cfq[x] = it.dns->cfq; S.efq[x] = it.dns->efq;
	    if (naive) for (int i = 0; i < sd->many; ++i) if (IsGap(it.res[i])) S.gapmask[x] |= 1u << i;
	    double*	v = &S.vec[(size_t) x * K];
	    bool	profile_part = my_mode == 2 && (is_a || ot_mode != 2);
	    bool	freq_part = my_mode == 2 && !profile_part;
	    if (profile_part) {
		for (int k = 0; k < K; ++k) v[k] = it.vss[it.felm + k];
	    } else if (freq_part) {
		for (int k = 0; k < it.felm; ++k) {
		    int	kk = dxd? pg_decompact[k]: k;
		    if (kk < K) v[kk] = it.vss[k];
		}
	    } else if (is_a && ot_mode != 2) {		// raw a against raw b: fold a through the matrix
		for (int i = 0; i < sd->many; ++i) {
		    double	w = sd->weight? sd->weight[i]: 1.;
		    for (int k = 0; k < K; ++k) v[k] += w * sm->mtx[it.res[i]][k];
		}
	    } else {					// residue counts (weights) per code
		for (int i = 0; i < sd->many; ++i) v[it.res[i]] += sd->weight? sd->weight[i]: 1.;
	    }
	    // Entries of residue codes that never occur may be uninitialised in Simmtx (and in the profiles convseq derives
	    // from it): the reference indexes by residue and never reads them, the contraction meets them with weight 0,
	    // and 0 x NaN is NaN.  Non-finite entries therefore count as 0.
	    for (int k = 0; k < K; ++k) if (!(fabs(v[k]) <= 1.7e308)) v[k] = 0;
	    if (it.sfq) {S.sfq[x] = S.pool(it.sfq); S.tfq[x] = S.pool(it.tfq); S.rfq[x] = S.pool(it.rfq);}
	}
	if (S.glen.empty()) {S.glen.push_back(-1); S.gfreq.push_back(0);}
}

static pg_group pg_view(mSeq* sd, PgSide& S)
{
	pg_group	g;
	g.many = sd->many; g.len = sd->len; g.left = sd->left; g.right = sd->right;
	g.hetero = sd->gfq? sd->gfq->hetero: -1;
	g.nils = sd->inex.nils;
	g.cfq = S.cfq.data(); g.efq = S.efq.data(); g.vec = S.vec.data();
	g.glen = S.glen.data(); g.gfreq = S.gfreq.data(); g.npool = (int32_t) S.glen.size();
	bool	lists = sd->gfq && sd->inex.dels;
	g.sfq = lists? S.sfq.data(): 0; g.tfq = lists? S.tfq.data(): 0; g.rfq = lists? S.rfq.data(): 0;
	g.gapmask = S.gapmask.empty()? 0: S.gapmask.data();
	g.weight = S.weight.empty()? 0: S.weight.data();
	return (g);
}

// ---- one staged call, and the rendezvous that turns concurrent calls into one batch ---------------------------
struct PgJob {
	PgSide	A, B;
	pg_group	ga, gb;
	pg_gparams	gp;
	bool	score_only;
	double	score;
	int64_t	rr[2];
	std::vector<pg_skl>	pts;
	bool	taken, done;
	PgJob() : score_only(false), score(0), taken(false), done(false) {rr[0] = rr[1] = 0;}
};

static void pg_run_batch(const std::vector<PgJob*>& batch)
{
	PgLease	ctx;
	const double	t1 = pg_stats.on? pg_now(): 0;
	double	kms = 0;
	for (int pass = 0; pass < 2; ++pass) {		// pass 0: alignments with path, pass 1: score only
	    std::vector<PgJob*>	jobs;
	    for (PgJob* j: batch) if ((int) j->score_only == pass) jobs.push_back(j);
	    if (jobs.empty()) continue;
	    const size_t	n = jobs.size();
	    std::vector<pg_group>	ga(n), gb(n);
	    std::vector<pg_gparams>	gp(n);
	    std::vector<double>	scr(n);
	    for (size_t i = 0; i < n; ++i) {ga[i] = jobs[i]->ga; gb[i] = jobs[i]->gb; gp[i] = jobs[i]->gp;}
	    if (pass == 0) {
		int64_t*	offs = 0;
		pg_skl*	pts = 0;
		if (pg_align_groups(ctx, ga.data(), gb.data(), gp.data(), (int64_t) n, scr.data(), &offs, &pts) != PG_OK)
		    fatal("prrn_gpu alignC: %s\n", pg_last_error(ctx));
		for (size_t i = 0; i < n; ++i) {
		    jobs[i]->score = scr[i];
		    jobs[i]->pts.assign(pts + offs[i], pts + offs[i + 1]);
		}
		pg_free(offs); pg_free(pts);
	    } else {
		std::vector<int64_t>	rr(2 * n);
		if (pg_score_groups(ctx, ga.data(), gb.data(), gp.data(), (int64_t) n, scr.data(), rr.data()) != PG_OK)
		    fatal("prrn_gpu HomScoreC: %s\n", pg_last_error(ctx));
		for (size_t i = 0; i < n; ++i) {jobs[i]->score = scr[i]; jobs[i]->rr[0] = rr[2 * i]; jobs[i]->rr[1] = rr[2 * i + 1];}
	    }
	    if (pg_stats.on) kms += pg_last_kernel_ms(ctx);
	}
	if (pg_stats.on) {
	    std::lock_guard<std::mutex>	lk(pg_stats.mu);
	    pg_stats.n_gpu += (long) batch.size(); ++pg_stats.n_batches;
	    if ((long) batch.size() > pg_stats.max_batch) pg_stats.max_batch = (long) batch.size();
	    pg_stats.t_gpu += pg_now() - t1; pg_stats.kernel_ms += kms;
	    for (PgJob* j: batch) {
		pg_stats.cells += (long) pg_group_cells(&j->ga, &j->gb, j->gp.sh);
		if (j->score_only) ++pg_stats.n_score;
	    }
	}
}

static std::mutex	rz_mu;
static std::condition_variable	rz_cv;
static std::vector<PgJob*>	rz_wait;		// staged calls not yet taken by a leader
static int	rz_live = 0;				// worker threads of the program that are alive
static thread_local bool	rz_worker = false;	// this thread was started by the program through pthread_create

static void pg_submit(PgJob* job)
{
	static const bool	enabled = !(getenv("PRRN_GPU_BATCH") && getenv("PRRN_GPU_BATCH")[0] == '0');
	static const long	wait_us = getenv("PRRN_GPU_BATCH_WAIT_US")? atol(getenv("PRRN_GPU_BATCH_WAIT_US")): 2000;
	if (!enabled || !rz_worker) {			// main thread (prrn5 -t0, aln): a batch of one, no waiting
	    std::vector<PgJob*>	one(1, job);
	    pg_run_batch(one);
	    return;
	}
	std::unique_lock<std::mutex>	lk(rz_mu);
	rz_wait.push_back(job);
	const auto	deadline = std::chrono::steady_clock::now() + std::chrono::microseconds(wait_us);
	while (!job->done) {
	    if (!job->taken && ((int) rz_wait.size() >= rz_live || std::chrono::steady_clock::now() >= deadline)) {
		std::vector<PgJob*>	batch;		// every live worker is parked here (or time is up): lead the batch
		batch.swap(rz_wait);
		for (PgJob* j: batch) j->taken = true;
		lk.unlock();
		pg_run_batch(batch);
		lk.lock();
		for (PgJob* j: batch) j->done = true;
		rz_cv.notify_all();
		break;
	    }
	    if (job->taken) rz_cv.wait(lk);
	    else rz_cv.wait_until(lk, deadline);
	}
}

// The reference starts its workers with pthread_create (src/prrn5.cc:606-609, src/calcserv.h:436-457, src/adjmat.cc:
// 255-273).  This definition in the executable is what those calls bind to; it forwards to the C library's and
// counts the threads whose start routine lies in the executable itself (not the CUDA runtime's or libstdc++'s).
struct PgTramp {void* (*fn)(void*); void* arg;};
static void* pg_trampoline(void* p)
{
	PgTramp	t = *(PgTramp*) p;
	delete (PgTramp*) p;
	rz_worker = true;
	void*	r = t.fn(t.arg);
	{
	    std::lock_guard<std::mutex>	lk(rz_mu);
	    --rz_live;
	}
	rz_cv.notify_all();				// the others may be complete without this thread now
	return r;
}

extern "C" int pthread_create(pthread_t* th, const pthread_attr_t* attr, void* (*fn)(void*), void* arg) noexcept
{
	typedef int (*create_t)(pthread_t*, const pthread_attr_t*, void* (*)(void*), void*);
	static create_t	real = (create_t) dlsym(RTLD_NEXT, "pthread_create");
	if (!real) fatal("prrn_gpu: cannot resolve pthread_create\n");
	Dl_info	me, it;
	const bool	own = dladdr((void*) &pg_trampoline, &me) && dladdr((void*) fn, &it) && me.dli_fbase == it.dli_fbase;
	if (!own) return real(th, attr, fn, arg);
	PgTramp*	t = new PgTramp;
	t->fn = fn; t->arg = arg;
	{
	    std::lock_guard<std::mutex>	lk(rz_mu);
	    ++rz_live;
	}
	const int	rc = real(th, attr, pg_trampoline, t);
	if (rc) {
	    {
		std::lock_guard<std::mutex>	lk(rz_mu);
		--rz_live;
	    }
	    delete t;
	}
	return rc;
}

// ---- what the library takes ---------------------------------------------------------------------------------------
static const char* pg_untaken(mSeq* seqs[], PwdM* pwd, bool rectangle, WINDOW* pwdw, bool score_only = false)
{
	// quick mode (-Q): align2 / HomScore run the DPunit (no gap profile) rules on whatever groups they get
	// (src/maln2.cc:1882-1896); the library stages by pwd->alnmode
	if ((algmode.qck & 1) && pwd->alnmode != NGP_ALB && pwd->alnmode != NGP_ALN) return "quick mode (-Q) on groups with gap profile";
	if (pwdw) return "caller-supplied window";
	if (rectangle) {		// forwardA + initA: taken for the groups without gap profile (NGP_ALN)
	    if (score_only) return "rectangle (forwardA) without path, with its island reports";
	    if (pwd->alnmode != NGP_ALN) return "rectangle (forwardA) with gap profiles or naive groups";
	    if (seqs[0]->inex.exgr || seqs[1]->inex.exgr) return "rectangle (forwardA) with free right ends (store_ild)";
	    if (seqs[1]->left != 0) return "rectangle (forwardA) on a sub-window of b";
	    return 0;
	}
	switch (pwd->alnmode) {
	    case NGP_ALB: case HLF_ALB: case RHF_ALB: case GPF_ALB: return 0;
	    case NTV_ALB:
		if (seqs[0]->inex.nils || seqs[1]->inex.nils) return "naive groups (NTV_ALB) with nil ends";
		if (seqs[0]->many > 32 || seqs[1]->many > 32) return "naive groups (NTV_ALB) of more than 32 members";
		return 0;
	    default: return "this alignment mode";
	}
}

static void pg_fill_gparams(PwdM* pwd, pg_gparams* gp)
{
	gp->alnmode = pwd->alnmode; gp->Noll = pwd->Noll; gp->codonk1 = pwd->codonk1; gp->sh = pwd->alnprm.sh;
	gp->kdim = pwd->simmtx->dim; gp->u = pwd->alnprm.u;
	gp->Weighted_GOP = (double) (VTYPE) -pwd->alnprm.v;	// PwdM::resetuab, src/maln2.cc:238
	gp->Basic_GOP = (double) pwd->vgop(1);
	gp->BasicGOP = pwd->BasicGOP; gp->BasicGEP = pwd->BasicGEP; gp->LongGOP = pwd->LongGOP; gp->LongGEP = pwd->LongGEP;
}

// Two plain single sequences (what `aln a b` and the pairwise stages hand to alignC<DPunit>): unit thickness on
// every column, no weights, no nil ends, sim11.  Then the fill is the two-sequence recurrence pg_align_pairs
// stands for: kernel K2 (exact integers on DPX) when matrix and penalties are integral, the group kernel otherwise.
static bool pg_plain_pair(mSeq* seqs[], PwdM* pwd)
{
	if (pwd->alnmode != NGP_ALB || pwd->a_mode == 2 || pwd->b_mode == 2) return false;
	for (int s = 0; s < 2; ++s) {
	    mSeq*	sd = seqs[s];
	    if (sd->many != 1 || sd->weight || sd->inex.nils || sd->inex.exgl || sd->inex.exgr) return false;
	    for (int p = sd->left - 1; p < sd->right; ++p) {
		mSeqItr	it(sd, p);
		if (it.dns->cfq != 1 || it.dns->efq != 1) return false;
	    }
	}
	return !(pwd->alnprm.tgapf != 1);
}

static SKL* pg_skl_from(const pg_skl* pts, int n)
{
	SKL*	skl = new SKL[n + 1];		// callers delete[] it (src/maln2.cc:1923,1948)
	skl->m = 0; skl->n = n;
	for (int k = 0; k < n; ++k) {skl[k + 1].m = pts[k].m; skl[k + 1].n = pts[k].n;}
	return (skl);				// Vmf back-walk order; align2 runs stdskl next
}

static SKL* pg_align_plain_pair(mSeq* seqs[], PwdM* pwd, VTYPE* scr)
{
	const double	t0 = pg_stats.on? pg_now(): 0;
	mSeq	*a = seqs[0], *b = seqs[1];
	std::vector<uint8_t>	res((size_t) a->len + b->len + 1);
	memcpy(&res[0], a->at(0), a->len);			// many == 1: contiguous residues
	memcpy(&res[a->len], b->at(0), b->len);
	int64_t	offs[2] = {0, a->len};
	int32_t	lens[2] = {a->len, b->len}, left[2] = {a->left, b->left}, right[2] = {a->right, b->right};
	pg_seqs	S = {res.data(), offs, lens, left, right, 0, 2};
	pg_params	P;
	memset(&P, 0, sizeof(P));
	const ALPRM&	ap = pwd->alnprm;
	P.alprm.u = ap.u; P.alprm.v = ap.v; P.alprm.u0 = ap.u0; P.alprm.u1 = ap.u1; P.alprm.v0 = ap.v0;
	P.alprm.tgapf = ap.tgapf; P.alprm.thr = ap.thr; P.alprm.scale = ap.scale; P.alprm.maxsp = ap.maxsp;
	P.alprm.gamma = ap.gamma; P.alprm.k1 = ap.k1; P.alprm.ls = ap.ls; P.alprm.sh = ap.sh; P.alprm.mtx_no = ap.mtx_no;
	P.lcl = 0;
	P.vtype = sizeof(VTYPE) == sizeof(double);
	const Simmtx*	sm = pwd->simmtx;
	std::vector<VTYPE>	flat((size_t) sm->dim * sm->dim);
	for (int i = 0; i < sm->dim; ++i)
	    for (int j = 0; j < sm->dim; ++j) flat[(size_t) i * sm->dim + j] = sm->mtx[i][j];
	int32_t	ia = 0, ib = 1;
	int64_t*	po = 0;
	pg_skl*	pts = 0;
	VTYPE	s = 0;
	PgLease	ctx;
	if (pg_align_pairs(ctx, &S, &ia, &ib, 1, &P, flat.data(), sm->dim, &s, &po, &pts) != PG_OK)
	    fatal("prrn_gpu alignC (pair): %s\n", pg_last_error(ctx));
	*scr = s;
	SKL*	skl = pg_skl_from(pts, (int) po[1]);
	pg_free(po); pg_free(pts);
	if (pg_stats.on) {
	    std::lock_guard<std::mutex>	lk(pg_stats.mu);
	    ++pg_stats.n_gpu; ++pg_stats.n_k2; ++pg_stats.n_batches; pg_stats.t_gpu += pg_now() - t0;
	    pg_stats.kernel_ms += pg_last_kernel_ms(ctx);
	}
	return (skl);
}

static void pg_stage_job(mSeq* seqs[], PwdM* pwd, PgJob* job, bool score_only)
{
	const double	t0 = pg_stats.on? pg_now(): 0;
	pg_stage(seqs[0], job->A, pwd, true, pwd->simmtx);
	pg_stage(seqs[1], job->B, pwd, false, pwd->simmtx, pwd->alnmode == NGP_ALN? 1: 0);
	job->ga = pg_view(seqs[0], job->A); job->gb = pg_view(seqs[1], job->B);
	pg_fill_gparams(pwd, &job->gp);
	job->score_only = score_only;
	if (pg_stats.on) {
	    std::lock_guard<std::mutex>	lk(pg_stats.mu);
	    pg_stats.t_stage += pg_now() - t0;
	}
}

// PRRN_GPU_TRACE=1: one line per call on stderr (what reaches the DP, how long the kernels took).
// PRRN_GPU_VERIFY=1: after every call, the reference's own Fwd2c runs on the same inputs and the two results are
// compared (score within 1e-5 relative, corner lists identical); differences are reported on stderr and counted.
// PRRN_GPU_VERIFY=2 additionally continues with the reference's result, so that one divergent call does not change
// the inputs of all later ones.  Debugging aids for end-to-end parity runs (off by default; the reference's own code).
static const int	pg_verify = getenv("PRRN_GPU_VERIFY")? atoi(getenv("PRRN_GPU_VERIFY")): 0;
static const bool	pg_trace = getenv("PRRN_GPU_TRACE") != 0;
static long	pg_vcalls = 0, pg_vbad = 0;
struct PgVerifyReport {~PgVerifyReport() {if (pg_verify) fprintf(stderr, "prrn_gpu verify: %ld calls compared with the reference's Fwd2c, %ld differ\n", pg_vcalls, pg_vbad);}};
static PgVerifyReport	pg_vreport;

template <class recd_t>
static SKL* pg_check(mSeq* seqs[], PwdM* pwd, VTYPE* scr, SKL* skl, const char* route)
{
	if (pg_trace)
	    fprintf(stderr, "prrn_gpu trace: %s alnmode %d Noll %d members %d x %d columns %d x %d hetero %d / %d nils %d %d sh %d score %.10g corners %d\n",
		route, pwd->alnmode, pwd->Noll, seqs[0]->many, seqs[1]->many, seqs[0]->right - seqs[0]->left, seqs[1]->right - seqs[1]->left,
		seqs[0]->gfq? seqs[0]->gfq->hetero: -1, seqs[1]->gfq? seqs[1]->gfq->hetero: -1, (int) seqs[0]->inex.nils, (int) seqs[1]->inex.nils,
		pwd->alnprm.sh, (double) *scr, skl? skl->n: -1);
	if (!pg_verify) return skl;
	const bool	rect = pwd->alnmode == NGP_ALN;
	Fwd2c<recd_t>	pwa(seqs, pwd, true, rect, 0);
	VTYPE	rs = rect? pwa.forwardA(0): pwa.forwardB(0);
	SKL*	rk = pwa.traceback();
	bool	same = fabs((double) rs - (double) *scr) <= 1e-5 * (fabs((double) rs) > 1? fabs((double) rs): 1.) && rk && skl && rk->n == skl->n;
	int	first = -1;
	if (same) for (int k = 1; k <= rk->n; ++k) if (rk[k].m != skl[k].m || rk[k].n != skl[k].n) {same = false; first = k; break;}
	static std::mutex	mu;
	std::lock_guard<std::mutex>	lk(mu);
	++pg_vcalls;
	if (!same) {
	    ++pg_vbad;
	    fprintf(stderr, "prrn_gpu verify: call %ld (%s) DIFFERS: alnmode %d Noll %d members %d x %d window [%d,%d) x [%d,%d) hetero %d / %d nils %d %d "
		"sh %d: score gpu %.12g ref %.12g, corners gpu %d ref %d, first differing corner %d", pg_vcalls, route, pwd->alnmode, pwd->Noll,
		seqs[0]->many, seqs[1]->many, seqs[0]->left, seqs[0]->right, seqs[1]->left, seqs[1]->right,
		seqs[0]->gfq? seqs[0]->gfq->hetero: -1, seqs[1]->gfq? seqs[1]->gfq->hetero: -1, (int) seqs[0]->inex.nils, (int) seqs[1]->inex.nils,
		pwd->alnprm.sh, (double) *scr, (double) rs, skl? skl->n: -1, rk? rk->n: -1, first);
	    if (first > 0) fprintf(stderr, " (gpu %d,%d ref %d,%d)", skl[first].m, skl[first].n, rk[first].m, rk[first].n);
	    fputc('\n', stderr);
	}
	if (pg_verify >= 2) {delete[] skl; *scr = rs; return rk;}
	delete[] rk;
	return skl;
}

template <class recd_t>
static SKL* pg_alignC(mSeq* seqs[], PwdM* pwd, VTYPE* scr, bool rectangle, WINDOW* pwdw)
{
	if (const char* why = pg_untaken(seqs, pwd, rectangle, pwdw)) {
	    pg_refused("alignC", why);			// fatal() unless PRRN_GPU_ALLOW_REF=1: then the reference's own Fwd2c
	    const double	t0 = pg_stats.on? pg_now(): 0;
	    Fwd2c<recd_t>	pwa(seqs, pwd, true, rectangle, pwdw);
	    *scr = rectangle? pwa.forwardA(0): pwa.forwardB(0);
	    SKL*	r = pwa.traceback();
	    if (pg_stats.on) {std::lock_guard<std::mutex> lk(pg_stats.mu); ++pg_stats.n_ref; pg_stats.t_ref += pg_now() - t0;}
	    return r;
	}
	if (pg_plain_pair(seqs, pwd)) return pg_check<recd_t>(seqs, pwd, scr, pg_align_plain_pair(seqs, pwd, scr), "pair");
	PgJob	job;
	pg_stage_job(seqs, pwd, &job, false);
	pg_submit(&job);
	*scr = (VTYPE) job.score;
	return pg_check<recd_t>(seqs, pwd, scr, pg_skl_from(job.pts.data(), (int) job.pts.size()), "groups");
}

template <class recd_t>
static VTYPE pg_HomScoreC(mSeq* seqs[], PwdM* pwd, long rr[], bool rectangle, WINDOW* pwdw)
{
	if (const char* why = pg_untaken(seqs, pwd, rectangle, pwdw, true)) {
	    pg_refused("HomScoreC", why);
	    const double	t0 = pg_stats.on? pg_now(): 0;
	    Fwd2c<recd_t>	pwa(seqs, pwd, false, rectangle, pwdw);
	    VTYPE	s = rectangle? pwa.forwardA(rr): pwa.forwardB(rr);
	    if (pg_stats.on) {std::lock_guard<std::mutex> lk(pg_stats.mu); ++pg_stats.n_ref; pg_stats.t_ref += pg_now() - t0;}
	    return s;
	}
	PgJob	job;
	pg_stage_job(seqs, pwd, &job, true);
	pg_submit(&job);
	if (rr) {rr[0] = (long) job.rr[0]; rr[1] = (long) job.rr[1];}
	return (VTYPE) job.score;
}

template <> SKL* alignC<DPunit>(mSeq* seqs[], PwdM* pwd, VTYPE* scr, bool rectangle, WINDOW* pwdw)
	{return pg_alignC<DPunit>(seqs, pwd, scr, rectangle, pwdw);}
template <> SKL* alignC<DPunit_hf>(mSeq* seqs[], PwdM* pwd, VTYPE* scr, bool rectangle, WINDOW* pwdw)
	{return pg_alignC<DPunit_hf>(seqs, pwd, scr, rectangle, pwdw);}
template <> SKL* alignC<DPunit_pf>(mSeq* seqs[], PwdM* pwd, VTYPE* scr, bool rectangle, WINDOW* pwdw)
	{return pg_alignC<DPunit_pf>(seqs, pwd, scr, rectangle, pwdw);}
template <> SKL* alignC<DPunit_nv>(mSeq* seqs[], PwdM* pwd, VTYPE* scr, bool rectangle, WINDOW* pwdw)
	{return pg_alignC<DPunit_nv>(seqs, pwd, scr, rectangle, pwdw);}

template <> VTYPE HomScoreC<DPunit>(mSeq* seqs[], PwdM* pwd, long rr[], bool rectangle, WINDOW* pwdw)
	{return pg_HomScoreC<DPunit>(seqs, pwd, rr, rectangle, pwdw);}
template <> VTYPE HomScoreC<DPunit_hf>(mSeq* seqs[], PwdM* pwd, long rr[], bool rectangle, WINDOW* pwdw)
	{return pg_HomScoreC<DPunit_hf>(seqs, pwd, rr, rectangle, pwdw);}
template <> VTYPE HomScoreC<DPunit_pf>(mSeq* seqs[], PwdM* pwd, long rr[], bool rectangle, WINDOW* pwdw)
	{return pg_HomScoreC<DPunit_pf>(seqs, pwd, rr, rectangle, pwdw);}
template <> VTYPE HomScoreC<DPunit_nv>(mSeq* seqs[], PwdM* pwd, long rr[], bool rectangle, WINDOW* pwdw)
	{return pg_HomScoreC<DPunit_nv>(seqs, pwd, rr, rectangle, pwdw);}

// HomScore (src/maln2.cc:1837-1872).  align2 reaches alignC<recd_t> through the out-of-line instantiations above,
// but the compiler folds HomScoreC<recd_t> into HomScore inside maln2.o, so there is no call left to interpose:
// the shim therefore carries HomScore's own dispatch (the same switch), and the link takes this definition.
VTYPE HomScore(mSeq* seqs[], PwdM* pwdm, long rr[])
{
	mSeq*&	a = seqs[0];
	mSeq*&	b = seqs[1];
	if (a->left == a->right || b->left == b->right) {
	    if (rr) {rr[0] = b->left - a->left; rr[1] = b->right - a->right;}
	    return (0);
	}
	switch (pwdm->alnmode) {
	    case NGP_ALB: return HomScoreC<DPunit>(seqs, pwdm, rr, false, 0);
	    case HLF_ALB:
	    case RHF_ALB: return HomScoreC<DPunit_hf>(seqs, pwdm, rr, false, 0);
	    case GPF_ALB: return HomScoreC<DPunit_pf>(seqs, pwdm, rr, false, 0);
	    case NTV_ALB: return HomScoreC<DPunit_nv>(seqs, pwdm, rr, false, 0);
	    case NGP_ALN: return HomScoreC<DPunit>(seqs, pwdm, rr, true, 0);
	    case NTV_ALN: return HomScoreC<DPunit_nv>(seqs, pwdm, rr, true, 0);
	    case HLF_ALN:
	    case RHF_ALN: return HomScoreC<DPunit_hf>(seqs, pwdm, rr, true, 0);
	    case GPF_ALN: return HomScoreC<DPunit_pf>(seqs, pwdm, rr, true, 0);
	    // spliced modes (fwd2h.h / fwd2s.h): SURVEY section 8 keeps them out of scope -- the reference's own templates
	    case NGP_ALH: pg_refused("HomScore", "spliced mode (HomScoreH)"); return HomScoreH<RVPDJ_nv>(seqs, pwdm);
	    case HLF_ALH:
	    case RHF_ALH: pg_refused("HomScore", "spliced mode (HomScoreH)"); return HomScoreH<RVPDJ_hf>(seqs, pwdm);
	    case NGP_ALS: pg_refused("HomScore", "spliced mode (HomScoreS)"); return HomScoreS<RVPDJ_nv>(seqs, pwdm);
	    case HLF_ALS:
	    case RHF_ALS: pg_refused("HomScore", "spliced mode (HomScoreH)"); return HomScoreH<RVPDJ_hf>(seqs, pwdm);
	    default:
		fatal("Mode %d is not supported !\n", pwdm->alnmode);
	}
	return (0);
}

// swg1st (src/maln2.cc:1999-2027): the first pass of the Smith-Waterman mode (aln.cc:287-311 under algmode.lcl & 16),
// swg1stC<SwgDPunit | _hf | _pf | _nv> = Fwd2c::initC + forwardC (src/fwd2c.h:178-207,483-659,697-701).  As with
// HomScore the compiler folds swg1stC into swg1st inside maln2.o, so the shim carries the dispatch.  The library takes
// algmode.mlt <= 1 (colony 0 only: the best local score and the box of its path); the second pass, swg2nd -> swg2ndC ->
// align2 inside that box, reaches alignC above.
Colonies* swg1st(mSeq** seqs, PwdM* pwd)
{
	if (seqs[0]->left == seqs[0]->right || seqs[1]->left == seqs[1]->right)
	    return (0);
	const char*	why = 0;
	if (algmode.qck & 1) why = "quick mode (-Q) Smith-Waterman";
	else if (algmode.mlt > 1) why = "Smith-Waterman with secondary colonies (algmode.mlt > 1)";
	else switch (pwd->alnmode) {
	    case NGP_ALB: case HLF_ALB: case RHF_ALB: case GPF_ALB: break;
	    case NTV_ALB:
		if (seqs[0]->inex.nils || seqs[1]->inex.nils) why = "naive groups (NTV_ALB) with nil ends";
		else if (seqs[0]->many > 32 || seqs[1]->many > 32) why = "naive groups (NTV_ALB) of more than 32 members";
		break;
	    default:
		fatal("Mode %d is not supported !\n", pwd->alnmode);
	}
	if (why) {
	    pg_refused("swg1st", why);
	    if (algmode.qck & 1) return swg1stC<SwgDPunit>(seqs, pwd);
	    switch (pwd->alnmode) {
		case NGP_ALB: return swg1stC<SwgDPunit>(seqs, pwd);
		case HLF_ALB:
		case RHF_ALB: return swg1stC<SwgDPunit_hf>(seqs, pwd);
		case GPF_ALB: return swg1stC<SwgDPunit_pf>(seqs, pwd);
		default: return swg1stC<SwgDPunit_nv>(seqs, pwd);
	    }
	}
	PgJob	job;
	pg_stage_job(seqs, pwd, &job, false);
	double	val = 0;
	int32_t	box[6];
	{
	    PgLease	ctx;
	    if (pg_local_groups(ctx, &job.ga, &job.gb, &job.gp, 1, &val, box) != PG_OK)
		fatal("prrn_gpu swg1st: %s\n", pg_last_error(ctx));
	}
	if (pg_stats.on) {std::lock_guard<std::mutex> lk(pg_stats.mu); ++pg_stats.n_local;}
	Colonies*	cls = new Colonies();
	COLONY*	c0 = cls->at();		// no secondary colony: sortcolonies() leaves colony 0 in place (aln2.cc:368-372)
	c0->val = (VTYPE) val;
	c0->mlb = box[0]; c0->nlb = box[1]; c0->mrb = box[2]; c0->nrb = box[3]; c0->lwr = box[4]; c0->upr = box[5];
	return (cls);
}
