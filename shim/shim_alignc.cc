// shim/shim_alignc.cc -- the reference-side binding of INTEGRATION.md section 3, made real.
//
// The reference instantiates   template<class recd_t> SKL* alignC(mSeq* seqs[], PwdM*, VTYPE*, bool, WINDOW*)
// (src/fwd2c.h:670-677) inside maln2.o as WEAK out-of-line functions; align2 (src/maln2.cc:1888-1910)
// and prrn5's refinement reach the DP through them.  This file provides STRONG explicit specialisations
// for recd_t = DPunit, DPunit_hf, DPunit_pf, DPunit_nv, so that once it is linked with the unmodified reference
// objects every one of those calls lands here: the groups are staged exactly as Fwd2c would read them
// (mSeqItr over left-1 .. right-1 after PwdM's own convseq / mkthick / Gfq) and the banded fill with path
// runs in libprrn_gpu.so (pg_align_groups: kernels K4 + K3).  Modes the library does not take yet
// (rectangle, caller-supplied window, local) run the reference's own Fwd2c -- the reference's code, not a
// port; nothing here re-implements the DP on the CPU.
#include "aln.h"
#include "mseq.h"
#include "maln.h"
#include "mgaps.h"
#include "gfreq.h"
#include "vmf.h"
#include "fwd2c.h"
#include "prrn_gpu.h"

#include <chrono>
#include <cstdlib>
#include <vector>

// PRRN_GPU_STATS=1: calls / seconds per route, printed to stderr at exit (where does a prrn run spend its time?)
struct PgStats {
	long	n_gpu, n_ref; double t_stage, t_gpu, t_ref, kernel_ms; long cells;
	bool	on;
	PgStats() : n_gpu(0), n_ref(0), t_stage(0), t_gpu(0), t_ref(0), kernel_ms(0), cells(0), on(getenv("PRRN_GPU_STATS") != 0) {}
	~PgStats() {
	    if (on) fprintf(stderr, "prrn_gpu alignC: %ld calls on the GPU (staging %.2f s, pg_align_groups %.2f s of which "
		"kernels %.2f s, %.3g cells), %ld calls left on the reference's Fwd2c (%.2f s)\n",
		n_gpu, t_stage, t_gpu, kernel_ms * 1e-3, (double) cells, n_ref, t_ref);
	}
};
static PgStats	pg_stats;
static double	pg_now() {return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();}

static pg_context* pg_ctx_groups()
{
	// one context (CUDA stream + workspace) per calling thread: the reference calls the DP concurrently from
	// pthread workers (CalcServer, src/calcserv.h:436-457; Prrn::best_of_n, src/prrn5.cc:606-612)
	static thread_local pg_context* c = 0;
	if (!c && pg_create(0, &c) != PG_OK) fatal("prrn_gpu: %s\n", pg_last_error(0));
	return (c);
}

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
static void pg_stage(mSeq* sd, PgSide& S, const PwdM* pwd, bool is_a, const Simmtx* sm)
{
	const int	npos = sd->right - sd->left + 1, K = sm->dim;
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

template <class recd_t>
static SKL* pg_alignC(mSeq* seqs[], PwdM* pwd, VTYPE* scr, bool rectangle, WINDOW* pwdw)
{
	bool	banded = pwd->alnmode == NGP_ALB || pwd->alnmode == HLF_ALB ||
			 pwd->alnmode == RHF_ALB || pwd->alnmode == GPF_ALB ||
			 (pwd->alnmode == NTV_ALB && !seqs[0]->inex.nils && !seqs[1]->inex.nils &&
			  seqs[0]->many <= 32 && seqs[1]->many <= 32);
	const double	t0 = pg_stats.on? pg_now(): 0;
	if (rectangle || pwdw || !banded || (algmode.lcl & 16)) {	// not built yet: the reference's own Fwd2c
	    Fwd2c<recd_t>	pwa(seqs, pwd, true, rectangle, pwdw);
	    *scr = rectangle? pwa.forwardA(0): pwa.forwardB(0);
	    SKL*	r = pwa.traceback();
	    if (pg_stats.on) {++pg_stats.n_ref; pg_stats.t_ref += pg_now() - t0;}
	    return r;
	}
	PgSide	A, B;
	pg_stage(seqs[0], A, pwd, true, pwd->simmtx);
	pg_stage(seqs[1], B, pwd, false, pwd->simmtx);
	pg_group	ga = pg_view(seqs[0], A), gb = pg_view(seqs[1], B);
	pg_gparams	gp;
	gp.alnmode = pwd->alnmode; gp.Noll = pwd->Noll; gp.codonk1 = pwd->codonk1; gp.sh = pwd->alnprm.sh;
	gp.kdim = pwd->simmtx->dim; gp.u = pwd->alnprm.u;
	gp.Weighted_GOP = (double) (VTYPE) -pwd->alnprm.v;	// PwdM::resetuab, src/maln2.cc:238
	gp.Basic_GOP = (double) pwd->vgop(1);
	gp.BasicGOP = pwd->BasicGOP; gp.BasicGEP = pwd->BasicGEP; gp.LongGOP = pwd->LongGOP; gp.LongGEP = pwd->LongGEP;
	double	s = 0;
	int64_t*	offs = 0;
	pg_skl*	pts = 0;
	const double	t1 = pg_stats.on? pg_now(): 0;
	if (pg_align_groups(pg_ctx_groups(), &ga, &gb, &gp, 1, &s, &offs, &pts) != PG_OK)
	    fatal("prrn_gpu alignC: %s\n", pg_last_error(pg_ctx_groups()));
	if (pg_stats.on) {
	    ++pg_stats.n_gpu; pg_stats.t_stage += t1 - t0; pg_stats.t_gpu += pg_now() - t1;
	    pg_stats.kernel_ms += pg_last_kernel_ms(pg_ctx_groups());
	    pg_stats.cells += (long) pg_group_cells(&ga, &gb, gp.sh);
	}
	*scr = (VTYPE) s;
	int	n = (int) offs[1];
	SKL*	skl = new SKL[n + 1];		// callers delete[] it (src/maln2.cc:1923,1948)
	skl->m = 0; skl->n = n;
	for (int k = 0; k < n; ++k) {skl[k + 1].m = pts[k].m; skl[k + 1].n = pts[k].n;}
	pg_free(offs); pg_free(pts);
	return (skl);				// Vmf back-walk order; align2 runs stdskl next
}

template <> SKL* alignC<DPunit>(mSeq* seqs[], PwdM* pwd, VTYPE* scr, bool rectangle, WINDOW* pwdw)
	{return pg_alignC<DPunit>(seqs, pwd, scr, rectangle, pwdw);}
template <> SKL* alignC<DPunit_hf>(mSeq* seqs[], PwdM* pwd, VTYPE* scr, bool rectangle, WINDOW* pwdw)
	{return pg_alignC<DPunit_hf>(seqs, pwd, scr, rectangle, pwdw);}
template <> SKL* alignC<DPunit_pf>(mSeq* seqs[], PwdM* pwd, VTYPE* scr, bool rectangle, WINDOW* pwdw)
	{return pg_alignC<DPunit_pf>(seqs, pwd, scr, rectangle, pwdw);}
template <> SKL* alignC<DPunit_nv>(mSeq* seqs[], PwdM* pwd, VTYPE* scr, bool rectangle, WINDOW* pwdw)
	{return pg_alignC<DPunit_nv>(seqs, pwd, scr, rectangle, pwdw);}
