// shim/shim_fwd2b1.cc -- the reference-side binding of INTEGRATION.md section 4: Aln2b1 on the GPU.
//
// The reference's pairwise single-sequence aligner with path pointers,
//     SKL*  alignB_ng(const Seq* seqs[], const PwdB* pwd, VTYPE* scr)          (src/fwd2b1.cc:1347-1353)
//     VTYPE HomScoreB_ng(const Seq* seqs[], const PwdB* pwd, long rr[])        (src/fwd2b1.cc:1317-1323)
// is what prrn5's DynAln distances (AdjacentMat::dist_align2, src/adjmat.cc:78-88) and iden.cc:193 call.  This
// file defines both symbols; linked IN FRONT OF the reference's archive (-Wl,--allow-multiple-definition: the
// first definition wins) every caller lands here and the fill -- Aln2b1::initB_ng / forwardB_ng / lastB_ng with
// the Vmf path records, src/fwd2b1.cc:64-279 -- runs in libprrn_gpu.so (pg_align_pairs_ng: kernel K3, record
// mode 3).  The rest of fwd2b1.o (skl_rngB_ng, the local / seeded variants) stays the reference's own.
//
// What globalB_ng does around the fill is reproduced on the host, with the reference's own functions: the corner
// records come back in Vmf back-walk order with the origin appended when the path does not end there
// (trcbkalignB_ng, :1025-1051), then stdskl (src/gaps.cc:139) normalises them, and an empty list is returned as 0.
//
// Refused (fatal() unless PRRN_GPU_ALLOW_REF=1, then the reference's own function, compiled once more as
// alignB_ng_ref / HomScoreB_ng_ref by oracle/Makefile): the seeded quick mode (algmode.qck, seededB_ng), the
// Smith-Waterman mode (lcl & 16, fwdswgB_ng), a band of one diagonal (diagonalB_ng) and -- for alignB_ng only --
// bands of MaxVmfSpace cells or more, where the reference switches to its linear-space recursion (lspB_ng /
// centerB_ng, :492-782,1053-1095) whose choice among co-optimal paths the direct traceback does not reproduce.
// PRRN_GPU_LSP_DIRECT=1 takes those on the GPU as well (same score; the path may differ at ties).
#include "aln.h"
#include "vmf.h"
#include "prrn_gpu.h"
#include "shim_ctx.h"

#include <stdlib.h>
#include <string.h>
#include <vector>

extern SKL*	alignB_ng_ref(const Seq* seqs[], const PwdB* pwd, VTYPE* scr);		// src/fwd2b1.cc:1347 under another name
extern VTYPE	HomScoreB_ng_ref(const Seq* seqs[], const PwdB* pwd, long rr[]);	// src/fwd2b1.cc:1317 under another name

static const char* pg_b1_untaken(const Seq* seqs[], bool with_path)
{
const	Seq*	a = seqs[0];
const	Seq*	b = seqs[1];
	if (a->many != 1 || b->many != 1) return "Aln2b1 on sequences with several members";
	if (algmode.lcl & 16) return "Smith-Waterman local mode (fwdswgB_ng)";
	if (with_path && algmode.qck) return "seeded quick alignment (seededB_ng)";
	WINDOW	wdw;
	stripe(seqs, &wdw, alprm.sh);
	if (with_path && wdw.up == wdw.lw) return "single-diagonal band (diagonalB_ng)";
	if (with_path && !(getenv("PRRN_GPU_LSP_DIRECT") && getenv("PRRN_GPU_LSP_DIRECT")[0] == '1')) {
	    long	m = a->right - a->left, n = b->right - b->left;			// lspB_ng, src/fwd2b1.cc:1062-1067
	    long	k = wdw.lw - b->left + a->right, q = b->right - a->left - wdw.up;
	    long	cvol = m * n - (k * k + q * q) / 2;
	    if (!(cvol < MaxVmfSpace || m == 1 || n <= 1))
		return "a band of MaxVmfSpace cells or more (linear-space recursion lspB_ng: PRRN_GPU_LSP_DIRECT=1 traces it directly)";
	}
	return 0;
}

// one pair through pg_align_pairs_ng: score, and (with_path) the corner records of trcbkalignB_ng
static VTYPE pg_b1_call(const Seq* seqs[], std::vector<pg_skl>* corners)
{
const	Seq*	a = seqs[0];
const	Seq*	b = seqs[1];
	std::vector<uint8_t>	res((size_t) a->len + b->len + 1);
	memcpy(&res[0], ((Seq*) a)->at(0), a->len);		// many == 1: contiguous residues
	memcpy(&res[a->len], ((Seq*) b)->at(0), b->len);
	int64_t	offs[2] = {0, a->len};
	int32_t	lens[2] = {a->len, b->len}, left[2] = {a->left, b->left}, right[2] = {a->right, b->right};
	uint8_t	exg[2] = {uint8_t((a->inex.exgl? 1: 0) | (a->inex.exgr? 2: 0)),
			  uint8_t((b->inex.exgl? 1: 0) | (b->inex.exgr? 2: 0))};
	pg_seqs	S = {res.data(), offs, lens, left, right, exg, 2};
	pg_params	P;
	memset(&P, 0, sizeof(P));
	P.alprm.u = alprm.u;   P.alprm.v = alprm.v;   P.alprm.u0 = alprm.u0; P.alprm.u1 = alprm.u1;
	P.alprm.v0 = alprm.v0; P.alprm.tgapf = alprm.tgapf; P.alprm.thr = alprm.thr;
	P.alprm.scale = alprm.scale; P.alprm.maxsp = alprm.maxsp; P.alprm.gamma = alprm.gamma;
	P.alprm.k1 = alprm.k1; P.alprm.ls = alprm.ls; P.alprm.sh = alprm.sh; P.alprm.mtx_no = alprm.mtx_no;
	P.lcl = 0;						// free ends travel per sequence (exg), as Aln2b1 reads inex
	P.vtype = sizeof(VTYPE) == sizeof(double);
const	Simmtx*	sm = getSimmtx(alprm.mtx_no);
	std::vector<VTYPE>	flat((size_t) sm->dim * sm->dim);
	for (int i = 0; i < sm->dim; ++i)
	    for (int j = 0; j < sm->dim; ++j) flat[(size_t) i * sm->dim + j] = sm->mtx[i][j];
	int32_t	ia = 0, ib = 1;
	int64_t*	po = 0;
	pg_skl*	pts = 0;
	VTYPE	s = 0;
	PgLease	ctx;
	if (pg_align_pairs_ng(ctx, &S, &ia, &ib, 1, &P, flat.data(), sm->dim, &s, &po, &pts) != PG_OK)
	    fatal("prrn_gpu alignB_ng: %s\n", pg_last_error(ctx));
	if (corners) corners->assign(pts, pts + po[1]);
	pg_free(po); pg_free(pts);
	return s;
}

SKL* alignB_ng(const Seq* seqs[], const PwdB* pwd, VTYPE* scr)
{
	if (const char* why = pg_b1_untaken(seqs, true)) {
	    pg_refused("alignB_ng", why);
	    return alignB_ng_ref(seqs, pwd, scr);
	}
	if (seqs[0]->left == seqs[0]->right || seqs[1]->left == seqs[1]->right) return alignB_ng_ref(seqs, pwd, scr);	// no DP
	std::vector<pg_skl>	c;
	*scr = pg_b1_call(seqs, &c);
	int	n = (int) c.size();
	if (n == 0) return (0);				// globalB_ng, src/fwd2b1.cc:1309-1312
	SKL*	skl = new SKL[n + 1];			// the record globalB_ng reserves in front + the corners
	skl->n = n; skl->m = 1;
	for (int k = 0; k < n; ++k) {skl[k + 1].m = c[k].m; skl[k + 1].n = c[k].n;}
	return (stdskl(&skl));
}

VTYPE HomScoreB_ng(const Seq* seqs[], const PwdB* pwd, long rr[])
{
	if (const char* why = pg_b1_untaken(seqs, false)) {
	    pg_refused("HomScoreB_ng", why);
	    return HomScoreB_ng_ref(seqs, pwd, rr);
	}
	if (seqs[0]->left == seqs[0]->right || seqs[1]->left == seqs[1]->right) return HomScoreB_ng_ref(seqs, pwd, rr);
	return pg_b1_call(seqs, 0);			// without Vmf forwardB_ng leaves rr untouched (src/fwd2b1.cc:268-275)
}
