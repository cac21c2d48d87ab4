// shim/shim_calcdist.cc -- the BATCH-level binding of INTEGRATION.md section 2, made real without a source patch.
//
// The reference computes the all-vs-all guide-tree distances in
//     FTYPE* calcdist(mSeq** sbuf, int nn, DistCal realign)                  (src/phyl.cc:318-342)
// as nn self scores + nn(nn-1)/2 independent dpscore jobs (src/phyl.cc:221-261), each ending in one alnScoreD.
// This file defines that symbol; linked IN FRONT OF the reference's archive with -Wl,--allow-multiple-definition
// the linker binds every caller -- the drivers and DistMat::DistMat (src/phyl.cc:521) -- to it, and the
// nn(nn-1)/2 pairs become ONE pg_calcdist call (kernels K1P / K1 / K1F with the fused distance epilogue).
// Everything the library does not take (realign != DynScr, groups, weights, the Smith-Waterman mode) goes to
// the reference's own function: oracle/Makefile compiles src/phyl.cc a second time with
// -Dcalcdist=calcdist_ref and makes every other symbol of that object local, so calcdist_ref is the unmodified
// reference code under another name.  No reference source is copied; no CPU re-implementation lives here.
#include "aln.h"
#include "mseq.h"
#include "maln.h"
#include "phyl.h"
#include "prrn_gpu.h"
#include "shim_ctx.h"

#include <stdlib.h>
#include <string.h>
#include <vector>

extern FTYPE*	calcdist_ref(mSeq** sbuf, int nn, DistCal realign);	// src/phyl.cc:318 compiled as calcdist_ref

FTYPE* calcdist(mSeq** sbuf, int nn, DistCal realign)
{
	Simmtx*	sm = getSimmtx(0);
	convertseqs(sbuf, nn, sm);			// as the reference does first (src/phyl.cc:320): sets thickness / sumwt
	// single unweighted sequences only (groups and weighted members go through PwdM in dpscore, src/phyl.cc:233-237)
	// (groups and weighted members: the reference's dpscore builds a PwdM per pair and calls HomScore / align2, which
	// land in shim_alignc.cc -- still the GPU; realign != DynScr is not a DP job at all)
	bool	take = realign == DynScr && nn >= 2;
	for (int i = 0; take && i < nn; ++i)
	    take = sbuf[i]->many == 1 && !sbuf[i]->weight;
	if (take && (algmode.lcl & 16)) {		// a DP mode the library refuses: fatal() unless PRRN_GPU_ALLOW_REF=1
	    pg_refused("calcdist", "Smith-Waterman distances (algmode.lcl & 16)");
	    take = false;
	}
	if (!take) {
	    if (getenv("PRRN_GPU_STATS"))
		fprintf(stderr, "prrn_gpu calcdist: %d sequences left on the reference's calcdist (realign %d, lcl %d, many %d, sumwt %g)\n",
		    nn, (int) realign, (int) algmode.lcl, nn? sbuf[0]->many: 0, nn? (double) sbuf[0]->sumwt: 0.);
	    return calcdist_ref(sbuf, nn, realign);
	}

	int64_t	total = 0;
	std::vector<int64_t>	offs(nn);
	std::vector<int32_t>	lens(nn), left(nn), right(nn);
	std::vector<uint8_t>	exg(nn);
	for (int i = 0; i < nn; ++i) {
	    offs[i] = total; lens[i] = sbuf[i]->len; left[i] = sbuf[i]->left; right[i] = sbuf[i]->right;
	    exg[i] = uint8_t((sbuf[i]->inex.exgl? 1: 0) | (sbuf[i]->inex.exgr? 2: 0));
	    total += sbuf[i]->len;
	}
	std::vector<uint8_t>	res(total + 1);
	for (int i = 0; i < nn; ++i) memcpy(&res[offs[i]], sbuf[i]->at(0), sbuf[i]->len);	// many == 1: contiguous
	pg_seqs	S = {res.data(), offs.data(), lens.data(), left.data(), right.data(), exg.data(), nn};
	pg_params	P;
	memset(&P, 0, sizeof(P));
	P.alprm.u = alprm.u;   P.alprm.v = alprm.v;   P.alprm.u0 = alprm.u0; P.alprm.u1 = alprm.u1;
	P.alprm.v0 = alprm.v0; P.alprm.tgapf = alprm.tgapf; P.alprm.thr = alprm.thr;
	P.alprm.scale = alprm.scale; P.alprm.maxsp = alprm.maxsp; P.alprm.gamma = alprm.gamma;
	P.alprm.k1 = alprm.k1; P.alprm.ls = alprm.ls; P.alprm.sh = alprm.sh; P.alprm.mtx_no = alprm.mtx_no;
	P.lcl = algmode.lcl;
	P.vtype = sizeof(VTYPE) == sizeof(double);
	std::vector<VTYPE>	flat((size_t) sm->dim * sm->dim);
	for (int i = 0; i < sm->dim; ++i)
	    for (int j = 0; j < sm->dim; ++j) flat[(size_t) i * sm->dim + j] = sm->mtx[i][j];
	FTYPE*	dist = new FTYPE[ncomb(nn)];		// the caller delete[]s it, as with the reference's own
	PgLease	ctx;
	int	rc = pg_calcdist(ctx, &S, &P, flat.data(), sm->dim, 0, (int64_t) ncomb(nn), dist);
	if (rc == PG_ERR_UNSUPPORTED) {			// a mode the library refuses: fatal() unless PRRN_GPU_ALLOW_REF=1
	    pg_refused("calcdist", pg_last_error(ctx));
	    delete[] dist;
	    return calcdist_ref(sbuf, nn, realign);
	}
	if (rc != PG_OK) fatal("prrn_gpu calcdist: %s\n", pg_last_error(ctx));
	if (getenv("PRRN_GPU_STATS"))
	    fprintf(stderr, "prrn_gpu calcdist: %d sequences, %d pairs in one pg_calcdist call\n", nn, ncomb(nn));
	return (dist);
}
