// oracle/ref_driver.cc -- TEST INFRASTRUCTURE (not product code).
//
// A small command-line driver of OUR OWN that links the UNMODIFIED reference objects
// (oracle/_ref/libref_{f,d}.a, built by oracle/Makefile from /root/reference/src) and calls the
// reference's own entry points for the DP hot path on in-memory sequences:
//
//   scores   : alnScoreD(const Seq*[2], const Simmtx*, int*)            reference src/fwd2d1.cc:324
//   dist     : calcdist(mSeq**, nn, DynScr)                             reference src/phyl.cc:318
//   align    : align2(mSeq*[2], PwdM*, VTYPE*, Gsinfo*)                 reference src/maln2.cc:1875
//   matrix   : dump of Simmtx::mtx as the reference built it            reference src/simmtx.cc:282-445
//   galign   : two GROUPS (MSA files): PwdM staging + alignC<recd_t> + align2; dumps what the DP
//              reads per column (thickness, profile vector, residues, gap-profile lists) so that the
//              oracle restatement and the CUDA kernel can be fed the reference's own staged inputs
//
// It is used (a) to freeze golden vectors under tests/golden/ (tools/make_golden.py), (b) to pin the
// C restatement in oracle/oracle.c, (c) as bench.py's cpu_baseline "reference" leg (dist/scores timing).
// Output is plain text on stdout, one record per line, parsed by tools/refio.py.
//
// usage: ref_driver <cmd> <multi-fasta> [key=value ...]
//   keys: mtx=blosum62|pam   u= v= sh= threads= lcl= ls= u1= k1= tgapf= molc=p|n rep=
#include "aln.h"
#include "vmf.h"
#include "wln.h"
#include "mseq.h"
#include "maln.h"
#include "mgaps.h"
#include "autocomp.h"
#include "phyl.h"
#include "consreg.h"
#include "fspscore.h"
#include "fwd2c.h"
#include <sys/time.h>
#include <string>
#include <vector>

// aln.h:334 declares HomScoreB_ng with two arguments; the definition (fwd2b1.cc:1317) takes three
extern VTYPE HomScoreB_ng(const Seq* seqs[], const PwdB* pwd, long rr[]);

void usage() {}
template <class seq_t> void AlnServer<seq_t>::setparam(int) {}
template <class seq_t> int AlnServer<seq_t>::localoption(int&, const char**&) { return 0; }

// galign mt=N: N workers call align2 at the same time, as the workers of Prrn::best_of_n do (src/prrn5.cc:565-612)
struct MtArg { mSeq* sq[2]; PwdM* pwd; VTYPE scr; SKL* skl; };
static void* mt_align2(void* p)
{
	MtArg*	a = (MtArg*) p;
	Gsinfo	gsi;
	a->skl = align2(a->sq, a->pwd, &a->scr, &gsi);
	gsi.skl = 0;
	return 0;
}

static double now_s()
{
	struct timeval tv; gettimeofday(&tv, 0);
	return tv.tv_sec + 1e-6 * tv.tv_usec;
}

static const char* kv(int argc, const char** argv, const char* key, const char* dflt)
{
	size_t	kl = strlen(key);
	for (int i = 3; i < argc; ++i)
	    if (!strncmp(argv[i], key, kl) && argv[i][kl] == '=') return argv[i] + kl + 1;
	return dflt;
}

static std::vector<mSeq*> read_all(const char* fn)
{
	std::vector<mSeq*> v;
	FILE*	fd = fopen(fn, "r");
	if (!fd) fatal("cannot open %s\n", fn);
	for (;;) {
	    mSeq*	sd = new mSeq();
	    if (!sd->fgetseq(fd)) { delete sd; break; }
	    sd->sid = (int) v.size();
	    v.push_back(sd);
	}
	fclose(fd);
	return v;
}

static void print_vt(double x) { printf("%.17g", x); }

static void dump_gfq(const char* tag, GFREQ** pp)
{
	if (!pp || !*pp) { printf(" %s -1", tag); return; }
	int	n = 0;
	for (const GFREQ* g = *pp; neogfq(g); ++g) ++n;
	printf(" %s %d", tag, n);
	for (const GFREQ* g = *pp; neogfq(g); ++g) {
	    printf(" %d ", g->glen); print_vt(g->freq); printf(" %d", g->nres);
	}
}

// everything Fwd2c reads from one group through mSeqItr, position by position
static void dump_group(int idx, mSeq* sd, bool one_more = false)
{
	printf("group %d many %d len %d left %d right %d nelm %d felm %d vect %d prof %d dels %d nils %d sngl %d exgl %d exgr %d hetero %d sumwt ",
	    idx, sd->many, sd->len, sd->left, sd->right, sd->nelm, sd->felm, (int) sd->inex.vect, (int) sd->inex.prof,
	    (int) sd->inex.dels, (int) sd->inex.nils, (int) sd->inex.sngl, (int) sd->inex.exgl, (int) sd->inex.exgr,
	    sd->gfq? sd->gfq->hetero: -1);
	print_vt(sd->sumwt); putchar('\n');
	printf("weight %d", sd->weight? sd->many: 0);
	if (sd->weight) for (int i = 0; i < sd->many; ++i) { putchar(' '); print_vt(sd->weight[i]); }
	putchar('\n');
	// rectangle mode (forwardA, fwd2c.h:231-356) also reads b's column at position `right`: dumped for groups without
	// gap profile only (NGP: thickness is all the boundary reads; the list pointers there are not always valid)
	for (int p = sd->left - 1; p < sd->right + (one_more && !sd->gfq? 1: 0); ++p) {
	    mSeqItr	it(sd, p);
	    printf("pos %d dns", p);
	    if (it.dns) { putchar(' '); print_vt(it.dns->cfq); putchar(' '); print_vt(it.dns->dfq); putchar(' '); print_vt(it.dns->efq); }
	    else printf(" nan nan nan");
	    printf(" res");
	    for (int i = 0; i < sd->many; ++i) printf(" %d", (p >= -1 && p <= sd->len)? it.res[i]: 0);
	    printf(" vss %d", it.vss? sd->nelm: 0);
	    if (it.vss) for (int i = 0; i < sd->nelm; ++i) { putchar(' '); print_vt(it.vss[i]); }
	    dump_gfq("sfq", it.sfq); dump_gfq("tfq", it.tfq); dump_gfq("rfq", it.rfq);
	    putchar('\n');
	}
}

int main(int argc, const char** argv)
{
	if (argc < 3) {
	    fputs("usage: ref_driver scores|dist|align|matrix fasta [key=value ...]\n", stderr);
	    return 1;
	}
	std::string cmd = argv[1];
	const char* mtxname = kv(argc, argv, "mtx", "blosum62");
	const char* molc = kv(argc, argv, "molc", "p");
	bool	isprot = *molc == 'p';

	optimize(GLOBAL, MAXIMUM);
	if (isprot) setdefPprm(250, 2., 9.);
	else setdefmolc(DNA);		// plain FASTA of ACGT is otherwise guessed per file (seq.cc:85)
	alprm.sh = atoi(kv(argc, argv, "sh", "-60"));
	const char* s;
	if ((s = kv(argc, argv, "u", 0)))  alprm.u = atof(s);
	if ((s = kv(argc, argv, "v", 0)))  alprm.v = atof(s);
	if ((s = kv(argc, argv, "u1", 0))) alprm.u1 = atof(s);
	if ((s = kv(argc, argv, "k1", 0))) alprm.k1 = atoi(s);
	if ((s = kv(argc, argv, "ls", 0))) alprm.ls = atoi(s);
	if ((s = kv(argc, argv, "tgapf", 0))) alprm.tgapf = atof(s);
	algmode.lcl = atoi(kv(argc, argv, "lcl", "0"));
	algmode.bnd = atoi(kv(argc, argv, "bnd", "1"));
	algmode.crs = atoi(kv(argc, argv, "crs", "0"));
	algmode.mlt = 1;
	algmode.mns = 1;
	algmode.thr = 0;
	OutPrm.trimend = false;
	thread_num = atoi(kv(argc, argv, "threads", "0"));
	int	rep = atoi(kv(argc, argv, "rep", "1"));
	if (strcmp(mtxname, "pam")) mdm_file[0] = mtxname;

	if (cmd == "galign") {		// argv[2] = group A (MSA file), fb=<group B>
	    const char* fb = kv(argc, argv, "fb", 0);
	    if (!fb) fatal("galign needs fb=<file>\n");
	    mSeq*	sq[2] = {new mSeq(), new mSeq()};
	    FILE*	fd = fopen(argv[2], "r");
	    if (!fd || !sq[0]->fgetseq(fd)) fatal("cannot read %s\n", argv[2]);
	    fclose(fd);
	    fd = fopen(fb, "r");
	    if (!fd || !sq[1]->fgetseq(fd)) fatal("cannot read %s\n", fb);
	    fclose(fd);
	    if (atoi(kv(argc, argv, "wt", "0"))) {	// deterministic, unequal sequence weights (sum = many)
		for (int g = 0; g < 2; ++g) {
		    int	nn = sq[g]->many;
		    if (nn < 2) continue;
		    sq[g]->weight = new FTYPE[nn];
		    FTYPE	tot = 0;
		    for (int i = 0; i < nn; ++i) tot += sq[g]->weight[i] = 0.5 + ((i * 37 + 11 * g) % 10) / 8.;
		    for (int i = 0; i < nn; ++i) sq[g]->weight[i] *= nn / tot;
		}
	    }
	    prePwd(sq[0]->inex.molc);
	    Simmtx*	sm = getSimmtx(0);
	    printf("#ref_driver cmd=galign vtype=%s u=%g v=%g u0=%g u1=%g k1=%d ls=%d sh=%d tgapf=%g scale=%g gamma=%g lcl=%d\n",
		sizeof(VTYPE) == 8? "f64": "f32", alprm.u, alprm.v, alprm.u0, alprm.u1, alprm.k1, alprm.ls, alprm.sh,
		alprm.tgapf, alprm.scale, alprm.gamma, (int) algmode.lcl);
	    PwdM*	pwd = new PwdM(sq);
	    printf("pwdm alnmode %d swp %d a_mode %d b_mode %d an %d bn %d Noll %d codonk1 %d DvsP %d\n", pwd->alnmode,
		(int) pwd->swp, (int) pwd->a_mode, (int) pwd->b_mode, pwd->an, pwd->bn, pwd->Noll, pwd->codonk1, pwd->DvsP);
	    printf("pwdc Vab "); print_vt(pwd->Vab);
	    printf(" BasicGOP "); print_vt(pwd->BasicGOP); printf(" BasicGEP "); print_vt(pwd->BasicGEP);
	    printf(" LongGOP "); print_vt(pwd->LongGOP); printf(" LongGEP "); print_vt(pwd->LongGEP);
	    printf(" vgop1 "); print_vt(pwd->vgop(1)); printf(" wgop1 "); print_vt(pwd->wgop(1));
	    printf(" u "); print_vt(pwd->alnprm.u); printf(" v "); print_vt(pwd->alnprm.v); putchar('\n');
	    printf("dim %d\n", sm->dim);
	    for (int i = 0; i < sm->dim; ++i) {
		for (int j = 0; j < sm->dim; ++j) { if (j) putchar(' '); print_vt(sm->mtx[i][j]); }
		putchar('\n');
	    }
	    dump_group(0, sq[0], algmode.bnd == 0);
	    dump_group(1, sq[1], algmode.bnd == 0);
	    WINDOW	wdw;
	    stripe((const Seq**) sq, &wdw, pwd->alnprm.sh);
	    printf("window %d %d %d\n", wdw.lw, wdw.up, wdw.width);
	    // the raw alignC result (Vmf back-walk order), as align2 dispatches it (maln2.cc:1899-1910)
	    VTYPE	scr = 0;
	    SKL*	skl = 0;
	    double	t0 = now_s();
	    int	rep = atoi(kv(argc, argv, "rep", "1"));
	    for (int r = 0; r < rep; ++r) {
		delete[] skl;
		switch (pwd->alnmode) {
		    case NGP_ALB: skl = alignC<DPunit>(sq, pwd, &scr); break;
		    case HLF_ALB: case RHF_ALB: skl = alignC<DPunit_hf>(sq, pwd, &scr); break;
		    case GPF_ALB: skl = alignC<DPunit_pf>(sq, pwd, &scr); break;
		    case NTV_ALB: skl = alignC<DPunit_nv>(sq, pwd, &scr); break;
		    case NGP_ALN: skl = alignC<DPunit>(sq, pwd, &scr, true); break;		// rectangle (bnd=0), maln2.cc:1911-1915
		    case NTV_ALN: skl = alignC<DPunit_nv>(sq, pwd, &scr, true); break;
		    case HLF_ALN: case RHF_ALN: skl = alignC<DPunit_hf>(sq, pwd, &scr, true); break;
		    case GPF_ALN: skl = alignC<DPunit_pf>(sq, pwd, &scr, true); break;
		    default: fatal("galign: alnmode %d not handled by the driver\n", pwd->alnmode);
		}
	    }
	    printf("time %.6f\n", (now_s() - t0) / rep);
	    printf("alignc "); print_vt(scr);
	    if (!skl) printf(" skl 0\n");
	    else {
		printf(" skl %d %d :", skl->n, skl->m);
		for (int k = 1; k <= skl->n; ++k) printf(" %d %d", skl[k].m, skl[k].n);
		putchar('\n');
	    }
	    delete[] skl;
	    {	// HomScore -> HomScoreC<recd_t>(seqs, pwd, rr): the score-only fill (maln2.cc:1837-1862, fwd2c.h:663-668)
		long	rr[2] = {0, 0};
		VTYPE	hs = HomScore(sq, pwd, rr);
		printf("homscore "); print_vt(hs); printf(" %ld %ld\n", rr[0], rr[1]);
	    }
	    if (algmode.lcl & 16) {	// Smith-Waterman (aln.cc:287-311): swg1st -> Fwd2c::forwardC, swg2nd -> align2 inside the colony's box
		Colonies*	clns = swg1st(sq, pwd);
		if (!clns) printf("swg none\n");
		else {
		    COLONY*	c0 = clns->at();
		    printf("swg size %d val ", clns->size()); print_vt(c0->val);
		    printf(" mlb %d nlb %d mrb %d nrb %d lwr %d upr %d\n", c0->mlb, c0->nlb, c0->mrb, c0->nrb, c0->lwr, c0->upr);
		    Gsinfo	g2;
		    SKL*	s2 = c0->val > 0? swg2nd(sq, pwd, &g2, c0): 0;
		    printf("swg2nd "); print_vt(s2? g2.scr: 0);
		    if (!s2) printf(" skl 0\n");
		    else {
			printf(" skl %d %d :", s2->n, s2->m);
			for (int k = 1; k <= s2->n; ++k) printf(" %d %d", s2[k].m, s2[k].n);
			putchar('\n');
		    }
		    g2.skl = 0;
		    delete[] s2;
		    delete clns;
		}
	    }
	    Gsinfo	gsi;
	    scr = 0;
	    skl = align2(sq, pwd, &scr, &gsi);
	    printf("align2 sh %d ", pwd->alnprm.sh); print_vt(scr);
	    if (!skl) printf(" skl 0\n");
	    else {
		printf(" skl %d %d :", skl->n, skl->m);
		for (int k = 1; k <= skl->n; ++k) printf(" %d %d", skl[k].m, skl[k].n);
		putchar('\n');
	    }
	    gsi.skl = 0;
	    delete[] skl;
	    int	mt = atoi(kv(argc, argv, "mt", "0"));
	    if (mt > 0) {
		std::vector<MtArg>	args(mt);
		std::vector<pthread_t>	th(mt);
		for (int k = 0; k < mt; ++k) {		// own sequences and PwdM per worker, built one after the other
		    for (int g = 0; g < 2; ++g) {
			args[k].sq[g] = new mSeq();
			FILE*	f = fopen(g? fb: argv[2], "r");
			if (!f || !args[k].sq[g]->fgetseq(f)) fatal("cannot read group %d\n", g);
			fclose(f);
			if (sq[g]->weight) {
			    args[k].sq[g]->weight = new FTYPE[sq[g]->many];
			    for (int i = 0; i < sq[g]->many; ++i) args[k].sq[g]->weight[i] = sq[g]->weight[i];
			}
		    }
		    args[k].pwd = new PwdM(args[k].sq);
		    args[k].scr = 0; args[k].skl = 0;
		}
		for (int k = 0; k < mt; ++k) pthread_create(&th[k], 0, mt_align2, &args[k]);
		for (int k = 0; k < mt; ++k) pthread_join(th[k], 0);
		for (int k = 0; k < mt; ++k) {
		    printf("mt %d ", k); print_vt(args[k].scr);
		    if (!args[k].skl) printf(" skl 0\n");
		    else {
			printf(" skl %d %d :", args[k].skl->n, args[k].skl->m);
			for (int q = 1; q <= args[k].skl->n; ++q) printf(" %d %d", args[k].skl[q].m, args[k].skl[q].n);
			putchar('\n');
		    }
		}
	    }
	    return 0;
	}

	std::vector<mSeq*> seqs = read_all(argv[2]);
	int	nn = (int) seqs.size();
	if (!nn) fatal("no sequence\n");
	if (!isprot && nn >= 2) {	// as aln does (aln.cc:368): prePwd(int) would pick the protein defaults for DNA (aln2.cc:66-77)
	    const Seq*	two[2] = {seqs[0], seqs[1]};
	    prePwd(two);
	} else prePwd(seqs[0]->inex.molc);
	Simmtx*	sm = getSimmtx(0);
	printf("#ref_driver cmd=%s n=%d vtype=%s u=%g v=%g u1=%g k1=%d ls=%d sh=%d tgapf=%g scale=%g lcl=%d threads=%d\n",
	    cmd.c_str(), nn, sizeof(VTYPE) == 8? "f64": "f32",
	    alprm.u, alprm.v, alprm.u1, alprm.k1, alprm.ls, alprm.sh, alprm.tgapf, alprm.scale,
	    (int) algmode.lcl, thread_num);

	if (cmd == "matrix") {
	    printf("dim %d\n", sm->dim);
	    for (int i = 0; i < sm->dim; ++i) {
		for (int j = 0; j < sm->dim; ++j) {
		    if (j) putchar(' ');
		    print_vt(sm->mtx[i][j]);
		}
		putchar('\n');
	    }
	} else if (cmd == "seqs") {		// residue codes as the reference encodes them
	    for (int i = 0; i < nn; ++i) {
		printf("seq %d %d %d %d :", i, seqs[i]->len, seqs[i]->left, seqs[i]->right);
		for (int p = 0; p < seqs[i]->len; ++p) printf(" %d", *seqs[i]->at(p));
		putchar('\n');
	    }
	} else if (cmd == "scores") {		// alnScoreD per pair, elem(i,j) order, i<j: a=seq i, b=seq j
	    double	t0 = now_s();
	    for (int r = 0; r < rep; ++r)
	    for (int j = 1; j < nn; ++j) {
		for (int i = 0; i < j; ++i) {
		    const Seq*	sq[2] = {seqs[i], seqs[j]};
		    if (algmode.lcl && !(algmode.lcl & 16)) {
			seqs[i]->exg_seq(algmode.lcl & 1, algmode.lcl & 2);
			seqs[j]->exg_seq(algmode.lcl & 4, algmode.lcl & 8);
			int	ends[2];
			VTYPE	scr = alnScoreD(sq, sm, ends);
			if (!r) { printf("score %d %d ", i, j); print_vt(scr); printf(" %d %d\n", ends[0], ends[1]); }
		    } else {
			VTYPE	scr = alnScoreD(sq, sm, 0);
			if (!r) { printf("score %d %d ", i, j); print_vt(scr); putchar('\n'); }
		    }
		}
	    }
	    printf("time %.6f\n", now_s() - t0);
	} else if (cmd == "dist") {		// calcdist(seqs, nn, DynScr): the all-vs-all guide-tree step
	    double	t0 = now_s();
	    FTYPE*	dist = 0;
	    for (int r = 0; r < rep; ++r) {
		delete[] dist;
		dist = calcdist(&seqs[0], nn, DynScr);
	    }
	    double	t1 = now_s();
	    if (!atoi(kv(argc, argv, "quiet", "0")))
	    for (int k = 0; k < ncomb(nn); ++k) { printf("dist %d ", k); print_vt(dist[k]); putchar('\n'); }
	    printf("time %.6f\n", (t1 - t0) / rep);
	    delete[] dist;
	} else if (cmd == "align") {		// align2 per pair (i<j): score + stdskl-normalised corner list
	    double	t0 = now_s();
	    for (int j = 1; j < nn; ++j) {
		for (int i = 0; i < j; ++i) {
		    mSeq*	sq[2] = {seqs[i], seqs[j]};
		    if (algmode.lcl & 16) { sq[0]->exg_seq(1, 1); sq[1]->exg_seq(1, 1); }
		    else {
			sq[0]->exg_seq(algmode.lcl & 1, algmode.lcl & 2);
			sq[1]->exg_seq(algmode.lcl & 4, algmode.lcl & 8);
		    }
		    PwdM*	pwd = new PwdM(sq);
		    Gsinfo	gsi;
		    VTYPE	scr = 0;
		    SKL*	skl = align2(sq, pwd, &scr, &gsi);
		    printf("align %d %d swp=%d mode=%d ", i, j, pwd->swp, pwd->alnmode);
		    print_vt(scr);
		    if (!skl) printf(" skl 0\n");
		    else {
			printf(" skl %d %d :", skl->n, skl->m);
			for (int k = 1; k <= skl->n; ++k) printf(" %d %d", skl[k].m, skl[k].n);
			putchar('\n');
		    }
		    printf("fstat %d %d ", i, j);
		    print_vt(gsi.fstat.val); printf(" %g %g %g %g\n", (double) gsi.fstat.mch,
			(double) gsi.fstat.mmc, (double) gsi.fstat.gap, (double) gsi.fstat.unp);
		    gsi.skl = 0;
		    delete[] skl;
		    delete pwd;
		}
	    }
	    printf("time %.6f\n", now_s() - t0);
	} else if (cmd == "alignb") {		// Aln2b1: alignB_ng (stdskl-normalised) + HomScoreB_ng per pair (i<j)
	    double	t0 = now_s();
	    for (int j = 1; j < nn; ++j) {
		for (int i = 0; i < j; ++i) {
		    const Seq*	sq[2] = {seqs[i], seqs[j]};
		    if (algmode.lcl && !(algmode.lcl & 16)) {	// semi-global: free end gaps per side (as aln does)
			seqs[i]->exg_seq(algmode.lcl & 1, algmode.lcl & 2);
			seqs[j]->exg_seq(algmode.lcl & 4, algmode.lcl & 8);
		    }
		    PwdB*	pwd = new PwdB(sq);
		    VTYPE	scr = 0;
		    SKL*	skl = alignB_ng(sq, pwd, &scr);
		    long	rr[2] = {0, 0};
		    VTYPE	hom = HomScoreB_ng(sq, pwd, rr);
		    printf("alignb %d %d ", i, j);
		    print_vt(scr); putchar(' '); print_vt(hom);
		    if (!skl) printf(" skl 0\n");
		    else {
			printf(" skl %d %d :", skl->n, skl->m);
			for (int k = 1; k <= skl->n; ++k) printf(" %d %d", skl[k].m, skl[k].n);
			putchar('\n');
		    }
		    delete[] skl;
		    delete pwd;
		}
	    }
	    printf("time %.6f\n", now_s() - t0);
	} else fatal("unknown command %s\n", cmd.c_str());
	return 0;
}
