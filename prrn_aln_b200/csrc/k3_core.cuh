// k3_core.cuh -- per-cell machinery of kernel K3: group-to-group banded fill WITH path and gap-profile
// state.  Stands behind alignC<DPunit | DPunit_hf | DPunit_pf> (reference src/fwd2c.h:359-482 forwardB,
// :138-176 initB; src/fwd2c.cc gapopen / update specialisations; src/gfreq.cc:507-605 newgap /
// newdelta / incdelta; src/maln.h:185-187 unp1, :262-312 newgap1/2/3; src/dpunit.cc reset / copy).
// Shared by the CUDA kernel (k3_groups.cu) and the host emulation (tests/host_emul/k3_emul.cc).
//
// A DP record (the reference's DPunit_hf / DPunit_pf) is a run of 32-bit words
//     [0-1] val (double)  [2] ptr  [3] dir | glb << 8  [4 ..) dla: (glen, nins) u16 pairs  [..) dlb
// with list capacities fixed per alignment (hetero + 3), so records can live in any memory space
// (shared memory for the wavefront, HBM/L2 for the parked rows) and be addressed by (slot, row).
// Gap lengths / insertion counts are 16-bit: sequences up to 65,000 columns (checked by the host).
// Arithmetic is double for both VTYPE flavours (the group path is floating point with a 1e-5 relative
// tolerance: BASELINE north_star).
#pragma once
#include <string.h>
#include <stdint.h>

#include "k1_core.cuh"

#define K3_NEVSEL (-(1.7976931348623157e+308 / 16 * 7))
#define K3_LAST 0xffff
typedef unsigned short k3u16;

enum { K3_DIAG = 2, K3_NEWD = 3, K3_VERT = 4, K3_HORI = 8, K3_NEWV = 12, K3_NEWH = 13 };   // aln.h:47-52
PG_HD bool k3_isdiag(int d) { d &= 15; return d == 2 || d == 3; }
PG_HD bool k3_isvert(int d) { d &= 15; return (d >= 4 && d <= 7) || d == 12; }
PG_HD bool k3_ishori(int d) { d &= 15; return (d >= 8 && d <= 11) || d == 13; }

// One group as the DP reads it, columns left-1 .. right-1 (entry x <-> position left-1+x)
struct K3Group {
    const double* cfq;      // SeqThk::cfq
    const double* efq;      // SeqThk::efq
    const double* prof;     // [npos][kdim] profile row (what sim2 takes from the `a` side)
    const double* freq;     // [npos][kdim] frequency row (what sim2 takes from the `b` side)
    const int32_t* glen;    // pooled GFREQ lists: glen < 0 terminates
    const double* gfreq;    //                    freq
    const int32_t* sfq;     // [npos] list offsets, -1 = empty
    const int32_t* tfq;
    const int32_t* rfq;
    int32_t L;              // right - left
    int32_t nils;
    // mode 4 (DPunit_nv, groups of raw residues without profile): per column, bit i = member i holds a gap
    const uint32_t* gapmask;
    const double* weight;   // [many] member weights (1 when the reference has none)
    int32_t many;
    int32_t pad;
    // register-list form (k3r_core.cuh): fixed-stride column blocks {cfq, efq, freq[3][CS], glen[3][CS]}, or null
    const int* blk;
};

struct K3Prm {
    int32_t mode;           // 0 DPunit, 1 DPunit_hf, 2 DPunit_pf, 3 Aln2b1 (RVPD records, fwd2b1.cc), 4 DPunit_nv
    int32_t Noll, codonk1;
    int32_t lw, up;         // band in window-relative coordinates (r = n - m)
    int32_t capa, capb;     // list capacities (pairs)
    int32_t kdim;
    double u;               // alnprm.u as the reference's float
    double wgop, bgop;      // Weighted_GOP, Basic_GOP
    double u2divu1, v2divv1;
    // mode 3 (Aln2b1, src/fwd2b1.cc): PwdB penalties and the leading-gap factors of initB_ng
    double gop1, gep1, gop2, gep2;      // BasicGOP, BasicGEP, LongGOP, LongGEP
    double ltg_a, ltg_b;                // a.left ? 1 : (a.exgl ? 0 : tgapf), same for b
    // lastB_ng: trailing-gap factors (exgr ? 0 : tgapf) and whether the relaxation of the last column (gaps after
    // the end of b) / last row (after the end of a) runs at all (true sequence end and factor < 1)
    double rtg_a, rtg_b;
    int32_t last_c, last_r;
    // score only (HomScoreC, src/fwd2c.h:663-668: Fwd2c without Vmf): no path records; ptr carries the diagonal
    // n - m of the path's last first-row cell instead (fwd2c.h:468-469), the origin's b.left - a.left otherwise
    int32_t novmf, origin_r;
    // register-list form (k3r_core.cuh): words per dynamic list (4, 6 or 8; capa / capb then describe the same
    // records), 0 = the list-walking form of this file
    int32_t rl;
    // rectangle (alnmode NGP_ALN: forwardA + initA, src/fwd2c.h:111-135,231-356; record mode 0 only): every cell of
    // the window, no first-row / first-column skips, the boundary column without the long-gap switch; staged b holds
    // one more column (position b.right)
    int32_t rect;
    // Smith-Waterman (Fwd2c<SwgDPunit*>::initC + forwardC, src/fwd2c.h:178-207,483-659, algmode.mlt <= 1): cells are
    // cleared when negative, records carry the box of their local path in their last four words (capb counts them),
    // the result is the best cell and its box (colony 0) instead of an end cell and a path
    int32_t swg, pad3;
};

// ---- record access -------------------------------------------------------------------------------
PG_HD int k3_stride(int capa, int capb) { return (4 + capa + capb + 1) & ~1; }      // even: val stays 8-byte aligned
#if defined(__CUDA_ARCH__)
PG_HD double k3_val(const int* r) { return *reinterpret_cast<const double*>(r); }
PG_HD void k3_setval(int* r, double v) { *reinterpret_cast<double*>(r) = v; }
#else       // host emulation (tests): no type punning under the host compiler's strict-aliasing rules
static inline double k3_val(const int* r) { double v; memcpy(&v, r, sizeof v); return v; }
static inline void k3_setval(int* r, double v) { memcpy(r, &v, sizeof v); }
#endif
#define K3_PTR(r) ((r)[2])
PG_HD int k3_dir(const int* r) { return r[3] & 0xff; }
PG_HD int k3_glb(const int* r) { return (int)((unsigned)r[3] >> 8); }
PG_HD void k3_setdg(int* r, int dir, int glb) { r[3] = dir | (glb << 8); }
PG_HD k3u16* k3_dla(int* r) { return reinterpret_cast<k3u16*>(r + 4); }
PG_HD const k3u16* k3_dla(const int* r) { return reinterpret_cast<const k3u16*>(r + 4); }
PG_HD k3u16* k3_dlb(int* r, int capa) { return reinterpret_cast<k3u16*>(r + 4 + capa); }
PG_HD const k3u16* k3_dlb(const int* r, int capa) { return reinterpret_cast<const k3u16*>(r + 4 + capa); }

// IDELTA lists: pairs (glen, nins), terminated by glen == K3_LAST (gfreq.h:26,31)
PG_HD void k3_cleardelta(k3u16* d) { d[0] = 0; d[1] = 0; d[2] = K3_LAST; d[3] = 0; }       // gfreq.cc:556
PG_HD void k3_copydelta(k3u16* dst, const k3u16* src)
{   // gfreq.cc:548-554
    do { dst[0] = src[0]; dst[1] = src[1]; dst += 2; src += 2; } while (src[0] < K3_LAST);
    dst[0] = src[0]; dst[1] = src[1];
}
PG_HD void k3_incdelta(k3u16* dlt, const k3u16* dln, int n)
{   // gfreq.cc:595-602 (dlt may alias dln)
    do { dlt[0] = dln[0]; dlt[1] = (k3u16)(dln[1] + n); dlt += 2; dln += 2; } while (dln[0] < K3_LAST);
    dlt[0] = dln[0]; dlt[1] = dln[1];
}
// (Measured in round 2: walking the lists entry by entry is cheaper than touching their whole capacity -- the lists
// hold one or two entries on average.  Word-wise capacity-wide copydelta / incdelta without the data-dependent loop:
// 384 pairs 30.9 -> 36.1 ms; the branch-free register form of the whole cell, k3r_core.cuh: no gain either.)
// gfreq.cc:567-584: filter the dynamic gap state through a column's static gap state; dlt may alias dln
PG_HD void k3_newdelta(k3u16* dlt, const int32_t* glen, const k3u16* dln)
{
    k3u16* dst = dlt;
    int tg = 0, tn = 0;
    if (glen)
        for (; *glen >= 0; ++glen) {
            const int gl = *glen;
            if (gl >= dln[0]) {
                while (gl >= dln[2]) dln += 2;
                if (dln[1] > tn) {
                    const int nins = dln[1];
                    dst[0] = (k3u16)tg; dst[1] = (k3u16)tn; dst += 2;
                    tn = nins;
                    tg = gl + 1;
                }
            }
        }
    dst[0] = (k3u16)tg; dst[1] = (k3u16)tn;
    dst[2] = K3_LAST; dst[3] = 0;
}
PG_HD int k3_gaplensd(int gl, const k3u16* dl)
{   // gfreq.h:67-71
    while (gl >= dl[2]) dl += 2;
    return gl + dl[1];
}

// a GFREQ list view
struct K3List { const int32_t* glen; const double* freq; };
PG_HD K3List k3_list(const K3Group& g, const int32_t* offs, int ix)
{
    K3List l;
    const int o = offs ? offs[ix] : -1;
    l.glen = o >= 0 ? g.glen + o : nullptr;
    l.freq = o >= 0 ? g.gfreq + o : nullptr;
    return l;
}
PG_HD bool k3_neo(const K3List& l, int i) { return l.glen && l.glen[i] >= 0; }

// newgap(cf, dlc, df, dld), gfreq.cc:507-521
PG_HD double k3_newgap4(const K3List& cf, const k3u16* dlc, const K3List& df, const k3u16* dld)
{
    double g = 0;
    int ic = 0;
    for (int id = 0; k3_neo(df, id); ++id) {
        const int j = k3_gaplensd(df.glen[id], dld);
        for (; k3_neo(cf, ic); ++ic)
            if (k3_gaplensd(cf.glen[ic], dlc) >= j) break;
        if (!k3_neo(cf, ic)) break;
        g += cf.freq[ic] * df.freq[id];
    }
    return g;
}
// PwdM::newgap1(acf, dla, glb), maln.h:288-291 with newgap(cf, dlc, j) gfreq.cc:523-531, newgapc maln.h:270
PG_HD double k3_newgap1(double wgop, const K3List& acf, const k3u16* dla, int glb)
{
    if (!k3_neo(acf, 0)) return 0;
    if (k3_neo(acf, 1)) {
        for (int i = 0; k3_neo(acf, i); ++i)
            if (k3_gaplensd(acf.glen[i], dla) >= glb) return wgop * acf.freq[i];
        return wgop * 0.0;
    }
    return (dla[1] + acf.glen[0] >= glb) ? wgop * acf.freq[0] : 0;
}
// PwdM::newgap2(adf, glb, dla), maln.h:297-300 with newgap(df, i, dld) gfreq.cc:533-544, newgapd maln.h:278
PG_HD double k3_newgap2(double wgop, const K3List& adf, int glb, const k3u16* dla)
{
    if (!k3_neo(adf, 0)) return 0;
    if (k3_neo(adf, 1)) {
        double g = 0;
        for (int i = 0; k3_neo(adf, i); ++i) {
            while (adf.glen[i] >= dla[2]) dla += 2;
            if (glb < adf.glen[i] + dla[1]) break;
            g += adf.freq[i];
        }
        return wgop * g;
    }
    return (glb >= dla[1] + adf.glen[0]) ? wgop * adf.freq[0] : 0;
}

// sim2 as one contraction: profile row of a x frequency row of b (covers sim11 .. sim33, maln2.cc:534-623)
PG_HD double k3_sim(const K3Group& a, const K3Group& b, const K3Prm& p, int ia, int ib)
{
    const double* pa = a.prof + (size_t)ia * p.kdim;
    const double* fb = b.freq + (size_t)ib * p.kdim;
    double s = 0;
    for (int k = 0; k < p.kdim; ++k) s += pa[k] * fb[k];
    return s;
}
// unp1, maln.h:185-187: x.cfq * y.efq * -u
PG_HD double k3_unp(const K3Group& x, int ix, const K3Group& y, int iy, double u) { return x.cfq[ix] * y.efq[iy] * -u; }

// gapopen (fwd2c.cc:52-91 without di-thickness, :152-161, :202-212)
// DPunit_nv: crg2 (src/maln2.cc:881-1024 crg11/12i/21i/22i, :1454-1614 the weighted forms) is, for every
// variant, the many x many weighted form with unit weights where the reference has none.  The record's
// two "lists" hold the members' current gap run lengths (gla / glb, 16 bit each).  No nil ends:
// gapdensity = IsGap, postgapdensity = 1 (src/mseq.h:150-160).
PG_HD double k3_crg(const K3Group& a, const K3Group& b, const k3u16* gla, const k3u16* glb, unsigned am, unsigned bm, int d3)
{
    double g = 0;
    if (d3 == 0) {
        for (int i = 0; i < a.many; ++i) {
            double s = 0;
            const bool agap = (am >> i) & 1u;
            for (int j = 0; j < b.many; ++j) {
                const bool bgap = (bm >> j) & 1u;
                if (!agap ? (bgap && gla[i] >= glb[j]) : (!bgap && glb[j] >= gla[i])) s += b.weight[j];
            }
            g += s * a.weight[i];
        }
    } else if (d3 > 0) {
        for (int i = 0; i < a.many; ++i)
            if (!((am >> i) & 1u)) {
                double s = 0;
                for (int j = 0; j < b.many; ++j)
                    if (gla[i] >= glb[j]) s += b.weight[j];
                g += s * a.weight[i];
            }
    } else {
        for (int j = 0; j < b.many; ++j)
            if (!((bm >> j) & 1u)) {
                double s = 0;
                for (int i = 0; i < a.many; ++i)
                    if (glb[j] >= gla[i]) s += a.weight[i];
                g += s * b.weight[j];
            }
    }
    return g;
}

PG_HD double k3_gapopen(const K3Prm& p, const K3Group& a, const K3Group& b, const int* r, int ia, int ib, int d3)
{
    if (p.mode == 4) return k3_crg(a, b, k3_dla(r), k3_dlb(r, p.capa), a.gapmask[ia], b.gapmask[ib], d3) * p.bgop;
    if (p.mode == 0) {
        double axb = 0;
        if (d3 > 0) { if (!k3_isvert(k3_dir(r))) axb = a.cfq[ia] * b.efq[ib]; }
        else if (d3 < 0) { if (!k3_ishori(k3_dir(r))) axb = b.cfq[ib] * a.efq[ia]; }
        else return 0;
        return p.bgop * axb;
    }
    if (p.mode == 1) {
        if (d3 == 0) return k3_newgap2(p.wgop, k3_list(a, a.tfq, ia), k3_glb(r), k3_dla(r));
        if (d3 > 0) return k3_newgap1(p.wgop, k3_list(a, a.sfq, ia), k3_dla(r), k3_glb(r));
        return k3_newgap2(p.wgop, k3_list(a, a.rfq, ia), k3_glb(r), k3_dla(r));
    }
    const k3u16* dla = k3_dla(r);
    const k3u16* dlb = k3_dlb(r, p.capa);
    if (d3 == 0)
        return k3_newgap4(k3_list(a, a.sfq, ia), dla, k3_list(b, b.tfq, ib), dlb) * p.bgop +
               k3_newgap4(k3_list(b, b.sfq, ib), dlb, k3_list(a, a.tfq, ia), dla) * p.bgop;
    if (d3 > 0) return k3_newgap4(k3_list(a, a.sfq, ia), dla, k3_list(b, b.rfq, ib), dlb) * p.bgop;
    return k3_newgap4(k3_list(b, b.sfq, ib), dlb, k3_list(a, a.rfq, ia), dla) * p.bgop;
}

// update (fwd2c.cc:93-102, :163-182, :214-233); dst may alias src
PG_HD void k3_update(const K3Prm& p, const K3Group& a, const K3Group& b, int* dst, const int* src, int ia, int ib,
                     double gpn, int d3)
{
    const int sdir = k3_dir(src);
    int dir;
    if (d3 > 0) dir = k3_ishori(sdir) ? K3_NEWV : K3_VERT;
    else if (d3 < 0) dir = k3_isvert(sdir) ? K3_NEWH : K3_HORI;
    else dir = k3_isdiag(sdir) ? K3_DIAG : K3_NEWD;
    int glb = 0;
    if (p.mode == 4) {          // elongap (src/mgaps.cc:442-451) on both run-length vectors
        const unsigned am = a.gapmask[ia], bm = b.gapmask[ib];
        k3u16* da = k3_dla(dst); const k3u16* sa = k3_dla(src);
        k3u16* db = k3_dlb(dst, p.capa); const k3u16* sb = k3_dlb(src, p.capa);
        for (int i = 0; i < a.many; ++i) da[i] = (k3u16)((d3 >= 0) ? (((am >> i) & 1u) ? sa[i] + 1 : 0) : sa[i] + 1);
        for (int j = 0; j < b.many; ++j) db[j] = (k3u16)((d3 <= 0) ? (((bm >> j) & 1u) ? sb[j] + 1 : 0) : sb[j] + 1);
    } else if (p.mode == 1) {
        const int o = a.tfq ? a.tfq[ia] : -1;
        if (d3 == 0) k3_newdelta(k3_dla(dst), o >= 0 ? a.glen + o : nullptr, k3_dla(src));
        else if (d3 > 0) { glb = k3_glb(src) + 1; k3_newdelta(k3_dla(dst), o >= 0 ? a.glen + o : nullptr, k3_dla(src)); }
        else k3_incdelta(k3_dla(dst), k3_dla(src), 1);
    } else if (p.mode == 2) {
        const int oa = a.tfq ? a.tfq[ia] : -1, ob = b.tfq ? b.tfq[ib] : -1;
        const int32_t* ta = oa >= 0 ? a.glen + oa : nullptr;
        const int32_t* tb = ob >= 0 ? b.glen + ob : nullptr;
        if (d3 == 0) { k3_newdelta(k3_dla(dst), ta, k3_dla(src)); k3_newdelta(k3_dlb(dst, p.capa), tb, k3_dlb(src, p.capa)); }
        else if (d3 > 0) { k3_newdelta(k3_dla(dst), ta, k3_dla(src)); k3_incdelta(k3_dlb(dst, p.capa), k3_dlb(src, p.capa), 1); }
        else { k3_newdelta(k3_dlb(dst, p.capa), tb, k3_dlb(src, p.capa)); k3_incdelta(k3_dla(dst), k3_dla(src), 1); }
    }
    const double v = k3_val(src) + gpn;
    const int ptr = K3_PTR(src);
    k3_setdg(dst, dir, glb);
    k3_setval(dst, v);
    K3_PTR(dst) = ptr;
}
PG_HD void k3_reset(const K3Prm& p, int* r)
{
    k3_setval(r, K3_NEVSEL);
    K3_PTR(r) = 0; k3_setdg(r, 0, 0);
    if (p.mode == 4) {          // reset<DPunit_nv>: vclear(gla, dtsize)
        for (int w = 0; w < p.capa + p.capb; ++w) r[4 + w] = 0;
        return;
    }
    k3_cleardelta(k3_dla(r));
    k3_cleardelta(k3_dlb(r, p.capa));
}
PG_HD void k3_copy(const K3Prm& p, int* d, const int* s)
{
    if (d == s) return;
    d[0] = s[0]; d[1] = s[1]; d[2] = s[2]; d[3] = s[3];
    if (p.mode == 4) { for (int w = 0; w < p.capa + p.capb; ++w) d[4 + w] = s[4 + w]; return; }
    if (p.mode == 1 || p.mode == 2) k3_copydelta(k3_dla(d), k3_dla(s));
    if (p.mode == 2) k3_copydelta(k3_dlb(d, p.capa), k3_dlb(s, p.capa));
}

// Path records (the reference's Vmf): (m, n, previous record); appended at NEWD / NEWV / NEWH cells
struct K3Vmf { int m, n, p; };

// The cell in four pieces.  The diagonal, vertical and horizontal candidates read disjoint inputs and
// write disjoint records, so one thread may run them in turn (k3_cell) or three threads side by side
// (the role-split latency kernel); k3_combine then applies the reference's selection order.
PG_HD void k3_part_diag(const K3Prm& p, const K3Group& a, const K3Group& b, int ia, int ib, double dab,
                        const int* hdiag, int* hout)
{   // diagonal (fwd2c.h:395-398), straight into the output record
    const double gop = k3_gapopen(p, a, b, hdiag, ia, ib, 0);
    k3_update(p, a, b, hout, hdiag, ia, ib, dab + gop, 0);
}
PG_HD void k3_part_vert(const K3Prm& p, const K3Group& a, const K3Group& b, int ia, int ib, bool first_row, double* pua,
                        const int* habove, const int* gabove, const int* g2above, int* gout, int* g2out, const int* black)
{
    if (!first_row) {       // vertical (fwd2c.h:401-409)
        if (a.nils || p.rect) *pua = k3_unp(a, ia, b, ib, p.u);      // forwardA: per cell (:272)
        double gnp = k3_gapopen(p, a, b, gabove, ia, ib, 1);
        double gop = k3_gapopen(p, a, b, habove, ia, ib, 1);
        if (!k3_isvert(k3_dir(habove)) && (k3_val(habove) + gop > k3_val(gabove) + gnp))
            k3_update(p, a, b, gout, habove, ia, ib, gop, 1);
        else k3_update(p, a, b, gout, gabove, ia, ib, gnp, 1);
        k3_setval(gout, k3_val(gout) + *pua);
        if (p.Noll == 3) {  // vertical2 (fwd2c.h:411-420)
            gnp = p.v2divv1 * k3_gapopen(p, a, b, g2above, ia, ib, 1);
            gop = p.rect ? p.v2divv1 + gop : p.v2divv1 * gop;           // forwardA adds (:276)
            if (!k3_isvert(k3_dir(habove)) && (k3_val(habove) + gop > k3_val(g2above) + gnp))
                k3_update(p, a, b, g2out, habove, ia, ib, gop, 1);
            else k3_update(p, a, b, g2out, g2above, ia, ib, gnp, 1);
            k3_setval(g2out, k3_val(g2out) + p.u2divu1 * *pua);
        }
    } else {                // first row: the G rows keep their untouched (black) records
        k3_copy(p, gout, black);
        if (p.Noll == 3) k3_copy(p, g2out, black);
    }
}
PG_HD void k3_part_hori(const K3Prm& p, const K3Group& a, const K3Group& b, int ia, int ib, bool first_col,
                        const int* hleft, int* f1, int* f2)
{
    if (first_col) return;  // horizontal (fwd2c.h:422-431)
    const double pub = k3_unp(b, ib, a, ia, p.u);
    double gnp = k3_gapopen(p, a, b, f1, ia, ib, -1);
    double gop = k3_gapopen(p, a, b, hleft, ia, ib, -1);
    if (!k3_ishori(k3_dir(hleft)) && (k3_val(hleft) + gop > k3_val(f1) + gnp))
        k3_update(p, a, b, f1, hleft, ia, ib, gop, -1);
    else k3_update(p, a, b, f1, f1, ia, ib, gnp, -1);
    k3_setval(f1, k3_val(f1) + pub);
    if (p.Noll == 3) {      // horizontal2 (fwd2c.h:433-442)
        gnp = p.v2divv1 * k3_gapopen(p, a, b, f2, ia, ib, -1);
        gop = p.v2divv1 * gop;
        if (!k3_ishori(k3_dir(hleft)) && (k3_val(hleft) + gop > k3_val(f2) + gnp))
            k3_update(p, a, b, f2, hleft, ia, ib, gop, -1);
        else k3_update(p, a, b, f2, f2, ia, ib, gnp, -1);
        k3_setval(f2, k3_val(f2) + p.u2divu1 * pub);
    }
}
// selection (fwd2c.h:409-453): G first, G2 beats it on >, F1 and F2 on >=; the gap state replaces the diagonal
// only if strictly better.  Returns true when the cell must append a path record (:465-467).
PG_HD bool k3_combine(const K3Prm& p, bool first_row, bool first_col, int* hout, const int* gout, const int* g2out,
                      const int* f1, const int* f2)
{
    const int* mx = gout;
    if (!first_row && p.Noll == 3 && k3_val(g2out) > k3_val(mx)) mx = g2out;
    if (!first_col) {
        if (k3_val(f1) >= k3_val(mx)) mx = f1;
        if (p.Noll == 3 && k3_val(f2) >= k3_val(mx)) mx = f2;
    }
    if (k3_val(mx) > k3_val(hout)) k3_copy(p, hout, mx);
    const int dir = k3_dir(hout);
    return dir == K3_NEWD || dir == K3_NEWV || dir == K3_NEWH;
}

// One DP cell (fwd2c.h:393-468).  Inputs are read-only records: hdiag = H(m-1,n-1), habove / gabove /
// g2above = row m-1 at column n (the black record outside the band), hleft = H(m,n-1).  f1 / f2 are the
// row's running horizontal states (updated in place); hout / gout / g2out receive H, G, G2 of the cell.
PG_HD bool k3_cell(const K3Prm& p, const K3Group& a, const K3Group& b, int ia, int ib, bool first_row, bool first_col,
                   double dab, double* pua, const int* hdiag, const int* habove, const int* gabove, const int* g2above,
                   const int* hleft, int* f1, int* f2, int* hout, int* gout, int* g2out, const int* black)
{
    k3_part_diag(p, a, b, ia, ib, dab, hdiag, hout);
    k3_part_vert(p, a, b, ia, ib, first_row, pua, habove, gabove, g2above, gout, g2out, black);
    k3_part_hori(p, a, b, ia, ib, first_col, hleft, f1, f2);
    return k3_combine(p, first_row, first_col, hout, gout, g2out, f1, f2);
}

// The same cell written as one function (the single-thread-per-row kernels use it: measured 30 % faster than
// the four pieces called in turn, whose records the compiler must re-read from shared memory in k3_combine).
PG_HD bool k3_cell_mono(const K3Prm& p, const K3Group& a, const K3Group& b, int ia, int ib, bool first_row, bool first_col,
                   double dab, double* pua, const int* hdiag, const int* habove, const int* gabove, const int* g2above,
                   const int* hleft, int* f1, int* f2, int* hout, int* gout, int* g2out, const int* black)
{
    // diagonal (fwd2c.h:395-398), straight into the output record
    double gop = k3_gapopen(p, a, b, hdiag, ia, ib, 0);
    k3_update(p, a, b, hout, hdiag, ia, ib, dab + gop, 0);
    double gnp;
    const int* mx;
    if (!first_row) {       // vertical (fwd2c.h:401-409)
        if (a.nils || p.rect) *pua = k3_unp(a, ia, b, ib, p.u);      // forwardA: per cell (:272)
        gnp = k3_gapopen(p, a, b, gabove, ia, ib, 1);
        gop = k3_gapopen(p, a, b, habove, ia, ib, 1);
        if (!k3_isvert(k3_dir(habove)) && (k3_val(habove) + gop > k3_val(gabove) + gnp))
            k3_update(p, a, b, gout, habove, ia, ib, gop, 1);
        else k3_update(p, a, b, gout, gabove, ia, ib, gnp, 1);
        k3_setval(gout, k3_val(gout) + *pua);
        mx = gout;
        if (p.Noll == 3) {  // vertical2 (fwd2c.h:411-420)
            gnp = p.v2divv1 * k3_gapopen(p, a, b, g2above, ia, ib, 1);
            gop = p.rect ? p.v2divv1 + gop : p.v2divv1 * gop;           // forwardA adds (:276)
            if (!k3_isvert(k3_dir(habove)) && (k3_val(habove) + gop > k3_val(g2above) + gnp))
                k3_update(p, a, b, g2out, habove, ia, ib, gop, 1);
            else k3_update(p, a, b, g2out, g2above, ia, ib, gnp, 1);
            k3_setval(g2out, k3_val(g2out) + p.u2divu1 * *pua);
            if (k3_val(g2out) > k3_val(mx)) mx = g2out;
        }
    } else {                // first row: the G rows keep their untouched (black) records
        k3_copy(p, gout, black);
        if (p.Noll == 3) k3_copy(p, g2out, black);
        mx = gout;
    }
    if (!first_col) {       // horizontal (fwd2c.h:422-431)
        const double pub = k3_unp(b, ib, a, ia, p.u);
        gnp = k3_gapopen(p, a, b, f1, ia, ib, -1);
        gop = k3_gapopen(p, a, b, hleft, ia, ib, -1);
        if (!k3_ishori(k3_dir(hleft)) && (k3_val(hleft) + gop > k3_val(f1) + gnp))
            k3_update(p, a, b, f1, hleft, ia, ib, gop, -1);
        else k3_update(p, a, b, f1, f1, ia, ib, gnp, -1);
        k3_setval(f1, k3_val(f1) + pub);
        if (k3_val(f1) >= k3_val(mx)) mx = f1;
        if (p.Noll == 3) {  // horizontal2 (fwd2c.h:433-442)
            gnp = p.v2divv1 * k3_gapopen(p, a, b, f2, ia, ib, -1);
            gop = p.v2divv1 * gop;
            if (!k3_ishori(k3_dir(hleft)) && (k3_val(hleft) + gop > k3_val(f2) + gnp))
                k3_update(p, a, b, f2, hleft, ia, ib, gop, -1);
            else k3_update(p, a, b, f2, f2, ia, ib, gnp, -1);
            k3_setval(f2, k3_val(f2) + p.u2divu1 * pub);
            if (k3_val(f2) >= k3_val(mx)) mx = f2;
        }
    }
    if (k3_val(mx) > k3_val(hout)) k3_copy(p, hout, mx);        // fwd2c.h:453
    const int dir = k3_dir(hout);
    return dir == K3_NEWD || dir == K3_NEWV || dir == K3_NEWH;   // fwd2c.h:465-467
}

// ---- Smith-Waterman cell (forwardC, src/fwd2c.h:483-659, without secondary colonies) ---------------------
// box of a record: {lwr, upr, mlb, nlb} (SwgDPunit, src/dpunit.h:53-63) in the record's last four words
#define K3S_POS_INT (0x7fffffff / 8 * 7)
#define K3S_NEG_INT ((-0x7fffffff - 1) / 8 * 7)
PG_HD int* k3s_box(const K3Prm& p, int* r) { return r + p.capa + p.capb; }
PG_HD const int* k3s_box(const K3Prm& p, const int* r) { return r + p.capa + p.capb; }
// clear / reset<SwgDPunit*> (src/dpunit.cc): blank (val 0) or black (val NEVSEL) record with an empty box
PG_HD void k3s_blank(const K3Prm& p, int* r, double v)
{
    k3_reset(p, r);
    k3_setval(r, v);
    int* bx = k3s_box(p, r);
    bx[0] = K3S_POS_INT; bx[1] = K3S_NEG_INT; bx[2] = 0; bx[3] = 0;
}
PG_HD void k3s_copy(const K3Prm& p, int* d, const int* s)
{
    if (d == s) return;
    k3_copy(p, d, s);
    int* bd = k3s_box(p, d); const int* bs = k3s_box(p, s);
    bd[0] = bs[0]; bd[1] = bs[1]; bd[2] = bs[2]; bd[3] = bs[3];
}
// gapopen: SwgDPunit's own rule for groups without gap profile (src/fwd2c.cc:268-271), the banded rules otherwise
PG_HD double k3s_gapopen(const K3Prm& p, const K3Group& a, const K3Group& b, const int* r, int ia, int ib, int d3)
{
    if (p.mode == 0) return (k3_isdiag(k3_dir(r)) && d3) ? p.gop1 : 0;
    return k3_gapopen(p, a, b, r, ia, ib, d3);
}
// swg_gdpunit_update (src/fwd2c.cc:273-289) + the list part of the record type; dst may alias src; r = n - m (absolute)
PG_HD void k3s_update(const K3Prm& p, const K3Group& a, const K3Group& b, int* dst, const int* src, int ia, int ib,
                      double gpn, int d3, int r)
{
    const int sdir = k3_dir(src);
    const int* bs = k3s_box(p, src);
    const int lwr = bs[0], upr = bs[1], mlb = bs[2], nlb = bs[3];
    k3_update(p, a, b, dst, src, ia, ib, gpn, d3);
    int* bd = k3s_box(p, dst);
    bd[0] = (d3 > 0 && r < lwr) ? r : lwr;
    bd[1] = (d3 < 0 && r > upr) ? r : upr;
    bd[2] = mlb; bd[3] = nlb;
    k3_setdg(dst, d3 > 0 ? K3_VERT : (d3 < 0 ? K3_HORI : (k3_isdiag(sdir) ? K3_DIAG : K3_NEWD)), k3_glb(dst));
}
// the best cell so far of one thread (colony 0 in the making)
struct K3Best { double val; int mlb, nlb, mrb, nrb, lwr, upr; };
PG_HD void k3s_part_diag(const K3Prm& p, const K3Group& a, const K3Group& b, int ia, int ib, int r, double dab,
                         const int* hdiag, int* hout)
{
    const double gop = k3s_gapopen(p, a, b, hdiag, ia, ib, 0);
    k3s_update(p, a, b, hout, hdiag, ia, ib, dab + gop, 0, r);
}
PG_HD void k3s_part_vert(const K3Prm& p, const K3Group& a, const K3Group& b, int ia, int ib, int r, bool first_row,
                         const int* habove, const int* gabove, const int* g2above, int* gout, int* g2out)
{
    if (first_row) {        // the G rows keep their untouched (black) records (:529-530)
        k3s_blank(p, gout, K3_NEVSEL);
        if (p.Noll == 3) k3s_blank(p, g2out, K3_NEVSEL);
        return;
    }
    const double pua = k3_unp(a, ia, b, ib, p.u);           // per cell (:531)
    double gnp = k3s_gapopen(p, a, b, gabove, ia, ib, 1);
    double gop = k3s_gapopen(p, a, b, habove, ia, ib, 1);
    if (!k3_isvert(k3_dir(habove)) && (k3_val(habove) + gop > k3_val(gabove) + gnp))
        k3s_update(p, a, b, gout, habove, ia, ib, gop, 1, r);
    else k3s_update(p, a, b, gout, gabove, ia, ib, gnp, 1, r);
    k3_setval(gout, k3_val(gout) + pua);
    if (p.Noll == 3) {      // vertical2 (:543-552)
        gnp = p.v2divv1 * k3s_gapopen(p, a, b, g2above, ia, ib, 1);
        gop = p.v2divv1 * gop;
        if (!k3_isvert(k3_dir(habove)) && (k3_val(habove) + gop > k3_val(g2above) + gnp))
            k3s_update(p, a, b, g2out, habove, ia, ib, gop, 1, r);
        else k3s_update(p, a, b, g2out, g2above, ia, ib, gnp, 1, r);
        k3_setval(g2out, k3_val(g2out) + p.u2divu1 * pua);
    }
}
PG_HD void k3s_part_hori(const K3Prm& p, const K3Group& a, const K3Group& b, int ia, int ib, int r, bool first_col,
                         const int* hleft, int* f1, int* f2)
{
    if (first_col) return;  // (:554)
    const double pub = k3_unp(b, ib, a, ia, p.u);
    double gnp = k3s_gapopen(p, a, b, f1, ia, ib, -1);
    double gop = k3s_gapopen(p, a, b, hleft, ia, ib, -1);
    if (!k3_ishori(k3_dir(hleft)) && (k3_val(hleft) + gop > k3_val(f1) + gnp))
        k3s_update(p, a, b, f1, hleft, ia, ib, gop, -1, r);
    else k3s_update(p, a, b, f1, f1, ia, ib, gnp, -1, r);
    k3_setval(f1, k3_val(f1) + pub);
    if (p.Noll == 3) {      // horizontal2 (:566-575)
        gnp = p.v2divv1 * k3s_gapopen(p, a, b, f2, ia, ib, -1);
        gop = p.v2divv1 * gop;
        if (!k3_ishori(k3_dir(hleft)) && (k3_val(hleft) + gop > k3_val(f2) + gnp))
            k3s_update(p, a, b, f2, hleft, ia, ib, gop, -1, r);
        else k3s_update(p, a, b, f2, f2, ia, ib, gnp, -1, r);
        k3_setval(f2, k3_val(f2) + p.u2divu1 * pub);
    }
}
// selection and book-keeping (:578-610): m, n absolute positions of the cell, r = n - m, diag = H(m-1, n-1)
PG_HD void k3s_combine(const K3Prm& p, bool first_row, bool first_col, int m, int n, double diag, int* hout, int* gout,
                       int* g2out, int* f1, int* f2, K3Best* best)
{
    const int r = n - m;
    const int* mx = gout;
    if (!first_row && p.Noll == 3 && k3_val(g2out) > k3_val(mx)) mx = g2out;
    if (!first_col) {
        if (k3_val(f1) >= k3_val(mx)) mx = f1;
        if (p.Noll == 3 && k3_val(f2) >= k3_val(mx)) mx = f2;
    }
    int* bx = k3s_box(p, hout);
    if (k3_val(mx) > k3_val(hout)) {            // non-diagonal
        k3s_copy(p, hout, mx);
        if (bx[0] > r) bx[0] = r;
        if (bx[1] < r) bx[1] = r;
    } else if (k3_val(hout) > diag) {
        if (diag == 0) { bx[0] = bx[1] = r; bx[2] = m; bx[3] = n; }     // new colony
        if (k3_val(hout) > best->val) {         // max local score (strict: the first cell of this thread's rows wins)
            best->val = k3_val(hout);
            best->mlb = bx[2]; best->nlb = bx[3]; best->mrb = m + 1; best->nrb = n + 1; best->lwr = bx[0]; best->upr = bx[1];
        }
    }
    if (k3_val(hout) < 0) {                     // reset to blank: h, f1 (twice in the reference), f2 and g2 -- not g
        k3s_blank(p, hout, 0);
        k3s_blank(p, f1, 0);
        if (p.Noll == 3) { k3s_blank(p, f2, 0); k3s_blank(p, g2out, 0); }
    }
}
// row-major order of two candidates for colony 0: the reference visits cells row by row and replaces on `>`
PG_HD bool k3s_better(const K3Best& x, const K3Best& y)
{
    if (x.val != y.val) return x.val > y.val;
    if (x.mrb != y.mrb) return x.mrb < y.mrb;
    return x.nrb < y.nrb;
}

// One DP cell of Aln2b1::forwardB_ng (src/fwd2b1.cc:176-249, global mode): a gap opens on >=, the
// vertical state displaces the diagonal on >, the horizontal one on >=; a gap state keeps the path
// pointer of the cell it opened from; only a resumed diagonal run (NEWD) appends a path record.
PG_HD bool k3_cell_b1(const K3Prm& p, double dab, const int* hdiag, const int* habove, const int* gabove,
                      const int* g2above, const int* hleft, int* f1, int* f2, int* hout, int* gout, int* g2out)
{
    k3_setval(hout, k3_val(hdiag) + dab);                               // :181-183
    K3_PTR(hout) = K3_PTR(hdiag);
    k3_setdg(hout, k3_isdiag(k3_dir(hdiag)) ? K3_DIAG : K3_NEWD, 0);
    const int* mx = hout;
    double x = k3_val(habove) + p.gop1;                                 // vertical :186-193
    if (x >= k3_val(gabove)) { k3_setval(gout, x); K3_PTR(gout) = K3_PTR(habove); k3_setdg(gout, K3_VERT, 0); }
    else { gout[0] = gabove[0]; gout[1] = gabove[1]; gout[2] = gabove[2]; gout[3] = gabove[3]; }
    k3_setval(gout, k3_val(gout) + p.gep1);
    if (k3_val(gout) > k3_val(mx)) mx = gout;
    if (p.Noll == 3) {                                                  // vertical2 :196-205
        x = k3_val(habove) + p.gop2;
        if (x >= k3_val(g2above)) { k3_setval(g2out, x); K3_PTR(g2out) = K3_PTR(habove); k3_setdg(g2out, K3_VERT, 0); }
        else { g2out[0] = g2above[0]; g2out[1] = g2above[1]; g2out[2] = g2above[2]; g2out[3] = g2above[3]; }
        k3_setval(g2out, k3_val(g2out) + p.gep2);
        if (k3_val(g2out) > k3_val(mx)) mx = g2out;
    }
    x = k3_val(hleft) + p.gop1;                                         // horizontal :207-214
    if (x >= k3_val(f1)) { k3_setval(f1, x); K3_PTR(f1) = K3_PTR(hleft); k3_setdg(f1, K3_HORI, 0); }
    k3_setval(f1, k3_val(f1) + p.gep1);
    if (k3_val(f1) >= k3_val(mx)) mx = f1;
    if (p.Noll == 3) {                                                  // horizontal2 :217-226
        x = k3_val(hleft) + p.gop2;
        if (x >= k3_val(f2)) { k3_setval(f2, x); K3_PTR(f2) = K3_PTR(hleft); k3_setdg(f2, 9 /* HORL */, 0); }
        k3_setval(f2, k3_val(f2) + p.gep2);
        if (k3_val(f2) >= k3_val(mx)) mx = f2;
    }
    if (mx != hout) { hout[0] = mx[0]; hout[1] = mx[1]; hout[2] = mx[2]; hout[3] = mx[3]; }      // :229
    const int dir = k3_dir(hout);
    return dir == K3_NEWD || dir == K3_NEWV || dir == K3_NEWH;          // :243-245
}
// boundary chains of initB_ng (src/fwd2b1.cc:64-98): k-th cell (1-based); records point at the origin
PG_HD void k3_boundary_b1(const K3Prm& p, int k, int* dst, const int* src, bool row)
{
    const double gpn = k == 1 ? ((1 > p.codonk1) ? p.gop2 + p.gep2 : p.gop1 + p.gep1) : (k > p.codonk1 ? p.gep2 : p.gep1);
    k3_setval(dst, k3_val(src) + gpn * (row ? p.ltg_a : p.ltg_b));
    K3_PTR(dst) = 1;
    k3_setdg(dst, row ? K3_HORI : K3_VERT, 0);
}

// Boundary cells (initB, fwd2c.h:138-176).  Row: asi at a.left-1 (ia = 0), k-th column (1-based).
PG_HD void k3_boundary_row(const K3Prm& p, const K3Group& a, const K3Group& b, int k, int* dst, const int* src)
{
    const int ia = 0, ib = k;           // b position left + k - 1  <->  staged index k
    const double pub = k3_unp(b, ib, a, ia, p.u);
    double gnp = k3_gapopen(p, a, b, src, ia, ib, -1);
    // initB tests the column count after its increment (:156-160), initA before (:125-131)
    gnp = (k - (p.rect ? 1 : 0) < p.codonk1) ? gnp + pub : (p.v2divv1 * gnp + p.u2divu1 * pub);
    k3_update(p, a, b, dst, src, ia, ib, gnp, -1);
}
// Column: bsi at b.left-1 (ib = 0), k-th row (1-based)
PG_HD void k3_boundary_col(const K3Prm& p, const K3Group& a, const K3Group& b, int k, int* dst, const int* src)
{
    // forwardA computes the boundary cell of a row in place (:245-249): no long-gap switch, and bsi is where the
    // previous row left it -- position 0 before the first row (`mSeqItr bsi(b, 0)`, :240; b.left = 0: staged index
    // 1), position b.right afterwards (staged index L + 1)
    const int ia = k, ib = p.rect ? (k == 1 ? 1 : b.L + 1) : 0;
    const double pua = k3_unp(a, ia, b, ib, p.u);
    double gnp = k3_gapopen(p, a, b, src, ia, ib, 1);
    gnp = (p.rect || k < p.codonk1) ? gnp + pua : (p.v2divv1 * gnp + p.u2divu1 * pua);
    k3_update(p, a, b, dst, src, ia, ib, gnp, 1);
}
