// k3r_core.cuh -- the "register list" form of kernel K3's cell for the gap-profile record types
// DPunit_hf (HLF_ALB / RHF_ALB) and DPunit_pf (GPF_ALB): same reference semantics as k3_core.cuh
// (src/fwd2c.h:359-482 forwardB, src/fwd2c.cc:151-233 gapopen / update, src/gfreq.cc:507-605 newgap / newdelta /
// incdelta, src/maln.h:262-312 newgap1/2/3), other machine form.
//
// k3_core.cuh walks the reference's variable-length lists entry by entry (a load, a compare and a branch per
// entry, 16 bits at a time): every step of a cell waits for the one before, and the threads of a warp -- each on
// another row with other lists -- diverge.  Here every list has a FIXED capacity chosen per launch and lives in
// registers for the time of a cell:
//   dynamic gap state (IDELTA lists)   CAP 32-bit words per list, one entry per word = glen << 16 | nins, ascending,
//                                      padded with K3R_TERM (glen 0xffff): a whole list is one or two 128-bit
//                                      shared-memory accesses, "which segment holds gl" is CAP - 1 compare-selects on
//                                      the packed words (no unpacking: the key is gl << 16 | 0xffff);
//   static gap profile (GFREQ lists)   per column a fixed-stride BLOCK {cfq, efq, freq[3][CS], glen[3][CS]} (CS =
//                                      CAP - 2 entries per list, padded with glen 0xffff / freq 0): the three lists of a
//                                      column sit at a computable address (no offset table, no pointer chase) and are
//                                      read with vector loads.
// The merges (newgap), the list rewrites (newdelta / incdelta) and the copies are unrolled over the capacity and
// branch-free: all lanes of a warp execute the same instruction stream, whose instructions are mostly independent.
// Padding is arranged so that it needs no validity tests: a padded static entry adjusts to 65,535 and carries
// frequency 0, a padded dynamic entry never matches a real length.  Lengths are below 32,768 (checked by the host).
//
// Results are bit-identical to k3_core.cuh (same operations on the same values in the same order); the kernel,
// the host emulation (tests/host_emul/k3_emul.cc) and the GPU tests run both forms against the reference's goldens.
#pragma once
#include <stdint.h>

#include "k3_core.cuh"

#define K3R_TERM 0xffff0000u
#define K3R_PAD 0xffff

// ---- static column blocks ------------------------------------------------------------------------------
// words per block: cfq, efq (2 doubles) | freq[3][CS] doubles | glen[3][CS] u16 | padding to a multiple of 4 words
PG_HD constexpr int k3r_block_words(int cs) { return (4 + 6 * cs + (3 * cs + 1) / 2 + 3) & ~3; }
PG_HD constexpr int k3r_rec_words(int cap, int mode) { return (4 + cap * (mode == 2 ? 2 : 1) + 3) & ~3; }

struct alignas(16) k3r_q4 { int x, y, z, w; };          // 16-byte moves (LDS.128 / STS.128 / LDG.128 on the device)
struct alignas(16) k3r_d2 { double a, b; };

template <int CS>
struct K3RList {            // one GFREQ list of a column in registers
    int g[CS];              // glen, K3R_PAD beyond the end
    double f[CS];           // freq, 0 beyond the end
};
template <int CS>
PG_HD const double* k3r_blk_freq(const int* blk, int which) { return reinterpret_cast<const double*>(blk + 4) + which * CS; }
template <int CS>
PG_HD const unsigned short* k3r_blk_glen(const int* blk, int which)
{
    return reinterpret_cast<const unsigned short*>(blk + 4 + 6 * CS) + which * CS;
}
// which: 0 sfq, 1 tfq, 2 rfq
// lengths: CS / 2 32-bit loads of two 16-bit lengths each; frequencies: CS / 2 16-byte loads of two doubles each
template <int CS>
PG_HD void k3r_load_glen(int (&g)[CS], const int* blk, int which)
{
    static_assert(CS % 2 == 0, "two lengths per word");
    const unsigned* p = reinterpret_cast<const unsigned*>(blk + 4 + 6 * CS) + which * (CS / 2);
#pragma unroll
    for (int i = 0; i < CS / 2; ++i) { const unsigned w = p[i]; g[2 * i] = (int)(w & 0xffffu); g[2 * i + 1] = (int)(w >> 16); }
}
template <int CS>
PG_HD void k3r_load_list(K3RList<CS>& l, const int* blk, int which)
{
    k3r_load_glen<CS>(l.g, blk, which);
    const k3r_d2* f = reinterpret_cast<const k3r_d2*>(blk + 4) + which * (CS / 2);
#pragma unroll
    for (int i = 0; i < CS / 2; ++i) { const k3r_d2 v = f[i]; l.f[2 * i] = v.a; l.f[2 * i + 1] = v.b; }
}
PG_HD double k3r_blk_cfq(const int* blk) { return reinterpret_cast<const double*>(blk)[0]; }
PG_HD double k3r_blk_efq(const int* blk) { return reinterpret_cast<const double*>(blk)[1]; }

// Host: the blocks of one group from the pooled lists of the C ABI (pg_group: glen < 0 ends a list, offset -1 = no
// list).  Returns the longest list met (so the caller can check it against CS); blocks are fully written.
static inline int k3r_build_blocks(int* out, int cs, int npos, const double* cfq, const double* efq, const int32_t* glen,
                                   const double* gfreq, const int32_t* sfq, const int32_t* tfq, const int32_t* rfq)
{
    const int bw = k3r_block_words(cs);
    int longest = 0;
    for (int x = 0; x < npos; ++x) {
        int* b = out + (size_t)x * bw;
        for (int w = 0; w < bw; ++w) b[w] = 0;
        reinterpret_cast<double*>(b)[0] = cfq[x];
        reinterpret_cast<double*>(b)[1] = efq[x];
        double* f = reinterpret_cast<double*>(b + 4);
        unsigned short* g = reinterpret_cast<unsigned short*>(b + 4 + 6 * cs);
        const int32_t* offs[3] = {sfq, tfq, rfq};
        for (int l = 0; l < 3; ++l) {
            int n = 0;
            const int o = offs[l] ? offs[l][x] : -1;
            if (o >= 0)
                for (const int32_t* p = glen + o; *p >= 0; ++p, ++n)
                    if (n < cs) { g[l * cs + n] = (unsigned short)*p; f[l * cs + n] = gfreq[o + n]; }
            if (n > longest) longest = n;
            for (int i = n < cs ? n : cs; i < cs; ++i) { g[l * cs + i] = K3R_PAD; f[l * cs + i] = 0.0; }
        }
    }
    return longest;
}
// longest static list and largest glen of a group (host-side choice of the capacity)
static inline void k3r_scan_lists(int npos, const int32_t* glen, const int32_t* sfq, const int32_t* tfq, const int32_t* rfq,
                                  int* longest, int* max_glen)
{
    const int32_t* offs[3] = {sfq, tfq, rfq};
    for (int l = 0; l < 3; ++l) {
        if (!offs[l]) continue;
        for (int x = 0; x < npos; ++x) {
            const int o = offs[l][x];
            if (o < 0) continue;
            int n = 0;
            for (const int32_t* p = glen + o; *p >= 0; ++p, ++n)
                if (*p > *max_glen) *max_glen = *p;
            if (n > *longest) *longest = n;
        }
    }
}

// ---- records -------------------------------------------------------------------------------------------
// words: [0-1] val  [2] ptr  [3] dir | glb << 8  [4, 4 + CAP) dla  [4 + CAP, 4 + 2 CAP) dlb (DPunit_pf only)
template <int CAP, int MODE>
struct K3RRec {
    double val;
    int ptr, dg;
    unsigned la[CAP];
    unsigned lb[MODE == 2 ? CAP : 1];
};
// a record travels as 16-byte quads (its stride is a multiple of 4 words, its address of 16 bytes)
template <int CAP, int MODE>
PG_HD void k3r_load(K3RRec<CAP, MODE>& r, const int* p)
{
    constexpr int NW = k3r_rec_words(CAP, MODE);
    int w[NW];
#pragma unroll
    for (int q = 0; q < NW / 4; ++q) {
        const k3r_q4 v = reinterpret_cast<const k3r_q4*>(p)[q];
        w[4 * q] = v.x; w[4 * q + 1] = v.y; w[4 * q + 2] = v.z; w[4 * q + 3] = v.w;
    }
    union { double d; int i[2]; } u;
    u.i[0] = w[0]; u.i[1] = w[1];
    r.val = u.d; r.ptr = w[2]; r.dg = w[3];
#pragma unroll
    for (int k = 0; k < CAP; ++k) r.la[k] = (unsigned)w[4 + k];
    if constexpr (MODE == 2) {
#pragma unroll
        for (int k = 0; k < CAP; ++k) r.lb[k] = (unsigned)w[4 + CAP + k];
    }
}
template <int CAP, int MODE>
PG_HD void k3r_store(int* p, const K3RRec<CAP, MODE>& r)
{
    constexpr int NW = k3r_rec_words(CAP, MODE);
    int w[NW];
    union { double d; int i[2]; } u;
    u.d = r.val;
    w[0] = u.i[0]; w[1] = u.i[1]; w[2] = r.ptr; w[3] = r.dg;
#pragma unroll
    for (int k = 0; k < CAP; ++k) w[4 + k] = (int)r.la[k];
    if constexpr (MODE == 2) {
#pragma unroll
        for (int k = 0; k < CAP; ++k) w[4 + CAP + k] = (int)r.lb[k];
    }
#pragma unroll
    for (int k = 4 + CAP * (MODE == 2 ? 2 : 1); k < NW; ++k) w[k] = 0;
#pragma unroll
    for (int q = 0; q < NW / 4; ++q) {
        k3r_q4 v;
        v.x = w[4 * q]; v.y = w[4 * q + 1]; v.z = w[4 * q + 2]; v.w = w[4 * q + 3];
        reinterpret_cast<k3r_q4*>(p)[q] = v;
    }
}
template <int CAP>
PG_HD void k3r_clear(unsigned (&d)[CAP])
{   // cleardelta, gfreq.cc:556
    d[0] = 0;
#pragma unroll
    for (int k = 1; k < CAP; ++k) d[k] = K3R_TERM;
}
template <int CAP, int MODE>
PG_HD void k3r_reset(int* p)
{   // reset<DPunit_hf / _pf> (dpunit.cc): val = NEVSEL, cleared lists
    K3RRec<CAP, MODE> r;
    r.val = K3_NEVSEL; r.ptr = 0; r.dg = 0;
    k3r_clear(r.la);
    if constexpr (MODE == 2) k3r_clear(r.lb); else r.lb[0] = 0;
    k3r_store(p, r);
}
// whole-record copy (parking, rings, hand-over): st words, 16-byte aligned
PG_HD void k3r_copy_words(int* d, const int* s, int st)
{
    for (int w = 0; w < st; w += 4) *reinterpret_cast<k3r_q4*>(d + w) = *reinterpret_cast<const k3r_q4*>(s + w);
}

// ---- list primitives -----------------------------------------------------------------------------------
// gaplensd (gfreq.h:67-71): gl + nins of the segment of dl that holds gl
template <int CAP>
PG_HD int k3r_adj(const unsigned (&d)[CAP], int gl)
{
    const unsigned key = ((unsigned)gl << 16) | 0xffffu;
    unsigned r = d[0];
#pragma unroll
    for (int k = 1; k < CAP; ++k) r = d[k] <= key ? d[k] : r;
    return gl + (int)(r & 0xffffu);
}
template <int CAP>
PG_HD int k3r_nins(const unsigned (&d)[CAP], int gl)
{
    const unsigned key = ((unsigned)gl << 16) | 0xffffu;
    unsigned r = d[0];
#pragma unroll
    for (int k = 1; k < CAP; ++k) r = d[k] <= key ? d[k] : r;
    return (int)(r & 0xffffu);
}
// newgap(cf, dlc, df, dld), gfreq.cc:507-521: for every entry of df in turn, the first entry of cf whose adjusted
// length reaches df's; adjusted lengths ascend in both lists, so "first from where the last search stopped" is
// "first overall", and once cf is exhausted nothing more is added (padding: adjusted 65,535, frequency 0)
template <int CAP, int CS>
PG_HD double k3r_newgap4(const K3RList<CS>& cf, const unsigned (&dlc)[CAP], const K3RList<CS>& df, const unsigned (&dld)[CAP])
{
    int c[CS], j[CS];
#pragma unroll
    for (int i = 0; i < CS; ++i) { c[i] = k3r_adj(dlc, cf.g[i]); j[i] = k3r_adj(dld, df.g[i]); }
    double g = 0;
#pragma unroll
    for (int id = 0; id < CS; ++id) {
        double f = 0;
#pragma unroll
        for (int ic = CS - 1; ic >= 0; --ic) f = c[ic] >= j[id] ? cf.f[ic] : f;
        g += f * df.f[id];
    }
    return g;
}
// PwdM::newgap1(acf, dla, glb), maln.h:288-291 with newgap(cf, dlc, j) gfreq.cc:523-531, newgapc maln.h:270
template <int CAP, int CS>
PG_HD double k3r_newgap1(double wgop, const K3RList<CS>& acf, const unsigned (&dla)[CAP], int glb)
{
    double f = 0;               // several entries: the first whose adjusted length reaches glb, else wgop * 0.0
#pragma unroll
    for (int i = CS - 1; i >= 0; --i) f = (acf.g[i] != K3R_PAD && k3r_adj(dla, acf.g[i]) >= glb) ? acf.f[i] : f;
    const double multi = wgop * f;
    const double single = ((int)(dla[0] & 0xffffu) + acf.g[0] >= glb) ? wgop * acf.f[0] : 0.0;
    if (acf.g[0] == K3R_PAD) return 0.0;
    return (CS > 1 && acf.g[CS > 1 ? 1 : 0] != K3R_PAD) ? multi : single;
}
// PwdM::newgap2(adf, glb, dla), maln.h:297-300 with newgap(df, i, dld) gfreq.cc:533-544, newgapd maln.h:278
template <int CAP, int CS>
PG_HD double k3r_newgap2(double wgop, const K3RList<CS>& adf, int glb, const unsigned (&dla)[CAP])
{
    double g = 0;               // several entries: the leading entries whose adjusted length stays within glb
    bool on = true;
#pragma unroll
    for (int i = 0; i < CS; ++i) {
        on = on && adf.g[i] != K3R_PAD && !(glb < k3r_adj(dla, adf.g[i]));
        g += on ? adf.f[i] : 0.0;
    }
    const double multi = wgop * g;
    const double single = (glb >= (int)(dla[0] & 0xffffu) + adf.g[0]) ? wgop * adf.f[0] : 0.0;
    if (adf.g[0] == K3R_PAD) return 0.0;
    return (CS > 1 && adf.g[CS > 1 ? 1 : 0] != K3R_PAD) ? multi : single;
}
// newdelta(dlt, df, dln), gfreq.cc:567-584: filter the dynamic gap state through a column's static gap lengths
template <int CAP, int CS>
PG_HD void k3r_newdelta(unsigned (&out)[CAP], const int (&g)[CS], const unsigned (&dln)[CAP])
{
    static_assert(CAP >= CS + 2, "a rewritten list holds up to CS + 1 entries and the terminator");
    unsigned o[CAP];
#pragma unroll
    for (int k = 0; k < CAP; ++k) o[k] = K3R_TERM;
    int tg = 0, tn = 0, cnt = 0;
#pragma unroll
    for (int e = 0; e < CS; ++e) {
        const int gl = g[e];
        const int nins = k3r_nins(dln, gl);             // padding looks up the terminator: 0, never a jump
        const bool jump = nins > tn;
        const unsigned cur = ((unsigned)tg << 16) | (unsigned)tn;
#pragma unroll
        for (int k = 0; k <= e; ++k) o[k] = (jump && cnt == k) ? cur : o[k];
        cnt += jump ? 1 : 0;
        tg = jump ? gl + 1 : tg;
        tn = jump ? nins : tn;
    }
    const unsigned last = ((unsigned)tg << 16) | (unsigned)tn;
#pragma unroll
    for (int k = 0; k <= CS; ++k) o[k] = cnt == k ? last : o[k];
#pragma unroll
    for (int k = 0; k < CAP; ++k) out[k] = o[k];
}
// incdelta(dlt, dln, n), gfreq.cc:595-602
template <int CAP>
PG_HD void k3r_incdelta(unsigned (&out)[CAP], const unsigned (&dln)[CAP], int n)
{
#pragma unroll
    for (int k = 0; k < CAP; ++k) out[k] = dln[k] < K3R_TERM ? dln[k] + (unsigned)n : dln[k];
}

// ---- gapopen / update (fwd2c.cc:152-161,163-182 DPunit_hf; :202-212,214-233 DPunit_pf) -----------------------
// The static lists a cell needs, by direction d3 (0 diagonal, > 0 vertical, < 0 horizontal):
//   DPunit_hf   d3 = 0: a.tfq   d3 > 0: a.sfq   d3 < 0: a.rfq                    update: a.tfq lengths (d3 >= 0)
//   DPunit_pf   d3 = 0: a.sfq x b.tfq + b.sfq x a.tfq   d3 > 0: a.sfq x b.rfq   d3 < 0: b.sfq x a.rfq
//               update: a.tfq lengths (d3 >= 0), b.tfq lengths (d3 <= 0)
template <int CAP, int CS, int MODE>
struct K3RCellLists {       // what one direction of one cell reads from the two column blocks
    K3RList<CS> x, y;       // see k3r_fetch
    K3RList<CS> x2, y2;     // DPunit_pf, d3 = 0: the second product
    int ta[CS], tb[CS];     // tfq lengths of a / b for the rewrite
};
template <int CAP, int CS, int MODE, int D3>
PG_HD void k3r_fetch(K3RCellLists<CAP, CS, MODE>& L, const int* ablk, const int* bblk)
{
    if constexpr (MODE == 1) {
        k3r_load_list(L.x, ablk, D3 == 0 ? 1 : (D3 > 0 ? 0 : 2));
        if (D3 >= 0) k3r_load_glen(L.ta, ablk, 1);
    } else {
        if (D3 == 0) {
            k3r_load_list(L.x, ablk, 0); k3r_load_list(L.y, bblk, 1);       // a.sfq x b.tfq
            k3r_load_list(L.x2, bblk, 0); k3r_load_list(L.y2, ablk, 1);     // b.sfq x a.tfq
#pragma unroll
            for (int i = 0; i < CS; ++i) { L.ta[i] = L.y2.g[i]; L.tb[i] = L.y.g[i]; }
        } else if (D3 > 0) {
            k3r_load_list(L.x, ablk, 0); k3r_load_list(L.y, bblk, 2);       // a.sfq x b.rfq
            k3r_load_glen(L.ta, ablk, 1);
        } else {
            k3r_load_list(L.x, bblk, 0); k3r_load_list(L.y, ablk, 2);       // b.sfq x a.rfq
            k3r_load_glen(L.tb, bblk, 1);
        }
    }
}
template <int CAP, int CS, int MODE, int D3>
PG_HD double k3r_gapopen(const K3Prm& p, const K3RCellLists<CAP, CS, MODE>& L, const K3RRec<CAP, MODE>& r)
{
    if constexpr (MODE == 1) {
        const int glb = (int)((unsigned)r.dg >> 8);
        if (D3 == 0) return k3r_newgap2<CAP, CS>(p.wgop, L.x, glb, r.la);
        if (D3 > 0) return k3r_newgap1<CAP, CS>(p.wgop, L.x, r.la, glb);
        return k3r_newgap2<CAP, CS>(p.wgop, L.x, glb, r.la);
    } else {
    if (D3 == 0) return k3r_newgap4<CAP, CS>(L.x, r.la, L.y, r.lb) * p.bgop + k3r_newgap4<CAP, CS>(L.x2, r.lb, L.y2, r.la) * p.bgop;
    if (D3 > 0) return k3r_newgap4<CAP, CS>(L.x, r.la, L.y, r.lb) * p.bgop;
    return k3r_newgap4<CAP, CS>(L.x, r.lb, L.y, r.la) * p.bgop;
    }
}
template <int CAP, int CS, int MODE, int D3>
PG_HD void k3r_update(K3RRec<CAP, MODE>& dst, const K3RRec<CAP, MODE>& src, const K3RCellLists<CAP, CS, MODE>& L, double gpn)
{
    const int sdir = src.dg & 0xff;
    int dir;
    if (D3 > 0) dir = k3_ishori(sdir) ? K3_NEWV : K3_VERT;
    else if (D3 < 0) dir = k3_isvert(sdir) ? K3_NEWH : K3_HORI;
    else dir = k3_isdiag(sdir) ? K3_DIAG : K3_NEWD;
    int glb = 0;
    K3RRec<CAP, MODE> o;
    if constexpr (MODE == 1) {
        if (D3 > 0) glb = (int)((unsigned)src.dg >> 8) + 1;
        if (D3 >= 0) k3r_newdelta<CAP, CS>(o.la, L.ta, src.la);
        else k3r_incdelta<CAP>(o.la, src.la, 1);
        o.lb[0] = 0;
    } else {
        if (D3 >= 0) k3r_newdelta<CAP, CS>(o.la, L.ta, src.la); else k3r_incdelta<CAP>(o.la, src.la, 1);
        if (D3 <= 0) k3r_newdelta<CAP, CS>(o.lb, L.tb, src.lb); else k3r_incdelta<CAP>(o.lb, src.lb, 1);
    }
    o.val = src.val + gpn;
    o.ptr = src.ptr;
    o.dg = dir | (glb << 8);
    dst = o;
}

// ---- the cell in pieces (k3_core.cuh: k3_part_diag / _vert / _hori / k3_combine), records in shared memory ---------
// ablk / bblk: the column blocks of a at row m (staged index m + 1) and of b at column n (index n + 1)
template <int CAP, int CS, int MODE>
PG_HD void k3r_part_diag(const K3Prm& p, const int* ablk, const int* bblk, double dab, const int* hdiag, int* hout)
{
    K3RCellLists<CAP, CS, MODE> L;
    k3r_fetch<CAP, CS, MODE, 0>(L, ablk, bblk);
    K3RRec<CAP, MODE> s, o;
    k3r_load(s, hdiag);
    const double gop = k3r_gapopen<CAP, CS, MODE, 0>(p, L, s);
    k3r_update<CAP, CS, MODE, 0>(o, s, L, dab + gop);
    k3r_store(hout, o);
}
template <int CAP, int CS, int MODE>
PG_HD void k3r_part_vert(const K3Prm& p, const int* ablk, const int* bblk, bool a_nils, bool first_row, double* pua,
                         const int* habove, const int* gabove, const int* g2above, int* gout, int* g2out, const int* black, int st)
{
    if (first_row) {            // the G rows keep their untouched (black) records
        k3r_copy_words(gout, black, st);
        if (p.Noll == 3) k3r_copy_words(g2out, black, st);
        return;
    }
    if (a_nils) *pua = k3r_blk_cfq(ablk) * k3r_blk_efq(bblk) * -p.u;
    K3RCellLists<CAP, CS, MODE> L;
    k3r_fetch<CAP, CS, MODE, 1>(L, ablk, bblk);
    K3RRec<CAP, MODE> h, g, o;
    k3r_load(h, habove); k3r_load(g, gabove);
    double gnp = k3r_gapopen<CAP, CS, MODE, 1>(p, L, g);
    double gop = k3r_gapopen<CAP, CS, MODE, 1>(p, L, h);
    const bool hv = !k3_isvert(h.dg & 0xff);
    if (hv && (h.val + gop > g.val + gnp)) k3r_update<CAP, CS, MODE, 1>(o, h, L, gop);
    else k3r_update<CAP, CS, MODE, 1>(o, g, L, gnp);
    o.val = o.val + *pua;
    k3r_store(gout, o);
    if (p.Noll == 3) {          // vertical2 (fwd2c.h:411-420)
        K3RRec<CAP, MODE> g2;
        k3r_load(g2, g2above);
        gnp = p.v2divv1 * k3r_gapopen<CAP, CS, MODE, 1>(p, L, g2);
        gop = p.v2divv1 * gop;
        if (hv && (h.val + gop > g2.val + gnp)) k3r_update<CAP, CS, MODE, 1>(o, h, L, gop);
        else k3r_update<CAP, CS, MODE, 1>(o, g2, L, gnp);
        o.val = o.val + p.u2divu1 * *pua;
        k3r_store(g2out, o);
    }
}
template <int CAP, int CS, int MODE>
PG_HD void k3r_part_hori(const K3Prm& p, const int* ablk, const int* bblk, bool first_col, const int* hleft, int* f1, int* f2)
{
    if (first_col) return;
    const double pub = k3r_blk_cfq(bblk) * k3r_blk_efq(ablk) * -p.u;
    K3RCellLists<CAP, CS, MODE> L;
    k3r_fetch<CAP, CS, MODE, -1>(L, ablk, bblk);
    K3RRec<CAP, MODE> h, f, o;
    k3r_load(h, hleft); k3r_load(f, f1);
    double gnp = k3r_gapopen<CAP, CS, MODE, -1>(p, L, f);
    double gop = k3r_gapopen<CAP, CS, MODE, -1>(p, L, h);
    const bool hh = !k3_ishori(h.dg & 0xff);
    if (hh && (h.val + gop > f.val + gnp)) k3r_update<CAP, CS, MODE, -1>(o, h, L, gop);
    else k3r_update<CAP, CS, MODE, -1>(o, f, L, gnp);
    o.val = o.val + pub;
    k3r_store(f1, o);
    if (p.Noll == 3) {          // horizontal2 (fwd2c.h:433-442)
        K3RRec<CAP, MODE> fb;
        k3r_load(fb, f2);
        gnp = p.v2divv1 * k3r_gapopen<CAP, CS, MODE, -1>(p, L, fb);
        gop = p.v2divv1 * gop;
        if (hh && (h.val + gop > fb.val + gnp)) k3r_update<CAP, CS, MODE, -1>(o, h, L, gop);
        else k3r_update<CAP, CS, MODE, -1>(o, fb, L, gnp);
        o.val = o.val + p.u2divu1 * pub;
        k3r_store(f2, o);
    }
}
// selection (fwd2c.h:409-453): G first, G2 beats it on >, F1 and F2 on >=; the gap state replaces the diagonal only
// if strictly better.  Returns true when the cell appends a path record (:465-467).
PG_HD bool k3r_combine(const K3Prm& p, bool first_row, bool first_col, int* hout, const int* gout, const int* g2out,
                       const int* f1, const int* f2, int st)
{
    const int* mx = gout;
    double mv = k3_val(gout);
    if (!first_row && p.Noll == 3) { const double v = k3_val(g2out); if (v > mv) { mx = g2out; mv = v; } }
    if (!first_col) {
        const double v1 = k3_val(f1);
        if (v1 >= mv) { mx = f1; mv = v1; }
        if (p.Noll == 3) { const double v2 = k3_val(f2); if (v2 >= mv) { mx = f2; mv = v2; } }
    }
    if (mv > k3_val(hout)) k3r_copy_words(hout, mx, st);
    const int dir = hout[3] & 0xff;
    return dir == K3_NEWD || dir == K3_NEWV || dir == K3_NEWH;
}
// the whole cell on one thread (host emulation; one-thread-per-row kernels)
template <int CAP, int CS, int MODE>
PG_HD bool k3r_cell(const K3Prm& p, const int* ablk, const int* bblk, bool a_nils, bool first_row, bool first_col, double dab,
                    double* pua, const int* hdiag, const int* habove, const int* gabove, const int* g2above, const int* hleft,
                    int* f1, int* f2, int* hout, int* gout, int* g2out, const int* black, int st)
{
    k3r_part_diag<CAP, CS, MODE>(p, ablk, bblk, dab, hdiag, hout);
    k3r_part_vert<CAP, CS, MODE>(p, ablk, bblk, a_nils, first_row, pua, habove, gabove, g2above, gout, g2out, black, st);
    k3r_part_hori<CAP, CS, MODE>(p, ablk, bblk, first_col, hleft, f1, f2);
    return k3r_combine(p, first_row, first_col, hout, gout, g2out, f1, f2, st);
}

// Boundary cells (initB, fwd2c.h:138-176).  Row: asi at a.left - 1 (a block 0), k-th column (b block k)
template <int CAP, int CS, int MODE>
PG_HD void k3r_boundary_row(const K3Prm& p, const int* ablk0, const int* bblk, int k, int* dst, const int* src)
{
    const double pub = k3r_blk_cfq(bblk) * k3r_blk_efq(ablk0) * -p.u;
    K3RCellLists<CAP, CS, MODE> L;
    k3r_fetch<CAP, CS, MODE, -1>(L, ablk0, bblk);
    K3RRec<CAP, MODE> s, o;
    k3r_load(s, src);
    double gnp = k3r_gapopen<CAP, CS, MODE, -1>(p, L, s);
    gnp = (k < p.codonk1) ? gnp + pub : (p.v2divv1 * gnp + p.u2divu1 * pub);
    k3r_update<CAP, CS, MODE, -1>(o, s, L, gnp);
    k3r_store(dst, o);
}
// Column: bsi at b.left - 1 (b block 0), k-th row (a block k)
template <int CAP, int CS, int MODE>
PG_HD void k3r_boundary_col(const K3Prm& p, const int* ablk, const int* bblk0, int k, int* dst, const int* src)
{
    const double pua = k3r_blk_cfq(ablk) * k3r_blk_efq(bblk0) * -p.u;
    K3RCellLists<CAP, CS, MODE> L;
    k3r_fetch<CAP, CS, MODE, 1>(L, ablk, bblk0);
    K3RRec<CAP, MODE> s, o;
    k3r_load(s, src);
    double gnp = k3r_gapopen<CAP, CS, MODE, 1>(p, L, s);
    gnp = (k < p.codonk1) ? gnp + pua : (p.v2divv1 * gnp + p.u2divu1 * pua);
    k3r_update<CAP, CS, MODE, 1>(o, s, L, gnp);
    k3r_store(dst, o);
}
