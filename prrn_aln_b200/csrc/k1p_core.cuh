// k1p_core.cuh -- packed (int16 x 2) form of the K1 recurrence: every lane computes TWO alignments
// at once, the low and high halves of each 32-bit register.  Both halves share the subject
// (columns) and the step counter; they differ in the query (rows), its length, band and result row.
// One __viaddmax_s16x2 / __vimax_s16x2 retires the work of two cells, so the DPX issue rate buys
// twice the cells of k1_core.cuh.  Same drifted formulation, same band "poke" per half.
//
// Range: values live in int16.  K1P_NEG is the -infinity; k1p_fits() is the host-side proof
// obligation that no real value leaves (K1P_NEG + growth, 32767) for a given batch.
#pragma once
#include "k1_core.cuh"

#define K1P_NEG (-30000)

PG_HD unsigned k1p_pack(int lo, int hi) { return ((unsigned)lo & 0xffffu) | ((unsigned)hi << 16); }
PG_HD int k1p_lo(unsigned x) { return (int)(short)(x & 0xffffu); }
PG_HD int k1p_hi(unsigned x) { return (int)(short)(x >> 16); }

#if defined(__CUDA_ARCH__)
#define K1P_ADDMAX(a, b, c) __viaddmax_s16x2((a), (b), (c))
#define K1P_MAX(a, b) __vmaxs2((a), (b))
#else
PG_HD unsigned k1p_addmax_host(unsigned a, unsigned b, unsigned c)
{
    short l = (short)(k1p_lo(a) + k1p_lo(b)), h = (short)(k1p_hi(a) + k1p_hi(b));     // wrap like the hardware
    int lo = l > k1p_lo(c) ? l : k1p_lo(c), hi = h > k1p_hi(c) ? h : k1p_hi(c);
    return k1p_pack(lo, hi);
}
PG_HD unsigned k1p_max_host(unsigned a, unsigned b)
{
    int lo = k1p_lo(a) > k1p_lo(b) ? k1p_lo(a) : k1p_lo(b), hi = k1p_hi(a) > k1p_hi(b) ? k1p_hi(a) : k1p_hi(b);
    return k1p_pack(lo, hi);
}
#define K1P_ADDMAX(a, b, c) k1p_addmax_host((a), (b), (c))
#define K1P_MAX(a, b) k1p_max_host((a), (b))
#endif

// 16-bit boundary values (same formulas as k1_top / k1_left with the 16-bit -infinity)
PG_HD int k1p_top(const K1Geom& g, int n)
{
    int k = n + 1;
    if (k == 0) return 0;
    return k <= g.up ? g.topOpen + k * (g.topExt + g.u) : K1P_NEG;
}
PG_HD int k1p_left(const K1Geom& g, int m)
{
    int k = m + 1;
    if (k == 0) return 0;
    return -k >= g.lw ? g.leftOpen + k * (g.leftExt + g.u) : K1P_NEG;
}

// Can the batch run in 16 bits?  smax/smin = extreme S' = S + 2u over residues present, lmax = the
// longest window.  Real drifted values lie in [-(3v + lmax*max(0,-smin)), max(0,smax)*lmax]; the
// poisoned out-of-band region may climb from K1P_NEG by at most max(0,smax)*lmax.
PG_HD bool k1p_fits(int smax, int smin, int v, int lmax)
{
    long long up = (long long)(smax > 0 ? smax : 0) * lmax;
    long long dn = 3LL * v + (long long)(smin < 0 ? -smin : 0) * lmax;
    return up + 64 < 32000 && -30000 + up + 64 < -dn - v && -30000 - v - 64 > -32768;
}

template <int R>
struct K1PLane {
    unsigned H[R];
    unsigned E[R];
    unsigned hdiag;
};

template <int R>
PG_HD void k1p_lane_init(K1PLane<R>& L, const K1Geom& g0, const K1Geom& g1, int mbase, unsigned negv2)
{
    const unsigned neg2 = k1p_pack(K1P_NEG, K1P_NEG);
#pragma unroll
    for (int k = 0; k < R; ++k) {
        unsigned h = k1p_pack(k1p_left(g0, mbase + k), k1p_left(g1, mbase + k));
        L.H[k] = h;
        L.E[k] = K1P_ADDMAX(h, negv2, neg2);
    }
    L.hdiag = k1p_pack(k1p_left(g0, mbase - 1), k1p_left(g1, mbase - 1));
}

template <int R>
PG_HD void k1p_lane_step(K1PLane<R>& L, unsigned* sc, unsigned negv2, unsigned h_up, unsigned f_up,
                         unsigned* h_dn, unsigned* f_dn)
{
    // phase 1 (independent per row): t_k = max(diag_k + S'_k, E_k), written over sc[k]; the old
    // H[k] is consumed here, so phase 2 can write the new H[k] in place without register moves
    sc[0] = K1P_ADDMAX(L.hdiag, sc[0], L.E[0]);
#pragma unroll
    for (int k = 1; k < R; ++k) sc[k] = K1P_ADDMAX(L.H[k - 1], sc[k], L.E[k]);
    // phase 2: the vertical chain (one dependent instruction per row) and the eager horizontal state
    unsigned f = f_up;
    unsigned h = h_up;
#pragma unroll
    for (int k = 0; k < R; ++k) {
        const unsigned t = sc[k];
        h = K1P_MAX(t, f);
        f = K1P_ADDMAX(t, negv2, f);
        L.H[k] = h;
        L.E[k] = K1P_ADDMAX(h, negv2, L.E[k]);
    }
    L.hdiag = h_up;
    *h_dn = h;
    *f_dn = f;
}

// Host emulation of one warp computing two alignments (q0 x s, q1 x s) in the packed form.
#if !defined(__CUDA_ARCH__)
template <int R>
static inline void k1p_emulate_pair(const uint8_t* q0, const uint8_t* q1, const uint8_t* s, const K1Geom& g0,
                                    const K1Geom& g1, const int* mtx, int dim, int* res0, int* res1)
{
    const int T = 32, rpp = T * R;
    const int LS = g0.LS;
    const unsigned negv2 = k1p_pack(-g0.v, -g0.v), neg2 = k1p_pack(K1P_NEG, K1P_NEG);
    const int LQ = g0.LQ > g1.LQ ? g0.LQ : g1.LQ;
    unsigned* rowH = new unsigned[LS + 1];
    unsigned* rowF = new unsigned[LS + 1];
    int r0 = 0, r1 = 0;
    for (int pass = 0; pass * rpp < LQ; ++pass) {
        const int pbase = pass * rpp;
        K1PLane<R> L[T];
        unsigned send_h[2][T], send_f[2][T];
        for (int t = 0; t < T; ++t) k1p_lane_init(L[t], g0, g1, pbase + t * R, negv2);
        const int rows_here = LQ - pbase < rpp ? LQ - pbase : rpp;
        const int lanes = (rows_here + R - 1) / R;
        for (int step = 0; step < LS + lanes - 1; ++step) {
            const int cur = step & 1, prv = cur ^ 1;
            for (int t = 0; t < lanes; ++t) {
                const int n = step - t;
                if (n < 0 || n >= LS) continue;
                const int mbase = pbase + t * R;
                unsigned h_up, f_up;
                if (t == 0) {
                    if (pass == 0) { h_up = k1p_pack(k1p_top(g0, n), k1p_top(g1, n)); f_up = K1P_ADDMAX(h_up, negv2, neg2); }
                    else { h_up = rowH[n]; f_up = rowF[n]; }
                } else { h_up = send_h[prv][t - 1]; f_up = send_f[prv][t - 1]; }
                int kL, kU;
                k1_poke_rows(g0, mbase, n, &kL, &kU);
                if (kL >= 0 && kL < R) L[t].E[kL] = k1p_pack(K1P_NEG, k1p_hi(L[t].E[kL]));
                if (kU >= 0 && kU < R) L[t].E[kU] = k1p_pack(K1P_NEG, k1p_hi(L[t].E[kU]));
                k1_poke_rows(g1, mbase, n, &kL, &kU);
                if (kL >= 0 && kL < R) L[t].E[kL] = k1p_pack(k1p_lo(L[t].E[kL]), K1P_NEG);
                if (kU >= 0 && kU < R) L[t].E[kU] = k1p_pack(k1p_lo(L[t].E[kU]), K1P_NEG);
                unsigned sc[R];
                for (int k = 0; k < R; ++k) {
                    int m = mbase + k;
                    int a = m < g0.LQ ? mtx[q0[m] * dim + s[n]] + 2 * g0.u : 0;
                    int b = m < g1.LQ ? mtx[q1[m] * dim + s[n]] + 2 * g0.u : 0;
                    sc[k] = k1p_pack(a, b);
                }
                unsigned h_dn, f_dn;
                k1p_lane_step(L[t], sc, negv2, h_up, f_up, &h_dn, &f_dn);
                send_h[cur][t] = h_dn; send_f[cur][t] = f_dn;
                if (t == T - 1) { rowH[n] = h_dn; rowF[n] = f_dn; }
            }
        }
        if (g0.LQ > pbase && g0.LQ <= pbase + rpp) { int r = g0.LQ - 1 - pbase; r0 = k1p_lo(L[r / R].H[r % R]); }
        if (g1.LQ > pbase && g1.LQ <= pbase + rpp) { int r = g1.LQ - 1 - pbase; r1 = k1p_hi(L[r / R].H[r % R]); }
    }
    delete[] rowH; delete[] rowF;
    *res0 = r0 - (g0.LQ + LS) * g0.u;
    *res1 = r1 - (g1.LQ + LS) * g1.u;
}
#endif
