// k1_core.cuh -- the per-lane recurrence of kernel K1 (batched score-only banded affine fill).
//
// Reference semantics: Fwd2d::Fwd2d / forwardD (src/fwd2d1.cc:57-90, 136-160): global affine Gotoh
// score over the stripe() band (src/aln2.cc:156-174).  This file is shared between the CUDA kernel
// (k1_score.cu) and a host emulation of one warp (tests/host_emul) so that the geometry, boundary
// and band logic can be checked against the oracle on a machine without a GPU.
//
// Formulation (exact in integers, see DESIGN.md "K1"):
//   kernel orientation: Q = rows (LQ residues, relative index m), S = columns (LS residues, index n),
//   relative diagonal r = n - m, band lw <= r <= up (stripe() shifted by r0 = S.left - Q.left).
//   Drifted values  Ht(m,n) = H(m,n) + (m+n+2)*u  turn "extend a gap" into a no-op:
//       Et(m,n) = max(Ht(m,n-1) - v, Et(m,n-1))          horizontal state  (reference ff[])
//       Ft(m,n) = max(Ht(m-1,n) - v, Ft(m-1,n))          vertical state    (reference gg[])
//       Ht(m,n) = max(Ht(m-1,n-1) + S'(q_m,s_n), Et, Ft) with S' = S + 2u
//   4 DPX-class instructions per cell: viaddmax x3 + max.
//   Band: a path leaves the band only through a horizontal move onto diagonal up+1 or onto
//   diagonal lw.  Clearing the (eagerly computed) Et of exactly those two cells per row ("poke")
//   makes every out-of-band cell that an in-band cell can read equal to -inf, which is what the
//   reference's sentinels hh[lw-1] = hh[up+1] = NEG_INT (fwd2d1.cc:76,85) do.
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define PG_HD __host__ __device__ __forceinline__
#else
#define PG_HD static inline
#endif

#define K1_NEG (-(1 << 30))

#if defined(__CUDA_ARCH__)
#define K1_ADDMAX(a, b, c) __viaddmax_s32((a), (b), (c))
#define K1_MAX(a, b) max((a), (b))
#else
PG_HD int k1_addmax_host(int a, int b, int c) { int t = a + b; return t > c ? t : c; }
#define K1_ADDMAX(a, b, c) k1_addmax_host((a), (b), (c))
#define K1_MAX(a, b) ((a) > (b) ? (a) : (b))
#endif

// Geometry and integer penalties of one pair in kernel orientation.
struct K1Geom {
    int LQ, LS;         // rows, columns
    int lw, up;         // relative band  lw <= n - m <= up
    int u, v;           // integer extension / open penalties (uu, vv of fwd2d1.cc:62-63)
    int topOpen, topExt;    // boundary row (columns consumed before any row): -(v,u)*f(Q end flags)
    int leftOpen, leftExt;  // boundary column: -(v,u)*f(S end flags)
};

// stripe() (src/aln2.cc:156-174) in relative coordinates (r - r0): depends on lengths only.
PG_HD void k1_band(int LQ, int LS, int sh, int* lw, int* up)
{
    if (sh < 0) {
        int shorter = LQ < LS ? LQ : LS;
        sh = -sh * shorter / 100;
    }
    int hi = LS - LQ, lo = 0;
    if (hi < lo) { int t = hi; hi = lo; lo = t; }
    hi += sh;
    lo -= sh;
    if (LS < hi) hi = LS;
    if (-LQ > lo) lo = -LQ;
    *lw = lo;
    *up = hi;
}

// cells the reference visits
PG_HD long long k1_cells(int LQ, int LS, int lw, int up)
{
    long long c = 0;
    for (int m = 0; m < LQ; ++m) {
        int n0 = m + lw > 0 ? m + lw : 0;
        int n9 = m + up + 1 < LS ? m + up + 1 : LS;
        if (n9 > n0) c += n9 - n0;
    }
    return c;
}

// drifted boundary row value Ht(-1, n), n >= -1
PG_HD int k1_top(const K1Geom& g, int n)
{
    int k = n + 1;
    if (k == 0) return 0;
    return k <= g.up ? g.topOpen + k * (g.topExt + g.u) : K1_NEG;
}
// drifted boundary column value Ht(m, -1), m >= -1
PG_HD int k1_left(const K1Geom& g, int m)
{
    int k = m + 1;
    if (k == 0) return 0;
    return -k >= g.lw ? g.leftOpen + k * (g.leftExt + g.u) : K1_NEG;
}

template <int R>
struct K1Lane {
    int H[R];   // Ht(mbase+k, n-1): previous column
    int E[R];   // Et(mbase+k, n):   eagerly computed horizontal state for the coming column
    int hdiag;  // Ht(mbase-1, n-1)
};

template <int R>
PG_HD void k1_lane_init(K1Lane<R>& L, const K1Geom& g, int mbase)
{
#pragma unroll
    for (int k = 0; k < R; ++k) {
        int h = k1_left(g, mbase + k);
        L.H[k] = h;
        L.E[k] = K1_ADDMAX(h, -g.v, K1_NEG);
    }
    L.hdiag = k1_left(g, mbase - 1);
}

// One column for one lane.  sc[k] = S'(q[mbase+k], s[n]).  (h_up, f_up) = Ht(mbase-1, n) and the
// vertical state entering row mbase.  Returns the pair to hand to the lane below.
template <int R>
PG_HD void k1_lane_step(K1Lane<R>& L, const int* sc, int negv, int h_up, int f_up, int* h_dn, int* f_dn)
{
    int diag = L.hdiag;
    int f = f_up;
    int h = h_up;
#pragma unroll
    for (int k = 0; k < R; ++k) {
        int t = K1_ADDMAX(diag, sc[k], L.E[k]);      // max(diagonal, horizontal)
        h = K1_MAX(t, f);                            // Ht(m, n)
        f = K1_ADDMAX(t, negv, f);                   // vertical state entering row m+1
        diag = L.H[k];
        L.H[k] = h;
        L.E[k] = K1_ADDMAX(h, negv, L.E[k]);         // horizontal state for column n+1
    }
    L.hdiag = h_up;
    *h_dn = h;
    *f_dn = f;
}

// Band "poke": before the column n is processed, the row whose cell (m, n) lies on diagonal lw
// (first in-band cell of that row) or on diagonal up+1 (first cell past the band) loses its
// horizontal input.  kL / kU are the row indices inside this lane's strip (may be out of [0,R)).
PG_HD void k1_poke_rows(const K1Geom& g, int mbase, int n, int* kL, int* kU)
{
    *kL = n - g.lw - mbase;
    *kU = n - g.up - 1 - mbase;
}

// Host-side emulation of one warp (32 lanes x R rows, multi-pass over Q): the control flow of the
// CUDA kernel with shuffles replaced by arrays.  mtx is dim x dim integer scores.
#if !defined(__CUDA_ARCH__)
template <int R>
static inline int k1_emulate_pair(const uint8_t* q, const uint8_t* s, const K1Geom& g, const int* mtx, int dim)
{
    const int T = 32;
    const int rows_per_pass = T * R;
    const int negv = -g.v;
    int* rowH = new int[g.LS > 0 ? g.LS : 1];   // bottom row of the previous pass, per column
    int* rowF = new int[g.LS > 0 ? g.LS : 1];
    int result = 0;
    if (g.LQ == 0 || g.LS == 0) {           // degenerate: score is the boundary value
        result = g.LQ == 0 ? k1_top(g, g.LS - 1) : k1_left(g, g.LQ - 1);
        delete[] rowH; delete[] rowF;
        return result - (g.LQ + g.LS) * g.u;
    }
    for (int pass = 0; pass * rows_per_pass < g.LQ; ++pass) {
        const int pbase = pass * rows_per_pass;
        K1Lane<R> L[T];
        int send_h[2][T], send_f[2][T];
        for (int t = 0; t < T; ++t) k1_lane_init(L[t], g, pbase + t * R);
        const int rows_here = g.LQ - pbase < rows_per_pass ? g.LQ - pbase : rows_per_pass;
        const int lanes = (rows_here + R - 1) / R;
        for (int step = 0; step < g.LS + lanes - 1; ++step) {
            const int cur = step & 1, prv = cur ^ 1;
            for (int t = 0; t < lanes; ++t) {
                const int n = step - t;
                if (n < 0 || n >= g.LS) continue;
                const int mbase = pbase + t * R;
                int h_up, f_up;
                if (t == 0) {
                    if (pass == 0) { h_up = k1_top(g, n); f_up = K1_ADDMAX(h_up, negv, K1_NEG); }
                    else { h_up = rowH[n]; f_up = rowF[n]; }
                } else { h_up = send_h[prv][t - 1]; f_up = send_f[prv][t - 1]; }
                int kL, kU;
                k1_poke_rows(g, mbase, n, &kL, &kU);
                if (kL >= 0 && kL < R) L[t].E[kL] = K1_NEG;
                if (kU >= 0 && kU < R) L[t].E[kU] = K1_NEG;
                int sc[R];
                for (int k = 0; k < R; ++k) {
                    int m = mbase + k;
                    sc[k] = m < g.LQ ? mtx[q[m] * dim + s[n]] + 2 * g.u : 0;
                }
                int h_dn, f_dn;
                k1_lane_step(L[t], sc, negv, h_up, f_up, &h_dn, &f_dn);
                send_h[cur][t] = h_dn;
                send_f[cur][t] = f_dn;
                if (t == T - 1) { rowH[n] = h_dn; rowF[n] = f_dn; }
            }
        }
        if (pbase + rows_here == g.LQ) {
            int tl = (rows_here - 1) / R, kf = (rows_here - 1) % R;
            result = L[tl].H[kf];
        }
    }
    delete[] rowH; delete[] rowF;
    return result - (g.LQ + g.LS) * g.u;
}
#endif
