// k1f_core.cuh -- floating-point form of the K1 recurrence (score-only banded affine fill): the
// general path of alnScoreD for everything the exact-integer kernels do not take -- non-integral
// scoring (the reference's default PAM matrices, src/simmtx.cc:282-334), discounted / free terminal
// gaps (tgapf < 1, inex.exgl / exgr: the Fwd2d ctor src/fwd2d1.cc:57-90 and lastD :97-134) and the
// Smith-Waterman-Gotoh score (algmode.lcl & 16: swgforwardD :162-189).
//
// Every value is produced by the same IEEE operation, in the same order, as Fwd2d::forwardD
// (src/fwd2d1.cc:147-150):   ff = max(h_left - vv, ff_left) - uu ;  gg = max(h_up - vv, gg_up) - uu ;
// h = h_diag + S ;  h = max(max(h, ff), gg)   -- only add/sub/max appear (no FMA, no re-association),
// so float and double results are bit-identical to the reference's VTYPE arithmetic.
// T = float (aln build) or double (prrn build, -DDVAL=1).
//
// Shared between the CUDA kernel (k1f_score.cu) and the host emulation (tests/host_emul/k1f_emul.cc).
#pragma once
#include "k1_core.cuh"

// scratch lines are written with st.cg by other lanes of the warp: read them at L2 as well
#if defined(__CUDA_ARCH__)
#define K1F_LD(p) __ldcg(p)
#else
#define K1F_LD(p) (*(p))
#endif

// a product that must round on its own (nvcc would otherwise contract a*b + c into one FMA)
#if defined(__CUDA_ARCH__)
__device__ __forceinline__ float k1f_mul(float a, float b) { return __fmul_rn(a, b); }
__device__ __forceinline__ double k1f_mul(double a, double b) { return __dmul_rn(a, b); }
#else
static inline float k1f_mul(float a, float b) { return a * b; }
static inline double k1f_mul(double a, double b) { return a * b; }
#endif

template <typename T> PG_HD T k1f_neg() { return (T)-1.0e30; }     // the "-infinity" of this kernel

// NEVSEL of the reference (src/cmn.h:56-60): what swgforwardD returns when no cell exists
PG_HD float k1f_nevsel(float) { return -(3.402823466e+38f / 16 * 7); }
PG_HD double k1f_nevsel(double) { return -(1.7976931348623157e+308 / 16 * 7); }

// max of two finite values.  float on the device: the one-instruction form (FMNMX, FMNMX3 for two in a row) --
// `a > b ? a : b` compiles to a compare and a select because the two differ on NaN and on the sign of a zero; scores
// are never NaN, and +0 / -0 compare equal everywhere they are used.  Measured on B200, C2 with PAM250: 943 -> 1,384
// GCUPS end to end, same checksum.  double keeps compare + select: DMNMX runs on the FP64 pipe, which the five DADD
// and four DSETP of a cell already fill (602 -> 416 GCUPS with fmax).
#if defined(__CUDA_ARCH__)
__device__ __forceinline__ float k1f_max(float a, float b) { return fmaxf(a, b); }
__device__ __forceinline__ double k1f_max(double a, double b) { return a > b ? a : b; }
#else
template <typename T> static inline T k1f_max(T a, T b) { return a > b ? a : b; }
#endif

// Per-pair description in KERNEL orientation: Q = rows, S = columns, r = n - m.
template <typename T>
struct K1FPair {
    int LQ, LS, lw, up;
    const T* topTab;    // boundary row:    H(-1, n) = topTab[n + 1]   (columns consumed before any row)
    const T* leftTab;   // boundary column: H(m, -1) = leftTab[m + 1]
    // lastD (fwd2d1.cc:97-134): which end-gap relaxations run, with which factor, in which order
    int doQ, doS;       // Q exhausted (scan of the last row) / S exhausted (scan of the last column)
    T rtgQ, rtgS;       // (VTYPE)-exact factors  rtgapf  of the two blocks
    int qFirst;         // 1: the Q block is the reference's first block (Q is the reference's b)
    int originR;        // VD: relative diagonal recorded at the origin cell (-r0: the vclear quirk, :226)
};

// boundary values with the band applied
template <typename T> PG_HD T k1f_top(const K1FPair<T>& g, int n)
{
    const int k = n + 1;
    return k <= g.up ? g.topTab[k] : k1f_neg<T>();
}
template <typename T> PG_HD T k1f_left(const K1FPair<T>& g, int m)
{
    const int k = m + 1;
    return -k >= g.lw ? g.leftTab[k] : k1f_neg<T>();
}

// VD = the Fwd2d_vd variant (fwd2d1.cc:191-322): every value carries the (relative) diagonal on which
// its path left the boundary; selections use the reference's compare-and-copy tie rules.
template <typename T, int R, bool VD = false>
struct K1FLane {
    T H[R];     // H(mbase+k, n-1)
    T E[R];     // ff of the coming column (eager): max(H(m,n-1) - vv, ff(m,n-1)) - uu
    T hdiag;    // H(mbase-1, n-1)
    int Hr[VD ? R : 1], Er[VD ? R : 1], hdiagr;
};

template <typename T, int R, bool VD>
PG_HD void k1f_lane_init(K1FLane<T, R, VD>& L, const K1FPair<T>& g, int mbase, T vv, T uu)
{
#pragma unroll
    for (int k = 0; k < R; ++k) {
        const T h = k1f_left(g, mbase + k);
        L.H[k] = h;
        // ff[r-1] = NEVSEL on the boundary column: the eager state is (h - vv) - uu
        L.E[k] = VD ? (h - vv) - uu : k1f_max(h - vv, k1f_neg<T>()) - uu;
        if (VD) { L.Hr[k] = -(mbase + k + 1); L.Er[k] = -(mbase + k + 1); }
    }
    L.hdiag = k1f_left(g, mbase - 1);
    if (VD) L.hdiagr = mbase == 0 ? g.originR : -mbase;
}

// One column for one lane.  sc[k] = S(q[mbase+k], s[n]);  (h_up, g_up) = H(mbase-1, n), gg(mbase-1, n).
// Hands (H, gg) of its last row to the lane below.
// SWG (swgforwardD, fwd2d1.cc:162-189): cells are clamped at 0 and the maximum over the in-band cells
// is tracked; out-of-band cells are masked to -inf (rel = n - mbase - lw, span = up - lw).
// Two IEEE single-precision additions in one instruction (add.rn.f32x2, SASS FADD2: sm_100 packed FP32).  Each half
// rounds exactly like a scalar add.rn.f32, so results stay bit-identical; it is used where two rows of a lane add
// independently of the vertical chain (h_diag + S, and the final "- uu" of the eager ff): 7 -> 6 instructions per
// cell.  Measured on B200 (tools/bench_k1f.py, C2 with PAM250, float): 1,517 -> 1,553 GCUPS end to end, the 300-aa
// set 1,298 -> 1,325, same checksums (-DK1F_NO_PACKED_ADD builds the scalar form).
#if !defined(K1F_NO_PACKED_ADD)
#define K1F_PACKED_ADD 1
#endif
#if defined(__CUDA_ARCH__)
__device__ __forceinline__ void k1f_add2(float& r0, float& r1, float a0, float a1, float b0, float b1)
{
    asm("{ .reg .b64 pa, pb, pc;\n\t"
        "mov.b64 pa, {%2, %3};\n\t"
        "mov.b64 pb, {%4, %5};\n\t"
        "add.rn.f32x2 pc, pa, pb;\n\t"
        "mov.b64 {%0, %1}, pc; }"
        : "=f"(r0), "=f"(r1) : "f"(a0), "f"(a1), "f"(b0), "f"(b1));
}
#endif
template <typename T> struct k1f_is_float { static constexpr bool value = false; };
template <> struct k1f_is_float<float> { static constexpr bool value = true; };

template <typename T, int R, bool SWG, bool VD>
PG_HD void k1f_lane_step(K1FLane<T, R, VD>& L, T* sc, T vv, T uu, T h_up, T g_up, T* h_dn, T* g_dn, int rel,
                         unsigned span, T* maxh, int hr_up = 0, int gr_up = 0, int* hr_dn = nullptr,
                         int* gr_dn = nullptr)
{
    // phase 1 (independent per row): h_diag + S, consumed old H[k-1]
#if defined(__CUDA_ARCH__) && defined(K1F_PACKED_ADD)
    if constexpr (k1f_is_float<T>::value && !VD && !SWG && R >= 2) {
#pragma unroll
        for (int k = 0; k + 1 < R; k += 2)
            k1f_add2(sc[k], sc[k + 1], k ? L.H[k - 1] : L.hdiag, L.H[k], sc[k], sc[k + 1]);
        if (R & 1) sc[R - 1] = L.H[R - 2] + sc[R - 1];
        const T uv = uu + vv;
        (void)uv;
        T habove = h_up, gabove = g_up, h = h_up, g = g_up;
        T t[R];
#pragma unroll
        for (int k = 0; k < R; ++k) {
            const T f = L.E[k];
            g = k1f_max(habove - vv, gabove) - uu;               // :148
            h = k1f_max(k1f_max(sc[k], f), g);                   // :150
            t[k] = k1f_max(h - vv, f);
            L.H[k] = h;
            habove = h;
            gabove = g;
        }
        const T nu = -uu;
#pragma unroll
        for (int k = 0; k + 1 < R; k += 2) k1f_add2(L.E[k], L.E[k + 1], t[k], t[k + 1], nu, nu);
        if (R & 1) L.E[R - 1] = t[R - 1] - uu;
        L.hdiag = h_up;
        *h_dn = h;
        *g_dn = g;
        return;
    }
#endif
    sc[0] = L.hdiag + sc[0];                                     // fwd2d1.cc:149
#pragma unroll
    for (int k = 1; k < R; ++k) sc[k] = L.H[k - 1] + sc[k];
    // phase 2: the vertical chain
    const T uv = uu + vv;                                        // Fwd2d_vd subtracts (uu + vv) at once (:309)
    T habove = h_up, gabove = g_up;
    T h = h_up, g = g_up;
    int hra = hr_up, gra = gr_up, hr = hr_up, gr = gr_up, dr = L.hdiagr;
#pragma unroll
    for (int k = 0; k < R; ++k) {
        const T f = L.E[k];                                      // :147 (computed one column early)
        if (VD) {
            const T ng = habove - uv, eg = gabove - uu;          // :308-309
            const bool sg = ng > eg;                             // :310
            g = sg ? ng : eg; gr = sg ? hra : gra;
            const int fr = L.Er[k];
            const bool sf = f > g;                               // :312 (tie: gg)
            const T x = sf ? f : g; const int xr = sf ? fr : gr;
            const bool sd = sc[k] > x;                           // :313 (tie: the gap state)
            h = sd ? sc[k] : x; hr = sd ? dr : xr;
            const T ng2 = h - vv - uu, eg2 = f - uu;             // :305-307 for the next column
            const bool se = ng2 > eg2;
            L.E[k] = se ? ng2 : eg2; L.Er[k] = se ? hr : fr;
            dr = L.Hr[k]; L.Hr[k] = hr;
            hra = hr; gra = gr;
        } else {
            g = k1f_max(habove - vv, gabove) - uu;               // :148
            h = k1f_max(k1f_max(sc[k], f), g);                   // :150
            if (SWG) {
                h = k1f_max(h, (T)0);                            // :179
                h = (unsigned)(rel - k) <= span ? h : k1f_neg<T>();
                *maxh = k1f_max(*maxh, h);                       // :180
            }
            L.E[k] = k1f_max(h - vv, f) - uu;
        }
        L.H[k] = h;
        habove = h;
        gabove = g;
    }
    L.hdiag = h_up;
    *h_dn = h;
    *g_dn = g;
    if (VD) { L.hdiagr = hr_up; *hr_dn = hr; *gr_dn = gr; }
}

// ---- lastD (fwd2d1.cc:97-134) ------------------------------------------------------------------
// A "line" holds the values the reference's in-place hh[] has on one side of the end corner when
// forwardD returns: line[0] = the boundary value in front of the line, line[i] = the i-th cell
// (last column: H(i-1, LS-1), i = 1..LQ; last row: H(LQ-1, i-1), i = 1..LS); the last entry is the
// end corner H(LQ-1, LS-1), shared by both lines (its current value travels as `corner`).
//   ext    band extent towards this line (up for the last column, -lw for the last row)
//   nline  number of cells on the line, nother = length of the other sequence
//   first_block: the reference's first loop (`while (--h >= h9)`, may start from the boundary value
//   or the hh[up+1] sentinel) or its second loop (`while (++h <= h9)`, starts from an in-band cell).
template <typename T>
PG_HD T k1f_lastd_scan(const T* line, int ext, int nline, int nother, bool first_block, T corner, T vv, T uu, T rtg,
                       const int* rline = nullptr, int* corner_r = nullptr, int* cnt_out = nullptr)
{
    int first;
    T run;
    int runr = 0, cr = corner_r ? *corner_r : 0;
    if (cnt_out) *cnt_out = 0;
    if (first_block) {
        int s = nother - 1 - ext;
        if (s < 0) s = 0;
        run = ext == nother ? K1F_LD(line) : k1f_neg<T>();
        if (rline && ext == nother) runr = K1F_LD(rline);
        first = s + 1;
    } else {
        int e = ext < nother - 1 ? ext : nother - 1;
        int s = nother - 1 - e;
        if (s + 1 >= nline) return corner;
        run = K1F_LD(line + s + 1);
        if (rline) runr = K1F_LD(rline + s + 1);
        first = s + 2;
    }
    if (first > nline) return corner;
    int cnt = 0;
    const T open = vv + uu;
    for (int i = first; i <= nline; ++i) {
        T hv = i == nline ? corner : K1F_LD(line + i);
        int hvr = rline ? (i == nline ? cr : K1F_LD(rline + i)) : 0;
        ++cnt;
        const T gpn = cnt == 1 ? open : uu;
        run = run + k1f_mul(gpn, rtg);                      // two roundings, as the reference
        if (hv < run) { hv = run; hvr = runr; }
        else cnt = 0;
        run = hv; runr = hvr;
    }
    if (corner_r) *corner_r = runr;
    if (cnt_out) *cnt_out = cnt;
    return run;
}

// both blocks in the reference's order; colLine / rowLine as described above.  With the r lines
// (VD) also returns ends[1] = dn ? dn : -dm (fwd2d1.cc:311) through ends1 and the corner's r.
template <typename T>
PG_HD T k1f_lastd(const K1FPair<T>& g, const T* colLine, const T* rowLine, T corner, T vv, T uu,
                  const int* colR = nullptr, const int* rowR = nullptr, int* corner_r = nullptr, int* ends1 = nullptr)
{
    int c1 = 0, c2 = 0;
    if (g.qFirst) {
        if (g.doQ) corner = k1f_lastd_scan(rowLine, -g.lw, g.LS, g.LQ, true, corner, vv, uu, g.rtgQ, rowR, corner_r, &c1);
        if (g.doS) corner = k1f_lastd_scan(colLine, g.up, g.LQ, g.LS, false, corner, vv, uu, g.rtgS, colR, corner_r, &c2);
    } else {
        if (g.doS) corner = k1f_lastd_scan(colLine, g.up, g.LQ, g.LS, true, corner, vv, uu, g.rtgS, colR, corner_r, &c1);
        if (g.doQ) corner = k1f_lastd_scan(rowLine, -g.lw, g.LS, g.LQ, false, corner, vv, uu, g.rtgQ, rowR, corner_r, &c2);
    }
    if (ends1) *ends1 = c2 ? c2 : -c1;
    return corner;
}

// Per-pair setup from the two sequences' flags (bit0 exgl, bit1 exgr, bit2 left != 0, bit3 right != len)
// bnd = T[3][stride]: table 0 zeros (exgl), 1 factor 1, 2 factor tgapf (fwd2d1.cc:67-87).
template <typename T>
PG_HD void k1f_pair_setup(K1FPair<T>& g, int LQ, int LS, int sh, uint8_t qflags, uint8_t sflags, const T* bnd,
                          int stride, float tgapf, bool swap, bool vd = false, int r0 = 0)
{
    g.LQ = LQ; g.LS = LS;
    k1_band(LQ, LS, sh, &g.lw, &g.up);
    // boundary row (columns consumed first) follows the ROW sequence's left end, and vice versa
    int tq = (qflags & 1) ? 0 : (qflags & 4) ? 1 : 2;
    int ts = (sflags & 1) ? 0 : (sflags & 4) ? 1 : 2;
    if (vd) {   // Fwd2d_vd ctor (:227,234): a's exgl is ignored, b's applies only when b.left == 0
        tq = (qflags & 4) ? 1 : 2;
        ts = (sflags & 4) ? 1 : (sflags & 1) ? 0 : 2;
    }
    g.topTab = bnd + (size_t)tq * stride;
    g.leftTab = bnd + (size_t)ts * stride;
    const float rq = (qflags & 2) ? 0.f : tgapf, rs = (sflags & 2) ? 0.f : tgapf;
    g.doQ = !(qflags & 8) && rq < 1.f;
    g.doS = !(sflags & 8) && rs < 1.f;
    g.rtgQ = (T)rq; g.rtgS = (T)rs;
    g.qFirst = swap ? 1 : 0;
    g.originR = -r0;
}

// Boundary tables (host): bnd[t * stride + k] = value of k leading residues against gaps with factor
// t = 0: free (exgl), 1: full penalty, 2: tgapf -- accumulated exactly as the reference does
// (gp = -vv*f; ge = -uu*f; hh[r] = gp += ge, fwd2d1.cc:70-74,80-84).
template <typename T>
static inline void k1f_build_tables(T* bnd, int stride, T uu, T vv, float tgapf)
{
    for (int t = 0; t < 3; ++t) {
        T* tab = bnd + (size_t)t * stride;
        const float f = t == 2 ? tgapf : 1.f;
        T gp = (T)(-vv * f);
        const T ge = (T)(-uu * f);
        tab[0] = 0;
        for (int k = 1; k < stride; ++k) tab[k] = t == 0 ? (T)0 : (gp += ge);
    }
}

// Host-side emulation of one warp (32 lanes x R rows, multi-pass over Q): the control flow of the
// CUDA kernel with shuffles replaced by arrays.  mtx[q * dim + s] = score of (row residue, column residue).
#if !defined(__CUDA_ARCH__)
template <typename T, int R, bool SWG, bool VD>
static inline T k1f_emulate_pair(const uint8_t* q, const uint8_t* s, const K1FPair<T>& g, const T* mtx, int dim,
                                 T vv, T uu, int* ends)
{
    const int NT = 32;
    const int rows_per_pass = NT * R;
    const int LQ = g.LQ, LS = g.LS;
    T* colLine = new T[LQ + 2];
    T* rowLine = new T[LS + 2];
    int* colR = new int[LQ + 2];
    int* rowR = new int[LS + 2];
    T* rowH = new T[LS + 1];
    T* rowG = new T[LS + 1];
    int* rowHr = new int[LS + 1];
    int* rowGr = new int[LS + 1];
    T result = 0, maxh = k1f_neg<T>();
    int result_r = 0, e1 = 0;
    auto topr = [&](int n) { return n + 1 == 0 ? g.originR : n + 1; };
    auto leftr = [&](int m) { return m + 1 == 0 ? g.originR : -(m + 1); };
    if (LQ == 0 || LS == 0) {
        // no cell: hh[] holds the boundary; the lines are the boundary column / row themselves
        for (int i = 0; i <= LQ; ++i) {
            colLine[i] = LS == 0 ? k1f_left(g, i - 1) : (i == 0 ? k1f_top(g, LS - 1) : k1f_neg<T>());
            colR[i] = LS == 0 ? leftr(i - 1) : (i == 0 ? topr(LS - 1) : 0);
        }
        for (int j = 0; j <= LS; ++j) {
            rowLine[j] = LQ == 0 ? k1f_top(g, j - 1) : (j == 0 ? k1f_left(g, LQ - 1) : k1f_neg<T>());
            rowR[j] = LQ == 0 ? topr(j - 1) : (j == 0 ? leftr(LQ - 1) : 0);
        }
        result = LQ == 0 ? k1f_top(g, LS - 1) : k1f_left(g, LQ - 1);
        result_r = LQ == 0 ? topr(LS - 1) : leftr(LQ - 1);
    } else {
        colLine[0] = k1f_top(g, LS - 1); colR[0] = topr(LS - 1);
        rowLine[0] = k1f_left(g, LQ - 1); rowR[0] = leftr(LQ - 1);
    }
    for (int pass = 0; LS > 0 && pass * rows_per_pass < LQ; ++pass) {
        const int pbase = pass * rows_per_pass;
        K1FLane<T, R, VD> L[NT];
        T send_h[2][NT], send_g[2][NT];
        int send_hr[2][NT], send_gr[2][NT];
        for (int t = 0; t < NT; ++t) k1f_lane_init(L[t], g, pbase + t * R, vv, uu);
        const int rows_here = LQ - pbase < rows_per_pass ? LQ - pbase : rows_per_pass;
        const int lanes = (rows_here + R - 1) / R;
        const bool last_pass = pbase + rows_here == LQ;
        for (int step = 0; step < LS + lanes - 1; ++step) {
            const int cur = step & 1, prv = cur ^ 1;
            for (int t = 0; t < lanes; ++t) {
                const int n = step - t;
                if (n < 0 || n >= LS) continue;
                const int mbase = pbase + t * R;
                T h_up, g_up;
                int hr_up = 0, gr_up = 0;
                if (t == 0) {
                    if (pass == 0) { h_up = k1f_top(g, n); g_up = k1f_neg<T>(); hr_up = n + 1; }
                    else { h_up = rowH[n]; g_up = rowG[n]; hr_up = rowHr[n]; gr_up = rowGr[n]; }
                } else { h_up = send_h[prv][t - 1]; g_up = send_g[prv][t - 1]; hr_up = send_hr[prv][t - 1]; gr_up = send_gr[prv][t - 1]; }
                if (!SWG) {
                    const int kL = n - g.lw - mbase, kU = n - g.up - 1 - mbase;
                    if (kL >= 0 && kL < R) L[t].E[kL] = k1f_neg<T>();
                    if (kU >= 0 && kU < R) L[t].E[kU] = k1f_neg<T>();
                }
                T sc[R];
                for (int k = 0; k < R; ++k) {
                    const int m = mbase + k;
                    sc[k] = m < LQ ? mtx[q[m] * dim + s[n]] : (SWG ? k1f_neg<T>() : (T)0);
                }
                T h_dn, g_dn;
                int hr_dn = 0, gr_dn = 0;
                k1f_lane_step<T, R, SWG, VD>(L[t], sc, vv, uu, h_up, g_up, &h_dn, &g_dn, n - mbase - g.lw,
                                             (unsigned)(g.up - g.lw), &maxh, hr_up, gr_up, &hr_dn, &gr_dn);
                send_h[cur][t] = h_dn; send_g[cur][t] = g_dn; send_hr[cur][t] = hr_dn; send_gr[cur][t] = gr_dn;
                if (t == NT - 1) { rowH[n] = h_dn; rowG[n] = g_dn; rowHr[n] = hr_dn; rowGr[n] = gr_dn; }
                if (last_pass && t == (rows_here - 1) / R) {
                    rowLine[n + 1] = L[t].H[(rows_here - 1) % R];
                    if (VD) rowR[n + 1] = L[t].Hr[(rows_here - 1) % R];
                }
            }
        }
        for (int t = 0; t < lanes; ++t)
            for (int k = 0; k < R; ++k)
                if (pbase + t * R + k < LQ) {
                    colLine[pbase + t * R + k + 1] = L[t].H[k];
                    if (VD) colR[pbase + t * R + k + 1] = L[t].Hr[k];
                }
        if (last_pass) {
            result = L[(rows_here - 1) / R].H[(rows_here - 1) % R];
            if (VD) result_r = L[(rows_here - 1) / R].Hr[(rows_here - 1) % R];
        }
    }
    if (SWG) result = maxh == k1f_neg<T>() ? k1f_nevsel((T)0) : maxh;
    else if (VD) result = k1f_lastd(g, colLine, rowLine, result, vv, uu, colR, rowR, &result_r, &e1);
    else result = k1f_lastd(g, colLine, rowLine, result, vv, uu);
    if (VD && ends) { ends[0] = result_r; ends[1] = e1; }
    delete[] colLine; delete[] rowLine; delete[] rowH; delete[] rowG;
    delete[] colR; delete[] rowR; delete[] rowHr; delete[] rowGr;
    return result;
}
#endif
