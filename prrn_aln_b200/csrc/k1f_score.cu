// k1f_score.cu -- kernel K1F: the floating-point, all-modes form of the batched score-only banded
// affine fill (sm_100a).  It takes every alnScoreD call the exact-integer kernels (k1_score.cu,
// k1p_score.cu) do not: non-integral scoring (default PAM matrices), tgapf < 1 / inex.exgl / exgr
// (Fwd2d ctor + lastD, reference src/fwd2d1.cc:57-134), the Smith-Waterman-Gotoh score
// (algmode.lcl & 16, swgforwardD :162-189) and the semi-global score with end points (Fwd2d_vd,
// :191-322), in float (aln build) or double (prrn build) VTYPE, bit-identical to the reference
// because every cell performs the same IEEE add / sub / max in the same order (k1f_core.cuh).
//
// Machine mapping = K1's: CTA = one query (rows, the reference's a) x a run of subjects (columns, b);
// query profile P[letter][row] = mtx[a_row][letter] in shared memory, [letter][chunk][lane][16 B] so
// that a lane's rows are conflict-free LDS.128; warp = one alignment; lane = R consecutive rows in
// registers; systolic wavefront with (H, gg) handed down by shuffle; persistent CTAs + atomic queue.
// Orientation is always the reference's (rows = a, columns = b): Fwd2d_vd is not symmetric.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdlib.h>

#include "k1f_core.cuh"
#include "pg_internal.h"

namespace {

constexpr int NW = 8;
constexpr int MAXDIM = 32;
constexpr unsigned FULL = 0xffffffffu;
enum { M_PLAIN = 0, M_LASTD = 1, M_SWG = 2, M_VD = 3 };

// RO: rows per lane chosen by the host for the plain mode (k1f_rows_per_pass: the smallest stripe of 32 x R rows that
// takes the longest query of the batch in one pass), 0 = the mode's default
template <typename T, int MODE, int RO = 0> struct Cfg {
    static constexpr int R = RO ? RO : (sizeof(T) == 4 ? (MODE == M_VD ? 8 : 16) : 8);
    static constexpr int BPS = (sizeof(T) == 8 && MODE == M_VD) ? 2 : 3;
};

// 64 bytes of profile per lane and letter in every configuration
template <typename T, int R> struct Lay {
    static constexpr int VPL = 16 / sizeof(T);      // values per 16-byte chunk
    static constexpr int CH = R / VPL;              // chunks per lane
    static constexpr int LETTER = CH * 32 * 16;     // bytes per letter
};

__host__ __device__ inline size_t smem_bytes(int dim, int bytes_per_letter)
{
    return (size_t)(dim + NW) * bytes_per_letter + 16;
}

template <typename T> struct Vec16;
template <> struct Vec16<float> { typedef float4 type; };
template <> struct Vec16<double> { typedef double2 type; };

__device__ __forceinline__ float4 pack16(const float* v) { return make_float4(v[0], v[1], v[2], v[3]); }
__device__ __forceinline__ double2 pack16(const double* v) { return make_double2(v[0], v[1]); }
__device__ __forceinline__ void unpack16(float* v, float4 x) { v[0] = x.x; v[1] = x.y; v[2] = x.z; v[3] = x.w; }
__device__ __forceinline__ void unpack16(double* v, double2 x) { v[0] = x.x; v[1] = x.y; }

// The compiler re-derives loop invariants (penalties converted from the argument block, the step count) inside the
// step loop instead of keeping them in registers; an empty asm with the value as in/out operand makes it a plain
// register value it cannot rematerialise.
__device__ __forceinline__ void pin(float& v) { asm volatile("" : "+f"(v)); }
__device__ __forceinline__ void pin(double& v) { asm volatile("" : "+d"(v)); }
__device__ __forceinline__ void pin(int& v) { asm volatile("" : "+r"(v)); }

// E[k] = neg for a run-time k in [0, R): a jump table with one move per case (the lanes that cut the band at a step
// are one or two of 32; the alternative -- the whole array through shared memory -- cost 8 LDS/STS.128 + addressing)
template <typename T, int R>
__device__ __forceinline__ void poke(T (&E)[R], int k, T neg)
{
    switch (k) {
#define K1F_POKE(i) case i: if (i < R) E[i < R ? i : 0] = neg; break;
    K1F_POKE(0) K1F_POKE(1) K1F_POKE(2) K1F_POKE(3) K1F_POKE(4) K1F_POKE(5) K1F_POKE(6) K1F_POKE(7)
    K1F_POKE(8) K1F_POKE(9) K1F_POKE(10) K1F_POKE(11) K1F_POKE(12) K1F_POKE(13) K1F_POKE(14) K1F_POKE(15)
#undef K1F_POKE
    default: break;
    }
}

template <typename T> __device__ __forceinline__ T shfl_up(T v) { return __shfl_up_sync(FULL, v, 1); }
template <typename T> __device__ __forceinline__ T shfl_idx(T v, int src) { return __shfl_sync(FULL, v, src); }

// alnscore2dist tail + dpscore's x100 (aln2.cc:332-333, phyl.cc:249) in FTYPE = T
template <typename T>
__device__ __forceinline__ T dist_value(T score, int dl, T self_a, T self_b, float u_f32)
{
    if (sizeof(T) == 4) {
        float denome = __fsqrt_rn(__fmul_rn((float)self_a, (float)self_b));
        float scr = __fadd_rn((float)score, __fdiv_rn(__fmul_rn(u_f32, (float)dl), 2.f));
        float dst = (float)__dsub_rn(1.0, (double)__fdiv_rn(scr, denome));
        return (T)(float)__dmul_rn(100.0, (double)dst);
    } else {
        double denome = __dsqrt_rn(__dmul_rn((double)self_a, (double)self_b));
        double scr = __dadd_rn((double)score, (double)__fdiv_rn(__fmul_rn(u_f32, (float)dl), 2.f));
        double dst = __dsub_rn(1.0, __ddiv_rn(scr, denome));
        return (T)__dmul_rn(100.0, dst);
    }
}

// selfAlnScr over [lo, hi) (aln2.cc:54-64): sequential VTYPE sum, as the reference adds
template <typename T>
__device__ __forceinline__ T self_window(const uint8_t* p, int lo, int hi, const T* mtx, int dim)
{
    T acc = 0;
    for (int k = lo; k < hi; ++k) {
        const int c = p[k];
        acc = acc + mtx[c * dim + c];
    }
    return acc;
}

template <typename T, int MODE, int RO = 0>
__global__ void __launch_bounds__(NW * 32, Cfg<T, MODE, RO>::BPS) k1f_score_kernel(const K1FArgs a)
{
    constexpr int R = Cfg<T, MODE, RO>::R;
    constexpr int VPL = Lay<T, R>::VPL, CH = Lay<T, R>::CH;
    constexpr int ROWS_PER_PASS = 32 * R;
    constexpr bool SWG = MODE == M_SWG, VD = MODE == M_VD, LINES = MODE == M_LASTD || MODE == M_VD;
    typedef typename Vec16<T>::type V16;

    extern __shared__ __align__(16) unsigned char smem_raw[];
    V16* const sm_prof = reinterpret_cast<V16*>(smem_raw);
    V16* const sm_poke = sm_prof + a.dim * (CH * 32);
    int* const sm_item = reinterpret_cast<int*>(sm_poke + NW * (CH * 32));
    const int tid = threadIdx.x;
    const int lane = tid & 31;
    const int warp = tid >> 5;
    const int gwarp = blockIdx.x * NW + warp;
    const T* const mtx = reinterpret_cast<const T*>(a.mtx);
    const T* const bnd = reinterpret_cast<const T*>(a.bnd);
    T uu = (T)a.uu, vv = (T)a.vv;
    if (sizeof(T) == 4) { pin(uu); pin(vv); }       // double: the argument block already holds them as DADD operands
    const T NEG = k1f_neg<T>();
    // per-warp scratch (multi-pass rows, lastD lines)
    unsigned char* const scratch = a.scratch ? reinterpret_cast<unsigned char*>(a.scratch) + (size_t)gwarp * a.scratch_stride : nullptr;
    T* const rowbufH = reinterpret_cast<T*>(scratch + a.off_rowv);          // [2 * n]: H, gg
    int* const rowbufR = reinterpret_cast<int*>(scratch + a.off_rowr);      // VD: [2 * n]: r of H, gg
    T* const colLine = reinterpret_cast<T*>(scratch + a.off_col);
    T* const rowLine = reinterpret_cast<T*>(scratch + a.off_row);
    int* const colR = reinterpret_cast<int*>(scratch + a.off_colr);
    int* const rowR = reinterpret_cast<int*>(scratch + a.off_rowr2);
    T* const misc = reinterpret_cast<T*>(scratch + a.off_misc);             // SWG: running max across passes

    for (;;) {
        if (tid == 0) *sm_item = atomicAdd(a.counter, 1);
        __syncthreads();
        const int it = *sm_item;
        if (it >= a.nitems) break;
        const PgItem item = a.items[it];
        const int qi = item.q;
        const uint8_t* q = a.seqs.res + a.seqs.offs[qi] + a.seqs.left[qi];
        const int LQ = a.seqs.wlen[qi];
        const uint8_t qflags = a.seqs.flags[qi];
        const int npass = LQ > 0 ? (LQ + ROWS_PER_PASS - 1) / ROWS_PER_PASS : 1;

        for (int pass = 0; pass < npass; ++pass) {
            const int pbase = pass * ROWS_PER_PASS;
            // ---- query profile of this pass: row = pbase + lane*R + chunk*VPL + c
            {
                T* p = reinterpret_cast<T*>(sm_prof);
                const int total = a.dim * 32 * R;
                for (int idx = tid; idx < total; idx += NW * 32) {
                    const int letter = idx / (32 * R);
                    const int rem = idx - letter * (32 * R);
                    const int j = rem / (32 * VPL), ln = (rem / VPL) & 31, c = rem % VPL;
                    const int row = pbase + ln * R + j * VPL + c;
                    p[idx] = row < LQ ? mtx[(int)q[row] * a.dim + letter] : (SWG ? NEG : (T)0);
                }
            }
            __syncthreads();

            const int rows_here = min(LQ - pbase, ROWS_PER_PASS);
            const int lanes = (rows_here + R - 1) / R;
            const int mbase = pbase + lane * R;
            const bool last_pass = pass == npass - 1;
            const int tl = (rows_here - 1) / R, kf = (rows_here - 1) % R;

            for (int sub = item.sub_begin + warp; sub < item.sub_end; sub += NW) {
                int si;
                int64_t slot;
                if (a.pair_s) {
                    si = a.pair_s[sub];
                    slot = a.pair_out[sub];
                } else {            // implicit all-vs-all: a = qi (rows), b = sub > qi, slot = elem(qi, sub)
                    si = sub;
                    slot = (int64_t)si * (si - 1) / 2 + qi;
                    if (slot < a.k_begin || slot >= a.k_end) continue;
                    slot -= a.k_begin;
                }
                const uint8_t* s = a.seqs.res + a.seqs.offs[si] + a.seqs.left[si];
                const int LS = a.seqs.wlen[si];
                uint8_t sflags = a.seqs.flags[si], qfl = qflags;
                if (a.exg_override >= 0) {      // exg_seq per role (alnscore2dist, aln2.cc:298-299)
                    qfl = (uint8_t)((qfl & ~3) | (a.exg_override & 3));
                    sflags = (uint8_t)((sflags & ~3) | ((a.exg_override >> 2) & 3));
                }
                K1FPair<T> g;
                k1f_pair_setup<T>(g, LQ, LS, a.sh, qfl, sflags, bnd, a.bnd_stride, a.tgapf, false, VD,
                                  a.seqs.left[si] - a.seqs.left[qi]);
                T result = 0;
                int result_r = 0;
                T maxh = NEG;
                const bool degenerate = LQ == 0 || LS == 0;

                if (!degenerate) {
                    K1FLane<T, R, VD> L;
                    k1f_lane_init(L, g, mbase, vv, uu);
                    const int lwm = g.lw + mbase;           // kL = n - lwm
                    const int upm = g.up + 1 + mbase;       // kU = n - upm
                    const unsigned span = (unsigned)(g.up - g.lw);
                    T recv_h = NEG, recv_g = NEG;
                    int recv_hr = 0, recv_gr = 0;
                    const V16* pp = sm_prof + lane;
                    V16* pk = sm_poke + warp * (CH * 32) + lane;
                    int nsteps = LS + lanes - 1;
                    if (sizeof(T) == 4) pin(nsteps);
                    if (SWG && pass > 0) maxh = __ldcg(misc);

                    for (int step = 0; step < nsteps; ++step) {
                        const int n = step - lane;
                        T h_dn = NEG, g_dn = NEG;
                        int hr_dn = 0, gr_dn = 0;
                        if (n >= 0 && n < LS && lane < lanes) {
                            T h_up = recv_h, g_up = recv_g;
                            int hr_up = recv_hr, gr_up = recv_gr;
                            if (lane == 0) {
                                if (pass == 0) { h_up = k1f_top(g, n); g_up = NEG; hr_up = n + 1; gr_up = 0; }
                                else {
                                    h_up = __ldcg(rowbufH + 2 * n); g_up = __ldcg(rowbufH + 2 * n + 1);
                                    if (VD) { hr_up = __ldcg(rowbufR + 2 * n); gr_up = __ldcg(rowbufR + 2 * n + 1); }
                                }
                            }
                            if (!SWG) {
                                // band cut: rows on diagonal lw / up+1 lose their horizontal input
                                const int kL = n - lwm, kU = n - upm;
                                if (sizeof(T) == 4) {       // measured: float 1,384 -> 1,500 GCUPS with the jump table,
                                    if ((unsigned)kL < (unsigned)R) poke<T, R>(L.E, kL, NEG);      // double 602 -> 568
                                    if ((unsigned)kU < (unsigned)R) poke<T, R>(L.E, kU, NEG);
                                } else if ((unsigned)kL < (unsigned)R || (unsigned)kU < (unsigned)R) {
                                    // double: the eight 16-byte halves of E through this warp's shared-memory scratch
#pragma unroll
                                    for (int j = 0; j < CH; ++j) pk[j * 32] = pack16(&L.E[j * VPL]);
                                    T* pks = reinterpret_cast<T*>(pk);
                                    if ((unsigned)kL < (unsigned)R) pks[(kL / VPL) * (32 * VPL) + (kL % VPL)] = NEG;
                                    if ((unsigned)kU < (unsigned)R) pks[(kU / VPL) * (32 * VPL) + (kU % VPL)] = NEG;
#pragma unroll
                                    for (int j = 0; j < CH; ++j) unpack16(&L.E[j * VPL], pk[j * 32]);
                                }
                            }
                            const int letter = __ldg(s + n);
                            const V16* pl = pp + letter * (CH * 32);
                            T sc[R];
#pragma unroll
                            for (int j = 0; j < CH; ++j) unpack16(&sc[j * VPL], pl[j * 32]);
                            k1f_lane_step<T, R, SWG, VD>(L, sc, vv, uu, h_up, g_up, &h_dn, &g_dn, n - lwm, span, &maxh,
                                                         hr_up, gr_up, &hr_dn, &gr_dn);
                            if (lane == 31 && !last_pass) {
                                __stcg(rowbufH + 2 * n, h_dn); __stcg(rowbufH + 2 * n + 1, g_dn);
                                if (VD) { __stcg(rowbufR + 2 * n, hr_dn); __stcg(rowbufR + 2 * n + 1, gr_dn); }
                            }
                            if (LINES && last_pass && lane == tl) {     // the last row, for lastD
                                T v = L.H[0];
                                int vr = VD ? L.Hr[0] : 0;
#pragma unroll
                                for (int k = 1; k < R; ++k)
                                    if (k == kf) { v = L.H[k]; if (VD) vr = L.Hr[k]; }
                                __stcg(rowLine + n + 1, v);
                                if (VD) __stcg(rowR + n + 1, vr);
                            }
                        }
                        recv_h = shfl_up(h_dn);
                        recv_g = shfl_up(g_dn);
                        if (VD) { recv_hr = shfl_up(hr_dn); recv_gr = shfl_up(gr_dn); }
                    }

                    if (LINES && lane < lanes) {        // the last column, for lastD
#pragma unroll
                        for (int k = 0; k < R; ++k)
                            if (mbase + k < LQ) {
                                __stcg(colLine + mbase + k + 1, L.H[k]);
                                if (VD) __stcg(colR + mbase + k + 1, L.Hr[k]);
                            }
                    }
                    if (SWG) {
#pragma unroll
                        for (int o = 16; o; o >>= 1) maxh = k1f_max(maxh, __shfl_xor_sync(FULL, maxh, o));
                        if (!last_pass && lane == 0) __stcg(misc, maxh);
                    }
                    if (last_pass) {
                        T val = L.H[0];
                        int vr = VD ? L.Hr[0] : 0;
#pragma unroll
                        for (int k = 1; k < R; ++k)
                            if (k == kf) { val = L.H[k]; if (VD) vr = L.Hr[k]; }
                        result = shfl_idx(val, tl);
                        if (VD) result_r = shfl_idx(vr, tl);
                    }
                }
                if (!last_pass) continue;
                __syncwarp();

                // ---- epilogue (lane 0, or lanes 0/1 for the two self sums): lastD, ends, score / distance
                int e1 = 0;
                if (lane == 0) {
                    if (SWG) result = maxh == NEG ? k1f_nevsel((T)0) : maxh;
                    else if (LINES) {
                        if (degenerate) {
                            // no cell: the lines are the boundary column / row themselves
                            for (int i = 0; i <= LQ; ++i) {
                                __stcg(colLine + i, LS == 0 ? k1f_left(g, i - 1) : (i == 0 ? k1f_top(g, LS - 1) : NEG));
                                if (VD) __stcg(colR + i, LS == 0 ? (i == 0 ? g.originR : -i) : (i == 0 ? LS : 0));
                            }
                            for (int j = 0; j <= LS; ++j) {
                                __stcg(rowLine + j, LQ == 0 ? k1f_top(g, j - 1) : (j == 0 ? k1f_left(g, LQ - 1) : NEG));
                                if (VD) __stcg(rowR + j, LQ == 0 ? (j == 0 ? g.originR : j) : (j == 0 ? -LQ : 0));
                            }
                            result = LQ == 0 ? k1f_top(g, LS - 1) : k1f_left(g, LQ - 1);
                            if (VD) result_r = LQ == 0 ? (LS == 0 ? g.originR : LS) : -LQ;
                        } else {
                            __stcg(colLine, k1f_top(g, LS - 1));
                            __stcg(rowLine, k1f_left(g, LQ - 1));
                            if (VD) { __stcg(colR, LS); __stcg(rowR, -LQ); }
                        }
                        if (VD) result = k1f_lastd(g, colLine, rowLine, result, vv, uu, colR, rowR, &result_r, &e1);
                        else result = k1f_lastd(g, colLine, rowLine, result, vv, uu);
                    } else if (degenerate) {
                        result = LQ == 0 ? k1f_top(g, LS - 1) : k1f_left(g, LQ - 1);
                    }
                }
                if (a.epilogue == 0) {
                    if (lane == 0) {
                        reinterpret_cast<T*>(a.out)[slot] = result;
                        if (VD && a.out_ends) { a.out_ends[2 * slot] = result_r; a.out_ends[2 * slot + 1] = e1; }
                    }
                } else if (a.epilogue == 1) {           // dpscore / alnscore2dist, global branch
                    if (lane == 0) {
                        const int dl = LQ > LS ? LQ - LS : LS - LQ;
                        const T* self = reinterpret_cast<const T*>(a.self);
                        reinterpret_cast<T*>(a.out)[slot] = dist_value<T>(result, dl, self[qi], self[si], a.u_f32);
                    }
                } else {                                // alnscore2dist, algmode.lcl branch (aln2.cc:296-320)
                    result = shfl_idx(result, 0);
                    const int e0 = shfl_idx(result_r, 0);
                    e1 = shfl_idx(e1, 0);
                    int al = 0, ar = LQ, bl = 0, br = LS;       // window-relative
                    if (e0 > 0) bl += e0; else if (e0 < 0) al -= e0;
                    if (e1 > 0) br -= e1; else if (e1 < 0) ar += e1;
                    T sv = 0;
                    if (lane == 0) sv = self_window<T>(q, al, ar, mtx, a.dim);
                    if (lane == 1) sv = self_window<T>(s, bl, br, mtx, a.dim);
                    const T sa = shfl_idx(sv, 0), sb = shfl_idx(sv, 1);
                    if (lane == 0) {
                        int dl = ar - al - br + bl;
                        dl = dl < 0 ? -dl : dl;
                        reinterpret_cast<T*>(a.out)[slot] = dist_value<T>(result, dl, sa, sb, a.u_f32);
                        if (a.out_ends) { a.out_ends[2 * slot] = e0; a.out_ends[2 * slot + 1] = e1; }
                    }
                }
                __syncwarp();
            }
            __syncthreads();    // profile (and scratch lines) are reused by the next pass / item
        }
    }
}

// self score per sequence (selfAlnScr, aln2.cc:54-64): one thread per sequence, sequential VTYPE sum
template <typename T>
__global__ void k1f_self_kernel(PgDevSeqs s, const T* mtx, int dim, T* self)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= s.nseq) return;
    self[i] = self_window<T>(s.res + s.offs[i] + s.left[i], 0, s.wlen[i], mtx, dim);
}

template <typename T, int MODE, int RO = 0>
cudaError_t launch_one(const K1FArgs& a, int sm_count, cudaStream_t st)
{
    constexpr int R = Cfg<T, MODE, RO>::R;
    const size_t smem = smem_bytes(a.dim, Lay<T, R>::LETTER);
    cudaError_t e = cudaFuncSetAttribute(k1f_score_kernel<T, MODE, RO>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)smem_bytes(MAXDIM, Lay<T, R>::LETTER));
    if (e != cudaSuccess) return e;
    k1f_score_kernel<T, MODE, RO><<<sm_count * Cfg<T, MODE, RO>::BPS, NW * 32, smem, st>>>(a);
    return cudaGetLastError();
}

}  // namespace

int k1f_warps_per_block() { return NW; }
// Rows of one pass = 32 lanes x R rows.  Plain mode: R follows the window lengths of the batch, so that short sets do not
// leave most lanes of the warp idle (R = 16 on 300-row queries: 19 of 32 lanes) and queries just over 256 rows in double
// do not need a second pass of a few lanes.  Cost of a query = passes x (R + the step's fixed instructions counted in
// rows); every sequence counts as a query once (calcdist: each is the row side of about half its pairs).
int k1f_rows_per_pass(int vtype, int mode, const int32_t* wlen, int nseq)
{
    const int dflt = vtype ? 8 : (mode == M_VD ? 8 : 16);
    static const bool fixed = getenv("PG_K1F_FIXED_ROWS") != nullptr;      // A/B switch: the mode's default R always
    if (mode != M_PLAIN || fixed || !wlen || nseq <= 0) return 32 * dflt;
    static const int cand_f[4] = {4, 8, 12, 16}, cand_d[4] = {4, 6, 8, 10};
    const int* cand = vtype ? cand_d : cand_f;
    const int nc = 4, fixed_rows = vtype ? 3 : 12;
    int best = dflt;
    double best_cost = -1;
    for (int c = 0; c < nc; ++c) {
        const int rpp = 32 * cand[c];
        double cost = 0;
        for (int i = 0; i < nseq; ++i) cost += (double)((wlen[i] + rpp - 1) / rpp) * (cand[c] + fixed_rows);
        if (best_cost < 0 || cost < best_cost || (cost == best_cost && cand[c] == dflt)) { best_cost = cost; best = cand[c]; }
    }
    return 32 * best;
}
int k1f_grid_blocks(int sm_count, int vtype, int mode) { return sm_count * ((vtype && mode == M_VD) ? 2 : 3); }

cudaError_t k1f_launch(const K1FArgs& a, int sm_count, cudaStream_t st)
{
    if (a.dim < 1 || a.dim > MAXDIM) return cudaErrorInvalidValue;
    if (a.mode == M_PLAIN) {
        switch (a.vtype ? -a.rows_per_lane : a.rows_per_lane) {
        case 4: return launch_one<float, M_PLAIN, 4>(a, sm_count, st);
        case 8: return launch_one<float, M_PLAIN, 8>(a, sm_count, st);
        case 12: return launch_one<float, M_PLAIN, 12>(a, sm_count, st);
        case -4: return launch_one<double, M_PLAIN, 4>(a, sm_count, st);
        case -6: return launch_one<double, M_PLAIN, 6>(a, sm_count, st);
        case -10: return launch_one<double, M_PLAIN, 10>(a, sm_count, st);
        default: break;
        }
    }
    if (a.vtype) {
        switch (a.mode) {
        case M_PLAIN: return launch_one<double, M_PLAIN>(a, sm_count, st);
        case M_LASTD: return launch_one<double, M_LASTD>(a, sm_count, st);
        case M_SWG: return launch_one<double, M_SWG>(a, sm_count, st);
        default: return launch_one<double, M_VD>(a, sm_count, st);
        }
    }
    switch (a.mode) {
    case M_PLAIN: return launch_one<float, M_PLAIN>(a, sm_count, st);
    case M_LASTD: return launch_one<float, M_LASTD>(a, sm_count, st);
    case M_SWG: return launch_one<float, M_SWG>(a, sm_count, st);
    default: return launch_one<float, M_VD>(a, sm_count, st);
    }
}

cudaError_t k1f_self_launch(const PgDevSeqs& s, const void* mtx, int dim, int vtype, void* self, cudaStream_t st)
{
    const int blocks = (s.nseq + 127) / 128;
    if (blocks < 1) return cudaSuccess;
    if (vtype) k1f_self_kernel<double><<<blocks, 128, 0, st>>>(s, (const double*)mtx, dim, (double*)self);
    else k1f_self_kernel<float><<<blocks, 128, 0, st>>>(s, (const float*)mtx, dim, (float*)self);
    return cudaGetLastError();
}
