// k1_score.cu -- kernel K1: batched score-only banded affine Gotoh fill (sm_100a).
//
// Stands behind alnScoreD / Fwd2d::forwardD (reference src/fwd2d1.cc:57-160, 324-337) and the
// dpscore/alnscore2dist epilogue (src/phyl.cc:221-251, src/aln2.cc:321-333).
//
// Mapping onto the machine
//   * one CTA  = one work item = one query (rows) x a run of subjects (columns).  The CTA builds
//     the query profile  P[letter][row] = S(q_row, letter) + 2u  once in shared memory, laid out
//     [letter][R/4][lane][4] so that each lane's 4 consecutive rows are one conflict-free LDS.128.
//   * one warp = one alignment at a time.  Lane t owns rows [t*R, (t+1)*R) of the pass in
//     registers (H, E per row) and walks the columns one step behind lane t-1 (systolic
//     wavefront); the bottom row's (H, F) go to the lane below by warp shuffle.
//   * the recurrence runs on DPX instructions: 3 x __viaddmax_s32 + 1 x max per cell (k1_core.cuh).
//   * queries longer than 32*R rows take several passes; the last row of a pass is kept in a
//     per-warp global scratch line (L2-resident) and re-enters as the top boundary of the next.
//   * persistent CTAs pull work items from an atomic counter; grid = SMs x resident CTAs.
#include <cuda_runtime.h>
#include <stdint.h>

#include <stdlib.h>

#include "k1_core.cuh"
#include "pg_internal.h"

namespace {

constexpr int NW = 8;           // warps per CTA
constexpr int MAXDIM = 32;      // profile letters (dim <= 32)
constexpr int BLOCKS_PER_SM = 3;
constexpr unsigned FULL = 0xffffffffu;

// dynamic shared memory: prof[dim][R/4][32] int4 (2 KB per letter at R = 16) | item
__host__ __device__ inline size_t smem_bytes(int dim, int R)
{
    return (size_t)dim * (R / 4) * 32 * sizeof(int4) + 16;
}

__device__ __forceinline__ void epilogue_store(const K1Args& a, int64_t slot, int score, int qi, int si,
                                               int LQ, int LS)
{
    switch (a.epilogue) {
    case PG_EPI_SCORE_F32:
        reinterpret_cast<float*>(a.out)[slot] = (float)score;
        break;
    case PG_EPI_SCORE_F64:
        reinterpret_cast<double*>(a.out)[slot] = (double)score;
        break;
    case PG_EPI_DIST_F32: {
        // phyl.cc:230  denome = sqrt(scr1[a] * scr1[b])            (FTYPE = float)
        // aln2.cc:332  scr += alprm.u * abs(dlen) / 2              (float)
        // aln2.cc:333  return 1. - scr / denome                    (double expr -> float)
        // phyl.cc:249  dist = 100. * dst                           (double expr -> float)
        float denome = __fsqrt_rn(__fmul_rn((float)a.self[qi], (float)a.self[si]));
        int dl = LQ > LS ? LQ - LS : LS - LQ;
        float scr = __fadd_rn((float)score, __fdiv_rn(__fmul_rn(a.u_f32, (float)dl), 2.f));
        float dst = (float)__dsub_rn(1.0, (double)__fdiv_rn(scr, denome));
        reinterpret_cast<float*>(a.out)[slot] = (float)__dmul_rn(100.0, (double)dst);
        break;
    }
    case PG_EPI_DIST_F64: {
        double denome = __dsqrt_rn(__dmul_rn((double)a.self[qi], (double)a.self[si]));
        int dl = LQ > LS ? LQ - LS : LS - LQ;
        double scr = __dadd_rn((double)score, (double)__fdiv_rn(__fmul_rn(a.u_f32, (float)dl), 2.f));
        double dst = __dsub_rn(1.0, __ddiv_rn(scr, denome));
        reinterpret_cast<double*>(a.out)[slot] = __dmul_rn(100.0, dst);
        break;
    }
    }
}

// gap factor of a boundary: 0 if that end is free (exgl) or tgapf == 0 at a true sequence start
__device__ __forceinline__ int lead_factor(const K1Args& a, uint8_t flags)
{
    if (flags & 1) return 0;                    // inex.exgl
    if (flags & 4) return 1;                    // left != 0 -> factor 1 (fwd2d1.cc:71,80)
    return a.tgapf_zero ? 0 : 1;
}

// E[k] = -inf for a run-time k in [0, RR)
template <int RR>
__device__ __forceinline__ void poke_row(int (&E)[RR], int k)
{
    switch (k) {
#define K1_POKE(i) case i: if (i < RR) E[i < RR ? i : 0] = K1_NEG; break;
    K1_POKE(0) K1_POKE(1) K1_POKE(2) K1_POKE(3) K1_POKE(4) K1_POKE(5) K1_POKE(6) K1_POKE(7)
    K1_POKE(8) K1_POKE(9) K1_POKE(10) K1_POKE(11) K1_POKE(12) K1_POKE(13) K1_POKE(14) K1_POKE(15)
#undef K1_POKE
    default: break;
    }
}

// R = rows per lane (4 / 8 / 12 / 16, chosen per batch by k1_rows_per_pass: the smallest stripe of 32 x R rows that takes
// the batch's queries in one pass leaves the fewest lanes idle)
template <int R>
__global__ void __launch_bounds__(NW * 32, BLOCKS_PER_SM) k1_score_kernel(const K1Args a)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    int4* const sm_prof = reinterpret_cast<int4*>(smem_raw);
    constexpr int ROWS_PER_PASS = 32 * R;
    int* const sm_item = reinterpret_cast<int*>(sm_prof + a.dim * (R / 4) * 32);
    const int tid = threadIdx.x;
    const int lane = tid & 31;
    const int warp = tid >> 5;
    const int gwarp = blockIdx.x * NW + warp;
    const int negv = -a.v;

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
            // ---- query profile of this pass: P[letter][j][lane][c], row = pbase + lane*R + 4j + c
            {
                int* p = reinterpret_cast<int*>(sm_prof);
                const int total = a.dim * 32 * R;
                for (int idx = tid; idx < total; idx += NW * 32) {
                    int letter = idx / (32 * R);
                    int rem = idx - letter * (32 * R);
                    int j = rem >> 7, ln = (rem >> 2) & 31, c = rem & 3;
                    int row = pbase + ln * R + j * 4 + c;
                    p[idx] = row < LQ ? a.mtx[(int)q[row] * a.dim + letter] + 2 * a.u : 0;
                }
            }
            __syncthreads();

            const int rows_here = min(LQ - pbase, ROWS_PER_PASS);
            const int lanes = (rows_here + R - 1) / R;
            const int mbase = pbase + lane * R;
            const bool last_pass = pass == npass - 1;

            for (int sub = item.sub_begin + warp, ord = 0; sub < item.sub_end; sub += NW, ++ord) {
                int si;
                int64_t slot;
                if (a.pair_s) {
                    si = a.pair_s[sub];
                    slot = a.pair_out[sub];
                } else {
                    si = sub;
                    slot = (int64_t)qi * (qi - 1) / 2 + si;
                    if (slot < a.k_begin || slot >= a.k_end) continue;
                    slot -= a.k_begin;
                }
                const uint8_t* s = a.seqs.res + a.seqs.offs[si] + a.seqs.left[si];
                const int LS = a.seqs.wlen[si];
                K1Geom g;
                g.LQ = LQ; g.LS = LS; g.u = a.u; g.v = a.v;
                k1_band(LQ, LS, a.sh, &g.lw, &g.up);
                {
                    int fq = lead_factor(a, qflags), fs = lead_factor(a, a.seqs.flags[si]);
                    g.topOpen = -a.v * fq; g.topExt = -a.u * fq;
                    g.leftOpen = -a.v * fs; g.leftExt = -a.u * fs;
                }
                if (LQ == 0 || LS == 0) {       // no cell: the score is a boundary value
                    if (lane == 0) {
                        int val = LQ == 0 ? k1_top(g, LS - 1) : k1_left(g, LQ - 1);
                        epilogue_store(a, slot, val - (LQ + LS) * a.u, qi, si, LQ, LS);
                    }
                    continue;
                }
                int2* rowbuf = a.rowbuf ? a.rowbuf + ((int64_t)gwarp * a.rowbuf_stride + (int64_t)ord * LS) : nullptr;

                K1Lane<R> L;
                k1_lane_init(L, g, mbase);
                const int lwm = g.lw + mbase;           // kL = n - lwm
                const int upm = g.up + 1 + mbase;       // kU = n - upm
                int recv_h = K1_NEG, recv_f = K1_NEG;
                const int4* pp = sm_prof + lane;
                const int nsteps = LS + lanes - 1;

                for (int step = 0; step < nsteps; ++step) {
                    const int n = step - lane;
                    int h_dn = K1_NEG, f_dn = K1_NEG;
                    if (n >= 0 && n < LS && lane < lanes) {
                        int h_up = recv_h, f_up = recv_f;
                        if (lane == 0) {
                            if (pass == 0) { h_up = k1_top(g, n); f_up = K1_ADDMAX(h_up, negv, K1_NEG); }  // gg may open from the boundary row (fwd2d1.cc:148)
                            else { int2 v = __ldcg(rowbuf + n); h_up = v.x; f_up = v.y; }
                        }
                        // band poke: rows on diagonal lw / up+1 lose their horizontal input
                        const int kL = n - lwm, kU = n - upm;
                        // (a jump table with one move per case: the lanes that cut at a step are one or two of 32, and the
                        //  round trip of the whole E array through shared memory cost 8 LDS / STS.128 in most steps)
                        if ((unsigned)kL < (unsigned)R) poke_row<R>(L.E, kL);
                        if ((unsigned)kU < (unsigned)R) poke_row<R>(L.E, kU);
                        const int letter = __ldg(s + n);
                        const int4* pl = pp + letter * ((R / 4) * 32);
                        int sc[R];
#pragma unroll
                        for (int j = 0; j < R / 4; ++j) {
                            int4 v = pl[j * 32];
                            sc[4 * j] = v.x; sc[4 * j + 1] = v.y; sc[4 * j + 2] = v.z; sc[4 * j + 3] = v.w;
                        }
                        k1_lane_step(L, sc, negv, h_up, f_up, &h_dn, &f_dn);
                        if (lane == 31 && !last_pass) __stcg(rowbuf + n, make_int2(h_dn, f_dn));
                    }
                    recv_h = __shfl_up_sync(FULL, h_dn, 1);
                    recv_f = __shfl_up_sync(FULL, f_dn, 1);
                }

                if (last_pass) {
                    const int tl = (rows_here - 1) / R, kf = (rows_here - 1) % R;
                    int val = 0;
#pragma unroll
                    for (int k = 0; k < R; ++k)
                        if (k == kf) val = L.H[k];
                    val = __shfl_sync(FULL, val, tl);
                    if (lane == 0) epilogue_store(a, slot, val - (LQ + LS) * a.u, qi, si, LQ, LS);
                }
            }
            __syncthreads();    // profile (and rowbuf lines) are reused by the next pass / item
        }
    }
}

// self score per sequence: sum of S(x, x) over the window (selfAlnScr, aln2.cc:54-64)
__global__ void k1_self_kernel(PgDevSeqs s, const int32_t* mtx, int dim, int32_t* self)
{
    const int warps = (gridDim.x * blockDim.x) >> 5;
    const int lane = threadIdx.x & 31;
    for (int i = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; i < s.nseq; i += warps) {
        const uint8_t* p = s.res + s.offs[i] + s.left[i];
        int acc = 0;
        for (int k = lane; k < s.wlen[i]; k += 32) {
            int c = p[k];
            acc += mtx[c * dim + c];
        }
        for (int o = 16; o; o >>= 1) acc += __shfl_xor_sync(FULL, acc, o);
        if (lane == 0) self[i] = acc;
    }
}

}  // namespace

// cost of a query = passes x (R + the step's fixed instructions counted in rows: a DPX cell is ~5 instructions, the
// step's bookkeeping ~80)
int k1_rows_per_pass(const int32_t* wlen, int nseq)
{
    static const bool fixed = getenv("PG_K1_FIXED_ROWS") != nullptr;       // A/B switch
    if (fixed || !wlen || nseq <= 0) return 32 * 16;
    static const int cand[4] = {4, 8, 12, 16};
    int best = 16;
    double best_cost = -1;
    for (int c = 0; c < 4; ++c) {
        const int rpp = 32 * cand[c];
        double cost = 0;
        for (int i = 0; i < nseq; ++i) cost += (double)((wlen[i] + rpp - 1) / rpp) * (cand[c] + 16);
        if (best_cost < 0 || cost < best_cost || (cost == best_cost && cand[c] == 16)) { best_cost = cost; best = cand[c]; }
    }
    return 32 * best;
}
int k1_warps_per_block() { return NW; }
int k1_blocks_per_sm() { return BLOCKS_PER_SM; }

template <int R>
static cudaError_t k1_launch_r(const K1Args& a, int grid_blocks, cudaStream_t st)
{
    const size_t smem = smem_bytes(a.dim, R);
    cudaError_t e = cudaFuncSetAttribute(k1_score_kernel<R>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)smem_bytes(MAXDIM, R));
    if (e != cudaSuccess) return e;
    k1_score_kernel<R><<<grid_blocks, NW * 32, smem, st>>>(a);
    return cudaGetLastError();
}

cudaError_t k1_launch(const K1Args& a, int grid_blocks, cudaStream_t st)
{
    if (a.dim < 1 || a.dim > MAXDIM) return cudaErrorInvalidValue;
    switch (a.rows_per_lane) {
    case 4: return k1_launch_r<4>(a, grid_blocks, st);
    case 8: return k1_launch_r<8>(a, grid_blocks, st);
    case 12: return k1_launch_r<12>(a, grid_blocks, st);
    default: return k1_launch_r<16>(a, grid_blocks, st);
    }
}

cudaError_t k1_self_launch(const PgDevSeqs& s, const int32_t* mtx, int dim, int32_t* self, cudaStream_t st)
{
    int blocks = (s.nseq + 7) / 8;
    if (blocks > 1184) blocks = 1184;
    if (blocks < 1) blocks = 1;
    k1_self_kernel<<<blocks, 256, 0, st>>>(s, mtx, dim, self);
    return cudaGetLastError();
}
