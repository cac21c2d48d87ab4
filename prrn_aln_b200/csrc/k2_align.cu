// k2_align.cu -- kernel K2: pairwise banded affine alignment WITH path for two single sequences.
//
// Stands behind alignC<DPunit> (reference src/fwd2c.h:670-677: Fwd2c ctor :81, initB :138,
// forwardB :359, Vmf::traceback src/vmf.cc:103): global NGP mode, thickness 1 (no nil ends),
// integer scoring.  Two kernels:
//   k2_fill_kernel   the K1 machine mapping (CTA = query profile, warp = alignment, lane = 16 rows,
//                    systolic wavefront) + 4 direction bits per cell, one coalesced 256-byte store
//                    per warp-step in wavefront order (k2_core.cuh);
//   k2_trace_kernel  one thread per alignment: walks the bits back, replays the path forward with
//                    the reference's record rules and emits the corner list in Vmf back-walk order.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#include "k2_core.cuh"
#include "pg_internal.h"

namespace {

constexpr int R = 16;
constexpr int NW = 8;
constexpr int ROWS_PER_PASS = 32 * R;
constexpr int MAXDIM = 32;
constexpr int BLOCKS_PER_SM = 2;
constexpr unsigned FULL = 0xffffffffu;

__host__ __device__ inline size_t smem_bytes(int dim)
{
    return (size_t)(dim + NW) * (R / 4) * 32 * sizeof(int4) + 16;
}

__global__ void __launch_bounds__(NW * 32, BLOCKS_PER_SM) k2_fill_kernel(const K2Args a)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    int4* const sm_prof = reinterpret_cast<int4*>(smem_raw);
    int4* const sm_poke = sm_prof + a.dim * (R / 4) * 32;
    int* const sm_item = reinterpret_cast<int*>(sm_poke + NW * (R / 4) * 32);
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
        const int npass = LQ > 0 ? (LQ + ROWS_PER_PASS - 1) / ROWS_PER_PASS : 1;

        for (int pass = 0; pass < npass; ++pass) {
            const int pbase = pass * ROWS_PER_PASS;
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
                const int si = a.pair_s[sub];
                const uint8_t* s = a.seqs.res + a.seqs.offs[si] + a.seqs.left[si];
                const int LS = a.seqs.wlen[si];
                K1Geom g;
                g.LQ = LQ; g.LS = LS; g.u = a.u; g.v = a.v;
                k1_band(LQ, LS, a.sh, &g.lw, &g.up);
                g.topOpen = -a.v; g.topExt = -a.u; g.leftOpen = -a.v; g.leftExt = -a.u;
                if (LQ == 0 || LS == 0) {
                    if (lane == 0 && pass == 0) a.score[sub] = 0;       // align2: nogap_skl, score untouched
                    continue;
                }
                int2* rowbuf = a.rowbuf ? a.rowbuf + ((int64_t)gwarp * a.rowbuf_stride + (int64_t)ord * LS) : nullptr;
                unsigned long long* words = a.dirs + a.dir_off[sub] + ((int64_t)pass * (LS + 31)) * 32 + lane;

                K2Lane<R> L;
                k2_lane_init(L, g, mbase);
                const int lwm = g.lw + mbase;
                const int upm = g.up + 1 + mbase;
                int recv_h = K1_NEG, recv_g = K1_NEG;
                const int4* pp = sm_prof + lane;
                int4* pk = sm_poke + warp * ((R / 4) * 32) + lane;
                const int nsteps = LS + lanes - 1;

                for (int step = 0; step < nsteps; ++step) {
                    const int n = step - lane;
                    int h_dn = K1_NEG, g_dn = K1_NEG;
                    if (n >= 0 && n < LS && lane < lanes) {
                        int h_up = recv_h, g_up = recv_g;
                        if (lane == 0) {
                            if (pass == 0) { h_up = k1_top(g, n); g_up = K1_NEG; }
                            else { int2 v = __ldcg(rowbuf + n); h_up = v.x; g_up = v.y; }
                        }
                        const int kL = n - lwm, kU = n - upm;
                        if ((unsigned)kL < (unsigned)R || (unsigned)kU < (unsigned)R) {
#pragma unroll
                            for (int j = 0; j < R / 4; ++j)
                                pk[j * 32] = make_int4(L.E[4 * j], L.E[4 * j + 1], L.E[4 * j + 2], L.E[4 * j + 3]);
                            int* pki = reinterpret_cast<int*>(pk);
                            if ((unsigned)kL < (unsigned)R) pki[(kL >> 2) * 128 + (kL & 3)] = K1_NEG;
                            if ((unsigned)kU < (unsigned)R) pki[(kU >> 2) * 128 + (kU & 3)] = K1_NEG;
#pragma unroll
                            for (int j = 0; j < R / 4; ++j) {
                                int4 v = pk[j * 32];
                                L.E[4 * j] = v.x; L.E[4 * j + 1] = v.y; L.E[4 * j + 2] = v.z; L.E[4 * j + 3] = v.w;
                            }
                        }
                        const int letter = __ldg(s + n);
                        const int4* pl = pp + letter * ((R / 4) * 32);
                        int sc[R];
#pragma unroll
                        for (int j = 0; j < R / 4; ++j) {
                            int4 v = pl[j * 32];
                            sc[4 * j] = v.x; sc[4 * j + 1] = v.y; sc[4 * j + 2] = v.z; sc[4 * j + 3] = v.w;
                        }
                        const unsigned long long bits =
                            k2_lane_step(L, sc, negv, h_up, g_up, mbase == 0, &h_dn, &g_dn);
                        __stcs(words + (int64_t)step * 32, bits);       // streaming: written once, read by the trace
                        if (lane == 31 && !last_pass) __stcg(rowbuf + n, make_int2(h_dn, g_dn));
                    }
                    recv_h = __shfl_up_sync(FULL, h_dn, 1);
                    recv_g = __shfl_up_sync(FULL, g_dn, 1);
                }

                if (last_pass) {
                    const int tl = (rows_here - 1) / R, kf = (rows_here - 1) % R;
                    int val = 0;
#pragma unroll
                    for (int k = 0; k < R; ++k)
                        if (k == kf) val = L.H[k];
                    val = __shfl_sync(FULL, val, tl);
                    if (lane == 0) a.score[sub] = val - (LQ + LS) * a.u;
                }
            }
            __syncthreads();
        }
    }
}

// ---- one LONG pair: striped wavefront across warps ------------------------------------------------
// A query of LQ rows is npass = ceil(LQ / 512) stripes.  k2_fill_kernel runs the stripes of a pair one
// after the other in one warp (fine for proteins, 59 x 30,000 dependent steps for a 30 kb pair).  Here
// every stripe is its own single-warp CTA with its own profile in shared memory, and stripe p+1 follows
// stripe p a few columns behind: the bottom row of p goes through an L2-resident row buffer and a
// progress counter (published every PUB columns with a fence) tells the stripe below how far it may
// read.  Stripes take their index from a ticket, so a waiting stripe only ever waits for one that has
// already started: no deadlock however many stripes are resident.  Same direction-word layout as
// k2_fill_kernel, so the traceback kernel is shared.
constexpr int PUB = 16;

// progress counters between stripes: release store / acquire load at GPU scope (a __threadfence() compiles to
// MEMBAR.SC.GPU and, on the reading side, an L1 invalidation that also throws away the subject's residues)
__device__ __forceinline__ void st_release(int* p, int v)
{
    asm volatile("st.release.gpu.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ int ld_acquire(const int* p)
{
    int v;
    asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}

// Hand-over slots of the second form (k2_fill_long2_kernel): the row buffer is filled with 0x80 bytes before the launch;
// a real H value never has that pattern (drifted scores stay above K1_NEG - (LQ + LS) * (u + v) > 0x80808080 as int).
constexpr unsigned K2_SLOT_EMPTY = 0x80808080u;
// the step loop of the second form is unrolled: the loop-carried values (what the shuffles hand on, the running pointers)
// otherwise change registers through ~14 moves per step (30 kb pair: 7.99 ms rolled, 7.50 ms unrolled twice)
#ifndef K2_LONG_UNROLL
#define K2_LONG_UNROLL 2
#endif
constexpr int K2_LONG_UNROLL_N = K2_LONG_UNROLL;
__device__ __forceinline__ unsigned long long ld_relaxed64(const int2* p)
{
    unsigned long long v;
    asm volatile("ld.relaxed.gpu.global.b64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}

// RR rows per lane: a stripe is 32 * RR rows.  One warp alone on its scheduler runs a lane-step as one dependent
// chain (about 30 cycles per row + 150 of fixed work), so narrow stripes (RR = 4: 128 rows, 235 warps for a 30 kb
// query) finish the matrix in LS + stripes * skew steps of a quarter of the length.  The bottom row of the stripe
// above is read in coalesced chunks of 32 columns (one 256-byte load per 32 steps, handed to lane 0 by shuffle)
// instead of one L2 round trip on the critical path of every step.
template <int RR>
__global__ void __launch_bounds__(32) k2_fill_long_kernel(const K2Args a, int npass)
{
    constexpr int RPP = 32 * RR;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    int4* const sm_prof = reinterpret_cast<int4*>(smem_raw);
    int4* const sm_poke = sm_prof + a.dim * (RR / 4) * 32;
    const int lane = threadIdx.x;
    const int negv = -a.v;
    const int qi = a.pair_q[0], si = a.pair_s[0];
    const uint8_t* q = a.seqs.res + a.seqs.offs[qi] + a.seqs.left[qi];
    const uint8_t* s = a.seqs.res + a.seqs.offs[si] + a.seqs.left[si];
    const int LQ = a.seqs.wlen[qi], LS = a.seqs.wlen[si];
    K1Geom g;
    g.LQ = LQ; g.LS = LS; g.u = a.u; g.v = a.v;
    k1_band(LQ, LS, a.sh, &g.lw, &g.up);
    g.topOpen = -a.v; g.topExt = -a.u; g.leftOpen = -a.v; g.leftExt = -a.u;
    unsigned char* const dir_base = reinterpret_cast<unsigned char*>(a.dirs + a.dir_off[0]);

    for (;;) {
        int pass = 0;
        if (lane == 0) pass = atomicAdd(a.ticket, 1);
        pass = __shfl_sync(FULL, pass, 0);
        if (pass >= npass) break;
        const int pbase = pass * RPP;
        {
            int* p = reinterpret_cast<int*>(sm_prof);
            const int total = a.dim * 32 * RR;
            for (int idx = lane; idx < total; idx += 32) {
                int letter = idx / (32 * RR);
                int rem = idx - letter * (32 * RR);
                int j = rem >> 7, ln = (rem >> 2) & 31, c = rem & 3;
                int row = pbase + ln * RR + j * 4 + c;
                p[idx] = row < LQ ? a.mtx[(int)q[row] * a.dim + letter] + 2 * a.u : 0;
            }
        }
        __syncwarp();
        const int rows_here = min(LQ - pbase, RPP);
        const int lanes = (rows_here + RR - 1) / RR;
        const int mbase = pbase + lane * RR;
        const bool last_pass = pass == npass - 1;
        const int2* const row_in = a.rowbuf + (int64_t)(pass - 1) * LS;      // bottom row of the stripe above
        int2* const row_out = a.rowbuf + (int64_t)pass * LS;
        const int* const prog_in = a.progress + (pass - 1);
        // direction bits: RR / 2 bytes per lane-step, slot (pass * (LS + 31) + step) * 32 + lane (k2_core.cuh)
        unsigned char* const words = dir_base + ((int64_t)pass * (LS + 31) * 32 + lane) * (RR / 2);

#ifdef K2_TIMERS
        unsigned long long tm0; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(tm0));
        long long tm_wait = 0;
#endif
        K2Lane<RR> L;
        k2_lane_init(L, g, mbase);
        const int lwm = g.lw + mbase;
        const int upm = g.up + 1 + mbase;
        int recv_h = K1_NEG, recv_g = K1_NEG;
        const int4* pp = sm_prof + lane;
        int4* pk = sm_poke + lane;
        const int nsteps = LS + lanes - 1;
        int avail = 0;          // columns of the stripe above known to be published
        int2 chunk = make_int2(K1_NEG, K1_NEG);         // lane i: row_in[32 * (step / 32) + i]
        // the residue of the column a lane works on next is fetched one step ahead (off the step's critical path;
        // s[LS] behind the last column is padding or the next sequence and never used)
        int nxt = (lane < lanes && LS > 0) ? (int)__ldg(s) : 0;

        for (int step = 0; step < nsteps; ++step) {
            const int n = step - lane;
            if (pass > 0 && (step & 31) == 0 && step < LS) {        // lane 0 enters a new chunk of 32 columns
                const int need = min(step + 32, LS);
#ifdef K2_TIMERS
                const long long tw0 = clock64();
#endif
                if (lane == 0) {
                    while (avail < need) avail = ld_acquire(prog_in);       // wait for the stripe above
                }
                __syncwarp();
#ifdef K2_TIMERS
                tm_wait += clock64() - tw0;
#endif
                if (step + lane < LS) chunk = __ldcg(row_in + step + lane);
            }
            const int in_h = __shfl_sync(FULL, chunk.x, step & 31);
            const int in_g = __shfl_sync(FULL, chunk.y, step & 31);
            int h_dn = K1_NEG, g_dn = K1_NEG;
            if (n >= 0 && n < LS && lane < lanes) {
                int h_up = recv_h, g_up = recv_g;
                if (lane == 0) {
                    if (pass == 0) { h_up = k1_top(g, n); g_up = K1_NEG; }
                    else { h_up = in_h; g_up = in_g; }
                }
                const int kL = n - lwm, kU = n - upm;
                if ((unsigned)kL < (unsigned)RR || (unsigned)kU < (unsigned)RR) {
#pragma unroll
                    for (int j = 0; j < RR / 4; ++j)
                        pk[j * 32] = make_int4(L.E[4 * j], L.E[4 * j + 1], L.E[4 * j + 2], L.E[4 * j + 3]);
                    int* pki = reinterpret_cast<int*>(pk);
                    if ((unsigned)kL < (unsigned)RR) pki[(kL >> 2) * 128 + (kL & 3)] = K1_NEG;
                    if ((unsigned)kU < (unsigned)RR) pki[(kU >> 2) * 128 + (kU & 3)] = K1_NEG;
#pragma unroll
                    for (int j = 0; j < RR / 4; ++j) {
                        int4 v = pk[j * 32];
                        L.E[4 * j] = v.x; L.E[4 * j + 1] = v.y; L.E[4 * j + 2] = v.z; L.E[4 * j + 3] = v.w;
                    }
                }
                const int letter = nxt;
                nxt = (int)__ldg(s + n + 1);
                const int4* pl = pp + letter * ((RR / 4) * 32);
                int sc[RR];
#pragma unroll
                for (int j = 0; j < RR / 4; ++j) {
                    int4 v = pl[j * 32];
                    sc[4 * j] = v.x; sc[4 * j + 1] = v.y; sc[4 * j + 2] = v.z; sc[4 * j + 3] = v.w;
                }
                const unsigned long long bits = k2_lane_step(L, sc, negv, h_up, g_up, mbase == 0, &h_dn, &g_dn);
                unsigned char* const w = words + (int64_t)step * 32 * (RR / 2);
                if (RR == 16) __stcs(reinterpret_cast<unsigned long long*>(w), bits);
                else if (RR == 8) __stcs(reinterpret_cast<unsigned*>(w), (unsigned)bits);
                else __stcs(reinterpret_cast<unsigned short*>(w), (unsigned short)bits);
                if (lane == lanes - 1 && !last_pass) {
                    __stcg(row_out + n, make_int2(h_dn, g_dn));
                    if ((n % PUB) == PUB - 1 || n == LS - 1)            // publish: data first, then the counter
                        st_release(a.progress + pass, n + 1);
                }
            }
            recv_h = __shfl_up_sync(FULL, h_dn, 1);
            recv_g = __shfl_up_sync(FULL, g_dn, 1);
        }
        if (last_pass) {
            const int tl = (rows_here - 1) / RR, kf = (rows_here - 1) % RR;
            int val = 0;
#pragma unroll
            for (int k = 0; k < RR; ++k)
                if (k == kf) val = L.H[k];
            val = __shfl_sync(FULL, val, tl);
            if (lane == 0) a.score[0] = val - (LQ + LS) * a.u;
        }
#ifdef K2_TIMERS
        { unsigned long long tm1; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(tm1));
          if (lane == 0 && (pass < 4 || pass % 16 == 0 || pass == npass - 1))
              printf("k2 stripe %d of %d: start %.1f us, ran %.1f us, waited %.1f us, %d steps\n", pass, npass, (double)(tm0 % 100000000000ull) * 1e-3,
                     (double)(tm1 - tm0) * 1e-3, tm_wait / 1965., nsteps); }
#endif
        __syncwarp();
    }
}

// one thread per alignment
__global__ void __launch_bounds__(128) k2_trace_kernel(const K2Args a, int npairs)
{
    for (int p = blockIdx.x * blockDim.x + threadIdx.x; p < npairs; p += gridDim.x * blockDim.x) {
        const int qi = a.pair_q[p], si = a.pair_s[p];
        const int LQ = a.seqs.wlen[qi], LS = a.seqs.wlen[si];
        const int ql = a.seqs.left[qi], sl = a.seqs.left[si];
        const int64_t off = a.len_off[p];
        int* out = a.out_pts + 2 * off;
        if (LQ == 0 || LS == 0) {       // nogap_skl (aln2.cc:30-40) in back-walk order
            out[0] = ql + LQ; out[1] = sl + LS; out[2] = ql; out[3] = sl;
            a.out_cnt[p] = 2;
            continue;
        }
        a.out_cnt[p] = k2_trace(a.dirs + a.dir_off[p], LQ, LS, a.rows_per_lane ? a.rows_per_lane : R, ql, sl, a.moves + off, a.recs + off, out);
    }
}


// ---- the same kernel, second form -----------------------------------------------------------------------------------
// Same stripes, same hand-over through L2 in 32-column chunks, same direction-word layout; what changed is inside a
// step: the query profile is laid out [letter][row][lane] as plain words (RR conflict-free LDS.32 instead of int4
// chunks that need RR % 4 == 0), and the band cut is an unrolled compare-and-move on the register array instead of a
// round trip of E[] through shared memory.  30 kb pair, RR = 8: 17.4 -> 15.3 ms.
// (The multi-warp variant of it, k2_fill_wide_kernel below, is opt-in.)
template <int RR, int CH, bool BULK>
__global__ void __launch_bounds__(32) k2_fill_long2_kernel(const K2Args a, int npass)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int lane = threadIdx.x;
    int* const prof = reinterpret_cast<int*>(smem_raw);            // [letter][k][lane]
    // BULK: the direction bits of DB steps (a contiguous DB * 32 * RR / 2 bytes of the wavefront-ordered store) are
    // collected in shared memory and leave the SM as ONE bulk copy (cp.async.bulk shared -> global, the TMA engine),
    // two buffers in flight; otherwise one 2 * RR-byte store per lane and step
    constexpr int DB = 16;
    constexpr unsigned DSTEP = 32 * (RR / 2);                       // bytes of direction bits per step
    const unsigned stage_sh = (unsigned)__cvta_generic_to_shared(smem_raw) + (((unsigned)a.dim * RR * 32 * 4 + 127u) & ~127u);
    const int negv = -a.v;
    const int qi = a.pair_q[0], si = a.pair_s[0];
    const uint8_t* q = a.seqs.res + a.seqs.offs[qi] + a.seqs.left[qi];
    const uint8_t* s = a.seqs.res + a.seqs.offs[si] + a.seqs.left[si];
    const int LQ = a.seqs.wlen[qi], LS = a.seqs.wlen[si];
    K1Geom g;
    g.LQ = LQ; g.LS = LS; g.u = a.u; g.v = a.v;
    k1_band(LQ, LS, a.sh, &g.lw, &g.up);
    g.topOpen = -a.v; g.topExt = -a.u; g.leftOpen = -a.v; g.leftExt = -a.u;
    unsigned char* const dir_base = reinterpret_cast<unsigned char*>(a.dirs + a.dir_off[0]);
    constexpr int RPP = 32 * RR;

    for (;;) {
        int pass = 0;
        if (lane == 0) pass = atomicAdd(a.ticket, 1);
        pass = __shfl_sync(FULL, pass, 0);
        if (pass >= npass) break;
        const int pbase = pass * RPP;
        for (int idx = lane; idx < a.dim * RPP; idx += 32) {
            const int letter = idx / RPP, rem = idx - letter * RPP;
            const int k = rem >> 5, ln = rem & 31;
            const int row = pbase + ln * RR + k;
            prof[idx] = row < LQ ? a.mtx[(int)q[row] * a.dim + letter] + 2 * a.u : 0;
        }
        __syncwarp();
        const int rows_here = min(LQ - pbase, RPP);
        const int lanes = (rows_here + RR - 1) / RR;
        const int mbase = pbase + lane * RR;
        const bool last_pass = pass == npass - 1;
        const int2* const row_in = a.rowbuf + (int64_t)(pass - 1) * LS;
        int2* const row_out = a.rowbuf + (int64_t)pass * LS;
        unsigned char* const words = dir_base + ((int64_t)pass * (LS + 31) * 32 + lane) * (RR / 2);

        K2Lane<RR> L;
        k2_lane_init(L, g, mbase);
        const int lwm = g.lw + mbase;
        const int upm = g.up + 1 + mbase;
        int recv_h = K1_NEG, recv_g = K1_NEG;
        const int nsteps = LS + lanes - 1;
        int2 chunk = make_int2(K1_NEG, K1_NEG);
        int nxt = (lane < lanes && LS > 0) ? (int)__ldg(s) : 0;

        // running pointers instead of index * stride products (64-bit multiplies per step otherwise)
        unsigned char* wp = words;                                  // direction bits of this lane at `step`
        const uint8_t* sp = s + 1 - lane;                           // residue of column n + 1 (running)
        int2* op = row_out - lane;                                  // slot of column n in the row below (running)
        const bool publisher = lane == lanes - 1 && !last_pass;     // the lane that owns the bottom row of the stripe
        // band cut as a bit mask per lane: bit k set <=> row k loses its horizontal input at this column
        // (k == n - lwm or k == n - upm); shifts beyond 31 (either side) give 0 in PTX
        const int cl0 = -lane - lwm, cu0 = -lane - upm;             // kL = step + cl0, kU = step + cu0
        const unsigned LSa = lane < lanes ? (unsigned)LS : 0u;      // this lane works at column n iff (unsigned)n < LSa

#pragma unroll K2_LONG_UNROLL_N
        for (int step = 0; step < nsteps; ++step, wp += 32 * (RR / 2), ++sp, ++op) {
            const int n = step - lane;
            int in_h = K1_NEG, in_g = K1_NEG;
            if (pass > 0) {                                         // uniform over the warp
                if ((step & (CH - 1)) == 0 && step < LS) {          // lane 0 enters a new chunk of CH columns
                    // the slots are their own flags: the warp re-reads its CH slots until none holds the fill pattern
                    const bool want = lane < CH && step + lane < LS;
                    unsigned long long w = 0;
                    do {
                        if (want) w = ld_relaxed64(row_in + step + lane);
                    } while (__any_sync(FULL, want && (unsigned)w == K2_SLOT_EMPTY));
                    chunk = make_int2((int)(unsigned)w, (int)(unsigned)(w >> 32));
                }
                in_h = __shfl_sync(FULL, chunk.x, step & (CH - 1));
                in_g = __shfl_sync(FULL, chunk.y, step & (CH - 1));
            }
            int h_dn = K1_NEG, g_dn = K1_NEG;
            if ((unsigned)n < LSa) {
                int h_up = recv_h, g_up = recv_g;
                if (lane == 0) {
                    if (pass == 0) { h_up = k1_top(g, n); g_up = K1_NEG; }
                    else { h_up = in_h; g_up = in_g; }
                }
                unsigned cut;
                asm("{ .reg .u32 ma, mb;\n\t"
                    "shl.b32 ma, 1, %1;\n\t"
                    "shl.b32 mb, 1, %2;\n\t"
                    "or.b32 %0, ma, mb; }" : "=r"(cut) : "r"(step + cl0), "r"(step + cu0));
#pragma unroll
                for (int k = 0; k < RR; ++k) if (cut & (1u << k)) L.E[k] = K1_NEG;
                const int letter = nxt;
                nxt = (int)__ldg(sp);
                int sc[RR];
#pragma unroll
                for (int k = 0; k < RR; ++k) sc[k] = prof[(letter * RR + k) * 32 + lane];
                const unsigned long long bits = k2_lane_step(L, sc, negv, h_up, g_up, mbase == 0, &h_dn, &g_dn);
                if (BULK) {
                    const unsigned sa = stage_sh + ((unsigned)step & (2 * DB - 1)) * DSTEP + (unsigned)lane * (RR / 2);
                    if (RR == 16) asm volatile("st.shared.b64 [%0], %1;" ::"r"(sa), "l"(bits) : "memory");
                    else if (RR == 8) asm volatile("st.shared.b32 [%0], %1;" ::"r"(sa), "r"((unsigned)bits) : "memory");
                    else asm volatile("st.shared.b16 [%0], %1;" ::"r"(sa), "h"((unsigned short)bits) : "memory");
                } else if (RR == 16) __stcs(reinterpret_cast<unsigned long long*>(wp), bits);
                else if (RR == 8) __stcs(reinterpret_cast<unsigned*>(wp), (unsigned)bits);
                else __stcs(reinterpret_cast<unsigned short*>(wp), (unsigned short)bits);
                // the bottom row goes to the stripe below: ONE predicated 64-bit store per column (single-copy atomic, so
                // (h, g) arrive together); no counter, no fence -- a slot that no longer holds the fill pattern is valid
                asm volatile("{ .reg .pred pp;\n\t"
                             ".reg .b64 pw;\n\t"
                             "setp.ne.b32 pp, %0, 0;\n\t"
                             "mov.b64 pw, {%2, %3};\n\t"
                             "@pp st.relaxed.gpu.global.b64 [%1], pw; }" ::"r"((int)publisher), "l"(op), "r"(h_dn), "r"(g_dn) : "memory");
            }
            recv_h = __shfl_up_sync(FULL, h_dn, 1);
            recv_g = __shfl_up_sync(FULL, g_dn, 1);
            if (BULK && (((step & (DB - 1)) == DB - 1) || step == nsteps - 1)) {        // uniform: a block of DB steps is complete
                const int s0 = step & ~(DB - 1);
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");          // every lane: its stores before the bulk copy
                __syncwarp();
                if (lane == 0) {
                    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;\n\t"
                                 "cp.async.bulk.commit_group;\n\t"
                                 "cp.async.bulk.wait_group.read 1;"                    // the other buffer has been read out
                                 ::"l"(words - (size_t)lane * (RR / 2) + (size_t)s0 * DSTEP),
                                   "r"(stage_sh + ((unsigned)s0 & (2 * DB - 1)) * DSTEP), "r"((unsigned)(step + 1 - s0) * DSTEP) : "memory");
                }
                __syncwarp();
            }
        }
        if (BULK) {
            if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
            __syncwarp();
        }
        if (last_pass) {
            const int tl = (rows_here - 1) / RR, kf = (rows_here - 1) % RR;
            int val = 0;
#pragma unroll
            for (int k = 0; k < RR; ++k)
                if (k == kf) val = L.H[k];
            val = __shfl_sync(FULL, val, tl);
            if (lane == 0) a.score[0] = val - (LQ + LS) * a.u;
        }
        __syncwarp();
    }
}


// ---- W consecutive stripes per CTA ---------------------------------------------------------------------------------
// Narrow stripes are faster per step (a lone warp runs a lane-step as one dependent chain), but every stripe of the
// kernels above trails the one above by ~87 steps (32-column chunks through L2).  Here W warps of one CTA, one per
// scheduler, hold W consecutive stripes and hand the bottom row down through a shared-memory ring, one column per step:
// warp w + 1 trails warp w by the 32 steps of the systolic skew and nothing more; only the last warp of a CTA goes through
// the L2 row buffer to the first warp of the next CTA.  Stripe = warp: the direction words keep k2_fill_kernel's layout.
// A ring slot = {h, g, column + 1, -} in one 16-byte shared-memory access, so value and tag travel together without a
// fence.  The lane that polls (lane 0) or publishes (the last lane) does so in a divergent branch: the __syncwarp() right
// after it matters -- without it the polling lane ran on ALONE through the whole cell code before the warp reconverged at
// the shuffles, i.e. every step was executed twice (38 - 47 ms instead of 14.5).
// OPT-IN (PG_K2_WIDE = 2 | 4 with PG_K2_LONG_ROWS = 4 | 8): measured against k2_fill_long2_kernel<4> on DNA pairs of 6 /
// 12 / 20 / 30 / 45 kb, 4 rows x 4 warps: 2.9 / 5.8 / 14.3 / 14.5 / 21.8 ms against 3.1 / 6.3 / 10.6 / 16.2 / 24.9 -- 8 to 13 %
// faster at four lengths and 35 % slower at one; the other shapes (8 x 2, 8 x 4, 4 x 2) run in the slow regime at 30 kb.
// Results are identical everywhere; what tips a run into the slow regime is not understood, so the default stays the
// single-warp stripe.
constexpr int WRING = 64;
__device__ __forceinline__ void ring_put(unsigned addr, int h, int g, int tag)
{
    asm volatile("st.volatile.shared.v4.s32 [%0], {%1, %2, %3, %3};" ::"r"(addr), "r"(h), "r"(g), "r"(tag) : "memory");
}
__device__ __forceinline__ int4 ring_get(unsigned addr)
{
    int4 v;
    asm volatile("ld.volatile.shared.v4.s32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr) : "memory");
    return v;
}
template <int RR, int W>
__global__ void __launch_bounds__(32 * W) k2_fill_wide_kernel(const K2Args a, int npass)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ __align__(16) int4 ring[W][WRING];
    __shared__ int cons[W];             // columns warp w has taken from the warp above
    __shared__ int sm_ticket;
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    int* const prof = reinterpret_cast<int*>(smem_raw) + (size_t)w * a.dim * RR * 32;      // [letter][k][lane]
    const int negv = -a.v;
    const int qi = a.pair_q[0], si = a.pair_s[0];
    const uint8_t* q = a.seqs.res + a.seqs.offs[qi] + a.seqs.left[qi];
    const uint8_t* s = a.seqs.res + a.seqs.offs[si] + a.seqs.left[si];
    const int LQ = a.seqs.wlen[qi], LS = a.seqs.wlen[si];
    K1Geom g;
    g.LQ = LQ; g.LS = LS; g.u = a.u; g.v = a.v;
    k1_band(LQ, LS, a.sh, &g.lw, &g.up);
    g.topOpen = -a.v; g.topExt = -a.u; g.leftOpen = -a.v; g.leftExt = -a.u;
    unsigned char* const dir_base = reinterpret_cast<unsigned char*>(a.dirs + a.dir_off[0]);
    constexpr int RPP = 32 * RR;
    volatile int* const vcons = cons;

    for (;;) {
        __syncthreads();
        if (threadIdx.x == 0) sm_ticket = atomicAdd(a.ticket, 1);
        if (lane == 0) cons[w] = 0;
        ring[w][lane] = make_int4(0, 0, 0, 0); ring[w][lane + 32] = make_int4(0, 0, 0, 0);      // tag 0 = nothing yet
        __syncthreads();
        const int pass0 = sm_ticket * W;
        if (pass0 >= npass) break;
        const int pass = pass0 + w;
        if (pass >= npass) continue;        // (the barrier at the top of the loop is reached by every warp)
        const int pbase = pass * RPP;
        for (int idx = lane; idx < a.dim * RPP; idx += 32) {
            const int letter = idx / RPP, rem = idx - letter * RPP;
            const int k = rem >> 5, ln = rem & 31;
            const int row = pbase + ln * RR + k;
            prof[idx] = row < LQ ? a.mtx[(int)q[row] * a.dim + letter] + 2 * a.u : 0;
        }
        __syncwarp();
        const int rows_here = min(LQ - pbase, RPP);
        const int lanes = (rows_here + RR - 1) / RR;
        const int mbase = pbase + lane * RR;
        const bool last_pass = pass == npass - 1;
        const bool from_l2 = w == 0 && pass > 0;                    // the stripe above lives in the previous CTA
        const bool from_ring = w > 0;
        const bool to_l2 = w == W - 1 && !last_pass;                // the stripe below lives in the next CTA
        const bool to_ring = w < W - 1 && !last_pass;
        const int2* const row_in = a.rowbuf + (int64_t)(pass - 1) * LS;
        int2* const row_out = a.rowbuf + (int64_t)pass * LS;
        const int* const prog_in = a.progress + (pass - 1);
        unsigned char* const words = dir_base + ((int64_t)pass * (LS + 31) * 32 + lane) * (RR / 2);

        K2Lane<RR> L;
        k2_lane_init(L, g, mbase);
        const int lwm = g.lw + mbase;
        const int upm = g.up + 1 + mbase;
        int recv_h = K1_NEG, recv_g = K1_NEG;
        const int nsteps = LS + lanes - 1;
        int avail = 0;
        int2 chunk = make_int2(K1_NEG, K1_NEG);
        int nxt = (lane < lanes && LS > 0) ? (int)__ldg(s) : 0;

        for (int step = 0; step < nsteps; ++step) {
            const int n = step - lane;
            int in_h = K1_NEG, in_g = K1_NEG;
            if (from_l2) {                                          // uniform over the warp
                if ((step & 31) == 0 && step < LS) {
                    const int need = min(step + 32, LS);
                    if (lane == 0) while (avail < need) avail = ld_acquire(prog_in);
                    __syncwarp();
                    if (step + lane < LS) chunk = __ldcg(row_in + step + lane);
                }
                in_h = __shfl_sync(FULL, chunk.x, step & 31);
                in_g = __shfl_sync(FULL, chunk.y, step & 31);
            } else if (from_ring) {                                 // uniform: the warp above, through the ring
                if (lane == 0 && step < LS) {
                    const unsigned slot = (unsigned)__cvta_generic_to_shared(&ring[w - 1][step & (WRING - 1)]);
                    int4 v = ring_get(slot);
                    while (v.z != step + 1) v = ring_get(slot);     // (a __nanosleep here oversleeps every step: 14.5 -> 26.7 ms)
                    in_h = v.x; in_g = v.y;
                    vcons[w] = step + 1;
                }
                __syncwarp();                                       // the other 31 lanes wait HERE, not at the shuffles below
            }
            int h_dn = K1_NEG, g_dn = K1_NEG;
            if (n >= 0 && n < LS && lane < lanes) {
                int h_up = recv_h, g_up = recv_g;
                if (lane == 0) {
                    if (pass == 0) { h_up = k1_top(g, n); g_up = K1_NEG; }
                    else { h_up = in_h; g_up = in_g; }
                }
                const int kL = n - lwm, kU = n - upm;
#pragma unroll
                for (int k = 0; k < RR; ++k) if (k == kL || k == kU) L.E[k] = K1_NEG;
                const int letter = nxt;
                nxt = (int)__ldg(s + n + 1);
                int sc[RR];
#pragma unroll
                for (int k = 0; k < RR; ++k) sc[k] = prof[(letter * RR + k) * 32 + lane];
                const unsigned long long bits = k2_lane_step(L, sc, negv, h_up, g_up, mbase == 0, &h_dn, &g_dn);
                unsigned char* const wp = words + (int64_t)step * 32 * (RR / 2);
                if (RR == 8) __stcs(reinterpret_cast<unsigned*>(wp), (unsigned)bits);
                else __stcs(reinterpret_cast<unsigned short*>(wp), (unsigned short)bits);
                if (lane == lanes - 1) {
                    if (to_l2) {
                        __stcg(row_out + n, make_int2(h_dn, g_dn));
                        if ((n % PUB) == PUB - 1 || n == LS - 1) st_release(a.progress + pass, n + 1);
                    } else if (to_ring) {
                        while (n - vcons[w + 1] >= WRING) { }       // the warp below has not read this slot's last value yet
                        ring_put((unsigned)__cvta_generic_to_shared(&ring[w][n & (WRING - 1)]), h_dn, g_dn, n + 1);
                    }
                }
            }
            __syncwarp();
            recv_h = __shfl_up_sync(FULL, h_dn, 1);
            recv_g = __shfl_up_sync(FULL, g_dn, 1);
        }
        if (last_pass) {
            const int tl = (rows_here - 1) / RR, kf = (rows_here - 1) % RR;
            int val = 0;
#pragma unroll
            for (int k = 0; k < RR; ++k)
                if (k == kf) val = L.H[k];
            val = __shfl_sync(FULL, val, tl);
            if (lane == 0) a.score[0] = val - (LQ + LS) * a.u;
        }
    }
}

// One WARP per alignment: the back-walk is a chain of dependent loads (the next cell is known only after the nibble of
// this one: ~0.35 us per move -- 20.9 ms for the 60,000 moves of a 30 kb pair, more than its fill).  Paths run straight
// most of the time, so the 32 lanes read the next 32 cells of the current run at once (diagonal in state H, up in G,
// left in F) and a ballot finds where the run really ends; one load latency per run (or per 32 moves) instead of per
// move.  Lane 0 then replays the moves (sequential memory, no dependent loads).
__device__ int k2_backwalk_warp(const unsigned long long* words, int LQ, int LS, int RL, unsigned char* moves)
{
    const int lane = threadIdx.x & 31;
    int nmv = 0, m = LQ - 1, n = LS - 1, state = 0;     // uniform over the warp
    while (m >= 0 && n >= 0) {
        const int mj = m - (state != 2 ? lane : 0), nj = n - (state != 1 ? lane : 0);
        const bool inr = mj >= 0 && nj >= 0;
        const unsigned nib = inr ? k2_nibble(words, LS, RL, mj, nj) : 0u;
        if (state == 0) {
            const unsigned bal = __ballot_sync(FULL, !inr || (nib & 3u) != 0u);
            const int f = bal ? __ffs(bal) - 1 : 32;            // lanes below f continue the diagonal run
            if (lane < f) moves[nmv + lane] = 1;
            nmv += f; m -= f; n -= f;
            if (f < 32) {
                const unsigned nf = __shfl_sync(FULL, nib, f);
                if (m >= 0 && n >= 0) state = (int)(nf & 3u);   // the cell where the run ends: a gap state takes over
            }
        } else {
            const unsigned obit = state == 1 ? 4u : 8u;         // the gap opened here: last cell of the run
            const unsigned bal = __ballot_sync(FULL, !inr || (nib & obit) != 0u);
            const int f = bal ? __ffs(bal) - 1 : 32;
            const bool opened = f < 32 && (state == 1 ? m - f >= 0 : n - f >= 0);
            if (lane < f) moves[nmv + lane] = (unsigned char)(state == 1 ? 3 : 5);
            if (opened && lane == f) moves[nmv + f] = (unsigned char)(state == 1 ? 2 : 4);
            const int used = f + (opened ? 1 : 0);
            nmv += used;
            if (state == 1) m -= used; else n -= used;
            if (opened) state = 0;
        }
    }
    for (int j = lane; j <= n; j += 32) moves[nmv + j] = 6;     // boundary row, then boundary column
    if (n >= 0) nmv += n + 1;
    for (int j = lane; j <= m; j += 32) moves[nmv + j] = 7;
    if (m >= 0) nmv += m + 1;
    __syncwarp();
    return nmv;
}

// The forward replay (k2_replay's record rules) by the whole warp: 32 moves per load, the runs inside a chunk found by
// ballot, one state update per RUN (a diagonal run can start a record only at its first move, a gap run only at its
// end).  Every lane carries the same state; lane 0 writes.  The path records chain strictly in order of creation (a gap
// run creates none between its opening and its end), so Vmf::traceback's list is the records in reverse order: the
// lanes copy it out side by side instead of walking the chain.
__device__ int k2_replay_warp(const unsigned char* moves, int nmv, int LQ, int LS, int ql, int sl, K2Rec* recs, int* out)
{
    const int lane = threadIdx.x & 31;
    int nrec = 0;
    if (lane == 0) { recs[0].m = ql; recs[0].n = sl; recs[0].p = -1; }
    nrec = 1;
    int hdir = K2_DIAG, hptr = 0, gdir = 0, gptr = 0, cm = -1, cn = -1;
    int base = nmv - 1;
    unsigned char mine = base - lane >= 0 ? moves[base - lane] : 0;     // forward order: lane j = j-th move of the chunk
    while (base >= 0) {
        const int cnt = base + 1 < 32 ? base + 1 : 32;
        const int nbase = base - 32;
        const unsigned char ahead = nbase - lane >= 0 ? moves[nbase - lane] : 0;    // next chunk, in flight while this one runs
        const int next_first = __shfl_sync(FULL, (int)ahead, 0);
        int j = 0;
        while (j < cnt) {
            const int c = __shfl_sync(FULL, (int)mine, j);
            const unsigned diff = __ballot_sync(FULL, lane >= j && lane < cnt && (int)mine != c);
            const int e = diff ? __ffs(diff) - 1 : cnt;
            int len = e - j;
            const int nxt = e < cnt ? __shfl_sync(FULL, (int)mine, e) : (nbase >= 0 ? next_first : 0);
            switch (c) {
            case 6: cn += len; hdir = K2_HORI; break;
            case 7: cm += len; hdir = K2_VERT; break;
            case 1:
                ++cm; ++cn;
                hdir = k2_isdiag(hdir) ? K2_DIAG : K2_NEWD;
                if (hdir == K2_NEWD) {
                    if (lane == 0) { recs[nrec].m = cm + ql; recs[nrec].n = cn + sl; recs[nrec].p = hptr; }
                    hptr = nrec++;
                }
                if (len > 1) { cm += len - 1; cn += len - 1; hdir = K2_DIAG; }
                break;
            case 2: case 4:         // gap openings: one cell each (several in a row are separate one-cell gaps)
                for (; len > 0; --len) {
                    if (c == 2) { ++cm; gdir = k2_ishori(hdir) ? K2_NEWV : K2_VERT; } else { ++cn; gdir = k2_isvert(hdir) ? K2_NEWH : K2_HORI; }
                    gptr = hptr;
                    const int after = len > 1 ? c : nxt;
                    if (after != c + 1) {   // the run ends here: H takes the gap state
                        hdir = gdir; hptr = gptr;
                        if (hdir == K2_NEWV || hdir == K2_NEWH) {
                            if (lane == 0) { recs[nrec].m = cm + ql; recs[nrec].n = cn + sl; recs[nrec].p = hptr; }
                            hptr = nrec++;
                        }
                    }
                }
                break;
            case 3: case 5:         // extensions
                if (c == 3) { cm += len; gdir = K2_VERT; } else { cn += len; gdir = K2_HORI; }
                if (nxt != c) { hdir = gdir; hptr = gptr; }
                break;
            default: break;
            }
            j = e;
        }
        base = nbase;
        mine = ahead;
    }
    __syncwarp();
    // final record (fwd2c.h:476) + the records in reverse order of creation (= Vmf::traceback, vmf.cc:103-119)
    if (lane == 0) { out[0] = LQ + ql; out[1] = LS + sl; }
    for (int q = lane; q < nrec; q += 32) {
        const K2Rec r = recs[nrec - 1 - q];
        out[2 * (q + 1)] = r.m; out[2 * (q + 1) + 1] = r.n;
    }
    return nrec + 1;
}

__global__ void __launch_bounds__(32) k2_trace_warp_kernel(const K2Args a, int npairs)
{
    for (int p = blockIdx.x; p < npairs; p += gridDim.x) {
        const int qi = a.pair_q[p], si = a.pair_s[p];
        const int LQ = a.seqs.wlen[qi], LS = a.seqs.wlen[si];
        const int ql = a.seqs.left[qi], sl = a.seqs.left[si];
        const int64_t off = a.len_off[p];
        int* out = a.out_pts + 2 * off;
        if (LQ == 0 || LS == 0) {
            if (threadIdx.x == 0) { out[0] = ql + LQ; out[1] = sl + LS; out[2] = ql; out[3] = sl; a.out_cnt[p] = 2; }
            continue;
        }
        const int nmv = k2_backwalk_warp(a.dirs + a.dir_off[p], LQ, LS, a.rows_per_lane ? a.rows_per_lane : R, a.moves + off);
        const int cnt = k2_replay_warp(a.moves + off, nmv, LQ, LS, ql, sl, a.recs + off, out);
        if (threadIdx.x == 0) a.out_cnt[p] = cnt;
        __syncwarp();
    }
}

}  // namespace

int k2_rows_per_lane() { return R; }
int k2_warps_per_block() { return NW; }
int k2_blocks_per_sm() { return BLOCKS_PER_SM; }

cudaError_t k2_fill_launch(const K2Args& a, int grid_blocks, cudaStream_t st)
{
    if (a.dim < 1 || a.dim > MAXDIM) return cudaErrorInvalidValue;
    const size_t smem = smem_bytes(a.dim);
    cudaError_t e = cudaFuncSetAttribute(k2_fill_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)smem_bytes(MAXDIM));
    if (e != cudaSuccess) return e;
    k2_fill_kernel<<<grid_blocks, NW * 32, smem, st>>>(a);
    return cudaGetLastError();
}

// rows per lane of the striped kernel for a query of LQ rows (PG_K2_LONG_ROWS = 4 / 8 / 16 overrides).  Measured on the
// 30 kb pair of config 5b (-DK2_TIMERS prints per-stripe times): with R = 8 the first stripe alone walks its 30,048
// steps in 14.1 ms (923 cycles per step: ~340 instructions of a warp that has its scheduler to itself, at its own
// instruction-level parallelism), the other 117 stripes finish 4.8 ms later (87 steps of skew each): R = 16 / 8 / 4:
// 23.1 / 17.0 / 18.3 ms (38 ms before the chunked input).  The per-step instruction count, not the hand-over, is the limit.
int k2_long_rows(int LQ, int LS)
{
    if (const char* e = getenv("PG_K2_LONG_ROWS")) {
        const int v = atoi(e);
        if (v == 4 || v == 8 || v == 16) return v;
    }
    (void)LS;
    // second form of the kernel (k2_fill_long2_kernel): 16 / 8 / 4 rows 21.1 / 16.9 / 16.1 ms on the 30 kb pair
    // with the flag-less hand-over narrow stripes win at every length (3 kb pair: 1.66 ms with 16 rows, 0.81 with 4)
    if (!getenv("PG_K2_LONG_V1")) return 4;
    return LQ >= 4096 ? 8 : 16;
}

template <int RR>
static cudaError_t long_launch(const K2Args& a, int npass, int sm_count, cudaStream_t st)
{
    const size_t smem = (size_t)(a.dim + 1) * (RR / 4) * 32 * sizeof(int4);
    cudaError_t e = cudaFuncSetAttribute(k2_fill_long_kernel<RR>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)((size_t)(MAXDIM + 1) * (RR / 4) * 32 * sizeof(int4)));
    if (e != cudaSuccess) return e;
    // every stripe in flight must be resident (a stripe waits for the one above): up to 16 single-warp CTAs per SM
    int blocks = npass < sm_count * 16 ? npass : sm_count * 16;
    k2_fill_long_kernel<RR><<<blocks, 32, smem, st>>>(a, npass);
    return cudaGetLastError();
}

template <int RR, int CH, bool BULK>
static cudaError_t long2_launch(const K2Args& a, int npass, int sm_count, size_t rowbuf_bytes, cudaStream_t st)
{
    const size_t stage = BULK ? (size_t)2 * 16 * 32 * (RR / 2) + 128 : 0;
    const size_t smem = (size_t)a.dim * RR * 32 * sizeof(int) + stage;
    cudaError_t e = cudaFuncSetAttribute(k2_fill_long2_kernel<RR, CH, BULK>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)((size_t)MAXDIM * RR * 32 * sizeof(int) + stage));
    if (e != cudaSuccess) return e;
    // every hand-over slot starts out "empty" (K2_SLOT_EMPTY)
    if ((e = cudaMemsetAsync(a.rowbuf, 0x80, rowbuf_bytes, st)) != cudaSuccess) return e;
    int blocks = npass < sm_count * 16 ? npass : sm_count * 16;
    k2_fill_long2_kernel<RR, CH, BULK><<<blocks, 32, smem, st>>>(a, npass);
    return cudaGetLastError();
}
template <int RR>
static cudaError_t long2_launch_ch(const K2Args& a, int npass, int sm_count, size_t rowbuf_bytes, cudaStream_t st)
{
    // A/B and test switches, read per call.  Hand-over chunk of 8 / 16 / 32 columns: 30 kb pair 8.58 / 8.09 / 8.30 ms.
    // PG_K2_BULK=1: the direction bits leave through bulk copies (UBLKCP.G.S, 1 KB per 16 steps) -- identical results,
    // 7.45 -> 8.79 ms: the proxy fence + hand-shake in front of every copy cost more than 16 two-byte stores per lane.
    const int ch = getenv("PG_K2_CHUNK") ? atoi(getenv("PG_K2_CHUNK")) : 16;
    const bool bulk = getenv("PG_K2_BULK") && getenv("PG_K2_BULK")[0] == '1';
    if (bulk) return long2_launch<RR, 16, true>(a, npass, sm_count, rowbuf_bytes, st);
    if (ch == 8) return long2_launch<RR, 8, false>(a, npass, sm_count, rowbuf_bytes, st);
    if (ch == 32) return long2_launch<RR, 32, false>(a, npass, sm_count, rowbuf_bytes, st);
    return long2_launch<RR, 16, false>(a, npass, sm_count, rowbuf_bytes, st);
}

template <int RR, int W>
static cudaError_t wide_launch(const K2Args& a, int npass, int sm_count, cudaStream_t st)
{
    const size_t smem = (size_t)W * a.dim * RR * 32 * sizeof(int);
    cudaError_t e = cudaFuncSetAttribute(k2_fill_wide_kernel<RR, W>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)((size_t)W * MAXDIM * RR * 32 * sizeof(int)));
    if (e != cudaSuccess) return e;
    const int ctas = (npass + W - 1) / W;       // a CTA waits only for one that took its ticket earlier: any grid works
    const int blocks = ctas < sm_count * 4 ? ctas : sm_count * 4;
    k2_fill_wide_kernel<RR, W><<<blocks, 32 * W, smem, st>>>(a, npass);
    return cudaGetLastError();
}

cudaError_t k2_fill_long_launch(const K2Args& a, int npass, int sm_count, size_t rowbuf_bytes, cudaStream_t st)
{
    if (const char* wv = getenv("PG_K2_WIDE")) {        // W stripes per CTA (2 or 4), rows per lane from PG_K2_LONG_ROWS (4 or 8)
        const int wn = atoi(wv);
        if (a.rows_per_lane == 4 && wn == 2) return wide_launch<4, 2>(a, npass, sm_count, st);
        if (a.rows_per_lane == 4 && wn == 4) return wide_launch<4, 4>(a, npass, sm_count, st);
        if (a.rows_per_lane == 8 && wn == 2) return wide_launch<8, 2>(a, npass, sm_count, st);
        if (a.rows_per_lane == 8 && wn == 4) return wide_launch<8, 4>(a, npass, sm_count, st);
    }
    if (!getenv("PG_K2_LONG_V1"))           // A/B switch: the first form of the kernel
        switch (a.rows_per_lane) {
        case 4: return long2_launch_ch<4>(a, npass, sm_count, rowbuf_bytes, st);
        case 8: return long2_launch_ch<8>(a, npass, sm_count, rowbuf_bytes, st);
        default: return long2_launch_ch<16>(a, npass, sm_count, rowbuf_bytes, st);
        }
    switch (a.rows_per_lane) {
    case 4: return long_launch<4>(a, npass, sm_count, st);
    case 8: return long_launch<8>(a, npass, sm_count, st);
    default: return long_launch<16>(a, npass, sm_count, st);
    }
}

cudaError_t k2_trace_launch(const K2Args& a, int npairs, cudaStream_t st)
{
    // few alignments (a long pair, a handful of candidates): one warp each, runs of the path read 32 cells at a time;
    // large batches: one thread each, the parallelism is across alignments
    const bool thread_only = getenv("PG_K2_TRACE_THREAD") != nullptr;           // A/B and test switch
    if (npairs <= 148 * 64 && !thread_only) {
        k2_trace_warp_kernel<<<npairs < 1 ? 1 : npairs, 32, 0, st>>>(a, npairs);
        return cudaGetLastError();
    }
    int blocks = (npairs + 127) / 128;
    if (blocks < 1) blocks = 1;
    if (blocks > 148 * 16) blocks = 148 * 16;
    k2_trace_kernel<<<blocks, 128, 0, st>>>(a, npairs);
    return cudaGetLastError();
}
