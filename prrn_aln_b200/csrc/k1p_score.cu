// k1p_score.cu -- kernel K1P: the packed (int16 x 2) form of K1.  Every lane register holds two
// alignments -- two QUERIES (rows) against the same subject (columns) -- so each DPX instruction
// (__viaddmax_s16x2 / __vimax_s16x2, SASS VIADDMNMX.S16x2 / VIMNMX.S16x2) retires two cells and all
// per-step overhead (profile fetch, band cut, shuffles, index math) is shared by both.
//
// Same reference semantics as k1_score.cu: alnScoreD / Fwd2d::forwardD (src/fwd2d1.cc:57-160) +
// the dpscore epilogue (src/phyl.cc:221-251, src/aln2.cc:321-333).
//   CTA  = one work item = a PAIR of queries (q0, q1) x a run of subjects; the packed profile
//          P[letter][row] = (S'(q0_row, letter), S'(q1_row, letter)) lives in shared memory in the
//          conflict-free [letter][R/4][lane][4] layout (one LDS.128 = 4 rows x 2 alignments).
//   half-warp (16 lanes) = one subject at a time; lane t of it owns rows [R t, R t + R) of both queries.
// The host (pg_api.cu) chooses query pairs of similar length and gives every unordered pair {x, y}
// to exactly one (query, subject) slot (round-robin tournament), with two "valid" bits per subject.
#include <cuda_runtime.h>
#include <stdint.h>

#include "k1p_core.cuh"
#include "pg_internal.h"

namespace {

constexpr int RMAX = 28;        // largest rows-per-lane variant
constexpr int NW = 8;
constexpr int HL = 16;          // lanes of one systolic array (half a warp)
constexpr int NHW = NW * 32 / HL;
constexpr int MAXDIM = 32;
constexpr int BLOCKS_PER_SM = 2;
constexpr unsigned FULL = 0xffffffffu;
constexpr int QMAX = RMAX / 4;

constexpr int POKE_WORDS = RMAX + 8;     // band-cut scratch words per lane (poke_stride(RMAX) at most)

__host__ __device__ inline size_t smem_bytes(int dim)
{
    return (size_t)dim * QMAX * HL * sizeof(uint4) + (size_t)NW * 32 * POKE_WORDS * 4 + 16;
}

__device__ __forceinline__ void store_result(const K1PArgs& a, int score, int qi, int si, int LQ, int LS)
{
    // condensed slot of the unordered pair {qi, si} (src/cmn.h:115)
    const int64_t hi = qi > si ? qi : si, lo = qi > si ? si : qi;
    int64_t slot = hi * (hi - 1) / 2 + lo;
    if (slot < a.k_begin || slot >= a.k_end) return;
    slot -= a.k_begin;
    const int dl = LQ > LS ? LQ - LS : LS - LQ;
    if (a.epilogue == PG_EPI_DIST_F32) {
        float denome = __fsqrt_rn(__fmul_rn((float)a.self[qi], (float)a.self[si]));
        float scr = __fadd_rn((float)score, __fdiv_rn(__fmul_rn(a.u_f32, (float)dl), 2.f));
        float dst = (float)__dsub_rn(1.0, (double)__fdiv_rn(scr, denome));
        reinterpret_cast<float*>(a.out)[slot] = (float)__dmul_rn(100.0, (double)dst);
    } else if (a.epilogue == PG_EPI_DIST_F64) {
        double denome = __dsqrt_rn(__dmul_rn((double)a.self[qi], (double)a.self[si]));
        double scr = __dadd_rn((double)score, (double)__fdiv_rn(__fmul_rn(a.u_f32, (float)dl), 2.f));
        double dst = __dsub_rn(1.0, __ddiv_rn(scr, denome));
        reinterpret_cast<double*>(a.out)[slot] = __dmul_rn(100.0, dst);
    } else if (a.epilogue == PG_EPI_SCORE_F32) {
        reinterpret_cast<float*>(a.out)[slot] = (float)score;
    } else {
        reinterpret_cast<double*>(a.out)[slot] = (double)score;
    }
}

// shared-memory accesses by 32-bit address (keeps the per-lane base in one register, no generic-pointer math)
__device__ __forceinline__ uint4 lds128(unsigned addr)
{
    uint4 v;
    asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr));
    return v;
}
__device__ __forceinline__ void sts128(unsigned addr, uint4 v)
{
    asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
__device__ __forceinline__ void sts16(unsigned addr, unsigned short v)
{
    asm volatile("st.shared.u16 [%0], %1;" ::"r"(addr), "h"(v) : "memory");
}
// words of band-cut scratch per lane: RP rows + padding that makes the lane stride an odd number of 16-byte
// chunks (conflict-free 128-bit accesses with rows contiguous per lane, so a row's address is linear in n)
__host__ __device__ constexpr int poke_stride(int rp) { return ((rp / 4 + 1) % 2 ? rp + 4 : rp + 8); }

// One work item with R rows per lane.  A HALF-WARP (16 lanes) is one systolic array: the two halves of a warp
// run two different subjects of the item against the same pair of queries (same profile), so every per-step
// instruction that is not a DPX instruction -- profile fetch, band cut, shuffles, loop control -- is paid once
// per 2 x R x 2 cells.  R in {16..28}: the host picks the smallest variant whose 16*R rows hold the longer query.
// The profile layout is [letter][quad][half-lane][4] with Q = ceil(R/4) quads per lane (a quarter-warp reads
// 128 contiguous bytes: conflict-free LDS.128).
// MP = the longer query needs more than one pass (bottom row parked in a.rowbuf between passes).
//
// The ALU pipe (DPX instructions included) issues at half the rate of the scheduler, so everything that is
// not DPX is kept off it: addresses are IMADs on a per-lane shared-memory base, the top boundary is a
// register that changes at two columns per alignment, one unsigned compare tests "lane inside the matrix".
template <int R, bool MP>
__device__ __forceinline__ void process_item(const K1PArgs& a, const PgItem2& item, uint4* sm_prof, unsigned* sm_poke)
{
    constexpr int Q = (R + 3) / 4;
    constexpr int RP = 4 * Q;
    constexpr int PS = poke_stride(RP);
    constexpr int ROWS_PER_PASS = HL * R;
    const int tid = threadIdx.x;
    const int lane = tid & 31;
    const int hl = lane & (HL - 1);                     // lane inside the half-warp
    const int hw = tid / HL;                            // half-warp of the CTA
    const int ghw = blockIdx.x * NHW + hw;
    const unsigned negv2 = k1p_pack(-a.v, -a.v);
    const unsigned neg2 = k1p_pack(K1P_NEG, K1P_NEG);
    const int qi0 = item.q0, qi1 = item.q1;
    const uint8_t* q0 = a.seqs.res + a.seqs.offs[qi0] + a.seqs.left[qi0];
    const uint8_t* q1 = a.seqs.res + a.seqs.offs[qi1] + a.seqs.left[qi1];
    const int LQ0 = a.seqs.wlen[qi0], LQ1 = a.seqs.wlen[qi1];
    const int LQ = max(LQ0, LQ1);
    const int npass = MP ? (LQ + ROWS_PER_PASS - 1) / ROWS_PER_PASS : 1;
    unsigned prof_sh = (unsigned)__cvta_generic_to_shared(sm_prof + hl);
    unsigned pk_sh = (unsigned)__cvta_generic_to_shared(sm_poke + tid * PS);
    asm volatile("" : "+r"(prof_sh), "+r"(pk_sh));      // opaque: keep both bases in registers, never recomputed per step
    const bool lane0 = hl == 0;

    for (int pass = 0; pass < npass; ++pass) {
        const int pbase = pass * ROWS_PER_PASS;
        {
            unsigned* p = reinterpret_cast<unsigned*>(sm_prof);
            const int total = a.dim * HL * RP;
            for (int idx = tid; idx < total; idx += NW * 32) {
                int letter = idx / (HL * RP);
                int rem = idx - letter * (HL * RP);
                int j = rem / (HL * 4), ln = (rem >> 2) & (HL - 1), c = rem & 3;
                int kk = j * 4 + c;
                int row = pbase + ln * R + kk;
                bool ok = kk < R;
                int s0 = ok && row < LQ0 ? a.mtx[(int)q0[row] * a.dim + letter] + 2 * a.u : 0;
                int s1 = ok && row < LQ1 ? a.mtx[(int)q1[row] * a.dim + letter] + 2 * a.u : 0;
                p[idx] = k1p_pack(s0, s1);
            }
        }
        __syncthreads();

        const int rows_here = min(LQ - pbase, ROWS_PER_PASS);
        const int lanes = (rows_here + R - 1) / R;
        const int mbase = pbase + hl * R;
        const bool last_pass = pass == npass - 1;
        const bool fin0 = LQ0 > pbase && LQ0 <= pbase + ROWS_PER_PASS;
        const bool fin1 = LQ1 > pbase && LQ1 <= pbase + ROWS_PER_PASS;
        const int r0 = LQ0 - 1 - pbase, r1 = LQ1 - 1 - pbase;
        const bool top_lane = lane0 && pass == 0;
        const unsigned tmask = top_lane ? 0xffffffffu : 0u;

        // both halves of a warp walk the subject list together; a half without a subject idles
        for (int sub0 = item.sub_begin + (hw & ~1), ord = 0; sub0 < item.sub_end; sub0 += NHW, ++ord) {
            const int sub = sub0 + (hw & 1);
            const bool have = sub < item.sub_end;
            const unsigned ent = have ? a.subs[sub] : 0u;
            const int si = (int)(ent & 0x3fffffffu);
            const bool v0 = (ent >> 30) & 1u, v1 = (ent >> 31) & 1u;
            const unsigned soff = (unsigned)(a.seqs.offs[si] + a.seqs.left[si]);    // < 2^31: checked by the host
            const int LS = have ? a.seqs.wlen[si] : 0;
            K1Geom g0, g1;
            g0.LQ = LQ0; g0.LS = LS; g0.u = a.u; g0.v = a.v;
            g1.LQ = LQ1; g1.LS = LS; g1.u = a.u; g1.v = a.v;
            k1_band(LQ0, LS, a.sh, &g0.lw, &g0.up);
            k1_band(LQ1, LS, a.sh, &g1.lw, &g1.up);
            g0.topOpen = g1.topOpen = g0.leftOpen = g1.leftOpen = -a.v;
            g0.topExt = g1.topExt = g0.leftExt = g1.leftExt = -a.u;
            // empty sequences never reach this kernel (the host routes such batches to the int32 kernel)
            const bool live = have && LS > 0 && LQ0 > 0 && LQ1 > 0;
            uint2* rowbuf = MP && a.rowbuf ? a.rowbuf + ((int64_t)ghw * a.rowbuf_stride + (int64_t)ord * LS) : nullptr;

            K1PLane<R> L;
            k1p_lane_init(L, g0, g1, mbase, negv2);
            // band cut: row k of this lane loses its horizontal input at column n when k == n - lwm (diagonal
            // lw) or k == n - upm (diagonal up + 1), per packed half
            const int lwm0 = g0.lw + mbase, upm0 = g0.up + 1 + mbase;
            const int lwm1 = g1.lw + mbase, upm1 = g1.up + 1 + mbase;
            // top boundary of pass 0 as seen by lane 0 (n == step): the drifted value is the constant
            // topOpen while n + 1 <= up, then -inf; it changes at columns up0 and up1 only
            const int tslope = g0.topExt + g0.u;
            unsigned topv = neg2, topf = neg2;
            int tev = -1;                               // next column at which topv changes (lane 0, pass 0)
            if (top_lane) {
                topv = k1p_pack(1 <= g0.up ? g0.topOpen + tslope : K1P_NEG, 1 <= g1.up ? g1.topOpen + tslope : K1P_NEG);
                topf = K1P_ADDMAX(topv, negv2, neg2);
                tev = tslope != 0 ? 1 : min(g0.up >= 1 ? g0.up : 0x7fffffff, g1.up >= 1 ? g1.up : 0x7fffffff);
            }
            unsigned recv_h = neg2, recv_f = neg2;
            int nsteps = live ? LS + lanes - 1 : 0;
            nsteps = max(nsteps, __shfl_xor_sync(FULL, nsteps, HL));
            const unsigned LSa = live && hl < lanes ? (unsigned)LS : 0u;    // lane works at column n iff (unsigned)n < LSa
            const unsigned sl = soff - (unsigned)hl;                        // letter of this lane at `step` = res[sl + step]
            unsigned nxt = __ldg(a.seqs.res + soff);                        // every lane starts at column 0

            for (int step = 0; step < nsteps; ++step) {
                const int n = step - hl;
                unsigned h_dn = neg2, f_dn = neg2;
                if ((unsigned)n < LSa) {
                    unsigned h_up = recv_h, f_up = recv_f;
                    if (n == tev) {                                     // rare: lane 0, a few times per alignment
                        const int kk = n + 1;
                        topv = k1p_pack(kk <= g0.up ? g0.topOpen + kk * tslope : K1P_NEG,
                                        kk <= g1.up ? g1.topOpen + kk * tslope : K1P_NEG);
                        topf = K1P_ADDMAX(topv, negv2, neg2);
                        tev = tslope != 0 ? n + 1
                                          : min(g0.up > n ? g0.up : 0x7fffffff, g1.up > n ? g1.up : 0x7fffffff);
                    }
                    h_up = (h_up & ~tmask) | (topv & tmask);           // lane 0 of pass 0 takes the top boundary
                    f_up = (f_up & ~tmask) | (topf & tmask);
                    if (MP) {
                        if (pass > 0 && lane0) { uint2 v = __ldcg(rowbuf + n); h_up = v.x; f_up = v.y; }
                    }
                    const unsigned kL0 = (unsigned)(n - lwm0), kU0 = (unsigned)(n - upm0);
                    const unsigned kL1 = (unsigned)(n - lwm1), kU1 = (unsigned)(n - upm1);
                    if (min(min(kL0, kU0), min(kL1, kU1)) < (unsigned)R) {
#pragma unroll
                        for (int j = 0; j < Q; ++j)
                            sts128(pk_sh + 16 * j, make_uint4(L.E[4 * j], 4 * j + 1 < R ? L.E[4 * j + 1] : 0u,
                                                              4 * j + 2 < R ? L.E[4 * j + 2] : 0u, 4 * j + 3 < R ? L.E[4 * j + 3] : 0u));
                        if (kL0 < (unsigned)R) sts16(pk_sh + 4 * kL0, (unsigned short)K1P_NEG);
                        if (kU0 < (unsigned)R) sts16(pk_sh + 4 * kU0, (unsigned short)K1P_NEG);
                        if (kL1 < (unsigned)R) sts16(pk_sh + 4 * kL1 + 2, (unsigned short)K1P_NEG);
                        if (kU1 < (unsigned)R) sts16(pk_sh + 4 * kU1 + 2, (unsigned short)K1P_NEG);
#pragma unroll
                        for (int j = 0; j < Q; ++j) {
                            uint4 v = lds128(pk_sh + 16 * j);
                            L.E[4 * j] = v.x;
                            if (4 * j + 1 < R) L.E[4 * j + 1] = v.y;
                            if (4 * j + 2 < R) L.E[4 * j + 2] = v.z;
                            if (4 * j + 3 < R) L.E[4 * j + 3] = v.w;
                        }
                    }
                    // the letter of the next column is fetched one step ahead (s[LS] at the last column is padding
                    // or the next sequence: never used), so no global-load latency sits in front of the profile rows
                    const unsigned letter = nxt;
                    nxt = __ldg(a.seqs.res + (sl + (unsigned)step + 1u));
                    const unsigned pa = prof_sh + letter * (unsigned)(Q * HL * sizeof(uint4));
                    unsigned sc[RP];
#pragma unroll
                    for (int j = 0; j < Q; ++j) {
                        uint4 v = lds128(pa + j * (unsigned)(HL * sizeof(uint4)));
                        sc[4 * j] = v.x; sc[4 * j + 1] = v.y; sc[4 * j + 2] = v.z; sc[4 * j + 3] = v.w;
                    }
                    k1p_lane_step(L, sc, negv2, h_up, f_up, &h_dn, &f_dn);
                    if (MP) {
                        if (hl == HL - 1 && !last_pass) __stcg(rowbuf + n, make_uint2(h_dn, f_dn));
                    }
                }
                recv_h = __shfl_up_sync(FULL, h_dn, 1, HL);
                recv_f = __shfl_up_sync(FULL, f_dn, 1, HL);
            }

            if (fin0 || fin1) {
                unsigned val0 = 0, val1 = 0;
                const int k0f = r0 >= 0 ? r0 % R : -1, k1f = r1 >= 0 ? r1 % R : -1;
#pragma unroll
                for (int k = 0; k < R; ++k) {
                    if (k == k0f) val0 = L.H[k];
                    if (k == k1f) val1 = L.H[k];
                }
                val0 = __shfl_sync(FULL, val0, (r0 >= 0 ? r0 / R : 0) & (HL - 1), HL);
                val1 = __shfl_sync(FULL, val1, (r1 >= 0 ? r1 / R : 0) & (HL - 1), HL);
                if (lane0 && live) {
                    if (fin0 && v0) store_result(a, k1p_lo(val0) - (LQ0 + LS) * a.u, qi0, si, LQ0, LS);
                    if (fin1 && v1) store_result(a, k1p_hi(val1) - (LQ1 + LS) * a.u, qi1, si, LQ1, LS);
                }
            }
        }
        __syncthreads();
    }
}

template <int R>
__device__ __forceinline__ void process_item_r(const K1PArgs& a, const PgItem2& item, uint4* sm_prof, unsigned* sm_poke)
{
    const int LQ = max(a.seqs.wlen[item.q0], a.seqs.wlen[item.q1]);
    if (LQ <= HL * R) process_item<R, false>(a, item, sm_prof, sm_poke);
    else process_item<R, true>(a, item, sm_prof, sm_poke);
}

__global__ void __launch_bounds__(NW * 32, BLOCKS_PER_SM) k1p_score_kernel(const K1PArgs a)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    uint4* const sm_prof = reinterpret_cast<uint4*>(smem_raw);
    unsigned* const sm_poke = reinterpret_cast<unsigned*>(sm_prof + a.dim * QMAX * HL);
    int* const sm_item = reinterpret_cast<int*>(sm_poke + NW * 32 * POKE_WORDS);
    for (;;) {
        if (threadIdx.x == 0) *sm_item = atomicAdd(a.counter, 1);
        __syncthreads();
        const int it = *sm_item;
        if (it >= a.nitems) break;
        const PgItem2 item = a.items[it];
        switch (item.rows) {
        case 16: process_item_r<16>(a, item, sm_prof, sm_poke); break;
        case 20: process_item_r<20>(a, item, sm_prof, sm_poke); break;
        case 24: process_item_r<24>(a, item, sm_prof, sm_poke); break;
        case 26: process_item_r<26>(a, item, sm_prof, sm_poke); break;
        default: process_item_r<28>(a, item, sm_prof, sm_poke); break;
        }
    }
}

}  // namespace

int k1p_rows_per_pass() { return HL * RMAX; }
int k1p_pick_rows(int lq)
{
    const int opts[5] = {16, 20, 24, 26, 28};
    for (int i = 0; i < 5; ++i)
        if (lq <= HL * opts[i]) return opts[i];
    return RMAX;
}
int k1p_warps_per_block() { return NHW; }      // systolic arrays (half-warps) per CTA: the host's unit of subjects
int k1p_blocks_per_sm() { return BLOCKS_PER_SM; }

cudaError_t k1p_launch(const K1PArgs& a, int grid_blocks, cudaStream_t st)
{
    if (a.dim < 1 || a.dim > MAXDIM) return cudaErrorInvalidValue;
    const size_t smem = smem_bytes(a.dim);
    cudaError_t e = cudaFuncSetAttribute(k1p_score_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)smem_bytes(MAXDIM));
    if (e != cudaSuccess) return e;
    k1p_score_kernel<<<grid_blocks, NW * 32, smem, st>>>(a);
    return cudaGetLastError();
}
