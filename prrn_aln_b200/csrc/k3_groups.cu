// k3_groups.cu -- kernel K3: group-to-group banded alignment WITH path and gap-profile state (sm_100a).
//
// Stands behind alignC<DPunit | DPunit_hf | DPunit_pf> (reference src/fwd2c.h:670-677: Fwd2c ctor :81,
// initB :138, forwardB :359, Vmf::traceback src/vmf.cc:103) as align2 dispatches it for NGP_ALB /
// HLF_ALB / RHF_ALB / GPF_ALB (src/maln2.cc:1899-1910), affine and two-piece (Noll 2 / 3).
//
// Machine mapping
//   CTA    = one alignment at a time (persistent CTAs pull pairs, heaviest first, from an atomic queue).
//   thread = one ROW of the DP matrix; the CTA sweeps ANTI-DIAGONALS: at step s thread t computes cell
//            (m = pass*T + t, n = s - t), one __syncthreads per step.  Rows beyond T take further passes;
//            the bottom row of a pass is parked in a row buffer and re-enters as the top of the next.
//   records  The reference's DPunit_hf / DPunit_pf records (value, direction, path pointer and the
//            IDELTA lists of the dynamic gap state) are fixed-stride word runs (k3_core.cuh) in an
//            L2-resident per-CTA arena; H records rotate through three generations per row so that
//            the diagonal / upper / left neighbours are plain reads of what the neighbouring thread
//            published one and two steps ago; G, G2 through two.
//   path   = the reference's Vmf: an append-only record store per CTA (shared-memory cursor); the same
//            thread block walks the pointer chain back at the end and emits the corner list in Vmf
//            back-walk order, so the output equals alignC's, quirks included.
//   sim2   = one contraction X_a[m] . Y_b[n] over residue codes (all sim11..sim33 variants, staged by
//            the host layer); evaluated per cell here, or read from the tile matrix K4 precomputed.
#include <cooperative_groups.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "k3_core.cuh"
#include "k3r_core.cuh"
#include "pg_internal.h"

namespace {

constexpr int CTA = K3_THREADS;     // threads per CTA; an alignment gets TG of them (TG = 32 .. 256)
constexpr int RING = 4;             // prefetch ring depth for the parked row (records n, n+1 are read at step n)
constexpr int PFN = 4;              // prefetch words per thread (3 * stride <= PFN * TG, else no ring)
constexpr int XD = 8;               // cluster variant: depth of the ring of neighbour records (steps a CTA may lead)
constexpr int TGCL = 192;           // cluster variant: rows per CTA (3 x 192 = 576 threads leave 112 registers per thread;
                                    // with 3 x 256 the extra hand-over state spilled into local memory: +2.4 us per step)
constexpr int TGRL = 128;           // register-list form (k3r_core.cuh): rows per CTA; 3 x 128 threads, two CTAs per SM,
                                    // 168 registers per thread (the records and lists of a cell live in registers)

// words of shared memory one alignment needs for its wavefront records
__host__ __device__ inline size_t k3_smem_words(int st, int Noll, int tg)
{
    return (size_t)st * ((Noll == 3 ? 9 : 6) * tg + 1 + 3 * RING + 4 + 3 * XD);
}
// SM variant of the kernel (every record operand of a cell lives in shared memory, so the compiler emits
// LDS / STS with 32-bit addresses instead of generic loads): needs the wavefront records in shared memory
// and the prefetch ring of the parked row
__host__ __device__ inline bool k3_sm_ok(int st, int Noll, int tg, size_t smem_words_per_group)
{
    return k3_smem_words(st, Noll, tg) <= smem_words_per_group && (Noll == 3 ? 3 : 2) * st <= PFN * tg;
}

// barrier over the TG threads that share one alignment (named barrier g + 1; a warp needs none)
template <int TG>
__device__ __forceinline__ void group_sync(int g)
{
    if (TG == 32) __syncwarp();
    else if (TG == CTA) __syncthreads();
    else asm volatile("bar.sync %0, %1;" ::"r"(g + 1), "r"(TG) : "memory");
}

// Hand-over between the CTAs of a cluster.  A cluster-wide barrier (or a neighbour hand-shake) per anti-diagonal
// keeps all NC x 256 rows in lock step, and a step then lasts as long as the slowest of ~1,000 cells (measured:
// 6-7 us against 3.5 us of one CTA).  The CTAs are therefore decoupled: CTA c PUSHES the records of its last row
// into a ring of XD slots in the shared memory of CTA c+1 (plain DSMEM stores by one warp, then
// mbarrier.arrive.release.cluster on that CTA's full[slot]); CTA c+1 waits for full[(S-1) % XD] when its first
// row needs step S-1 of the row above, and returns the slot with a remote arrive on empty[] once it has used it
// as "above" (step +1) and "diagonal" (step +2).  A CTA may lead its successor by XD - 2 steps; inside a CTA
// the usual two block barriers per step remain.
__device__ __forceinline__ unsigned smem_addr(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
// rank / size of the cluster straight from the special registers (pure: the compiler re-reads them instead of
// keeping -- and spilling -- loop-invariant copies; the cluster kernel runs at the 80-register limit)
__device__ __forceinline__ int cl_rank() { unsigned r; asm("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return (int)r; }
__device__ __forceinline__ int cl_size() { unsigned r; asm("mov.u32 %0, %%cluster_nctarank;" : "=r"(r)); return (int)r; }
// 8-byte store into the shared memory of another CTA of the cluster (address of the same variable here + rank)
__device__ __forceinline__ void st_remote_v2(unsigned local_addr, unsigned rank, int2 v)
{
    asm volatile("{ .reg .b32 ra; mapa.shared::cluster.u32 ra, %0, %1; st.shared::cluster.v2.u32 [ra], {%2, %3}; }"
                 ::"r"(local_addr), "r"(rank), "r"(v.x), "r"(v.y) : "memory");
}
__device__ __forceinline__ void mbar_init(unsigned addr, unsigned count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(addr), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_remote_arrive(unsigned local_addr, unsigned rank)
{
    asm volatile("{ .reg .b32 ra; mapa.shared::cluster.u32 ra, %0, %1; "
                 "mbarrier.arrive.release.cluster.shared::cluster.b64 _, [ra]; }" ::"r"(local_addr), "r"(rank) : "memory");
}
// The release above is MEMBAR.ALL.CTA + MEMBAR.ALL.GPU in SASS, in front of every per-step signal.  The default
// hand-over (K3Args.cluster_fence == 0) therefore lets the DATA complete the barrier: the records travel as st.async
// (each store performs complete_tx of its 8 bytes on full[slot] of the receiving CTA), the receiving CTA arms
// full[slot] itself for every use (arrive.expect_tx with the bytes of one push, a local operation), and the only
// remote arrive left -- "slot free again", which orders no data -- is relaxed.
__device__ __forceinline__ void mbar_remote_arrive_relaxed(unsigned local_addr, unsigned rank)
{
    asm volatile("{ .reg .b32 ra; mapa.shared::cluster.u32 ra, %0, %1; "
                 "mbarrier.arrive.relaxed.cluster.shared::cluster.b64 _, [ra]; }" ::"r"(local_addr), "r"(rank) : "memory");
}
__device__ __forceinline__ void mbar_arm_tx(unsigned addr, unsigned bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(addr), "r"(bytes) : "memory");
}
__device__ __forceinline__ void st_remote_async_v2(unsigned local_addr, unsigned local_mbar, unsigned rank, int2 v)
{
    asm volatile("{ .reg .b32 ra, rm; mapa.shared::cluster.u32 ra, %0, %2; mapa.shared::cluster.u32 rm, %1, %2; "
                 "st.async.shared::cluster.mbarrier::complete_tx::bytes.v2.b32 [ra], {%3, %4}, [rm]; }"
                 ::"r"(local_addr), "r"(local_mbar), "r"(rank), "r"(v.x), "r"(v.y) : "memory");
}
// bytes one push puts into a ring slot: nrec records of ceil(st / 2) 8-byte stores
__device__ __forceinline__ unsigned cl_push_bytes(int st, bool three) { return (three ? 3u : 2u) * (unsigned)((st + 1) / 2) * 8u; }
// The wait acquires at CTA scope by default: what a neighbour hands over per step are shared-memory records, read
// through DSMEM (never cached in L1) after this wait and a __syncthreads, and parked rows in global memory that are
// read with ld.global.cg hundreds of steps later; a cluster-scope acquire makes ptxas invalidate L1 (CCTL.IVALL)
// on every step, where the latency mode keeps the gap-profile lists (measured: prrn5 C3 kernels 22.8 s against
// 25.1 s).  fenced = true (PG_K3_CLUSTER_FENCE=1) selects the cluster-scope form; the tests run both.
__device__ __forceinline__ void mbar_wait(unsigned addr, unsigned parity, bool fenced)
{
    unsigned ok;
    if (fenced) {
        do {
            asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                         : "=r"(ok) : "r"(addr), "r"(parity) : "memory");
        } while (!ok);
    } else {
        do {
            asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                         : "=r"(ok) : "r"(addr), "r"(parity) : "memory");
        } while (!ok);
    }
}

// CL, end of step S, last warp of role 0: push the last row (H, G, G2 records of thread 255) into slot S % XD of the
// CTA below once that slot is free again, then signal it.  Not inlined: its temporaries must not lengthen live
// ranges inside the per-cell code (the kernel sits at the 80-register limit).
__device__ __noinline__ void cl_push(const int* srcH, const int* srcG, const int* srcG2, unsigned ring_addr, unsigned full_addr,
                                     unsigned empty_addr, int S, int st, int lane, int fence_mode)
{
    const int xs = S % XD;
    if (S >= XD && lane == 0) mbar_wait(empty_addr + 8u * xs, (unsigned)(S / XD - 1) & 1u, fence_mode == 1);
    __syncwarp();
    const unsigned below = (unsigned)(cl_rank() + 1);
    const unsigned dst = ring_addr + 4u * (unsigned)(xs * 3 * st);
    if (fence_mode == 0) {                      // the stores themselves complete full[xs] of the CTA below
        const unsigned mb = full_addr + 8u * xs;
        for (int w = 2 * lane; w < st; w += 64) {
            st_remote_async_v2(dst + 4 * w, mb, below, *reinterpret_cast<const int2*>(srcH + w));
            st_remote_async_v2(dst + 4 * (st + w), mb, below, *reinterpret_cast<const int2*>(srcG + w));
            if (srcG2) st_remote_async_v2(dst + 4 * (2 * st + w), mb, below, *reinterpret_cast<const int2*>(srcG2 + w));
        }
        return;
    }
    for (int w = 2 * lane; w < st; w += 64) {
        st_remote_v2(dst + 4 * w, below, *reinterpret_cast<const int2*>(srcH + w));
        st_remote_v2(dst + 4 * (st + w), below, *reinterpret_cast<const int2*>(srcG + w));
        if (srcG2) st_remote_v2(dst + 4 * (2 * st + w), below, *reinterpret_cast<const int2*>(srcG2 + w));
    }
    __syncwarp();
    if (lane == 0) mbar_remote_arrive(full_addr + 8u * xs, below);
}
__device__ __noinline__ void cl_wait_above(unsigned full_addr, int S, bool fenced)
{
    mbar_wait(full_addr + 8u * (unsigned)((S - 1) % XD), (unsigned)((S - 1) / XD) & 1u, fenced);
}
__device__ __noinline__ void cl_release(unsigned empty_addr, unsigned full_addr, int S, int fence_mode, unsigned tx)
{
    const unsigned slot = (unsigned)((S - 2) % XD);
    if (fence_mode == 0) {
        mbar_arm_tx(full_addr + 8u * slot, tx);         // the next use of the slot: one arrival (this one) + the bytes of a push
        mbar_remote_arrive_relaxed(empty_addr + 8u * slot, (unsigned)(cl_rank() - 1));
    } else mbar_remote_arrive(empty_addr + 8u * slot, (unsigned)(cl_rank() - 1));
}

// TG threads per alignment: the whole CTA for few pairs (latency), down to one warp per alignment for
// large batches (no block barrier at all, 95 % of the lane-steps inside the matrix instead of 72 %).
// SPLIT (latency mode, TG = 256 only): three threads per row -- the diagonal, vertical and horizontal
// candidates of a cell are computed side by side by threads of three warp groups ("roles"), a second
// barrier, then role 0 applies the selection.  Same records, same results, ~1/3 of the per-step chain.
// MODE (record type) is a template parameter: a batch is launched once per mode present, so every
// `p.mode ==` test inside the per-cell code folds away (p.mode is overwritten with the constant below).
// SM: all record operands in shared memory (see k3_sm_ok; the host picks the variant per launch).
// CL (latency mode only: SPLIT, SM): a THREAD-BLOCK CLUSTER of NC CTAs works on one alignment -- CTA c holds the
// rows c*256 .. c*256+255 (+ k*NC*256), so NC*256 rows are in flight and a 1,100-row alignment needs
// LQ + LS steps instead of five stripes one after the other.  The first thread of CTA c > 0 reads what the last
// thread of CTA c-1 published one and two steps ago through distributed shared memory (copied to a local
// slot, so the cell code still sees shared-memory records); every step ends in a cluster barrier; the path
// records of CTA c live in its own part of the store (id = c * vmf_cap + local id).
// RL (register lists, k3r_core.cuh): 0 = the list-walking cell of k3_core.cuh; 4 / 6 / 8 = words per dynamic list of the
// branch-free register form (record modes 1 and 2, role-split shared-memory kernels, TGRL rows per CTA).
// SWG: the Smith-Waterman form (forwardC without secondary colonies, k3s_* in k3_core.cuh) -- a compile-time switch
// so that the banded kernels carry none of it; it runs the one-thread-per-row geometry.
template <int TG, bool SPLIT, int MODE, bool SM, bool CL, int RL, bool SWG = false>
__global__ void __launch_bounds__(SPLIT ? 3 * TG : CTA, SPLIT ? 1 : 2) k3_fill_kernel(const K3Args a)
{
    static_assert(!SWG || (!SPLIT && !CL && !RL && MODE != 3), "Smith-Waterman: one thread per row, single CTA, list-walking cell");
    static_assert(!CL || (SPLIT && SM), "the cluster variant is the role-split shared-memory kernel");
    static_assert(!SPLIT || CL || TG == CTA || RL, "role-split without clusters runs 3 x 256 threads");
    static_assert(!RL || (SPLIT && SM && (MODE == 1 || MODE == 2) && TG == TGRL), "register lists: modes 1 / 2, role-split, shared memory");
    constexpr int RCAP = RL ? RL : 4, RCS = RCAP - 2, RMODE = MODE == 2 ? 2 : 1;      // (valid template arguments also when RL == 0)
    namespace cg = cooperative_groups;
    constexpr int NG = SPLIT ? 1 : CTA / TG;        // alignments in flight per CTA
#define NC (CL ? cl_size() : 1)                     /* CTAs per alignment */
#define crank (CL ? cl_rank() : 0)
#define TGC (NC * TG)                               /* rows in flight per alignment */
    extern __shared__ __align__(16) int sm_dyn[];
    __shared__ int sm_pair[NG];
    __shared__ int sm_vmf[NG];
    __shared__ int sm_last_ptr[NG];
    __shared__ double sm_last_val[NG];
    const int role = SPLIT ? threadIdx.x / TG : 0;  // 0 diagonal (+ selection), 1 vertical, 2 horizontal
    const int g = SPLIT ? 0 : threadIdx.x / TG;
    const int t = threadIdx.x - (SPLIT ? role : g) * TG;
    const int slot = CL ? (int)blockIdx.x / NC : blockIdx.x * NG + g;
    int* const arena = a.arena + (size_t)slot * a.arena_words;
    K3Vmf* const vmf = a.vmf + (size_t)slot * (CL ? NC : 1) * a.vmf_cap;       // CL: NC consecutive parts
    int* const sm_grp = sm_dyn + (size_t)g * (a.smem_bytes / 4 / NG);
    const size_t sm_grp_words = (size_t)(a.smem_bytes / 4 / NG);

#define GSYNC() do { if (CL) cg::this_cluster().sync(); else if (SPLIT) __syncthreads(); else group_sync<TG>(g); } while (0)
#define GSTEP() do { if (SPLIT) __syncthreads(); else group_sync<TG>(g); } while (0)       /* end of one step */
#define RESET(ptr) do { if (RL) k3r_reset<RCAP, RMODE>(ptr); else if (SWG) k3s_blank(p, ptr, K3_NEVSEL); else k3_reset(p, ptr); } while (0)
#define RCOPY(d, s) do { if (RL) k3r_copy_words(d, s, st); else if (SWG) k3s_copy(p, d, s); else k3_copy(p, d, s); } while (0)
    // CL: ring hand-over (see above): full[slot] counts the push of the CTA above, empty[slot] the release by the CTA below
    __shared__ __align__(8) unsigned long long sm_full[XD], sm_empty[XD];
    for (;;) {
        if (t == 0 && role == 0 && crank == 0) sm_pair[g] = atomicAdd(a.counter, 1);
        GSYNC();
        const int pi = CL ? *cg::this_cluster().map_shared_rank(&sm_pair[0], 0) : sm_pair[g];
        if (CL) {
            if (threadIdx.x == 0) {             // fresh barriers per alignment (every CTA runs the same number of steps)
                for (int i = 0; i < XD; ++i) { mbar_init(smem_addr(&sm_full[i]), 1); mbar_init(smem_addr(&sm_empty[i]), 1); }
                asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
            }
            cg::this_cluster().sync();          // nobody leaves (or rewrites sm_pair) before every CTA has read it
        }
        if (pi >= a.npairs) break;
        const K3Pair& P_ = a.pairs[pi];
        const K3Group A = P_.a;
        const K3Group B = P_.b;
        K3Prm p = P_.prm;
        p.mode = MODE;
        // the per-column data of both groups sit in the staged blob (global memory; never null, pg_groups.cu
        // dev_side): telling the compiler turns the generic loads through these struct members into LDG
#define K3_GLOBAL(ptr) __builtin_assume(__isGlobal(ptr))
        K3_GLOBAL(A.cfq); K3_GLOBAL(A.efq); K3_GLOBAL(A.prof); K3_GLOBAL(A.freq); K3_GLOBAL(A.glen); K3_GLOBAL(A.gfreq);
        K3_GLOBAL(A.sfq); K3_GLOBAL(A.tfq); K3_GLOBAL(A.rfq); K3_GLOBAL(A.gapmask); K3_GLOBAL(A.weight);
        K3_GLOBAL(B.cfq); K3_GLOBAL(B.efq); K3_GLOBAL(B.prof); K3_GLOBAL(B.freq); K3_GLOBAL(B.glen); K3_GLOBAL(B.gfreq);
        K3_GLOBAL(B.sfq); K3_GLOBAL(B.tfq); K3_GLOBAL(B.rfq); K3_GLOBAL(B.gapmask); K3_GLOBAL(B.weight);
#undef K3_GLOBAL
        const int LQ = A.L, LS = B.L;
        const int st = k3_stride(p.capa, p.capb);
        constexpr int BW = k3r_block_words(RCS);              // words per static column block (register-list form)
        if (RL) { __builtin_assume(__isGlobal(A.blk)); __builtin_assume(__isGlobal(B.blk)); }
        const bool n3 = p.Noll == 3;
        // parked rows / boundary column: global arena (L2).  rowH[k] = H(pbase-1, k-1), colH[k] = H(k-1, -1)
        int* const rowH = arena;
        int* const rowG = rowH + (size_t)(LS + 2) * st;
        int* const rowG2 = rowG + (size_t)(LS + 2) * st;
        int* const colH = rowG2 + (size_t)(LS + 2) * st;
        // mode 3 with relaxed trailing gaps (lastB_ng): lastC[m+1] = H(m, LS-1), lastC[0] = boundary above the last
        // column; lastR[n+1] = H(LQ-1, n), lastR[0] = boundary left of the last row
        int* const lastC = colH + (size_t)(LQ + 2) * st;
        int* const lastR = lastC + (size_t)(LQ + 2) * st;
        int* const gwave = lastR + (size_t)(LS + 2) * st;
        int* const ghdr = gwave;                            // CL: {last ptr, overflow flag, last val (double)} of the cluster
        const bool want_last = MODE == 3 && (p.last_c || p.last_r);
        // wavefront records: shared memory when they fit, else the arena
        const bool in_smem = SM || k3_smem_words(st, p.Noll, TG) <= sm_grp_words;
        int* const wave = SM ? sm_grp : (in_smem ? sm_grp : gwave);
        int* const ringH = wave;                            // [RING] prefetched rowH records
        int* const ringG = ringH + (size_t)RING * st;
        int* const ringG2 = ringG + (size_t)RING * st;
        int* const black = ringG2 + (size_t)RING * st;
        int* const colbuf = black + st;                     // [2][2] boundary-column records of the row that starts (SM)
        int* const xring = colbuf + (size_t)4 * st;         // [XD][3] H, G, G2 of the last row of the CTA above (CL)
        int* const pubH = xring + (size_t)3 * XD * st;      // [3][TG]
        int* const pubG = pubH + (size_t)3 * TG * st;       // [2][TG]
        int* const F1 = pubG + (size_t)2 * TG * st;         // [TG]
        int* const pubG2 = F1 + (size_t)TG * st;            // [2][TG]   (two-piece only)
        int* const F2 = pubG2 + (size_t)2 * TG * st;        // [TG]
        // ---- reset the records this pair can read before writing
        if (role == 0) {
        if (crank == 0) {
        for (int i = t; i < LS + 2; i += TG) {
            RESET(rowH + (size_t)i * st); RESET(rowG + (size_t)i * st);
            if (n3) RESET(rowG2 + (size_t)i * st);
        }
        for (int i = t; i < LQ + 2; i += TG) RESET(colH + (size_t)i * st);
        if (CL && t == 0) { ghdr[0] = 0; ghdr[1] = 0; }
        }
        for (int k = 0; k < 3; ++k) RESET(pubH + ((size_t)k * TG + t) * st);
        for (int k = 0; k < 2; ++k) { RESET(pubG + ((size_t)k * TG + t) * st); if (n3) RESET(pubG2 + ((size_t)k * TG + t) * st); }
        if (t < RING) { RESET(ringH + (size_t)t * st); RESET(ringG + (size_t)t * st); RESET(ringG2 + (size_t)t * st); }
        if (t == 0) RESET(black);
        }
        GSYNC();
        // ---- initB (fwd2c.h:138-176): origin, then the two boundary chains (one thread each)
        if (CL && crank != 0 && t == 0 && role == 0) sm_vmf[g] = 0;
        if (t == 0 && role == 0 && crank == 0) {
            vmf[0].m = 0; vmf[0].n = 0; vmf[0].p = 0;                       // skip 0-th record (:361)
            vmf[1].m = P_.al; vmf[1].n = P_.bl; vmf[1].p = 0;                 // origin
            sm_vmf[g] = 2;
            k3_setval(colH, 0); k3_setdg(colH, p.mode == 3 ? K3_NEWD : ((p.rect || SWG) ? 0 : K3_DIAG), 0); K3_PTR(colH) = p.novmf ? p.origin_r : 1;
            if (SWG) { int* bx = k3s_box(p, colH); bx[0] = bx[1] = p.origin_r; bx[2] = P_.al; bx[3] = P_.bl; }
            const int rr = LQ < -p.lw ? LQ : -p.lw;
            for (int k = 1; k <= rr; ++k) {
                if (SWG) {        // initC (fwd2c.h:196-206): blank records on the diagonals below the origin
                    int* h = colH + (size_t)k * st; k3s_blank(p, h, 0);
                    int* bx = k3s_box(p, h); bx[0] = bx[1] = p.origin_r - k; bx[2] = P_.al + k; bx[3] = P_.bl;
                } else if (p.mode == 3) k3_boundary_b1(p, k, colH + (size_t)k * st, colH + (size_t)(k - 1) * st, false);
                else if (RL) k3r_boundary_col<RCAP, RCS, RMODE>(p, A.blk + (size_t)k * BW, B.blk, k, colH + (size_t)k * st, colH + (size_t)(k - 1) * st);
                else k3_boundary_col(p, A, B, k, colH + (size_t)k * st, colH + (size_t)(k - 1) * st);
            }
        }
        if (t == TG / 2 && role == (SPLIT ? 1 : 0) && crank == 0) {
            k3_setval(rowH, 0); k3_setdg(rowH, p.mode == 3 ? K3_NEWD : ((p.rect || SWG) ? 0 : K3_DIAG), 0); K3_PTR(rowH) = p.novmf ? p.origin_r : 1;
            if (SWG) { int* bx = k3s_box(p, rowH); bx[0] = bx[1] = p.origin_r; bx[2] = P_.al; bx[3] = P_.bl; }
            const int rr = LS < p.up ? LS : p.up;
            for (int k = 1; k <= rr; ++k) {
                if (SWG) {        // initC (:186-194): ... and above it
                    int* h = rowH + (size_t)k * st; k3s_blank(p, h, 0);
                    int* bx = k3s_box(p, h); bx[0] = bx[1] = p.origin_r + k; bx[2] = P_.al; bx[3] = P_.bl + k;
                } else if (p.mode == 3) k3_boundary_b1(p, k, rowH + (size_t)k * st, rowH + (size_t)(k - 1) * st, true);
                else if (RL) k3r_boundary_row<RCAP, RCS, RMODE>(p, A.blk, B.blk + (size_t)k * BW, k, rowH + (size_t)k * st, rowH + (size_t)(k - 1) * st);
                else k3_boundary_row(p, A, B, k, rowH + (size_t)k * st, rowH + (size_t)(k - 1) * st);
            }
        }
        GSYNC();
        if (want_last && t == 0 && role == 0 && crank == 0) {
            RCOPY(lastC, rowH + (size_t)LS * st);                   // black unless the band reaches it
            RCOPY(lastR, colH + (size_t)LQ * st);
        }

        // prefetch of the parked row: word w of the three records (rowH, rowG, rowG2) of one column
        const int pf_words = (n3 ? 3 : 2) * st;
        const bool ring_ok = SM || pf_words <= PFN * TG;

        // ---- continuous stripes: thread t takes rows t, t+TG, t+2TG, ...; its k-th row meets column n at
        //      global step S = k*P + t + n with the period P = max(LS, TG + 4): a thread starts its next row the
        //      step after it finished the previous one, so the wavefront never drains between stripes.  Ring
        //      slot of the parked record (stripe k, index i) = (k*P + i) % RING: thread 0 reads slots S, S+1
        //      at step S while slot S+3 is being filled.
        {
            const int PMIN = TGC + 4 + (CL ? 4 + NC * XD : 0);     // CL: CTA 0 may lead the last CTA by (NC - 1) * (XD - 1) steps
            const int P = LS > PMIN ? LS : PMIN;
            const int npass = (LQ + TGC - 1) / TGC;
            const int rows_last = LQ - (npass - 1) * TGC;
            const int total_steps = (npass - 1) * P + (rows_last - 1) + LS;
            int* const f1 = F1 + (size_t)t * st;
            int* const f2 = F2 + (size_t)t * st;
            double pua = 0;
            K3Best best = {0.0, 0, 0, 0, 0, 0, 0};         // Smith-Waterman: this thread's candidate for colony 0
            double diag_v = 0;
            // ring: records 1 and 2 of the first parked row (the boundary row) before the first step
            if (ring_ok && role == 0 && crank == 0) {
                for (int w = t; w < pf_words; w += TG) {
                    const int pa = w / st, pw = w - pa * st;
                    const int* src = pa == 0 ? rowH : (pa == 1 ? rowG : rowG2);
                    int* dst = pa == 0 ? ringH : (pa == 1 ? ringG : ringG2);
                    dst[(size_t)(1 % RING) * st + pw] = __ldcg(src + st + pw);
                    if (LS + 1 >= 2) dst[(size_t)(2 % RING) * st + pw] = __ldcg(src + 2 * st + pw);
                }
            }
            GSYNC();
            // this thread's share of the ring prefetch: word pf_pw[q] of parked array pf_src[q]
            const int* pf_src[PFN];
            int* pf_dst[PFN];
            bool pf_on[PFN];
#pragma unroll
            for (int q = 0; q < PFN; ++q) {
                const int w = t + q * TG;
                pf_on[q] = ring_ok && w < pf_words && role == (SPLIT ? 2 : 0) && crank == 0;
                const int pa = w / st, pw = w - pa * st;
                pf_src[q] = (pa == 0 ? rowH : (pa == 1 ? rowG : rowG2)) + pw;
                pf_dst[q] = (pa == 0 ? ringH : (pa == 1 ? ringG : ringG2)) + pw;
            }
            // this thread's position: q = S - t = k*P + n
            const int gt = crank * TG + t;                  // row slot of this thread inside the alignment
            int k = 0, n = -gt, m = gt;
            // CL: the same records in the previous CTA of the cluster (its last thread is the row above)
            // thread 0's position two steps ahead (what the ring must hold by then)
            int k2 = 0, n2 = 2;
            if (n2 >= P) { n2 -= P; ++k2; }
#ifdef K3_TIMERS
            long long tm_wait = 0, tm_c1 = 0, tm_b1 = 0, tm_c2 = 0, tm_b2 = 0, tm_push = 0, tm_cmb = 0, tm_vmf = 0, tm_pre = 0, tm_t;
#define TM_START() (tm_t = clock64())
#define TM_ADD(acc) do { const long long now_ = clock64(); acc += now_ - tm_t; tm_t = now_; } while (0)
#else
#define TM_START() do {} while (0)
#define TM_ADD(acc) do {} while (0)
#endif
            // CL, data-completed hand-over: this CTA arms every slot of its ring for its first use (later uses are armed
            // when the slot is given back); a push that arrives earlier just leaves the count negative until then
            if (CL && crank > 0 && t == 0 && role == 0 && a.cluster_fence == 0)
                for (int i = 0; i < XD; ++i) mbar_arm_tx(smem_addr(&sm_full[i]), cl_push_bytes(st, n3));
            for (int S = 0; S < total_steps; ++S) {
                TM_START();
                // prefetch the parked record thread 0 reads as "above" at step S + 2: index n2 + 1 of stripe k2 - 1
                int pf_val[PFN];
                const bool pf_now = ring_ok && n2 < LS && k2 * TGC < LQ;
                if (pf_now) {
#pragma unroll
                    for (int q = 0; q < PFN; ++q)
                        if (pf_on[q]) pf_val[q] = __ldcg(pf_src[q] + (size_t)(n2 + 1) * st);
                }
                const int r = n - m;
                const bool active = n >= 0 && n < LS && m < LQ && r >= p.lw && r <= p.up;
                bool rec = false;
                int* hout = pubH + ((size_t)(S % 3) * TG + t) * st;
                int* gout = pubG + ((size_t)(S & 1) * TG + t) * st;
                int* g2out = pubG2 + ((size_t)(S & 1) * TG + t) * st;
                if (CL && crank > 0 && t == 0 && role <= 1 && S >= 1)     // the row above has finished step S - 1 (pushed into our ring)
                    cl_wait_above(smem_addr(sm_full), S, a.cluster_fence == 1);
                TM_ADD(tm_wait);
                if (active) {
                    const int ia = m + 1, ib = n + 1;
                    const bool fr = m == 0 && !p.rect, fc = n == 0 && !p.rect;      // initA: no first-row / -column skips
                    if (n == 0 || r == p.lw) {                      // first in-band column of this row
                        if (!SPLIT || role == 1) pua = k3_unp(A, ia, B, ib, p.u);   // once per row (:377)
                        if (!SPLIT || role == 2) { RESET(f1); if (n3) RESET(f2); }
                    }
                    const int g3a = (S + 2) % 3, g3d = (S + 1) % 3, g2a = (S + 1) & 1;
                    const int seq = k * P + n;                      // == S for thread 0
                    const int* parkedH0 = ring_ok ? ringH + (size_t)(seq % RING) * st : rowH + (size_t)n * st;
                    const int* parkedH1 = ring_ok ? ringH + (size_t)((seq + 1) % RING) * st : rowH + (size_t)(n + 1) * st;
                    const int* parkedG1 = ring_ok ? ringG + (size_t)((seq + 1) % RING) * st : rowG + (size_t)(n + 1) * st;
                    const int* parkedG21 = ring_ok ? ringG2 + (size_t)((seq + 1) % RING) * st : rowG2 + (size_t)(n + 1) * st;
                    if (CL && crank > 0) {          // the row above lives in the CTA above: its records of steps S - 2 and S - 1
                        const int sd = (S + XD - 2) % XD, sa = (S + XD - 1) % XD;
                        parkedH0 = xring + (size_t)(sd * 3) * st; parkedH1 = xring + (size_t)(sa * 3) * st;
                        parkedG1 = xring + (size_t)(sa * 3 + 1) * st; parkedG21 = xring + (size_t)(sa * 3 + 2) * st;
                    }
                    // boundary column: H(m-1, -1) and H(m, -1).  SM: one row starts per step; its two boundary records
                    // come to shared memory first, so that every operand of the cell is a shared-memory record
                    int* const cb = colbuf + (size_t)(S & 1) * 2 * st;
                    const int* const colD = SM ? cb : colH + (size_t)m * st;
                    const int* const colL = SM ? cb + st : colH + (size_t)(m + 1) * st;
                    if (SM && n == 0) {
                        const int* gD = colH + (size_t)m * st;
                        const int* gL = colH + (size_t)(m + 1) * st;
                        if (!SPLIT || role == 0) for (int w = 0; w < st; w += 2) *reinterpret_cast<int2*>(cb + w) = __ldcg(reinterpret_cast<const int2*>(gD + w));
                        if (!SPLIT || role == 2) for (int w = 0; w < st; w += 2) *reinterpret_cast<int2*>(cb + st + w) = __ldcg(reinterpret_cast<const int2*>(gL + w));
                    }
                    const int* hdiag = n == 0 ? colD : (t == 0 ? parkedH0 : pubH + ((size_t)g3d * TG + (t - 1)) * st);
                    const bool above_in = r + 1 <= p.up;
                    const int* habove = !above_in ? black : (t == 0 ? parkedH1 : pubH + ((size_t)g3a * TG + (t - 1)) * st);
                    const int* gabove = (!above_in || m == 0) ? black : (t == 0 ? parkedG1 : pubG + ((size_t)g2a * TG + (t - 1)) * st);
                    const int* g2above = (!above_in || m == 0) ? black : (t == 0 ? parkedG21 : pubG2 + ((size_t)g2a * TG + (t - 1)) * st);
                    const bool left_in = r - 1 >= p.lw;
                    const int* hleft = n == 0 ? colL : (left_in ? pubH + ((size_t)g3a * TG + t) * st : black);
                    if (!SPLIT) {
                        const double dab = P_.simmat ? __ldg(P_.simmat + (size_t)m * LS + n) : k3_sim(A, B, p, ia, ib);
                        if (SWG) {
                            const int ra = r + p.origin_r;
                            diag_v = k3_val(hdiag);
                            k3s_part_diag(p, A, B, ia, ib, ra, dab, hdiag, hout);
                            k3s_part_vert(p, A, B, ia, ib, ra, m == 0, habove, gabove, g2above, gout, g2out);
                            k3s_part_hori(p, A, B, ia, ib, ra, n == 0, hleft, f1, f2);
                            k3s_combine(p, m == 0, n == 0, m + P_.al, n + P_.bl, diag_v, hout, gout, g2out, f1, f2, &best);
                        } else
                        rec = p.mode == 3
                            ? k3_cell_b1(p, dab, hdiag, habove, gabove, g2above, hleft, f1, f2, hout, gout, g2out)
                            : k3_cell_mono(p, A, B, ia, ib, fr, fc, dab, &pua, hdiag, habove, gabove, g2above, hleft, f1, f2,
                                           hout, gout, g2out, black);
                    } else if (p.mode == 3) {
                        if (role == 0) {
                            const double dab = P_.simmat ? __ldg(P_.simmat + (size_t)m * LS + n) : k3_sim(A, B, p, ia, ib);
                            rec = k3_cell_b1(p, dab, hdiag, habove, gabove, g2above, hleft, f1, f2, hout, gout, g2out);
                        }
                    } else if (RL) {
                        const int* const ablk = A.blk + (size_t)ia * BW;
                        const int* const bblk = B.blk + (size_t)ib * BW;
                        if (role == 0) {
                            const double dab = P_.simmat ? __ldg(P_.simmat + (size_t)m * LS + n) : k3_sim(A, B, p, ia, ib);
                            k3r_part_diag<RCAP, RCS, RMODE>(p, ablk, bblk, dab, hdiag, hout);
                        } else if (role == 1) {
                            k3r_part_vert<RCAP, RCS, RMODE>(p, ablk, bblk, A.nils != 0, m == 0, &pua, habove, gabove, g2above, gout, g2out, black, st);
                        } else {
                            k3r_part_hori<RCAP, RCS, RMODE>(p, ablk, bblk, n == 0, hleft, f1, f2);
                        }
                    } else if (SWG) {
                        const int ra = r + p.origin_r;
                        if (role == 0) {
                            const double dab = P_.simmat ? __ldg(P_.simmat + (size_t)m * LS + n) : k3_sim(A, B, p, ia, ib);
                            diag_v = k3_val(hdiag);
                            k3s_part_diag(p, A, B, ia, ib, ra, dab, hdiag, hout);
                        } else if (role == 1) k3s_part_vert(p, A, B, ia, ib, ra, m == 0, habove, gabove, g2above, gout, g2out);
                        else k3s_part_hori(p, A, B, ia, ib, ra, n == 0, hleft, f1, f2);
                    } else if (role == 0) {
                        const double dab = P_.simmat ? __ldg(P_.simmat + (size_t)m * LS + n) : k3_sim(A, B, p, ia, ib);
                        k3_part_diag(p, A, B, ia, ib, dab, hdiag, hout);
                    } else if (role == 1) {
                        k3_part_vert(p, A, B, ia, ib, fr, &pua, habove, gabove, g2above, gout, g2out, black);
                    } else {
                        k3_part_hori(p, A, B, ia, ib, fc, hleft, f1, f2);
                    }
                }
                TM_ADD(tm_c1);
                if (SPLIT) {
                    __syncthreads();                                // the three candidates are in shared memory
                    TM_ADD(tm_b1);
                    if (active && role == 0 && SWG)
                        k3s_combine(p, m == 0, n == 0, m + P_.al, n + P_.bl, diag_v, hout, gout, g2out, f1, f2, &best);
                    else if (active && role == 0 && p.mode != 3)
                        rec = RL ? k3r_combine(p, m == 0, n == 0, hout, gout, g2out, f1, f2, st) : k3_combine(p, m == 0 && !p.rect, n == 0 && !p.rect, hout, gout, g2out, f1, f2);
                    TM_ADD(tm_cmb);
                }
                if (active && role == 0) {
                    if (p.novmf) {                                  // HomScoreC: no Vmf; first-row cells remember their diagonal
                        if (m == 0) K3_PTR(hout) = n + p.origin_r;
                    } else if (rec) {
                        int id = atomicAdd(&sm_vmf[g], 1);          // Vmf::add (fwd2c.h:465-467)
                        const bool fits = id < a.vmf_cap;
                        if (CL) { if (!fits) ghdr[1] = 1; id += crank * a.vmf_cap; }
                        if (fits) { vmf[id].m = m + P_.al; vmf[id].n = n + P_.bl; vmf[id].p = K3_PTR(hout); }
                        K3_PTR(hout) = id;
                    }
                    if (want_last) {
                        if (n == LS - 1) RCOPY(lastC + (size_t)(m + 1) * st, hout);
                        if (m == LQ - 1) RCOPY(lastR + (size_t)(n + 1) * st, hout);
                    }
                    if (m == LQ - 1) {
                        if (n == LS - 1) {
                            if (CL) { ghdr[0] = K3_PTR(hout); *reinterpret_cast<double*>(ghdr + 2) = k3_val(hout); }
                            else { sm_last_ptr[g] = K3_PTR(hout); sm_last_val[g] = k3_val(hout); }
                        }
                    } else if (t == TG - 1 && crank == NC - 1) {    // bottom row of a stripe: park it
                        RCOPY(rowH + (size_t)(n + 1) * st, hout);
                        RCOPY(rowG + (size_t)(n + 1) * st, gout);
                        if (n3) RCOPY(rowG2 + (size_t)(n + 1) * st, g2out);
                    }
                }
                TM_ADD(tm_vmf);
                if (pf_now) {
                    const int slot = (k2 * P + n2 + 1) % RING;
#pragma unroll
                    for (int q = 0; q < PFN; ++q)
                        if (pf_on[q]) pf_dst[q][(size_t)slot * st] = pf_val[q];
                }
                // advance the two positions
                if (++n == P) { n = 0; ++k; m += TGC; }
                if (++n2 == P) { n2 = 0; ++k2; }
                TM_ADD(tm_c2);
                if (CL && role == 0) {
                    if (crank < NC - 1 && t >= TG - 32)
                        cl_push(pubH + ((size_t)(S % 3) * TG + (TG - 1)) * st, pubG + ((size_t)(S & 1) * TG + (TG - 1)) * st,
                                n3 ? pubG2 + ((size_t)(S & 1) * TG + (TG - 1)) * st : nullptr, smem_addr(xring), smem_addr(sm_full),
                                smem_addr(sm_empty), S, st, t - (TG - 32), a.cluster_fence);
                    // the slot of step S - 2 has now served as "above" (S - 1) and "diagonal" (S): give it back
                    if (crank > 0 && t == 0 && S >= 2) cl_release(smem_addr(sm_empty), smem_addr(sm_full), S, a.cluster_fence, cl_push_bytes(st, n3));
                }
                TM_ADD(tm_push);
                GSTEP();
                TM_ADD(tm_b2);
            }
#ifdef K3_TIMERS
            if (t == 0 || t == TG - 1)
                printf("k3 timers role %d pair %d cta %d/%d t %d steps %d LQ %d LS %d us: wait %.0f c1 %.0f b1 %.0f combine %.0f records %.0f rest %.0f push %.0f b2 %.0f\n", role, pi, crank, NC, t,
                       total_steps, LQ, LS, tm_wait / 1965., tm_c1 / 1965., tm_b1 / 1965., tm_cmb / 1965., tm_vmf / 1965., tm_c2 / 1965., tm_push / 1965., tm_b2 / 1965.);
#endif
            if (CL) cg::this_cluster().sync();      // the last cell, the path parts and the parked rows of every CTA
            if (SWG) {        // colony 0 (Colonies::at(0)): the first cell in row-major order that holds the maximum
                K3Best* const sb = reinterpret_cast<K3Best*>(pubH);         // the wavefront records are done with
                GSYNC();
                if (role == 0) sb[t] = best;
                GSYNC();
                if (t == 0 && role == 0 && crank == 0) {
                    K3Best c0 = sb[0];
                    for (int i = 1; i < TG; ++i) if (k3s_better(sb[i], c0)) c0 = sb[i];
                    int* out = a.out_pts + 2 * P_.out_off;
                    out[0] = c0.mlb; out[1] = c0.nlb; out[2] = c0.mrb; out[3] = c0.nrb; out[4] = c0.lwr; out[5] = c0.upr;
                    a.out_cnt[pi] = 3;
                    a.out_score[pi] = c0.val;
                }
            }
        }
        // ---- Aln2b1::lastB_ng (fwd2b1.cc:100-143): trailing gaps at true sequence ends cost rtgapf times the
        //      penalty: the last column is relaxed downwards, then the last row rightwards, in place
        if (CL && t == 0 && role == 0 && crank == 0) {            // the last cell was written by some CTA of the cluster
            sm_last_ptr[g] = __ldcg(ghdr);
            sm_last_val[g] = __ldcg(reinterpret_cast<const double*>(ghdr + 2));
        }
        if (want_last && t == 0 && role == 0 && crank == 0) {
            __threadfence();
            int dm = 0, dn = 0;
            if (p.last_c) {
                const int rw = p.up < LS ? p.up : LS;
                for (int m = LS - rw; m <= LQ - 1; ++m) {
                    int* gq = lastC + (size_t)m * st;
                    int* hq = lastC + (size_t)(m + 1) * st;
                    ++dm;
                    const double gpn = !k3_isvert(k3_dir(gq)) ? ((1 > p.codonk1) ? p.gop2 + p.gep2 : p.gop1 + p.gep1)
                                                              : (dm > p.codonk1 ? p.gep2 : p.gep1);
                    k3_setval(gq, k3_val(gq) + gpn * p.rtg_b);
                    if (k3_val(gq) > k3_val(hq)) { k3_copy(p, hq, gq); k3_setdg(hq, K3_VERT, 0); } else dm = 0;
                }
                k3_copy(p, lastR + (size_t)LS * st, lastC + (size_t)LQ * st);       // the corner cell is shared
            }
            if (p.last_r) {
                const int rw = p.lw > -LQ ? p.lw : -LQ;
                for (int n = rw + LQ; n <= LS - 1; ++n) {
                    int* fq = lastR + (size_t)n * st;
                    int* hq = lastR + (size_t)(n + 1) * st;
                    ++dn;
                    const double gpn = !k3_ishori(k3_dir(fq)) ? ((1 > p.codonk1) ? p.gop2 + p.gep2 : p.gop1 + p.gep1)
                                                              : (dn > p.codonk1 ? p.gep2 : p.gep1);
                    k3_setval(fq, k3_val(fq) + gpn * p.rtg_a);
                    if (k3_val(fq) > k3_val(hq)) { k3_copy(p, hq, fq); k3_setdg(hq, K3_VERT, 0); } else dn = 0;
                }
            }
            int* const h9 = lastR + (size_t)LS * st;
            int ptr = K3_PTR(h9);
            if (dn || dm) {
                if (dn) dm = 0;
                const int id = sm_vmf[g]++;
                if (id < a.vmf_cap) { vmf[id].m = LQ - dm + P_.al; vmf[id].n = LS - dn + P_.bl; vmf[id].p = ptr; }
                ptr = id;
            }
            sm_last_ptr[g] = ptr;
            sm_last_val[g] = k3_val(h9);
        }
        // ---- final record + Vmf::traceback (fwd2c.h:475-481, vmf.cc:103-119)
        if (t == 0 && role == 0 && crank == 0 && !SWG) {
            int* out = a.out_pts + 2 * P_.out_off;
            int cnt = 0;
            const int nrec = sm_vmf[g];
            if (p.novmf) out[0] = sm_last_ptr[g];                             // pp[0] of forwardB (fwd2c.h:476-479); no corners
            else if (nrec >= a.vmf_cap || (CL && __ldcg(ghdr + 1))) cnt = -1;   // record store overflow: reported, never silent
            else {
                out[0] = LQ + P_.al; out[1] = LS + P_.bl; cnt = 1;
                for (int q = sm_last_ptr[g];; q = vmf[q].p) {
                    if (cnt >= P_.out_cap) { cnt = -1; break; }
                    out[2 * cnt] = vmf[q].m; out[2 * cnt + 1] = vmf[q].n; ++cnt;
                    if (!vmf[q].p) break;
                }
            }
            a.out_cnt[pi] = cnt;
            a.out_score[pi] = sm_last_val[g];
        }
        GSYNC();
    }
}

#undef NC
#undef crank
#undef TGC

template <int TG, bool SPLIT, int MODE, bool SM>
cudaError_t launch_tgm(const K3Args& a, int grid_blocks, cudaStream_t st)
{
    cudaError_t e = cudaFuncSetAttribute(k3_fill_kernel<TG, SPLIT, MODE, SM, false, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, a.smem_bytes);
    if (e != cudaSuccess) return e;
    // The gap-profile lists and score rows stream through L1.  Few pairs (latency: Prrn::best_of_n sized
    // batches): one CTA per SM and the rest of the unified array as L1 (measured 28 vs 37 ms for 24
    // pairs); more CTAs than SMs (throughput): two CTAs per SM win (59 vs 68 ms for 384 pairs).
    // PG_K3_CARVEOUT = percent of shared memory overrides.
    int carve = (!SPLIT && grid_blocks > 148) ? 100 : (int)((a.smem_bytes + 2048) * 100LL / (228 * 1024)) + 1;
    if (const char* cv = getenv("PG_K3_CARVEOUT")) carve = atoi(cv);
    if (carve > 100) carve = 100;
    e = cudaFuncSetAttribute(k3_fill_kernel<TG, SPLIT, MODE, SM, false, 0>, cudaFuncAttributePreferredSharedMemoryCarveout, carve);
    if (e != cudaSuccess) return e;
    k3_fill_kernel<TG, SPLIT, MODE, SM, false, 0><<<grid_blocks, SPLIT ? 3 * TG : CTA, a.smem_bytes, st>>>(a);
    return cudaGetLastError();
}

// cluster variant: `clusters` alignments in flight, a.cluster CTAs each
template <int MODE>
cudaError_t launch_cluster(const K3Args& a, int clusters, cudaStream_t st)
{
    auto kern = k3_fill_kernel<TGCL, true, MODE, true, true, 0>;
    K3Args ac = a;
    { const char* f = getenv("PG_K3_CLUSTER_FENCE"); ac.cluster_fence = f && f[0] >= '0' && f[0] <= '2' ? f[0] - '0' : 0; }
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, a.smem_bytes);
    if (e != cudaSuccess) return e;
    int carve = (int)((a.smem_bytes + 2048) * 100LL / (228 * 1024)) + 1;
    if (const char* cv = getenv("PG_K3_CARVEOUT")) carve = atoi(cv);
    if (carve > 100) carve = 100;
    e = cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, carve);
    if (e != cudaSuccess) return e;
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = dim3((unsigned)(clusters * a.cluster));
    cfg.blockDim = dim3(3 * TGCL);
    cfg.dynamicSmemBytes = (size_t)a.smem_bytes;
    cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = (unsigned)a.cluster; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, kern, ac);
}

// register-list form: TGRL rows x 3 roles per CTA, two CTAs per SM; `units` CTAs, or clusters of a.cluster CTAs
template <int MODE, int RL>
cudaError_t launch_rl(const K3Args& a, int units, cudaStream_t st)
{
    const bool cl = a.cluster > 1;
    K3Args ac = a;
    { const char* f = getenv("PG_K3_CLUSTER_FENCE"); ac.cluster_fence = f && f[0] >= '0' && f[0] <= '2' ? f[0] - '0' : 0; }
    const void* kern = cl ? (const void*)k3_fill_kernel<TGRL, true, MODE, true, true, RL>
                          : (const void*)k3_fill_kernel<TGRL, true, MODE, true, false, RL>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, a.smem_bytes);
    if (e != cudaSuccess) return e;
    int carve = (int)(2 * (a.smem_bytes + 2048) * 100LL / (228 * 1024)) + 1;       // room for two CTAs per SM
    if (const char* cv = getenv("PG_K3_CARVEOUT")) carve = atoi(cv);
    if (carve > 100) carve = 100;
    e = cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, carve);
    if (e != cudaSuccess) return e;
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = dim3((unsigned)(units * (cl ? a.cluster : 1)));
    cfg.blockDim = dim3(3 * TGRL);
    cfg.dynamicSmemBytes = (size_t)a.smem_bytes;
    cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = (unsigned)(cl ? a.cluster : 1); at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at; cfg.numAttrs = cl ? 1 : 0;
    void* args[1] = {(void*)&ac};
    return cudaLaunchKernelExC(&cfg, kern, args);
}
template <int MODE>
cudaError_t launch_rl_cap(const K3Args& a, int units, cudaStream_t st)
{
    switch (a.rl) {
    case 4: return launch_rl<MODE, 4>(a, units, st);
    case 6: return launch_rl<MODE, 6>(a, units, st);
    default: return launch_rl<MODE, 8>(a, units, st);
    }
}

template <int TG, bool SPLIT, bool SM>
cudaError_t launch_tg(const K3Args& a, int mode, int grid_blocks, cudaStream_t st)
{
    if (SPLIT && SM && (a.cluster > 1 || a.rows192)) {
        cudaError_t e;
        switch (mode) {
        case 0: e = launch_cluster<0>(a, grid_blocks, st); break;
        case 1: e = launch_cluster<1>(a, grid_blocks, st); break;
        case 2: e = launch_cluster<2>(a, grid_blocks, st); break;
        case 3: e = launch_cluster<3>(a, grid_blocks, st); break;
        default: e = launch_cluster<4>(a, grid_blocks, st); break;
        }
        if (e == cudaSuccess) return e;
        (void)cudaGetLastError();       // a cluster of this size cannot be co-scheduled here: the single-CTA kernel below
        if (a.rows192) return launch_tg<256, false, false>(a, mode, grid_blocks, st);   // (256 rows do not fit shared memory)
    }
    switch (mode) {
    case 0: return launch_tgm<TG, SPLIT, 0, SM>(a, grid_blocks, st);
    case 1: return launch_tgm<TG, SPLIT, 1, SM>(a, grid_blocks, st);
    case 2: return launch_tgm<TG, SPLIT, 2, SM>(a, grid_blocks, st);
    case 3: return launch_tgm<TG, SPLIT, 3, SM>(a, grid_blocks, st);
    default: return launch_tgm<TG, SPLIT, 4, SM>(a, grid_blocks, st);
    }
}

}  // namespace

int k3_threads() { return CTA; }
int k3_cluster_rows() { return TGCL; }
int k3_rl_rows() { return TGRL; }
int k3_blocks_per_sm() { return 2; }
size_t k3_wave_words(int stride, int Noll, int tg) { return k3_smem_words(stride, Noll, tg); }

// Threads per alignment for a batch (PG_K3_TG overrides).  Measured on B200, partitions of a 200 x ~500 family:
// the role-split kernel (768 = 3 threads per row, one CTA per SM, the rest of the unified array as L1 for the
// gap-profile lists) wins at every batch size -- 24 pairs 8.7 ms, 384 pairs 30.4 ms against 46.4 ms (one thread
// per row, 256 rows, two CTAs per SM, no L1 left), 1,536 pairs 111 ms against 147 ms (128 rows per alignment).
// The one-thread-per-row variants stay for PG_K3_TG = 128 / 256 and for records too long for shared memory.
int k3_pick_tg(int64_t npairs, int sm_count)
{
    (void)npairs; (void)sm_count;
    if (const char* e = getenv("PG_K3_TG")) {
        const int v = atoi(e);
        if (v == 128 || v == 256 || v == 768) return v;
    }
    return 768;
}

// a.smem_bytes: dynamic shared memory per CTA = (CTA / tg) x the largest wavefront of the batch (capped)
// One launch processes the pairs of ONE mode (a.pairs / a.npairs / a.counter / a.out_* are that mode's slice).
cudaError_t k3_launch(const K3Args& a, int tg, int mode, int grid_blocks, cudaStream_t st)
{
    // a.all_sm: every pair of the launch keeps its wavefront records and prefetch ring in shared memory for this
    // tg (k3_all_sm): the variant whose cell operands are all shared-memory records.  Otherwise (very long
    // gap-state lists) the generic-address variant, one thread per row, 256 rows per stripe.
    if (a.swg) {           // Smith-Waterman pairs: their own instantiations of the generic one-thread-per-row kernel
        auto go = [&](auto kern) -> cudaError_t {
            cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, a.smem_bytes);
            if (e != cudaSuccess) return e;
            kern<<<grid_blocks, CTA, a.smem_bytes, st>>>(a);
            return cudaGetLastError();
        };
        switch (mode) {
        case 0: return go(k3_fill_kernel<256, false, 0, false, false, 0, true>);
        case 1: return go(k3_fill_kernel<256, false, 1, false, false, 0, true>);
        case 2: return go(k3_fill_kernel<256, false, 2, false, false, 0, true>);
        case 4: return go(k3_fill_kernel<256, false, 4, false, false, 0, true>);
        default: return cudaErrorInvalidValue;
        }
    }
    if (a.rl) return mode == 1 ? launch_rl_cap<1>(a, grid_blocks, st) : launch_rl_cap<2>(a, grid_blocks, st);
    if (!a.all_sm) return launch_tg<256, false, false>(a, mode, grid_blocks, st);
    if (getenv("PG_K3_GENERIC") && tg != 768 && tg != 128) return launch_tg<256, false, false>(a, mode, grid_blocks, st);
    switch (tg) {
    case 128: return launch_tg<128, false, true>(a, mode, grid_blocks, st);
    case 768: return launch_tg<256, true, true>(a, mode, grid_blocks, st);      // role-split latency kernel
    default: return launch_tg<256, false, true>(a, mode, grid_blocks, st);
    }
}

bool k3_sm_fits_rows(int stride, int Noll, int rows, size_t smem_bytes) { return k3_sm_ok(stride, Noll, rows, smem_bytes / 4); }

bool k3_sm_fits(int stride, int Noll, int tg, size_t smem_bytes)
{
    const int ngrp = tg == 768 ? 1 : K3_THREADS / tg;
    return k3_sm_ok(stride, Noll, tg == 768 ? 256 : tg, smem_bytes / 4 / ngrp);
}
