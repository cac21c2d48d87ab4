// pg_internal.h -- shared declarations between the C ABI (pg_api.cu) and the kernels.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <string>
#include <vector>

#include "../../include/prrn_gpu.h"

// ---- device view of a sequence set ---------------------------------------------------------------
struct PgDevSeqs {
    const uint8_t* res;     // concatenated residue codes
    const int64_t* offs;    // [nseq]
    const int32_t* left;    // [nseq] window start (absolute within the sequence)
    const int32_t* wlen;    // [nseq] window length  right - left
    const uint8_t* flags;   // [nseq] bit0 exgl, bit1 exgr, bit2 left != 0, bit3 right != len
    int32_t nseq;
};

// epilogue selector of the score kernel
enum PgEpilogue {
    PG_EPI_SCORE_F32 = 0,   // VTYPE float score        (alnScoreD, aln build)
    PG_EPI_SCORE_F64 = 1,   // VTYPE double score       (alnScoreD, prrn build)
    PG_EPI_DIST_F32 = 2,    // 100*(1 - (scr + u|dl|/2)/sqrt(self_a self_b)) in float  (dpscore)
    PG_EPI_DIST_F64 = 3,    // same in double
};

// One work item = one query (rows) against a run of subjects (columns); one CTA builds the query
// profile once and its warps take the subjects.
struct PgItem {
    int32_t q;          // sequence index of the rows
    int32_t sub_begin;  // implicit mode: first subject sequence index; explicit: index into pair_s[]
    int32_t sub_end;
    int32_t pad;
};

struct K1Args {
    PgDevSeqs seqs;
    const PgItem* items;
    int32_t nitems;
    int32_t* counter;           // dynamic item counter (zeroed before launch)
    // explicit pair mode (pg_score_pairs): subject index and output slot per sorted pair
    const int32_t* pair_s;      // nullptr => implicit all-vs-all: subject = sub index, out = elem(q,s)-k_begin
    const int64_t* pair_out;
    const uint8_t* pair_swap;   // explicit mode: 1 if (rows, cols) = (b, a) of the caller's pair
    int64_t k_begin, k_end;     // implicit mode: only k in [k_begin, k_end) is written
    const int32_t* mtx;         // dim x dim integer scores
    int32_t dim;
    int32_t u, v;               // integer uu, vv
    int32_t sh;
    int32_t tgapf_zero;         // 1 when tgapf == 0 (free terminal gaps at true sequence ends)
    float u_f32;                // alprm.u as float for the distance epilogue
    const int32_t* self;        // per-sequence self scores (dist epilogues)
    int32_t epilogue;
    void* out;
    int2* rowbuf;               // multi-pass scratch: [grid warps][rowbuf_stride]
    int64_t rowbuf_stride;
    int32_t rows_per_lane;      // k1_rows_per_pass(window lengths) / 32: which instantiation runs (0 = 16)
    int32_t pad_rows;
};

// packed (int16 x 2) score kernel: a pair of queries against a run of subjects
struct PgItem2 {
    int32_t q0, q1;
    int32_t sub_begin, sub_end;     // range in subs[]
    int32_t rows;                   // rows per lane of the kernel variant (8, 10, 12, 14 or 16)
    int32_t pad[3];
};

struct K1PArgs {
    PgDevSeqs seqs;
    const PgItem2* items;
    int32_t nitems;
    int32_t* counter;
    const uint32_t* subs;       // subject index | half-0 valid << 30 | half-1 valid << 31
    int64_t k_begin, k_end;
    const int32_t* mtx;
    int32_t dim;
    int32_t u, v, sh;
    float u_f32;
    const int32_t* self;
    int32_t epilogue;
    void* out;
    uint2* rowbuf;
    int64_t rowbuf_stride;
};

// floating-point / all-modes score kernel (k1f_score.cu); orientation rows = a, columns = b
struct K1FArgs {
    PgDevSeqs seqs;
    const PgItem* items;        // q = a (rows); implicit mode: subjects are sequence indices j > q
    int32_t nitems;
    int32_t* counter;
    const int32_t* pair_s;      // explicit mode: subject (b) per sorted pair; nullptr => implicit
    const int64_t* pair_out;    // explicit mode: output slot per sorted pair
    int64_t k_begin, k_end;     // implicit mode: slot = elem(q, j) - k_begin, written if inside the range
    const void* mtx;            // VTYPE[dim * dim], mtx[a][b]
    int32_t dim;
    const void* bnd;            // VTYPE[3][bnd_stride] boundary tables (k1f_build_tables)
    int32_t bnd_stride;
    double uu, vv;              // (VTYPE)(alprm.u * scale), (VTYPE)(alprm.v * scale)  (fwd2d1.cc:62-63)
    float tgapf;
    int32_t sh;
    int32_t mode;               // 0 forwardD, 1 forwardD + lastD, 2 swgforwardD, 3 Fwd2d_vd (ends)
    int32_t vtype;              // 0 float, 1 double
    int32_t epilogue;           // 0 score, 1 distance (global branch), 2 distance (lcl branch; mode 3)
    float u_f32;
    const void* self;           // VTYPE[nseq] self scores (epilogue 1)
    void* out;                  // VTYPE / FTYPE per slot
    int32_t* out_ends;          // mode 3: ends[2] per slot (nullable)
    void* scratch;              // per-warp scratch: multi-pass rows, lastD lines
    int64_t scratch_stride;     // bytes per warp
    int32_t off_rowv, off_rowr, off_col, off_row, off_colr, off_rowr2, off_misc;
    int32_t exg_override;       // -1: inex flags from the sequences; else bits 0-1 = a's exgl/exgr, 2-3 = b's
    int32_t rows_per_lane;      // k1f_rows_per_pass(vtype, mode, window lengths) / 32: which instantiation runs
};

struct K2Rec;

// arguments of the alignment-with-path kernels (k2_align.cu); pairs are sorted by query
struct K2Args {
    PgDevSeqs seqs;
    const PgItem* items;
    int32_t nitems;
    int32_t* counter;
    const int32_t* pair_q;      // [npairs] rows sequence (a)
    const int32_t* pair_s;      // [npairs] columns sequence (b)
    const int64_t* dir_off;     // [npairs] offset (in 64-bit words) of the pair's direction words
    const int64_t* len_off;     // [npairs] prefix sum of (LQ + LS + 8): scratch / output offsets
    unsigned long long* dirs;   // direction bits, wavefront order (k2_core.cuh)
    const int32_t* mtx;
    int32_t dim;
    int32_t u, v, sh;
    int32_t* score;             // [npairs] integer DP score (sorted order)
    int2* rowbuf;
    int64_t rowbuf_stride;
    unsigned char* moves;       // trace scratch
    K2Rec* recs;                // trace scratch
    int32_t* out_pts;           // corner lists, 2 ints per corner, Vmf back-walk order
    int32_t* out_cnt;           // [npairs]
    // striped long-pair kernel (k2_fill_long_kernel): one pair, one warp per 512-row stripe
    int32_t* progress;          // [npass] columns whose bottom row a stripe has published
    int32_t* ticket;            // stripe dispenser (scheduling order = dependency order)
    int32_t rows_per_lane;      // rows per lane of the fill that wrote `dirs` (0 = k2_rows_per_lane(); the striped kernel: 4 / 8 / 16)
};

// group-to-group alignment kernel (k3_groups.cu)
#include "k3_core.cuh"
#define K3_THREADS 256
struct K3Pair {
    K3Group a, b;               // device pointers
    K3Prm prm;
    const double* simmat;       // optional precomputed sim2 matrix [LQ][LS] (kernel K4), else nullptr
    int32_t al, bl;             // a.left, b.left (absolute coordinates of the corner list)
    int64_t out_off;            // corner output offset (in corners)
    int32_t out_cap;
    int32_t pad;
};
struct K3Args {
    const K3Pair* pairs;
    int32_t npairs;
    int32_t* counter;
    int* arena;                 // per-CTA record arena
    int64_t arena_words;
    K3Vmf* vmf;                 // per-CTA path record store
    int32_t vmf_cap;
    int32_t* out_pts;           // corner lists (2 ints per corner), Vmf back-walk order
    int32_t* out_cnt;           // [npairs] corners, -1 on overflow
    double* out_score;          // [npairs]
    int32_t smem_bytes;         // dynamic shared memory per CTA for the wavefront records
    int32_t all_sm;             // every pair fits the all-shared-memory variant (k3_sm_fits)
    int32_t cluster;            // latency mode: CTAs (thread-block cluster size) per alignment, 1 = none
    int32_t cluster_fence;      // cluster variant, per-step hand-over (PG_K3_CLUSTER_FENCE): 0 st.async records complete the barrier (default), 1 release arrives + cluster-scope acquires, 2 release arrives + CTA-scope waits
    int32_t rl;                 // register-list form (k3r_core.cuh): words per dynamic list (4 / 6 / 8) of this launch, 0 = classic
    int32_t rows192;            // the records fit shared memory only with 192 rows per CTA: the cluster kernel's geometry,
                                // also for a "cluster" of one CTA (long gap-state lists: high hetero, two-piece)
    int32_t swg;                // Smith-Waterman launch (forwardC): the SWG instantiation, results = colony 0
    int32_t pad_swg;
};

// profile contraction (k4_contract.cu): S = X_a . Y_b^T per pair, written to K3Pair::simmat
struct K4Args {
    const K3Pair* pairs;
    int32_t npairs;
    const int32_t* block_off;   // [npairs + 1] prefix sums of ceil(LQ / 8)
};

struct pg_dev_seqs {
    PgDevSeqs v;
    void* blob;                 // single device allocation holding everything
    std::vector<int32_t> h_wlen;    // host copies needed to build work items
    int32_t max_wlen;
    int32_t max_code;           // largest residue code inside the windows
    uint8_t present[256];       // which residue codes occur inside the windows
    int32_t min_wlen;
    int64_t res_bytes;              // bytes of concatenated residues
    bool any_exg;               // some sequence carries inex.exgl / exgr
    // cached packed plan of the last calcdist range (schedule only, no results)
    int64_t plan_k0, plan_k1;
    void* d_plan;               // PgItem2[nitems] | subs[nsubs]
    int32_t plan_nitems;
    int64_t plan_nsubs;
    bool plan_multipass;
    bool owns;                  // false: blob / plan live in the context's reusable workspace
    size_t plan_cap;
};

struct pg_context {
    int device;
    int sm_count;
    cudaStream_t stream;
    std::string err;
    // reusable device workspace
    void* d_items; size_t items_cap;
    void* d_mtx; size_t mtx_cap;
    void* d_self; size_t self_cap;
    void* d_rowbuf; size_t rowbuf_cap;
    void* d_out; size_t out_cap;
    void* d_pairs; size_t pairs_cap;
    void* d_dirs; size_t dirs_cap;
    void* d_seqblob; size_t seqblob_cap;
    void* d_planbuf; size_t planbuf_cap;
    void* d_trace; size_t trace_cap;
    void* d_bnd; size_t bnd_cap;
    void* d_scratch; size_t scratch_cap;
    void* d_ends; size_t ends_cap;
    void* d_gblob; size_t gblob_cap;
    void* h_gstage; size_t gstage_cap;      // pinned host staging of the group blob (grow-only)
    void* d_garena; size_t garena_cap;
    void* d_gvmf; size_t gvmf_cap;
    void* d_gout; size_t gout_cap;
    void* d_gsim; size_t gsim_cap;
    int32_t* d_counter;
    cudaStream_t aux[5];        // K3 launches one kernel per record mode: they run side by side
    cudaEvent_t ev_fork, ev_join[5];
    cudaEvent_t ev0, ev1;       // device time of the last fill launch (pg_last_kernel_ms)
    bool ev_valid;
    // The workspace above is shared by every call on this context.  Calls are ordered on c->stream, except
    // pg_calcdist_dev on a caller's stream: whoever touches the workspace next first waits (on the device) for the
    // last stream that used it (pg_int_order_stream).
    cudaStream_t ws_stream;     // stream of the last call that used the workspace
    cudaEvent_t ev_ws;
};

int pg_int_fail(pg_context* ctx, int code, const char* msg);
// Make `st` the stream that owns the context's workspace: if another stream used it last, `st` waits for the work
// queued there.  Called at the top of every entry point before the workspace is written.
cudaError_t pg_int_order_stream(pg_context* c, cudaStream_t st);
int pg_int_ensure_cap(pg_context* c, void** p, size_t* cap, size_t need);

int pg_int_align_pairs_fp(pg_context* c, const pg_seqs* s, const int32_t* a_idx, const int32_t* b_idx, int64_t npairs,
                          const pg_params* prm, const void* mtx, int32_t dim, void* out_scores, int64_t** out_offs,
                          pg_skl** out_pts, int b1);

// kernels (k1_score.cu)
cudaError_t k1_launch(const K1Args& a, int grid_blocks, cudaStream_t st);
int k1_rows_per_pass(const int32_t* wlen, int nseq);
int k1_warps_per_block();
int k1_blocks_per_sm();
cudaError_t k1_self_launch(const PgDevSeqs& s, const int32_t* mtx, int dim, int32_t* self, cudaStream_t st);
// k1p_score.cu
cudaError_t k1p_launch(const K1PArgs& a, int grid_blocks, cudaStream_t st);
int k1p_rows_per_pass();
int k1p_pick_rows(int lq);
int k1p_warps_per_block();
int k1p_blocks_per_sm();
// k1f_score.cu
cudaError_t k1f_launch(const K1FArgs& a, int sm_count, cudaStream_t st);
cudaError_t k1f_self_launch(const PgDevSeqs& s, const void* mtx, int dim, int vtype, void* self, cudaStream_t st);
int k1f_warps_per_block();
int k1f_rows_per_pass(int vtype, int mode, const int32_t* wlen, int nseq);
int k1f_grid_blocks(int sm_count, int vtype, int mode);
// k2_align.cu
cudaError_t k2_fill_launch(const K2Args& a, int grid_blocks, cudaStream_t st);
cudaError_t k2_trace_launch(const K2Args& a, int npairs, cudaStream_t st);
cudaError_t k2_fill_long_launch(const K2Args& a, int npass, int sm_count, size_t rowbuf_bytes, cudaStream_t st);
int k2_rows_per_lane();
int k2_long_rows(int LQ, int LS);     // rows per lane of the striped long-pair kernel
int k2_warps_per_block();
int k2_blocks_per_sm();
// k3_groups.cu
cudaError_t k3_launch(const K3Args& a, int tg, int mode, int grid_blocks, cudaStream_t st);
int k3_threads();
int k3_cluster_rows();        // rows per CTA of the cluster latency kernel
int k3_rl_rows();             // rows per CTA of the register-list kernels
int k3_blocks_per_sm();
int k3_pick_tg(int64_t npairs, int sm_count);
bool k3_sm_fits(int stride, int Noll, int tg, size_t smem_bytes);
bool k3_sm_fits_rows(int stride, int Noll, int rows, size_t smem_bytes);
size_t k3_wave_words(int stride, int Noll, int tg);
// k4_contract.cu
cudaError_t k4_launch(const K4Args& a, int total_blocks, cudaStream_t st);
// dpx_peak.cu
cudaError_t dpx_peak_run(int sm_count, cudaStream_t st, double* gops_s32, double* gops_s16x2);
