// pg_api.cu -- the extern "C" layer of libprrn_gpu.so (include/prrn_gpu.h): argument checks,
// staging of sequences / score matrix / work items into HBM, kernel launches.  No CPU fallback:
// every compute entry fails loudly when the device or a kernel is unavailable.
#include <cuda_runtime.h>
#include <math.h>
#include <stdio.h>
#include <string.h>
#include <time.h>

#include <algorithm>
#include <numeric>
#include <string>
#include <thread>
#include <vector>

#include "k1f_core.cuh"
#include "k1p_core.cuh"
#include "k2_core.cuh"
#include "pg_internal.h"

static std::string g_create_err;
static inline size_t up256(size_t x) { return (x + 255) & ~(size_t)255; }

#define PG_CUDA(ctx, call)                                                                      \
    do {                                                                                        \
        cudaError_t e__ = (call);                                                               \
        if (e__ != cudaSuccess) {                                                               \
            (ctx)->err = std::string(#call) + ": " + cudaGetErrorString(e__);                   \
            return PG_ERR_CUDA;                                                                 \
        }                                                                                       \
    } while (0)

static int fail(pg_context* ctx, int code, const std::string& msg)
{
    if (ctx) ctx->err = msg;
    else g_create_err = msg;
    return code;
}

extern "C" const char* pg_version(void) { return "prrn_aln_b200 0.1 (sm_100a)"; }

extern "C" const char* pg_last_error(const pg_context* ctx) { return ctx ? ctx->err.c_str() : g_create_err.c_str(); }

extern "C" int pg_create(int device, pg_context** out)
{
    if (!out) return fail(nullptr, PG_ERR_ARG, "pg_create: out is NULL");
    *out = nullptr;
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev <= 0)
        return fail(nullptr, PG_ERR_NO_DEVICE,
                    std::string("pg_create: no CUDA device (") + cudaGetErrorString(e) + "); there is no CPU fallback");
    if (device < 0 || device >= ndev) return fail(nullptr, PG_ERR_ARG, "pg_create: bad device index");
    if ((e = cudaSetDevice(device)) != cudaSuccess)
        return fail(nullptr, PG_ERR_CUDA, std::string("cudaSetDevice: ") + cudaGetErrorString(e));
    cudaDeviceProp prop;
    if ((e = cudaGetDeviceProperties(&prop, device)) != cudaSuccess)
        return fail(nullptr, PG_ERR_CUDA, std::string("cudaGetDeviceProperties: ") + cudaGetErrorString(e));
    if (prop.major < 10)
        return fail(nullptr, PG_ERR_NO_DEVICE, "pg_create: device is not sm_100-class; kernels are built for sm_100a only");
    pg_context* c = new pg_context();
    c->device = device;
    c->sm_count = prop.multiProcessorCount;
    c->d_items = c->d_mtx = c->d_self = c->d_rowbuf = c->d_out = c->d_pairs = c->d_dirs = c->d_trace = c->d_seqblob = c->d_planbuf = nullptr;
    c->dirs_cap = c->trace_cap = c->seqblob_cap = c->planbuf_cap = 0;
    c->d_bnd = c->d_scratch = c->d_ends = nullptr;
    c->bnd_cap = c->scratch_cap = c->ends_cap = 0;
    c->d_gblob = c->d_garena = c->d_gvmf = c->d_gout = c->d_gsim = nullptr;
    c->gblob_cap = c->garena_cap = c->gvmf_cap = c->gout_cap = c->gsim_cap = 0;
    c->h_gstage = nullptr; c->gstage_cap = 0;
    c->items_cap = c->mtx_cap = c->self_cap = c->rowbuf_cap = c->out_cap = c->pairs_cap = 0;
    c->d_counter = nullptr;
    c->ev_valid = false;
    if ((e = cudaEventCreate(&c->ev0)) != cudaSuccess || (e = cudaEventCreate(&c->ev1)) != cudaSuccess ||
        (e = cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking)) != cudaSuccess ||
        (e = cudaMalloc(&c->d_counter, 8 * sizeof(int32_t))) != cudaSuccess) {     // [0] work queue; [1..5] K3 per-mode queues
        std::string m = std::string("pg_create: ") + cudaGetErrorString(e);
        delete c;
        return fail(nullptr, PG_ERR_CUDA, m);
    }
    for (int i = 0; i < 5 && e == cudaSuccess; ++i) {
        e = cudaStreamCreateWithFlags(&c->aux[i], cudaStreamNonBlocking);
        if (e == cudaSuccess) e = cudaEventCreateWithFlags(&c->ev_join[i], cudaEventDisableTiming);
    }
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&c->ev_fork, cudaEventDisableTiming);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&c->ev_ws, cudaEventDisableTiming);
    c->ws_stream = c->stream;
    if (e != cudaSuccess) {
        std::string m = std::string("pg_create: ") + cudaGetErrorString(e);
        return fail(nullptr, PG_ERR_CUDA, m);
    }
    *out = c;
    return PG_OK;
}

extern "C" void pg_destroy(pg_context* c)
{
    if (!c) return;
    cudaSetDevice(c->device);
    cudaStreamSynchronize(c->stream);
    cudaFree(c->d_items); cudaFree(c->d_mtx); cudaFree(c->d_self); cudaFree(c->d_rowbuf);
    cudaFree(c->d_out); cudaFree(c->d_pairs); cudaFree(c->d_counter); cudaFree(c->d_dirs); cudaFree(c->d_trace); cudaFree(c->d_seqblob); cudaFree(c->d_planbuf);
    cudaFree(c->d_bnd); cudaFree(c->d_scratch); cudaFree(c->d_ends);
    if (c->h_gstage) cudaFreeHost(c->h_gstage);
    cudaFree(c->d_gblob); cudaFree(c->d_garena); cudaFree(c->d_gvmf); cudaFree(c->d_gout); cudaFree(c->d_gsim);
    if (c->ws_stream != c->stream) cudaStreamSynchronize(c->ws_stream);     // a caller's stream may still use the workspace
    cudaEventDestroy(c->ev0); cudaEventDestroy(c->ev1); cudaEventDestroy(c->ev_fork); cudaEventDestroy(c->ev_ws);
    for (int i = 0; i < 5; ++i) { cudaStreamDestroy(c->aux[i]); cudaEventDestroy(c->ev_join[i]); }
    cudaStreamDestroy(c->stream);
    delete c;
}

static int ensure_cap(pg_context* c, void** p, size_t* cap, size_t need)
{
    if (need <= *cap && *p) return PG_OK;
    if (*p) { PG_CUDA(c, cudaFree(*p)); *p = nullptr; *cap = 0; }
    size_t n = need + need / 4 + 256;
    PG_CUDA(c, cudaMalloc(p, n));
    *cap = n;
    return PG_OK;
}

cudaError_t pg_int_order_stream(pg_context* c, cudaStream_t st)
{
    if (c->ws_stream == st) return cudaSuccess;
    cudaError_t e = cudaEventRecord(c->ev_ws, c->ws_stream);
    if (e == cudaSuccess) e = cudaStreamWaitEvent(st, c->ev_ws, 0);
    if (e == cudaSuccess) c->ws_stream = st;
    return e;
}

// internal linkage helpers for the other translation units of the library (pg_groups.cu)
int pg_int_fail(pg_context* ctx, int code, const char* msg) { return fail(ctx, code, msg); }
int pg_int_ensure_cap(pg_context* c, void** p, size_t* cap, size_t need) { return ensure_cap(c, p, cap, need); }

// ---- parameter / matrix validation for the integer kernels -------------------------------------
struct IntScoring {
    int u, v;
    std::vector<int32_t> mtx;
};

static bool is_integral(double x) { return x == rint(x) && fabs(x) < (1 << 20); }

static int make_int_scoring(pg_context* c, const pg_params* prm, const void* mtx, int dim, const uint8_t* present,
                            IntScoring* out)
{
    if (!prm || !mtx) return fail(c, PG_ERR_ARG, "params / mtx is NULL");
    if (dim < 1 || dim > 32) return fail(c, PG_ERR_ARG, "dim must be in [1, 32]");
    // fwd2d1.cc:62-63: uu = (VTYPE)(alprm.u * alprm.scale) -- float product
    float uu = prm->alprm.u * prm->alprm.scale, vv = prm->alprm.v * prm->alprm.scale;
    if (!is_integral(uu) || !is_integral(vv) || uu < 0 || vv < 0)
        return fail(c, PG_ERR_UNSUPPORTED, "non-integral gap penalties: floating-point fill is not built yet (no CPU fallback)");
    out->u = (int)uu;
    out->v = (int)vv;
    out->mtx.resize((size_t)dim * dim);
    for (int i = 0; i < dim * dim; ++i) {
        // Entries of residue codes that do not occur in the batch are never read by the fill.  The
        // reference leaves some of them uninitialised (e.g. mtx[SEC][UNP], mtx[SEC][SEC] after
        // Simmtx::Pmtx(fname), simmtx.cc:336-445), so they must not take part in validation.
        if (!present[i / dim] || !present[i % dim]) { out->mtx[i] = 0; continue; }
        double x = prm->vtype ? ((const double*)mtx)[i] : (double)((const float*)mtx)[i];
        if (!is_integral(x) || fabs(x) > 32000)
            return fail(c, PG_ERR_UNSUPPORTED, "non-integral substitution matrix: floating-point fill is not built yet (no CPU fallback)");
        out->mtx[i] = (int32_t)x;
    }
    if (prm->lcl != 0)
        return fail(c, PG_ERR_UNSUPPORTED, "algmode.lcl != 0 (semi-global / local score) is not built yet");
    if (!(prm->alprm.tgapf == 1.0f))
        return fail(c, PG_ERR_UNSUPPORTED, "tgapf != 1 (lastD terminal-gap relaxation, fwd2d1.cc:97-134) is not built yet");
    return PG_OK;
}

// ---- which kernel takes an alnScoreD batch ------------------------------------------------------
// The exact-integer DPX kernels (K1 / K1P) take the global score with integral matrix and penalties,
// tgapf == 1 and no free ends; everything else of alnScoreD's dispatch (fwd2d1.cc:324-337) goes to the
// floating-point kernel K1F in the caller's VTYPE.  mode: 0 forwardD, 1 forwardD + lastD,
// 2 swgforwardD, 3 Fwd2d_vd.  No path leaves the GPU.
struct ScorePlan {
    bool integer;       // K1 / K1P
    int mode;           // K1F mode otherwise
};

static int plan_scoring(pg_context* c, const pg_params* prm, const void* mtx, int dim, const pg_dev_seqs* d,
                        bool want_ends, IntScoring* sc, ScorePlan* plan)
{
    if (!prm || !mtx) return fail(c, PG_ERR_ARG, "params / mtx is NULL");
    if (dim < 1 || dim > 32) return fail(c, PG_ERR_ARG, "dim must be in [1, 32]");
    plan->integer = false;
    if (prm->lcl & 16) { plan->mode = 2; return PG_OK; }
    if (want_ends) { plan->mode = 3; return PG_OK; }
    const bool lastd = d->any_exg || !(prm->alprm.tgapf == 1.0f);
    plan->mode = lastd ? 1 : 0;
    if (!lastd) {
        const char* force = getenv("PG_FORCE_FLOAT");
        if (!(force && force[0] == '1')) {
            pg_params p0 = *prm;
            p0.lcl = 0;
            std::string keep = c->err;
            if (make_int_scoring(c, &p0, mtx, dim, d->present, sc) == PG_OK) plan->integer = true;
            else c->err = keep;     // not integral: K1F takes it
            // K1 / K1P pick rows and columns freely (rows = the longer list's query, tournament halves); that equals
            // the reference's mtx[a][b] (fwd2d1.cc:149) only for a matrix symmetric over the codes present.  K1F
            // keeps the orientation rows = a.
            for (int i = 0; plan->integer && i < dim; ++i)
                for (int j = 0; j < i; ++j)
                    if (d->present[i] && d->present[j] && sc->mtx[i * dim + j] != sc->mtx[j * dim + i]) { plan->integer = false; break; }
        }
    }
    return PG_OK;
}

// ---- K1F staging: matrix in VTYPE, boundary tables, per-warp scratch, launch ----------------------
static int k1f_run(pg_context* c, cudaStream_t st, pg_dev_seqs* d, const pg_params* prm, const void* mtx, int dim,
                   int mode, int epilogue, const std::vector<PgItem>& items, bool multipass, const int32_t* d_pair_s,
                   const int64_t* d_pair_out, int64_t k0, int64_t k1, int exg_override, void* d_out, int32_t* d_out_ends,
                   int32_t* n_launches)
{
    int rc;
    const int vt = prm->vtype ? 1 : 0;
    const size_t esz = vt ? sizeof(double) : sizeof(float);
    K1FArgs a;
    memset(&a, 0, sizeof(a));
    a.seqs = d->v;
    // fwd2d1.cc:62-63: uu = (VTYPE)(alprm.u * alprm.scale) -- float products
    const float fu = prm->alprm.u * prm->alprm.scale, fv = prm->alprm.v * prm->alprm.scale;
    a.uu = (double)fu;
    a.vv = (double)fv;
    a.tgapf = prm->alprm.tgapf;
    // score matrix as the caller's VTYPE
    if ((rc = ensure_cap(c, &c->d_mtx, &c->mtx_cap, esz * (size_t)dim * dim))) return rc;
    PG_CUDA(c, cudaMemcpyAsync(c->d_mtx, mtx, esz * (size_t)dim * dim, cudaMemcpyHostToDevice, st));
    // boundary tables (fwd2d1.cc:67-87), accumulated in VTYPE as the reference does
    const int stride = d->max_wlen + 2;
    std::vector<unsigned char> bnd(3 * (size_t)stride * esz);
    if (vt) k1f_build_tables<double>((double*)bnd.data(), stride, (double)fu, (double)fv, prm->alprm.tgapf);
    else k1f_build_tables<float>((float*)bnd.data(), stride, fu, fv, prm->alprm.tgapf);
    if ((rc = ensure_cap(c, &c->d_bnd, &c->bnd_cap, bnd.size()))) return rc;
    PG_CUDA(c, cudaMemcpyAsync(c->d_bnd, bnd.data(), bnd.size(), cudaMemcpyHostToDevice, st));
    if ((rc = ensure_cap(c, &c->d_items, &c->items_cap, sizeof(PgItem) * std::max<size_t>(items.size(), 1)))) return rc;
    if (!items.empty())
        PG_CUDA(c, cudaMemcpyAsync(c->d_items, items.data(), sizeof(PgItem) * items.size(), cudaMemcpyHostToDevice, st));
    // per-warp scratch
    const bool lines = mode == 1 || mode == 3, vd = mode == 3;
    auto al16 = [](size_t x) { return (x + 15) & ~(size_t)15; };
    const size_t n = (size_t)d->max_wlen + 4;
    size_t off = 0;
    a.off_rowv = (int32_t)off; off += multipass ? al16(2 * n * esz) : 0;
    a.off_rowr = (int32_t)off; off += (multipass && vd) ? al16(2 * n * 4) : 0;
    a.off_col = (int32_t)off; off += lines ? al16(n * esz) : 0;
    a.off_row = (int32_t)off; off += lines ? al16(n * esz) : 0;
    a.off_colr = (int32_t)off; off += vd ? al16(n * 4) : 0;
    a.off_rowr2 = (int32_t)off; off += vd ? al16(n * 4) : 0;
    a.off_misc = (int32_t)off; off += 16;
    if (off > 0x7fffffff) return fail(c, PG_ERR_RANGE, "sequence too long for the score kernel's scratch layout");
    const int grid = k1f_grid_blocks(c->sm_count, vt, mode);
    a.scratch_stride = (int64_t)al16(off);
    if ((rc = ensure_cap(c, &c->d_scratch, &c->scratch_cap, (size_t)a.scratch_stride * grid * k1f_warps_per_block()))) return rc;
    a.scratch = c->d_scratch;
    int launches = 1;
    if (epilogue == 1) {
        if ((rc = ensure_cap(c, &c->d_self, &c->self_cap, esz * std::max<size_t>(d->v.nseq, 1)))) return rc;
        PG_CUDA(c, k1f_self_launch(d->v, c->d_mtx, dim, vt, c->d_self, st));
        a.self = c->d_self;
        ++launches;
    }
    PG_CUDA(c, cudaMemsetAsync(c->d_counter, 0, sizeof(int32_t), st));
    a.items = (const PgItem*)c->d_items;
    a.nitems = (int32_t)items.size();
    a.counter = c->d_counter;
    a.pair_s = d_pair_s;
    a.pair_out = d_pair_out;
    a.k_begin = k0; a.k_end = k1;
    a.mtx = c->d_mtx; a.dim = dim;
    a.bnd = c->d_bnd; a.bnd_stride = stride;
    a.sh = prm->alprm.sh;
    a.mode = mode; a.vtype = vt; a.epilogue = epilogue;
    a.u_f32 = prm->alprm.u;
    a.out = d_out; a.out_ends = d_out_ends;
    a.exg_override = exg_override;
    a.rows_per_lane = k1f_rows_per_pass(vt, mode, d->h_wlen.data(), (int)d->h_wlen.size()) / 32;
    PG_CUDA(c, k1f_launch(a, c->sm_count, st));
    if (n_launches) *n_launches = launches;
    return PG_OK;
}

// ---- sequences --------------------------------------------------------------------------------
static int seqs_upload_impl(pg_context* c, const pg_seqs* s, pg_dev_seqs** out, bool in_ctx);

extern "C" int pg_seqs_upload(pg_context* c, const pg_seqs* s, pg_dev_seqs** out)
{
    return seqs_upload_impl(c, s, out, false);
}

// in_ctx: place the blob in the context's reusable workspace (host-buffer entry points call this
// once per call; a cudaMalloc/cudaFree pair per call costs milliseconds and occasional long stalls)
static int seqs_upload_impl(pg_context* c, const pg_seqs* s, pg_dev_seqs** out, bool in_ctx)
{
    if (!c) return PG_ERR_ARG;
    if (!s || !out || s->nseq < 0 || (s->nseq > 0 && (!s->res || !s->offs || !s->lens)))
        return fail(c, PG_ERR_ARG, "pg_seqs_upload: bad sequence set");
    PG_CUDA(c, cudaSetDevice(c->device));
    PG_CUDA(c, pg_int_order_stream(c, c->stream));
    const int n = s->nseq;
    int64_t total = 0;
    for (int i = 0; i < n; ++i) {
        if (s->lens[i] < 0 || s->offs[i] < 0) return fail(c, PG_ERR_ARG, "negative length / offset");
        total = std::max<int64_t>(total, s->offs[i] + s->lens[i]);
    }
    std::vector<int32_t> left(n), wlen(n);
    std::vector<uint8_t> flags(n);
    int32_t maxw = 0;
    for (int i = 0; i < n; ++i) {
        int l = s->left ? s->left[i] : 0, r = s->right ? s->right[i] : s->lens[i];
        if (l < 0 || r > s->lens[i] || l > r) return fail(c, PG_ERR_ARG, "window outside the sequence");
        uint8_t ex = s->exg ? s->exg[i] : 0;
        left[i] = l;
        wlen[i] = r - l;
        flags[i] = (uint8_t)((ex & 3) | (l ? 4 : 0) | (r != s->lens[i] ? 8 : 0));
        maxw = std::max(maxw, wlen[i]);
    }
    auto up16 = [](size_t x) { return (x + 255) & ~(size_t)255; };
    size_t o_res = 0, o_offs = up16(o_res + (size_t)total + 16), o_left = up16(o_offs + sizeof(int64_t) * n),
           o_wlen = up16(o_left + sizeof(int32_t) * n), o_flags = up16(o_wlen + sizeof(int32_t) * n),
           bytes = up16(o_flags + n + 16);
    pg_dev_seqs* d = new pg_dev_seqs();
    d->blob = nullptr;
    d->owns = !in_ctx;
    cudaError_t e = cudaSuccess;
    if (in_ctx) {
        int rc0 = ensure_cap(c, &c->d_seqblob, &c->seqblob_cap, bytes);
        if (rc0) { delete d; return rc0; }
        d->blob = c->d_seqblob;
    } else {
        e = cudaMalloc(&d->blob, bytes);
        if (e != cudaSuccess) { delete d; return fail(c, PG_ERR_CUDA, std::string("cudaMalloc(seqs): ") + cudaGetErrorString(e)); }
    }
    char* b = (char*)d->blob;
    e = cudaMemcpyAsync(b + o_res, s->res, (size_t)total, cudaMemcpyHostToDevice, c->stream);
    if (e == cudaSuccess && n) e = cudaMemcpyAsync(b + o_offs, s->offs, sizeof(int64_t) * n, cudaMemcpyHostToDevice, c->stream);
    if (e == cudaSuccess && n) e = cudaMemcpyAsync(b + o_left, left.data(), sizeof(int32_t) * n, cudaMemcpyHostToDevice, c->stream);
    if (e == cudaSuccess && n) e = cudaMemcpyAsync(b + o_wlen, wlen.data(), sizeof(int32_t) * n, cudaMemcpyHostToDevice, c->stream);
    if (e == cudaSuccess && n) e = cudaMemcpyAsync(b + o_flags, flags.data(), n, cudaMemcpyHostToDevice, c->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);   // host staging vectors die at return
    if (e != cudaSuccess) {
        if (d->owns) cudaFree(d->blob);
        delete d;
        return fail(c, PG_ERR_CUDA, std::string("upload seqs: ") + cudaGetErrorString(e));
    }
    d->v.res = (const uint8_t*)(b + o_res);
    d->v.offs = (const int64_t*)(b + o_offs);
    d->v.left = (const int32_t*)(b + o_left);
    d->v.wlen = (const int32_t*)(b + o_wlen);
    d->v.flags = (const uint8_t*)(b + o_flags);
    d->v.nseq = n;
    d->h_wlen = wlen;
    d->max_wlen = maxw;
    d->min_wlen = n ? *std::min_element(wlen.begin(), wlen.end()) : 0;
    d->res_bytes = (int64_t)total;
    d->any_exg = false;
    for (int i = 0; i < n; ++i) d->any_exg = d->any_exg || (flags[i] & 3);
    d->plan_k0 = d->plan_k1 = -1;
    d->d_plan = nullptr;
    d->plan_nitems = 0;
    d->plan_nsubs = 0;
    d->plan_multipass = false;
    d->plan_cap = 0;
    // residues must index the matrix: remember the largest code
    uint8_t mx = 0;
    memset(d->present, 0, sizeof(d->present));
    for (int i = 0; i < n; ++i) {
        const uint8_t* p = s->res + s->offs[i];
        for (int k = left[i]; k < left[i] + wlen[i]; ++k) { mx = std::max(mx, p[k]); d->present[p[k]] = 1; }
    }
    d->max_code = mx;
    *out = d;
    return PG_OK;
}

extern "C" void pg_seqs_free(pg_context* c, pg_dev_seqs* d)
{
    if (!d) return;
    if (c) cudaSetDevice(c->device);
    if (d->owns) {
        cudaFree(d->blob);
        cudaFree(d->d_plan);
    }
    delete d;
}

// ---- work items -------------------------------------------------------------------------------
static inline int64_t tri(int64_t j) { return j * (j - 1) / 2; }

static int row_of_k(int64_t k)
{   // largest j with j(j-1)/2 <= k
    int64_t j = (int64_t)((1.0 + sqrt(1.0 + 8.0 * (double)k)) / 2.0);
    while (tri(j) > k) --j;
    while (tri(j + 1) <= k) ++j;
    return (int)j;
}

static void build_calcdist_items(const pg_dev_seqs* d, int64_t k0, int64_t k1, int grid_blocks,
                                 std::vector<PgItem>* items, bool* multipass)
{
    items->clear();
    *multipass = false;
    if (k1 <= k0) return;
    const int NWv = k1_warps_per_block(), rpp = k1_rows_per_pass(d->h_wlen.data(), (int)d->h_wlen.size());
    const int jlo = row_of_k(k0), jhi = row_of_k(k1 - 1);
    int64_t total = k1 - k0;
    int64_t ch = (total + (int64_t)16 * grid_blocks - 1) / ((int64_t)16 * grid_blocks);
    ch = std::max<int64_t>(NWv, std::min<int64_t>(ch, 32 * NWv));
    ch = (ch + NWv - 1) / NWv * NWv;
    for (int j = jlo; j <= jhi; ++j) {
        int64_t i0 = j == jlo ? k0 - tri(j) : 0;
        int64_t i1 = j == jhi ? k1 - tri(j) : j;
        const bool mp = d->h_wlen[j] > rpp;
        if (mp) *multipass = true;
        const int64_t c = mp ? NWv : ch;
        for (int64_t i = i0; i < i1; i += c) {
            PgItem it;
            it.q = j;
            it.sub_begin = (int32_t)i;
            it.sub_end = (int32_t)std::min<int64_t>(i + c, i1);
            it.pad = 0;
            items->push_back(it);
        }
    }
    // longest-processing-time-first: heavy items early so the persistent CTAs finish together
    std::stable_sort(items->begin(), items->end(), [&](const PgItem& a, const PgItem& b) {
        int64_t ca = (int64_t)d->h_wlen[a.q] * (a.sub_end - a.sub_begin);
        int64_t cb = (int64_t)d->h_wlen[b.q] * (b.sub_end - b.sub_begin);
        return ca > cb;
    });
}

// K1F orientation: rows = a = the smaller index i, subjects j > i restricted to the rows [jlo, jhi] of
// the condensed range (partial first / last rows are trimmed by the kernel's slot test).
static void build_calcdist_items_rows_a(const pg_dev_seqs* d, int64_t k0, int64_t k1, int grid_blocks, int rpp,
                                        std::vector<PgItem>* items, bool* multipass)
{
    items->clear();
    *multipass = false;
    if (k1 <= k0) return;
    const int NWv = k1f_warps_per_block();
    const int jlo = row_of_k(k0), jhi = row_of_k(k1 - 1);
    int64_t ch = ((k1 - k0) + (int64_t)16 * grid_blocks - 1) / ((int64_t)16 * grid_blocks);
    ch = std::max<int64_t>(NWv, std::min<int64_t>(ch, 32 * NWv));
    ch = (ch + NWv - 1) / NWv * NWv;
    for (int i = 0; i < jhi; ++i) {
        const int j0 = std::max(i + 1, jlo), j1 = jhi + 1;
        const bool mp = d->h_wlen[i] > rpp;
        if (mp) *multipass = true;
        const int64_t c = mp ? NWv : ch;
        for (int64_t j = j0; j < j1; j += c) {
            PgItem it;
            it.q = i;
            it.sub_begin = (int32_t)j;
            it.sub_end = (int32_t)std::min<int64_t>(j + c, j1);
            it.pad = 0;
            items->push_back(it);
        }
    }
    std::stable_sort(items->begin(), items->end(), [&](const PgItem& a, const PgItem& b) {
        return (int64_t)d->h_wlen[a.q] * (a.sub_end - a.sub_begin) > (int64_t)d->h_wlen[b.q] * (b.sub_end - b.sub_begin);
    });
}

// ---- packed plan: query pairs of similar length x subjects, every pair {x, y} exactly once ---------
// Rows J = [jlo, jhi] of the condensed triangle are the queries of this range.  Subjects are
//   (a) the sequences below jlo: every (j in J, i < jlo) pair -- both halves valid;
//   (b) the other members of J, assigned by a round-robin tournament over J sorted by length:
//       {x, y} goes to x iff (pos(y) - pos(x)) mod |J| in [1, (|J|-1)/2] (antipodal ties to the
//       smaller position), so two queries adjacent in the sorted order share all but one subject.
// Partial first / last rows of the range are trimmed by the per-pair range test.
static void build_packed_plan(const pg_dev_seqs* d, int64_t k0, int64_t k1, int grid_blocks,
                              std::vector<PgItem2>* items, std::vector<uint32_t>* subs, bool* multipass)
{
    items->clear();
    subs->clear();
    *multipass = false;
    if (k1 <= k0) return;
    const int NWv = k1p_warps_per_block(), rpp = k1p_rows_per_pass();
    const int jlo = row_of_k(k0), jhi = row_of_k(k1 - 1);
    const int nJ = jhi - jlo + 1;
    std::vector<int> J(nJ);
    std::iota(J.begin(), J.end(), jlo);
    std::stable_sort(J.begin(), J.end(), [&](int x, int y) { return d->h_wlen[x] < d->h_wlen[y]; });
    // pair {x, y} is inside [k0, k1) iff min(x, y) lies in [rlo, rhi) of the row max(x, y) of the triangle
    std::vector<int32_t> rlo(jhi + 1), rhi(jhi + 1);
    for (int64_t j = 0; j <= jhi; ++j) {
        const int64_t t = j * (j - 1) / 2;
        rlo[j] = (int32_t)std::min<int64_t>(std::max<int64_t>(k0 - t, 0), j);
        rhi[j] = (int32_t)std::min<int64_t>(std::max<int64_t>(k1 - t, 0), j);
    }
    auto in_range = [&](int x, int y) {
        const int hi = x > y ? x : y, lo = x > y ? y : x;
        return lo >= rlo[hi] && lo < rhi[hi];
    };
    const int h = (nJ - 1) / 2;
    auto assigned = [&](int px, int py) {       // does the pair of positions (px, py) belong to px ?
        int dd = py - px;
        if (dd < 0) dd += nJ;
        if (dd >= 1 && dd <= h) return true;
        return (nJ % 2 == 0) && dd == nJ / 2 && px < nJ / 2;
    };
    int64_t total = k1 - k0;
    int64_t ch = (total / 2 + (int64_t)16 * grid_blocks - 1) / ((int64_t)16 * grid_blocks);
    ch = std::max<int64_t>(NWv, std::min<int64_t>(ch, 32 * NWv));
    ch = (ch + NWv - 1) / NWv * NWv;
    // subjects are visited longest first, so that every item holds subjects of similar length: the two halves of
    // a warp (and the systolic arrays of a CTA) then run the same number of steps
    const int nseq = (int)d->h_wlen.size();
    std::vector<int> by_len(jhi + 1), posJ(nseq, -1);
    std::iota(by_len.begin(), by_len.end(), 0);
    std::stable_sort(by_len.begin(), by_len.end(), [&](int x, int y) { return d->h_wlen[x] > d->h_wlen[y]; });
    for (int p = 0; p < nJ; ++p) posJ[J[p]] = p;
    // the whole triangle of the whole set (what calcdist asks for on one GPU): no range tests needed
    const bool full = k0 == 0 && k1 == tri((int64_t)nseq) && jhi == nseq - 1;
    std::vector<int> by_len_lo;                                  // sequences below jlo, longest first
    for (int x : by_len) if (x < jlo) by_len_lo.push_back(x);
    // the subject lists of the query pairs are independent: host threads build them side by side
    const int npq = (nJ + 1) / 2;
    std::vector<std::vector<uint32_t>> lists(npq);
    auto build_lists = [&](int p_begin, int p_end) {
        for (int pq = p_begin; pq < p_end; ++pq) {
            const int p = 2 * pq;
            const int qa = J[p], qb = p + 1 < nJ ? J[p + 1] : J[p];
            const bool has_b = p + 1 < nJ;
            std::vector<uint32_t>& list = lists[pq];
            list.reserve((size_t)jhi / 2 + 16);
            if (full) {
                // Whole triangle: every pair is in range, so the subjects of (p, p + 1) are known without scanning all
                // sequences -- the positions p + 1 .. p + h + 2 (mod |J|) of the tournament, longest first (J is sorted
                // by length: descending position, the wrapped positions last), and the sequences below jlo (sequence 0)
                // merged in by length.
                int r = 0;                                       // next rectangle-part subject (by_len_lo: longest first)
                auto emit_lo = [&](int upto_len) {               // rectangle subjects at least as long as upto_len
                    while (r < (int)by_len_lo.size() && d->h_wlen[by_len_lo[r]] >= upto_len) {
                        list.push_back((uint32_t)by_len_lo[r] | (1u << 30) | ((has_b ? 1u : 0u) << 31));
                        ++r;
                    }
                };
                const int last = p + h + 2;                      // dd_a in [1, h + 2] covers both queries, antipodal ties included
                auto visit = [&](int py) {
                    if (py == p || py < 0 || py >= nJ) return;
                    const uint32_t va = assigned(p, py);
                    const uint32_t vb = has_b && py != p + 1 && assigned(p + 1, py);
                    if (!(va | vb)) return;
                    emit_lo(d->h_wlen[J[py]]);
                    list.push_back((uint32_t)J[py] | (va << 30) | (vb << 31));
                };
                if (nJ <= h + 3) { for (int py = nJ - 1; py >= 0; --py) visit(py); }
                else {
                    for (int q = std::min(last, nJ - 1); q >= p + 1; --q) visit(q);
                    for (int q = last - nJ; q >= 0; --q) visit(q);              // wrapped around: the short end of J
                }
                emit_lo(-1);
                continue;
            }
            for (int s : by_len) {
                uint32_t va, vb;
                if (s < jlo) {                                   // (a) rectangle part
                    va = in_range(qa, s); vb = has_b && in_range(qb, s);
                } else {                                         // (b) tournament part
                    const int py = posJ[s];
                    if (py == p) continue;
                    va = assigned(p, py) && in_range(qa, s);
                    vb = has_b && py != p + 1 && assigned(p + 1, py) && in_range(qb, s);
                }
                if (va | vb) list.push_back((uint32_t)s | (va << 30) | (vb << 31));
            }
        }
    };
    {
        const int64_t work = (int64_t)npq * (jhi + 1);
        int nth = work < 200000 ? 1 : (int)std::min<int64_t>(4, std::max(1u, std::thread::hardware_concurrency()));
        if (const char* ev = getenv("PG_PLAN_THREADS")) nth = std::max(1, atoi(ev));
        nth = std::min(nth, npq);
        if (nth <= 1) build_lists(0, npq);
        else {
            std::vector<std::thread> th;
            for (int k = 0; k < nth; ++k)
                th.emplace_back(build_lists, (int)((int64_t)npq * k / nth), (int)((int64_t)npq * (k + 1) / nth));
            for (auto& t : th) t.join();
        }
    }
    for (int p = 0; p < nJ; p += 2) {
        const int qa = J[p], qb = p + 1 < nJ ? J[p + 1] : J[p];
        const std::vector<uint32_t>& list = lists[p / 2];
        const bool mp = std::max(d->h_wlen[qa], d->h_wlen[qb]) > rpp;
        if (mp) *multipass = true;
        // the shortest quarter of the queries is cut into items of a quarter of the size: they sort to the end of
        // the queue and level the tail of the persistent CTAs
        const int64_t cs = p >= nJ / 4 ? ch : std::max<int64_t>(NWv, (ch / 4 + NWv - 1) / NWv * NWv);
        const int64_t cc = mp ? NWv : cs;
        for (size_t i = 0; i < list.size(); i += cc) {
            PgItem2 it;
            it.q0 = qa; it.q1 = qb;
            it.rows = k1p_pick_rows(std::max(d->h_wlen[qa], d->h_wlen[qb]));
            it.pad[0] = it.pad[1] = it.pad[2] = 0;
            it.sub_begin = (int32_t)subs->size();
            const size_t e = std::min(list.size(), i + (size_t)cc);
            subs->insert(subs->end(), list.begin() + i, list.begin() + e);
            it.sub_end = (int32_t)subs->size();
            items->push_back(it);
        }
    }
    std::stable_sort(items->begin(), items->end(), [&](const PgItem2& x, const PgItem2& y) {
        int64_t cx = (int64_t)std::max(d->h_wlen[x.q0], d->h_wlen[x.q1]) * (x.sub_end - x.sub_begin);
        int64_t cy = (int64_t)std::max(d->h_wlen[y.q0], d->h_wlen[y.q1]) * (y.sub_end - y.sub_begin);
        return cx > cy;
    });
}

static int stage_common(pg_context* c, cudaStream_t st, const IntScoring& sc, int dim, const std::vector<PgItem>& items,
                        bool multipass, int grid_blocks, int max_wlen, K1Args* a)
{
    int rc;
    if ((rc = ensure_cap(c, &c->d_mtx, &c->mtx_cap, sizeof(int32_t) * sc.mtx.size()))) return rc;
    PG_CUDA(c, cudaMemcpyAsync(c->d_mtx, sc.mtx.data(), sizeof(int32_t) * sc.mtx.size(), cudaMemcpyHostToDevice, st));
    if ((rc = ensure_cap(c, &c->d_items, &c->items_cap, sizeof(PgItem) * std::max<size_t>(items.size(), 1)))) return rc;
    if (!items.empty())
        PG_CUDA(c, cudaMemcpyAsync(c->d_items, items.data(), sizeof(PgItem) * items.size(), cudaMemcpyHostToDevice, st));
    a->rowbuf = nullptr;
    a->rowbuf_stride = 0;
    if (multipass) {
        size_t stride = (size_t)max_wlen + 8;
        size_t need = sizeof(int2) * stride * (size_t)grid_blocks * k1_warps_per_block();
        if ((rc = ensure_cap(c, &c->d_rowbuf, &c->rowbuf_cap, need))) return rc;
        a->rowbuf = (int2*)c->d_rowbuf;
        a->rowbuf_stride = (int64_t)stride;
    }
    PG_CUDA(c, cudaMemsetAsync(c->d_counter, 0, sizeof(int32_t), st));
    a->items = (const PgItem*)c->d_items;
    a->nitems = (int32_t)items.size();
    a->counter = c->d_counter;
    a->mtx = (const int32_t*)c->d_mtx;
    a->dim = dim;
    a->u = sc.u;
    a->v = sc.v;
    return PG_OK;
}

// ---- calcdist ---------------------------------------------------------------------------------
extern "C" int pg_calcdist_dev(pg_context* c, pg_dev_seqs* d, const pg_params* prm, const void* mtx, int32_t dim,
                               int64_t k_begin, int64_t k_end, void* d_out_dist, void* stream, int32_t* n_launches)
{
    if (!c) return PG_ERR_ARG;
    if (!d || !prm || !mtx) return fail(c, PG_ERR_ARG, "pg_calcdist_dev: NULL argument");
    const int64_t n = d->v.nseq, npair = n * (n - 1) / 2;
    if (k_begin < 0 || k_end > npair || k_begin > k_end) return fail(c, PG_ERR_ARG, "pg_calcdist_dev: bad k range");
    if (n_launches) *n_launches = 0;
    if (k_begin == k_end) return PG_OK;
    if (!d_out_dist) return fail(c, PG_ERR_ARG, "pg_calcdist_dev: output is NULL");
    PG_CUDA(c, cudaSetDevice(c->device));
    if (d->max_code >= dim) return fail(c, PG_ERR_ARG, "residue code outside the substitution matrix");
    if (prm->lcl & 16)
        return fail(c, PG_ERR_UNSUPPORTED, "calcdist with algmode.lcl & 16: the reference reads `ends` uninitialised "
                                           "there (aln2.cc:296-305), there is no defined result to reproduce");
    IntScoring sc;
    ScorePlan plan;
    int rc = plan_scoring(c, prm, mtx, dim, d, prm->lcl != 0, &sc, &plan);
    if (rc) return rc;
    // everything (staging copies, self-score kernel, fill kernel) is ordered on one stream; the context's workspace
    // (work-queue counter, matrix, items, self scores, scratch) passes from the stream that used it last to this one
    cudaStream_t st = stream ? (cudaStream_t)stream : c->stream;
    PG_CUDA(c, pg_int_order_stream(c, st));
    if (!plan.integer) {
        std::vector<PgItem> fitems;
        bool fmp = false;
        const int vt = prm->vtype ? 1 : 0;
        build_calcdist_items_rows_a(d, k_begin, k_end, k1f_grid_blocks(c->sm_count, vt, plan.mode),
                                    k1f_rows_per_pass(vt, plan.mode, d->h_wlen.data(), (int)d->h_wlen.size()), &fitems, &fmp);
        return k1f_run(c, st, d, prm, mtx, dim, plan.mode, prm->lcl ? 2 : 1, fitems, fmp, nullptr, nullptr, k_begin, k_end,
                       prm->lcl ? (prm->lcl & 15) : -1, d_out_dist, nullptr, n_launches);
    }
    const int grid = c->sm_count * k1_blocks_per_sm();
    // ---- packed int16x2 path when every value provably fits 16 bits (k1p_fits)
    {
        int smax = -1 << 30, smin = 1 << 30;
        for (int i = 0; i < dim; ++i)
            for (int j = 0; j < dim; ++j)
                if (d->present[i] && d->present[j]) {
                    smax = std::max(smax, sc.mtx[i * dim + j] + 2 * sc.u);
                    smin = std::min(smin, sc.mtx[i * dim + j] + 2 * sc.u);
                }
        const char* force = getenv("PG_FORCE_INT32");
        const bool packed = !(force && force[0] == '1') && d->min_wlen >= 1 && n >= 2 &&
                            d->res_bytes < ((int64_t)1 << 31) &&       // K1P addresses residues by 32-bit offsets
                            k1p_fits(smax, smin, sc.v, d->max_wlen);
        if (packed) {
            if (d->plan_k0 != k_begin || d->plan_k1 != k_end || !d->d_plan) {
                std::vector<PgItem2> items2;
                std::vector<uint32_t> subs;
                bool mp = false;
                build_packed_plan(d, k_begin, k_end, grid, &items2, &subs, &mp);
                const size_t ib = up256(sizeof(PgItem2) * std::max<size_t>(items2.size(), 1));
                const size_t need = ib + sizeof(uint32_t) * std::max<size_t>(subs.size(), 1);
                if (d->owns) {
                    if (d->d_plan && d->plan_cap < need) {
                        PG_CUDA(c, cudaStreamSynchronize(st)); PG_CUDA(c, cudaFree(d->d_plan)); d->d_plan = nullptr;
                    }
                    if (!d->d_plan) { PG_CUDA(c, cudaMalloc(&d->d_plan, need)); d->plan_cap = need; }
                } else {
                    if ((rc = ensure_cap(c, &c->d_planbuf, &c->planbuf_cap, need))) return rc;
                    d->d_plan = c->d_planbuf;
                }
                if (!items2.empty())
                    PG_CUDA(c, cudaMemcpyAsync(d->d_plan, items2.data(), sizeof(PgItem2) * items2.size(), cudaMemcpyHostToDevice, st));
                if (!subs.empty())
                    PG_CUDA(c, cudaMemcpyAsync((char*)d->d_plan + ib, subs.data(), sizeof(uint32_t) * subs.size(), cudaMemcpyHostToDevice, st));
                d->plan_k0 = k_begin; d->plan_k1 = k_end;
                d->plan_nitems = (int32_t)items2.size();
                d->plan_nsubs = (int64_t)subs.size();
                d->plan_multipass = mp;
            }
            K1PArgs a;
            memset(&a, 0, sizeof(a));
            a.seqs = d->v;
            if ((rc = ensure_cap(c, &c->d_mtx, &c->mtx_cap, sizeof(int32_t) * sc.mtx.size()))) return rc;
            PG_CUDA(c, cudaMemcpyAsync(c->d_mtx, sc.mtx.data(), sizeof(int32_t) * sc.mtx.size(), cudaMemcpyHostToDevice, st));
            if (d->plan_multipass) {
                size_t stride = (size_t)d->max_wlen + 8;
                if ((rc = ensure_cap(c, &c->d_rowbuf, &c->rowbuf_cap, sizeof(uint2) * stride * (size_t)grid * k1p_warps_per_block()))) return rc;
                a.rowbuf = (uint2*)c->d_rowbuf;
                a.rowbuf_stride = (int64_t)stride;
            }
            PG_CUDA(c, cudaMemsetAsync(c->d_counter, 0, sizeof(int32_t), st));
            if ((rc = ensure_cap(c, &c->d_self, &c->self_cap, sizeof(int32_t) * std::max<size_t>(n, 1)))) return rc;
            PG_CUDA(c, k1_self_launch(d->v, (const int32_t*)c->d_mtx, dim, (int32_t*)c->d_self, st));
            const size_t ib = up256(sizeof(PgItem2) * std::max<size_t>((size_t)d->plan_nitems, 1));
            a.items = (const PgItem2*)d->d_plan;
            a.nitems = d->plan_nitems;
            a.counter = c->d_counter;
            a.subs = (const uint32_t*)((char*)d->d_plan + ib);
            a.k_begin = k_begin; a.k_end = k_end;
            a.mtx = (const int32_t*)c->d_mtx; a.dim = dim; a.u = sc.u; a.v = sc.v; a.sh = prm->alprm.sh;
            a.u_f32 = prm->alprm.u;
            a.self = (const int32_t*)c->d_self;
            a.epilogue = prm->vtype ? PG_EPI_DIST_F64 : PG_EPI_DIST_F32;
            a.out = d_out_dist;
            PG_CUDA(c, k1p_launch(a, c->sm_count * k1p_blocks_per_sm(), st));
            if (n_launches) *n_launches = 2;
            return PG_OK;
        }
    }
    std::vector<PgItem> items;
    bool multipass = false;
    build_calcdist_items(d, k_begin, k_end, grid, &items, &multipass);
    K1Args a;
    memset(&a, 0, sizeof(a));
    a.seqs = d->v;
    if ((rc = stage_common(c, st, sc, dim, items, multipass, grid, d->max_wlen, &a))) return rc;
    if ((rc = ensure_cap(c, &c->d_self, &c->self_cap, sizeof(int32_t) * std::max<size_t>(n, 1)))) return rc;
    PG_CUDA(c, k1_self_launch(d->v, a.mtx, dim, (int32_t*)c->d_self, st));
    a.pair_s = nullptr;
    a.pair_out = nullptr;
    a.k_begin = k_begin;
    a.k_end = k_end;
    a.sh = prm->alprm.sh;
    a.tgapf_zero = 0;
    a.u_f32 = prm->alprm.u;
    a.self = (const int32_t*)c->d_self;
    a.epilogue = prm->vtype ? PG_EPI_DIST_F64 : PG_EPI_DIST_F32;
    a.out = d_out_dist;
    a.rows_per_lane = k1_rows_per_pass(d->h_wlen.data(), (int)d->h_wlen.size()) / 32;
    PG_CUDA(c, k1_launch(a, grid, st));
    if (n_launches) *n_launches = 2;
    // Host staging vectors (items, integer matrix) are pageable: cudaMemcpyAsync has already copied
    // them into the driver's staging buffer when it returned, so they may die here.  The kernels
    // themselves stay asynchronous on `st`.
    return PG_OK;
}

static double now_ms()
{
    struct timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return ts.tv_sec * 1e3 + ts.tv_nsec * 1e-6;
}

extern "C" int pg_calcdist(pg_context* c, const pg_seqs* s, const pg_params* prm, const void* mtx, int32_t dim,
                           int64_t k_begin, int64_t k_end, void* out_dist)
{
    if (!c) return PG_ERR_ARG;
    if (!s || !prm) return fail(c, PG_ERR_ARG, "pg_calcdist: NULL argument");
    static const bool timing = getenv("PG_TIMING") != nullptr;     // PG_TIMING=1: host-side breakdown of this call on stderr
    const double t0 = timing ? now_ms() : 0;
    pg_dev_seqs* d = nullptr;
    int rc = seqs_upload_impl(c, s, &d, true);
    const double t1 = timing ? now_ms() : 0;
    if (rc) return rc;
    const size_t esz = prm->vtype ? sizeof(double) : sizeof(float);
    const size_t cnt = k_end > k_begin ? (size_t)(k_end - k_begin) : 0;
    if (cnt && !out_dist) { pg_seqs_free(c, d); return fail(c, PG_ERR_ARG, "pg_calcdist: output is NULL"); }
    rc = ensure_cap(c, &c->d_out, &c->out_cap, std::max<size_t>(cnt * esz, 16));
    if (!rc) rc = pg_calcdist_dev(c, d, prm, mtx, dim, k_begin, k_end, c->d_out, nullptr, nullptr);
    const double t2 = timing ? now_ms() : 0;
    if (!rc && cnt) {
        cudaError_t e = cudaMemcpyAsync(out_dist, c->d_out, cnt * esz, cudaMemcpyDeviceToHost, c->stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
        if (e != cudaSuccess) rc = fail(c, PG_ERR_CUDA, std::string("pg_calcdist D2H: ") + cudaGetErrorString(e));
    }
    pg_seqs_free(c, d);
    if (timing)
        fprintf(stderr, "pg_calcdist: upload %.3f ms, schedule + launches %.3f ms, kernels + D2H + sync %.3f ms\n", t1 - t0, t2 - t1,
                now_ms() - t2);
    return rc;
}

// ---- explicit pairs (alnScoreD batch; with `dist` the distance epilogue of alnscore2dist on top) ---
static int score_pairs_impl(pg_context* c, const pg_seqs* s, const int32_t* a_idx, const int32_t* b_idx,
                            int64_t npairs, const pg_params* prm, const void* mtx, int32_t dim,
                            void* out_scores, int32_t* out_ends, bool dist);

extern "C" int pg_score_pairs(pg_context* c, const pg_seqs* s, const int32_t* a_idx, const int32_t* b_idx,
                              int64_t npairs, const pg_params* prm, const void* mtx, int32_t dim,
                              void* out_scores, int32_t* out_ends)
{
    return score_pairs_impl(c, s, a_idx, b_idx, npairs, prm, mtx, dim, out_scores, out_ends, false);
}

extern "C" int pg_dist_pairs(pg_context* c, const pg_seqs* s, const int32_t* a_idx, const int32_t* b_idx,
                             int64_t npairs, const pg_params* prm, const void* mtx, int32_t dim, void* out_dist)
{
    if (c && prm && (prm->lcl & 16))
        return fail(c, PG_ERR_UNSUPPORTED, "pg_dist_pairs with algmode.lcl & 16: the reference reads `ends` uninitialised "
                                           "there (aln2.cc:296-305), there is no defined result to reproduce");
    return score_pairs_impl(c, s, a_idx, b_idx, npairs, prm, mtx, dim, out_dist, nullptr, true);
}

static int score_pairs_impl(pg_context* c, const pg_seqs* s, const int32_t* a_idx, const int32_t* b_idx,
                            int64_t npairs, const pg_params* prm, const void* mtx, int32_t dim,
                            void* out_scores, int32_t* out_ends, bool dist)
{
    if (!c) return PG_ERR_ARG;
    if (!s || !prm || !mtx || npairs < 0 || (npairs && (!a_idx || !b_idx || !out_scores)))
        return fail(c, PG_ERR_ARG, "pg_score_pairs: NULL / bad argument");
    if (npairs == 0) return PG_OK;
    if (npairs > 0x7fffffff) return fail(c, PG_ERR_ARG, "pg_score_pairs: too many pairs in one call");
    if (dim < 1 || dim > 32) return fail(c, PG_ERR_ARG, "dim must be in [1, 32]");
    int rc;
    for (int64_t p = 0; p < npairs; ++p)
        if (a_idx[p] < 0 || a_idx[p] >= s->nseq || b_idx[p] < 0 || b_idx[p] >= s->nseq)
            return fail(c, PG_ERR_ARG, "pg_score_pairs: sequence index out of range");
    pg_dev_seqs* d = nullptr;
    if ((rc = seqs_upload_impl(c, s, &d, true))) return rc;
    if (d->max_code >= dim) { pg_seqs_free(c, d); return fail(c, PG_ERR_ARG, "residue code outside the substitution matrix"); }
    IntScoring sc;
    ScorePlan plan;
    // dist: alnscore2dist's algmode.lcl branch runs the semi-global score with end points (aln2.cc:296-320)
    if ((rc = plan_scoring(c, prm, mtx, dim, d, dist ? prm->lcl != 0 : out_ends != nullptr, &sc, &plan))) { pg_seqs_free(c, d); return rc; }
    // rows = a (query of the work item), columns = b; sort pairs by a so that one CTA reuses the profile
    std::vector<int32_t> order(npairs);
    std::iota(order.begin(), order.end(), 0);
    std::stable_sort(order.begin(), order.end(), [&](int32_t x, int32_t y) { return a_idx[x] < a_idx[y]; });
    std::vector<int32_t> pair_s(npairs);
    std::vector<int64_t> pair_out(npairs);
    for (int64_t p = 0; p < npairs; ++p) { pair_s[p] = b_idx[order[p]]; pair_out[p] = order[p]; }
    const int vt = prm->vtype ? 1 : 0;
    const int grid = plan.integer ? c->sm_count * k1_blocks_per_sm() : k1f_grid_blocks(c->sm_count, vt, plan.mode);
    const int NWv = k1_warps_per_block(), rpp = plan.integer ? k1_rows_per_pass(d->h_wlen.data(), (int)d->h_wlen.size()) : k1f_rows_per_pass(vt, plan.mode, d->h_wlen.data(), (int)d->h_wlen.size());
    int64_t ch = (npairs + (int64_t)16 * grid - 1) / ((int64_t)16 * grid);
    ch = std::max<int64_t>(NWv, std::min<int64_t>(ch, 32 * NWv));
    ch = (ch + NWv - 1) / NWv * NWv;
    std::vector<PgItem> items;
    bool multipass = false;
    for (int64_t p = 0; p < npairs;) {
        int q = a_idx[order[p]];
        int64_t e = p;
        while (e < npairs && a_idx[order[e]] == q) ++e;
        const bool mp = d->h_wlen[q] > rpp;
        if (mp) multipass = true;
        const int64_t cc = mp ? NWv : ch;
        for (int64_t i = p; i < e; i += cc) {
            PgItem it;
            it.q = q; it.sub_begin = (int32_t)i; it.sub_end = (int32_t)std::min<int64_t>(i + cc, e); it.pad = 0;
            items.push_back(it);
        }
        p = e;
    }
    std::stable_sort(items.begin(), items.end(), [&](const PgItem& x, const PgItem& y) {
        return (int64_t)d->h_wlen[x.q] * (x.sub_end - x.sub_begin) > (int64_t)d->h_wlen[y.q] * (y.sub_end - y.sub_begin);
    });
    const size_t esz = prm->vtype ? sizeof(double) : sizeof(float);
    if (!plan.integer) {
        // floating-point / semi-global / local kernel, explicit pairs
        cudaError_t e = cudaSuccess;
        rc = ensure_cap(c, &c->d_pairs, &c->pairs_cap, (sizeof(int32_t) + sizeof(int64_t)) * (size_t)npairs + 64);
        if (!rc) rc = ensure_cap(c, &c->d_out, &c->out_cap, esz * (size_t)npairs);
        if (!rc && out_ends) rc = ensure_cap(c, &c->d_ends, &c->ends_cap, 2 * sizeof(int32_t) * (size_t)npairs);
        if (!rc) {
            int64_t* d_po = (int64_t*)c->d_pairs;
            int32_t* d_ps = (int32_t*)((char*)c->d_pairs + sizeof(int64_t) * (size_t)npairs);
            e = cudaMemcpyAsync(d_po, pair_out.data(), sizeof(int64_t) * npairs, cudaMemcpyHostToDevice, c->stream);
            if (e == cudaSuccess) e = cudaMemcpyAsync(d_ps, pair_s.data(), sizeof(int32_t) * npairs, cudaMemcpyHostToDevice, c->stream);
            if (e != cudaSuccess) rc = fail(c, PG_ERR_CUDA, std::string("pg_score_pairs: ") + cudaGetErrorString(e));
            if (!rc) rc = k1f_run(c, c->stream, d, prm, mtx, dim, plan.mode, dist ? (prm->lcl ? 2 : 1) : 0, items, multipass, d_ps, d_po,
                                  0, 0, dist && prm->lcl ? (prm->lcl & 15) : -1, c->d_out,
                                  out_ends ? (int32_t*)c->d_ends : nullptr, nullptr);
            if (!rc) {
                e = cudaMemcpyAsync(out_scores, c->d_out, esz * npairs, cudaMemcpyDeviceToHost, c->stream);
                if (e == cudaSuccess && out_ends)
                    e = cudaMemcpyAsync(out_ends, c->d_ends, 2 * sizeof(int32_t) * npairs, cudaMemcpyDeviceToHost, c->stream);
                if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
                if (e != cudaSuccess) rc = fail(c, PG_ERR_CUDA, std::string("pg_score_pairs: ") + cudaGetErrorString(e));
            }
        }
        pg_seqs_free(c, d);
        return rc;
    }
    K1Args a;
    memset(&a, 0, sizeof(a));
    a.seqs = d->v;
    cudaError_t e = cudaSuccess;
    rc = stage_common(c, c->stream, sc, dim, items, multipass, grid, d->max_wlen, &a);
    if (!rc) rc = ensure_cap(c, &c->d_pairs, &c->pairs_cap, (sizeof(int32_t) + sizeof(int64_t)) * (size_t)npairs + 64);
    if (!rc) rc = ensure_cap(c, &c->d_out, &c->out_cap, esz * (size_t)npairs);
    if (!rc && dist) rc = ensure_cap(c, &c->d_self, &c->self_cap, sizeof(int32_t) * std::max<size_t>(d->v.nseq, 1));
    if (!rc) {
        int64_t* d_po = (int64_t*)c->d_pairs;
        int32_t* d_ps = (int32_t*)((char*)c->d_pairs + sizeof(int64_t) * (size_t)npairs);
        e = cudaMemcpyAsync(d_po, pair_out.data(), sizeof(int64_t) * npairs, cudaMemcpyHostToDevice, c->stream);
        if (e == cudaSuccess) e = cudaMemcpyAsync(d_ps, pair_s.data(), sizeof(int32_t) * npairs, cudaMemcpyHostToDevice, c->stream);
        if (e == cudaSuccess && dist) e = k1_self_launch(d->v, a.mtx, dim, (int32_t*)c->d_self, c->stream);
        a.pair_s = d_ps;
        a.pair_out = d_po;
        a.sh = prm->alprm.sh;
        a.u_f32 = prm->alprm.u;
        a.self = dist ? (const int32_t*)c->d_self : nullptr;
        a.epilogue = dist ? (prm->vtype ? PG_EPI_DIST_F64 : PG_EPI_DIST_F32) : (prm->vtype ? PG_EPI_SCORE_F64 : PG_EPI_SCORE_F32);
        a.out = c->d_out;
        a.rows_per_lane = k1_rows_per_pass(d->h_wlen.data(), (int)d->h_wlen.size()) / 32;
        if (e == cudaSuccess) e = k1_launch(a, grid, c->stream);
        if (e == cudaSuccess) e = cudaMemcpyAsync(out_scores, c->d_out, esz * npairs, cudaMemcpyDeviceToHost, c->stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
        if (e != cudaSuccess) rc = fail(c, PG_ERR_CUDA, std::string("pg_score_pairs: ") + cudaGetErrorString(e));
    }
    pg_seqs_free(c, d);
    return rc;
}

// ---- alignment with path (alignC<DPunit> batch) ------------------------------------------------
extern "C" void pg_free(void* p) { free(p); }


extern "C" int pg_align_pairs(pg_context* c, const pg_seqs* s, const int32_t* a_idx, const int32_t* b_idx,
                              int64_t npairs, const pg_params* prm, const void* mtx, int32_t dim,
                              void* out_scores, int64_t** out_offs, pg_skl** out_pts)
{
    if (!c) return PG_ERR_ARG;
    if (!s || !prm || !mtx || npairs < 0 || !out_offs || !out_pts || (npairs && (!a_idx || !b_idx || !out_scores)))
        return fail(c, PG_ERR_ARG, "pg_align_pairs: NULL / bad argument");
    *out_offs = nullptr;
    *out_pts = nullptr;
    if (npairs > 0x7fffffff) return fail(c, PG_ERR_ARG, "pg_align_pairs: too many pairs in one call");
    if (dim < 1 || dim > 32) return fail(c, PG_ERR_ARG, "dim must be in [1, 32]");
    for (int64_t p = 0; p < npairs; ++p)
        if (a_idx[p] < 0 || a_idx[p] >= s->nseq || b_idx[p] < 0 || b_idx[p] >= s->nseq)
            return fail(c, PG_ERR_ARG, "pg_align_pairs: sequence index out of range");
    {   // the exact-integer bit kernel (K2) takes integral affine scoring; two-piece penalties and
        // non-integral tables (default PAM) run as groups of one on the floating-point kernel K3
        bool fp = prm->alprm.ls >= 3;
        const char* force = getenv("PG_FORCE_FLOAT");
        if (force && force[0] == '1') fp = true;
        if (!fp) {
            uint8_t present[256];
            memset(present, 0, sizeof(present));
            for (int64_t p = 0; p < npairs; ++p)
                for (int side = 0; side < 2; ++side) {
                    const int i = side ? b_idx[p] : a_idx[p];
                    if (s->lens[i] < 0 || s->offs[i] < 0) return fail(c, PG_ERR_ARG, "negative length / offset");
                    const int l = s->left ? s->left[i] : 0, r = s->right ? s->right[i] : s->lens[i];
                    if (l < 0 || r > s->lens[i] || l > r) return fail(c, PG_ERR_ARG, "window outside the sequence");
                    for (int k = l; k < r; ++k) present[s->res[s->offs[i] + k]] = 1;
                }
            IntScoring probe;
            std::string keep = c->err;
            pg_params p0 = *prm;
            p0.lcl = 0; p0.alprm.tgapf = 1.0f;
            if (make_int_scoring(c, &p0, mtx, dim, present, &probe) != PG_OK) { fp = true; c->err = keep; }
        }
        if (fp) return pg_int_align_pairs_fp(c, s, a_idx, b_idx, npairs, prm, mtx, dim, out_scores, out_offs, out_pts, 0);
    }
    int64_t* offs = (int64_t*)malloc(sizeof(int64_t) * (size_t)(npairs + 1));
    if (!offs) return fail(c, PG_ERR_ARG, "out of host memory");
    offs[0] = 0;
    if (npairs == 0) { *out_offs = offs; *out_pts = (pg_skl*)malloc(sizeof(pg_skl)); return PG_OK; }
    pg_dev_seqs* d = nullptr;
    int rc = seqs_upload_impl(c, s, &d, true);
    if (rc) { free(offs); return rc; }
    IntScoring sc;
    if (d->max_code >= dim) rc = fail(c, PG_ERR_ARG, "residue code outside the substitution matrix");
    if (!rc) rc = make_int_scoring(c, prm, mtx, dim, d->present, &sc);
    if (rc) { pg_seqs_free(c, d); free(offs); return rc; }

    // rows = a, columns = b (the tie rules of forwardB are not symmetric); sort by a for profile reuse
    std::vector<int32_t> order(npairs);
    std::iota(order.begin(), order.end(), 0);
    std::stable_sort(order.begin(), order.end(), [&](int32_t x, int32_t y) { return a_idx[x] < a_idx[y]; });
    const int Rr = k2_rows_per_lane(), NWv = k2_warps_per_block(), rpp = 32 * Rr;
    const int grid = c->sm_count * k2_blocks_per_sm();
    const size_t esz = prm->vtype ? sizeof(double) : sizeof(float);
    const size_t DIR_BUDGET_WORDS = (size_t)3 << 30;      // 24 GB of direction words per launch
    std::vector<int32_t> h_score(npairs), h_cnt(npairs);
    std::vector<int64_t> h_lenoff(npairs + 1);
    std::vector<int32_t> h_pts;
    cudaError_t e = cudaSuccess;
    std::vector<std::vector<pg_skl>> lists;    // not used; corner lists are compacted below
    std::vector<pg_skl> all_pts;
    std::vector<int64_t> cnt_by_orig(npairs);
    std::vector<int64_t> start_sorted(npairs);

    for (int64_t c0 = 0; c0 < npairs && !rc;) {
        // ---- chunk of sorted pairs whose direction words fit the budget
        std::vector<int32_t> pq, ps;
        std::vector<int64_t> diroff, lenoff;
        size_t words = 0;
        int64_t lens = 0;
        int64_t c1 = c0;
        bool multipass = false;
        int max_ls = 0;
        // a pair whose query spans LONG_PASSES stripes or more gets the striped multi-warp kernel, alone
        const int LONG_PASSES = 6;
        bool long_pair = false;
        while (c1 < npairs) {
            const int qa = a_idx[order[c1]], sb = b_idx[order[c1]];
            const bool is_long = d->h_wlen[qa] > (LONG_PASSES - 1) * rpp && d->h_wlen[sb] > 0;
            // (the striped kernel chooses its own rows per lane, and with them the size of a lane-step's direction bits)
            const size_t w = (size_t)k2_words_per_pair(d->h_wlen[qa], d->h_wlen[sb],
                                                        is_long ? k2_long_rows(d->h_wlen[qa], d->h_wlen[sb]) : Rr);
            if (is_long) {
                if (c1 > c0) break;             // close the chunk before it
                long_pair = true;
            }
            if (c1 > c0 && words + w > DIR_BUDGET_WORDS) break;
            pq.push_back(qa); ps.push_back(sb);
            diroff.push_back((int64_t)words); lenoff.push_back(lens);
            words += w;
            lens += d->h_wlen[qa] + d->h_wlen[sb] + 8;
            if (d->h_wlen[qa] > rpp) multipass = true;
            max_ls = std::max(max_ls, d->h_wlen[sb]);
            ++c1;
            if (long_pair) break;
        }
        const int64_t np = c1 - c0;
        // ---- work items: runs of equal query, NW subjects per warp round
        std::vector<PgItem> items;
        int64_t ch = (np + (int64_t)8 * grid - 1) / ((int64_t)8 * grid);
        ch = std::max<int64_t>(NWv, std::min<int64_t>(ch, 8 * NWv));
        ch = (ch + NWv - 1) / NWv * NWv;
        for (int64_t p = 0; p < np;) {
            int64_t e2 = p;
            while (e2 < np && pq[e2] == pq[p]) ++e2;
            const int64_t cc = d->h_wlen[pq[p]] > rpp ? NWv : ch;
            for (int64_t i = p; i < e2; i += cc) {
                PgItem it; it.q = pq[p]; it.sub_begin = (int32_t)i; it.sub_end = (int32_t)std::min<int64_t>(i + cc, e2); it.pad = 0;
                items.push_back(it);
            }
            p = e2;
        }
        std::stable_sort(items.begin(), items.end(), [&](const PgItem& x, const PgItem& y) {
            return (int64_t)d->h_wlen[x.q] * (x.sub_end - x.sub_begin) > (int64_t)d->h_wlen[y.q] * (y.sub_end - y.sub_begin);
        });
        // ---- device buffers
        K1Args k1a;
        memset(&k1a, 0, sizeof(k1a));
        rc = stage_common(c, c->stream, sc, dim, items, multipass, grid, max_ls, &k1a);
        if (rc) break;
        // pairs blob: pair_q | pair_s | dir_off | len_off | score | cnt
        size_t o_q = 0, o_s = up256(o_q + 4 * np), o_do = up256(o_s + 4 * np), o_lo = up256(o_do + 8 * np),
               o_sc = up256(o_lo + 8 * np), o_cn = up256(o_sc + 4 * np), pbytes = up256(o_cn + 4 * np);
        if ((rc = ensure_cap(c, &c->d_pairs, &c->pairs_cap, pbytes))) break;
        if ((rc = ensure_cap(c, &c->d_dirs, &c->dirs_cap, std::max<size_t>(words, 1) * 8))) break;
        // trace blob: moves (1 B) | recs (12 B) | out (8 B) per unit of len
        size_t o_mv = 0, o_rc = up256(o_mv + (size_t)lens), o_out = up256(o_rc + 12 * (size_t)lens),
               tbytes = up256(o_out + 8 * (size_t)lens);
        if ((rc = ensure_cap(c, &c->d_trace, &c->trace_cap, tbytes))) break;
        char* pb = (char*)c->d_pairs;
        e = cudaMemcpyAsync(pb + o_q, pq.data(), 4 * np, cudaMemcpyHostToDevice, c->stream);
        if (e == cudaSuccess) e = cudaMemcpyAsync(pb + o_s, ps.data(), 4 * np, cudaMemcpyHostToDevice, c->stream);
        if (e == cudaSuccess) e = cudaMemcpyAsync(pb + o_do, diroff.data(), 8 * np, cudaMemcpyHostToDevice, c->stream);
        if (e == cudaSuccess) e = cudaMemcpyAsync(pb + o_lo, lenoff.data(), 8 * np, cudaMemcpyHostToDevice, c->stream);
        K2Args a;
        memset(&a, 0, sizeof(a));
        a.seqs = d->v;
        a.items = k1a.items; a.nitems = k1a.nitems; a.counter = k1a.counter;
        a.pair_q = (const int32_t*)(pb + o_q); a.pair_s = (const int32_t*)(pb + o_s);
        a.dir_off = (const int64_t*)(pb + o_do); a.len_off = (const int64_t*)(pb + o_lo);
        a.dirs = (unsigned long long*)c->d_dirs;
        a.mtx = k1a.mtx; a.dim = dim; a.u = sc.u; a.v = sc.v; a.sh = prm->alprm.sh;
        a.score = (int32_t*)(pb + o_sc);
        a.rowbuf = k1a.rowbuf; a.rowbuf_stride = k1a.rowbuf_stride;
        char* tb = (char*)c->d_trace;
        a.moves = (unsigned char*)(tb + o_mv); a.recs = (K2Rec*)(tb + o_rc); a.out_pts = (int32_t*)(tb + o_out);
        a.out_cnt = (int32_t*)(pb + o_cn);
        int long_npass = 0;
        size_t long_rb = 0;
        if (long_pair) {
            // row buffers of all stripes + progress counters + ticket, zeroed
            a.rows_per_lane = k2_long_rows(d->h_wlen[pq[0]], max_ls);
            long_npass = (d->h_wlen[pq[0]] + 32 * a.rows_per_lane - 1) / (32 * a.rows_per_lane);
            const size_t rb = sizeof(int2) * (size_t)long_npass * (size_t)max_ls;
            const size_t need = up256(rb) + sizeof(int32_t) * (size_t)(long_npass + 2);
            long_rb = rb;
            if ((rc = ensure_cap(c, &c->d_rowbuf, &c->rowbuf_cap, need))) break;
            a.rowbuf = (int2*)c->d_rowbuf;
            a.progress = (int32_t*)((char*)c->d_rowbuf + up256(rb));
            a.ticket = a.progress + long_npass;
            if (e == cudaSuccess) e = cudaMemsetAsync(a.progress, 0, sizeof(int32_t) * (size_t)(long_npass + 2), c->stream);
        }
        if (e == cudaSuccess) e = cudaEventRecord(c->ev0, c->stream);
        if (e == cudaSuccess) e = long_pair ? k2_fill_long_launch(a, long_npass, c->sm_count, long_rb, c->stream)
                                            : k2_fill_launch(a, grid, c->stream);
        if (e == cudaSuccess) e = cudaEventRecord(c->ev1, c->stream);
        c->ev_valid = e == cudaSuccess;
        if (e == cudaSuccess) e = k2_trace_launch(a, (int)np, c->stream);
        h_pts.resize(2 * (size_t)lens);
        if (e == cudaSuccess) e = cudaMemcpyAsync(h_score.data() + c0, pb + o_sc, 4 * np, cudaMemcpyDeviceToHost, c->stream);
        if (e == cudaSuccess) e = cudaMemcpyAsync(h_cnt.data() + c0, pb + o_cn, 4 * np, cudaMemcpyDeviceToHost, c->stream);
        if (e == cudaSuccess) e = cudaMemcpyAsync(h_pts.data(), tb + o_out, 8 * (size_t)lens, cudaMemcpyDeviceToHost, c->stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
        if (e != cudaSuccess) { rc = fail(c, PG_ERR_CUDA, std::string("pg_align_pairs: ") + cudaGetErrorString(e)); break; }
        // ---- compact this chunk's corner lists (sorted order) into all_pts
        for (int64_t p = 0; p < np; ++p) {
            start_sorted[c0 + p] = (int64_t)all_pts.size();
            const int32_t* src = h_pts.data() + 2 * lenoff[p];
            for (int k = 0; k < h_cnt[c0 + p]; ++k) { pg_skl q; q.m = src[2 * k]; q.n = src[2 * k + 1]; all_pts.push_back(q); }
        }
        c0 = c1;
    }
    pg_seqs_free(c, d);
    if (rc) { free(offs); return rc; }
    // ---- back to the caller's pair order
    for (int64_t p = 0; p < npairs; ++p) cnt_by_orig[order[p]] = h_cnt[p];
    for (int64_t p = 0; p < npairs; ++p) offs[p + 1] = offs[p] + cnt_by_orig[p];
    pg_skl* pts = (pg_skl*)malloc(sizeof(pg_skl) * (size_t)std::max<int64_t>(offs[npairs], 1));
    if (!pts) { free(offs); return fail(c, PG_ERR_ARG, "out of host memory"); }
    for (int64_t p = 0; p < npairs; ++p) {
        const int64_t o = order[p];
        memcpy(pts + offs[o], all_pts.data() + start_sorted[p], sizeof(pg_skl) * (size_t)h_cnt[p]);
        if (prm->vtype) ((double*)out_scores)[o] = (double)h_score[p];
        else ((float*)out_scores)[o] = (float)h_score[p];
    }
    (void)esz;
    *out_offs = offs;
    *out_pts = pts;
    return PG_OK;
}

// ---- measurement helpers ----------------------------------------------------------------------
// sum_{m=0}^{n-1} clamp(a0 + m, 0, hi)
static int64_t sum_clamped(int64_t a0, int64_t n, int64_t hi)
{
    if (n <= 0 || hi <= 0) return 0;
    // m < m0: value 0 ; m0 <= m < m1: a0 + m ; m >= m1: hi
    int64_t m0 = std::max<int64_t>(0, -a0);            // first m with a0 + m >= 0
    int64_t m1 = std::max<int64_t>(m0, hi - a0);       // first m with a0 + m >= hi
    m0 = std::min(m0, n);
    m1 = std::min(m1, n);
    int64_t cnt = m1 - m0;
    int64_t lin = cnt > 0 ? cnt * a0 + (m0 + m1 - 1) * cnt / 2 : 0;
    return lin + (n - m1) * hi;
}

static int64_t band_cells_closed(int LQ, int LS, int sh)
{
    if (LQ <= 0 || LS <= 0) return 0;
    int lw, up;
    k1_band(LQ, LS, sh, &lw, &up);
    // row m: columns [clamp(m+lw,0,LS), clamp(m+up+1,0,LS))
    int64_t hi = sum_clamped((int64_t)up + 1, LQ, LS);
    int64_t lo = sum_clamped((int64_t)lw, LQ, LS);
    return hi - lo;
}

extern "C" int64_t pg_calcdist_cells(const pg_seqs* s, const pg_params* prm, int64_t k_begin, int64_t k_end)
{
    if (!s || !prm || k_end <= k_begin) return 0;
    int64_t cells = 0;
    const int jlo = row_of_k(k_begin), jhi = row_of_k(k_end - 1);
    auto wl = [&](int i) { return (s->right ? s->right[i] : s->lens[i]) - (s->left ? s->left[i] : 0); };
    for (int j = jlo; j <= jhi; ++j) {
        int64_t i0 = j == jlo ? k_begin - tri(j) : 0;
        int64_t i1 = j == jhi ? k_end - tri(j) : j;
        const int lb = wl(j);
        for (int64_t i = i0; i < i1; ++i) cells += band_cells_closed(wl((int)i), lb, prm->alprm.sh);
    }
    return cells;
}

// Host-only view of the packed schedule (no device needed): how many (query pair, subject) slots the
// plan has and how often every condensed index k in [k_begin, k_end) is covered (must be exactly 1).
extern "C" int pg_debug_packed_plan(const pg_seqs* s, int64_t k_begin, int64_t k_end, int32_t grid_blocks,
                                    int64_t* nitems, int64_t* nslots, uint8_t* cover)
{
    if (!s || k_end < k_begin) return PG_ERR_ARG;
    pg_dev_seqs d;
    d.owns = false; d.blob = nullptr; d.d_plan = nullptr;
    d.h_wlen.resize(s->nseq);
    d.max_wlen = 0;
    for (int i = 0; i < s->nseq; ++i) {
        d.h_wlen[i] = (s->right ? s->right[i] : s->lens[i]) - (s->left ? s->left[i] : 0);
        d.max_wlen = std::max(d.max_wlen, d.h_wlen[i]);
    }
    std::vector<PgItem2> items;
    std::vector<uint32_t> subs;
    bool mp = false;
    build_packed_plan(&d, k_begin, k_end, grid_blocks, &items, &subs, &mp);
    if (nitems) *nitems = (int64_t)items.size();
    if (nslots) *nslots = (int64_t)subs.size();
    if (cover) {
        memset(cover, 0, (size_t)(k_end - k_begin));
        for (const PgItem2& it : items)
            for (int32_t p = it.sub_begin; p < it.sub_end; ++p) {
                const int64_t si = subs[p] & 0x3fffffffu;
                const int64_t q[2] = {it.q0, it.q1};
                for (int h = 0; h < 2; ++h)
                    if ((subs[p] >> (30 + h)) & 1u) {
                        const int64_t hi = std::max(q[h], si), lo = std::min(q[h], si);
                        const int64_t k = hi * (hi - 1) / 2 + lo;
                        if (k >= k_begin && k < k_end && cover[k - k_begin] < 255) ++cover[k - k_begin];
                    }
            }
    }
    return PG_OK;
}

extern "C" double pg_last_kernel_ms(pg_context* c)
{
    if (!c || !c->ev_valid) return -1.0;
    float ms = 0;
    if (cudaEventElapsedTime(&ms, c->ev0, c->ev1) != cudaSuccess) return -1.0;
    return (double)ms;
}

extern "C" int pg_dpx_peak(pg_context* c, double* gops_s32, double* gops_s16x2)
{
    if (!c || !gops_s32 || !gops_s16x2) return PG_ERR_ARG;
    PG_CUDA(c, cudaSetDevice(c->device));
    PG_CUDA(c, dpx_peak_run(c->sm_count, c->stream, gops_s32, gops_s16x2));
    return PG_OK;
}
