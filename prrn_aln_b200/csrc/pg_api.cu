// pg_api.cu -- the extern "C" layer of libprrn_gpu.so (include/prrn_gpu.h): argument checks,
// staging of sequences / score matrix / work items into HBM, kernel launches.  No CPU fallback:
// every compute entry fails loudly when the device or a kernel is unavailable.
#include <cuda_runtime.h>
#include <math.h>
#include <stdio.h>
#include <string.h>

#include <algorithm>
#include <numeric>
#include <string>
#include <vector>

#include "k1_core.cuh"
#include "pg_internal.h"

static std::string g_create_err;

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
    c->d_items = c->d_mtx = c->d_self = c->d_rowbuf = c->d_out = c->d_pairs = nullptr;
    c->items_cap = c->mtx_cap = c->self_cap = c->rowbuf_cap = c->out_cap = c->pairs_cap = 0;
    c->d_counter = nullptr;
    if ((e = cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking)) != cudaSuccess ||
        (e = cudaMalloc(&c->d_counter, sizeof(int32_t))) != cudaSuccess) {
        std::string m = std::string("pg_create: ") + cudaGetErrorString(e);
        delete c;
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
    cudaFree(c->d_out); cudaFree(c->d_pairs); cudaFree(c->d_counter);
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

// ---- sequences --------------------------------------------------------------------------------
extern "C" int pg_seqs_upload(pg_context* c, const pg_seqs* s, pg_dev_seqs** out)
{
    if (!c) return PG_ERR_ARG;
    if (!s || !out || s->nseq < 0 || (s->nseq > 0 && (!s->res || !s->offs || !s->lens)))
        return fail(c, PG_ERR_ARG, "pg_seqs_upload: bad sequence set");
    PG_CUDA(c, cudaSetDevice(c->device));
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
        if (ex & 3) return fail(c, PG_ERR_UNSUPPORTED, "inex.exgl / exgr (free end gaps) is not built yet");
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
    cudaError_t e = cudaMalloc(&d->blob, bytes);
    if (e != cudaSuccess) { delete d; return fail(c, PG_ERR_CUDA, std::string("cudaMalloc(seqs): ") + cudaGetErrorString(e)); }
    char* b = (char*)d->blob;
    e = cudaMemcpyAsync(b + o_res, s->res, (size_t)total, cudaMemcpyHostToDevice, c->stream);
    if (e == cudaSuccess && n) e = cudaMemcpyAsync(b + o_offs, s->offs, sizeof(int64_t) * n, cudaMemcpyHostToDevice, c->stream);
    if (e == cudaSuccess && n) e = cudaMemcpyAsync(b + o_left, left.data(), sizeof(int32_t) * n, cudaMemcpyHostToDevice, c->stream);
    if (e == cudaSuccess && n) e = cudaMemcpyAsync(b + o_wlen, wlen.data(), sizeof(int32_t) * n, cudaMemcpyHostToDevice, c->stream);
    if (e == cudaSuccess && n) e = cudaMemcpyAsync(b + o_flags, flags.data(), n, cudaMemcpyHostToDevice, c->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);   // host staging vectors die at return
    if (e != cudaSuccess) {
        cudaFree(d->blob); delete d;
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
    cudaFree(d->blob);
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
    const int NWv = k1_warps_per_block(), rpp = k1_rows_per_pass();
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
    IntScoring sc;
    int rc = make_int_scoring(c, prm, mtx, dim, d->present, &sc);
    if (rc) return rc;
    // everything (staging copies, self-score kernel, fill kernel) is ordered on one stream
    cudaStream_t st = stream ? (cudaStream_t)stream : c->stream;
    const int grid = c->sm_count * k1_blocks_per_sm();
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
    PG_CUDA(c, k1_launch(a, grid, st));
    if (n_launches) *n_launches = 2;
    // Host staging vectors (items, integer matrix) are pageable: cudaMemcpyAsync has already copied
    // them into the driver's staging buffer when it returned, so they may die here.  The kernels
    // themselves stay asynchronous on `st`.
    return PG_OK;
}

extern "C" int pg_calcdist(pg_context* c, const pg_seqs* s, const pg_params* prm, const void* mtx, int32_t dim,
                           int64_t k_begin, int64_t k_end, void* out_dist)
{
    if (!c) return PG_ERR_ARG;
    if (!s || !prm) return fail(c, PG_ERR_ARG, "pg_calcdist: NULL argument");
    pg_dev_seqs* d = nullptr;
    int rc = pg_seqs_upload(c, s, &d);
    if (rc) return rc;
    const size_t esz = prm->vtype ? sizeof(double) : sizeof(float);
    const size_t cnt = k_end > k_begin ? (size_t)(k_end - k_begin) : 0;
    if (cnt && !out_dist) { pg_seqs_free(c, d); return fail(c, PG_ERR_ARG, "pg_calcdist: output is NULL"); }
    rc = ensure_cap(c, &c->d_out, &c->out_cap, std::max<size_t>(cnt * esz, 16));
    if (!rc) rc = pg_calcdist_dev(c, d, prm, mtx, dim, k_begin, k_end, c->d_out, nullptr, nullptr);
    if (!rc && cnt) {
        cudaError_t e = cudaMemcpyAsync(out_dist, c->d_out, cnt * esz, cudaMemcpyDeviceToHost, c->stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
        if (e != cudaSuccess) rc = fail(c, PG_ERR_CUDA, std::string("pg_calcdist D2H: ") + cudaGetErrorString(e));
    }
    pg_seqs_free(c, d);
    return rc;
}

// ---- explicit pairs (alnScoreD batch) -----------------------------------------------------------
extern "C" int pg_score_pairs(pg_context* c, const pg_seqs* s, const int32_t* a_idx, const int32_t* b_idx,
                              int64_t npairs, const pg_params* prm, const void* mtx, int32_t dim,
                              void* out_scores, int32_t* out_ends)
{
    if (!c) return PG_ERR_ARG;
    if (!s || !prm || !mtx || npairs < 0 || (npairs && (!a_idx || !b_idx || !out_scores)))
        return fail(c, PG_ERR_ARG, "pg_score_pairs: NULL / bad argument");
    if (out_ends) return fail(c, PG_ERR_UNSUPPORTED, "`ends` output (Fwd2d_vd, fwd2d1.cc:191-322) is not built yet");
    if (npairs == 0) return PG_OK;
    if (npairs > 0x7fffffff) return fail(c, PG_ERR_ARG, "pg_score_pairs: too many pairs in one call");
    if (dim < 1 || dim > 32) return fail(c, PG_ERR_ARG, "dim must be in [1, 32]");
    int rc;
    for (int64_t p = 0; p < npairs; ++p)
        if (a_idx[p] < 0 || a_idx[p] >= s->nseq || b_idx[p] < 0 || b_idx[p] >= s->nseq)
            return fail(c, PG_ERR_ARG, "pg_score_pairs: sequence index out of range");
    pg_dev_seqs* d = nullptr;
    if ((rc = pg_seqs_upload(c, s, &d))) return rc;
    if (d->max_code >= dim) { pg_seqs_free(c, d); return fail(c, PG_ERR_ARG, "residue code outside the substitution matrix"); }
    IntScoring sc;
    if ((rc = make_int_scoring(c, prm, mtx, dim, d->present, &sc))) { pg_seqs_free(c, d); return rc; }
    // rows = a (query of the work item), columns = b; sort pairs by a so that one CTA reuses the profile
    std::vector<int32_t> order(npairs);
    std::iota(order.begin(), order.end(), 0);
    std::stable_sort(order.begin(), order.end(), [&](int32_t x, int32_t y) { return a_idx[x] < a_idx[y]; });
    std::vector<int32_t> pair_s(npairs);
    std::vector<int64_t> pair_out(npairs);
    for (int64_t p = 0; p < npairs; ++p) { pair_s[p] = b_idx[order[p]]; pair_out[p] = order[p]; }
    const int grid = c->sm_count * k1_blocks_per_sm();
    const int NWv = k1_warps_per_block(), rpp = k1_rows_per_pass();
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
    K1Args a;
    memset(&a, 0, sizeof(a));
    a.seqs = d->v;
    cudaError_t e = cudaSuccess;
    const size_t esz = prm->vtype ? sizeof(double) : sizeof(float);
    rc = stage_common(c, c->stream, sc, dim, items, multipass, grid, d->max_wlen, &a);
    if (!rc) rc = ensure_cap(c, &c->d_pairs, &c->pairs_cap, (sizeof(int32_t) + sizeof(int64_t)) * (size_t)npairs + 64);
    if (!rc) rc = ensure_cap(c, &c->d_out, &c->out_cap, esz * (size_t)npairs);
    if (!rc) {
        int64_t* d_po = (int64_t*)c->d_pairs;
        int32_t* d_ps = (int32_t*)((char*)c->d_pairs + sizeof(int64_t) * (size_t)npairs);
        e = cudaMemcpyAsync(d_po, pair_out.data(), sizeof(int64_t) * npairs, cudaMemcpyHostToDevice, c->stream);
        if (e == cudaSuccess) e = cudaMemcpyAsync(d_ps, pair_s.data(), sizeof(int32_t) * npairs, cudaMemcpyHostToDevice, c->stream);
        a.pair_s = d_ps;
        a.pair_out = d_po;
        a.sh = prm->alprm.sh;
        a.u_f32 = prm->alprm.u;
        a.epilogue = prm->vtype ? PG_EPI_SCORE_F64 : PG_EPI_SCORE_F32;
        a.out = c->d_out;
        if (e == cudaSuccess) e = k1_launch(a, grid, c->stream);
        if (e == cudaSuccess) e = cudaMemcpyAsync(out_scores, c->d_out, esz * npairs, cudaMemcpyDeviceToHost, c->stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
        if (e != cudaSuccess) rc = fail(c, PG_ERR_CUDA, std::string("pg_score_pairs: ") + cudaGetErrorString(e));
    }
    pg_seqs_free(c, d);
    return rc;
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

extern "C" int pg_dpx_peak(pg_context* c, double* gops_s32, double* gops_s16x2)
{
    if (!c || !gops_s32 || !gops_s16x2) return PG_ERR_ARG;
    PG_CUDA(c, cudaSetDevice(c->device));
    PG_CUDA(c, dpx_peak_run(c->sm_count, c->stream, gops_s32, gops_s16x2));
    return PG_OK;
}
