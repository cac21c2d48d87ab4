// pg_groups.cu -- the extern "C" layer of the group-to-group path (include/prrn_gpu.h: pg_align_groups):
// argument checks, flattening of the staged groups into one HBM blob, launch of kernel K3
// (k3_groups.cu), collection of scores and corner lists.  No CPU fallback.
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <numeric>
#include <string>
#include <thread>
#include <vector>

#include <atomic>

#include "pg_internal.h"
#include "k3r_core.cuh"

static inline size_t up16(size_t x) { return (x + 15) & ~(size_t)15; }

// stripe() (src/aln2.cc:156-174) in absolute diagonal coordinates
static void group_band(const pg_group* a, const pg_group* b, int sh, int* lw, int* up)
{
    if (sh < 0) {
        int shorter = std::min(a->right - a->left, b->right - b->left);
        sh = -sh * shorter / 100;
    }
    int hi = b->right - a->right, lo = b->left - a->left;
    if (hi < lo) std::swap(hi, lo);
    hi += sh; lo -= sh;
    hi = std::min(hi, b->right - a->left);
    lo = std::max(lo, b->left - a->right);
    *lw = lo; *up = hi;
}

extern "C" int64_t pg_group_cells(const pg_group* a, const pg_group* b, int32_t sh)
{
    if (!a || !b) return 0;
    int lw, up;
    group_band(a, b, sh, &lw, &up);
    int64_t c = 0;
    for (int m = a->left; m < a->right; ++m) {
        int n0 = std::max(m + lw, b->left), n9 = std::min(m + up + 1, b->right);
        if (n9 > n0) c += n9 - n0;
    }
    return c;
}

namespace {
struct SideOff { size_t cfq, efq, vec, glen, gfreq, sfq, tfq, rfq, gapmask, weight, blk; int cs, extra; };

// cs > 0: also the fixed-stride column blocks of the register-list kernels (k3r_core.cuh), cs entries per list
// extra: columns staged beyond right - 1 (1 for side b of a rectangle pair: forwardA reads position b.right)
size_t place_side(const pg_group& g, int kdim, size_t off, SideOff* so, int cs, int extra = 0)
{
    const size_t npos = (size_t)(g.right - g.left + 1 + extra);
    so->extra = extra;
    const bool lists = g.sfq && g.tfq && g.rfq && g.glen && g.gfreq && g.npool > 0;
    so->cfq = off; off = up16(off + 8 * npos);
    so->efq = off; off = up16(off + 8 * npos);
    so->vec = off; off = up16(off + 8 * npos * kdim);
    so->gfreq = off; off = up16(off + 8 * (lists ? (size_t)g.npool : 1));
    so->glen = off; off = up16(off + 4 * (lists ? (size_t)g.npool : 1));
    so->sfq = off; off = up16(off + 4 * npos);
    so->tfq = off; off = up16(off + 4 * npos);
    so->rfq = off; off = up16(off + 4 * npos);
    so->gapmask = off; off = up16(off + 4 * npos);
    so->weight = off; off = up16(off + 8 * (size_t)std::max(g.many, 1));
    so->cs = cs;
    so->blk = off;
    if (cs > 0) off = up16(off + 4 * npos * (size_t)k3r_block_words(cs));
    return off;
}

void fill_side(const pg_group& g, int kdim, const SideOff& so, char* h)
{
    const size_t npos = (size_t)(g.right - g.left + 1 + so.extra);
    const bool lists = g.sfq && g.tfq && g.rfq && g.glen && g.gfreq && g.npool > 0;
    memcpy(h + so.cfq, g.cfq, 8 * npos);
    memcpy(h + so.efq, g.efq, 8 * npos);
    memcpy(h + so.vec, g.vec, 8 * npos * kdim);
    if (lists) {
        memcpy(h + so.gfreq, g.gfreq, 8 * (size_t)g.npool);
        memcpy(h + so.glen, g.glen, 4 * (size_t)g.npool);
        memcpy(h + so.sfq, g.sfq, 4 * npos);
        memcpy(h + so.tfq, g.tfq, 4 * npos);
        memcpy(h + so.rfq, g.rfq, 4 * npos);
    } else {
        *(double*)(h + so.gfreq) = 0;
        *(int32_t*)(h + so.glen) = -1;
        memset(h + so.sfq, 0xff, 4 * npos);
        memset(h + so.tfq, 0xff, 4 * npos);
        memset(h + so.rfq, 0xff, 4 * npos);
    }
    if (g.gapmask) memcpy(h + so.gapmask, g.gapmask, 4 * npos); else memset(h + so.gapmask, 0, 4 * npos);
    for (int i = 0; i < std::max(g.many, 1); ++i) ((double*)(h + so.weight))[i] = (g.weight && i < g.many) ? g.weight[i] : 1.0;
    if (so.cs > 0)
        k3r_build_blocks((int*)(h + so.blk), so.cs, (int)npos, g.cfq, g.efq, lists ? g.glen : nullptr, lists ? g.gfreq : nullptr,
                         lists ? g.sfq : nullptr, lists ? g.tfq : nullptr, lists ? g.rfq : nullptr);
}

K3Group dev_side(const pg_group& g, const SideOff& so, const char* d)
{
    K3Group k;
    k.cfq = (const double*)(d + so.cfq); k.efq = (const double*)(d + so.efq);
    k.prof = k.freq = (const double*)(d + so.vec);
    k.glen = (const int32_t*)(d + so.glen); k.gfreq = (const double*)(d + so.gfreq);
    k.sfq = (const int32_t*)(d + so.sfq); k.tfq = (const int32_t*)(d + so.tfq); k.rfq = (const int32_t*)(d + so.rfq);
    k.L = g.right - g.left;
    k.nils = g.nils;
    k.gapmask = (const uint32_t*)(d + so.gapmask);
    k.weight = (const double*)(d + so.weight);
    k.many = g.many;
    k.pad = 0;
    k.blk = so.cs > 0 ? (const int*)(d + so.blk) : nullptr;
    return k;
}
}  // namespace

// Terminal-gap factors of one Aln2b1 pair (initB_ng / lastB_ng, src/fwd2b1.cc:64-143); nullptr = all 1, no relaxation
struct PgTerm {
    double ltg_a, ltg_b, rtg_a, rtg_b;
    int32_t last_c, last_r;
};

static int pg_int_align_groups(pg_context* c, const pg_group* a, const pg_group* b, const pg_gparams* prm, int64_t npairs,
                               double* out_scores, int64_t** out_offs, pg_skl** out_pts, const PgTerm* term,
                               int64_t* out_rr = nullptr, bool score_only = false, bool swg = false);

extern "C" int pg_align_groups(pg_context* c, const pg_group* a, const pg_group* b, const pg_gparams* prm, int64_t npairs,
                               double* out_scores, int64_t** out_offs, pg_skl** out_pts)
{
    return pg_int_align_groups(c, a, b, prm, npairs, out_scores, out_offs, out_pts, nullptr);
}

// HomScoreC (src/fwd2c.h:663-668): the same fill without the path store
extern "C" int pg_score_groups(pg_context* c, const pg_group* a, const pg_group* b, const pg_gparams* prm, int64_t npairs,
                               double* out_scores, int64_t* out_rr)
{
    int64_t* offs = nullptr;
    pg_skl* pts = nullptr;
    const int rc = pg_int_align_groups(c, a, b, prm, npairs, out_scores, &offs, &pts, nullptr, out_rr, true);
    free(offs); free(pts);
    return rc;
}

// swg1stC<SwgDPunit | _hf | _pf | _nv> (src/fwd2c.h:697-701) for algmode.mlt <= 1: Fwd2c::forwardC without secondary
// colonies; per pair the best local score and the box of colony 0 (src/aln.h:150-160)
extern "C" int pg_local_groups(pg_context* c, const pg_group* a, const pg_group* b, const pg_gparams* prm, int64_t npairs,
                               double* out_val, int32_t* out_box)
{
    if (c && npairs > 0 && !out_box) return pg_int_fail(c, PG_ERR_ARG, "pg_local_groups: NULL output");
    int64_t* offs = nullptr;
    pg_skl* pts = nullptr;
    const int rc = pg_int_align_groups(c, a, b, prm, npairs, out_val, &offs, &pts, nullptr, nullptr, false, true);
    if (!rc)
        for (int64_t i = 0; i < npairs; ++i) {      // the kernel returns the box as three "corners"
            const pg_skl* q = pts + offs[i];
            int32_t* o = out_box + 6 * i;
            o[0] = q[0].m; o[1] = q[0].n; o[2] = q[1].m; o[3] = q[1].n; o[4] = q[2].m; o[5] = q[2].n;
        }
    free(offs); free(pts);
    return rc;
}

static int pg_int_align_groups(pg_context* c, const pg_group* a, const pg_group* b, const pg_gparams* prm, int64_t npairs,
                               double* out_scores, int64_t** out_offs, pg_skl** out_pts, const PgTerm* term,
                               int64_t* out_rr, bool score_only, bool swg)
{
    if (!c) return PG_ERR_ARG;
    if (npairs < 0 || !out_offs || !out_pts || (npairs && (!a || !b || !prm || !out_scores)))
        return pg_int_fail(c, PG_ERR_ARG, "pg_align_groups: NULL / bad argument");
    *out_offs = nullptr;
    *out_pts = nullptr;
    if (npairs > 0x3fffffff) return pg_int_fail(c, PG_ERR_ARG, "pg_align_groups: too many pairs in one call");
    int64_t* offs = (int64_t*)malloc(sizeof(int64_t) * (size_t)(npairs + 1));
    if (!offs) return pg_int_fail(c, PG_ERR_ARG, "out of host memory");
    offs[0] = 0;
    if (npairs == 0) { *out_offs = offs; *out_pts = (pg_skl*)malloc(sizeof(pg_skl)); return PG_OK; }
    cudaError_t e = cudaSetDevice(c->device);
    if (e == cudaSuccess) e = pg_int_order_stream(c, c->stream);
    if (e != cudaSuccess) { free(offs); return pg_int_fail(c, PG_ERR_CUDA, cudaGetErrorString(e)); }

    // ---- validate, lay out the blob
    std::vector<SideOff> soa(npairs), sob(npairs);
    std::vector<K3Pair> pairs(npairs);
    std::vector<int64_t> cells(npairs), outoff(npairs + 1);
    size_t blob = 0, arena_words = 0, wave_bytes = 0;
    const bool nopath = score_only || swg;
    // threads per alignment (768 = 3 roles x 256 rows); Smith-Waterman runs the one-thread-per-row kernel
    const int tg_sel = swg ? 256 : k3_pick_tg(npairs, c->sm_count);
    const int tg = tg_sel == 768 ? 256 : tg_sel;            // rows per stripe
    const int ngrp = k3_threads() / tg;                     // alignments in flight per CTA
    int64_t max_cells = 0;
    outoff[0] = 0;
    // ---- register-list kernels (k3r_core.cuh) for the gap-profile record modes, on request (PG_K3_RL=1): one list
    //      capacity per mode and call, the smallest of 4 / 6 / 8 words that holds every dynamic list (hetero + 1 entries
    //      and the terminator) and every static list (CAP - 2 entries) of the mode's pairs; longer lists or sequences
    //      run the list-walking form.  Measured on B200: bit-identical results, warps 27 of 32
    //      lanes wide instead of 19 and barrier stalls halved, but the capacity-wide unrolled merges execute as many
    //      warp instructions per cell as the list walk (~100) -- 384 pairs 29.6 ms against 30.9 ms, 24 pairs 10.2 ms
    //      against 8.8 ms -- so the list-walking kernels stay the default.
    int rl_cap[5] = {0, 0, 0, 0, 0};
    {
        const char* ev = getenv("PG_K3_RL");
        const bool rl_on = tg_sel == 768 && ev && ev[0] == '1' && !swg;
        int need[5] = {0, 0, 0, 0, 0};
        bool ok[5] = {false, rl_on, rl_on, false, false};
        for (int64_t i = 0; i < npairs && rl_on; ++i) {
            const int m5 = (prm[i].alnmode == 7 || prm[i].alnmode == 8) ? 1 : (prm[i].alnmode == 9 ? 2 : -1);
            if (m5 < 0 || !ok[m5]) continue;
            const pg_group* gs[2] = {&a[i], &b[i]};
            for (int sd = 0; sd < 2 && ok[m5]; ++sd) {
                const pg_group& G = *gs[sd];
                if (G.len > 30000 || G.left < 0 || G.right > G.len || G.left >= G.right) { ok[m5] = false; break; }
                int longest = 0, max_glen = 0;
                if (G.sfq && G.tfq && G.rfq && G.glen && G.gfreq && G.npool > 0)
                    k3r_scan_lists(G.right - G.left + 1, G.glen, G.sfq, G.tfq, G.rfq, &longest, &max_glen);
                if (max_glen > 30000) { ok[m5] = false; break; }
                need[m5] = std::max(need[m5], std::max(std::max(G.hetero, 0), longest));
            }
        }
        for (int m5 = 1; m5 <= 2; ++m5)
            if (ok[m5] && need[m5] <= 6) rl_cap[m5] = need[m5] <= 2 ? 4 : (need[m5] <= 4 ? 6 : 8);
    }
    for (int64_t i = 0; i < npairs; ++i) {
        const pg_group& A = a[i];
        const pg_group& B = b[i];
        const pg_gparams& P = prm[i];
        int mode;
        bool rect = false;
        switch (P.alnmode) {
        case 1: mode = 0; rect = true; break;   // NGP_ALN: forwardA over the whole rectangle
        case 6: mode = 0; break;
        case 7: case 8: mode = 1; break;
        case 9: mode = 2; break;
        case 10: mode = 4; break;       // NTV_ALB: DPunit_nv
        case 100: mode = 3; break;      // PG_ALN_B1_NG: Aln2b1 (pg_align_pairs_ng), two single sequences
        default:
            free(offs);
            return pg_int_fail(c, PG_ERR_UNSUPPORTED, "pg_align_groups: alnmode is not one of NGP_ALN / NGP_ALB / HLF_ALB / RHF_ALB / "
                                                      "GPF_ALB / NTV_ALB (the rectangle modes with gap profiles, local and spliced "
                                                      "modes are not built yet)");
        }
        if (swg && (rect || mode == 3)) {
            free(offs);
            return pg_int_fail(c, PG_ERR_UNSUPPORTED, "pg_local_groups: alnmode is not one of NGP_ALB / HLF_ALB / RHF_ALB / GPF_ALB / NTV_ALB");
        }
        if (rect && (score_only || B.left != 0)) {
            free(offs);
            return pg_int_fail(c, score_only ? PG_ERR_UNSUPPORTED : PG_ERR_ARG,
                               score_only ? "pg_score_groups: the rectangle form of HomScoreC (forwardA with island reports) is not built"
                                          : "pg_align_groups: NGP_ALN needs b.left = 0 (forwardA starts its b iterator at position 0, "
                                            "src/fwd2c.h:240) and b's arrays one column longer (position b.right)");
        }
        if (A.left < 0 || A.right > A.len || A.left >= A.right || B.left < 0 || B.right > B.len || B.left >= B.right ||
            !A.cfq || !A.efq || !A.vec || !B.cfq || !B.efq || !B.vec || P.kdim < 1 || P.kdim > 64 || P.Noll < 2 || P.Noll > 3) {
            free(offs);
            return pg_int_fail(c, PG_ERR_ARG, "pg_align_groups: bad group window / missing arrays (align2 handles empty windows "
                                              "with nogap_skl before alignC)");
        }
        if (((mode == 1 || mode == 2) && !(A.sfq && A.tfq && A.rfq)) || (mode == 2 && !(B.sfq && B.tfq && B.rfq))) {
            free(offs);
            return pg_int_fail(c, PG_ERR_ARG, "pg_align_groups: gap-profile lists missing for a half / full profile mode");
        }
        const int rl = rl_cap[mode];
        blob = place_side(A, P.kdim, blob, &soa[i], rl ? rl - 2 : 0);
        blob = place_side(B, P.kdim, blob, &sob[i], rl ? rl - 2 : 0, rect ? 1 : 0);
        K3Prm& kp = pairs[i].prm;
        kp.mode = mode; kp.Noll = P.Noll; kp.codonk1 = P.codonk1; kp.kdim = P.kdim;
        int lw, up;
        group_band(&A, &B, P.sh, &lw, &up);
        const int r0 = B.left - A.left;
        kp.lw = lw - r0; kp.up = up - r0;
        if (rect) { kp.lw = -(A.right - A.left); kp.up = B.right - B.left; }
        kp.capa = std::max(A.hetero, 0) + 3; kp.capb = std::max(B.hetero, 0) + 3;
        if (mode == 4) {
            if (!A.gapmask || !B.gapmask || A.many < 1 || B.many < 1 || A.many > 32 || B.many > 32 || A.nils || B.nils) {
                free(offs);
                return pg_int_fail(c, (A.nils || B.nils) ? PG_ERR_UNSUPPORTED : PG_ERR_ARG,
                                   "pg_align_groups: NTV_ALB needs gap masks, 1..32 members per group and no nil ends "
                                   "(semi-global naive mode is not built yet)");
            }
            kp.capa = (A.many + 1) / 2 + 1; kp.capb = (B.many + 1) / 2 + 1;     // 16-bit run lengths, two per word
        }
        kp.rl = rl; kp.rect = rect ? 1 : 0;
        kp.swg = swg ? 1 : 0; kp.pad3 = 0;
        if (swg) kp.capb += 4;          // the box of a Smith-Waterman record: its last four words (k3s_box)
        if (rl) {               // the records of the register-list form: 4 header words + rl words per list, a multiple of 4
            kp.capa = rl;
            kp.capb = mode == 2 ? rl : k3r_rec_words(rl, 1) - 4 - rl;
        }
        kp.u = (double)(float)P.u;
        kp.wgop = P.Weighted_GOP; kp.bgop = P.Basic_GOP;
        kp.u2divu1 = P.BasicGEP < 0 ? P.LongGEP / P.BasicGEP : 0;      // fwd2c.h:85-86
        kp.v2divv1 = P.BasicGOP < 0 ? P.LongGOP / P.BasicGOP : 0;
        kp.gop1 = P.BasicGOP; kp.gep1 = P.BasicGEP; kp.gop2 = P.LongGOP; kp.gep2 = P.LongGEP;
        kp.ltg_a = kp.ltg_b = kp.rtg_a = kp.rtg_b = 1.0;
        kp.last_c = kp.last_r = 0;
        kp.novmf = nopath ? 1 : 0;
        kp.origin_r = B.left - A.left;
        if (score_only && mode == 3) {
            free(offs);
            return pg_int_fail(c, PG_ERR_ARG, "pg_score_groups: the Aln2b1 recurrence has its own entry (pg_align_pairs_ng)");
        }
        if (term && mode == 3) {
            kp.ltg_a = term[i].ltg_a; kp.ltg_b = term[i].ltg_b; kp.rtg_a = term[i].rtg_a; kp.rtg_b = term[i].rtg_b;
            kp.last_c = term[i].last_c; kp.last_r = term[i].last_r;
        }
        const int LQ = A.right - A.left, LS = B.right - B.left;
        const size_t st = (size_t)k3_stride(kp.capa, kp.capb);
        arena_words = std::max(arena_words, (st * (size_t)(4 * (LS + 2) + 2 * (LQ + 2)) + k3_wave_words((int)st, 3, tg) + 3) & ~(size_t)3);
        wave_bytes = std::max(wave_bytes, 4 * k3_wave_words((int)st, P.Noll, tg));
        if (A.len > 65000 || B.len > 65000) {
            free(offs);
            return pg_int_fail(c, PG_ERR_RANGE, "pg_align_groups: groups longer than 65,000 columns exceed the 16-bit gap-state lists");
        }
        cells[i] = rect ? (int64_t)(A.right - A.left) * (B.right - B.left) : pg_group_cells(&A, &B, P.sh);
        max_cells = std::max(max_cells, cells[i]);
        pairs[i].al = A.left; pairs[i].bl = B.left;
        pairs[i].simmat = nullptr;
        pairs[i].out_cap = LQ + LS + 4;
        pairs[i].out_off = outoff[i];
        pairs[i].pad = 0;
        outoff[i + 1] = outoff[i] + pairs[i].out_cap;
    }
    const int grid = (int)std::min<int64_t>((npairs + ngrp - 1) / ngrp,
                                            (int64_t)c->sm_count * (tg_sel == 768 ? 1 : k3_blocks_per_sm()));
    // arenas / path stores: the kernels of the record modes present run side by side, each on its own slots
    int64_t n_mode[5] = {0, 0, 0, 0, 0};
    int g_mode[5], slot0_mode[5], nc_mode[5], vslot0_mode[5], maxlq_mode[5] = {0, 0, 0, 0, 0};
    for (int64_t i = 0; i < npairs; ++i) {
        ++n_mode[pairs[i].prm.mode];
        maxlq_mode[pairs[i].prm.mode] = std::max(maxlq_mode[pairs[i].prm.mode], a[i].right - a[i].left);
    }
    // a thread-block cluster of 2 / 4 / 8 CTAs per alignment (192 rows each) when the longest group of the mode
    // needs more than one 256-row stripe and the SMs can hold all clusters of the batch at once (otherwise
    // alignments side by side use the SMs better).  Measured with the decoupled ring hand-over: ~650-row groups
    // 12.0 -> 8.8 ms (24 pairs), ~1,100 rows 32 -> 14.5 ms, 300-450 rows unchanged (prrn5 on 80 x 300 aa).
    // PG_K3_CLUSTER = 1 turns it off, 2 / 4 / 8 cap the size.
    {
        const char* ce = getenv("PG_K3_CLUSTER");
        const int cmax = tg_sel == 768 ? (ce ? atoi(ce) : 8) : 1;
        int64_t ctas = 0;
        for (int m5 = 0; m5 < 5; ++m5) {
            int nc = 1;
            const int rows1 = rl_cap[m5] ? k3_rl_rows() : k3_threads();         // rows of a single CTA / of a cluster CTA
            const int rowsc = rl_cap[m5] ? k3_rl_rows() : k3_cluster_rows();
            if (maxlq_mode[m5] > rows1)
                while (nc < cmax && nc < 8 && maxlq_mode[m5] > nc * rowsc) nc *= 2;
            nc_mode[m5] = n_mode[m5] ? nc : 1;
            ctas += n_mode[m5] * nc_mode[m5] * (rl_cap[m5] ? 1 : 2);            // in half SMs: the register-list CTAs pair up
        }
        while (ctas > 2 * c->sm_count) {        // more clusters than SMs: halve the widest ones
            int w = 0;
            for (int m5 = 1; m5 < 5; ++m5) if (nc_mode[m5] > nc_mode[w]) w = m5;
            if (nc_mode[w] == 1) break;
            ctas -= n_mode[w] * (nc_mode[w] / 2) * (rl_cap[w] ? 1 : 2);
            nc_mode[w] /= 2;
        }
    }
    size_t slots = 0, vslots = 0;
    for (int m5 = 0; m5 < 5; ++m5) {
        g_mode[m5] = rl_cap[m5] ? (int)std::min<int64_t>(n_mode[m5], (int64_t)2 * c->sm_count / nc_mode[m5] > 0 ? (int64_t)2 * c->sm_count / nc_mode[m5] : 1)
                                : (int)std::min<int64_t>((n_mode[m5] + ngrp - 1) / ngrp, grid);
        slot0_mode[m5] = (int)slots;
        vslot0_mode[m5] = (int)vslots;
        slots += (size_t)g_mode[m5] * ngrp;
        vslots += (size_t)g_mode[m5] * ngrp * nc_mode[m5];      // path store: one part per CTA of a cluster
    }
    const int64_t vmf_cap = nopath ? 8 : max_cells + 8;
    if (vmf_cap * 8 > 0x7fffffff) { free(offs); return pg_int_fail(c, PG_ERR_RANGE, "pg_align_groups: DP matrix too large for the path store"); }

    // ---- stage
    int rc = pg_int_ensure_cap(c, &c->d_garena, &c->garena_cap, arena_words * 4 * slots);
    if (!rc) rc = pg_int_ensure_cap(c, &c->d_gvmf, &c->gvmf_cap, sizeof(K3Vmf) * (size_t)vmf_cap * vslots);
    const size_t o_pts = 0, o_cnt = up16(o_pts + 8 * (size_t)outoff[npairs]), o_scr = up16(o_cnt + 4 * (size_t)npairs),
                 obytes = up16(o_scr + 8 * (size_t)npairs);
    if (!rc) rc = pg_int_ensure_cap(c, &c->d_gout, &c->gout_cap, obytes);
    if (rc) { free(offs); return rc; }
    // ---- column score matrices S = X_a . Y_b^T (kernel K4, FP64 tensor cores) when they fit the budget;
    //      otherwise (or with PG_K3_INLINE_SIM=1) K3 evaluates the contraction per cell
    std::vector<size_t> sim_off(npairs, (size_t)-1);
    size_t sim_bytes = 0;
    {
        const char* inl = getenv("PG_K3_INLINE_SIM");
        const size_t SIM_BUDGET = (size_t)24 << 30;
        if (!(inl && inl[0] == '1')) {
            size_t tot = 0;
            for (int64_t i = 0; i < npairs; ++i)
                tot += up16(8 * (size_t)(a[i].right - a[i].left) * (size_t)(b[i].right - b[i].left));
            if (tot <= SIM_BUDGET) {
                for (int64_t i = 0; i < npairs; ++i) {
                    sim_off[i] = sim_bytes;
                    sim_bytes += up16(8 * (size_t)(a[i].right - a[i].left) * (size_t)(b[i].right - b[i].left));
                }
            }
        }
    }
    if (sim_bytes) { rc = pg_int_ensure_cap(c, &c->d_gsim, &c->gsim_cap, sim_bytes); if (rc) { free(offs); return rc; } }
    const size_t o_boff = up16(blob + sizeof(K3Pair) * (size_t)npairs);
    rc = pg_int_ensure_cap(c, &c->d_gblob, &c->gblob_cap, o_boff + 4 * (size_t)(npairs + 1) + 256);
    if (rc) { free(offs); return rc; }
    // host staging in pinned memory owned by the context: one async copy at PCIe rate, no zero fill
    const size_t h_bytes = o_boff + 4 * (size_t)(npairs + 1);
    if (c->gstage_cap < h_bytes) {
        if (c->h_gstage) cudaFreeHost(c->h_gstage);
        c->h_gstage = nullptr; c->gstage_cap = 0;
        const size_t want = h_bytes + h_bytes / 4 + 4096;
        if (cudaHostAlloc(&c->h_gstage, want, cudaHostAllocDefault) != cudaSuccess) {
            free(offs);
            return pg_int_fail(c, PG_ERR_CUDA, "pg_align_groups: cudaHostAlloc(staging) failed");
        }
        c->gstage_cap = want;
    }
    struct { char* p; size_t n; char* data() const { return p; } size_t size() const { return n; } } h = {(char*)c->h_gstage, h_bytes};
    char* d = (char*)c->d_gblob;
    // heaviest pairs first (persistent CTAs finish together); the kernel indexes results by this order
    std::vector<int32_t> order(npairs);
    std::iota(order.begin(), order.end(), 0);
    // ... grouped by record mode: one launch per mode present, each with its own queue
    std::stable_sort(order.begin(), order.end(), [&](int32_t x, int32_t y) {
        return pairs[x].prm.mode != pairs[y].prm.mode ? pairs[x].prm.mode < pairs[y].prm.mode : cells[x] > cells[y];
    });
    std::vector<K3Pair> sorted(npairs);
    std::atomic<int> up_err(0);
    {   // staging copies of the pairs are independent: large batches are filled by up to 4 host threads
        // ... and every finished piece (about 8 MB) goes to the device at once: the upload of the blob runs under the
        //     staging of the rest instead of after it (the pieces are disjoint byte ranges of one pinned buffer)
        auto fill_range = [&](int64_t i0, int64_t i1) {
            cudaSetDevice(c->device);
            size_t sent = soa[i0].cfq;
            for (int64_t i = i0; i < i1; ++i) {
                fill_side(a[i], prm[i].kdim, soa[i], h.data());
                fill_side(b[i], prm[i].kdim, sob[i], h.data());
                const size_t end = i + 1 < npairs ? soa[i + 1].cfq : blob;
                if (end - sent >= ((size_t)8 << 20) || i + 1 == i1) {
                    if (end > sent && cudaMemcpyAsync(d + sent, h.data() + sent, end - sent, cudaMemcpyHostToDevice, c->stream) != cudaSuccess)
                        up_err.store(1);
                    sent = end;
                }
            }
        };
        static const int nth_max = getenv("PG_STAGE_THREADS") ? std::max(1, atoi(getenv("PG_STAGE_THREADS"))) : 4;
        int nth = h_bytes < ((size_t)8 << 20) ? 1 : (int)std::min<int64_t>(nth_max, std::max(1u, std::thread::hardware_concurrency()));
        nth = (int)std::min<int64_t>(nth, npairs);
        if (nth <= 1) fill_range(0, npairs);
        else {
            std::vector<std::thread> th;
            for (int k = 0; k < nth; ++k) th.emplace_back(fill_range, npairs * k / nth, npairs * (k + 1) / nth);
            for (auto& t : th) t.join();
        }
    }
    for (int64_t i = 0; i < npairs; ++i) {
        pairs[i].a = dev_side(a[i], soa[i], d);
        pairs[i].b = dev_side(b[i], sob[i], d);
        pairs[i].simmat = sim_off[i] == (size_t)-1 ? nullptr : (const double*)((char*)c->d_gsim + sim_off[i]);
    }
    for (int64_t k = 0; k < npairs; ++k) sorted[k] = pairs[order[k]];
    memcpy(h.data() + blob, sorted.data(), sizeof(K3Pair) * (size_t)npairs);
    int32_t* boff = (int32_t*)(h.data() + o_boff);
    boff[0] = 0;
    for (int64_t k = 0; k < npairs; ++k) boff[k + 1] = boff[k] + (sorted[k].simmat ? (sorted[k].a.L + 7) / 8 : 0);
    K4Args k4;
    k4.pairs = (const K3Pair*)(d + blob);
    k4.npairs = (int32_t)npairs;
    k4.block_off = (const int32_t*)(d + o_boff);
    const int k4_blocks = boff[npairs];
    K3Args ka;
    memset(&ka, 0, sizeof(ka));
    ka.pairs = (const K3Pair*)(d + blob);
    ka.npairs = (int32_t)npairs;
    ka.counter = c->d_counter;
    ka.arena = (int*)c->d_garena;
    ka.arena_words = (int64_t)arena_words;
    ka.vmf = (K3Vmf*)c->d_gvmf;
    ka.vmf_cap = (int32_t)vmf_cap;
    char* go = (char*)c->d_gout;
    ka.out_pts = (int32_t*)(go + o_pts);
    ka.out_cnt = (int32_t*)(go + o_cnt);
    ka.out_score = (double*)(go + o_scr);
    // wavefront records live in shared memory when they fit (two CTAs per SM up to ~110 KB each, one up to
    // 220 KB); pairs whose records are larger fall back to the L2-resident arena inside the kernel
    {
        const size_t unit = 16 * (size_t)ngrp, cap = (size_t)220 * 1024 / unit * unit;
        ka.smem_bytes = (int32_t)std::min((wave_bytes * ngrp + unit - 1) / unit * unit, cap);
    }
    std::vector<int32_t> h_pts(2 * (size_t)outoff[npairs]), h_cnt(npairs);
    std::vector<double> h_scr(npairs);
    // the group data went up piece by piece while it was staged; what is left is the pair table and the K4 block offsets
    e = up_err.load() ? cudaErrorUnknown
                      : cudaMemcpyAsync(d + blob, h.data() + blob, h.size() - blob, cudaMemcpyHostToDevice, c->stream);
    if (e == cudaSuccess) e = cudaMemsetAsync(c->d_counter, 0, 8 * sizeof(int32_t), c->stream);
    if (e == cudaSuccess) e = cudaEventRecord(c->ev0, c->stream);
    if (e == cudaSuccess && sim_bytes) e = k4_launch(k4, k4_blocks, c->stream);
    // one kernel per record mode present, side by side on auxiliary streams (fork / join around them)
    if (e == cudaSuccess) e = cudaEventRecord(c->ev_fork, c->stream);
    int used[5] = {0, 0, 0, 0, 0};
    for (int64_t k0 = 0; k0 < npairs && e == cudaSuccess;) {
        const int mode = sorted[k0].prm.mode;
        int64_t k1 = k0;
        while (k1 < npairs && sorted[k1].prm.mode == mode) ++k1;
        K3Args km = ka;
        km.pairs = ka.pairs + k0;
        km.npairs = (int32_t)(k1 - k0);
        km.counter = c->d_counter + 1 + mode;
        km.out_cnt = ka.out_cnt + k0;
        km.out_score = ka.out_score + k0;
        const int gm = g_mode[mode];
        km.all_sm = 1;
        for (int64_t k = k0; k < k1 && km.all_sm; ++k)
            km.all_sm = k3_sm_fits(k3_stride(sorted[k].prm.capa, sorted[k].prm.capb), sorted[k].prm.Noll, tg_sel,
                                   (size_t)ka.smem_bytes) ? 1 : 0;
        km.cluster = nc_mode[mode];
        km.rl = rl_cap[mode];
        km.swg = swg ? 1 : 0;
        if (km.rl) {                            // register-list kernels: 128 rows per CTA, their own shared-memory size
            size_t wb = 0;
            for (int64_t k = k0; k < k1; ++k)
                wb = std::max(wb, 4 * k3_wave_words(k3_stride(sorted[k].prm.capa, sorted[k].prm.capb), sorted[k].prm.Noll, k3_rl_rows()));
            km.smem_bytes = (int32_t)((wb + 15) / 16 * 16);
            km.all_sm = 1;
        }
        if (!km.all_sm && tg_sel == 768) {      // long records (high hetero, two-piece): 192 rows per CTA may still fit
            km.rows192 = 1;
            for (int64_t k = k0; k < k1 && km.rows192; ++k)
                km.rows192 = k3_sm_fits_rows(k3_stride(sorted[k].prm.capa, sorted[k].prm.capb), sorted[k].prm.Noll,
                                             k3_cluster_rows(), (size_t)ka.smem_bytes) ? 1 : 0;
            km.all_sm = km.rows192;
        }
        km.arena = ka.arena + (size_t)slot0_mode[mode] * arena_words;
        km.vmf = ka.vmf + (size_t)vslot0_mode[mode] * (size_t)vmf_cap;
        if (getenv("PG_K3_DEBUG"))
            fprintf(stderr, "pg_align_groups: mode %d pairs %d rl %d cluster %d units %d smem %d all_sm %d rows192 %d\n", mode, km.npairs,
                    km.rl, km.cluster, gm, km.smem_bytes, km.all_sm, km.rows192);
        e = cudaStreamWaitEvent(c->aux[mode], c->ev_fork, 0);
        if (e == cudaSuccess) e = k3_launch(km, tg_sel, mode, gm, c->aux[mode]);
        if (e == cudaSuccess) e = cudaEventRecord(c->ev_join[mode], c->aux[mode]);
        used[mode] = 1;
        k0 = k1;
    }
    for (int m5 = 0; m5 < 5 && e == cudaSuccess; ++m5)
        if (used[m5]) e = cudaStreamWaitEvent(c->stream, c->ev_join[m5], 0);
    if (e == cudaSuccess) e = cudaEventRecord(c->ev1, c->stream);
    c->ev_valid = e == cudaSuccess;
    if (e == cudaSuccess) e = cudaMemcpyAsync(h_pts.data(), ka.out_pts, 8 * (size_t)outoff[npairs], cudaMemcpyDeviceToHost, c->stream);
    if (e == cudaSuccess) e = cudaMemcpyAsync(h_cnt.data(), ka.out_cnt, 4 * (size_t)npairs, cudaMemcpyDeviceToHost, c->stream);
    if (e == cudaSuccess) e = cudaMemcpyAsync(h_scr.data(), ka.out_score, 8 * (size_t)npairs, cudaMemcpyDeviceToHost, c->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
    if (e != cudaSuccess) { free(offs); return pg_int_fail(c, PG_ERR_CUDA, (std::string("pg_align_groups: ") + cudaGetErrorString(e)).c_str()); }
    // ---- back to the caller's order
    std::vector<int32_t> cnt_by_orig(npairs);
    for (int64_t k = 0; k < npairs; ++k) {
        if (h_cnt[k] < 0) { free(offs); return pg_int_fail(c, PG_ERR_RANGE, "pg_align_groups: path record store overflow"); }
        cnt_by_orig[order[k]] = h_cnt[k];
    }
    for (int64_t i = 0; i < npairs; ++i) offs[i + 1] = offs[i] + cnt_by_orig[i];
    pg_skl* pts = (pg_skl*)malloc(sizeof(pg_skl) * (size_t)std::max<int64_t>(offs[npairs], 1));
    if (!pts) { free(offs); return pg_int_fail(c, PG_ERR_ARG, "out of host memory"); }
    for (int64_t k = 0; k < npairs; ++k) {
        const int64_t i = order[k];
        const int32_t* src = h_pts.data() + 2 * pairs[i].out_off;
        for (int q = 0; q < h_cnt[k]; ++q) { pts[offs[i] + q].m = src[2 * q]; pts[offs[i] + q].n = src[2 * q + 1]; }
        out_scores[i] = h_scr[k];
        if (score_only && out_rr) {         // pp[2] of forwardB (fwd2c.h:476-479)
            out_rr[2 * i] = src[0];
            out_rr[2 * i + 1] = (int64_t)(b[i].left - a[i].left) + (b[i].right - a[i].right);
        }
    }
    *out_offs = offs;
    *out_pts = pts;
    return PG_OK;
}


// ---- alignC<DPunit> for two single sequences on the floating-point path --------------------------
// pg_align_pairs sends here what the exact-integer bit kernel (K2) does not take: non-integral
// matrices / penalties (the default PAM tables) and the two-piece gap function (alprm.ls == 3,
// fwd2c.h:411-442).  A single sequence is a group of one: thickness 1 on every column (unit_dns,
// mseq.h:89), no gap profile, sim2 = sim11 = mtx[a][b] -> X_a[m] = matrix row of a's residue,
// Y_b[n] = one-hot of b's residue; constants as PwdB::PwdB / PwdM::resetuab derive them in the caller's
// VTYPE (aln2.cc:97-117, maln2.cc:227-243).  Runs kernel K3 in its DPunit mode.
namespace {
template <typename VT>
void single_gparams(const pg_params* prm, int dim, pg_gparams* gp)
{
    const float fu = prm->alprm.u, fv = prm->alprm.v, fu1 = prm->alprm.u1, fsc = prm->alprm.scale;
    const VT Vab = (VT)(fsc * 1 * 1);
    const VT BasicGOP = (VT)(-fv * Vab), BasicGEP = (VT)(-fu * Vab), LongGEP = (VT)(-fu1 * Vab);
    const VT diffu = LongGEP - BasicGEP;
    const VT LongGOP = BasicGOP - diffu * prm->alprm.k1;
    gp->alnmode = 6;    // NGP_ALB
    gp->Noll = prm->alprm.ls < 2 ? 2 : (prm->alprm.ls > 3 ? 3 : prm->alprm.ls);
    gp->codonk1 = prm->alprm.ls == 3 ? prm->alprm.k1 : (0x7fffffff / 8 * 7);
    gp->sh = prm->alprm.sh;
    gp->kdim = dim;
    gp->u = fu;
    gp->Weighted_GOP = (double)(VT)-fv;
    gp->Basic_GOP = (double)(VT)(-fsc * fv);
    gp->BasicGOP = (double)BasicGOP; gp->BasicGEP = (double)BasicGEP;
    gp->LongGOP = (double)LongGOP; gp->LongGEP = (double)LongGEP;
}
}  // namespace

int pg_int_align_pairs_fp(pg_context* c, const pg_seqs* s, const int32_t* a_idx, const int32_t* b_idx, int64_t npairs,
                          const pg_params* prm, const void* mtx, int32_t dim, void* out_scores, int64_t** out_offs,
                          pg_skl** out_pts, int b1)
{
    // Fwd2c<DPunit> on nil-ended sequences is not built; Aln2b1 (b1) takes tgapf < 1 and free ends (initB_ng /
    // lastB_ng), but not the Smith-Waterman branch (lcl & 16: fwdswgB_ng is another function)
    if (b1 ? (prm->lcl & 16) != 0 : (prm->lcl != 0 || !(prm->alprm.tgapf == 1.0f)))
        return pg_int_fail(c, PG_ERR_UNSUPPORTED, "pg_align_pairs: semi-global / local alignment with path (nil-ended sequences, "
                                                  "thickness at the ends) is not built yet");
    pg_gparams gp;
    if (prm->vtype) single_gparams<double>(prm, dim, &gp); else single_gparams<float>(prm, dim, &gp);
    if (b1) gp.alnmode = 100;
    // per sequence and role: ones, matrix rows (as a), one-hot rows (as b)
    const int n = s->nseq;
    std::vector<std::vector<double>> rowv(n), onehot(n), ones(n);
    std::vector<char> need_a(n, 0), need_b(n, 0);
    for (int64_t p = 0; p < npairs; ++p) { need_a[a_idx[p]] = 1; need_b[b_idx[p]] = 1; }
    auto wl = [&](int i) { return std::make_pair(s->left ? s->left[i] : 0, s->right ? s->right[i] : s->lens[i]); };
    for (int i = 0; i < n; ++i) {
        if (!need_a[i] && !need_b[i]) continue;
        if (!b1 && s->exg && (s->exg[i] & 3))
            return pg_int_fail(c, PG_ERR_UNSUPPORTED, "pg_align_pairs: inex.exgl / exgr with path is not built yet");
        const auto w = wl(i);
        if (w.first < 0 || w.second > s->lens[i] || w.first > w.second) return pg_int_fail(c, PG_ERR_ARG, "window outside the sequence");
        const size_t npos = (size_t)(w.second - w.first + 1);
        ones[i].assign(npos, 1.0);
        const uint8_t* r = s->res + s->offs[i] + w.first;
        for (size_t x = 1; x < npos; ++x)
            if (r[x - 1] >= dim) return pg_int_fail(c, PG_ERR_ARG, "residue code outside the substitution matrix");
        if (need_a[i]) {
            rowv[i].assign(npos * dim, 0.0);
            for (size_t x = 1; x < npos; ++x)
                for (int k = 0; k < dim; ++k) {
                    const double v = prm->vtype ? ((const double*)mtx)[r[x - 1] * dim + k] : (double)((const float*)mtx)[r[x - 1] * dim + k];
                    rowv[i][x * dim + k] = v == v ? v : 0.0;    // uninitialised (NaN) entries of unused codes
                }
        }
        if (need_b[i]) {
            onehot[i].assign(npos * dim, 0.0);
            for (size_t x = 1; x < npos; ++x) onehot[i][x * dim + r[x - 1]] = 1.0;
        }
    }
    int64_t* offs = (int64_t*)malloc(sizeof(int64_t) * (size_t)(npairs + 1));
    if (!offs) return pg_int_fail(c, PG_ERR_ARG, "out of host memory");
    offs[0] = 0;
    std::vector<pg_skl> all;
    const int64_t CHUNK = 2048;
    std::vector<pg_group> ga, gb;
    std::vector<pg_gparams> gps;
    std::vector<PgTerm> terms;
    std::vector<double> sc;
    for (int64_t c0 = 0; c0 < npairs; c0 += CHUNK) {
        const int64_t c1 = std::min(npairs, c0 + CHUNK);
        ga.clear(); gb.clear(); gps.clear(); terms.clear();
        std::vector<int64_t> live;      // pairs with two non-empty windows (align2 answers the others with nogap_skl)
        for (int64_t p = c0; p < c1; ++p) {
            const int ia = a_idx[p], ib = b_idx[p];
            const auto wa = wl(ia), wb = wl(ib);
            if (wa.first == wa.second || wb.first == wb.second) continue;
            pg_group A, B;
            memset(&A, 0, sizeof(A)); memset(&B, 0, sizeof(B));
            A.many = 1; A.len = s->lens[ia]; A.left = wa.first; A.right = wa.second; A.hetero = -1;
            A.cfq = A.efq = ones[ia].data(); A.vec = rowv[ia].data();
            B.many = 1; B.len = s->lens[ib]; B.left = wb.first; B.right = wb.second; B.hetero = -1;
            B.cfq = B.efq = ones[ib].data(); B.vec = onehot[ib].data();
            ga.push_back(A); gb.push_back(B); gps.push_back(gp);
            if (b1) {
                // inex flags of the two roles: the caller's per-sequence flags, or algmode.lcl as aln applies it
                // (exg_seq(lcl&1, lcl&2) on a, exg_seq(lcl&4, lcl&8) on b)
                const int ea = (s->exg ? s->exg[ia] & 3 : 0) | (prm->lcl & 3);
                const int eb = (s->exg ? s->exg[ib] & 3 : 0) | ((prm->lcl >> 2) & 3);
                const double tg = (double)prm->alprm.tgapf;
                PgTerm t;
                t.ltg_a = wa.first ? 1.0 : ((ea & 1) ? 0.0 : tg);
                t.ltg_b = wb.first ? 1.0 : ((eb & 1) ? 0.0 : tg);
                t.rtg_a = (ea & 2) ? 0.0 : tg;
                t.rtg_b = (eb & 2) ? 0.0 : tg;
                t.last_c = wb.second == s->lens[ib] && t.rtg_b < 1.0;
                t.last_r = wa.second == s->lens[ia] && t.rtg_a < 1.0;
                terms.push_back(t);
            }
            live.push_back(p);
        }
        sc.assign(live.size(), 0.0);
        int64_t* o2 = nullptr;
        pg_skl* p2 = nullptr;
        if (!live.empty()) {
            int rc = pg_int_align_groups(c, ga.data(), gb.data(), gps.data(), (int64_t)live.size(), sc.data(), &o2, &p2,
                                         b1 ? terms.data() : nullptr);
            if (rc) { free(offs); return rc; }
        }
        size_t li = 0;
        for (int64_t p = c0; p < c1; ++p) {
            double score = 0;
            if (li < live.size() && live[li] == p) {
                for (int64_t q = o2[li]; q < o2[li + 1]; ++q) all.push_back(p2[q]);
                score = sc[li];
                ++li;
            } else {            // nogap_skl (aln2.cc:30-40) in back-walk order; score untouched (0)
                const auto wa = wl(a_idx[p]), wb = wl(b_idx[p]);
                pg_skl e1 = {wa.second, wb.second}, e0 = {wa.first, wb.first};
                all.push_back(e1); all.push_back(e0);
            }
            offs[p + 1] = (int64_t)all.size();
            if (prm->vtype) ((double*)out_scores)[p] = score; else ((float*)out_scores)[p] = (float)score;
        }
        pg_free(o2); pg_free(p2);
    }
    pg_skl* pts = (pg_skl*)malloc(sizeof(pg_skl) * std::max<size_t>(all.size(), 1));
    if (!pts) { free(offs); return pg_int_fail(c, PG_ERR_ARG, "out of host memory"); }
    memcpy(pts, all.data(), sizeof(pg_skl) * all.size());
    *out_offs = offs;
    *out_pts = pts;
    return PG_OK;
}


// ---- Aln2b1: alignB_ng / HomScoreB_ng ------------------------------------------------------------
extern "C" int pg_align_pairs_ng(pg_context* c, const pg_seqs* s, const int32_t* a_idx, const int32_t* b_idx, int64_t npairs,
                                 const pg_params* prm, const void* mtx, int32_t dim, void* out_scores, int64_t** out_offs,
                                 pg_skl** out_pts)
{
    if (!c) return PG_ERR_ARG;
    if (!s || !prm || !mtx || npairs < 0 || !out_offs || !out_pts || (npairs && (!a_idx || !b_idx || !out_scores)))
        return pg_int_fail(c, PG_ERR_ARG, "pg_align_pairs_ng: NULL / bad argument");
    if (dim < 1 || dim > 32) return pg_int_fail(c, PG_ERR_ARG, "dim must be in [1, 32]");
    for (int64_t p = 0; p < npairs; ++p)
        if (a_idx[p] < 0 || a_idx[p] >= s->nseq || b_idx[p] < 0 || b_idx[p] >= s->nseq)
            return pg_int_fail(c, PG_ERR_ARG, "pg_align_pairs_ng: sequence index out of range");
    return pg_int_align_pairs_fp(c, s, a_idx, b_idx, npairs, prm, mtx, dim, out_scores, out_offs, out_pts, 1);
}
