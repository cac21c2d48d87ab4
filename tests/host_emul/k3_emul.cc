// Host emulation of kernel K3's CTA (tests only): the anti-diagonal wavefront of k3_groups.cu with
// threads replaced by a loop and barriers by loop boundaries, calling the very per-cell code the kernel
// runs (k3_core.cuh), so that slot rotation, band guards, pass hand-over, boundary chains and the path
// record store can be checked against the oracle without a GPU.
#include <stdlib.h>
#include <string.h>
#include <vector>
#include "../../prrn_aln_b200/csrc/k3_core.cuh"
#include "../../prrn_aln_b200/csrc/k3r_core.cuh"

// RL = 0: the list-walking cell (k3_core.cuh); 4 / 6 / 8: the register-list form (k3r_core.cuh) with that many words
// per dynamic list, column blocks built here from the pooled lists
template <int RL, int RMODE>
static int emul(const K3Group* ga, const K3Group* gb, const K3Prm* prm, int Tsigned, int al, int bl,
                double* score, int* out_pts, int cap)
{
    constexpr int RCAP = RL ? RL : 4, RCS = RCAP - 2;
    constexpr int BW = k3r_block_words(RCS);
    const int T = Tsigned < 0 ? -Tsigned : Tsigned;
    const bool rev = Tsigned < 0;       // run the "threads" of a step in reverse order: results must not depend on it
    const K3Group& a = *ga; const K3Group& b = *gb;
    K3Prm p = *prm;
    const int LQ = a.L, LS = b.L;
    std::vector<int> ablk, bblk;
    int *ablk_p = nullptr, *bblk_p = nullptr;
    if (RL) {
        p.rl = RL; p.capa = RL; p.capb = RMODE == 2 ? RL : k3r_rec_words(RL, 1) - 4 - RL;
        ablk.resize((size_t)(LQ + 1) * BW + 4); bblk.resize((size_t)(LS + 1) * BW + 4);
        ablk_p = ablk.data(); bblk_p = bblk.data();
        while (((uintptr_t)ablk_p & 15) != 0) ++ablk_p;     // blocks are read 16 bytes at a time
        while (((uintptr_t)bblk_p & 15) != 0) ++bblk_p;
        const int la = k3r_build_blocks(ablk_p, RCS, LQ + 1, a.cfq, a.efq, a.glen, a.gfreq, a.sfq, a.tfq, a.rfq);
        const int lb = k3r_build_blocks(bblk_p, RCS, LS + 1, b.cfq, b.efq, b.glen, b.gfreq, b.sfq, b.tfq, b.rfq);
        if (la > RCS || lb > RCS) return -2;
    }
    const int st = k3_stride(p.capa, p.capb);
#define RESET(ptr) do { if (RL) k3r_reset<RCAP, RMODE>(ptr); else if (p.swg) k3s_blank(p, ptr, K3_NEVSEL); else k3_reset(p, ptr); } while (0)
#define RCOPY(d, s) do { if (RL) k3r_copy_words(d, s, st); else if (p.swg) k3s_copy(p, d, s); else k3_copy(p, d, s); } while (0)
    std::vector<int> mem((size_t)st * (3 * (LS + 2) + (LQ + 2) + 9 * T + 2) + 4);
    int* base = mem.data();
    while (((uintptr_t)base & 15) != 0) ++base;     // records are read and written 16 bytes at a time in the RL form
    auto rec = [&](size_t i) { return base + i * st; };
    size_t o = 0;
    int* rowH = rec(o); o += LS + 2;
    int* rowG = rec(o); o += LS + 2;
    int* rowG2 = rec(o); o += LS + 2;
    int* colH = rec(o); o += LQ + 2;
    int* pubH = rec(o); o += 3 * T;
    int* pubG = rec(o); o += 2 * T;
    int* pubG2 = rec(o); o += 2 * T;
    int* F1 = rec(o); o += T;
    int* F2 = rec(o); o += T;
    int* black = rec(o); o += 1;
    for (size_t i = 0; i < o; ++i) RESET(rec(i));
    std::vector<K3Vmf> vmf;
    vmf.push_back({0, 0, 0});                       // skip 0-th record (fwd2c.h:361)
    vmf.push_back({al, bl, 0});                     // origin (initB)
    // origin + boundary chains
    k3_setval(colH, 0); k3_setdg(colH, p.mode == 3 ? K3_NEWD : ((p.rect || p.swg) ? 0 : K3_DIAG), 0); K3_PTR(colH) = 1;
    const int r0 = bl - al;
    if (p.swg) { int* bx = k3s_box(p, colH); bx[0] = bx[1] = r0; bx[2] = al; bx[3] = bl; }
    RCOPY(rowH, colH);
    { int rr = LS < p.up ? LS : p.up; for (int k = 1; k <= rr; ++k) {
        if (p.swg) {        // initC (fwd2c.h:186-194): blank records on the diagonals of the first row
            int* h = rowH + (size_t)k * st; k3s_blank(p, h, 0); int* bx = k3s_box(p, h); bx[0] = bx[1] = r0 + k; bx[2] = al; bx[3] = bl + k;
        } else if (p.mode == 3) k3_boundary_b1(p, k, rowH + (size_t)k * st, rowH + (size_t)(k - 1) * st, true);
        else if (RL) k3r_boundary_row<RCAP, RCS, RMODE>(p, ablk_p, bblk_p + (size_t)k * BW, k, rowH + (size_t)k * st, rowH + (size_t)(k - 1) * st);
        else k3_boundary_row(p, a, b, k, rowH + (size_t)k * st, rowH + (size_t)(k - 1) * st); } }
    { int rr = LQ < -p.lw ? LQ : -p.lw; for (int k = 1; k <= rr; ++k) {
        if (p.swg) {        // (:196-206): ... and of the first column
            int* h = colH + (size_t)k * st; k3s_blank(p, h, 0); int* bx = k3s_box(p, h); bx[0] = bx[1] = r0 - k; bx[2] = al + k; bx[3] = bl;
        } else if (p.mode == 3) k3_boundary_b1(p, k, colH + (size_t)k * st, colH + (size_t)(k - 1) * st, false);
        else if (RL) k3r_boundary_col<RCAP, RCS, RMODE>(p, ablk_p + (size_t)k * BW, bblk_p, k, colH + (size_t)k * st, colH + (size_t)(k - 1) * st);
        else k3_boundary_col(p, a, b, k, colH + (size_t)k * st, colH + (size_t)(k - 1) * st); } }
    std::vector<double> pua(T, 0.0);
    std::vector<K3Best> best(T, K3Best{0.0, 0, 0, 0, 0, 0, 0});
    int last_ptr = 0; double last_val = 0;
    // Continuous schedule (k3_groups.cu): thread t takes rows t, t+T, t+2T, ...; its k-th row meets column n
    // at global step S = k*P + t + n with the period P = max(LS, T + 4), so a thread starts its next row
    // the step after it finished the previous one and the wavefront never drains between passes.
    const int P = LS > T + 4 ? LS : T + 4;
    const int npass = (LQ + T - 1) / T;
    const int rows_last = LQ - (npass - 1) * T;
    const int total_steps = (npass - 1) * P + (rows_last - 1) + LS;
    for (int S = 0; S < total_steps; ++S) {
        for (int tt = 0; tt < T; ++tt) {                // "threads"; no intra-step dependencies
            const int t = rev ? T - 1 - tt : tt;
            const int q = S - t;
            if (q < 0) continue;
            const int k = q / P, n = q - k * P;
            const int m = k * T + t;
            if (m >= LQ || n >= LS) continue;
            const int r = n - m;
            if (r < p.lw || r > p.up) { continue; }
            const int ia = m + 1, ib = n + 1;           // staged indices (entry 0 = position left-1)
            const bool row_start = n == 0 || r == p.lw; // first in-band column of this row
            if (row_start) {
                pua[t] = k3_unp(a, ia, b, ib, p.u);     // once per row, at its first column (fwd2c.h:377)
                RESET(F1 + (size_t)t * st);
                RESET(F2 + (size_t)t * st);
            }
            const bool first_row = m == 0 && !p.rect, first_col = n == 0 && !p.rect;
            const int* hdiag = n == 0 ? colH + (size_t)m * st
                             : (t == 0 ? rowH + (size_t)n * st : pubH + ((size_t)((S + 1) % 3) * T + (t - 1)) * st);
            const bool above_in = r + 1 <= p.up;
            const int* habove = !above_in ? black : (t == 0 ? rowH + (size_t)(n + 1) * st : pubH + ((size_t)((S + 2) % 3) * T + (t - 1)) * st);
            const int* gabove = (!above_in || m == 0) ? black : (t == 0 ? rowG + (size_t)(n + 1) * st : pubG + ((size_t)((S + 1) & 1) * T + (t - 1)) * st);
            const int* g2above = (!above_in || m == 0) ? black : (t == 0 ? rowG2 + (size_t)(n + 1) * st : pubG2 + ((size_t)((S + 1) & 1) * T + (t - 1)) * st);
            const bool left_in = r - 1 >= p.lw;
            const int* hleft = n == 0 ? colH + (size_t)(m + 1) * st : (left_in ? pubH + ((size_t)((S + 2) % 3) * T + t) * st : black);
            int* hout = pubH + ((size_t)(S % 3) * T + t) * st;
            int* gout = pubG + ((size_t)(S & 1) * T + t) * st;
            int* g2out = pubG2 + ((size_t)(S & 1) * T + t) * st;
            const double dab = k3_sim(a, b, p, ia, ib);
            bool rec;
            if (p.swg) {
                const int ra = r + r0;
                const double diag = k3_val(hdiag);
                k3s_part_diag(p, a, b, ia, ib, ra, dab, hdiag, hout);
                k3s_part_vert(p, a, b, ia, ib, ra, first_row, habove, gabove, g2above, gout, g2out);
                k3s_part_hori(p, a, b, ia, ib, ra, first_col, hleft, F1 + (size_t)t * st, F2 + (size_t)t * st);
                k3s_combine(p, first_row, first_col, m + al, n + bl, diag, hout, gout, g2out, F1 + (size_t)t * st, F2 + (size_t)t * st, &best[t]);
                rec = false;
            } else if (p.mode == 3) rec = k3_cell_b1(p, dab, hdiag, habove, gabove, g2above, hleft, F1 + (size_t)t * st, F2 + (size_t)t * st, hout, gout, g2out);
            else if (RL) rec = k3r_cell<RCAP, RCS, RMODE>(p, ablk_p + (size_t)ia * BW, bblk_p + (size_t)ib * BW, a.nils != 0, first_row, first_col,
                                                          dab, &pua[t], hdiag, habove, gabove, g2above, hleft, F1 + (size_t)t * st,
                                                          F2 + (size_t)t * st, hout, gout, g2out, black, st);
            else rec = k3_cell(p, a, b, ia, ib, first_row, first_col, dab, &pua[t], hdiag, habove, gabove, g2above, hleft,
                               F1 + (size_t)t * st, F2 + (size_t)t * st, hout, gout, g2out, black);
            if (rec) {
                vmf.push_back({m + al, n + bl, K3_PTR(hout)});
                K3_PTR(hout) = (int)vmf.size() - 1;
            }
            if (m == LQ - 1) {
                if (n == LS - 1) { last_ptr = K3_PTR(hout); last_val = k3_val(hout); }
            } else if (t == T - 1) {                    // bottom row of a stripe: park it for thread 0's next row
                RCOPY(rowH + (size_t)(n + 1) * st, hout);
                RCOPY(rowG + (size_t)(n + 1) * st, gout);
                if (p.Noll == 3) RCOPY(rowG2 + (size_t)(n + 1) * st, g2out);
            }
        }
    }
    if (p.swg) {            // colony 0: the first cell in row-major order that holds the maximum
        K3Best c0 = best[0];
        for (int t = 1; t < T; ++t) if (k3s_better(best[t], c0)) c0 = best[t];
        *score = c0.val;
        if (cap < 3) return -1;
        out_pts[0] = c0.mlb; out_pts[1] = c0.nlb; out_pts[2] = c0.mrb; out_pts[3] = c0.nrb; out_pts[4] = c0.lwr; out_pts[5] = c0.upr;
        return 3;
    }
    vmf.push_back({LQ + al, LS + bl, last_ptr});
    *score = last_val;
    int cnt = 0;
    for (int q = (int)vmf.size() - 1;; q = vmf[q].p) {      // Vmf::traceback (vmf.cc:103-119)
        if (cnt >= cap) return -1;
        out_pts[2 * cnt] = vmf[q].m; out_pts[2 * cnt + 1] = vmf[q].n; ++cnt;
        if (!vmf[q].p) break;
    }
    return cnt;
}
#undef RESET
#undef RCOPY

extern "C" int k3_emul_align(const K3Group* ga, const K3Group* gb, const K3Prm* prm, int Tsigned, int al, int bl,
                             double* score, int* out_pts, int cap)
{
    return emul<0, 1>(ga, gb, prm, Tsigned, al, bl, score, out_pts, cap);
}

// register-list form; rl = 4 / 6 / 8 words per dynamic list; record modes 1 (DPunit_hf) and 2 (DPunit_pf) only.
// Returns -2 when a static list is longer than rl - 2 entries.
extern "C" int k3_emul_align_rl(const K3Group* ga, const K3Group* gb, const K3Prm* prm, int Tsigned, int al, int bl,
                                double* score, int* out_pts, int cap, int rl)
{
    const int m = prm->mode;
    if (m != 1 && m != 2) return -3;
    switch (rl) {
    case 4: return m == 1 ? emul<4, 1>(ga, gb, prm, Tsigned, al, bl, score, out_pts, cap) : emul<4, 2>(ga, gb, prm, Tsigned, al, bl, score, out_pts, cap);
    case 6: return m == 1 ? emul<6, 1>(ga, gb, prm, Tsigned, al, bl, score, out_pts, cap) : emul<6, 2>(ga, gb, prm, Tsigned, al, bl, score, out_pts, cap);
    case 8: return m == 1 ? emul<8, 1>(ga, gb, prm, Tsigned, al, bl, score, out_pts, cap) : emul<8, 2>(ga, gb, prm, Tsigned, al, bl, score, out_pts, cap);
    default: return -3;
    }
}
