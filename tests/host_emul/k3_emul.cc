// Host emulation of kernel K3's CTA (tests only): the anti-diagonal wavefront of k3_groups.cu with
// threads replaced by a loop and barriers by loop boundaries, calling the very per-cell code the kernel
// runs (k3_core.cuh), so that slot rotation, band guards, pass hand-over, boundary chains and the path
// record store can be checked against the oracle without a GPU.
#include <stdlib.h>
#include <string.h>
#include <vector>
#include "../../prrn_aln_b200/csrc/k3_core.cuh"

extern "C" int k3_emul_align(const K3Group* ga, const K3Group* gb, const K3Prm* prm, int Tsigned, int al, int bl,
                             double* score, int* out_pts, int cap)
{
    const int T = Tsigned < 0 ? -Tsigned : Tsigned;
    const bool rev = Tsigned < 0;       // run the "threads" of a step in reverse order: results must not depend on it
    const K3Group& a = *ga; const K3Group& b = *gb; const K3Prm& p = *prm;
    const int LQ = a.L, LS = b.L;
    const int st = k3_stride(p.capa, p.capb);
    std::vector<int> mem((size_t)st * (3 * (LS + 2) + (LQ + 2) + 9 * T + 2));
    int* base = mem.data();
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
    for (size_t i = 0; i < o; ++i) k3_reset(p, rec(i));
    std::vector<K3Vmf> vmf;
    vmf.push_back({0, 0, 0});                       // skip 0-th record (fwd2c.h:361)
    vmf.push_back({al, bl, 0});                     // origin (initB)
    // origin + boundary chains
    k3_setval(colH, 0); k3_setdg(colH, p.mode == 3 ? K3_NEWD : K3_DIAG, 0); K3_PTR(colH) = 1;
    k3_copy(p, rowH, colH);
    { int rr = LS < p.up ? LS : p.up; for (int k = 1; k <= rr; ++k) { if (p.mode == 3) k3_boundary_b1(p, k, rowH + (size_t)k * st, rowH + (size_t)(k - 1) * st, true); else k3_boundary_row(p, a, b, k, rowH + (size_t)k * st, rowH + (size_t)(k - 1) * st); } }
    { int rr = LQ < -p.lw ? LQ : -p.lw; for (int k = 1; k <= rr; ++k) { if (p.mode == 3) k3_boundary_b1(p, k, colH + (size_t)k * st, colH + (size_t)(k - 1) * st, false); else k3_boundary_col(p, a, b, k, colH + (size_t)k * st, colH + (size_t)(k - 1) * st); } }
    std::vector<double> pua(T, 0.0);
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
                k3_reset(p, F1 + (size_t)t * st);
                k3_reset(p, F2 + (size_t)t * st);
            }
            const bool first_row = m == 0, first_col = n == 0;
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
            const bool rec = p.mode == 3
                ? k3_cell_b1(p, dab, hdiag, habove, gabove, g2above, hleft, F1 + (size_t)t * st, F2 + (size_t)t * st, hout, gout, g2out)
                : k3_cell(p, a, b, ia, ib, first_row, first_col, dab, &pua[t], hdiag, habove, gabove, g2above, hleft,
                          F1 + (size_t)t * st, F2 + (size_t)t * st, hout, gout, g2out, black);
            if (rec) {
                vmf.push_back({m + al, n + bl, K3_PTR(hout)});
                K3_PTR(hout) = (int)vmf.size() - 1;
            }
            if (m == LQ - 1) {
                if (n == LS - 1) { last_ptr = K3_PTR(hout); last_val = k3_val(hout); }
            } else if (t == T - 1) {                    // bottom row of a stripe: park it for thread 0's next row
                k3_copy(p, rowH + (size_t)(n + 1) * st, hout);
                k3_copy(p, rowG + (size_t)(n + 1) * st, gout);
                if (p.Noll == 3) k3_copy(p, rowG2 + (size_t)(n + 1) * st, g2out);
            }
        }
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
