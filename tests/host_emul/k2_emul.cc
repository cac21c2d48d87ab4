// Host emulation of kernel K2 (tests only): the warp wavefront of k2_core.cuh with shuffles replaced
// by arrays, direction words in the kernel's wavefront layout, then the same k2_trace the device runs.
#include <vector>
#include "../../prrn_aln_b200/csrc/k2_core.cuh"

template <int R>
static int emulate(const uint8_t* q, const uint8_t* s, const K1Geom& g, const int* mtx, int dim, int* out, int* score)
{
    const int T = 32, rpp = T * R, negv = -g.v;
    std::vector<unsigned long long> words((size_t)k2_words_per_pair(g.LQ, g.LS, R) + 64, 0ull);
    std::vector<int> rowH(g.LS + 1), rowG(g.LS + 1);
    int result = 0;
    for (int pass = 0; pass * rpp < g.LQ; ++pass) {
        const int pbase = pass * rpp;
        std::vector<K2Lane<R>> L(T);
        int send_h[2][T], send_g[2][T];
        for (int t = 0; t < T; ++t) k2_lane_init(L[t], g, pbase + t * R);
        const int rows_here = g.LQ - pbase < rpp ? g.LQ - pbase : rpp;
        const int lanes = (rows_here + R - 1) / R;
        for (int step = 0; step < g.LS + lanes - 1; ++step) {
            const int cur = step & 1, prv = cur ^ 1;
            for (int t = 0; t < lanes; ++t) {
                const int n = step - t;
                if (n < 0 || n >= g.LS) continue;
                const int mbase = pbase + t * R;
                int h_up, g_up;
                if (t == 0) {
                    if (pass == 0) { h_up = k1_top(g, n); g_up = K1_NEG; }
                    else { h_up = rowH[n]; g_up = rowG[n]; }
                } else { h_up = send_h[prv][t - 1]; g_up = send_g[prv][t - 1]; }
                int kL, kU;
                k1_poke_rows(g, mbase, n, &kL, &kU);
                if (kL >= 0 && kL < R) L[t].E[kL] = K1_NEG;
                if (kU >= 0 && kU < R) L[t].E[kU] = K1_NEG;
                int sc[R];
                for (int k = 0; k < R; ++k) {
                    int m = mbase + k;
                    sc[k] = m < g.LQ ? mtx[q[m] * dim + s[n]] + 2 * g.u : 0;
                }
                int h_dn, g_dn;
                unsigned long long bits = k2_lane_step(L[t], sc, negv, h_up, g_up, mbase == 0, &h_dn, &g_dn);
                {   // R / 2 bytes per lane-step, slot (pass * (LS + 31) + step) * 32 + lane (k2_core.cuh)
                    unsigned char* w = reinterpret_cast<unsigned char*>(words.data()) + (((size_t)pass * (g.LS + 31) + step) * 32 + t) * (R / 2);
                    for (int b = 0; b < R / 2; ++b) w[b] = (unsigned char)(bits >> (8 * b));
                }
                send_h[cur][t] = h_dn; send_g[cur][t] = g_dn;
                if (t == T - 1) { rowH[n] = h_dn; rowG[n] = g_dn; }
            }
        }
        if (pbase + rows_here == g.LQ) {
            int tl = (rows_here - 1) / R, kf = (rows_here - 1) % R;
            result = L[tl].H[kf];
        }
    }
    *score = result - (g.LQ + g.LS) * g.u;
    std::vector<unsigned char> moves(g.LQ + g.LS + 4);
    std::vector<K2Rec> recs(g.LQ + g.LS + 8);
    return k2_trace(words.data(), g.LQ, g.LS, R, 0, 0, moves.data(), recs.data(), out);
}

// out: 2 ints per corner, Vmf back-walk order; returns corner count
extern "C" int k2_emul_align(const uint8_t* q, int LQ, const uint8_t* s, int LS, const int* mtx, int dim, int u, int v,
                             int sh, int R, int* out, int* score)
{
    K1Geom g;
    g.LQ = LQ; g.LS = LS; g.u = u; g.v = v;
    g.topOpen = -v; g.topExt = -u; g.leftOpen = -v; g.leftExt = -u;
    k1_band(LQ, LS, sh, &g.lw, &g.up);
    switch (R) {
        case 4: return emulate<4>(q, s, g, mtx, dim, out, score);
        case 8: return emulate<8>(q, s, g, mtx, dim, out, score);
        default: return emulate<16>(q, s, g, mtx, dim, out, score);
    }
}
