// Host emulation of the packed (int16 x 2) K1 warp: two alignments (q0 x s, q1 x s) per call.
#include "../../prrn_aln_b200/csrc/k1p_core.cuh"

extern "C" int k1p_emul_score2(const uint8_t* q0, int LQ0, const uint8_t* q1, int LQ1, const uint8_t* s, int LS,
                               const int* mtx, int dim, int u, int v, int sh, int R, int* out0, int* out1)
{
    K1Geom g[2];
    const int LQ[2] = {LQ0, LQ1};
    for (int h = 0; h < 2; ++h) {
        g[h].LQ = LQ[h]; g[h].LS = LS; g[h].u = u; g[h].v = v;
        g[h].topOpen = -v; g[h].topExt = -u; g[h].leftOpen = -v; g[h].leftExt = -u;
        k1_band(LQ[h], LS, sh, &g[h].lw, &g[h].up);
    }
    switch (R) {
        case 4: k1p_emulate_pair<4>(q0, q1, s, g[0], g[1], mtx, dim, out0, out1); break;
        case 8: k1p_emulate_pair<8>(q0, q1, s, g[0], g[1], mtx, dim, out0, out1); break;
        default: k1p_emulate_pair<16>(q0, q1, s, g[0], g[1], mtx, dim, out0, out1); break;
    }
    return 0;
}

extern "C" int k1p_emul_fits(int smax, int smin, int v, int lmax) { return k1p_fits(smax, smin, v, lmax) ? 1 : 0; }
