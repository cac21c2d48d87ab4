// Host emulation of kernel K1's warp (tests only): compiles k1_core.cuh with g++ and exposes one C
// entry so that tests/test_k1_emulation.py can compare the kernel's arithmetic, boundary and band
// logic with the oracle on a machine without a GPU.
#include "../../prrn_aln_b200/csrc/k1_core.cuh"

extern "C" int k1_emul_score(const uint8_t* q, int LQ, const uint8_t* s, int LS, const int* mtx, int dim,
                             int u, int v, int sh, int topOpen, int topExt, int leftOpen, int leftExt, int R)
{
    K1Geom g;
    g.LQ = LQ; g.LS = LS; g.u = u; g.v = v;
    g.topOpen = topOpen; g.topExt = topExt; g.leftOpen = leftOpen; g.leftExt = leftExt;
    k1_band(LQ, LS, sh, &g.lw, &g.up);
    switch (R) {
        case 4: return k1_emulate_pair<4>(q, s, g, mtx, dim);
        case 8: return k1_emulate_pair<8>(q, s, g, mtx, dim);
        default: return k1_emulate_pair<16>(q, s, g, mtx, dim);
    }
}

extern "C" long long k1_emul_cells(int LQ, int LS, int sh)
{
    int lw, up;
    k1_band(LQ, LS, sh, &lw, &up);
    return k1_cells(LQ, LS, lw, up);
}
