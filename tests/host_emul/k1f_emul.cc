// Host emulation of kernel K1F's warp (tests only): compiles k1f_core.cuh with g++ so that the
// floating-point recurrence, boundary tables, band logic and lastD scans can be compared with the
// oracle on a machine without a GPU.
#include <vector>
#include "../../prrn_aln_b200/csrc/k1f_core.cuh"

template <typename T>
static double run(const uint8_t* q, int LQ, uint8_t qf, const uint8_t* s, int LS, uint8_t sf, const double* mtx, int dim,
                  double u, double v, float tgapf, int sh, int R, int mode, int swap, int r0, int* ends)
{
    std::vector<T> m((size_t)dim * dim);
    for (int i = 0; i < dim * dim; ++i) m[i] = (T)mtx[i];
    const T uu = (T)((float)u * 1.f), vv = (T)((float)v * 1.f);   // fwd2d1.cc:62-63: float products
    const int stride = (LQ > LS ? LQ : LS) + 2;
    std::vector<T> bnd(3 * (size_t)stride);
    k1f_build_tables<T>(bnd.data(), stride, uu, vv, tgapf);
    K1FPair<T> g;
    k1f_pair_setup<T>(g, LQ, LS, sh, qf, sf, bnd.data(), stride, tgapf, swap != 0, mode == 2, r0);
#define RUN(RR, SWG, VD) return k1f_emulate_pair<T, RR, SWG, VD>(q, s, g, m.data(), dim, vv, uu, ends)
    if (mode == 1) { if (R == 4) RUN(4, true, false); if (R == 8) RUN(8, true, false); RUN(16, true, false); }
    if (mode == 2) { if (R == 4) RUN(4, false, true); if (R == 8) RUN(8, false, true); RUN(16, false, true); }
    if (R == 4) RUN(4, false, false);
    if (R == 8) RUN(8, false, false);
    RUN(16, false, false);
#undef RUN
}

// mtx[row residue * dim + column residue]; flags bit0 exgl, bit1 exgr, bit2 left != 0, bit3 right != len
// mode 0: forwardD + lastD, 1: swgforwardD, 2: Fwd2d_vd (ends[2] written; r0 = b.left - a.left)
extern "C" double k1f_emul_score(const uint8_t* q, int LQ, int qf, const uint8_t* s, int LS, int sf, const double* mtx,
                                 int dim, double u, double v, float tgapf, int sh, int R, int vtype, int mode, int swap,
                                 int r0, int* ends)
{
    return vtype ? run<double>(q, LQ, (uint8_t)qf, s, LS, (uint8_t)sf, mtx, dim, u, v, tgapf, sh, R, mode, swap, r0, ends)
                 : run<float>(q, LQ, (uint8_t)qf, s, LS, (uint8_t)sf, mtx, dim, u, v, tgapf, sh, R, mode, swap, r0, ends);
}
