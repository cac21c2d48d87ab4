// Host emulation of kernel K1F's warp (tests only): compiles k1f_core.cuh with g++ so that the
// floating-point recurrence, boundary tables, band logic and lastD scans can be compared with the
// oracle on a machine without a GPU.
#include <vector>
#include "../../prrn_aln_b200/csrc/k1f_core.cuh"

template <typename T>
static double run(const uint8_t* q, int LQ, uint8_t qf, const uint8_t* s, int LS, uint8_t sf, const double* mtx, int dim,
                  double u, double v, float tgapf, int sh, int R, int swg, int swap)
{
    std::vector<T> m((size_t)dim * dim);
    for (int i = 0; i < dim * dim; ++i) m[i] = (T)mtx[i];
    const T uu = (T)((float)u * 1.f), vv = (T)((float)v * 1.f);   // fwd2d1.cc:62-63: float products
    const int stride = (LQ > LS ? LQ : LS) + 2;
    std::vector<T> bnd(3 * (size_t)stride);
    k1f_build_tables<T>(bnd.data(), stride, uu, vv, tgapf);
    K1FPair<T> g;
    k1f_pair_setup<T>(g, LQ, LS, sh, qf, sf, bnd.data(), stride, tgapf, swap != 0);
    if (swg) {
        if (R == 4) return k1f_emulate_pair<T, 4, true>(q, s, g, m.data(), dim, vv, uu);
        if (R == 8) return k1f_emulate_pair<T, 8, true>(q, s, g, m.data(), dim, vv, uu);
        return k1f_emulate_pair<T, 16, true>(q, s, g, m.data(), dim, vv, uu);
    }
    if (R == 4) return k1f_emulate_pair<T, 4, false>(q, s, g, m.data(), dim, vv, uu);
    if (R == 8) return k1f_emulate_pair<T, 8, false>(q, s, g, m.data(), dim, vv, uu);
    return k1f_emulate_pair<T, 16, false>(q, s, g, m.data(), dim, vv, uu);
}

// mtx[row residue * dim + column residue]; flags bit0 exgl, bit1 exgr, bit2 left != 0, bit3 right != len
extern "C" double k1f_emul_score(const uint8_t* q, int LQ, int qf, const uint8_t* s, int LS, int sf, const double* mtx,
                                 int dim, double u, double v, float tgapf, int sh, int R, int vtype, int swg, int swap)
{
    return vtype ? run<double>(q, LQ, (uint8_t)qf, s, LS, (uint8_t)sf, mtx, dim, u, v, tgapf, sh, R, swg, swap)
                 : run<float>(q, LQ, (uint8_t)qf, s, LS, (uint8_t)sf, mtx, dim, u, v, tgapf, sh, R, swg, swap);
}
