// k4_contract.cu -- kernel K4: the one dense contraction of the path, on the tensor cores.
//
// sim2(asi, bsi) of the group-to-group DP (reference src/maln2.cc:534-623, 1230-1296: sim11 .. sim33 and
// their weighted forms) is, for every variant, the dot product over residue codes of a per-column
// vector of a (profile part of the profile vector, or member residues folded through the matrix) with
// a per-column vector of b (frequency part, or member weights per code).  The reference evaluates it
// inside the DP cell; here the whole column score matrix
//         S[m][n] = X_a[m] . Y_b[n]          (LQ x LS, K = kdim <= 32)
// is precomputed once per group pair and K3 reads one number per cell.
//
// VTYPE is double in the prrn build and the parity tolerance is 1e-5 relative, so the contraction runs
// on the FP64 tensor-core path: mma.sync.aligned.m8n8k4 .f64 (SASS DMMA).  tcgen05 has no FP64 kind and
// a TF32 / BF16 split would need 3-6 passes to reach the tolerance for a K = 25 contraction that is
// bound by its 8-byte-per-cell output store, not by the multiply.
//   CTA = 4 warps = one 8-row block of one pair; each warp sweeps 32-column panels (4 MMA n-tiles),
//   K in steps of 4.  Fragment layout (PTX ISA, m8n8k4 .f64): lane l holds A[l/4][l%4], B[l%4][l/4],
//   C[l/4][2*(l%4) + {0,1}].
#include <cuda_runtime.h>
#include <stdint.h>

#include "pg_internal.h"

namespace {

__device__ __forceinline__ void dmma8x8x4(double& c0, double& c1, double a, double b)
{
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(c0), "+d"(c1)
                 : "d"(a), "d"(b));
}

__global__ void __launch_bounds__(128) k4_contract_kernel(const K4Args a)
{
    // which pair owns this row block: binary search over the prefix sums of 8-row blocks
    int lo = 0, hi = a.npairs - 1;
    const int blk = blockIdx.x;
    while (lo < hi) {
        const int mid = (lo + hi + 1) >> 1;
        if (a.block_off[mid] <= blk) lo = mid; else hi = mid - 1;
    }
    const K3Pair& P = a.pairs[lo];
    if (!P.simmat) return;
    const int LQ = P.a.L, LS = P.b.L, K = P.prm.kdim;
    const int m0 = (blk - a.block_off[lo]) * 8;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int row = lane >> 2, kk = lane & 3;
    // staged entry x <-> sequence position left-1+x: DP row m is entry m+1
    const double* xa = P.a.prof + (size_t)(m0 + row + 1) * K;
    const bool row_ok = m0 + row < LQ;
    double* out = const_cast<double*>(P.simmat);
    for (int n0 = warp * 32; n0 < LS; n0 += 128) {
        double c[4][2] = {{0, 0}, {0, 0}, {0, 0}, {0, 0}};
        for (int k0 = 0; k0 < K; k0 += 4) {
            const int k = k0 + kk;
            const double av = (row_ok && k < K) ? __ldg(xa + k) : 0.0;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int n = n0 + 8 * j + row;             // B[k][col]: col = lane/4
                const double bv = (n < LS && k < K) ? __ldg(P.b.freq + (size_t)(n + 1) * K + k) : 0.0;
                dmma8x8x4(c[j][0], c[j][1], av, bv);
            }
        }
        if (row_ok) {
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int n = n0 + 8 * j + 2 * kk;
                double* o = out + (size_t)(m0 + row) * LS + n;
                if (n < LS) o[0] = c[j][0];
                if (n + 1 < LS) o[1] = c[j][1];
            }
        }
    }
}

}  // namespace

cudaError_t k4_launch(const K4Args& a, int total_blocks, cudaStream_t st)
{
    if (total_blocks <= 0) return cudaSuccess;
    k4_contract_kernel<<<total_blocks, 128, 0, st>>>(a);
    return cudaGetLastError();
}
