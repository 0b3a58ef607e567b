// Jacobi preconditioner of the fused operator: the exact diagonal of
//   K~ = W (sum_q B_q (x) T_q) W^T + diag(noise)
// without forming K~.  Row i of W holds the (at most 4^d) cubic-convolution weights of point i on the
// cells clamp(i0 - 1 + t), so
//   K_ii = noise_d + sum_q B_q[d][d] * sum_{ox, oy} cx[ox] cy[oy] top_q(ox, oy),
// cx[o] = sum over tap pairs (t, t') whose cells lie o apart of w_t w_t'  (clamped taps share a cell and
// land in o = 0, like the CSR `+=` of reference approx/interpolation.py:105-115).  Only the 4 x 4 top
// values at offsets 0..3 are needed per kernel.  The reference forwards K.preconditioner to scipy as M
// (approx/iterative.py:47-50) and never builds one; this is the opt-in M the device solver offers.
#include "op.cuh"
#include "interp_weights.cuh"

namespace lmc {

__global__ void __launch_bounds__(256) jacobi_kernel(const double* __restrict__ u0, const double* __restrict__ u1,
                                                      const int* __restrict__ i00, const int* __restrict__ i01,
                                                      const long* __restrict__ out_start, int D, int ndim, int m0,
                                                      int m1, int Q, const double* __restrict__ t16,
                                                      const double* __restrict__ B, const double* __restrict__ noise,
                                                      long n, double* diag, double* inv_diag) {
    const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    int d = 0;
    while (d + 1 < D && i >= out_start[d + 1]) ++d;
    double c[2][4];
#pragma unroll
    for (int ax = 0; ax < 2; ++ax) {
        c[ax][0] = 1.0; c[ax][1] = c[ax][2] = c[ax][3] = 0.0;
        if (ax >= ndim) continue;
        double w[4];
        keys_weights(ax == 0 ? u0[i] : u1[i], w);
        const int base = (ax == 0 ? i00[i] : i01[i]) - 1, m = ax == 0 ? m0 : m1;
        int cell[4];
#pragma unroll
        for (int t = 0; t < 4; ++t) cell[t] = min(max(base + t, 0), m - 1);
        c[ax][0] = 0.0;
#pragma unroll
        for (int t = 0; t < 4; ++t)
#pragma unroll
            for (int s = 0; s < 4; ++s) {
                const int o = abs(cell[t] - cell[s]);
                const double p = w[t] * w[s];
#pragma unroll
                for (int k = 0; k < 4; ++k)
                    if (o == k) c[ax][k] += p;
            }
    }
    double acc = noise[d];
    for (int q = 0; q < Q; ++q) {
        double sq = 0.0;
#pragma unroll
        for (int ox = 0; ox < 4; ++ox)
#pragma unroll
            for (int oy = 0; oy < 4; ++oy) sq = fma(c[0][ox] * c[1][oy], t16[(q * 4 + ox) * 4 + oy], sq);
        acc = fma(B[((long)q * D + d) * D + d], sq, acc);
    }
    if (diag) diag[i] = acc;
    if (inv_diag) inv_diag[i] = 1.0 / acc;
}

// diag (optional, [n] device) and/or inv_diag in the operator's sorted point order
int op_jacobi(lmc_op* op, double* diag_sorted, double* inv_sorted, cudaStream_t st) {
    LMC_REQUIRE(op->Q > 0, "operator parameters not set");
    LMC_REQUIRE(op->Q <= 64, "Q must be <= 64");
    const long n = op->ps.n;
    if (n == 0) return 0;
    double* t16 = nullptr;
    LMC_CHECK(cudaMallocAsync(&t16, sizeof(double) * 16 * op->Q, st));
    LMC_CHECK(cudaMemcpyAsync(t16, op->t16, sizeof(double) * 16 * op->Q, cudaMemcpyHostToDevice, st));
    jacobi_kernel<<<ceil_div(n, 256), 256, 0, st>>>(op->ps.u[0], op->ps.u[1], op->ps.i0[0], op->ps.i0[1],
                                                    op->ps.out_start_dev, op->D, op->ndim, op->ps.m[0], op->ps.m[1],
                                                    op->Q, t16, op->B, op->noise, n, diag_sorted, inv_sorted);
    count_launch();
    LMC_CHECK(cudaGetLastError());
    LMC_CHECK(cudaFreeAsync(t16, st));
    return 0;
}

}  // namespace lmc
