// Gradient / Hutchinson-trace contractions for all hyper-parameters at once.
//
// The reference evaluates, per hyper-parameter j, (N+1) full SKI products with
// dK_j (StochasticDeriv.d_normal_quadratic / d_logdet_K,
// runlmc/lmc/stochastic_deriv.py:69-78, with the dK operators of
// runlmc/lmc/likelihood.py:112-128).  Every dK_j except the noise ones has the
// form W (C (x) T_t) W^T with a D x D matrix C and a BTTB T_t, so
//     u^T dK_j z = sum_{d,e} C[d][e] * (W^T u)_d^T T_t (W^T z)_e = <C, Gram_t(u, z)>.
// By Parseval on the zero-padded circulant embedding,
//     (W^T u)_d^T T_t (W^T z)_e = sum_k spec_t[k] Re(conj(U_d[k]) Z_e[k]),
// so one forward transform per vector (no inverse) and a per-bin D x D
// cross-spectrum give the Gram matrices of ALL tops.  Probes are processed as
// complex pairs (u_a + i u_b, z_a + i z_b); the cross terms cancel in the sum
// over k because spec_t is real and even.
#include "op.cuh"

#include <algorithm>
#include <vector>

namespace lmc {

// C[d][e][bin] += sum_pairs Re(conj(U[pair][d][bin]) Z[pair][e][bin])
template <int D>
__global__ void __launch_bounds__(128) cross_spectrum_kernel(const cplx* __restrict__ U,
                                                             const cplx* __restrict__ Z, long bins,
                                                             int npairs, double* C, int accumulate) {
    const long bin = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (bin >= bins) return;
    double acc[D][D];
#pragma unroll
    for (int d = 0; d < D; ++d)
#pragma unroll
        for (int e = 0; e < D; ++e) acc[d][e] = accumulate ? C[((long)d * D + e) * bins + bin] : 0.0;
    for (int p = 0; p < npairs; ++p) {
        cplx u[D], z[D];
#pragma unroll
        for (int d = 0; d < D; ++d) {
            u[d] = U[((long)p * D + d) * bins + bin];
            z[d] = Z[((long)p * D + d) * bins + bin];
        }
#pragma unroll
        for (int d = 0; d < D; ++d)
#pragma unroll
            for (int e = 0; e < D; ++e) acc[d][e] = fma(u[d].x, z[e].x, fma(u[d].y, z[e].y, acc[d][e]));
    }
#pragma unroll
    for (int d = 0; d < D; ++d)
#pragma unroll
        for (int e = 0; e < D; ++e) C[((long)d * D + e) * bins + bin] = acc[d][e];
}

__global__ void cross_spectrum_generic_kernel(const cplx* __restrict__ U, const cplx* __restrict__ Z,
                                              long bins, int npairs, int D, double* C, int accumulate) {
    const long bin = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (bin >= bins) return;
    for (int d = 0; d < D; ++d)
        for (int e = 0; e < D; ++e) {
            double acc = accumulate ? C[((long)d * D + e) * bins + bin] : 0.0;
            for (int p = 0; p < npairs; ++p) {
                const cplx u = U[((long)p * D + d) * bins + bin];
                const cplx z = Z[((long)p * D + e) * bins + bin];
                acc = fma(u.x, z.x, fma(u.y, z.y, acc));
            }
            C[((long)d * D + e) * bins + bin] = acc;
        }
}

// part[t][de][blk] = sum_{bin in chunk} spec[t][bin] * C[de][bin]
__global__ void __launch_bounds__(256) contract_kernel(const double* __restrict__ C,
                                                       const double* __restrict__ spec, long bins, int T,
                                                       double* part, int nblk) {
    __shared__ double s_red[8];
    const int de = blockIdx.y;
    const long base = (long)blockIdx.x * 2048;
    for (int t = 0; t < T; ++t) {
        double acc = 0.0;
        for (int k = 0; k < 8; ++k) {
            const long bin = base + k * 256 + threadIdx.x;
            if (bin < bins) acc = fma(spec[(long)t * bins + bin], C[(long)de * bins + bin], acc);
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) acc += __shfl_down_sync(0xffffffffu, acc, o);
        __syncthreads();
        if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = acc;
        __syncthreads();
        if (threadIdx.x == 0) {
            double s = 0.0;
            for (int w = 0; w < 8; ++w) s += s_red[w];
            part[((long)t * gridDim.y + de) * nblk + blockIdx.x] = s;
        }
    }
}

__global__ void final_sum_kernel(const double* __restrict__ part, int nblk, int count, double* out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= count) return;
    double s = 0.0;
    for (int b = 0; b < nblk; ++b) s += part[(long)i * nblk + b];
    out[i] = s;
}

// per-output dot products  out[d] = sum_cols sum_{i in output d} A[c][i] * B[c][i]
__global__ void __launch_bounds__(256) output_dot_kernel(const double* __restrict__ A,
                                                         const double* __restrict__ Bv, long ld, int ncols,
                                                         const long* __restrict__ out_start, double* part,
                                                         int nblk) {
    __shared__ double s_red[8];
    const int d = blockIdx.y;
    const long lo = out_start[d], hi = out_start[d + 1];
    double acc = 0.0;
    for (int c = 0; c < ncols; ++c)
        for (long i = lo + (long)blockIdx.x * 256 + threadIdx.x; i < hi; i += (long)nblk * 256)
            acc = fma(A[(long)c * ld + i], Bv[(long)c * ld + i], acc);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_down_sync(0xffffffffu, acc, o);
    if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        double s = 0.0;
        for (int w = 0; w < 8; ++w) s += s_red[w];
        part[(long)d * nblk + blockIdx.x] = s;
    }
}

// Scratch of the Gram stage, carved out of ONE grow-only allocation owned by the operator handle: at config E
// the stage needs ~2.3 GB (a second set of grid / spectrum slabs, two D x D cross-spectra) and allocating and
// freeing that per gradient evaluation cost 4-6x the 16 ms its kernels take.
struct GradArena {
    char* base;
    size_t cap, used = 0;
    GradArena(void* p, size_t c) : base(static_cast<char*>(p)), cap(c) {}
    template <class T> T* take(size_t count) {
        const size_t bytes = (sizeof(T) * (count ? count : 1) + 255) & ~(size_t)255;
        T* p = reinterpret_cast<T*>(base + used);
        used += bytes;
        return used <= cap ? p : nullptr;
    }
};
static size_t arena_round(size_t bytes) { return (bytes + 255) & ~(size_t)255; }

static int cross_spectrum(int D, const cplx* U, const cplx* Z, long bins, int npairs, double* C,
                          int accumulate, cudaStream_t st) {
    const unsigned grid = (unsigned)ceil_div(bins, 128);
    ProfScope prof(PROF_GRAD, st);
    switch (D) {
#define LMC_CS_CASE(DD)                                                                        \
    case DD:                                                                                   \
        cross_spectrum_kernel<DD><<<grid, 128, 0, st>>>(U, Z, bins, npairs, C, accumulate);    \
        break;
        LMC_CS_CASE(1) LMC_CS_CASE(2) LMC_CS_CASE(3) LMC_CS_CASE(4) LMC_CS_CASE(5) LMC_CS_CASE(6)
        LMC_CS_CASE(7) LMC_CS_CASE(8) LMC_CS_CASE(9) LMC_CS_CASE(10)
#undef LMC_CS_CASE
        default:
            cross_spectrum_generic_kernel<<<grid, 128, 0, st>>>(U, Z, bins, npairs, D, C, accumulate);
    }
    count_launch();
    LMC_CHECK(cudaGetLastError());
    return 0;
}

int grad_grams(lmc_op* op, const double* alpha, const double* R, const double* RINV, long ld, int N,
               int ntops_extra, const double* tops_extra, bool extra_on_device, double* quad, double* trace,
               double* nquad, double* ntrace, cudaStream_t st) {
    LMC_REQUIRE(op->Q > 0, "operator parameters not set");
    LMC_REQUIRE(N >= 0 && ntops_extra >= 0, "negative count");
    LMC_TRY(op_ensure_workspace(op));
    const int D = op->D, Q = op->Q, T = Q + ntops_extra;
    const long bins = op->emb.bins, cells = op->emb.cells, n = op->ps.n;
    LMC_REQUIRE(ld >= n, "leading dimension < n");
    const int DD = D * D;

    // one arena for everything below (grow-only, kept in the handle between evaluations)
    const int tile = op->tile_pairs;
    const int nblk = ceil_div(bins, 2048);
    const int nb2 = 64;
    const size_t need = arena_round(sizeof(double) * (size_t)T * bins) + arena_round(sizeof(cplx) * (size_t)bins) +
                        arena_round(sizeof(double) * (size_t)std::max(ntops_extra, 1) * cells) +
                        arena_round(sizeof(cplx) * (size_t)tile * D * op->emb.grid_pitch) +
                        arena_round(sizeof(cplx) * (size_t)tile * D * bins) +
                        2 * arena_round(sizeof(double) * (size_t)DD * bins) +
                        arena_round(sizeof(double) * (size_t)T * DD * nblk) +
                        arena_round(sizeof(double) * (size_t)(2 * T * DD + 2 * D)) +
                        arena_round(sizeof(double) * (size_t)D * nb2) + 4096;
    if (need > op->grad_ws_cap) {
        cudaFree(op->grad_ws);
        op->grad_ws = nullptr; op->grad_ws_cap = 0;
        LMC_CHECK(cudaMalloc(&op->grad_ws, need));
        op->grad_ws_cap = need;
    }
    GradArena arena(op->grad_ws, op->grad_ws_cap);

    // spectra of all tops: the Q kernels already live in op->spec; derivative tops are transformed here
    double* spec = arena.take<double>((size_t)T * bins);
    LMC_CHECK(cudaMemcpyAsync(spec, op->spec, sizeof(double) * (size_t)Q * bins, cudaMemcpyDeviceToDevice, st));
    if (ntops_extra) {
        cplx* work = arena.take<cplx>((size_t)bins);
        const double* tops_dev = tops_extra;   // derivative tops evaluated on the device (setup.cu) ...
        if (!extra_on_device) {                // ... or uploaded by the caller
            double* topb = arena.take<double>((size_t)ntops_extra * cells);
            LMC_CHECK(cudaMemcpyAsync(topb, tops_extra, sizeof(double) * (size_t)ntops_extra * cells,
                                      cudaMemcpyHostToDevice, st));
            tops_dev = topb;
        }
        for (int t = 0; t < ntops_extra; ++t)
            LMC_TRY(op->eng.spectrum(tops_dev + (size_t)t * cells, spec + (size_t)(Q + t) * bins, work, st));
    }

    // second set of grid / spectrum slabs for the z side
    cplx* G2 = arena.take<cplx>((size_t)tile * D * op->emb.grid_pitch);
    LMC_CHECK(cudaMemsetAsync(G2, 0, sizeof(cplx) * (size_t)tile * D * op->emb.grid_pitch, st));
    cplx* S2 = arena.take<cplx>((size_t)tile * D * bins);
    double* cq = arena.take<double>((size_t)DD * bins);
    double* ct = arena.take<double>((size_t)DD * bins);
    double* part = arena.take<double>((size_t)T * DD * nblk);
    double* out = arena.take<double>((size_t)(2 * T * DD + 2 * D));
    double* pn = arena.take<double>((size_t)D * nb2);
    LMC_REQUIRE(pn != nullptr, "internal: gradient scratch arena too small");

    // quadratic term: u = z = alpha (one "pair" with zero imaginary part)
    ColumnView cv;
    cv.ld = ld;
    cv.in = alpha; cv.ncols = 1;
    LMC_TRY(to_grid(op->ps, cv, op->G, st));
    LMC_TRY(op->eng.forward(op->G, op->S, D, st));
    LMC_TRY(cross_spectrum(D, op->S, op->S, bins, 1, cq, 0, st));

    // trace term: pairs of probes
    const int npairs = (N + 1) / 2;
    if (npairs == 0) LMC_CHECK(cudaMemsetAsync(ct, 0, sizeof(double) * (size_t)DD * bins, st));
    for (int p0 = 0; p0 < npairs; p0 += tile) {
        const int cnt = std::min(tile, npairs - p0);
        const int c0 = 2 * p0;
        const int ncols = std::min(2 * cnt, N - c0);
        cv.in = RINV + (long)c0 * ld; cv.ncols = ncols;
        LMC_TRY(to_grid(op->ps, cv, op->G, st));
        LMC_TRY(op->eng.forward(op->G, op->S, cnt * D, st));
        cv.in = R + (long)c0 * ld;
        LMC_TRY(to_grid(op->ps, cv, G2, st));
        LMC_TRY(op->eng.forward(G2, S2, cnt * D, st));
        LMC_TRY(cross_spectrum(D, op->S, S2, bins, cnt, ct, p0 > 0 ? 1 : 0, st));
    }

    // contract with every spectrum
    for (int which = 0; which < 2; ++which) {
        const double* C = which ? ct : cq;
        contract_kernel<<<dim3((unsigned)nblk, (unsigned)DD), 256, 0, st>>>(C, spec, bins, T, part, nblk);
        final_sum_kernel<<<ceil_div(T * DD, 128), 128, 0, st>>>(part, nblk, T * DD, out + (size_t)which * T * DD);
        count_launch(2);
    }
    // noise terms
    {
        output_dot_kernel<<<dim3(nb2, (unsigned)D), 256, 0, st>>>(alpha, alpha, ld, 1, op->ps.out_start_dev, pn, nb2);
        final_sum_kernel<<<1, 128, 0, st>>>(pn, nb2, D, out + (size_t)2 * T * DD);
        if (N > 0) {
            output_dot_kernel<<<dim3(nb2, (unsigned)D), 256, 0, st>>>(RINV, R, ld, N, op->ps.out_start_dev, pn, nb2);
            final_sum_kernel<<<1, 128, 0, st>>>(pn, nb2, D, out + (size_t)2 * T * DD + D);
        } else {
            LMC_CHECK(cudaMemsetAsync(out + (size_t)2 * T * DD + D, 0, sizeof(double) * D, st));
        }
        count_launch(4);
        LMC_CHECK(cudaStreamSynchronize(st));
    }
    LMC_CHECK(cudaGetLastError());
    std::vector<double> h((size_t)(2 * T * DD + 2 * D));
    LMC_CHECK(cudaMemcpy(h.data(), out, sizeof(double) * h.size(), cudaMemcpyDeviceToHost));
    for (int i = 0; i < T * DD; ++i) {
        quad[i] = h[i];
        trace[i] = h[(size_t)T * DD + i];
    }
    for (int d = 0; d < D; ++d) {
        nquad[d] = h[(size_t)2 * T * DD + d];
        ntrace[d] = h[(size_t)2 * T * DD + D + d];
    }
    return 0;
}

}  // namespace lmc
