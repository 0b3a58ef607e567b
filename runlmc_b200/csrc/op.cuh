// The fused SKI-LMC operator handle.
#pragma once
#include "common.cuh"
#include "interp.cuh"
#include "spectral.cuh"
#include <vector>

namespace lmc {
// Host-side state of the block solvers that outlives a solve: pinned flags and events of the
// asynchronous stop test, and the stream iterations are captured on.
struct SolverHost {
    static const int kRing = 4;
    int* flags = nullptr;                  // pinned, [kRing]
    cudaEvent_t ev[kRing] = {nullptr, nullptr, nullptr, nullptr};
    cudaEvent_t ev_join = nullptr;
    cudaStream_t stream = nullptr;
    ~SolverHost();
};
}  // namespace lmc

struct lmc_op {
    int D = 0, ndim = 0, Q = 0;
    lmc::Embedding emb;
    lmc::SpectralEngine eng;
    lmc::PointSet ps;
    double* spec = nullptr;   // [Q][bins] real circulant spectra / bins (digit-reversed layout)
    double* specL = nullptr;  // [Q][line][pos] line-major copy for the fused spectral kernel
    double* specP = nullptr;  // same in the mix layout of the 512-point column kernel (when used)
    double origin[2] = {0.0, 0.0}, delta[2] = {1.0, 1.0};   // grid axes: origin + k * delta
    std::vector<double> B_host;
    // kernel descriptors of the last lmc_op_set_kernels (empty after lmc_op_set_params): kinds[Q],
    // kparams[Q][2] = (inv_lengthscale, period)
    std::vector<int> kinds;
    std::vector<double> kparams;
    std::vector<int> ranks;          // optional factors B_q = A_q^T A_q + diag(kappa_q)
    std::vector<double> A_host, kappa_host;
    lmc::MixSpec mix_spec() const {
        lmc::MixSpec m;
        m.B = B_host.data();
        if (!ranks.empty()) { m.ranks = ranks.data(); m.A = A_host.data(); m.kappa = kappa_host.data(); }
        return m;
    }
    bool fused = false;       // fused spectral path usable for this geometry / D / Q
    int fused_tile_pairs = 0;
    double* B = nullptr;      // [Q][D][D]
    double* noise = nullptr;  // [D]
    int spec_cap = 0;         // allocated Q
    // grid-stage workspace for `tile_pairs` RHS pairs
    cplx* G = nullptr;        // [g_pairs][D][grid_pitch]  grid-side vectors of a block of RHS pairs
    cplx* S = nullptr;        // [tile_pairs][D][bins]     spectra of one L2-sized sub-tile of pairs
    double* Vs = nullptr;     // [2 g_pairs][n] sorted-order copy of a block of caller-ordered columns
    int g_pairs = 0;
    int tile_pairs = 0;
    // host-buffer entry points: double-buffered device staging + 3 streams (copy in / compute / copy out)
    double* stage_in[2] = {nullptr, nullptr};
    double* stage_out[2] = {nullptr, nullptr};
    size_t stage_cap = 0;   // doubles per staging buffer
    // block-solver state vectors (MINRES: 8 [P][n] blocks, CG: 5), grow-only, kept between solves:
    // cudaMalloc/cudaFree of ~8 n P doubles per solve would cost as much as tens of iterations
    void* solver_ws = nullptr;
    size_t solver_ws_cap = 0;
    lmc::SolverHost solver_host;
    void* grad_ws = nullptr;       // scratch of the gradient Gram stage (grad.cu), grow-only
    size_t grad_ws_cap = 0;
    double* jacobi = nullptr;      // [n] 1 / diag(K~) in sorted order, valid for the current parameters if jacobi_valid
    bool jacobi_valid = false;
    double t16[64 * 16] = {0};     // top_q at the offsets (0..3, 0..3) a cubic stencil spans, per kernel (Q <= 64)
    cudaStream_t hs[3] = {nullptr, nullptr, nullptr};
    cudaEvent_t ev_in[2] = {nullptr, nullptr}, ev_cmp[2] = {nullptr, nullptr}, ev_out[2] = {nullptr, nullptr};
    ~lmc_op();
};

struct lmc_bttb {
    lmc::Embedding emb;
    lmc::SpectralEngine eng;
    double* spec = nullptr;  // [bins]
    double* one = nullptr;   // 1x1 identity "B"
    cplx* G = nullptr;
    cplx* S = nullptr;
    int cap_pairs = 0;
    ~lmc_bttb();
};

namespace lmc {

int op_ensure_workspace(lmc_op* op);
// OUT = K~ V for ncols columns described by cv (cv.in / cv.out), noise included
int op_mvm(lmc_op* op, const ColumnView& cv, cudaStream_t st);
int op_mvm_rows(lmc_op* op, const double* X, long ldx, int P, double* Y, long ldy, cudaStream_t st);
// same without the noise term and with explicit spectra / mixing matrices
// (used by the gradient stage); spec/B device pointers, Q kernels
int op_grid_block(lmc_op* op, cplx* G, int cnt, cudaStream_t st);
int op_grid_apply(lmc_op* op, cplx* G, int npairs, int Q, const double* spec, const double* B,
                  cudaStream_t st);

// jacobi: optional [n] device vector 1 / diag(K~) in sorted order = scipy's M (iterative.py:47)
int minres_solve(lmc_op* op, const double* RHS, long ld, int P, double* X, double tol, int maxiter,
                 int check_every, int* iters, double* resid, int* istop, cudaStream_t st,
                 const double* jacobi = nullptr);

// same solve, also returning the Lanczos tridiagonals the recurrence computes (stochastic Lanczos quadrature)
int minres_solve_lanczos(lmc_op* op, const double* RHS, long ld, int P, double* X, double tol, int maxiter,
                         int check_every, int* iters, double* resid, int* istop, int k, double* tridiag,
                         double* beta1, cudaStream_t st);

int cg_solve(lmc_op* op, const double* RHS, long ld, int P, double* X, double tol, int maxiter,
             int check_every, int* iters, double* resid, int* info, cudaStream_t st);

// diag(K~) and / or its reciprocal in sorted order (either pointer may be null); precond.cu
int op_jacobi(lmc_op* op, double* diag_sorted, double* inv_sorted, cudaStream_t st);

// tops_extra: [ntops_extra][cells] derivative tops, host memory unless extra_on_device
int grad_grams(lmc_op* op, const double* alpha, const double* R, const double* RINV, long ld, int N,
               int ntops_extra, const double* tops_extra, bool extra_on_device, double* quad, double* trace,
               double* nquad, double* ntrace, cudaStream_t st);
int op_set_params_dev(lmc_op* op, int Q, const double* top_dev, const double* B_host, const double* noise_host);

}  // namespace lmc
