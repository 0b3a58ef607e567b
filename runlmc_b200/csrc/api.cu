// C ABI of the fused SKI-LMC operator (include/lmc_b200.h).
#include "../../include/lmc_b200.h"
#include "op.cuh"

#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <vector>

using namespace lmc;

namespace {
struct HostBlock {  // device staging for the *_host entry points
    double* p = nullptr;
    ~HostBlock() { if (p) cudaFree(p); }
};
}  // namespace

namespace lmc {

// Spectra of the Q tops (already on the device), B and noise upload: everything a new set of
// hyper-parameters changes.  Shared by lmc_op_set_params (host tops) and lmc_op_set_kernels
// (tops evaluated on the device, setup.cu).
int op_set_params_dev(lmc_op* op, int Q, const double* top_dev, const double* B_host, const double* noise_host) {
    const long bins = op->emb.bins, cells = op->emb.cells;
    const int D = op->D;
    if (Q > op->spec_cap) {
        cudaFree(op->spec); cudaFree(op->B); cudaFree(op->specL); cudaFree(op->specP);
        op->spec = nullptr; op->B = nullptr; op->specL = nullptr; op->specP = nullptr; op->spec_cap = 0;
        LMC_CHECK(cudaMalloc(&op->spec, sizeof(double) * (size_t)Q * bins));
        LMC_CHECK(cudaMalloc(&op->specL, sizeof(double) * (size_t)Q * bins));
        if (op->eng.col512()) LMC_CHECK(cudaMalloc(&op->specP, sizeof(double) * (size_t)Q * bins));
        LMC_CHECK(cudaMalloc(&op->B, sizeof(double) * (size_t)Q * D * D));
        op->spec_cap = Q;
    }
    if (!op->noise) LMC_CHECK(cudaMalloc(&op->noise, sizeof(double) * D));
    cplx* work = nullptr;
    LMC_CHECK(cudaMalloc(&work, sizeof(cplx) * (size_t)bins));
    int rc = 0;
    cudaError_t e = cudaMemcpy(op->B, B_host, sizeof(double) * (size_t)Q * D * D, cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMemcpy(op->noise, noise_host, sizeof(double) * D, cudaMemcpyHostToDevice);
    for (int q = 0; q < Q && e == cudaSuccess && rc == 0; ++q)
        rc = op->eng.spectrum(top_dev + (size_t)q * cells, op->spec + (size_t)q * bins, work, 0);
    op->fused = op->eng.fused_supported(D, Q) && getenv("LMC_NO_FUSED") == nullptr;
    if (e == cudaSuccess && rc == 0 && op->fused) rc = op->eng.spectrum_lines(op->spec, op->specL, op->specP, Q, 0);
    // top values at the offsets (0..3, 0..3) a cubic stencil spans: all the Jacobi diagonal needs (precond.cu)
    for (int q = 0; q < Q && e == cudaSuccess && rc == 0; ++q) {
        double* t = op->t16 + 16 * q;
        for (int i = 0; i < 16; ++i) t[i] = 0.0;
        if (op->ndim == 2) {
            e = cudaMemcpy2D(t, 4 * sizeof(double), top_dev + (size_t)q * cells, op->emb.m[1] * sizeof(double),
                             4 * sizeof(double), 4, cudaMemcpyDeviceToHost);
        } else {
            double t4[4];
            e = cudaMemcpy(t4, top_dev + (size_t)q * cells, sizeof(t4), cudaMemcpyDeviceToHost);
            for (int i = 0; i < 4; ++i) t[4 * i] = t4[i];
        }
    }
    op->jacobi_valid = false;
    if (e == cudaSuccess && rc == 0) e = cudaDeviceSynchronize();
    cudaFree(work);
    if (rc != 0) return rc;
    LMC_CHECK(e);
    op->B_host.assign(B_host, B_host + (size_t)Q * D * D);
    op->ranks.clear();
    op->A_host.clear();
    op->kappa_host.clear();
    op->Q = Q;
    return 0;
}

}  // namespace lmc

extern "C" {

int lmc_version(void) { return 100; }
const char* lmc_last_error(void) { return get_error(); }
unsigned long long lmc_launch_count(void) { return g_launches.load(std::memory_order_relaxed); }

int lmc_op_create(lmc_op** out, int D, int ndim, const int* grid_sizes, const double* origin,
                  const double* delta, const int* lens, const double* X_host) {
    LMC_REQUIRE(out && grid_sizes && origin && delta && lens && X_host, "null argument");
    LMC_REQUIRE(ndim == 1 || ndim == 2, "ndim must be 1 or 2");
    lmc_op* op = new lmc_op();
    op->D = D;
    op->ndim = ndim;
    for (int p = 0; p < ndim; ++p) { op->origin[p] = origin[p]; op->delta[p] = delta[p]; }
    int rc = embedding_init(&op->emb, ndim, grid_sizes);
    if (rc == 0) rc = op->eng.init(op->emb);
    if (rc == 0) rc = build_points(&op->ps, D, ndim, grid_sizes, origin, delta, lens, X_host, op->emb.grid_pitch);
    if (rc != 0) { delete op; return rc; }
    *out = op;
    return 0;
}

int lmc_op_destroy(lmc_op* op) {
    delete op;
    return 0;
}

int lmc_op_set_params(lmc_op* op, int Q, const double* tops_host, const double* B_host,
                      const double* noise_host) {
    LMC_REQUIRE(op && tops_host && B_host && noise_host, "null argument");
    LMC_REQUIRE(Q >= 1 && Q <= 64, "Q must be in 1..64");
    const long cells = op->emb.cells;
    double* top_dev = nullptr;
    LMC_CHECK(cudaMalloc(&top_dev, sizeof(double) * (size_t)Q * cells));
    cudaError_t e = cudaMemcpy(top_dev, tops_host, sizeof(double) * (size_t)Q * cells, cudaMemcpyHostToDevice);
    int rc = 0;
    if (e == cudaSuccess) rc = op_set_params_dev(op, Q, top_dev, B_host, noise_host);
    cudaFree(top_dev);
    LMC_CHECK(e);
    if (rc == 0) { op->kinds.clear(); op->kparams.clear(); }   // tops came from the host: no kernel descriptors
    return rc;
}

int lmc_op_set_coreg_factors(lmc_op* op, const int* ranks_host, const double* A_host,
                             const double* kappa_host) {
    LMC_REQUIRE(op && ranks_host && A_host && kappa_host, "null argument");
    LMC_REQUIRE(op->Q > 0, "call lmc_op_set_params first");
    const int D = op->D, Q = op->Q;
    int total = 0;
    for (int q = 0; q < Q; ++q) {
        LMC_REQUIRE(ranks_host[q] >= 0, "negative rank");
        total += ranks_host[q];
    }
    // the factors must reproduce the dense matrices uploaded by lmc_op_set_params
    int r0 = 0;
    for (int q = 0; q < Q; ++q) {
        for (int i = 0; i < D; ++i)
            for (int j = 0; j < D; ++j) {
                double v = (i == j) ? kappa_host[(size_t)q * D + i] : 0.0;
                for (int r = r0; r < r0 + ranks_host[q]; ++r) v += A_host[(size_t)r * D + i] * A_host[(size_t)r * D + j];
                const double b = op->B_host[((size_t)q * D + i) * D + j];
                LMC_REQUIRE(std::fabs(v - b) <= 1e-12 * (1.0 + std::fabs(b)),
                            "coregionalisation factors do not match B_q = A_q^T A_q + diag(kappa_q)");
            }
        r0 += ranks_host[q];
    }
    op->ranks.assign(ranks_host, ranks_host + Q);
    op->A_host.assign(A_host, A_host + (size_t)total * D);
    op->kappa_host.assign(kappa_host, kappa_host + (size_t)Q * D);
    return 0;
}

long lmc_op_n(const lmc_op* op) { return op ? op->ps.n : -1; }
long lmc_op_grid_cells(const lmc_op* op) { return op ? op->emb.cells : -1; }
long lmc_op_embed_bins(const lmc_op* op) { return op ? op->emb.bins : -1; }
int lmc_op_max_tile_points(const lmc_op* op) { return op ? op->ps.max_tile_pts_8x8 : -1; }

int lmc_op_perm(const lmc_op* op, int* perm_host) {
    LMC_REQUIRE(op && perm_host, "null argument");
    LMC_CHECK(cudaMemcpy(perm_host, op->ps.perm, sizeof(int) * op->ps.n, cudaMemcpyDeviceToHost));
    return 0;
}

int lmc_mvm(lmc_op* op, const double* V_dev, long ld, int P, double* OUT_dev, void* stream) {
    LMC_REQUIRE(op, "null argument");
    if (P == 0) return 0;   // an empty block has no storage to point at
    LMC_REQUIRE(V_dev && OUT_dev, "null argument");
    LMC_REQUIRE(P >= 0 && ld >= op->ps.n, "bad block shape");
    LMC_REQUIRE(V_dev != OUT_dev, "in-place product not supported");
    ColumnView cv;
    cv.in = V_dev; cv.out = OUT_dev; cv.ld = ld; cv.ncols = P;
    return op_mvm(op, cv, (cudaStream_t)stream);
}

int lmc_mvm_sorted(lmc_op* op, const double* V_dev, long ld, int P, double* OUT_dev, void* stream) {
    LMC_REQUIRE(op, "null argument");
    if (P == 0) return 0;   // an empty block has no storage to point at
    LMC_REQUIRE(V_dev && OUT_dev, "null argument");
    LMC_REQUIRE(P >= 0 && ld >= op->ps.n, "bad block shape");
    LMC_REQUIRE(V_dev != OUT_dev, "in-place product not supported");
    ColumnView cv;
    cv.in = V_dev; cv.out = OUT_dev; cv.ld = ld; cv.ncols = P; cv.sorted_in = cv.sorted_out = true;
    return op_mvm(op, cv, (cudaStream_t)stream);
}

int lmc_mvm_rows(lmc_op* op, const double* X_dev, long ldx, int P, double* Y_dev, long ldy, void* stream) {
    LMC_REQUIRE(op, "null argument");
    if (P == 0) return 0;
    LMC_REQUIRE(X_dev && Y_dev, "null argument");
    LMC_REQUIRE(P >= 0 && ldx >= P && ldy >= P, "bad block shape");
    LMC_REQUIRE(X_dev != Y_dev, "in-place product not supported");
    return op_mvm_rows(op, X_dev, ldx, P, Y_dev, ldy, (cudaStream_t)stream);
}

// Host side of the point-major product.  Every column needs every row, so the pipeline runs over chunks of
// COLUMNS like lmc_mvm_host: a chunk is a strided 2-D copy on the host side (rows of `chunk` doubles out of rows of
// ldx) and a compact [n][chunk] block on the device; H2D copy | product | D2H copy on three streams with
// double-buffered staging.
int lmc_mvm_rows_host(lmc_op* op, const double* X_host, long ldx, int P, double* Y_host, long ldy) {
    LMC_REQUIRE(op, "null argument");
    if (P == 0) return 0;
    LMC_REQUIRE(X_host && Y_host, "null argument");
    LMC_REQUIRE(P >= 0 && ldx >= P && ldy >= P, "bad block shape");
    LMC_REQUIRE(op->Q > 0, "operator parameters not set (call lmc_op_set_params)");
    const long n = op->ps.n;
    LMC_CHECK(cudaDeviceSynchronize());   // shares the grid workspace with the stream-ordered entry points
    // chunk width: 32 columns (one group of 16 pairs for the scatter, 256-byte row pieces for the copy engine)
    // unless that needs more than 512 MB per staging buffer
    int chunk = (int)std::max<long>(2, std::min<long>(32, (512L << 20) / (8 * n)));
    chunk &= ~1;
    if (P <= chunk + 1) chunk = P;        // a block of 32 k + 1 columns keeps its odd column in the last chunk
    const size_t need = (size_t)(chunk + 1) * n;
    if (!op->hs[0]) {
        for (int i = 0; i < 3; ++i) LMC_CHECK(cudaStreamCreateWithFlags(&op->hs[i], cudaStreamNonBlocking));
        for (int i = 0; i < 2; ++i) {
            LMC_CHECK(cudaEventCreateWithFlags(&op->ev_in[i], cudaEventDisableTiming));
            LMC_CHECK(cudaEventCreateWithFlags(&op->ev_cmp[i], cudaEventDisableTiming));
            LMC_CHECK(cudaEventCreateWithFlags(&op->ev_out[i], cudaEventDisableTiming));
        }
    }
    if (need > op->stage_cap) {
        for (int i = 0; i < 2; ++i) {
            cudaFree(op->stage_in[i]); cudaFree(op->stage_out[i]);
            op->stage_in[i] = op->stage_out[i] = nullptr;
        }
        op->stage_cap = 0;
        for (int i = 0; i < 2; ++i) {
            LMC_CHECK(cudaMalloc(&op->stage_in[i], sizeof(double) * need));
            LMC_CHECK(cudaMalloc(&op->stage_out[i], sizeof(double) * need));
        }
        op->stage_cap = need;
    }
    cudaStream_t s_in = op->hs[0], s_cmp = op->hs[1], s_out = op->hs[2];
    int k = 0;
    for (int c0 = 0; c0 < P; ++k) {
        const int b = k & 1;
        int cnt = std::min(chunk, P - c0);
        if (P - c0 - cnt == 1) cnt += 1;   // never leave a single column for a chunk of its own
        if (k >= 2) LMC_CHECK(cudaStreamWaitEvent(s_in, op->ev_cmp[b], 0));    // staging buffer consumed
        LMC_CHECK(cudaMemcpy2DAsync(op->stage_in[b], sizeof(double) * cnt, X_host + c0, sizeof(double) * ldx,
                                    sizeof(double) * cnt, n, cudaMemcpyHostToDevice, s_in));
        LMC_CHECK(cudaEventRecord(op->ev_in[b], s_in));
        LMC_CHECK(cudaStreamWaitEvent(s_cmp, op->ev_in[b], 0));
        if (k >= 2) LMC_CHECK(cudaStreamWaitEvent(s_cmp, op->ev_out[b], 0));  // previous result copied out
        LMC_TRY(op_mvm_rows(op, op->stage_in[b], cnt, cnt, op->stage_out[b], cnt, s_cmp));
        LMC_CHECK(cudaEventRecord(op->ev_cmp[b], s_cmp));
        LMC_CHECK(cudaStreamWaitEvent(s_out, op->ev_cmp[b], 0));
        LMC_CHECK(cudaMemcpy2DAsync(Y_host + c0, sizeof(double) * ldy, op->stage_out[b], sizeof(double) * cnt,
                                    sizeof(double) * cnt, n, cudaMemcpyDeviceToHost, s_out));
        LMC_CHECK(cudaEventRecord(op->ev_out[b], s_out));
        c0 += cnt;
    }
    LMC_CHECK(cudaStreamSynchronize(s_out));
    LMC_CHECK(cudaStreamSynchronize(s_cmp));
    LMC_CHECK(cudaStreamSynchronize(s_in));
    return 0;
}

// Host buffers in, host buffers out.  The block is cut into chunks of columns that flow through a
// 3-stage pipeline (H2D copy | product | D2H copy) on three streams with double-buffered device
// staging, so PCIe traffic in both directions overlaps the kernels.  With pinned host memory the
// copies are truly asynchronous; pageable memory works too (the driver stages it).
int lmc_mvm_host(lmc_op* op, const double* V_host, long ld, int P, double* OUT_host) {
    LMC_REQUIRE(op, "null argument");
    if (P == 0) return 0;
    LMC_REQUIRE(V_host && OUT_host, "null argument");
    LMC_REQUIRE(P >= 0 && ld >= op->ps.n, "bad block shape");
    LMC_REQUIRE(op->Q > 0, "operator parameters not set (call lmc_op_set_params)");
    const long n = op->ps.n;
    // The pipeline runs on the handle's own non-blocking streams but shares its grid workspace with the
    // stream-ordered entry points (lmc_mvm, lmc_minres, ...): wait for whatever they still have in
    // flight, on any stream.  This entry point is synchronous anyway (it returns host data).
    LMC_CHECK(cudaDeviceSynchronize());
    // chunk width: even, ~64 MB per buffer, at least 2 and at most 32 columns
    int chunk = (int)std::max<long>(2, std::min<long>(32, (64L << 20) / (8 * n)));
    chunk &= ~1;
    chunk = std::min(chunk, (P + 1) & ~1);
    const size_t need = (size_t)chunk * n;
    if (!op->hs[0]) {
        for (int i = 0; i < 3; ++i) LMC_CHECK(cudaStreamCreateWithFlags(&op->hs[i], cudaStreamNonBlocking));
        for (int i = 0; i < 2; ++i) {
            LMC_CHECK(cudaEventCreateWithFlags(&op->ev_in[i], cudaEventDisableTiming));
            LMC_CHECK(cudaEventCreateWithFlags(&op->ev_cmp[i], cudaEventDisableTiming));
            LMC_CHECK(cudaEventCreateWithFlags(&op->ev_out[i], cudaEventDisableTiming));
        }
    }
    if (need > op->stage_cap) {
        for (int i = 0; i < 2; ++i) {
            cudaFree(op->stage_in[i]); cudaFree(op->stage_out[i]);
            op->stage_in[i] = op->stage_out[i] = nullptr;
        }
        op->stage_cap = 0;
        for (int i = 0; i < 2; ++i) {
            LMC_CHECK(cudaMalloc(&op->stage_in[i], sizeof(double) * need));
            LMC_CHECK(cudaMalloc(&op->stage_out[i], sizeof(double) * need));
        }
        op->stage_cap = need;
    }
    cudaStream_t s_in = op->hs[0], s_cmp = op->hs[1], s_out = op->hs[2];
    int k = 0;
    for (int c0 = 0; c0 < P; c0 += chunk, ++k) {
        const int b = k & 1;
        const int cnt = std::min(chunk, P - c0);
        if (k >= 2) LMC_CHECK(cudaStreamWaitEvent(s_in, op->ev_cmp[b], 0));    // staging buffer consumed
        LMC_CHECK(cudaMemcpy2DAsync(op->stage_in[b], sizeof(double) * n, V_host + (size_t)c0 * ld,
                                    sizeof(double) * ld, sizeof(double) * n, cnt, cudaMemcpyHostToDevice, s_in));
        LMC_CHECK(cudaEventRecord(op->ev_in[b], s_in));
        LMC_CHECK(cudaStreamWaitEvent(s_cmp, op->ev_in[b], 0));
        if (k >= 2) LMC_CHECK(cudaStreamWaitEvent(s_cmp, op->ev_out[b], 0));  // previous result copied out
        ColumnView cv;
        cv.in = op->stage_in[b]; cv.out = op->stage_out[b]; cv.ld = n; cv.ncols = cnt;
        LMC_TRY(op_mvm(op, cv, s_cmp));
        LMC_CHECK(cudaEventRecord(op->ev_cmp[b], s_cmp));
        LMC_CHECK(cudaStreamWaitEvent(s_out, op->ev_cmp[b], 0));
        LMC_CHECK(cudaMemcpy2DAsync(OUT_host + (size_t)c0 * ld, sizeof(double) * ld, op->stage_out[b],
                                    sizeof(double) * n, sizeof(double) * n, cnt, cudaMemcpyDeviceToHost, s_out));
        LMC_CHECK(cudaEventRecord(op->ev_out[b], s_out));
    }
    LMC_CHECK(cudaStreamSynchronize(s_out));
    LMC_CHECK(cudaStreamSynchronize(s_cmp));
    LMC_CHECK(cudaStreamSynchronize(s_in));
    return 0;
}

int lmc_to_grid(lmc_op* op, const double* V_dev, long ld, int P, double* G_dev, void* stream) {
    LMC_REQUIRE(op && V_dev && G_dev, "null argument");
    LMC_REQUIRE(P >= 0 && ld >= op->ps.n, "bad block shape");
    LMC_TRY(op_ensure_workspace(op));
    cudaStream_t st = (cudaStream_t)stream;
    const int npairs = (P + 1) / 2;
    const long gm = (long)op->D * op->emb.cells;
    for (int p0 = 0; p0 < npairs; p0 += op->tile_pairs) {
        const int cnt = std::min(op->tile_pairs, npairs - p0);
        ColumnView cv;
        cv.in = V_dev + (long)2 * p0 * ld; cv.ld = ld; cv.ncols = std::min(2 * cnt, P - 2 * p0);
        LMC_TRY(to_grid(op->ps, cv, op->G, st));
        LMC_TRY(unpack_pairs(op->G, op->emb.grid_pitch, G_dev + (long)2 * p0 * gm, cv.ncols, op->D,
                             op->emb.cells, st));
    }
    return 0;
}

int lmc_grid_mvm(lmc_op* op, const double* GIN_dev, int P, double* GOUT_dev, void* stream) {
    LMC_REQUIRE(op && GIN_dev && GOUT_dev, "null argument");
    LMC_REQUIRE(op->Q > 0, "operator parameters not set");
    LMC_TRY(op_ensure_workspace(op));
    cudaStream_t st = (cudaStream_t)stream;
    const int npairs = (P + 1) / 2;
    const long gm = (long)op->D * op->emb.cells;
    for (int p0 = 0; p0 < npairs; p0 += op->tile_pairs) {
        const int cnt = std::min(op->tile_pairs, npairs - p0);
        const int ncols = std::min(2 * cnt, P - 2 * p0);
        LMC_TRY(pack_pairs(GIN_dev + (long)2 * p0 * gm, ncols, op->D, op->emb.cells, op->G,
                           op->emb.grid_pitch, st));
        LMC_TRY(op_grid_block(op, op->G, cnt, st));
        LMC_TRY(unpack_pairs(op->G, op->emb.grid_pitch, GOUT_dev + (long)2 * p0 * gm, ncols, op->D,
                             op->emb.cells, st));
    }
    return 0;
}

int lmc_from_grid(lmc_op* op, const double* G_dev, int P, double* OUT_dev, long ld, void* stream) {
    LMC_REQUIRE(op && G_dev && OUT_dev, "null argument");
    LMC_REQUIRE(P >= 0 && ld >= op->ps.n, "bad block shape");
    LMC_TRY(op_ensure_workspace(op));
    cudaStream_t st = (cudaStream_t)stream;
    const int npairs = (P + 1) / 2;
    const long gm = (long)op->D * op->emb.cells;
    for (int p0 = 0; p0 < npairs; p0 += op->tile_pairs) {
        const int cnt = std::min(op->tile_pairs, npairs - p0);
        const int ncols = std::min(2 * cnt, P - 2 * p0);
        LMC_TRY(pack_pairs(G_dev + (long)2 * p0 * gm, ncols, op->D, op->emb.cells, op->G,
                           op->emb.grid_pitch, st));
        ColumnView cv;
        cv.out = OUT_dev + (long)2 * p0 * ld; cv.ld = ld; cv.ncols = ncols;
        LMC_TRY(from_grid(op->ps, cv, op->G, nullptr, st));
    }
    return 0;
}

int lmc_minres(lmc_op* op, const double* RHS_dev, long ld, int P, double* X_dev, double tol,
               int maxiter, int check_every, int* iters_host, double* resid_host, int* istop_host,
               void* stream) {
    LMC_REQUIRE(op && RHS_dev && X_dev, "null argument");
    return minres_solve(op, RHS_dev, ld, P, X_dev, tol, maxiter, check_every, iters_host, resid_host,
                        istop_host, (cudaStream_t)stream);
}

int lmc_minres_lanczos(lmc_op* op, const double* RHS_dev, long ld, int P, double* X_dev, double tol,
                       int maxiter, int check_every, int* iters_host, double* resid_host, int* istop_host,
                       int k, double* tridiag_host, double* beta1_host, void* stream) {
    LMC_REQUIRE(op && RHS_dev && X_dev && tridiag_host, "null argument");
    return minres_solve_lanczos(op, RHS_dev, ld, P, X_dev, tol, maxiter, check_every, iters_host, resid_host,
                                istop_host, k, tridiag_host, beta1_host, (cudaStream_t)stream);
}

// 1 / diag(K~) of the current parameters in sorted order, computed on first use after a parameter update
static int ensure_jacobi(lmc_op* op, cudaStream_t st) {
    if (op->jacobi_valid) return 0;
    if (!op->jacobi) LMC_CHECK(cudaMalloc(&op->jacobi, sizeof(double) * (size_t)std::max<long>(op->ps.n, 1)));
    LMC_TRY(op_jacobi(op, nullptr, op->jacobi, st));
    LMC_CHECK(cudaStreamSynchronize(st));
    op->jacobi_valid = true;
    return 0;
}

int lmc_op_diagonal(lmc_op* op, double* diag_host) {
    LMC_REQUIRE(op && diag_host, "null argument");
    const long n = op->ps.n;
    if (n == 0) return 0;
    HostBlock sorted, caller;
    LMC_CHECK(cudaMalloc(&sorted.p, sizeof(double) * n));
    LMC_CHECK(cudaMalloc(&caller.p, sizeof(double) * n));
    LMC_TRY(op_jacobi(op, sorted.p, nullptr, nullptr));
    if (op->ps.identity) {
        LMC_CHECK(cudaMemcpy(diag_host, sorted.p, sizeof(double) * n, cudaMemcpyDeviceToHost));
    } else {
        LMC_TRY(permute_cols(op->ps, false, sorted.p, n, 1, caller.p, n, nullptr));
        LMC_CHECK(cudaMemcpy(diag_host, caller.p, sizeof(double) * n, cudaMemcpyDeviceToHost));
    }
    return 0;
}

int lmc_minres_pre(lmc_op* op, const double* RHS_dev, long ld, int P, double* X_dev, double tol, int maxiter,
                   int check_every, int precond, int* iters_host, double* resid_host, int* istop_host,
                   void* stream) {
    LMC_REQUIRE(op && RHS_dev && X_dev, "null argument");
    LMC_REQUIRE(precond == LMC_PRECOND_NONE || precond == LMC_PRECOND_JACOBI, "unknown preconditioner");
    const double* jac = nullptr;
    if (precond == LMC_PRECOND_JACOBI) {
        LMC_REQUIRE(op->Q > 0, "operator parameters not set");
        LMC_TRY(ensure_jacobi(op, (cudaStream_t)stream));
        jac = op->jacobi;
    }
    return minres_solve(op, RHS_dev, ld, P, X_dev, tol, maxiter, check_every, iters_host, resid_host,
                        istop_host, (cudaStream_t)stream, jac);
}

int lmc_minres_host(lmc_op* op, const double* RHS_host, long ld, int P, double* X_host, double tol,
                    int maxiter, int check_every, int* iters_host, double* resid_host, int* istop_host) {
    LMC_REQUIRE(op && RHS_host && X_host, "null argument");
    LMC_REQUIRE(P >= 1 && ld >= op->ps.n, "bad block shape");
    HostBlock in, out;
    const size_t bytes = sizeof(double) * (size_t)P * ld;
    LMC_CHECK(cudaMalloc(&in.p, bytes));
    LMC_CHECK(cudaMalloc(&out.p, bytes));
    LMC_CHECK(cudaMemcpy(in.p, RHS_host, bytes, cudaMemcpyHostToDevice));
    LMC_CHECK(cudaMemset(out.p, 0, bytes));
    LMC_TRY(minres_solve(op, in.p, ld, P, out.p, tol, maxiter, check_every, iters_host, resid_host,
                         istop_host, nullptr));
    LMC_CHECK(cudaMemcpy(X_host, out.p, bytes, cudaMemcpyDeviceToHost));
    return 0;
}

int lmc_cg(lmc_op* op, const double* RHS_dev, long ld, int P, double* X_dev, double tol, int maxiter,
           int check_every, int* iters_host, double* resid_host, int* info_host, void* stream) {
    LMC_REQUIRE(op && RHS_dev && X_dev, "null argument");
    return cg_solve(op, RHS_dev, ld, P, X_dev, tol, maxiter, check_every, iters_host, resid_host, info_host,
                    (cudaStream_t)stream);
}

int lmc_cg_host(lmc_op* op, const double* RHS_host, long ld, int P, double* X_host, double tol, int maxiter,
                int check_every, int* iters_host, double* resid_host, int* info_host) {
    LMC_REQUIRE(op && RHS_host && X_host, "null argument");
    LMC_REQUIRE(P >= 1 && ld >= op->ps.n, "bad block shape");
    HostBlock in, out;
    const size_t bytes = sizeof(double) * (size_t)P * ld;
    LMC_CHECK(cudaMalloc(&in.p, bytes));
    LMC_CHECK(cudaMalloc(&out.p, bytes));
    LMC_CHECK(cudaMemcpy(in.p, RHS_host, bytes, cudaMemcpyHostToDevice));
    LMC_CHECK(cudaMemset(out.p, 0, bytes));
    LMC_TRY(cg_solve(op, in.p, ld, P, out.p, tol, maxiter, check_every, iters_host, resid_host, info_host,
                     nullptr));
    LMC_CHECK(cudaMemcpy(X_host, out.p, bytes, cudaMemcpyDeviceToHost));
    return 0;
}

int lmc_grad_grams(lmc_op* op, const double* alpha_dev, const double* R_dev, const double* RINV_dev,
                   long ld, int N, int ntops_extra, const double* tops_extra_host, double* quad_host,
                   double* trace_host, double* nquad_host, double* ntrace_host, void* stream) {
    LMC_REQUIRE(op && alpha_dev && quad_host && trace_host && nquad_host && ntrace_host, "null argument");
    LMC_REQUIRE(N == 0 || (R_dev && RINV_dev), "null probe blocks");
    LMC_REQUIRE(ntops_extra == 0 || tops_extra_host, "null derivative tops");
    return grad_grams(op, alpha_dev, R_dev, RINV_dev, ld, N, ntops_extra, tops_extra_host, false, quad_host,
                      trace_host, nquad_host, ntrace_host, (cudaStream_t)stream);
}

}  // extern "C"
