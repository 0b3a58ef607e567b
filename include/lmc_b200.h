/* lmc_b200.h -- C ABI of the B200-native SKI-LMC hot path.
 *
 * The reference (vlad17/runlmc) is pure Python and has no FFI; the seam this
 * library plugs into is the duck-typed operator protocol of runlmc/linalg
 * (Matrix.matvec / matmat, matrix.py:43-67), Iterative.solve
 * (approx/iterative.py:24) and StochasticDerivService/StochasticDeriv
 * (lmc/stochastic_deriv.py:27-78).  Each entry point below names the
 * reference code it replaces.  INTEGRATION.md shows the ctypes binding.
 *
 * Conventions
 *  - every function returns 0 on success, 1 for an invalid argument, 2 for a
 *    CUDA failure; lmc_last_error() gives the text.  Non-convergence of MINRES
 *    is NOT an error (reference iterative.py:54-58 only logs).
 *  - all floating point data is float64.  A block of P vectors of length n is
 *    stored vector-major: V[p*ld + i] (ld >= n), i.e. each right-hand side is
 *    contiguous, like the rows of the reference's probe matrix `rs`
 *    (stochastic_deriv.py:35).
 *  - pointers named *_dev are device pointers on the current CUDA device,
 *    *_host are host pointers.  `stream` is a cudaStream_t passed as void*
 *    (NULL = default stream).  Device entry points are stream-ordered and do not
 *    synchronise unless stated.
 *  - vectors of length n are in the caller's point order: outputs concatenated,
 *    np.hstack(Ys) (reference lmc/likelihood.py:30).
 *  - one CUDA device per process (the model is one process per GPU; kernel attributes
 *    are set once per process).  Handles own all their device memory, including the
 *    solver state kept between solves, and are independent of each other: different
 *    handles may be driven from different host threads, one handle from one thread at a
 *    time.  lmc_last_error() is per host thread; the launch counter and the profiling
 *    switch are process-wide diagnostics.
 */
#ifndef LMC_B200_H
#define LMC_B200_H

#ifdef __cplusplus
extern "C" {
#endif

typedef struct lmc_op lmc_op;     /* fused SKI-LMC operator  K~ = W (sum_q B_q (x) T_q) W^T + diag(noise) */
typedef struct lmc_bttb lmc_bttb; /* stand-alone symmetric (block-)Toeplitz operator */

int lmc_version(void);
const char* lmc_last_error(void);
/* number of CUDA kernels launched by this library so far in this process */
unsigned long long lmc_launch_count(void);

/* Measured dense fp64 FMA rate of the current GPU in TFLOP/s (a register-resident DFMA
 * microbenchmark, best of 5): the compute roofline bench.py holds the FFT + mix stage against. */
int lmc_fp64_peak(double* tflops_host);

/* Optional per-kernel-family device timing (CUDA events on the launching stream).
 * lmc_profile_begin() starts recording; lmc_profile_end() synchronises and returns,
 * per family, total milliseconds and launch counts (arrays of lmc_profile_ncat()). */
int lmc_profile_ncat(void);
const char* lmc_profile_name(int cat);
int lmc_profile_begin(void);
int lmc_profile_end(double* ms, int* counts);

/* ---- fused operator ------------------------------------------------------
 * Replaces the operator tree built by gen_grid_kernel (lmc/grid_kernel.py:49-74):
 * SumMatrix([GridKernel(SKI(sum|bt|slfm grid kernel, W, W^T)), Diag(noise)]).
 * X_host: [n][ndim] coordinates, outputs concatenated; lens[D] points per output.
 * The grid along axis p is origin[p] + k*delta[p], k < grid_sizes[p]
 * (origin = grid[0], delta = grid[1]-grid[0], as interpolation.py:98).
 * ndim is 1 or 2, D <= 16.  The X-dependent sort is done here, once.          */
int lmc_op_create(lmc_op** out, int D, int ndim, const int* grid_sizes, const double* origin,
                  const double* delta, const int* lens, const double* X_host);
/* Same operator from coordinates that already are on the device (X_dev [n][ndim]): base indices,
 * fractional offsets, the stable sort by (output, grid bin) and the bin offsets are computed on the
 * device -- the work of multi_interpolant (interpolation.py:119-176; 4 s on the host at n = 1M) without
 * building a CSR.  The result is identical, bit for bit, to lmc_op_create on the same coordinates.
 * Synchronises `stream` before returning.                                                          */
int lmc_op_create_dev(lmc_op** out, int D, int ndim, const int* grid_sizes, const double* origin,
                      const double* delta, const int* lens, const double* X_dev, void* stream);
int lmc_op_destroy(lmc_op* op);
/* Per-optimiser-step update.  tops_host [Q][prod grid_sizes]: kernel values
 * k_q(|z - z_0|) on the grid (grid_kernel.py:26-27); B_host [Q][D][D]:
 * coregionalisation matrices (functional_kernel.py:280-287); noise_host [D].    */
int lmc_op_set_params(lmc_op* op, int Q, const double* tops_host, const double* B_host,
                      const double* noise_host);
/* Optional, after lmc_op_set_params: the LMC factors B_q = A_q^T A_q + diag(kappa_q)
 * (coreg_vecs / coreg_diags of functional_kernel.py:280-287; ranks[Q], A [sum ranks][D],
 * kappa [Q][D]).  They must reproduce the dense B; the spectral mix then costs O(R D) instead
 * of O(D^2) per frequency bin.  Results are unchanged up to rounding.                         */
int lmc_op_set_coreg_factors(lmc_op* op, const int* ranks_host, const double* A_host,
                             const double* kappa_host);
/* Per-optimiser-step update with the kernel values evaluated on the device instead of uploaded:
 * kinds_host[Q] in LMC_KERN_*, kparams_host[Q][2] = (inv_lengthscale, period; period ignored unless
 * periodic).  k_q is evaluated at r = ||z - z_0|| of the operator's own grid (interpolated_llgp.py:431):
 *   LMC_KERN_RBF           exp(-gamma r^2 / 2)                        (kern/rbf.py:39-40)
 *   LMC_KERN_MATERN32      (1 + sqrt3 gamma r) exp(-sqrt3 gamma r)    (kern/matern32.py:39-41)
 *   LMC_KERN_STD_PERIODIC  exp(-gamma sin^2(pi r / T) / 2)            (kern/std_periodic.py:44-48)
 * followed by everything lmc_op_set_params does.  The descriptors are remembered for
 * lmc_grad_grams_kernels; lmc_op_set_params forgets them.                                          */
enum { LMC_KERN_RBF = 0, LMC_KERN_MATERN32 = 1, LMC_KERN_STD_PERIODIC = 2 };
int lmc_op_set_kernels(lmc_op* op, int Q, const int* kinds_host, const double* kparams_host,
                       const double* B_host, const double* noise_host);
/* The tops the device evaluates for the remembered descriptors, copied to the host (tests,
 * materialized_kernels): deriv == 0: the Q kernels [Q][m]; deriv != 0: their parameter derivatives
 * [sum_q p_q][m] in kernel order (kernel_gradient of kern/rbf.py:50-54, matern32.py:50-57,
 * std_periodic.py:58-67).  lmc_op_num_kernel_tops gives the first dimension.                       */
int lmc_op_num_kernel_tops(const lmc_op* op, int deriv);
int lmc_op_kernel_tops(lmc_op* op, int deriv, double* tops_host);
long lmc_op_n(const lmc_op* op);          /* total points */
long lmc_op_grid_cells(const lmc_op* op); /* m = prod grid_sizes */
long lmc_op_embed_bins(const lmc_op* op); /* prod of the power-of-two embedding sizes */
int lmc_op_max_tile_points(const lmc_op* op); /* 2-D: most points staged by one scatter tile (diagnostic) */
/* sorted position -> caller index (int32[n]); the solver keeps its state in this order */
int lmc_op_perm(const lmc_op* op, int* perm_host);

/* OUT = K~ V  (SumMatrix.matvec, sum_matrix.py:31-32, over the whole tree)     */
int lmc_mvm(lmc_op* op, const double* V_dev, long ld, int P, double* OUT_dev, void* stream);
int lmc_mvm_host(lmc_op* op, const double* V_host, long ld, int P, double* OUT_host);
/* Same product with V and OUT in the operator's own point order (sorted by output and grid cell:
 * element i is point lmc_op_perm()[i] of the caller).  This is the layout the batched solver
 * keeps its state in; every access is then coalesced.                                          */
int lmc_mvm_sorted(lmc_op* op, const double* V_dev, long ld, int P, double* OUT_dev, void* stream);
/* Same product on a point-major block: X[i * ldx + c], Y[i * ldy + c] for point i (caller's order) and
 * column c < P -- numpy's C order for the [n, P] argument of Matrix.matmat (linalg/matrix.py:27-41), so a
 * caller holding such an array passes it as it is.  A point's columns are contiguous there: the scatter
 * stages them straight from the caller's rows and the permutation into the operator's point order costs
 * no pass of its own (column-major blocks pay two, lmc_mvm).  ldx, ldy >= P.  lmc_mvm_rows_host pipelines
 * chunks of columns (strided 2-D copies on the host side) through copy-in | product | copy-out.        */
int lmc_mvm_rows(lmc_op* op, const double* X_dev, long ldx, int P, double* Y_dev, long ldy, void* stream);
int lmc_mvm_rows_host(lmc_op* op, const double* X_host, long ldx, int P, double* Y_host, long ldy);
/* unit-testable stages.  Grid vectors are [P][D*m], output-major (likelihood.py:30):
 *   lmc_to_grid    G = W^T V      (WT.dot, ski.py:16)
 *   lmc_grid_mvm   GOUT = (sum_q B_q (x) T_q) GIN   (grid_kernel.py:126-136)
 *   lmc_from_grid  OUT = W G      (W.dot, ski.py:14)                           */
int lmc_to_grid(lmc_op* op, const double* V_dev, long ld, int P, double* G_dev, void* stream);
int lmc_grid_mvm(lmc_op* op, const double* GIN_dev, int P, double* GOUT_dev, void* stream);
int lmc_from_grid(lmc_op* op, const double* G_dev, int P, double* OUT_dev, long ld, void* stream);

/* ---- batched MINRES --------------------------------------------------------
 * Replaces pool.starmap(Iterative.solve, [(K, rhs_p, True, True, tol)...])
 * (stochastic_deriv.py:39-52): scipy's Paige-Saunders recurrence and stopping
 * rules per column with rtol = min(1e-10, tol), maxiter (reference: n), plus the
 * reference's true-residual test ||b - K x||_2 < tol on every `check_every`-th
 * iteration (iterative.py:36-42, 100 in the reference).
 * iters[P]  <- number of iterations (callback count) per column
 * resid[P]  <- final ||b - K x||_2 per column (iterative.py:53)
 * istop[P]  <- scipy istop code, or 10 if stopped by the residual test         */
int lmc_minres(lmc_op* op, const double* RHS_dev, long ld, int P, double* X_dev, double tol,
               int maxiter, int check_every, int* iters_host, double* resid_host,
               int* istop_host, void* stream);
int lmc_minres_host(lmc_op* op, const double* RHS_host, long ld, int P, double* X_host,
                    double tol, int maxiter, int check_every, int* iters_host,
                    double* resid_host, int* istop_host);

/* ---- preconditioned MINRES ----------------------------------------------------
 * The reference forwards an optional K.preconditioner to scipy as M (iterative.py:47-50: an SPD
 * approximation of the INVERSE, y = M r); nothing in runlmc builds one.  The block solver runs scipy's
 * preconditioned recurrence (minres.py:252-316: y = M r2, beta = sqrt(r2 . y), v = y / beta) per column;
 * istop = 9 stands for scipy's ValueError on an indefinite M (the column keeps its last iterate).
 * lmc_minres_pre: `precond` selects a preconditioner the library builds itself from the operator:
 *   LMC_PRECOND_JACOBI  M = diag(K~)^-1, the exact diagonal (lmc_op_diagonal), refreshed after every
 *                       parameter update.
 * lmc_op_diagonal: diag(K~) in the caller's point order, host buffer [n].
 * lmc_minres_generic_pre: lmc_minres_generic with a second callback that must leave M * scratch_in in
 * scratch_out (any SPD operator composed by the caller).                                            */
#define LMC_PRECOND_NONE 0
#define LMC_PRECOND_JACOBI 1
int lmc_minres_pre(lmc_op* op, const double* RHS_dev, long ld, int P, double* X_dev, double tol,
                   int maxiter, int check_every, int precond, int* iters_host, double* resid_host,
                   int* istop_host, void* stream);
int lmc_op_diagonal(lmc_op* op, double* diag_host);
int lmc_minres_generic_pre(int (*apply_cb)(void*), int (*precond_cb)(void*), void* ctx, long n,
                           double* scratch_in_dev, double* scratch_out_dev, const double* RHS_dev, long ld,
                           int P, double* X_dev, double tol, int maxiter, int check_every, int* iters_host,
                           double* resid_host, int* istop_host, void* stream);

/* ---- log-determinant by-product ------------------------------------------------
 * The reference's roadmap (README.md:88-93, "Lanczos") and its model's log_det_K (models/
 * interpolated_llgp.py:262-276, a dense Cholesky) ask for a matrix-free log det K~.  MINRES is a Lanczos
 * process: lmc_minres_lanczos is lmc_minres that also hands back the tridiagonal of every column,
 * tridiag_host[c][j] = (alfa_{j+1}, beta_{j+2}) for the first min(k, iterations of c) iterations (zero
 * beyond), and beta1_host[c] = ||b_c||.  With Rademacher right-hand sides z_c,
 *   z_c^T log(K~) z_c ~= beta1_c^2 * e1^T log(T_c) e1
 * (stochastic Lanczos quadrature), so the probe solves of one gradient evaluation also estimate
 * log det K~ = E[z^T log(K~) z] at no extra product (runlmc_b200/approx/logdet.py).               */
int lmc_minres_lanczos(lmc_op* op, const double* RHS_dev, long ld, int P, double* X_dev, double tol,
                       int maxiter, int check_every, int* iters_host, double* resid_host, int* istop_host,
                       int k, double* tridiag_host, double* beta1_host, void* stream);

/* ---- batched conjugate gradients ---------------------------------------------
 * Iterative.solve(..., minres=False) (iterative.py:44-51): scipy.sparse.linalg.cg with M = I, x0 = 0,
 * rtol = min(1e-10, tol), atol = 0, maxiter, behind the same true-residual test every `check_every`-th
 * iteration.  iters/resid as for lmc_minres; info[P] <- scipy's info (0 converged, maxiter if the
 * iterations ran out), or 10 if stopped by the residual test.                                      */
int lmc_cg(lmc_op* op, const double* RHS_dev, long ld, int P, double* X_dev, double tol, int maxiter,
           int check_every, int* iters_host, double* resid_host, int* info_host, void* stream);
int lmc_cg_host(lmc_op* op, const double* RHS_host, long ld, int P, double* X_host, double tol, int maxiter,
                int check_every, int* iters_host, double* resid_host, int* info_host);
int lmc_cg_generic(int (*apply_cb)(void*), void* ctx, long n, double* scratch_in_dev, double* scratch_out_dev,
                   const double* RHS_dev, long ld, int P, double* X_dev, double tol, int maxiter,
                   int check_every, int* iters_host, double* resid_host, int* info_host, void* stream);

/* Same solver for an operator tree composed by the caller (the runlmc.linalg mirror
 * classes): for every product the solver writes the [P][n] input block to
 * scratch_in_dev, calls apply_cb(ctx) -- which must leave K * scratch_in in
 * scratch_out_dev, on the same stream -- and reads scratch_out_dev.            */
int lmc_minres_generic(int (*apply_cb)(void*), void* ctx, long n, double* scratch_in_dev,
                       double* scratch_out_dev, const double* RHS_dev, long ld, int P, double* X_dev,
                       double tol, int maxiter, int check_every, int* iters_host, double* resid_host,
                       int* istop_host, void* stream);
/* out = sum_c sum_i A[c][i] * B[c][i]  (the dot products of StochasticDeriv, stochastic_deriv.py:69-78) */
int lmc_block_dot(const double* A_dev, long lda, const double* B_dev, long ldb, long n, int ncols,
                  double* out_host, void* stream);

/* ---- gradient contractions -------------------------------------------------
 * Everything StochasticDeriv.derivative (stochastic_deriv.py:69-78) needs for all
 * hyper-parameters of ApproxLMCLikelihood (likelihood.py:48-96) at once.
 * For each "top" t (the Q kernels first, then ntops_extra derivative tops
 * d k_q / d theta, likelihood.py:121-124) returns the D x D Gram matrices
 *   quad [t][d][e] = (W^T alpha)_d^T  T_t (W^T alpha)_e
 *   trace[t][d][e] = sum_i (W^T Kinv r_i)_d^T T_t (W^T r_i)_e      (sum over N probes)
 * and per output   nquad[d] = sum_{i in d} alpha_i^2,  ntrace[d] = sum_probes sum_{i in d} (Kinv r)_i r_i.
 * dK for A_q[r,j], kappa_q[i], kernel parameters and noise are then
 * <C, Gram> Frobenius products formed on the host (see runlmc_b200/lmc).       */
int lmc_grad_grams(lmc_op* op, const double* alpha_dev, const double* R_dev, const double* RINV_dev,
                   long ld, int N, int ntops_extra, const double* tops_extra_host,
                   double* quad_host, double* trace_host, double* nquad_host, double* ntrace_host,
                   void* stream);

/* Same with the derivative tops d k_q / d theta evaluated on the device from the descriptors of
 * lmc_op_set_kernels: T = Q + lmc_op_num_kernel_tops(op, 1).                                       */
int lmc_grad_grams_kernels(lmc_op* op, const double* alpha_dev, const double* R_dev, const double* RINV_dev,
                           long ld, int N, double* quad_host, double* trace_host, double* nquad_host,
                           double* ntrace_host, void* stream);

/* ---- stand-alone structured operators (runlmc/linalg mirror) ----------------
 * lmc_bttb: BTTB(top, sizes).matvec (bttb.py:91-148), ndim <= 3; ndim == 1 also
 * serves Toeplitz(top).matvec (toeplitz.py:34-67).  X/Y are [k][m] blocks.     */
int lmc_bttb_create(lmc_bttb** out, int ndim, const int* sizes, const double* top_host);
int lmc_bttb_destroy(lmc_bttb* h);
int lmc_bttb_apply(lmc_bttb* h, const double* X_dev, int k, double* Y_dev, void* stream);
/* Y[k][r][i] = sum_c A[r][c] X[k][c][i], small dense A on device (NumpyMatrix.matmat,
 * numpy_matrix.py:30-31; inner > 1 is the A-side contraction of Kronecker.matvec, kronecker.py:39-46) */
int lmc_dense_apply(const double* A_dev, int rows, int cols, const double* X_dev, int k, int inner,
                    double* Y_dev, void* stream);
/* Y[k][b][a] = X[k][a][b]  (the reshape/transposes of Kronecker.matvec, kronecker.py:41-44) */
int lmc_transpose(const double* X_dev, int k, int a, int b, double* Y_dev, void* stream);
/* Y[k][rows] = CSR(rows x cols) * X[k][cols]   (scipy CSR .dot used by SKI, ski.py:14-16) */
int lmc_csr_apply(int rows, const int* indptr_dev, const int* indices_dev, const double* data_dev,
                  const double* X_dev, long ldx, int k, double* Y_dev, long ldy, void* stream);
/* Y = a*X + b*Y elementwise over len entries (SumMatrix accumulation) */
int lmc_axpby(long len, double a, const double* X_dev, double b, double* Y_dev, void* stream);
/* Y[k][len] = v[len] * X[k][len]  (Diag.matmat, diag.py:27-28) */
int lmc_diag_apply(const double* v_dev, long len, const double* X_dev, int k, double* Y_dev, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* LMC_B200_H */
