// Multi-right-hand-side MINRES on the fused SKI-LMC operator.
//
// Replaces pool.starmap(Iterative.solve, ...) of the reference
// (runlmc/lmc/stochastic_deriv.py:39-52 -> runlmc/approx/iterative.py:24-62 ->
// scipy.sparse.linalg.minres, scipy 1.18.1 _isolve/minres.py:210-364).
// Every column runs scipy's Paige-Saunders recurrence with scipy's stopping
// rules (rtol = min(1e-10, tol)) and the reference wrapper's true-residual test
// every `check_every` iterations; columns stop independently (frozen once
// stopped) while the block advances in lock step.
//
// State lives in the operator's sorted point order, vector-major [P][n]; the
// Lanczos vector v is never materialised (v = r2 / beta is applied as a column
// scale when r2 is read).  r1/r2/y and w1/w2/w rotate by pointer.  Dot products
// are two-stage (per-CTA partials, fixed-order final sum) => deterministic.
#include "op.cuh"

#include <cfloat>
#include <cmath>
#include <cstdlib>
#include <vector>

namespace lmc {

struct ColState {
    double beta1, beta, oldb, dbar, epsln, phibar, rhs1, rhs2, tnorm2, gmax, gmin, cs, sn;
    double alfa, root;
    double c_r1;                                   // beta / oldb for the next Lanczos step
    double oldeps, delta, denom, phi, inv_oldb;    // coefficients of the w / x update
    double resid;
    int istop, itn, done;
};

static const int kVecThreads = 256;
static const int kVecPerThread = 8;
static const int kVecChunk = kVecThreads * kVecPerThread;

__device__ __forceinline__ double block_reduce_sum(double v) {
    __shared__ double s_red[32];
    __syncthreads();
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    if (lane == 0) s_red[wid] = v;
    __syncthreads();
    const int nw = (blockDim.x + 31) >> 5;
    v = (threadIdx.x < nw) ? s_red[threadIdx.x] : 0.0;
    if (wid == 0) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
    }
    if (threadIdx.x == 0) s_red[0] = v;
    __syncthreads();
    return s_red[0];
}

// fixed-order sum of `cnt` partials by the whole block (same value in every CTA)
__device__ __forceinline__ double block_sum_partials(const double* part, int cnt) {
    double v = 0.0;
    for (int i = threadIdx.x; i < cnt; i += blockDim.x) v += part[i];
    return block_reduce_sum(v);
}

// b_sorted[c][i] = RHS[c][perm[i]];  r2 = b; x = w = w1 = w2 = 0; partial ||b||^2
__global__ void __launch_bounds__(kVecThreads) minres_init_kernel(
    const double* __restrict__ RHS, long ld, const int* __restrict__ perm, long n, double* b,
    double* r2, double* x, double* w, double* w1, double* w2, double* part, int nblk) {
    const int col = blockIdx.y;
    const long base = (long)blockIdx.x * kVecChunk;
    double acc = 0.0;
    for (int k = 0; k < kVecPerThread; ++k) {
        const long i = base + k * kVecThreads + threadIdx.x;
        if (i < n) {
            const double v = RHS[(long)col * ld + (perm ? perm[i] : i)];
            const long o = (long)col * n + i;
            b[o] = v; r2[o] = v; x[o] = 0.0; w[o] = 0.0; w1[o] = 0.0; w2[o] = 0.0;
            acc = fma(v, v, acc);
        }
    }
    acc = block_reduce_sum(acc);
    if (threadIdx.x == 0) part[(long)col * nblk + blockIdx.x] = acc;
}

// part_bb: partials of ||b||^2 (scipy returns x = 0 for b = 0); part: partials of beta1^2 = b . M b (the
// same array without a preconditioner)
__global__ void minres_init_scalars_kernel(ColState* st, double* inv_beta, int* active, const double* part_bb,
                                           const double* part, int nblk, int P, int* n_active) {
    const int col = blockIdx.x * blockDim.x + threadIdx.x;
    if (col < P) {
        double s = 0.0, bb = 0.0;
        for (int i = 0; i < nblk; ++i) { s += part[(long)col * nblk + i]; bb += part_bb[(long)col * nblk + i]; }
        ColState c = {};
        if (bb == 0.0) s = 0.0;
        if (s < 0.0) { c.istop = 9; s = 0.0; }      // scipy: ValueError('indefinite preconditioner')
        c.beta1 = sqrt(s);
        c.beta = c.beta1; c.oldb = 0.0; c.dbar = 0.0; c.epsln = 0.0; c.phibar = c.beta1;
        c.rhs1 = c.beta1; c.rhs2 = 0.0; c.tnorm2 = 0.0; c.gmax = 0.0; c.gmin = DBL_MAX;
        c.cs = -1.0; c.sn = 0.0; c.c_r1 = 0.0; c.itn = 0;
        c.done = (c.beta1 == 0.0) ? 1 : 0;  // scipy returns x = 0 at once
        st[col] = c;
        inv_beta[col] = c.done ? 0.0 : 1.0 / c.beta1;
        active[col] = c.done ? 0 : 1;
    }
    __syncthreads();
    if (blockIdx.x == 0 && threadIdx.x == 0) *n_active = -1;  // recomputed by the first status kernel
}

// y -= (beta/oldb) r1 ; alfa partial = sum (v/beta) * y   (v = r2, or M r2 with a preconditioner)
__global__ void __launch_bounds__(kVecThreads) minres_k1_kernel(
    double* y, const double* __restrict__ r1, const double* __restrict__ r2_unused, const double* __restrict__ vsrc,
    const ColState* st, const double* inv_beta, const int* active, long n, double* part, int nblk) {
    const int col = blockIdx.y;
    if (!active[col]) return;
    const double c = st[col].c_r1, s = inv_beta[col];
    const long base = (long)blockIdx.x * kVecChunk;
    double acc = 0.0;
    for (int k = 0; k < kVecPerThread; ++k) {
        const long i = base + k * kVecThreads + threadIdx.x;
        if (i < n) {
            const long o = (long)col * n + i;
            double yv = y[o];
            if (c != 0.0) yv = __dsub_rn(yv, __dmul_rn(c, r1[o]));
            y[o] = yv;
            acc = fma(__dmul_rn(s, vsrc[o]), yv, acc);
        }
    }
    acc = block_reduce_sum(acc);
    if (threadIdx.x == 0) part[(long)col * nblk + blockIdx.x] = acc;
}

// alfa = sum partials ; y -= (alfa/beta) r2 ; beta^2 partial = sum y^2
__global__ void __launch_bounds__(kVecThreads) minres_k2_kernel(
    double* y, const double* __restrict__ r2, ColState* st, const int* active, long n,
    const double* part_a, double* part_b, int nblk) {
    const int col = blockIdx.y;
    if (!active[col]) return;
    const double alfa = block_sum_partials(part_a + (long)col * nblk, nblk);
    if (blockIdx.x == 0 && threadIdx.x == 0) st[col].alfa = alfa;
    const double c = alfa / st[col].beta;
    const long base = (long)blockIdx.x * kVecChunk;
    double acc = 0.0;
    for (int k = 0; k < kVecPerThread; ++k) {
        const long i = base + k * kVecThreads + threadIdx.x;
        if (i < n) {
            const long o = (long)col * n + i;
            const double yv = __dsub_rn(y[o], __dmul_rn(c, r2[o]));
            y[o] = yv;
            acc = fma(yv, yv, acc);
        }
    }
    acc = block_reduce_sum(acc);
    if (threadIdx.x == 0) part_b[(long)col * nblk + blockIdx.x] = acc;
}

// fixed-order sum of `cnt` partials by one (small) block: strided per-thread sums, then a tree
__device__ __forceinline__ double column_sum(const double* part, int cnt) {
    double v = 0.0;
    for (int i = threadIdx.x; i < cnt; i += blockDim.x) v += part[i];
    return block_reduce_sum(v);
}

static const int kScalarThreads = 64;   // threads per column in the scalar kernels (one block per column)

// scalar recurrences after the Lanczos step (minres.py:236-283); one block per column so the sum of
// the per-CTA partials is a parallel reduction (a single thread walking ~500 partials cost 40 us)
__global__ void minres_s1_kernel(ColState* st, const int* active, const double* part_b, int nblk, int P,
                                 int* n_active, double* rec, int rec_k) {
    const int col = blockIdx.x;
    if (col == 0 && threadIdx.x == 0) *n_active = 0;   // recounted by the s2 kernel that follows
    if (!active[col]) return;
    const double s = column_sum(part_b + (long)col * nblk, nblk);
    if (threadIdx.x != 0) return;
    ColState c = st[col];
    c.itn += 1;
    c.oldb = c.beta;
    double s_pos = s;
    if (s < 0.0) { c.istop = 9; s_pos = 0.0; }   // r2 . M r2 < 0: scipy raises (indefinite preconditioner)
    c.beta = sqrt(s_pos);
    c.tnorm2 += c.alfa * c.alfa + c.oldb * c.oldb + c.beta * c.beta;
    if (rec && c.itn <= rec_k) {
        // the Lanczos tridiagonal of this column: T[k][k] = alfa_k, T[k][k+1] = beta_{k+1} (k = itn)
        double* r = rec + ((long)col * rec_k + (c.itn - 1)) * 2;
        r[0] = c.alfa;
        r[1] = c.beta;
    }
    if (c.itn == 1 && c.beta / c.beta1 <= 10 * DBL_EPSILON) c.istop = -1;
    c.oldeps = c.epsln;
    c.delta = c.cs * c.dbar + c.sn * c.alfa;
    const double gbar = c.sn * c.dbar - c.cs * c.alfa;
    c.epsln = c.sn * c.beta;
    c.dbar = -c.cs * c.beta;
    c.root = sqrt(gbar * gbar + c.dbar * c.dbar);
    double gamma = sqrt(gbar * gbar + c.beta * c.beta);
    gamma = fmax(gamma, DBL_EPSILON);
    c.cs = gbar / gamma;
    c.sn = c.beta / gamma;
    c.phi = c.cs * c.phibar;
    c.phibar = c.sn * c.phibar;
    c.denom = 1.0 / gamma;
    c.inv_oldb = 1.0 / c.oldb;
    c.gmax = fmax(c.gmax, gamma);
    c.gmin = fmin(c.gmin, gamma);
    const double z = c.rhs1 / gamma;
    c.rhs1 = c.rhs2 - c.delta * z;
    c.rhs2 = -c.epsln * z;
    st[col] = c;
}

// scipy: w1 = w2; w2 = w; w = (v - oldeps*w1 - delta*w2)*denom, i.e. with the two most
// recent directions w_older (= old w2) and w_newer (= old w):
//   w_out = (v - oldeps w_older - delta w_newer) * denom ; x += phi w_out ; ||x||^2 partial
__global__ void __launch_bounds__(kVecThreads) minres_k3_kernel(
    double* __restrict__ w_out, const double* __restrict__ w_older, const double* __restrict__ w_newer,
    const double* __restrict__ vsrc, double* x, const ColState* st, const int* active, long n,
    double* part, int nblk) {
    const int col = blockIdx.y;
    if (!active[col]) return;
    const ColState* c = st + col;
    if (c->istop == 9) return;                   // x keeps the iterate before the failed step
    const double oldeps = c->oldeps, delta = c->delta, denom = c->denom, phi = c->phi, s = c->inv_oldb;
    const long base = (long)blockIdx.x * kVecChunk;
    double acc = 0.0;
    for (int k = 0; k < kVecPerThread; ++k) {
        const long i = base + k * kVecThreads + threadIdx.x;
        if (i < n) {
            const long o = (long)col * n + i;
            const double v = __dmul_rn(s, vsrc[o]);
            double t = __dsub_rn(v, __dmul_rn(oldeps, w_older[o]));
            t = __dsub_rn(t, __dmul_rn(delta, w_newer[o]));
            const double wn = __dmul_rn(t, denom);
            w_out[o] = wn;
            const double xv = __dadd_rn(x[o], __dmul_rn(phi, wn));
            x[o] = xv;
            acc = fma(xv, xv, acc);
        }
    }
    acc = block_reduce_sum(acc);
    if (threadIdx.x == 0) part[(long)col * nblk + blockIdx.x] = acc;
}

// norms + stopping rules (minres.py:285-330); prepares the next step's scales.  One block per
// column; the number of columns still running is an integer count (order independent).
__global__ void minres_s2_kernel(ColState* st, double* inv_beta, int* active, const double* part_c, int nblk,
                                 int P, double rtol, int maxiter, int* n_active) {
    const int col = blockIdx.x;
    if (!active[col]) return;
    const double s = column_sum(part_c + (long)col * nblk, nblk);
    if (threadIdx.x != 0) return;
    ColState c = st[col];
    const double ynorm = sqrt(s);
    const double Anorm = sqrt(c.tnorm2);
    const double epsx = Anorm * ynorm * DBL_EPSILON;
    const double rnorm = c.phibar;
    const double test1 = (ynorm == 0.0 || Anorm == 0.0) ? INFINITY : rnorm / (Anorm * ynorm);
    const double test2 = (Anorm == 0.0) ? INFINITY : c.root / Anorm;
    const double Acond = c.gmax / c.gmin;
    if (c.istop == 0) {
        if (1.0 + test2 <= 1.0) c.istop = 2;
        if (1.0 + test1 <= 1.0) c.istop = 1;
        if (c.itn >= maxiter) c.istop = 6;
        if (Acond >= 0.1 / DBL_EPSILON) c.istop = 4;
        if (epsx >= c.beta1) c.istop = 3;
        if (test2 <= rtol) c.istop = 2;
        if (test1 <= rtol) c.istop = 1;
    }
    c.c_r1 = c.beta / c.oldb;
    if (c.istop != 0) { c.done = 1; active[col] = 0; }
    else { inv_beta[col] = 1.0 / c.beta; atomicAdd(n_active, 1); }
    st[col] = c;
}

// partial ||b - y||^2
__global__ void __launch_bounds__(kVecThreads) minres_resid_kernel(
    const double* __restrict__ b, const double* __restrict__ y, const int* active, long n, double* part,
    int nblk) {
    const int col = blockIdx.y;
    if (active && !active[col]) return;
    const long base = (long)blockIdx.x * kVecChunk;
    double acc = 0.0;
    for (int k = 0; k < kVecPerThread; ++k) {
        const long i = base + k * kVecThreads + threadIdx.x;
        if (i < n) {
            const long o = (long)col * n + i;
            const double dlt = __dsub_rn(b[o], y[o]);
            acc = fma(dlt, dlt, acc);
        }
    }
    acc = block_reduce_sum(acc);
    if (threadIdx.x == 0) part[(long)col * nblk + blockIdx.x] = acc;
}

// reference early termination (iterative.py:36-42): residual < tol stops the column.  One block per
// column; *n_active must be zero on entry of a non-final pass.
__global__ void minres_s3_kernel(ColState* st, int* active, const double* part, int nblk, int P, double tol,
                                 int final_pass, int* n_active) {
    const int col = blockIdx.x;
    if (!final_pass && !active[col]) return;
    const double s = column_sum(part + (long)col * nblk, nblk);
    if (threadIdx.x != 0) return;
    const double r = sqrt(s);
    st[col].resid = r;
    if (!final_pass) {
        if (r < tol) { st[col].istop = 10; st[col].done = 1; active[col] = 0; }
        else atomicAdd(n_active, 1);
    }
}

__global__ void __launch_bounds__(kVecThreads) minres_finish_kernel(
    const double* __restrict__ x, const int* __restrict__ perm, long n, double* X, long ld) {
    const int col = blockIdx.y;
    const long base = (long)blockIdx.x * kVecChunk;
    for (int k = 0; k < kVecPerThread; ++k) {
        const long i = base + k * kVecThreads + threadIdx.x;
        if (i < n) X[(long)col * ld + (perm ? perm[i] : i)] = x[(long)col * n + i];
    }
}

// Y = s * X per column (materialises v = r2 / beta for callback operators)
__global__ void __launch_bounds__(kVecThreads) scale_cols_kernel(const double* __restrict__ X,
                                                                const double* __restrict__ scale, long n,
                                                                double* Y) {
    const int col = blockIdx.y;
    const double s = scale ? scale[col] : 1.0;
    const long base = (long)blockIdx.x * kVecChunk;
    for (int k = 0; k < kVecPerThread; ++k) {
        const long i = base + k * kVecThreads + threadIdx.x;
        if (i < n) Y[(long)col * n + i] = __dmul_rn(s, X[(long)col * n + i]);
    }
}

// Solver workspace: grow-only, owned by whoever owns the operator (the fused handle keeps it between
// solves; a callback operator's lives for one call)
static int reserve_workspace(void** p, size_t* cap, size_t bytes) {
    if (bytes <= *cap) return 0;
    if (*p) { cudaFree(*p); *p = nullptr; *cap = 0; }
    LMC_CHECK(cudaMalloc(p, bytes));
    *cap = bytes;
    return 0;
}

SolverHost::~SolverHost() {
    for (auto& e : ev) if (e) cudaEventDestroy(e);
    if (ev_join) cudaEventDestroy(ev_join);
    if (flags) cudaFreeHost(flags);
    if (stream) cudaStreamDestroy(stream);
}

// pinned flags + events of the asynchronous stop test, and the solver's own stream (stream capture is
// not allowed on the legacy default stream, which is what most callers pass)
static int solver_host_init(SolverHost* h, bool own_stream) {
    if (!h->flags) {
        LMC_CHECK(cudaHostAlloc(&h->flags, sizeof(int) * SolverHost::kRing, cudaHostAllocDefault));
        for (auto& e : h->ev) LMC_CHECK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
        LMC_CHECK(cudaEventCreateWithFlags(&h->ev_join, cudaEventDisableTiming));
    }
    if (own_stream && !h->stream) LMC_CHECK(cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking));
    return 0;
}

// The product the solver iterates with: out = A (in * in_scale) on [P][n] blocks
struct MinresOperator {
    long n = 0;
    const int* perm = nullptr;  // solver order -> caller order (nullptr = identity)
    virtual int apply(const double* in, const double* in_scale, const int* active, double* out, int P,
                      cudaStream_t st) = 0;
    // device memory for the solver's state vectors, valid until the operator goes away
    virtual int workspace(size_t bytes, void** p) = 0;
    virtual SolverHost* host() = 0;
    // every product is a fixed sequence of launches of this library on the given stream, so an
    // iteration can be captured into a CUDA graph and the solver may run on its own stream
    virtual bool capturable() const { return false; }
    virtual ~MinresOperator() {}
};

struct FusedOperator : MinresOperator {
    lmc_op* op;
    explicit FusedOperator(lmc_op* o) : op(o) {
        n = o->ps.n;
        perm = o->ps.identity ? nullptr : o->ps.perm;
    }
    int apply(const double* in, const double* in_scale, const int* active, double* out, int P,
              cudaStream_t st) override {
        ColumnView cv;
        cv.in = in; cv.out = out; cv.ld = n; cv.ncols = P; cv.sorted_in = cv.sorted_out = true;
        cv.in_scale = in_scale; cv.active = active;
        return op_mvm(op, cv, st);
    }
    int workspace(size_t bytes, void** p) override {
        LMC_TRY(reserve_workspace(&op->solver_ws, &op->solver_ws_cap, bytes));
        *p = op->solver_ws;
        return 0;
    }
    SolverHost* host() override { return &op->solver_host; }
    bool capturable() const override { return true; }
};

// Operator trees composed on the Python side: the solver fills `scratch_in`,
// calls back, and reads `scratch_out` (both [P][n] device blocks).
struct CallbackOperator : MinresOperator {
    int (*cb)(void*);
    void* ctx;
    double* scratch_in;
    double* scratch_out;
    void* ws = nullptr;
    size_t ws_cap = 0;
    SolverHost sh;
    ~CallbackOperator() override { cudaFree(ws); }
    int workspace(size_t bytes, void** p) override {
        LMC_TRY(reserve_workspace(&ws, &ws_cap, bytes));
        *p = ws;
        return 0;
    }
    SolverHost* host() override { return &sh; }
    int apply(const double* in, const double* in_scale, const int* active, double* out, int P,
              cudaStream_t st) override {
        (void)active;
        const dim3 grid((unsigned)ceil_div(n, kVecChunk), (unsigned)P);
        scale_cols_kernel<<<grid, kVecThreads, 0, st>>>(in, in_scale, n, scratch_in);
        count_launch();
        LMC_CHECK(cudaGetLastError());
        const int rc = cb(ctx);
        if (rc != 0) { set_error("operator callback failed"); return 3; }
        LMC_CHECK(cudaMemcpyAsync(out, scratch_out, sizeof(double) * (size_t)P * n,
                                  cudaMemcpyDeviceToDevice, st));
        return 0;
    }
};

// The preconditioner M of scipy's minres (an SPD approximation of the inverse; y = M r, minres.py:256
// and :313), forwarded by the reference from K.preconditioner (iterative.py:47-50).
struct Preconditioner {
    virtual int apply(const double* in, const int* active, double* out, long n, int P, cudaStream_t st) = 0;
    virtual bool capturable() const { return false; }
    virtual ~Preconditioner() {}
};

// out[c][i] = d[i] * in[c][i]  (Jacobi: d = 1 / diag K in the solver's point order)
__global__ void __launch_bounds__(kVecThreads) diag_precond_kernel(const double* __restrict__ d,
                                                                  const double* __restrict__ in,
                                                                  const int* active, long n, double* out) {
    const int col = blockIdx.y;
    if (active && !active[col]) return;
    const long base = (long)blockIdx.x * kVecChunk;
    for (int k = 0; k < kVecPerThread; ++k) {
        const long i = base + k * kVecThreads + threadIdx.x;
        if (i < n) out[(long)col * n + i] = __dmul_rn(d[i], in[(long)col * n + i]);
    }
}

struct DiagPreconditioner : Preconditioner {
    const double* d;   // [n] device, solver order
    explicit DiagPreconditioner(const double* dd) : d(dd) {}
    int apply(const double* in, const int* active, double* out, long n, int P, cudaStream_t st) override {
        const dim3 grid((unsigned)ceil_div(n, kVecChunk), (unsigned)P);
        diag_precond_kernel<<<grid, kVecThreads, 0, st>>>(d, in, active, n, out);
        count_launch();
        LMC_CHECK(cudaGetLastError());
        return 0;
    }
    bool capturable() const override { return true; }
};

struct CallbackPreconditioner : Preconditioner {
    int (*cb)(void*);
    void* ctx;
    double* scratch_in;
    double* scratch_out;
    int apply(const double* in, const int* active, double* out, long n, int P, cudaStream_t st) override {
        (void)active;
        LMC_CHECK(cudaMemcpyAsync(scratch_in, in, sizeof(double) * (size_t)P * n, cudaMemcpyDeviceToDevice, st));
        const int rc = cb(ctx);
        if (rc != 0) { set_error("preconditioner callback failed"); return 3; }
        LMC_CHECK(cudaMemcpyAsync(out, scratch_out, sizeof(double) * (size_t)P * n, cudaMemcpyDeviceToDevice, st));
        return 0;
    }
};

// partial sum_i a[c][i] * b[c][i]
__global__ void __launch_bounds__(kVecThreads) minres_dot_kernel(const double* __restrict__ a,
                                                                const double* __restrict__ b, const int* active,
                                                                long n, double* part, int nblk) {
    const int col = blockIdx.y;
    if (active && !active[col]) return;
    const long base = (long)blockIdx.x * kVecChunk;
    double acc = 0.0;
    for (int k = 0; k < kVecPerThread; ++k) {
        const long i = base + k * kVecThreads + threadIdx.x;
        if (i < n) acc = fma(a[(long)col * n + i], b[(long)col * n + i], acc);
    }
    acc = block_reduce_sum(acc);
    if (threadIdx.x == 0) part[(long)col * nblk + blockIdx.x] = acc;
}

struct DevBuf {
    void* p = nullptr;
    ~DevBuf() { if (p) cudaFree(p); }
    int alloc(size_t bytes) { LMC_CHECK(cudaMalloc(&p, bytes ? bytes : 8)); return 0; }
    template <class T> T* as() { return static_cast<T*>(p); }
};

struct WsSlice {
    char* p;
    template <class T> T* as() { return reinterpret_cast<T*>(p); }
};

// Everything one MINRES iteration touches; the vector buffers rotate by pointer, so the launches of
// iteration k depend on k only through (k - 1) mod period (3 without, 6 with a preconditioner).
struct MinresState {
    long n; int P, nblk;
    double *b, *x, *r1, *r2, *y, *wa, *wb, *wc;   // wa = older direction (scipy's w2), wb = newer (w), wc = free
    double *z, *zold;                             // M r2 and the previous one (preconditioned only)
    double *pa, *pb, *pc;
    ColState* cs; double* inv_beta; int* active; int* n_active;
    double rtol; int maxiter;
    double* rec; int rec_k;                       // optional record of the Lanczos coefficients, [P][rec_k][2]
};

static int minres_iteration(MinresOperator& A, Preconditioner* M, MinresState& s, cudaStream_t st) {
    const dim3 vgrid((unsigned)s.nblk, (unsigned)s.P);
    const long n = s.n;
    const double* v = M ? s.z : s.r2;             // v = (M r2) / beta, applied as a column scale
    LMC_TRY(A.apply(v, s.inv_beta, s.active, s.y, s.P, st));
    {
        ProfScope prof(PROF_MINRES_VEC, st);
        minres_k1_kernel<<<vgrid, kVecThreads, 0, st>>>(s.y, s.r1, s.r2, v, s.cs, s.inv_beta, s.active, n, s.pa, s.nblk);
        minres_k2_kernel<<<vgrid, kVecThreads, 0, st>>>(s.y, s.r2, s.cs, s.active, n, s.pa, s.pb, s.nblk);
    }
    { double* t = s.r1; s.r1 = s.r2; s.r2 = s.y; s.y = t; }   // r1 <- r2 <- y ; old r1 buffer is the next y
    count_launch(2);
    if (M) {
        { double* t = s.zold; s.zold = s.z; s.z = t; }
        LMC_TRY(M->apply(s.r2, s.active, s.z, n, s.P, st));
        ProfScope prof(PROF_MINRES_VEC, st);
        minres_dot_kernel<<<vgrid, kVecThreads, 0, st>>>(s.r2, s.z, s.active, n, s.pb, s.nblk);
        count_launch();
    }
    {
        ProfScope prof(PROF_MINRES_SCALAR, st);
        minres_s1_kernel<<<s.P, kScalarThreads, 0, st>>>(s.cs, s.active, s.pb, s.nblk, s.P, s.n_active, s.rec, s.rec_k);
    }
    {
        // v = (previous M r2, or r1) / oldb
        ProfScope prof(PROF_MINRES_VEC, st);
        minres_k3_kernel<<<vgrid, kVecThreads, 0, st>>>(s.wc, s.wa, s.wb, M ? s.zold : s.r1, s.x, s.cs, s.active, n,
                                                        s.pc, s.nblk);
    }
    { double* t = s.wa; s.wa = s.wb; s.wb = s.wc; s.wc = t; }
    {
        ProfScope prof(PROF_MINRES_SCALAR, st);
        minres_s2_kernel<<<s.P, kScalarThreads, 0, st>>>(s.cs, s.inv_beta, s.active, s.pc, s.nblk, s.P, s.rtol,
                                                         s.maxiter, s.n_active);
    }
    count_launch(3);
    LMC_CHECK(cudaGetLastError());
    return 0;
}

struct GraphSet {
    cudaGraphExec_t exec[6] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
    unsigned long long launches[6] = {0, 0, 0, 0, 0, 0};   // kernels inside each captured iteration
    ~GraphSet() { for (auto& e : exec) if (e) cudaGraphExecDestroy(e); }
};

static bool graphs_enabled() {
    static const bool off = getenv("LMC_NO_GRAPH") != nullptr;
    return !off;
}

// Optional by-product of a solve: the Lanczos coefficients MINRES computes anyway.  tridiag is [P][k][2] on the
// host (alfa_j, beta_{j+1} of iteration j; zero past a column's last iteration), beta1 is [P] (||b||, or
// sqrt(b . M b) with a preconditioner).
struct LanczosRecord {
    int k = 0;
    double* tridiag = nullptr;
    double* beta1 = nullptr;
};

static int minres_core(MinresOperator& A, Preconditioner* M, const double* RHS, long ld, int P, double* X,
                       double tol, int maxiter, int check_every, int* iters, double* resid, int* istop,
                       cudaStream_t caller_st, const LanczosRecord* record = nullptr) {
    LMC_REQUIRE(P >= 1, "need at least one right-hand side");
    LMC_REQUIRE(maxiter >= 1 && check_every >= 1, "maxiter/check_every must be positive");
    const long n = A.n;
    LMC_REQUIRE(ld >= n, "leading dimension < n");
    const int nblk = ceil_div(n, kVecChunk);
    const size_t vec = sizeof(double) * (size_t)P * n;
    const size_t vec_al = (vec + 255) & ~(size_t)255;
    const int nvec = M ? 10 : 8;
    // one grow-only allocation: state vectors, then the per-CTA partials and the per-column scalars
    const size_t part_al = (sizeof(double) * (size_t)P * nblk + 255) & ~(size_t)255;
    const size_t cs_al = (sizeof(ColState) * (size_t)P + 255) & ~(size_t)255;
    const size_t col_al = (sizeof(double) * (size_t)P + 255) & ~(size_t)255;
    const int rec_k = (record && record->tridiag) ? std::min(record->k, maxiter) : 0;
    const size_t rec_al = (sizeof(double) * (size_t)P * rec_k * 2 + 255) & ~(size_t)255;
    void* ws = nullptr;
    LMC_TRY(A.workspace(vec_al * nvec + 3 * part_al + cs_al + 2 * col_al + 256 + rec_al, &ws));
    char* base = static_cast<char*>(ws);
    WsSlice bufs[10];
    for (int i = 0; i < nvec; ++i) bufs[i].p = base + vec_al * i;
    char* tail = base + vec_al * nvec;
    MinresState s = {};
    s.n = n; s.P = P; s.nblk = nblk;
    s.b = bufs[0].as<double>(); s.x = bufs[1].as<double>();
    s.r1 = bufs[2].as<double>(); s.r2 = bufs[3].as<double>(); s.y = bufs[4].as<double>();
    s.wa = bufs[5].as<double>(); s.wb = bufs[6].as<double>(); s.wc = bufs[7].as<double>();
    s.z = M ? bufs[8].as<double>() : nullptr; s.zold = M ? bufs[9].as<double>() : nullptr;
    s.pa = reinterpret_cast<double*>(tail); s.pb = reinterpret_cast<double*>(tail + part_al);
    s.pc = reinterpret_cast<double*>(tail + 2 * part_al);
    s.cs = reinterpret_cast<ColState*>(tail + 3 * part_al);
    s.inv_beta = reinterpret_cast<double*>(tail + 3 * part_al + cs_al);
    s.active = reinterpret_cast<int*>(tail + 3 * part_al + cs_al + col_al);
    s.n_active = reinterpret_cast<int*>(tail + 3 * part_al + cs_al + 2 * col_al);
    s.rtol = std::fmin(1e-10, tol); s.maxiter = maxiter;
    s.rec = rec_k ? reinterpret_cast<double*>(tail + 3 * part_al + cs_al + 2 * col_al + 256) : nullptr;
    s.rec_k = rec_k;
    const int* perm = A.perm;
    const dim3 vgrid((unsigned)nblk, (unsigned)P);
    const int sthreads = 128, sblocks = ceil_div(P, sthreads);

    // Graph capture needs a capturing-capable stream and products made of this library's launches only;
    // the per-family event timing (lmc_profile_*) wants the plain launches.
    const bool use_graph = graphs_enabled() && A.capturable() && (!M || M->capturable()) && !g_prof_on;
    SolverHost* sh = A.host();
    LMC_TRY(solver_host_init(sh, use_graph));
    cudaStream_t st = caller_st;
    if (use_graph) {
        st = sh->stream;
        LMC_CHECK(cudaEventRecord(sh->ev_join, caller_st));
        LMC_CHECK(cudaStreamWaitEvent(st, sh->ev_join, 0));
    }

    if (rec_k) LMC_CHECK(cudaMemsetAsync(s.rec, 0, sizeof(double) * (size_t)P * rec_k * 2, st));
    minres_init_kernel<<<vgrid, kVecThreads, 0, st>>>(RHS, ld, perm, n, s.b, s.r2, s.x, s.wa, s.wb, s.wc, s.pa, nblk);
    count_launch();
    if (M) {
        LMC_TRY(M->apply(s.r2, nullptr, s.z, n, P, st));                 // y = M r1, beta1^2 = r1 . y
        minres_dot_kernel<<<vgrid, kVecThreads, 0, st>>>(s.r2, s.z, nullptr, n, s.pb, nblk);
        count_launch();
    }
    minres_init_scalars_kernel<<<sblocks, sthreads, 0, st>>>(s.cs, s.inv_beta, s.active, s.pa, M ? s.pb : s.pa, nblk, P,
                                                           s.n_active);
    count_launch();
    LMC_CHECK(cudaGetLastError());

    const int period = M ? 6 : 3;
    GraphSet graphs;
    int head = 0, tail_i = 0;          // ring of outstanding stop-flag reads: [tail_i, head)
    bool stop = false;
    auto drain = [&](bool block) -> int {
        while (tail_i < head) {
            const int slot = tail_i % SolverHost::kRing;
            if (block || head - tail_i >= SolverHost::kRing) {
                LMC_CHECK(cudaEventSynchronize(sh->ev[slot]));
            } else {
                const cudaError_t q = cudaEventQuery(sh->ev[slot]);
                if (q == cudaErrorNotReady) break;
                LMC_CHECK(q);
            }
            if (sh->flags[slot] == 0) stop = true;
            ++tail_i;
        }
        return 0;
    };
    for (int itn = 1; itn <= maxiter && !stop; ++itn) {
        const int phase = (itn - 1) % period;
        if (use_graph && itn >= 2) {
            if (!graphs.exec[phase]) {
                // capture this iteration (the buffer rotation inside minres_iteration is host state and
                // advances exactly as in the eager path)
                cudaGraph_t g = nullptr;
                LMC_CHECK(cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal));
                const unsigned long long l0 = g_launches.load(std::memory_order_relaxed);
                const int rc = minres_iteration(A, M, s, st);
                graphs.launches[phase] = g_launches.load(std::memory_order_relaxed) - l0;
                const cudaError_t e = cudaStreamEndCapture(st, &g);
                if (rc != 0) { if (g) cudaGraphDestroy(g); return rc; }
                LMC_CHECK(e);
                const cudaError_t ei = cudaGraphInstantiate(&graphs.exec[phase], g, 0);
                cudaGraphDestroy(g);
                LMC_CHECK(ei);
            } else {
                // same launches as the captured ones: only advance the host-side rotation
                { double* t = s.r1; s.r1 = s.r2; s.r2 = s.y; s.y = t; }
                if (M) { double* t = s.zold; s.zold = s.z; s.z = t; }
                { double* t = s.wa; s.wa = s.wb; s.wb = s.wc; s.wc = t; }
                count_launch((int)graphs.launches[phase]);   // the replayed graph runs the same kernels
            }
            LMC_CHECK(cudaGraphLaunch(graphs.exec[phase], st));
        } else {
            LMC_TRY(minres_iteration(A, M, s, st));
        }
        if (itn % check_every == 0) {
            // reference callback: true residual of the columns still running
            LMC_TRY(A.apply(s.x, nullptr, s.active, s.y, P, st));
            minres_resid_kernel<<<vgrid, kVecThreads, 0, st>>>(s.b, s.y, s.active, n, s.pa, nblk);
            LMC_CHECK(cudaMemsetAsync(s.n_active, 0, sizeof(int), st));
            minres_s3_kernel<<<P, kScalarThreads, 0, st>>>(s.cs, s.active, s.pa, nblk, P, tol, 0, s.n_active);
            count_launch(2);
        }
        // asynchronous stop test: the count of running columns lands in pinned memory; the host reads
        // it when the copy's event has fired, at most kRing iterations later.  Iterations enqueued
        // after every column has stopped find no active column and do nothing.
        {
            const int slot = head % SolverHost::kRing;
            LMC_CHECK(cudaMemcpyAsync(&sh->flags[slot], s.n_active, sizeof(int), cudaMemcpyDeviceToHost, st));
            LMC_CHECK(cudaEventRecord(sh->ev[slot], st));
            ++head;
        }
        LMC_TRY(drain(itn == maxiter));
    }
    // final residual of every column (iterative.py:53)
    LMC_TRY(A.apply(s.x, nullptr, nullptr, s.y, P, st));
    minres_resid_kernel<<<vgrid, kVecThreads, 0, st>>>(s.b, s.y, nullptr, n, s.pa, nblk);
    minres_s3_kernel<<<P, kScalarThreads, 0, st>>>(s.cs, s.active, s.pa, nblk, P, tol, 1, s.n_active);
    minres_finish_kernel<<<vgrid, kVecThreads, 0, st>>>(s.x, perm, n, X, ld);
    count_launch(3);
    LMC_CHECK(cudaGetLastError());
    std::vector<ColState> h((size_t)P);
    LMC_CHECK(cudaMemcpyAsync(h.data(), s.cs, sizeof(ColState) * P, cudaMemcpyDeviceToHost, st));
    if (rec_k) {
        // rows of record->k entries on the host, rec_k <= k of them filled
        LMC_CHECK(cudaMemcpy2DAsync(record->tridiag, sizeof(double) * 2 * record->k, s.rec, sizeof(double) * 2 * rec_k,
                                    sizeof(double) * 2 * rec_k, P, cudaMemcpyDeviceToHost, st));
    }
    LMC_CHECK(cudaStreamSynchronize(st));
    for (int c = 0; c < P; ++c) {
        if (record && record->beta1) record->beta1[c] = h[c].beta1;
        if (iters) iters[c] = h[c].itn;
        if (resid) resid[c] = h[c].resid;
        if (istop) istop[c] = h[c].istop;
    }
    return 0;
}

int minres_solve_lanczos(lmc_op* op, const double* RHS, long ld, int P, double* X, double tol, int maxiter,
                         int check_every, int* iters, double* resid, int* istop, int k, double* tridiag,
                         double* beta1, cudaStream_t st) {
    LMC_REQUIRE(op->Q > 0, "operator parameters not set");
    LMC_REQUIRE(k >= 1 && tridiag, "need room for at least one Lanczos step");
    FusedOperator A(op);
    LanczosRecord rec;
    rec.k = k; rec.tridiag = tridiag; rec.beta1 = beta1;
    for (long i = 0; i < (long)P * k * 2; ++i) tridiag[i] = 0.0;
    return minres_core(A, nullptr, RHS, ld, P, X, tol, maxiter, check_every, iters, resid, istop, st, &rec);
}

int minres_solve(lmc_op* op, const double* RHS, long ld, int P, double* X, double tol, int maxiter,
                 int check_every, int* iters, double* resid, int* istop, cudaStream_t st, const double* jacobi) {
    LMC_REQUIRE(op->Q > 0, "operator parameters not set");
    FusedOperator A(op);
    if (jacobi) {
        DiagPreconditioner M(jacobi);
        return minres_core(A, &M, RHS, ld, P, X, tol, maxiter, check_every, iters, resid, istop, st);
    }
    return minres_core(A, nullptr, RHS, ld, P, X, tol, maxiter, check_every, iters, resid, istop, st);
}


// ---------------------------------------------------------------------------
// Batched conjugate gradients: Iterative.solve(..., minres=False) (iterative.py:44-51 ->
// scipy.sparse.linalg.cg, scipy 1.18.1 _isolve/iterative.py) with M = I, x0 = 0, atol = 0,
// rtol = min(1e-10, tol), maxiter (reference: n) and the wrapper's true-residual test every
// `check_every` iterations.  Same layout and conventions as the MINRES block solver above.
// ---------------------------------------------------------------------------
struct CgState {
    double atol, rho, rho_prev, beta, resid;
    int itn, done, info;   // info: scipy's second return value (0 converged, maxiter if exhausted); 10 = residual test
};

// r = b_sorted, x = 0 ; partial ||b||^2
__global__ void __launch_bounds__(kVecThreads) cg_init_kernel(const double* __restrict__ RHS, long ld,
                                                              const int* __restrict__ perm, long n, double* b,
                                                              double* r, double* x, double* part, int nblk) {
    const int col = blockIdx.y;
    const long base = (long)blockIdx.x * kVecChunk;
    double acc = 0.0;
    for (int k = 0; k < kVecPerThread; ++k) {
        const long i = base + k * kVecThreads + threadIdx.x;
        if (i < n) {
            const double v = RHS[(long)col * ld + (perm ? perm[i] : i)];
            const long o = (long)col * n + i;
            b[o] = v; r[o] = v; x[o] = 0.0;
            acc = fma(v, v, acc);
        }
    }
    acc = block_reduce_sum(acc);
    if (threadIdx.x == 0) part[(long)col * nblk + blockIdx.x] = acc;
}

__global__ void cg_init_scalars_kernel(CgState* st, int* active, const double* part, int nblk, int P, double rtol,
                                       int* n_active) {
    const int col = blockIdx.x;
    const double s = column_sum(part + (long)col * nblk, nblk);
    if (threadIdx.x != 0) return;
    CgState c = {};
    const double bnrm2 = sqrt(s);
    c.atol = rtol * bnrm2;               // _get_atol_rtol with atol = 0
    c.rho = s;                           // r = b, z = r
    c.beta = 0.0;
    c.done = (bnrm2 == 0.0 || bnrm2 < c.atol) ? 1 : 0;   // scipy returns b (= 0) at once
    st[col] = c;
    active[col] = c.done ? 0 : 1;
    if (col == 0) *n_active = -1;
}

// p = beta p + r   (first iteration: beta = 0, p is not read)
__global__ void __launch_bounds__(kVecThreads) cg_c1_kernel(double* p, const double* __restrict__ r, const CgState* st,
                                                            const int* active, long n) {
    const int col = blockIdx.y;
    if (!active[col]) return;
    const double beta = st[col].beta;
    const bool first = st[col].itn == 0;
    const long base = (long)blockIdx.x * kVecChunk;
    for (int k = 0; k < kVecPerThread; ++k) {
        const long i = base + k * kVecThreads + threadIdx.x;
        if (i < n) {
            const long o = (long)col * n + i;
            p[o] = first ? r[o] : __dadd_rn(__dmul_rn(p[o], beta), r[o]);
        }
    }
}

// partial p . q
__global__ void __launch_bounds__(kVecThreads) cg_c2_kernel(const double* __restrict__ p, const double* __restrict__ q,
                                                            const int* active, long n, double* part, int nblk) {
    const int col = blockIdx.y;
    if (!active[col]) return;
    const long base = (long)blockIdx.x * kVecChunk;
    double acc = 0.0;
    for (int k = 0; k < kVecPerThread; ++k) {
        const long i = base + k * kVecThreads + threadIdx.x;
        if (i < n) acc = fma(p[(long)col * n + i], q[(long)col * n + i], acc);
    }
    acc = block_reduce_sum(acc);
    if (threadIdx.x == 0) part[(long)col * nblk + blockIdx.x] = acc;
}

// alpha = rho / (p . q) ; x += alpha p ; r -= alpha q ; partial r . r
__global__ void __launch_bounds__(kVecThreads) cg_c3_kernel(double* x, double* r, const double* __restrict__ p,
                                                            const double* __restrict__ q, const CgState* st,
                                                            const int* active, long n, const double* part_pq,
                                                            double* part_rr, int nblk) {
    const int col = blockIdx.y;
    if (!active[col]) return;
    const double pq = block_sum_partials(part_pq + (long)col * nblk, nblk);
    const double alpha = st[col].rho / pq;
    const long base = (long)blockIdx.x * kVecChunk;
    double acc = 0.0;
    for (int k = 0; k < kVecPerThread; ++k) {
        const long i = base + k * kVecThreads + threadIdx.x;
        if (i < n) {
            const long o = (long)col * n + i;
            x[o] = __dadd_rn(x[o], __dmul_rn(alpha, p[o]));
            const double rv = __dsub_rn(r[o], __dmul_rn(alpha, q[o]));
            r[o] = rv;
            acc = fma(rv, rv, acc);
        }
    }
    acc = block_reduce_sum(acc);
    if (threadIdx.x == 0) part_rr[(long)col * nblk + blockIdx.x] = acc;
}

// end of an iteration: rho bookkeeping, scipy's stopping test (top of its next iteration), beta
__global__ void cg_scalars_kernel(CgState* st, int* active, const double* part_rr, int nblk, int P, int maxiter,
                                  int* n_active) {
    const int col = blockIdx.x;
    if (col == 0 && threadIdx.x == 0) *n_active = 0;
    if (!active[col]) return;
    const double s = column_sum(part_rr + (long)col * nblk, nblk);
    if (threadIdx.x != 0) return;
    CgState c = st[col];
    c.itn += 1;
    c.rho_prev = c.rho;
    c.rho = s;
    c.beta = c.rho / c.rho_prev;
    if (c.itn >= maxiter) { c.done = 1; c.info = maxiter; }
    else if (sqrt(s) < c.atol) { c.done = 1; c.info = 0; }
    if (c.done) active[col] = 0;
    st[col] = c;
}

__global__ void cg_count_active_kernel(const int* active, int P, int* n_active) {
    int cnt = 0;
    for (int c = threadIdx.x; c < P; c += blockDim.x) cnt += active[c] != 0;
    if (cnt) atomicAdd(n_active, cnt);
}

__global__ void cg_resid_scalars_kernel(CgState* st, int* active, const double* part, int nblk, int P, double tol,
                                        int final_pass, int* n_active) {
    const int col = blockIdx.x;
    if (!final_pass && !active[col]) return;
    const double s = column_sum(part + (long)col * nblk, nblk);
    if (threadIdx.x != 0) return;
    const double rn = sqrt(s);
    st[col].resid = rn;
    if (!final_pass) {
        if (rn < tol) { st[col].info = 10; st[col].done = 1; active[col] = 0; }
        else atomicAdd(n_active, 1);
    }
}

static int cg_core(MinresOperator& A, const double* RHS, long ld, int P, double* X, double tol, int maxiter,
                   int check_every, int* iters, double* resid, int* info, cudaStream_t st) {
    LMC_REQUIRE(P >= 1, "need at least one right-hand side");
    LMC_REQUIRE(maxiter >= 1 && check_every >= 1, "maxiter/check_every must be positive");
    const long n = A.n;
    LMC_REQUIRE(ld >= n, "leading dimension < n");
    const int nblk = ceil_div(n, kVecChunk);
    const double rtol = std::fmin(1e-10, tol);
    const size_t vec = sizeof(double) * (size_t)P * n;
    const size_t vec_al = (vec + 255) & ~(size_t)255;
    void* ws = nullptr;
    LMC_TRY(A.workspace(vec_al * 5, &ws));
    WsSlice bufs[5];
    for (int i = 0; i < 5; ++i) bufs[i].p = static_cast<char*>(ws) + vec_al * i;
    double *b = bufs[0].as<double>(), *x = bufs[1].as<double>(), *r = bufs[2].as<double>();
    double *p = bufs[3].as<double>(), *q = bufs[4].as<double>();
    DevBuf parts[2], stb, act, nact;
    for (auto& pb : parts) LMC_TRY(pb.alloc(sizeof(double) * (size_t)P * nblk));
    LMC_TRY(stb.alloc(sizeof(CgState) * P));
    LMC_TRY(act.alloc(sizeof(int) * P));
    LMC_TRY(nact.alloc(sizeof(int)));
    double *pa = parts[0].as<double>(), *pr = parts[1].as<double>();
    CgState* cs = stb.as<CgState>();
    int* active = act.as<int>();
    int* n_active = nact.as<int>();
    const int* perm = A.perm;
    const dim3 vgrid((unsigned)nblk, (unsigned)P);

    cg_init_kernel<<<vgrid, kVecThreads, 0, st>>>(RHS, ld, perm, n, b, r, x, pa, nblk);
    cg_init_scalars_kernel<<<P, kScalarThreads, 0, st>>>(cs, active, pa, nblk, P, rtol, n_active);
    count_launch(2);
    LMC_CHECK(cudaGetLastError());

    int h_active = P;
    const int poll = 8;
    for (int itn = 1; itn <= maxiter; ++itn) {
        {
            ProfScope prof(PROF_MINRES_VEC, st);
            cg_c1_kernel<<<vgrid, kVecThreads, 0, st>>>(p, r, cs, active, n);
        }
        LMC_TRY(A.apply(p, nullptr, active, q, P, st));
        {
            ProfScope prof(PROF_MINRES_VEC, st);
            cg_c2_kernel<<<vgrid, kVecThreads, 0, st>>>(p, q, active, n, pa, nblk);
            cg_c3_kernel<<<vgrid, kVecThreads, 0, st>>>(x, r, p, q, cs, active, n, pa, pr, nblk);
        }
        {
            ProfScope prof(PROF_MINRES_SCALAR, st);
            cg_scalars_kernel<<<P, kScalarThreads, 0, st>>>(cs, active, pr, nblk, P, maxiter, n_active);
            cg_count_active_kernel<<<1, 128, 0, st>>>(active, P, n_active);
        }
        count_launch(5);
        bool polled = false;
        if (itn % check_every == 0) {
            // reference callback: true residual of the columns still running (iterative.py:36-42)
            LMC_TRY(A.apply(x, nullptr, active, q, P, st));
            minres_resid_kernel<<<vgrid, kVecThreads, 0, st>>>(b, q, active, n, pa, nblk);
            LMC_CHECK(cudaMemsetAsync(n_active, 0, sizeof(int), st));
            cg_resid_scalars_kernel<<<P, kScalarThreads, 0, st>>>(cs, active, pa, nblk, P, tol, 0, n_active);
            count_launch(2);
            polled = true;
        }
        if (polled || itn % poll == 0 || itn == maxiter) {
            LMC_CHECK(cudaMemcpyAsync(&h_active, n_active, sizeof(int), cudaMemcpyDeviceToHost, st));
            LMC_CHECK(cudaStreamSynchronize(st));
            if (h_active == 0) break;
        }
    }
    // final residual of every column (iterative.py:53)
    LMC_TRY(A.apply(x, nullptr, nullptr, q, P, st));
    minres_resid_kernel<<<vgrid, kVecThreads, 0, st>>>(b, q, nullptr, n, pa, nblk);
    cg_resid_scalars_kernel<<<P, kScalarThreads, 0, st>>>(cs, active, pa, nblk, P, tol, 1, n_active);
    minres_finish_kernel<<<vgrid, kVecThreads, 0, st>>>(x, perm, n, X, ld);
    count_launch(3);
    LMC_CHECK(cudaGetLastError());
    std::vector<CgState> h((size_t)P);
    LMC_CHECK(cudaMemcpyAsync(h.data(), cs, sizeof(CgState) * P, cudaMemcpyDeviceToHost, st));
    LMC_CHECK(cudaStreamSynchronize(st));
    for (int c = 0; c < P; ++c) {
        if (iters) iters[c] = h[c].itn;
        if (resid) resid[c] = h[c].resid;
        if (info) info[c] = h[c].info;
    }
    return 0;
}

int cg_solve(lmc_op* op, const double* RHS, long ld, int P, double* X, double tol, int maxiter,
             int check_every, int* iters, double* resid, int* info, cudaStream_t st) {
    LMC_REQUIRE(op->Q > 0, "operator parameters not set");
    FusedOperator A(op);
    return cg_core(A, RHS, ld, P, X, tol, maxiter, check_every, iters, resid, info, st);
}

// sum over a [ncols][n] block of A .* B, two-stage deterministic
__global__ void __launch_bounds__(kVecThreads) block_dot_kernel(const double* __restrict__ A,
                                                               const double* __restrict__ B, long lda,
                                                               long ldb, long n, double* part, int nblk) {
    const int col = blockIdx.y;
    const long base = (long)blockIdx.x * kVecChunk;
    double acc = 0.0;
    for (int k = 0; k < kVecPerThread; ++k) {
        const long i = base + k * kVecThreads + threadIdx.x;
        if (i < n) acc = fma(A[(long)col * lda + i], B[(long)col * ldb + i], acc);
    }
    acc = block_reduce_sum(acc);
    if (threadIdx.x == 0) part[(long)col * nblk + blockIdx.x] = acc;
}

}  // namespace lmc

using namespace lmc;

extern "C" {

int lmc_minres_generic(int (*apply_cb)(void*), void* ctx, long n, double* scratch_in_dev,
                       double* scratch_out_dev, const double* RHS_dev, long ld, int P, double* X_dev,
                       double tol, int maxiter, int check_every, int* iters_host, double* resid_host,
                       int* istop_host, void* stream) {
    LMC_REQUIRE(apply_cb && scratch_in_dev && scratch_out_dev && RHS_dev && X_dev, "null argument");
    LMC_REQUIRE(n >= 1 && ld >= n, "bad block shape");
    CallbackOperator A;
    A.n = n; A.perm = nullptr; A.cb = apply_cb; A.ctx = ctx;
    A.scratch_in = scratch_in_dev; A.scratch_out = scratch_out_dev;
    return minres_core(A, nullptr, RHS_dev, ld, P, X_dev, tol, maxiter, check_every, iters_host, resid_host,
                       istop_host, (cudaStream_t)stream);
}

int lmc_minres_generic_pre(int (*apply_cb)(void*), int (*precond_cb)(void*), void* ctx, long n,
                           double* scratch_in_dev, double* scratch_out_dev, const double* RHS_dev, long ld,
                           int P, double* X_dev, double tol, int maxiter, int check_every, int* iters_host,
                           double* resid_host, int* istop_host, void* stream) {
    LMC_REQUIRE(apply_cb && precond_cb && scratch_in_dev && scratch_out_dev && RHS_dev && X_dev, "null argument");
    LMC_REQUIRE(n >= 1 && ld >= n, "bad block shape");
    CallbackOperator A;
    A.n = n; A.perm = nullptr; A.cb = apply_cb; A.ctx = ctx;
    A.scratch_in = scratch_in_dev; A.scratch_out = scratch_out_dev;
    CallbackPreconditioner M;
    M.cb = precond_cb; M.ctx = ctx; M.scratch_in = scratch_in_dev; M.scratch_out = scratch_out_dev;
    return minres_core(A, &M, RHS_dev, ld, P, X_dev, tol, maxiter, check_every, iters_host, resid_host,
                       istop_host, (cudaStream_t)stream);
}

int lmc_cg_generic(int (*apply_cb)(void*), void* ctx, long n, double* scratch_in_dev, double* scratch_out_dev,
                   const double* RHS_dev, long ld, int P, double* X_dev, double tol, int maxiter,
                   int check_every, int* iters_host, double* resid_host, int* info_host, void* stream) {
    LMC_REQUIRE(apply_cb && scratch_in_dev && scratch_out_dev && RHS_dev && X_dev, "null argument");
    LMC_REQUIRE(n >= 1 && ld >= n, "bad block shape");
    CallbackOperator A;
    A.n = n; A.perm = nullptr; A.cb = apply_cb; A.ctx = ctx;
    A.scratch_in = scratch_in_dev; A.scratch_out = scratch_out_dev;
    return cg_core(A, RHS_dev, ld, P, X_dev, tol, maxiter, check_every, iters_host, resid_host, info_host,
                   (cudaStream_t)stream);
}

int lmc_block_dot(const double* A_dev, long lda, const double* B_dev, long ldb, long n, int ncols,
                  double* out_host, void* stream) {
    LMC_REQUIRE(A_dev && B_dev && out_host && n >= 1 && ncols >= 0, "bad argument");
    *out_host = 0.0;
    if (ncols == 0) return 0;
    cudaStream_t st = (cudaStream_t)stream;
    const int nblk = ceil_div(n, kVecChunk);
    DevBuf part;
    LMC_TRY(part.alloc(sizeof(double) * (size_t)ncols * nblk));
    block_dot_kernel<<<dim3((unsigned)nblk, (unsigned)ncols), kVecThreads, 0, st>>>(
        A_dev, B_dev, lda, ldb, n, part.as<double>(), nblk);
    count_launch();
    LMC_CHECK(cudaGetLastError());
    std::vector<double> h((size_t)ncols * nblk);
    LMC_CHECK(cudaMemcpyAsync(h.data(), part.p, sizeof(double) * h.size(), cudaMemcpyDeviceToHost, st));
    LMC_CHECK(cudaStreamSynchronize(st));
    double s = 0.0;
    for (double v : h) s += v;
    *out_host = s;
    return 0;
}

}  // extern "C"
