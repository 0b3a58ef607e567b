// Per-model and per-step operator setup on the device.
//
//  * kernel tops: the values k_q(|z - z_0|) of the stationary kernels on the inducing grid and
//    their parameter derivatives, evaluated where they are consumed instead of on the host and
//    uploaded (reference: kern/rbf.py:39-54, kern/matern32.py:39-57, kern/std_periodic.py:44-67
//    applied to dists = la.norm(grid - grid[0], axis=-1), models/interpolated_llgp.py:431;
//    consumed by grid_kernel.py:26-27 and likelihood.py:121-124);
//  * point set: base index / fractional offset of every point, the sort by (output, grid bin),
//    the bin CSR and the inverse permutation (reference: the CSR construction of
//    multi_interpolant, approx/interpolation.py:119-176 -- here there is no CSR, only the sort).
//    The arithmetic is the host builder's (interp.cu build_points), IEEE operation by operation,
//    and the sort is stable, so both builders produce the same operator bit for bit.
#include "../../include/lmc_b200.h"
#include "op.cuh"

#include <cub/device/device_radix_sort.cuh>

#include <cmath>
#include <vector>

namespace lmc {

// ---------------------------------------------------------------------------
// kernel values on the grid
// ---------------------------------------------------------------------------
struct TopSpec {
    int kind;    // LMC_KERN_*
    int which;   // 0: k(r); 1: dk/d inv_lengthscale; 2: dk/d period
    double g;    // inv_lengthscale
    double T;    // period (std_periodic only)
};

// operation order of the reference's numpy expressions; exp/sin/cos are the CUDA double-precision
// functions (<= 2 ulp), numpy's are libm's: tops agree to ~1e-15 relative
__device__ double eval_top(const TopSpec s, double r) {
    const double kSqrt3 = 1.7320508075688772;
    const double kPi = 3.141592653589793;
    switch (s.kind) {
        case LMC_KERN_RBF: {
            const double sq = r * r;
            const double e = exp(-0.5 * sq * s.g);
            return s.which == 0 ? e : e * -0.5 * sq;
        }
        case LMC_KERN_MATERN32: {
            const double ds = r * kSqrt3;
            const double sd = ds * s.g;
            const double e = exp(-sd);
            if (s.which == 0) return (1.0 + sd) * e;
            const double dexp = e * -ds;
            return (1.0 + sd) * dexp + ds * e;
        }
        default: {  // LMC_KERN_STD_PERIODIC
            const double sc = (kPi / s.T) * r;
            const double sn = sin(sc);
            const double sq = sn * sn;
            const double e = exp(-0.5 * sq * s.g);
            if (s.which == 0) return e;
            if (s.which == 1) return e * -0.5 * sq;
            double dsn = cos(sc) * sc;
            dsn *= -1.0 / s.T * s.g;
            return e * -1.0 * sn * dsn;
        }
    }
}

static const int kMaxTops = 192;   // Q <= 64 kernels, <= 2 parameters each, plus the values
struct TopSpecs {
    TopSpec s[16];
};

// tops[t][cell] for up to 16 descriptors per launch; cell = ix * m1 + iy (interpolation.py:309)
__global__ void __launch_bounds__(256) eval_tops_kernel(const TopSpecs specs, int ntops, int ndim, int m1,
                                                        double d0, double d1, long cells, double* tops) {
    const long c = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= cells) return;
    double r;
    if (ndim == 1) {
        r = __dmul_rn((double)c, d0);
    } else {
        const double dx = __dmul_rn((double)(c / m1), d0), dy = __dmul_rn((double)(c % m1), d1);
        r = sqrt(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)));
    }
    for (int t = 0; t < ntops; ++t) tops[(size_t)t * cells + c] = eval_top(specs.s[t], r);
}

static int n_kernel_params(int kind) { return kind == LMC_KERN_STD_PERIODIC ? 2 : 1; }

static int eval_tops(const lmc_op* op, const std::vector<TopSpec>& specs, double* tops_dev, cudaStream_t st) {
    const long cells = op->emb.cells;
    for (size_t t0 = 0; t0 < specs.size(); t0 += 16) {
        TopSpecs pack;
        const int cnt = (int)std::min<size_t>(16, specs.size() - t0);
        for (int i = 0; i < cnt; ++i) pack.s[i] = specs[t0 + i];
        eval_tops_kernel<<<ceil_div(cells, 256), 256, 0, st>>>(pack, cnt, op->ndim, op->emb.m[1], op->delta[0],
                                                              op->delta[1], cells, tops_dev + t0 * cells);
        count_launch();
    }
    LMC_CHECK(cudaGetLastError());
    return 0;
}

static void kernel_specs(const lmc_op* op, bool deriv, std::vector<TopSpec>* out) {
    out->clear();
    for (size_t q = 0; q < op->kinds.size(); ++q) {
        const int np = deriv ? n_kernel_params(op->kinds[q]) : 1;
        for (int w = 0; w < np; ++w)
            out->push_back({op->kinds[q], deriv ? w + 1 : 0, op->kparams[2 * q], op->kparams[2 * q + 1]});
    }
}

struct SetupBuf {
    void* p = nullptr;
    ~SetupBuf() { if (p) cudaFree(p); }
    int alloc(size_t bytes) { LMC_CHECK(cudaMalloc(&p, bytes ? bytes : 8)); return 0; }
    template <class T> T* as() { return static_cast<T*>(p); }
};

// ---------------------------------------------------------------------------
// point set
// ---------------------------------------------------------------------------
struct PointGeom {
    long out_start[17];
    double origin[2], delta[2];
    int m[2], nb[2];
    long NB;
    int D, ndim;
};

// caller order: key = output * NB + bin, base index and fractional offset per axis
__global__ void __launch_bounds__(256) point_keys_kernel(const double* __restrict__ X, long n, const PointGeom g,
                                                         unsigned* keys, int* idx, int* i0a, int* i0b,
                                                         double* ua, double* ub, int* bad) {
    const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    int d = 0;
    while (d + 1 < g.D && i >= g.out_start[d + 1]) ++d;
    long bin = 0;
    for (int p = 0; p < g.ndim; ++p) {
        const double s = X[i * g.ndim + p];
        if (!isfinite(s)) *bad = 1;
        const double f = __ddiv_rn(__dsub_rn(s, g.origin[p]), g.delta[p]);
        const double fl = floor(f);
        const double hi = (double)g.m[p];
        const double cl = fl < -2.0 ? -2.0 : (fl > hi ? hi : fl);   // a NaN falls through to itself; flagged above
        const int ci = isfinite(s) ? (int)cl : 0;
        (p == 0 ? i0a : i0b)[i] = ci;
        (p == 0 ? ua : ub)[i] = __dsub_rn(f, fl);
        bin = bin * g.nb[p] + (ci + 2);
    }
    keys[i] = (unsigned)((long)d * g.NB + bin);
    idx[i] = (int)i;
}

// sorted order: gather the per-point data, invert the permutation, note whether it is the identity
__global__ void __launch_bounds__(256) point_gather_kernel(const int* __restrict__ perm, long n, int ndim,
                                                           const int* i0a, const int* i0b, const double* ua,
                                                           const double* ub, int* s_i0a, int* s_i0b, double* s_ua,
                                                           double* s_ub, int* iperm, int* moved) {
    const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int src = perm[i];
    s_i0a[i] = i0a[src];
    s_ua[i] = ua[src];
    if (ndim == 2) {
        s_i0b[i] = i0b[src];
        s_ub[i] = ub[src];
    }
    iperm[src] = (int)i;
    if (src != (int)i) *moved = 1;
}

// bin_start[b] = first sorted position with key >= b, b in [0, nbins]
__global__ void __launch_bounds__(256) bin_start_kernel(const unsigned* __restrict__ skeys, long n, long nbins,
                                                        int* bin_start) {
    const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i > n) return;
    const long lo = (i == 0) ? 0 : (long)skeys[i - 1] + 1;
    const long hi = (i == n) ? nbins : (long)skeys[i];
    for (long b = lo; b <= hi; ++b) bin_start[b] = (int)i;
}

int build_points_dev(PointSet* ps, int D, int ndim, const int* grid_sizes, const double* origin,
                     const double* delta, const int* lens, const double* X_dev, long grid_pitch, cudaStream_t st) {
    LMC_REQUIRE(D >= 1 && D <= 16, "number of outputs D must be in 1..16");
    LMC_REQUIRE(ndim == 1 || ndim == 2, "interpolation supports 1-D and 2-D inputs");
    *ps = PointSet();
    ps->D = D;
    ps->ndim = ndim;
    ps->grid_pitch = grid_pitch;
    PointGeom g = {};
    g.D = D;
    g.ndim = ndim;
    long n = 0;
    for (int d = 0; d < D; ++d) {
        LMC_REQUIRE(lens[d] >= 0, "negative output length");
        ps->out_start[d] = g.out_start[d] = n;
        n += lens[d];
    }
    ps->out_start[D] = g.out_start[D] = n;
    LMC_REQUIRE(n >= 1 && n < 2147483647L, "total number of points out of range");
    ps->n = n;
    ps->NB = 1;
    for (int p = 0; p < ndim; ++p) {
        LMC_REQUIRE(grid_sizes[p] >= 4, "grid size must be >= 4");
        LMC_REQUIRE(delta[p] > 0 && std::isfinite(delta[p]), "grid spacing must be positive");
        ps->m[p] = g.m[p] = grid_sizes[p];
        ps->nb[p] = g.nb[p] = grid_sizes[p] + 3;
        ps->NB *= ps->nb[p];
        g.origin[p] = origin[p];
        g.delta[p] = delta[p];
    }
    g.NB = ps->NB;
    const long nbins = (long)D * ps->NB;
    LMC_REQUIRE(nbins < 2147483647L, "too many grid bins for 32-bit sort keys");

    SetupBuf keys, skeys, idx, ci0[2], cu[2], flags, tmp;
    LMC_TRY(keys.alloc(sizeof(unsigned) * n));
    LMC_TRY(skeys.alloc(sizeof(unsigned) * n));
    LMC_TRY(idx.alloc(sizeof(int) * n));
    for (int p = 0; p < 2; ++p) {
        LMC_TRY(ci0[p].alloc(sizeof(int) * (p < ndim ? n : 1)));
        LMC_TRY(cu[p].alloc(sizeof(double) * (p < ndim ? n : 1)));
    }
    LMC_TRY(flags.alloc(sizeof(int) * 2));
    LMC_CHECK(cudaMemsetAsync(flags.p, 0, sizeof(int) * 2, st));
    int* bad = flags.as<int>();
    int* moved = bad + 1;
    const unsigned nblk = (unsigned)ceil_div(n, 256);
    point_keys_kernel<<<nblk, 256, 0, st>>>(X_dev, n, g, keys.as<unsigned>(), idx.as<int>(), ci0[0].as<int>(),
                                            ci0[1].as<int>(), cu[0].as<double>(), cu[1].as<double>(), bad);
    count_launch();
    LMC_CHECK(cudaGetLastError());

    // stable LSD radix sort of (key, caller index): points of one bin keep the caller's order, like
    // the host's counting sort, so both builders fix the same summation order inside a grid cell
    LMC_CHECK(cudaMalloc(&ps->perm, sizeof(int) * n));
    int end_bit = 1;
    while (end_bit < 32 && (1L << end_bit) < nbins) ++end_bit;
    size_t tmp_bytes = 0;
    LMC_CHECK(cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, keys.as<unsigned>(), skeys.as<unsigned>(),
                                              idx.as<int>(), ps->perm, (int)n, 0, end_bit, st));
    LMC_TRY(tmp.alloc(tmp_bytes));
    LMC_CHECK(cub::DeviceRadixSort::SortPairs(tmp.p, tmp_bytes, keys.as<unsigned>(), skeys.as<unsigned>(),
                                              idx.as<int>(), ps->perm, (int)n, 0, end_bit, st));
    count_launch();

    LMC_CHECK(cudaMalloc(&ps->iperm, sizeof(int) * n));
    for (int p = 0; p < ndim; ++p) {
        LMC_CHECK(cudaMalloc(&ps->i0[p], sizeof(int) * n));
        LMC_CHECK(cudaMalloc(&ps->u[p], sizeof(double) * n));
    }
    point_gather_kernel<<<nblk, 256, 0, st>>>(ps->perm, n, ndim, ci0[0].as<int>(), ci0[1].as<int>(),
                                              cu[0].as<double>(), cu[1].as<double>(), ps->i0[0], ps->i0[1],
                                              ps->u[0], ps->u[1], ps->iperm, moved);
    LMC_CHECK(cudaMalloc(&ps->bin_start, sizeof(int) * (nbins + 1)));
    bin_start_kernel<<<(unsigned)ceil_div(n + 1, 256), 256, 0, st>>>(skeys.as<unsigned>(), n, nbins, ps->bin_start);
    count_launch(2);
    LMC_CHECK(cudaGetLastError());

    int h_flags[2] = {0, 0};
    std::vector<int> start((size_t)nbins + 1);
    LMC_CHECK(cudaMemcpyAsync(h_flags, flags.p, sizeof(int) * 2, cudaMemcpyDeviceToHost, st));
    LMC_CHECK(cudaMemcpyAsync(start.data(), ps->bin_start, sizeof(int) * (nbins + 1), cudaMemcpyDeviceToHost, st));
    LMC_CHECK(cudaStreamSynchronize(st));
    LMC_REQUIRE(h_flags[0] == 0, "non-finite input coordinate");
    ps->identity = h_flags[1] == 0;
    if (ps->identity) {
        cudaFree(ps->iperm);
        ps->iperm = nullptr;
    }
    // shared-memory capacities of the tiled 2-D kernels: O(bins) work on the bin CSR, not O(n)
    tile_populations(ps, start);
    LMC_CHECK(cudaMalloc(&ps->out_start_dev, sizeof(long) * (D + 1)));
    LMC_CHECK(cudaMemcpy(ps->out_start_dev, ps->out_start, sizeof(long) * (D + 1), cudaMemcpyHostToDevice));
    return 0;
}

}  // namespace lmc

using namespace lmc;

extern "C" {

int lmc_op_create_dev(lmc_op** out, int D, int ndim, const int* grid_sizes, const double* origin,
                      const double* delta, const int* lens, const double* X_dev, void* stream) {
    LMC_REQUIRE(out && grid_sizes && origin && delta && lens && X_dev, "null argument");
    LMC_REQUIRE(ndim == 1 || ndim == 2, "ndim must be 1 or 2");
    lmc_op* op = new lmc_op();
    op->D = D;
    op->ndim = ndim;
    for (int p = 0; p < ndim; ++p) { op->origin[p] = origin[p]; op->delta[p] = delta[p]; }
    int rc = embedding_init(&op->emb, ndim, grid_sizes);
    if (rc == 0) rc = op->eng.init(op->emb);
    if (rc == 0)
        rc = build_points_dev(&op->ps, D, ndim, grid_sizes, origin, delta, lens, X_dev, op->emb.grid_pitch,
                              (cudaStream_t)stream);
    if (rc != 0) { delete op; return rc; }
    *out = op;
    return 0;
}

int lmc_op_set_kernels(lmc_op* op, int Q, const int* kinds_host, const double* kparams_host,
                       const double* B_host, const double* noise_host) {
    LMC_REQUIRE(op && kinds_host && kparams_host && B_host && noise_host, "null argument");
    LMC_REQUIRE(Q >= 1 && Q <= 64, "Q must be in 1..64");
    for (int q = 0; q < Q; ++q) {
        LMC_REQUIRE(kinds_host[q] >= LMC_KERN_RBF && kinds_host[q] <= LMC_KERN_STD_PERIODIC, "unknown kernel kind");
        LMC_REQUIRE(std::isfinite(kparams_host[2 * q]), "non-finite inverse lengthscale");
        if (kinds_host[q] == LMC_KERN_STD_PERIODIC)
            LMC_REQUIRE(std::isfinite(kparams_host[2 * q + 1]) && kparams_host[2 * q + 1] > 0, "period must be positive");
    }
    std::vector<int> old_kinds = op->kinds;
    std::vector<double> old_params = op->kparams;
    op->kinds.assign(kinds_host, kinds_host + Q);
    op->kparams.assign(kparams_host, kparams_host + 2 * (size_t)Q);
    std::vector<TopSpec> specs;
    kernel_specs(op, false, &specs);
    SetupBuf tops;
    int rc = tops.alloc(sizeof(double) * (size_t)Q * op->emb.cells);
    if (rc == 0) rc = eval_tops(op, specs, tops.as<double>(), 0);
    if (rc == 0) rc = op_set_params_dev(op, Q, tops.as<double>(), B_host, noise_host);
    if (rc != 0) { op->kinds = old_kinds; op->kparams = old_params; }
    return rc;
}

int lmc_op_num_kernel_tops(const lmc_op* op, int deriv) {
    if (!op || op->kinds.empty()) return 0;
    int cnt = 0;
    for (int k : op->kinds) cnt += deriv ? n_kernel_params(k) : 1;
    return cnt;
}

int lmc_op_kernel_tops(lmc_op* op, int deriv, double* tops_host) {
    LMC_REQUIRE(op && tops_host, "null argument");
    LMC_REQUIRE(!op->kinds.empty(), "no kernel descriptors (call lmc_op_set_kernels)");
    std::vector<TopSpec> specs;
    kernel_specs(op, deriv != 0, &specs);
    SetupBuf tops;
    const size_t bytes = sizeof(double) * specs.size() * (size_t)op->emb.cells;
    LMC_TRY(tops.alloc(bytes));
    LMC_TRY(eval_tops(op, specs, tops.as<double>(), 0));
    LMC_CHECK(cudaMemcpy(tops_host, tops.p, bytes, cudaMemcpyDeviceToHost));
    return 0;
}

int lmc_grad_grams_kernels(lmc_op* op, const double* alpha_dev, const double* R_dev, const double* RINV_dev,
                           long ld, int N, double* quad_host, double* trace_host, double* nquad_host,
                           double* ntrace_host, void* stream) {
    LMC_REQUIRE(op && alpha_dev && quad_host && trace_host && nquad_host && ntrace_host, "null argument");
    LMC_REQUIRE(N == 0 || (R_dev && RINV_dev), "null probe blocks");
    LMC_REQUIRE(!op->kinds.empty() && (int)op->kinds.size() == op->Q,
                "no kernel descriptors (call lmc_op_set_kernels)");
    cudaStream_t st = (cudaStream_t)stream;
    std::vector<TopSpec> specs;
    kernel_specs(op, true, &specs);
    LMC_REQUIRE((int)specs.size() <= kMaxTops, "too many kernel parameters");
    SetupBuf tops;
    LMC_TRY(tops.alloc(sizeof(double) * specs.size() * (size_t)op->emb.cells));
    LMC_TRY(eval_tops(op, specs, tops.as<double>(), st));
    return grad_grams(op, alpha_dev, R_dev, RINV_dev, ld, N, (int)specs.size(), tops.as<double>(), true,
                      quad_host, trace_host, nquad_host, ntrace_host, st);
}

}  // extern "C"
