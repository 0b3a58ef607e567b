// Sparse cubic-interpolation gather (W g) and scatter (W^T v) without a stored
// CSR: weights are recomputed from each point's fractional grid offset.
//
// Replaces scipy CSR SpMV `W.dot` / `WT.dot` inside SKI (reference
// runlmc/approx/ski.py:14-16) where W comes from interp_cubic / interp_bicubic /
// multi_interpolant (approx/interpolation.py:56-116, 218-328, 119-176).
//
// Points are sorted once by (output, grid bin).  W^T v is then a segmented
// reduction: one thread owns one bin, accumulates the 4^d per-tap partial sums
// of its (contiguous) points in registers, partial sums are exchanged through
// shared memory and each grid cell adds up the taps that land on it in a fixed
// order -- deterministic and free of atomics.  Clamped stencils at the grid
// edge accumulate onto the edge cell exactly like the reference's CSR `+=`
// (interpolation.py:105-115).
#include "interp.cuh"

#include <algorithm>
#include <cmath>
#include <vector>

namespace lmc {

// ---------------------------------------------------------------------------
// host: sort points by bin
// ---------------------------------------------------------------------------
int build_points(PointSet* ps, int D, int ndim, const int* grid_sizes, const double* origin,
                 const double* delta, const int* lens, const double* X, long grid_pitch) {
    LMC_REQUIRE(D >= 1 && D <= 16, "number of outputs D must be in 1..16");
    LMC_REQUIRE(ndim == 1 || ndim == 2, "interpolation supports 1-D and 2-D inputs");
    *ps = PointSet();
    ps->D = D;
    ps->ndim = ndim;
    ps->grid_pitch = grid_pitch;
    long n = 0;
    for (int d = 0; d < D; ++d) {
        LMC_REQUIRE(lens[d] >= 0, "negative output length");
        ps->out_start[d] = n;
        n += lens[d];
    }
    ps->out_start[D] = n;
    LMC_REQUIRE(n >= 1 && n < 2147483647L, "total number of points out of range");
    ps->n = n;
    ps->NB = 1;
    for (int p = 0; p < ndim; ++p) {
        LMC_REQUIRE(grid_sizes[p] >= 4, "grid size must be >= 4");
        LMC_REQUIRE(delta[p] > 0 && std::isfinite(delta[p]), "grid spacing must be positive");
        ps->m[p] = grid_sizes[p];
        ps->nb[p] = grid_sizes[p] + 3;
        ps->NB *= ps->nb[p];
    }
    std::vector<int> i0h[2];
    std::vector<double> uh[2];
    std::vector<long> binof((size_t)n);
    for (int p = 0; p < ndim; ++p) {
        i0h[p].resize((size_t)n);
        uh[p].resize((size_t)n);
    }
    for (int d = 0; d < D; ++d) {
        for (long g = ps->out_start[d]; g < ps->out_start[d + 1]; ++g) {
            long bin = 0;
            for (int p = 0; p < ndim; ++p) {
                const double s = X[g * ndim + p];
                LMC_REQUIRE(std::isfinite(s), "non-finite input coordinate");
                const double f = (s - origin[p]) / delta[p];
                const double fl = std::floor(f);
                double lo = -2.0, hi = (double)ps->m[p];
                const double cl = fl < lo ? lo : (fl > hi ? hi : fl);
                i0h[p][g] = (int)cl;
                uh[p][g] = f - fl;
                bin = bin * ps->nb[p] + ((int)cl + 2);
            }
            binof[g] = (long)d * ps->NB + bin;
        }
    }
    const long nbins = (long)D * ps->NB;
    std::vector<int> start((size_t)nbins + 1, 0);
    for (long g = 0; g < n; ++g) start[binof[g] + 1]++;
    for (long b = 0; b < nbins; ++b) start[b + 1] += start[b];
    std::vector<int> cursor(start.begin(), start.end() - 1);
    std::vector<int> perm((size_t)n);
    for (long g = 0; g < n; ++g) perm[cursor[binof[g]]++] = (int)g;
    bool ident = true;
    for (long g = 0; g < n; ++g) ident = ident && perm[g] == (int)g;
    ps->identity = ident;
    std::vector<int> si(n);
    std::vector<double> su(n);
    LMC_CHECK(cudaMalloc(&ps->perm, sizeof(int) * n));
    LMC_CHECK(cudaMemcpy(ps->perm, perm.data(), sizeof(int) * n, cudaMemcpyHostToDevice));
    for (int p = 0; p < ndim; ++p) {
        for (long g = 0; g < n; ++g) {
            si[g] = i0h[p][perm[g]];
            su[g] = uh[p][perm[g]];
        }
        LMC_CHECK(cudaMalloc(&ps->i0[p], sizeof(int) * n));
        LMC_CHECK(cudaMalloc(&ps->u[p], sizeof(double) * n));
        LMC_CHECK(cudaMemcpy(ps->i0[p], si.data(), sizeof(int) * n, cudaMemcpyHostToDevice));
        LMC_CHECK(cudaMemcpy(ps->u[p], su.data(), sizeof(double) * n, cudaMemcpyHostToDevice));
    }
    LMC_CHECK(cudaMalloc(&ps->bin_start, sizeof(int) * (nbins + 1)));
    LMC_CHECK(cudaMemcpy(ps->bin_start, start.data(), sizeof(int) * (nbins + 1), cudaMemcpyHostToDevice));
    LMC_CHECK(cudaMalloc(&ps->out_start_dev, sizeof(long) * (D + 1)));
    LMC_CHECK(cudaMemcpy(ps->out_start_dev, ps->out_start, sizeof(long) * (D + 1), cudaMemcpyHostToDevice));
    return 0;
}

void free_points(PointSet* ps) {
    cudaFree(ps->perm);
    for (int p = 0; p < 2; ++p) {
        cudaFree(ps->i0[p]);
        cudaFree(ps->u[p]);
    }
    cudaFree(ps->bin_start);
    cudaFree(ps->out_start_dev);
    *ps = PointSet();
}

// ---------------------------------------------------------------------------
// device helpers
// ---------------------------------------------------------------------------
// Keys weights of the 4 taps at cells i0-1, i0, i0+1, i0+2 for fractional offset u.
// Written with explicit rounding steps (no FMA contraction) so the weights are
// bit-identical to numpy's evaluation in the reference (interpolation.py:48-52).
__device__ __forceinline__ double keys_near(double x) {  // |x| <= 1
    double t = __dadd_rn(__dmul_rn(1.5, x), -2.5);
    t = __dmul_rn(__dmul_rn(t, x), x);
    return __dadd_rn(t, 1.0);
}
__device__ __forceinline__ double keys_far(double x) {  // 1 < |x| <= 2
    double t = __dadd_rn(__dmul_rn(-0.5, x), 2.5);
    t = __dadd_rn(__dmul_rn(t, x), -4.0);
    return __dadd_rn(__dmul_rn(t, x), 2.0);
}
__device__ __forceinline__ void keys_weights(double u, double* w) {
    const double x0 = __dadd_rn(u, 1.0);
    w[0] = (x0 <= 1.0) ? keys_near(x0) : keys_far(x0);
    w[1] = keys_near(u);
    w[2] = keys_near(fabs(__dadd_rn(u, -1.0)));
    w[3] = keys_far(fabs(__dadd_rn(u, -2.0)));
}
__device__ __forceinline__ int clampi(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }

struct InterpArgs {
    const double* u0;
    const double* u1;
    const int* i00;
    const int* i01;
    const int* perm;
    const int* bin_start;
    const long* out_start;
    const double* in;
    double* out;
    long ld;
    int ncols;
    const double* in_scale;
    const int* active;
    const double* noise;
    cplx* G;
    const cplx* Gc;
    long grid_pitch;
    int D, m0, m1;
    long NB;
    int nb1;
    int tiles, tiles1;
};

__device__ __forceinline__ bool group_active(const InterpArgs& a, int c0, int cnt) {
    if (!a.active) return true;
    bool any = false;
    for (int c = c0; c < c0 + cnt && c < a.ncols; ++c) any = any || a.active[c] != 0;
    return any;
}

// ---------------------------------------------------------------------------
// 1-D scatter: one CTA owns TC = blockDim-3 consecutive cells of one output and
// the TC+3 bins whose stencils touch them; handles PT columns (PT/2 pairs).
// Points stream through shared memory in coalesced chunks.
// ---------------------------------------------------------------------------
static const int kCap1 = 1024;

template <int PT, bool PERM>
__global__ void __launch_bounds__(256) to_grid_1d_kernel(const InterpArgs a) {
    __shared__ double s_u[kCap1];
    __shared__ double s_v[PT][kCap1];  // reused for the per-bin partial sums (needs 4*PT*256 <= PT*1024 + 1024)
    const int TC = blockDim.x - 3;
    const int tile = blockIdx.x % a.tiles;
    const int d = blockIdx.x / a.tiles;
    const int col0 = blockIdx.y * PT;
    if (!group_active(a, col0, PT)) return;
    const int m = a.m0;
    const int c0 = tile * TC;
    const int* bs = a.bin_start + (long)d * a.NB;
    const int last_bin = min(c0 + TC + 2, m + 2);
    const int my_bin = c0 + threadIdx.x;  // bin index = i0 + 2
    const bool has_bin = my_bin <= last_bin;
    const int pbeg = bs[c0], pend = bs[last_bin + 1];
    const int my_beg = has_bin ? bs[my_bin] : 0;
    const int my_end = has_bin ? bs[my_bin + 1] : 0;
    double scale[PT];
#pragma unroll
    for (int p = 0; p < PT; ++p)
        scale[p] = (col0 + p < a.ncols) ? (a.in_scale ? a.in_scale[col0 + p] : 1.0) : 0.0;
    double acc[4][PT];
#pragma unroll
    for (int t = 0; t < 4; ++t)
#pragma unroll
        for (int p = 0; p < PT; ++p) acc[t][p] = 0.0;

    for (int chunk = pbeg; chunk < pend; chunk += kCap1) {
        const int cnt = min(kCap1, pend - chunk);
        for (int i = threadIdx.x; i < cnt; i += blockDim.x) {
            s_u[i] = a.u0[chunk + i];
            const long src = PERM ? (long)a.perm[chunk + i] : (long)(chunk + i);
#pragma unroll
            for (int p = 0; p < PT; ++p)
                s_v[p][i] = (col0 + p < a.ncols) ? a.in[(long)(col0 + p) * a.ld + src] * scale[p] : 0.0;
        }
        __syncthreads();
        const int lo = max(my_beg, chunk) - chunk, hi = min(my_end, chunk + cnt) - chunk;
        for (int i = lo; i < hi; ++i) {
            double w[4];
            keys_weights(s_u[i], w);
#pragma unroll
            for (int p = 0; p < PT; ++p) {
                const double v = s_v[p][i];
#pragma unroll
                for (int t = 0; t < 4; ++t) acc[t][p] = fma(w[t], v, acc[t][p]);
            }
        }
        __syncthreads();
    }
    // exchange partial sums: sA[t][p][bin]
    double* sA = &s_v[0][0];
    double* sA2 = s_u;  // overflow region for PT == 1..: layout below stays within s_v + s_u
    (void)sA2;
    const int nthr = blockDim.x;
#pragma unroll
    for (int t = 0; t < 4; ++t)
#pragma unroll
        for (int p = 0; p < PT; ++p) sA[(t * PT + p) * nthr + threadIdx.x] = acc[t][p];
    __syncthreads();
    const int j = c0 + threadIdx.x;
    if (threadIdx.x < TC && j < m) {
        double sum[PT];
#pragma unroll
        for (int p = 0; p < PT; ++p) sum[p] = 0.0;
        const int ilo = max(j - 2, -2), ihi = min(j + 1, m);
        for (int i0 = ilo; i0 <= ihi; ++i0) {
            const int lb = i0 + 2 - c0;
#pragma unroll
            for (int t = 0; t < 4; ++t) {
                if (clampi(i0 - 1 + t, 0, m - 1) == j) {
#pragma unroll
                    for (int p = 0; p < PT; ++p) sum[p] += sA[(t * PT + p) * nthr + lb];
                }
            }
        }
#pragma unroll
        for (int p = 0; p < PT; p += 2) {
            const int pair = (col0 + p) >> 1;
            if (col0 + p < a.ncols)
                a.G[((long)pair * a.D + d) * a.grid_pitch + j] = make_double2(sum[p], (p + 1 < PT) ? sum[p + 1] : 0.0);
        }
    }
}

// 1-D gather: one thread per bin keeps the 4 taps of PT columns in registers
// and walks its points; results are staged in shared memory so the global
// stores (and the fused  + noise * in  epilogue) are coalesced.
template <int PT, bool PERM>
__global__ void __launch_bounds__(256) from_grid_1d_kernel(const InterpArgs a) {
    __shared__ double s_u[kCap1];
    __shared__ double s_o[PT][kCap1];
    const int tile = blockIdx.x % a.tiles;
    const int d = blockIdx.x / a.tiles;
    const int col0 = blockIdx.y * PT;
    if (!group_active(a, col0, PT)) return;
    const int m = a.m0;
    const int b0 = tile * blockDim.x;
    const int* bs = a.bin_start + (long)d * a.NB;
    const int last_bin = min(b0 + (int)blockDim.x - 1, m + 2);
    const int my_bin = b0 + threadIdx.x;
    const bool has_bin = my_bin <= last_bin;
    const int pbeg = bs[b0], pend = bs[last_bin + 1];
    const int my_beg = has_bin ? bs[my_bin] : 0;
    const int my_end = has_bin ? bs[my_bin + 1] : 0;
    double tap[4][PT];
    if (has_bin && my_end > my_beg) {
        const int i0 = my_bin - 2;
#pragma unroll
        for (int t = 0; t < 4; ++t) {
            const int cell = clampi(i0 - 1 + t, 0, m - 1);
#pragma unroll
            for (int p = 0; p < PT; p += 2) {
                const int pair = (col0 + p) >> 1;
                cplx g = make_double2(0.0, 0.0);
                if (col0 + p < a.ncols) g = a.Gc[((long)pair * a.D + d) * a.grid_pitch + cell];
                tap[t][p] = g.x;
                if (p + 1 < PT) tap[t][p + 1] = g.y;
            }
        }
    }
    const double nz = a.noise ? a.noise[d] : 0.0;
    for (int chunk = pbeg; chunk < pend; chunk += kCap1) {
        const int cnt = min(kCap1, pend - chunk);
        for (int i = threadIdx.x; i < cnt; i += blockDim.x) s_u[i] = a.u0[chunk + i];
        __syncthreads();
        const int lo = max(my_beg, chunk) - chunk, hi = min(my_end, chunk + cnt) - chunk;
        for (int i = lo; i < hi; ++i) {
            double w[4];
            keys_weights(s_u[i], w);
#pragma unroll
            for (int p = 0; p < PT; ++p) {
                double o = 0.0;
#pragma unroll
                for (int t = 0; t < 4; ++t) o = fma(w[t], tap[t][p], o);
                s_o[p][i] = o;
            }
        }
        __syncthreads();
        for (int i = threadIdx.x; i < cnt; i += blockDim.x) {
            const long dst = PERM ? (long)a.perm[chunk + i] : (long)(chunk + i);
#pragma unroll
            for (int p = 0; p < PT; ++p) {
                const int c = col0 + p;
                if (c >= a.ncols) continue;
                if (a.active && !a.active[c]) continue;
                double o = s_o[p][i];
                if (a.noise) {
                    const double sc = a.in_scale ? a.in_scale[c] : 1.0;
                    o = fma(nz, a.in[(long)c * a.ld + dst] * sc, o);
                }
                a.out[(long)c * a.ld + dst] = o;
            }
        }
        __syncthreads();
    }
}

// ---------------------------------------------------------------------------
// 2-D scatter: CTA owns a TX x TY tile of cells of one output and the
// (TX+3) x (TY+3) bins around it; one thread per bin, one RHS pair per CTA.
// ---------------------------------------------------------------------------
static const int kTX = 16, kTY = 16;
static const int kBX = kTX + 3, kBY = kTY + 3;  // 19 x 19 = 361 bins

template <bool PERM>
__global__ void __launch_bounds__(384) to_grid_2d_kernel(const InterpArgs a) {
    extern __shared__ double sA[];  // [16 taps][2][kBX*kBY]
    const int tile = blockIdx.x % a.tiles;
    const int d = blockIdx.x / a.tiles;
    const int pair = blockIdx.y;
    const int col0 = pair * 2;
    if (!group_active(a, col0, 2)) return;
    const int mx = a.m0, my = a.m1;
    const int tx = tile / a.tiles1, ty = tile % a.tiles1;
    const int cx0 = tx * kTX, cy0 = ty * kTY;
    const int nbins = kBX * kBY;
    const int lbx = threadIdx.x / kBY, lby = threadIdx.x % kBY;
    const int bx = cx0 + lbx, by = cy0 + lby;  // bin indices (= i0 + 2)
    double acc[4][4][2];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;
    if (threadIdx.x < nbins && bx <= mx + 2 && by <= my + 2) {
        const int* bs = a.bin_start + (long)d * a.NB + (long)bx * a.nb1 + by;
        const int beg = bs[0], end = bs[1];
        const bool c1 = col0 + 1 < a.ncols;
        const double s0 = a.in_scale ? a.in_scale[col0] : 1.0;
        const double s1 = (c1 && a.in_scale) ? a.in_scale[col0 + 1] : 1.0;
        for (int i = beg; i < end; ++i) {
            double wx[4], wy[4];
            keys_weights(a.u0[i], wx);
            keys_weights(a.u1[i], wy);
            const long src = PERM ? (long)a.perm[i] : (long)i;
            const double v0 = a.in[(long)col0 * a.ld + src] * s0;
            const double v1 = c1 ? a.in[(long)(col0 + 1) * a.ld + src] * s1 : 0.0;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const double y0 = wy[j] * v0, y1 = wy[j] * v1;
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    acc[k][j][0] = fma(wx[k], y0, acc[k][j][0]);
                    acc[k][j][1] = fma(wx[k], y1, acc[k][j][1]);
                }
            }
        }
    }
    if (threadIdx.x < nbins) {
#pragma unroll
        for (int k = 0; k < 4; ++k)
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                sA[((k * 4 + j) * 2 + 0) * nbins + threadIdx.x] = acc[k][j][0];
                sA[((k * 4 + j) * 2 + 1) * nbins + threadIdx.x] = acc[k][j][1];
            }
    }
    __syncthreads();
    if (threadIdx.x < kTX * kTY) {
        const int jx = cx0 + threadIdx.x / kTY, jy = cy0 + threadIdx.x % kTY;
        if (jx < mx && jy < my) {
            double s0 = 0.0, s1 = 0.0;
            const int xlo = max(jx - 2, -2), xhi = min(jx + 1, mx);
            const int ylo = max(jy - 2, -2), yhi = min(jy + 1, my);
            for (int ix = xlo; ix <= xhi; ++ix) {
                for (int k = 0; k < 4; ++k) {
                    if (clampi(ix - 1 + k, 0, mx - 1) != jx) continue;
                    for (int iy = ylo; iy <= yhi; ++iy) {
                        const int lb = (ix + 2 - cx0) * kBY + (iy + 2 - cy0);
                        for (int j = 0; j < 4; ++j) {
                            if (clampi(iy - 1 + j, 0, my - 1) != jy) continue;
                            s0 += sA[((k * 4 + j) * 2 + 0) * nbins + lb];
                            s1 += sA[((k * 4 + j) * 2 + 1) * nbins + lb];
                        }
                    }
                }
            }
            a.G[((long)pair * a.D + d) * a.grid_pitch + (long)jx * my + jy] = make_double2(s0, s1);
        }
    }
}

// 2-D gather: one thread per point (sorted order => neighbouring threads read
// neighbouring cells), one RHS pair per thread.
template <bool PERM>
__global__ void __launch_bounds__(128) from_grid_2d_kernel(const InterpArgs a) {
    const int d = blockIdx.y;
    const int pair = blockIdx.z;
    const int col0 = pair * 2;
    if (!group_active(a, col0, 2)) return;
    const long i = a.out_start[d] + (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= a.out_start[d + 1]) return;
    const int mx = a.m0, my = a.m1;
    double wx[4], wy[4];
    keys_weights(a.u0[i], wx);
    keys_weights(a.u1[i], wy);
    const int ix0 = a.i00[i] - 1, iy0 = a.i01[i] - 1;
    const cplx* g = a.Gc + ((long)pair * a.D + d) * a.grid_pitch;
    int cy[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) cy[j] = clampi(iy0 + j, 0, my - 1);
    double o0 = 0.0, o1 = 0.0;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const long row = (long)clampi(ix0 + k, 0, mx - 1) * my;
        double r0 = 0.0, r1 = 0.0;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const cplx v = __ldg(&g[row + cy[j]]);
            r0 = fma(wy[j], v.x, r0);
            r1 = fma(wy[j], v.y, r1);
        }
        o0 = fma(wx[k], r0, o0);
        o1 = fma(wx[k], r1, o1);
    }
    const long dst = PERM ? (long)a.perm[i] : i;
    const double nz = a.noise ? a.noise[d] : 0.0;
#pragma unroll
    for (int p = 0; p < 2; ++p) {
        const int c = col0 + p;
        if (c >= a.ncols) continue;
        if (a.active && !a.active[c]) continue;
        double o = p ? o1 : o0;
        if (a.noise) {
            const double sc = a.in_scale ? a.in_scale[c] : 1.0;
            o = fma(nz, a.in[(long)c * a.ld + dst] * sc, o);
        }
        a.out[(long)c * a.ld + dst] = o;
    }
}

// ---------------------------------------------------------------------------
// host launchers
// ---------------------------------------------------------------------------
static InterpArgs make_args(const PointSet& ps, const ColumnView& cv) {
    InterpArgs a = {};
    a.u0 = ps.u[0]; a.u1 = ps.u[1];
    a.i00 = ps.i0[0]; a.i01 = ps.i0[1];
    a.perm = ps.perm;
    a.bin_start = ps.bin_start;
    a.out_start = ps.out_start_dev;
    a.in = cv.in; a.out = cv.out; a.ld = cv.ld; a.ncols = cv.ncols;
    a.in_scale = cv.in_scale; a.active = cv.active;
    a.grid_pitch = ps.grid_pitch;
    a.D = ps.D; a.m0 = ps.m[0]; a.m1 = ps.m[1];
    a.NB = ps.NB; a.nb1 = ps.nb[1];
    return a;
}

int to_grid(const PointSet& ps, const ColumnView& cv, cplx* G, cudaStream_t st) {
    if (cv.ncols == 0) return 0;
    InterpArgs a = make_args(ps, cv);
    a.G = G;
    ProfScope prof(PROF_TO_GRID, st);
    const bool perm = !(cv.sorted_io || ps.identity);
    if (ps.ndim == 1) {
        const int threads = 256, TC = threads - 3, PT = 4;
        a.tiles = ceil_div(ps.m[0], TC);
        dim3 grid((unsigned)(a.tiles * ps.D), (unsigned)ceil_div(cv.ncols, PT));
        if (perm) to_grid_1d_kernel<4, true><<<grid, threads, 0, st>>>(a);
        else to_grid_1d_kernel<4, false><<<grid, threads, 0, st>>>(a);
    } else {
        a.tiles1 = ceil_div(ps.m[1], kTY);
        a.tiles = ceil_div(ps.m[0], kTX) * a.tiles1;
        dim3 grid((unsigned)(a.tiles * ps.D), (unsigned)((cv.ncols + 1) / 2));
        const size_t smem = sizeof(double) * 32 * kBX * kBY;
        static bool attr = false;
        if (!attr) {
            LMC_CHECK(cudaFuncSetAttribute(to_grid_2d_kernel<true>,
                                           cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            LMC_CHECK(cudaFuncSetAttribute(to_grid_2d_kernel<false>,
                                           cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            attr = true;
        }
        if (perm) to_grid_2d_kernel<true><<<grid, 384, smem, st>>>(a);
        else to_grid_2d_kernel<false><<<grid, 384, smem, st>>>(a);
    }
    count_launch();
    LMC_CHECK(cudaGetLastError());
    return 0;
}

int from_grid(const PointSet& ps, const ColumnView& cv, const cplx* G, const double* noise,
              cudaStream_t st) {
    if (cv.ncols == 0) return 0;
    InterpArgs a = make_args(ps, cv);
    a.Gc = G;
    ProfScope prof(PROF_FROM_GRID, st);
    a.noise = noise;
    const bool perm = !(cv.sorted_io || ps.identity);
    if (ps.ndim == 1) {
        const int threads = 256, PT = 4;
        a.tiles = ceil_div(ps.m[0] + 3, threads);
        dim3 grid((unsigned)(a.tiles * ps.D), (unsigned)ceil_div(cv.ncols, PT));
        if (perm) from_grid_1d_kernel<4, true><<<grid, threads, 0, st>>>(a);
        else from_grid_1d_kernel<4, false><<<grid, threads, 0, st>>>(a);
    } else {
        long maxlen = 0;
        for (int d = 0; d < ps.D; ++d) maxlen = std::max(maxlen, ps.out_start[d + 1] - ps.out_start[d]);
        if (maxlen == 0) return 0;
        dim3 grid((unsigned)ceil_div(maxlen, 128), (unsigned)ps.D, (unsigned)((cv.ncols + 1) / 2));
        if (perm) from_grid_2d_kernel<true><<<grid, 128, 0, st>>>(a);
        else from_grid_2d_kernel<false><<<grid, 128, 0, st>>>(a);
    }
    count_launch();
    LMC_CHECK(cudaGetLastError());
    return 0;
}

}  // namespace lmc
