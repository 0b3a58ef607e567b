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
    if (ndim == 2) {
        // population of every (16+3) x (16+3) bin window the tiled scatter kernel stages in shared memory
        const int T = 16, B = T + 3;
        int worst = 0;
        for (int d = 0; d < D; ++d)
            for (int cx0 = 0; cx0 < ps->m[0]; cx0 += T)
                for (int cy0 = 0; cy0 < ps->m[1]; cy0 += T) {
                    long cnt = 0;
                    for (int bx = cx0; bx < cx0 + B && bx < ps->nb[0]; ++bx) {
                        const long base = (long)d * ps->NB + (long)bx * ps->nb[1];
                        const int hi = std::min(cy0 + B, ps->nb[1]);
                        cnt += start[base + hi] - start[base + cy0];
                    }
                    worst = std::max<long>(worst, cnt);
                }
        ps->max_tile_pts = worst;
    }
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
// 1-D scatter.  A CTA of 256 threads owns NBIN = 256/TPB consecutive bins of one
// output (TPB = threads per bin, a power of two chosen from the point density so
// that every thread has a few points) and the NBIN-3 cells whose stencils are
// completely covered by them; it handles PT columns (PT/2 RHS pairs).  Points
// stream through shared memory in coalesced chunks; each thread accumulates the
// 4 per-tap partial sums of its share of its bin's points in registers, the TPB
// partials are combined with a fixed-order shuffle tree, exchanged through
// shared memory, and every cell adds up the (bin, tap) pairs that land on it in
// a fixed order.  Deterministic, no atomics.
// ---------------------------------------------------------------------------
static const int kCap1 = 1024;

template <int PT, bool PERM>
__global__ void __launch_bounds__(256) to_grid_1d_kernel(const InterpArgs a, int ltpb) {
    __shared__ double s_u[kCap1];
    __shared__ double s_v[PT][kCap1];  // reused for the per-bin partial sums (4*PT*NBIN <= PT*1024)
    const int tpb = 1 << ltpb;
    const int NBIN = blockDim.x >> ltpb;
    const int TC = NBIN - 3;
    const int tile = blockIdx.x % a.tiles;
    const int d = blockIdx.x / a.tiles;
    const int col0 = blockIdx.y * PT;
    if (!group_active(a, col0, PT)) return;
    const int m = a.m0;
    const int c0 = tile * TC;
    const int* bs = a.bin_start + (long)d * a.NB;
    const int last_bin = min(c0 + TC + 2, m + 2);
    const int lbin = threadIdx.x >> ltpb, sub = threadIdx.x & (tpb - 1);
    const int my_bin = c0 + lbin;  // bin index = i0 + 2
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
                s_v[p][i] = (col0 + p < a.ncols) ? a.in[(long)(col0 + p) * a.ld + src] : 0.0;
        }
        __syncthreads();
        const int lo = max(my_beg, chunk) - chunk, hi = min(my_end, chunk + cnt) - chunk;
        // the bin's points are dealt round-robin to its TPB threads, aligned to the bin start
        int first = lo + ((sub - (lo + chunk - my_beg)) & (tpb - 1));
        for (int i = first; i < hi; i += tpb) {
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
    // combine the TPB partials of a bin (fixed xor tree), apply the column scale
#pragma unroll
    for (int t = 0; t < 4; ++t)
#pragma unroll
        for (int p = 0; p < PT; ++p) {
            double v = acc[t][p];
            for (int o = 1; o < tpb; o <<= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
            acc[t][p] = v * scale[p];
        }
    double* sA = &s_v[0][0];
    if (sub == 0) {
#pragma unroll
        for (int t = 0; t < 4; ++t)
#pragma unroll
            for (int p = 0; p < PT; ++p) sA[(t * PT + p) * NBIN + lbin] = acc[t][p];
    }
    __syncthreads();
    for (int c = threadIdx.x; c < TC; c += blockDim.x) {
        const int j = c0 + c;
        if (j >= m) break;
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
                    for (int p = 0; p < PT; ++p) sum[p] += sA[(t * PT + p) * NBIN + lb];
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

// 1-D gather: the TPB threads of a bin keep the bin's 4 taps of PT columns in
// registers and walk their share of the bin's points; results are staged in
// shared memory so the global stores (and the fused  + noise * in  epilogue) are
// coalesced.
template <int PT, bool PERM>
__global__ void __launch_bounds__(256) from_grid_1d_kernel(const InterpArgs a, int ltpb) {
    __shared__ double s_u[kCap1];
    __shared__ double s_o[PT][kCap1];
    const int tpb = 1 << ltpb;
    const int NBIN = blockDim.x >> ltpb;
    const int tile = blockIdx.x % a.tiles;
    const int d = blockIdx.x / a.tiles;
    const int col0 = blockIdx.y * PT;
    if (!group_active(a, col0, PT)) return;
    const int m = a.m0;
    const int b0 = tile * NBIN;
    const int* bs = a.bin_start + (long)d * a.NB;
    const int last_bin = min(b0 + NBIN - 1, m + 2);
    const int lbin = threadIdx.x >> ltpb, sub = threadIdx.x & (tpb - 1);
    const int my_bin = b0 + lbin;
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
        int first = lo + ((sub - (lo + chunk - my_beg)) & (tpb - 1));
        for (int i = first; i < hi; i += tpb) {
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
// 2-D scatter, v2: "cell-owner gather".  A CTA owns a 16x16 tile of cells and
// stages the points of the 19x19 surrounding bins in shared memory together
// with their 4+4 Keys weights (computed once per point, balanced over threads,
// reused for every RHS pair the CTA loops over).  Each thread then owns one
// cell and sums, in fixed order, the contributions of the points in its 4x4
// neighbouring bins: every point is visited with exactly the one (kx, ky) tap
// that lands on the cell.  Summing 16 bins per thread evens out the Poisson
// imbalance of sparse bins (thread-per-bin wastes ~3.4x of the lanes at 1.5
// points per bin).  Deterministic, no atomics.
// ---------------------------------------------------------------------------
static const int kCap2 = 800;            // staged points per tile (host checks PointSet::max_tile_pts)
static const int kCap2P = kCap2 + 4;     // row pitch: shifts the 4 tap rows onto different banks

struct Tile2Smem {
    double wx[4][kCap2P];
    double wy[4][kCap2P];
    double2 v[2][2][kCap2];  // [buffer][pair of the pass][point]: double-buffered cp.async target
    int src[kCap2];
    int bin[kBX][kBY + 1];   // local offsets of every bin of the window (+ end of row)
    int row_beg[kBX];        // global sorted index of the first point of each bin row
    int row_off[kBX + 1];    // local prefix
};

__device__ __forceinline__ void cp_async8(void* smem_dst, const void* gsrc) {
    const unsigned sa = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;\n" ::"r"(sa), "l"(gsrc));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;\n" ::); }

template <bool PERM>
__global__ void __launch_bounds__(256, 2) to_grid_2d_v2_kernel(const InterpArgs a, int pairs_per_cta) {
    extern __shared__ __align__(16) unsigned char smem_raw2[];
    Tile2Smem& s = *reinterpret_cast<Tile2Smem*>(smem_raw2);
    const int tile = blockIdx.x % a.tiles;
    const int d = blockIdx.x / a.tiles;
    const int pair0 = blockIdx.y * pairs_per_cta;
    const int npairs_tot = (a.ncols + 1) >> 1;
    const int mx = a.m0, my = a.m1;
    const int tx = tile / a.tiles1, ty = tile % a.tiles1;
    const int cx0 = tx * kTX, cy0 = ty * kTY;
    const int* bs = a.bin_start + (long)d * a.NB;
    const int tid = threadIdx.x;

    // ---- window geometry ----
    if (tid < kBX) {
        const int bx = cx0 + tid;
        int beg = 0, end = 0;
        if (bx <= mx + 2) {
            const int by_hi = min(cy0 + kBY, my + 3);
            beg = bs[(long)bx * a.nb1 + cy0];
            end = bs[(long)bx * a.nb1 + by_hi];
        }
        s.row_beg[tid] = beg;
        s.row_off[tid + 1] = end - beg;
    }
    __syncthreads();
    if (tid == 0) {
        s.row_off[0] = 0;
        for (int r = 0; r < kBX; ++r) s.row_off[r + 1] += s.row_off[r];
    }
    __syncthreads();
    for (int i = tid; i < kBX * (kBY + 1); i += blockDim.x) {
        const int r = i / (kBY + 1), c = i % (kBY + 1);
        const int bx = cx0 + r;
        int off = s.row_off[r + 1];
        if (bx <= mx + 2) {
            const int by = min(cy0 + c, my + 3);
            off = s.row_off[r] + (bs[(long)bx * a.nb1 + by] - s.row_beg[r]);
        }
        s.bin[r][c] = off;
    }
    const int npts = s.row_off[kBX];
    const int npass = (min(pairs_per_cta, npairs_tot - pair0) + 1) >> 1;

    // stage the RHS values of pass `ps` (two pairs = four columns) with cp.async; columns past the
    // end of the block are zero-filled by ordinary stores
    auto stage = [&](int ps, int buf) {
        const int pair = pair0 + 2 * ps;
        for (int i = tid; i < npts; i += blockDim.x) {
            const long src = s.src[i];
#pragma unroll
            for (int h = 0; h < 4; ++h) {
                const int c = 2 * pair + h;
                double* dst = reinterpret_cast<double*>(&s.v[buf][h >> 1][i]) + (h & 1);
                if (c < a.ncols && (h < 2 || 2 * ps + 1 < pairs_per_cta)) cp_async8(dst, a.in + (long)c * a.ld + src);
                else *dst = 0.0;
            }
        }
        cp_async_commit();
    };

    // ---- per-point weights (once per CTA) ----
    for (int i = tid; i < npts; i += blockDim.x) {
        int r = 0;
        while (i >= s.row_off[r + 1]) ++r;
        const int g = s.row_beg[r] + (i - s.row_off[r]);
        double wx[4], wy[4];
        keys_weights(a.u0[g], wx);
        keys_weights(a.u1[g], wy);
#pragma unroll
        for (int k = 0; k < 4; ++k) { s.wx[k][i] = wx[k]; s.wy[k][i] = wy[k]; }
        s.src[i] = PERM ? a.perm[g] : g;
    }
    __syncthreads();
    if (npass > 0) stage(0, 0);

    const int lcx = tid / kTY, lcy = tid % kTY;
    const int jx = cx0 + lcx, jy = cy0 + lcy;
    const bool in_grid = jx < mx && jy < my;
    const bool interior = jx >= 1 && jx <= mx - 2 && jy >= 1 && jy <= my - 2;

    for (int ps = 0; ps < npass; ++ps) {
        const int buf = ps & 1;
        const int pair = pair0 + 2 * ps;
        const bool second = (2 * ps + 1 < pairs_per_cta) && (pair + 1 < npairs_tot);
        cp_async_wait_all();
        __syncthreads();                       // pass ps is staged; every thread finished pass ps-1
        if (ps + 1 < npass) stage(ps + 1, buf ^ 1);   // overlaps with the gather below
        if (!in_grid || !group_active(a, 2 * pair, second ? 4 : 2)) continue;
        double acc[4] = {0.0, 0.0, 0.0, 0.0};
        const double2* va_ = s.v[buf][0];
        const double2* vb_ = s.v[buf][1];
        if (interior) {
#pragma unroll
            for (int aa = 0; aa < 4; ++aa) {
                const int r = lcx + aa;          // bin row i0x = jx - 2 + aa  ->  tap kx = 3 - aa
                const double* wxr = s.wx[3 - aa];
                // the 4 bins (r, lcy..lcy+3) are one contiguous run of points; ky from the bin a point is in
                const int b0 = s.bin[r][lcy], b1 = s.bin[r][lcy + 1], b2 = s.bin[r][lcy + 2];
                const int b3 = s.bin[r][lcy + 3], b4 = s.bin[r][lcy + 4];
                for (int i = b0; i < b4; ++i) {
                    const int ky = 3 - ((i >= b1) + (i >= b2) + (i >= b3));
                    const double w = wxr[i] * s.wy[ky][i];
                    const double2 va = va_[i], vb = vb_[i];
                    acc[0] = fma(w, va.x, acc[0]);
                    acc[1] = fma(w, va.y, acc[1]);
                    acc[2] = fma(w, vb.x, acc[2]);
                    acc[3] = fma(w, vb.y, acc[3]);
                }
            }
        } else {
            // grid-edge cell: clamped taps of several bins / several taps of one bin land here
            const int xlo = max(jx - 2, -2), xhi = min(jx + 1, mx);
            const int ylo = max(jy - 2, -2), yhi = min(jy + 1, my);
            for (int ix = xlo; ix <= xhi; ++ix) {
                const int r = ix + 2 - cx0;
                for (int iy = ylo; iy <= yhi; ++iy) {
                    const int c = iy + 2 - cy0;
                    for (int i = s.bin[r][c]; i < s.bin[r][c + 1]; ++i) {
                        double wxs = 0.0, wys = 0.0;
#pragma unroll
                        for (int k = 0; k < 4; ++k) {
                            if (clampi(ix - 1 + k, 0, mx - 1) == jx) wxs += s.wx[k][i];
                            if (clampi(iy - 1 + k, 0, my - 1) == jy) wys += s.wy[k][i];
                        }
                        const double w = wxs * wys;
                        const double2 va = va_[i], vb = vb_[i];
                        acc[0] = fma(w, va.x, acc[0]);
                        acc[1] = fma(w, va.y, acc[1]);
                        acc[2] = fma(w, vb.x, acc[2]);
                        acc[3] = fma(w, vb.y, acc[3]);
                    }
                }
            }
        }
        if (a.in_scale) {   // W^T (v s) = s W^T v: the column scale is applied to the sums
#pragma unroll
            for (int h = 0; h < 4; ++h) {
                const int c = 2 * pair + h;
                if (c < a.ncols) acc[h] *= a.in_scale[c];
            }
        }
        const long cell = (long)jx * my + jy;
        a.G[((long)pair * a.D + d) * a.grid_pitch + cell] = make_double2(acc[0], acc[1]);
        if (second) a.G[((long)(pair + 1) * a.D + d) * a.grid_pitch + cell] = make_double2(acc[2], acc[3]);
    }
}

// 2-D gather, v2: one thread per point loops over the RHS pairs of its group, so the
// 4+4 weights and the 16 clamped cell offsets are computed once per point.
template <bool PERM>
__global__ void __launch_bounds__(128) from_grid_2d_v2_kernel(const InterpArgs a, int pairs_per_cta) {
    const int d = blockIdx.y;
    const long i = a.out_start[d] + (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= a.out_start[d + 1]) return;
    const int mx = a.m0, my = a.m1;
    const int npairs_tot = (a.ncols + 1) >> 1;
    double wx[4], wy[4];
    keys_weights(a.u0[i], wx);
    keys_weights(a.u1[i], wy);
    const int ix0 = a.i00[i] - 1, iy0 = a.i01[i] - 1;
    int cy[4];
    long row[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        cy[j] = clampi(iy0 + j, 0, my - 1);
        row[j] = (long)clampi(ix0 + j, 0, mx - 1) * my;
    }
    const long dst = PERM ? (long)a.perm[i] : i;
    const double nz = a.noise ? a.noise[d] : 0.0;
    const int pair0 = blockIdx.z * pairs_per_cta;
    for (int pp = 0; pp < pairs_per_cta; ++pp) {
        const int pair = pair0 + pp;
        if (pair >= npairs_tot) break;
        const int col0 = 2 * pair;
        if (!group_active(a, col0, 2)) continue;
        const cplx* g = a.Gc + ((long)pair * a.D + d) * a.grid_pitch;
        double o0 = 0.0, o1 = 0.0;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            double r0 = 0.0, r1 = 0.0;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const cplx v = __ldg(&g[row[k] + cy[j]]);
                r0 = fma(wy[j], v.x, r0);
                r1 = fma(wy[j], v.y, r1);
            }
            o0 = fma(wx[k], r0, o0);
            o1 = fma(wx[k], r1, o1);
        }
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
}

// ---------------------------------------------------------------------------
// host launchers
// ---------------------------------------------------------------------------
// threads per bin of the 1-D kernels: keep ~4 points per thread at the average density
static int threads_per_bin_log2(const PointSet& ps) {
    const double density = (double)ps.n / ((double)ps.D * ps.m[0]);
    int l = 0;
    while (l < 5 && density > 4.0 * (1 << l)) ++l;
    return l;
}

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
        const int threads = 256, PT = 4;
        const int ltpb = threads_per_bin_log2(ps);
        const int TC = (threads >> ltpb) - 3;
        a.tiles = ceil_div(ps.m[0], TC);
        dim3 grid((unsigned)(a.tiles * ps.D), (unsigned)ceil_div(cv.ncols, PT));
        if (perm) to_grid_1d_kernel<4, true><<<grid, threads, 0, st>>>(a, ltpb);
        else to_grid_1d_kernel<4, false><<<grid, threads, 0, st>>>(a, ltpb);
    } else {
        a.tiles1 = ceil_div(ps.m[1], kTY);
        a.tiles = ceil_div(ps.m[0], kTX) * a.tiles1;
        const int npairs = (cv.ncols + 1) / 2;
        if (ps.max_tile_pts <= kCap2) {
            static bool attr2 = false;
            if (!attr2) {
                LMC_CHECK(cudaFuncSetAttribute(to_grid_2d_v2_kernel<true>,
                                               cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(Tile2Smem)));
                LMC_CHECK(cudaFuncSetAttribute(to_grid_2d_v2_kernel<false>,
                                               cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(Tile2Smem)));
                attr2 = true;
            }
            // enough CTAs to fill the machine a few times over, as many pairs per CTA as that allows
            const long ctas1 = (long)a.tiles * ps.D;
            int ppc = (int)std::max<long>(1, std::min<long>(8, (ctas1 * npairs) / (148L * 3 * 4)));
            ppc = std::min(ppc, npairs);
            dim3 grid2((unsigned)ctas1, (unsigned)ceil_div(npairs, ppc));
            if (perm) to_grid_2d_v2_kernel<true><<<grid2, 256, sizeof(Tile2Smem), st>>>(a, ppc);
            else to_grid_2d_v2_kernel<false><<<grid2, 256, sizeof(Tile2Smem), st>>>(a, ppc);
            count_launch();
            LMC_CHECK(cudaGetLastError());
            return 0;
        }
        dim3 grid((unsigned)(a.tiles * ps.D), (unsigned)npairs);
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
        const int ltpb = threads_per_bin_log2(ps);
        a.tiles = ceil_div(ps.m[0] + 3, threads >> ltpb);
        dim3 grid((unsigned)(a.tiles * ps.D), (unsigned)ceil_div(cv.ncols, PT));
        if (perm) from_grid_1d_kernel<4, true><<<grid, threads, 0, st>>>(a, ltpb);
        else from_grid_1d_kernel<4, false><<<grid, threads, 0, st>>>(a, ltpb);
    } else {
        long maxlen = 0;
        for (int d = 0; d < ps.D; ++d) maxlen = std::max(maxlen, ps.out_start[d + 1] - ps.out_start[d]);
        if (maxlen == 0) return 0;
        const int npairs = (cv.ncols + 1) / 2;
        const long ctas1 = (long)ceil_div(maxlen, 128) * ps.D;
        int ppc = (int)std::max<long>(1, std::min<long>(8, (ctas1 * npairs) / (148L * 16 * 4)));
        ppc = std::min(ppc, npairs);
        dim3 grid((unsigned)ceil_div(maxlen, 128), (unsigned)ps.D, (unsigned)ceil_div(npairs, ppc));
        if (perm) from_grid_2d_v2_kernel<true><<<grid, 128, 0, st>>>(a, ppc);
        else from_grid_2d_v2_kernel<false><<<grid, 128, 0, st>>>(a, ppc);
    }
    count_launch();
    LMC_CHECK(cudaGetLastError());
    return 0;
}

}  // namespace lmc
