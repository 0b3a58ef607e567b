// Sparse cubic-interpolation gather (W g) and scatter (W^T v) without a stored
// CSR: weights are recomputed from each point's fractional grid offset.
//
// Replaces scipy CSR SpMV `W.dot` / `WT.dot` inside SKI (reference
// runlmc/approx/ski.py:14-16) where W comes from interp_cubic / interp_bicubic /
// multi_interpolant (approx/interpolation.py:56-116, 218-328, 119-176).
//
// Points are sorted once by (output, grid bin).  W^T v is then a segmented
// reduction over contiguous runs of points, staged in shared memory and summed
// per grid cell in a fixed order -- deterministic and free of atomics (kernel
// designs are described above each kernel).  Clamped stencils at the grid edge
// accumulate onto the edge cell exactly like the reference's CSR `+=`
// (interpolation.py:105-115).
#include "interp.cuh"
#include "interp_weights.cuh"

#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <vector>

namespace lmc {

// ---------------------------------------------------------------------------
// host: sort points by bin
// ---------------------------------------------------------------------------
// Largest populations of the point windows the tiled 2-D kernels stage in shared memory, from the bin
// CSR `start` ([D*NB + 1]); they pick the kernel variants at launch time.
void tile_populations(PointSet* ps, const std::vector<int>& start) {
    if (ps->ndim != 2) return;
    const int D = ps->D;
    // population of every (TX+3) x (TY+3) bin window the tiled scatter kernels stage in shared memory
    auto worst_window = [&](int TX, int TY) {
        const int BX = TX + 3, BY = TY + 3;
        long worst = 0;
        for (int d = 0; d < D; ++d)
            for (int cx0 = 0; cx0 < ps->m[0]; cx0 += TX)
                for (int cy0 = 0; cy0 < ps->m[1]; cy0 += TY) {
                    long cnt = 0;
                    for (int bx = cx0; bx < cx0 + BX && bx < ps->nb[0]; ++bx) {
                        const long base = (long)d * ps->NB + (long)bx * ps->nb[1];
                        const int hi = std::min(cy0 + BY, ps->nb[1]);
                        cnt += start[base + hi] - start[base + cy0];
                    }
                    worst = std::max(worst, cnt);
                }
        return (int)worst;
    };
    ps->max_tile_pts_8x8 = worst_window(8, 8);
    ps->max_tile_pts_16x8 = worst_window(16, 8);
    // population of every 16 x 16 tile of bins (no halo) the tiled gather kernel hands to one CTA
    long worst = 0;
    for (int d = 0; d < D; ++d)
        for (int bx0 = 0; bx0 < ps->nb[0]; bx0 += 16)
            for (int by0 = 0; by0 < ps->nb[1]; by0 += 16) {
                long cnt = 0;
                for (int bx = bx0; bx < bx0 + 16 && bx < ps->nb[0]; ++bx) {
                    const long base = (long)d * ps->NB + (long)bx * ps->nb[1];
                    cnt += start[base + std::min(by0 + 16, ps->nb[1])] - start[base + by0];
                }
                worst = std::max(worst, cnt);
            }
    ps->max_gather_tile_pts = (int)worst;
}

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
    tile_populations(ps, start);
    std::vector<int> si(n);
    std::vector<double> su(n);
    LMC_CHECK(cudaMalloc(&ps->perm, sizeof(int) * n));
    LMC_CHECK(cudaMemcpy(ps->perm, perm.data(), sizeof(int) * n, cudaMemcpyHostToDevice));
    if (!ident) {
        std::vector<int> inv((size_t)n);
        for (long g = 0; g < n; ++g) inv[perm[g]] = (int)g;
        LMC_CHECK(cudaMalloc(&ps->iperm, sizeof(int) * n));
        LMC_CHECK(cudaMemcpy(ps->iperm, inv.data(), sizeof(int) * n, cudaMemcpyHostToDevice));
    }
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
    cudaFree(ps->iperm);
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
struct InterpArgs {
    const double* u0;
    const double* u1;
    const int* i00;
    const int* i01;
    const int* perm_in;    // sorted position -> index into the columns of `in` (nullptr: sorted order)
    const int* perm_out;   // same for `out`
    const int* bin_start;
    const long* out_start;
    const double* in;
    double* out;
    long ld, ldo;          // column strides of in / out
    int ncols;
    const double* in_scale;
    const int* active;
    const double* noise;
    cplx* G;
    const cplx* Gc;
    long grid_pitch;
    int D, m0, m1;
    long NB;
    int nb0, nb1;
    int tiles, tiles1;
    int extra;             // 2-D strip scatter: the odd last column rides along with the last group of pairs
    const double* in_rows; // 1-D and 2-D strip scatter: point-major block in[point][column] (row stride ldr), caller's order
    long ldr;
};

__device__ __forceinline__ bool group_active(const InterpArgs& a, int c0, int cnt) {
    if (!a.active) return true;
    bool any = false;
    for (int c = c0; c < c0 + cnt && c < a.ncols; ++c) any = any || a.active[c] != 0;
    return any;
}

__device__ __forceinline__ void cp_async8(void* smem_dst, const void* gsrc) {
    const unsigned sa = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;\n" ::"r"(sa), "l"(gsrc));
}
__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc) {
    const unsigned sa = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(sa), "l"(gsrc));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;\n" ::); }
__device__ __forceinline__ void cp_async_wait_1() { asm volatile("cp.async.wait_group 1;\n" ::); }

// Copies of the 2G columns of one pair group for one staged point, transposed on the way in to
// v[re | im plane][point][pair] (pitch VP = G + 1 doubles: conflict-free for these strided stores and
// for the lane-contiguous loads of the scatter loops).  Columns past the block are zero-filled.
template <int G, int CAP>
__device__ __forceinline__ void stage_point_values(double* v, int i, const double* gp, long ld, int ncol) {
    constexpr int VP = G + 1;
    double* dst = v + i * VP;
#pragma unroll 8
    for (int c = 0; c < ncol; ++c) cp_async8(dst + (c & 1) * (CAP * VP) + (c >> 1), gp + (long)c * ld);
    for (int c = ncol; c < 2 * G; ++c) dst[(c & 1) * (CAP * VP) + (c >> 1)] = 0.0;
}

// Same destination layout from a point-major block (numpy's C order for an [n, P] array): the 2G columns
// of a staged point are 16 G contiguous bytes of its row, so a group of 2G lanes copies one point with one
// coalesced request, whatever row of the caller's array the point lives in -- the permutation into sorted
// order costs nothing extra here (column-major blocks need a pass of their own for it, permute_cols).
template <int G, int CAP>
__device__ __forceinline__ void stage_rows(double* v, const int* src, int npts, const double* rows, long ldr,
                                           int ncol, bool extra, int tid) {
    constexpr int VP = G + 1, LPP = 2 * G, PPW = 32 / LPP;
    const int lane = tid & 31, warp = tid >> 5;
    const int sub = lane / LPP, c = lane % LPP;
    double* dst0 = v + (c & 1) * (CAP * VP) + (c >> 1);
    for (int i = warp * PPW + sub; i < npts; i += 8 * PPW) {
        const double* row = rows + (long)src[i] * ldr;
        if (c < ncol) cp_async8(dst0 + i * VP, row + c);
        else dst0[i * VP] = 0.0;
        if (extra && c == 0) cp_async8(v + i * VP + G, row + 2 * G);
    }
}

// ---------------------------------------------------------------------------
// The scatter kernels (W^T v) are "pair parallel": the lanes of a (part of a)
// warp are G different RHS pairs working on the SAME grid bins / cells.  That
// makes every weight load a shared-memory broadcast, every value load
// v[point][pair] contiguous across lanes (conflict free), and gives all lanes of
// a group the same control flow, whatever the number of points per bin.  (An
// earlier cell-per-lane kernel was bound by shared-memory bank conflicts --
// 2.1 wavefronts per ideal one in ncu -- and Poisson divergence.)
// Every cell adds up its contributions in a fixed order: deterministic, no
// atomics; clamped stencils at the grid edge accumulate onto the edge cell like
// the reference's CSR `+=` (interpolation.py:105-115).
// ---------------------------------------------------------------------------

// ---------------------------------------------------------------------------
// 1-D scatter.  A CTA owns NB = 256 / G consecutive bins of one output and the
// NB - 3 cells whose stencils they cover, for one group of G pairs.  Thread
// (bin, pair) accumulates the bin's 4 per-tap partial sums over the bin's
// (contiguous) points, streamed through shared memory in chunks of CAP points;
// partials are exchanged through shared memory and thread (cell, pair) adds the
// (bin, tap) partials that land on its cell.
// ---------------------------------------------------------------------------
template <int G, int CAP>
struct Scatter1Smem {
    double2 w[CAP][2];            // Keys weights of the 4 taps
    double v[2][CAP * (G + 1)];   // reused for the partial sums
};

template <int G, int CAP, bool ROWS>
__global__ void __launch_bounds__(256, CAP <= 256 ? 4 : 2) to_grid_1d_v3_kernel(const InterpArgs a) {
    typedef Scatter1Smem<G, CAP> Smem;
    constexpr int NBIN = 256 / G, TC = NBIN - 3, VP = G + 1;
    static_assert(CAP <= 512 && 8 * NBIN * G <= 2 * CAP * VP, "staging / exchange layout");
    extern __shared__ __align__(16) unsigned char smem_raw1[];
    Smem& s = *reinterpret_cast<Smem*>(smem_raw1);
    const int tid = threadIdx.x;
    // point-major blocks: the pair groups of one tile are neighbours in launch order (group index fastest), so
    // the 16 G-byte pieces they take out of the same rows meet in L2 instead of being fetched from DRAM per group
    const int bt = ROWS ? blockIdx.y : blockIdx.x;
    const int tile = bt % a.tiles;
    const int d = bt / a.tiles;
    const int grp = ROWS ? blockIdx.x : blockIdx.y;
    const int col0 = 2 * G * grp;
    if (!group_active(a, col0, 2 * G)) return;
    const int ncol = min(2 * G, a.ncols - col0);
    const int m = a.m0;
    const int c0 = tile * TC;                       // first cell == first bin index (bin = i0 + 2)
    const int* bs = a.bin_start + (long)d * a.NB;
    const int last_bin = min(c0 + NBIN - 1, m + 2);
    const int g = tid & (G - 1), lb = tid / G;
    const int my_bin = c0 + lb;
    const bool has_bin = my_bin <= last_bin;
    const int pbeg = bs[c0], pend = bs[last_bin + 1];
    const int my_beg = has_bin ? bs[my_bin] : 0;
    const int my_end = has_bin ? bs[my_bin + 1] : 0;
    double acc[4][2];
#pragma unroll
    for (int t = 0; t < 4; ++t) acc[t][0] = acc[t][1] = 0.0;
    const double* vre = s.v[0] + g;
    const double* vim = s.v[1] + g;

    for (int chunk = pbeg; chunk < pend; chunk += CAP) {
        const int cnt = min(CAP, pend - chunk);
        if (ROWS) {
            // a warp stages its own 32 points: weights by the owning lane, values by groups of 2 G lanes that
            // copy one point's 16 G contiguous bytes each (row index handed round by shuffle)
            constexpr int VP1 = G + 1, LPP = 2 * G, PPW = 32 / LPP;
            const int lane = tid & 31, sub = lane / LPP, c = lane % LPP;
            double* dst0 = s.v[0] + (c & 1) * (CAP * VP1) + (c >> 1);
            for (int ib = (tid & ~31); ib < cnt; ib += 256) {
                const int i = ib + lane;
                long src = 0;
                if (i < cnt) {
                    const int gi = chunk + i;
                    double w[4];
                    keys_weights(a.u0[gi], w);
                    s.w[i][0] = make_double2(w[0], w[1]);
                    s.w[i][1] = make_double2(w[2], w[3]);
                    src = a.perm_in ? (long)a.perm_in[gi] : (long)gi;
                }
#pragma unroll 4
                for (int t = 0; t < 32; t += PPW) {
                    const long r = __shfl_sync(0xffffffffu, src, t + sub);
                    const int j = ib + t + sub;
                    if (j < cnt) {
                        if (c < ncol) cp_async8(dst0 + j * VP1, a.in_rows + r * a.ldr + col0 + c);
                        else dst0[j * VP1] = 0.0;
                    }
                }
            }
        } else {
            for (int i = tid; i < cnt; i += 256) {
                const int gi = chunk + i;
                double w[4];
                keys_weights(a.u0[gi], w);
                s.w[i][0] = make_double2(w[0], w[1]);
                s.w[i][1] = make_double2(w[2], w[3]);
                const long src = a.perm_in ? (long)a.perm_in[gi] : (long)gi;
                stage_point_values<G, CAP>(s.v[0], i, a.in + (long)col0 * a.ld + src, a.ld, ncol);
            }
        }
        cp_async_commit();
        cp_async_wait_all();
        __syncthreads();
        const int lo = max(my_beg, chunk) - chunk, hi = min(my_end, chunk + cnt) - chunk;
#pragma unroll 2
        for (int i = lo; i < hi; ++i) {
            const double2 wa = s.w[i][0], wb = s.w[i][1];
            const double vr = vre[i * VP], vi = vim[i * VP];
            acc[0][0] = fma(wa.x, vr, acc[0][0]); acc[0][1] = fma(wa.x, vi, acc[0][1]);
            acc[1][0] = fma(wa.y, vr, acc[1][0]); acc[1][1] = fma(wa.y, vi, acc[1][1]);
            acc[2][0] = fma(wb.x, vr, acc[2][0]); acc[2][1] = fma(wb.x, vi, acc[2][1]);
            acc[3][0] = fma(wb.y, vr, acc[3][0]); acc[3][1] = fma(wb.y, vi, acc[3][1]);
        }
        __syncthreads();
    }
    // exchange the per-bin partials: part[tap][re|im][bin][pair]
    double* part = s.v[0];
#pragma unroll
    for (int t = 0; t < 4; ++t) {
        part[((t * 2 + 0) * NBIN + lb) * G + g] = acc[t][0];
        part[((t * 2 + 1) * NBIN + lb) * G + g] = acc[t][1];
    }
    __syncthreads();
    const int j = c0 + lb;                          // thread (cell lb, pair g)
    const int pair = grp * G + g;
    const int npairs_tot = (a.ncols + 1) >> 1;
    if (lb < TC && j < m && pair < npairs_tot) {
        const int cA = 2 * pair, cB = 2 * pair + 1;
        if (a.active && !a.active[cA] && !(cB < a.ncols && a.active[cB])) return;
        double sr = 0.0, si = 0.0;
        const int ilo = max(j - 2, -2), ihi = min(j + 1, m);
        for (int i0 = ilo; i0 <= ihi; ++i0) {
            const int b = i0 + 2 - c0;
#pragma unroll
            for (int t = 0; t < 4; ++t) {
                if (clampi(i0 - 1 + t, 0, m - 1) == j) {
                    sr += part[((t * 2 + 0) * NBIN + b) * G + g];
                    si += part[((t * 2 + 1) * NBIN + b) * G + g];
                }
            }
        }
        double s0 = 1.0, s1 = 1.0;   // W^T (v s) = s W^T v: the column scale is applied to the sums
        if (a.in_scale) {
            s0 = a.in_scale[cA];
            s1 = (cB < a.ncols) ? a.in_scale[cB] : 0.0;
        }
        a.G[((long)pair * a.D + d) * a.grid_pitch + j] = make_double2(sr * s0, si * s1);
    }
}

// ---------------------------------------------------------------------------
// 1-D gather: one thread per point (sorted order: neighbouring lanes read the
// same or neighbouring cells, which L1 serves), looping over the RHS pairs so
// the weights are computed once per point.  Epilogue: + noise_d * in.
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(256) from_grid_1d_v3_kernel(const InterpArgs a, int pairs_per_cta) {
    const int d = blockIdx.y;
    const long i = a.out_start[d] + (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= a.out_start[d + 1]) return;
    const int m = a.m0;
    const int npairs_tot = (a.ncols + 1) >> 1;
    double w[4];
    keys_weights(a.u0[i], w);
    const int i0 = a.i00[i];
    int cell[4];
#pragma unroll
    for (int t = 0; t < 4; ++t) cell[t] = clampi(i0 - 1 + t, 0, m - 1);
    const long si = a.perm_in ? (long)a.perm_in[i] : i;
    const long so = a.perm_out ? (long)a.perm_out[i] : i;
    const double nz = a.noise ? a.noise[d] : 0.0;
    const int pair0 = blockIdx.z * pairs_per_cta;
    const int pair1 = min(npairs_tot, pair0 + pairs_per_cta);
    const cplx* gbase = a.Gc + (long)d * a.grid_pitch;
    const long pstride = (long)a.D * a.grid_pitch;
    // Chunks of CH pairs: all global loads of a chunk (inputs of the noise term and grid cells) are
    // issued before its first store -- the compiler cannot hoist loads over stores to `out` (possible
    // alias), and one pair at a time left the kernel waiting on DRAM (82 % long-scoreboard stalls).
    constexpr int CH = 4;
    for (int pc = pair0; pc < pair1; pc += CH) {
        double inA[CH], inB[CH];
        cplx v[CH][4];
        bool actA[CH], actB[CH];
#pragma unroll
        for (int u = 0; u < CH; ++u) {
            const int pair = pc + u;
            const int cA = 2 * pair, cB = cA + 1;
            const bool in_rng = pair < pair1;
            actA[u] = in_rng && (!a.active || a.active[cA]);
            actB[u] = in_rng && cB < a.ncols && (!a.active || a.active[cB]);
            inA[u] = inB[u] = 0.0;
            if (a.noise) {
                if (actA[u]) inA[u] = a.in[(long)cA * a.ld + si];
                if (actB[u]) inB[u] = a.in[(long)cB * a.ld + si];
            }
            if (actA[u] || actB[u]) {
                const cplx* gp = gbase + pair * pstride;
#pragma unroll
                for (int t = 0; t < 4; ++t) v[u][t] = __ldg(gp + cell[t]);
            }
        }
#pragma unroll
        for (int u = 0; u < CH; ++u) {
            if (!actA[u] && !actB[u]) continue;
            const int cA = 2 * (pc + u), cB = cA + 1;
            double iA = inA[u], iB = inB[u];
            if (a.noise && a.in_scale) { iA *= a.in_scale[cA]; if (cB < a.ncols) iB *= a.in_scale[cB]; }
            double oA = fma(w[3], v[u][3].x, fma(w[2], v[u][2].x, fma(w[1], v[u][1].x, w[0] * v[u][0].x)));
            double oB = fma(w[3], v[u][3].y, fma(w[2], v[u][2].y, fma(w[1], v[u][1].y, w[0] * v[u][0].y)));
            if (a.noise) { oA = fma(nz, iA, oA); oB = fma(nz, iB, oB); }
            if (actA[u]) a.out[(long)cA * a.ldo + so] = oA;
            if (actB[u]) a.out[(long)cB * a.ldo + so] = oB;
        }
    }
}

// 8 x 8 transpose across each group of 8 lanes: before, lane t of a group holds acc[p] = element (t, p);
// afterwards acc[i] = element (i, g) for its own position g in the group.  Three butterfly stages of
// __shfl_xor, every register index a compile-time constant.
__device__ __forceinline__ void lane_group_transpose8(double (&acc)[8][2], int g) {
#pragma unroll
    for (int sft = 1; sft < 8; sft <<= 1) {
        const bool up = (g & sft) != 0;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            if (i & sft) continue;
            const double sx = up ? acc[i][0] : acc[i | sft][0];
            const double sy = up ? acc[i][1] : acc[i | sft][1];
            const double rx = __shfl_xor_sync(0xffffffffu, sx, sft);
            const double ry = __shfl_xor_sync(0xffffffffu, sy, sft);
            if (up) { acc[i][0] = rx; acc[i][1] = ry; }
            else { acc[i | sft][0] = rx; acc[i | sft][1] = ry; }
        }
    }
}

// Row pieces of 8 points after the transpose: lane g of a group holds pair `pair` of the points whose rows
// the lanes gb .. gb + 7 own (row index `so`, liveness bit in `hv`); 8 lanes store 128 contiguous bytes
// of one row and read the input row beside them for the noise term.
__device__ __forceinline__ void store_row_pieces8(const double (&acc)[8][2], int so, unsigned hv, int gb, int cA,
                                                  bool okA, bool okB, bool noise, double nz,
                                                  const double* __restrict__ xin, long ldr,
                                                  double* __restrict__ yout, long ldo) {
#pragma unroll
    for (int h = 0; h < 8; h += 4) {
        long row[4];
        double x0[4], x1[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            row[i] = __shfl_sync(0xffffffffu, so, gb + h + i);
            const bool live = (hv >> (gb + h + i)) & 1u;
            x0[i] = (noise && live && okA) ? __ldcs(xin + row[i] * ldr + cA) : 0.0;
            x1[i] = (noise && live && okB) ? __ldcs(xin + row[i] * ldr + cA + 1) : 0.0;
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const bool live = (hv >> (gb + h + i)) & 1u;
            double v0 = acc[h + i][0], v1 = acc[h + i][1];
            if (noise) { v0 = fma(nz, x0[i], v0); v1 = fma(nz, x1[i], v1); }
            if (live && okA) __stcs(yout + row[i] * ldo + cA, v0);
            if (live && okB) __stcs(yout + row[i] * ldo + cA + 1, v1);
        }
    }
}

// The input-row pieces store_row_pieces8 is going to read, requested into L2 ahead of the tap arithmetic
// (no registers held; the two end lanes of a group cover the 128-byte piece).
__device__ __forceinline__ void prefetch_row_pieces8(int so, unsigned hv, int gb, int g, int cA, bool okA,
                                                     const double* __restrict__ xin, long ldr) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const long row = __shfl_sync(0xffffffffu, so, gb + i);
        if ((g == 0 || g == 7) && okA && ((hv >> (gb + i)) & 1u))
            asm volatile("prefetch.global.L2 [%0];" ::"l"(xin + row * ldr + cA));
    }
}

// 1-D gather writing a point-major block (out[point][column], the caller's rows): results of 8 pairs per
// point, transposed across each group of 8 lanes, leave as 128-byte row pieces (see the 2-D variant below).
__global__ void __launch_bounds__(256) from_grid_1d_rows_kernel(const InterpArgs a, int pairs_per_cta) {
    const int d = blockIdx.y;
    const long end = a.out_start[d + 1];
    const long i_raw = a.out_start[d] + (long)blockIdx.x * blockDim.x + threadIdx.x;
    const bool have = i_raw < end;
    const unsigned hv = __ballot_sync(0xffffffffu, have);
    if (hv == 0) return;
    const long i = have ? i_raw : end - 1;           // idle lanes shadow a live point and store nothing
    const int m = a.m0;
    const int npairs_tot = (a.ncols + 1) >> 1;
    double w[4];
    keys_weights(a.u0[i], w);
    const int i0 = a.i00[i];
    int cell[4];
#pragma unroll
    for (int t = 0; t < 4; ++t) cell[t] = clampi(i0 - 1 + t, 0, m - 1);
    const int so = a.perm_out ? a.perm_out[i] : (int)i;
    const double nz = a.noise ? a.noise[d] : 0.0;
    const int pair0 = blockIdx.z * pairs_per_cta;
    const int pair1 = min(npairs_tot, pair0 + pairs_per_cta);
    const cplx* gbase = a.Gc + (long)d * a.grid_pitch;
    const long pstride = (long)a.D * a.grid_pitch;
    const int lane = threadIdx.x & 31, gb = lane & ~7, g = lane & 7;
    for (int pc = pair0; pc < pair1; pc += 8) {
        if (a.noise) prefetch_row_pieces8(so, hv, gb, g, 2 * (pc + g), pc + g < pair1, a.in_rows, a.ldr);
        double acc[8][2];
#pragma unroll
        for (int hf = 0; hf < 8; hf += 4) {
            cplx v[4][4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int pair = min(pc + hf + u, pair1 - 1);
                const cplx* gp = gbase + pair * pstride;
#pragma unroll
                for (int t = 0; t < 4; ++t) v[u][t] = __ldg(gp + cell[t]);
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                acc[hf + u][0] = fma(w[3], v[u][3].x, fma(w[2], v[u][2].x, fma(w[1], v[u][1].x, w[0] * v[u][0].x)));
                acc[hf + u][1] = fma(w[3], v[u][3].y, fma(w[2], v[u][2].y, fma(w[1], v[u][1].y, w[0] * v[u][0].y)));
            }
        }
        lane_group_transpose8(acc, g);
        const int pair = pc + g;
        const int cA = 2 * pair;
        const bool okA = pair < pair1, okB = okA && cA + 1 < a.ncols;
        store_row_pieces8(acc, so, hv, gb, cA, okA, okB, a.noise != nullptr, nz, a.in_rows, a.ldr, a.out, a.ldo);
    }
}

// ---------------------------------------------------------------------------
// 2-D scatter, fallback for badly clustered points (no per-tile capacity):
// CTA owns a 16 x 16 tile of cells of one output and the 19 x 19 bins around
// it; one thread per bin, one RHS pair per CTA; per-bin 4x4 tap partials are
// exchanged through shared memory.
// ---------------------------------------------------------------------------
static const int kTX = 16, kTY = 16;
static const int kBX = kTX + 3, kBY = kTY + 3;  // 19 x 19 = 361 bins

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
            const long src = a.perm_in ? (long)a.perm_in[i] : (long)i;
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

// ---------------------------------------------------------------------------
// 2-D scatter, "pair-parallel strips".  Thread (strip, pair): a strip is 4 cells
// consecutive in x.  A CTA stages the points of the (TX+3) x (TY+3) bins around
// its TX x TY tile of cells once (weights) and their values once per group of G
// pairs.  One visit of a point serves up to 4 cells (7 visits per point instead
// of 16 for cell-per-thread ownership), x taps are compile-time constants of the
// unrolled bin-row / bin loops.
// ---------------------------------------------------------------------------
template <int G, int TX, int TY, int CAP>
struct Tile3Smem {
    static const int BX = TX + 3, BY = TY + 3;
    double wx[4][CAP];
    double wy[4][CAP];
    double v[2][CAP * (G + 1)];   // [re | im][point][pair]
    int src[CAP];
    unsigned char by[CAP];        // bin column of the point inside the window
    int bin[BX][BY + 1];          // local offset of every bin of the window (+ end of row)
    int row_beg[BX];
    int row_off[BX + 1];
    int live[64];                 // per group of this CTA: any column still active
};

template <int G, int TX, int TY, int CAP, bool ROWS>
__global__ void __launch_bounds__(256, 2) to_grid_2d_v3_kernel(const InterpArgs a, int groups_per_cta) {
    typedef Tile3Smem<G, TX, TY, CAP> Smem;
    static_assert((TX / 4) * TY * G == 256, "one thread per (strip, pair)");
    static_assert(CAP <= 512, "two staged points per thread");
    constexpr int BX = Smem::BX, BY = Smem::BY, VP = G + 1;
    extern __shared__ __align__(16) unsigned char smem_raw3[];
    Smem& s = *reinterpret_cast<Smem*>(smem_raw3);
    const int tile = blockIdx.x % a.tiles;
    const int d = blockIdx.x / a.tiles;
    // a.extra: ncols = 2 G k + 1 -- the k full groups carry the odd last column as a third accumulator of the
    // last group (all lanes of a strip compute it, lane 0 stores it) instead of a nearly empty group of its own
    const int npairs_tot = a.extra ? (a.ncols >> 1) : ((a.ncols + 1) >> 1);
    const int ngroups = (npairs_tot + G - 1) / G;
    const int grp0 = blockIdx.y * groups_per_cta;
    const int grp1 = min(ngroups, grp0 + groups_per_cta);
    const int mx = a.m0, my = a.m1;
    const int tx = tile / a.tiles1, ty = tile % a.tiles1;
    const int cx0 = tx * TX, cy0 = ty * TY;
    const int* bs = a.bin_start + (long)d * a.NB;
    const int tid = threadIdx.x;

    // ---- window geometry (bin index = i0 + 2; cell j collects from bins j .. j+3) ----
    if (tid < BX) {
        const int bx = cx0 + tid;
        int beg = 0, end = 0;
        if (bx <= mx + 2) {
            const int by_hi = min(cy0 + BY, my + 3);
            beg = bs[(long)bx * a.nb1 + cy0];
            end = bs[(long)bx * a.nb1 + by_hi];
        }
        s.row_beg[tid] = beg;
        s.row_off[tid + 1] = end - beg;
    }
    // one thread per group looks at the group's activity flags (instead of every thread, every group)
    if (tid >= 64 && tid < 128 && grp0 + (tid - 64) < grp1)
        s.live[tid - 64] = group_active(a, 2 * G * (grp0 + tid - 64),
                                        2 * G + ((a.extra && grp0 + tid - 64 == ngroups - 1) ? 1 : 0)) ? 1 : 0;
    __syncthreads();
    if (tid == 0) {
        s.row_off[0] = 0;
        for (int r = 0; r < BX; ++r) s.row_off[r + 1] += s.row_off[r];
    }
    __syncthreads();
    for (int i = tid; i < BX * (BY + 1); i += blockDim.x) {
        const int r = i / (BY + 1), c = i % (BY + 1);
        const int bx = cx0 + r;
        int off = s.row_off[r + 1];
        if (bx <= mx + 2) {
            const int by = min(cy0 + c, my + 3);
            off = s.row_off[r] + (bs[(long)bx * a.nb1 + by] - s.row_beg[r]);
        }
        s.bin[r][c] = off;
    }
    const int npts = s.row_off[BX];
    // ---- per-point weights, once per CTA ----
    for (int i = tid; i < npts; i += blockDim.x) {
        int r = 0;
        while (i >= s.row_off[r + 1]) ++r;
        const int gidx = s.row_beg[r] + (i - s.row_off[r]);
        double wx[4], wy[4];
        keys_weights(a.u0[gidx], wx);
        keys_weights(a.u1[gidx], wy);
#pragma unroll
        for (int k = 0; k < 4; ++k) { s.wx[k][i] = wx[k]; s.wy[k][i] = wy[k]; }
        s.src[i] = a.perm_in ? a.perm_in[gidx] : gidx;
        s.by[i] = (unsigned char)(a.i01[gidx] + 2 - cy0);
    }
    __syncthreads();

    const int i_a = tid, i_b = tid + 256;   // the (up to) two staged points this thread copies
    const long src_a = (!ROWS && i_a < npts) ? s.src[i_a] : 0, src_b = (!ROWS && i_b < npts) ? s.src[i_b] : 0;
    const int g = tid & (G - 1);
    const int strip = tid / G;
    const int sy = strip % TY, sx = strip / TY;
    const int jx0 = cx0 + 4 * sx, jy = cy0 + sy;
    const bool in_grid = jx0 < mx && jy < my;
    const bool interior = jx0 >= 1 && jx0 + 3 <= mx - 2 && jy >= 1 && jy <= my - 2;
    const double* vre = s.v[0] + g;
    const double* vim = s.v[1] + g;
    const double* vex = s.v[0] + G;        // the padding slot of every staged point holds the extra column

    for (int grp = grp0; grp < grp1; ++grp) {
        const int col0 = 2 * G * grp;
        const bool live = s.live[grp - grp0] != 0;           // uniform over the CTA
        const bool extra = a.extra && grp == ngroups - 1;    // uniform over the CTA
        if (live) {
            const int ncol = min(2 * G + (extra ? 1 : 0), a.ncols - col0);
            if (ROWS) {
                stage_rows<G, CAP>(s.v[0], s.src, npts, a.in_rows + col0, a.ldr, min(ncol, 2 * G), extra, tid);
            } else {
                const double* gp = a.in + (long)col0 * a.ld;
                if (i_a < npts) stage_point_values<G, CAP>(s.v[0], i_a, gp + src_a, a.ld, ncol);
                if (i_b < npts) stage_point_values<G, CAP>(s.v[0], i_b, gp + src_b, a.ld, ncol);
            }
            cp_async_commit();
            cp_async_wait_all();
        }
        __syncthreads();
        const int pair = grp * G + g;
        if (live && in_grid && pair < npairs_tot) {
            double acc[4][2], acx[4];
#pragma unroll
            for (int c = 0; c < 4; ++c) acc[c][0] = acc[c][1] = acx[c] = 0.0;
            if (interior && extra) {
#pragma unroll
                for (int aa = 0; aa < 7; ++aa) {
                    const int* brow = &s.bin[4 * sx + aa][sy];
                    const int ibeg = brow[0], iend = brow[4];
                    const double* pw = &s.wx[0][0] + ibeg;
                    const double* pwy = &s.wy[0][0] + (3 + sy) * CAP + ibeg;
                    const double* pv = vre + ibeg * VP;
                    const double* px = vex + ibeg * VP;
                    const unsigned char* pby = s.by + ibeg;
                    const unsigned char* pby_end = s.by + iend;
#pragma unroll 2
                    for (; pby < pby_end; ++pby, ++pw, ++pwy, pv += VP, px += VP) {
                        const double wyv = pwy[-(int)pby[0] * CAP];
                        const double t0 = wyv * pv[0], t1 = wyv * pv[CAP * VP], t2 = wyv * px[0];
#pragma unroll
                        for (int c = 0; c < 4; ++c) {
                            const int kx = c - aa + 3;
                            if (kx >= 0 && kx <= 3) {
                                const double wxv = pw[kx * CAP];
                                acc[c][0] = fma(wxv, t0, acc[c][0]);
                                acc[c][1] = fma(wxv, t1, acc[c][1]);
                                acx[c] = fma(wxv, t2, acx[c]);
                            }
                        }
                    }
                }
            } else if (interior) {
#pragma unroll
                for (int aa = 0; aa < 7; ++aa) {
                    // bin row 4 sx + aa: cell c of the strip takes x tap c - aa + 3 from it
                    // the 4 bins (row, sy .. sy + 3) are one contiguous run of points; the point in bin
                    // column by feeds cell jy with y tap 3 - (by - sy)
                    const int* brow = &s.bin[4 * sx + aa][sy];
                    const int ibeg = brow[0], iend = brow[4];
                    // running pointers (one add each per point) instead of index arithmetic:
                    // wx[kx][i] = pw[kx * CAP], wy[3 + sy - by][i] = pwy[-by * CAP], v = pv[0], pv[CAP * VP]
                    const double* pw = &s.wx[0][0] + ibeg;
                    const double* pwy = &s.wy[0][0] + (3 + sy) * CAP + ibeg;
                    const double* pv = vre + ibeg * VP;
                    const unsigned char* pby = s.by + ibeg;
                    const unsigned char* pby_end = s.by + iend;
#pragma unroll 2   // two points in flight: their shared-memory loads overlap (1.43 -> 1.41 ms at config E)
                    for (; pby < pby_end; ++pby, ++pw, ++pwy, pv += VP) {
                        const double wyv = pwy[-(int)pby[0] * CAP];
                        const double t0 = wyv * pv[0], t1 = wyv * pv[CAP * VP];
#pragma unroll
                        for (int c = 0; c < 4; ++c) {
                            const int kx = c - aa + 3;
                            if (kx >= 0 && kx <= 3) {
                                const double wxv = pw[kx * CAP];
                                acc[c][0] = fma(wxv, t0, acc[c][0]);
                                acc[c][1] = fma(wxv, t1, acc[c][1]);
                            }
                        }
                    }
                }
            } else {
                // strip touching the grid edge: clamped taps of several bins / several taps of one bin
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    const int jx = jx0 + c;
                    if (jx >= mx) continue;
                    const int xlo = max(jx - 2, -2), xhi = min(jx + 1, mx);
                    const int ylo = max(jy - 2, -2), yhi = min(jy + 1, my);
                    double a0 = 0.0, a1 = 0.0, a2 = 0.0;
                    for (int ix = xlo; ix <= xhi; ++ix) {
                        const int r = ix + 2 - cx0;
                        for (int iy = ylo; iy <= yhi; ++iy) {
                            const int cc = iy + 2 - cy0;
#pragma unroll 1
                            for (int i = s.bin[r][cc]; i < s.bin[r][cc + 1]; ++i) {
                                double wxs = 0.0, wys = 0.0;
#pragma unroll
                                for (int k = 0; k < 4; ++k) {
                                    if (clampi(ix - 1 + k, 0, mx - 1) == jx) wxs += s.wx[k][i];
                                    if (clampi(iy - 1 + k, 0, my - 1) == jy) wys += s.wy[k][i];
                                }
                                const double w = wxs * wys;
                                a0 = fma(w, vre[i * VP], a0);
                                a1 = fma(w, vim[i * VP], a1);
                                if (extra) a2 = fma(w, vex[i * VP], a2);
                            }
                        }
                    }
                    acc[c][0] = a0;
                    acc[c][1] = a1;
                    acx[c] = a2;
                }
            }
            if (extra && g == 0) {
                // the odd last column: pair slot npairs_tot, imaginary part zero
                const int cE = 2 * npairs_tot;
                if (!a.active || a.active[cE]) {
                    const double sE = a.in_scale ? a.in_scale[cE] : 1.0;
                    cplx* gpx = a.G + ((long)npairs_tot * a.D + d) * a.grid_pitch + jy;
#pragma unroll
                    for (int c = 0; c < 4; ++c)
                        if (jx0 + c < mx) gpx[(long)(jx0 + c) * my] = make_double2(acx[c] * sE, 0.0);
                }
            }
            const int cA = 2 * pair, cB = 2 * pair + 1;
            const bool act = !a.active || a.active[cA] || (cB < a.ncols && a.active[cB]);
            if (act) {
                double s0 = 1.0, s1 = 1.0;   // W^T (v s) = s W^T v: the column scale is applied to the sums
                if (a.in_scale) {
                    s0 = a.in_scale[cA];
                    s1 = (cB < a.ncols) ? a.in_scale[cB] : 0.0;
                }
                cplx* gp = a.G + ((long)pair * a.D + d) * a.grid_pitch + jy;
#pragma unroll
                for (int c = 0; c < 4; ++c)
                    if (jx0 + c < mx) gp[(long)(jx0 + c) * my] = make_double2(acc[c][0] * s0, acc[c][1] * s1);
            }
        }
        __syncthreads();   // every thread is done with this group's values
    }
}

// ---------------------------------------------------------------------------
// 2-D gather, fallback without a per-tile capacity: one thread per point loops
// over the RHS pairs of its group and reads the 16 cells through L1/L2.
// ---------------------------------------------------------------------------
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
    const long si = a.perm_in ? (long)a.perm_in[i] : i;
    const long so = a.perm_out ? (long)a.perm_out[i] : i;
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
                o = fma(nz, a.in[(long)c * a.ld + si] * sc, o);
            }
            a.out[(long)c * a.ldo + so] = o;
        }
    }
}

// ---------------------------------------------------------------------------
// 2-D gather with the cells staged in shared memory.  A CTA owns the points of a
// TB x TB tile of bins of one output (<= 512 points, two per thread) and the
// (TB+3)^2 cells their stencils touch; the cells of GP pairs at a time are copied
// in with cp.async (double buffered: the copies of the next pass overlap the
// arithmetic of this one), so the 16 taps per point and pair are shared-memory
// reads instead of L1 misses served by L2 (ncu on the L1 variant: 86 % of the
// stalls were waits on those loads, L2 -> SM traffic ~7x the cell data).
// ---------------------------------------------------------------------------
template <int GP, int TB>
struct Gather3Smem {
    static const int W = TB + 3;
    cplx cell[2][GP][W * W];
    int row_beg[TB];
    int row_off[TB + 1];
    double scale[272];           // per column of this CTA's pair range: in_scale (1 if none)
    unsigned char act[272];      // ... and activity flag (0 also for columns past the block)
};

template <int GP, int TB>
__global__ void __launch_bounds__(256, 2) from_grid_2d_v3_kernel(const InterpArgs a, int passes_per_cta) {
    typedef Gather3Smem<GP, TB> Smem;
    constexpr int W = Smem::W;
    extern __shared__ __align__(16) unsigned char smem_raw4[];
    Smem& s = *reinterpret_cast<Smem*>(smem_raw4);
    const int tid = threadIdx.x;
    const int tile = blockIdx.x % a.tiles;
    const int d = blockIdx.x / a.tiles;
    const int mx = a.m0, my = a.m1;
    const int BX0 = (tile / a.tiles1) * TB, BY0 = (tile % a.tiles1) * TB;   // first bin (bin = i0 + 2)
    const int X0 = BX0 - 3, Y0 = BY0 - 3;                                   // first cell of the window
    const int* bs = a.bin_start + (long)d * a.NB;
    const int npairs_tot = (a.ncols + 1) >> 1;
    const int pair_lo = blockIdx.y * passes_per_cta * GP;
    const int pair_hi = min(npairs_tot, pair_lo + passes_per_cta * GP);
    if (tid < TB) {
        const int bx = BX0 + tid;
        int beg = 0, end = 0;
        if (bx < a.nb0) {
            beg = bs[(long)bx * a.nb1 + BY0];
            end = bs[(long)bx * a.nb1 + min(BY0 + TB, a.nb1)];
        }
        s.row_beg[tid] = beg;
        s.row_off[tid + 1] = end - beg;
    }
    for (int c = tid; c < 2 * (pair_hi - pair_lo) + 2 && c < 272; c += 256) {
        const int col = 2 * pair_lo + c;
        const bool in = col < a.ncols;
        s.act[c] = (in && (!a.active || a.active[col])) ? 1 : 0;
        s.scale[c] = (in && a.in_scale) ? a.in_scale[col] : 1.0;
    }
    __syncthreads();
    if (tid == 0) {
        s.row_off[0] = 0;
        for (int r = 0; r < TB; ++r) s.row_off[r + 1] += s.row_off[r];
    }
    __syncthreads();
    const int npts = s.row_off[TB];
    if (npts == 0) return;

    auto stage = [&](int pbase, int buf) {
        for (int p = 0; p < GP; ++p) {
            const int pair = pbase + p;
            if (pair >= pair_hi) break;
            if (!(s.act[2 * (pair - pair_lo)] | s.act[2 * (pair - pair_lo) + 1])) continue;
            const cplx* gsl = a.Gc + ((long)pair * a.D + d) * a.grid_pitch;
            // window cell (x, y) holds G[clamp(X0 + x)][clamp(Y0 + y)]: stencils index the window
            // without clamping and still pick up the edge cell for every clamped tap
            // (interpolation.py:105-115), so the tap addresses below are compile-time offsets
            for (int c = tid; c < W * W; c += 256) {
                const int x = c / W, y = c - x * W;
                const int gx = clampi(X0 + x, 0, mx - 1), gy = clampi(Y0 + y, 0, my - 1);
                cp_async16(&s.cell[buf][p][c], gsl + (long)gx * my + gy);
            }
        }
        cp_async_commit();
    };
    stage(pair_lo, 0);

    // ---- this thread's (up to) two points ----
    double wx[2][4], wy[2][4];
    int o0[2];       // window offset of the point's first tap
    long si[2], so[2];
    bool have[2];
#pragma unroll
    for (int q = 0; q < 2; ++q) {
        const int li = tid + 256 * q;
        have[q] = li < npts;
        si[q] = so[q] = 0;
        if (have[q]) {
            int r = 0;
            while (li >= s.row_off[r + 1]) ++r;
            const long gi = s.row_beg[r] + (li - s.row_off[r]);
            keys_weights(a.u0[gi], wx[q]);
            keys_weights(a.u1[gi], wy[q]);
            const int ix0 = a.i00[gi] - 1, iy0 = a.i01[gi] - 1;
            o0[q] = (ix0 - X0) * W + (iy0 - Y0);
            si[q] = a.perm_in ? (long)a.perm_in[gi] : gi;
            so[q] = a.perm_out ? (long)a.perm_out[gi] : gi;
        } else {
#pragma unroll
            for (int k = 0; k < 4; ++k) wx[q][k] = wy[q][k] = 0.0;
            o0[q] = 0;
        }
    }
    const double nz = a.noise ? a.noise[d] : 0.0;

    // noise term inputs in[c][i] of the two columns of a pair, fetched one pair ahead of their use so
    // the global-load latency hides behind the 16-tap arithmetic of the current pair
    auto fetch_in = [&](int pair, double (&v)[2][2]) {
#pragma unroll
        for (int q = 0; q < 2; ++q) v[q][0] = v[q][1] = 0.0;
        if (!a.noise || pair >= pair_hi) return;
        const int cA = 2 * pair, cB = cA + 1;
        const bool actA = s.act[cA - 2 * pair_lo], actB = s.act[cB - 2 * pair_lo];
#pragma unroll
        for (int q = 0; q < 2; ++q) {
            if (!have[q]) continue;
            if (actA) v[q][0] = a.in[(long)cA * a.ld + si[q]];
            if (actB) v[q][1] = a.in[(long)cB * a.ld + si[q]];
        }
    };
    double cur[2][2], nxt[2][2];
    fetch_in(pair_lo, cur);

    int buf = 0;
    for (int pbase = pair_lo; pbase < pair_hi; pbase += GP, buf ^= 1) {
        if (pbase + GP < pair_hi) { stage(pbase + GP, buf ^ 1); cp_async_wait_1(); }
        else cp_async_wait_all();
        __syncthreads();
#pragma unroll 1
        for (int p = 0; p < GP; ++p) {
            const int pair = pbase + p;
            if (pair >= pair_hi) break;
            fetch_in(pair + 1, nxt);
            const int cA = 2 * pair, cB = cA + 1;
            const bool actA = s.act[cA - 2 * pair_lo], actB = s.act[cB - 2 * pair_lo];
            if (actA || actB) {
                const cplx* cl = s.cell[buf][p];
                const double scA = s.scale[cA - 2 * pair_lo], scB = s.scale[cB - 2 * pair_lo];
#pragma unroll
                for (int q = 0; q < 2; ++q) {
                    if (!have[q]) continue;
                    const cplx* cq = cl + o0[q];
                    double r0s = 0.0, r1s = 0.0;
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        double r0 = 0.0, r1 = 0.0;
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            const cplx v = cq[k * W + j];
                            r0 = fma(wy[q][j], v.x, r0);
                            r1 = fma(wy[q][j], v.y, r1);
                        }
                        r0s = fma(wx[q][k], r0, r0s);
                        r1s = fma(wx[q][k], r1, r1s);
                    }
                    if (a.noise) {
                        r0s = fma(nz, cur[q][0] * scA, r0s);
                        r1s = fma(nz, cur[q][1] * scB, r1s);
                    }
                    if (actA) a.out[(long)cA * a.ldo + so[q]] = r0s;
                    if (actB) a.out[(long)cB * a.ldo + so[q]] = r1s;
                }
            }
#pragma unroll
            for (int q = 0; q < 2; ++q) { cur[q][0] = nxt[q][0]; cur[q][1] = nxt[q][1]; }
        }
        __syncthreads();   // buffer `buf` is free for the pass after next
    }
}

// ---------------------------------------------------------------------------
// The same gather writing a point-major block (out[point][column], caller's rows) itself.  Threads own
// points as above, so the 2 GP results of a point and pass sit in ONE thread while a row piece wants them
// side by side in GP lanes: an 8 x 8 transpose across each group of 8 lanes (three butterfly stages of
// __shfl_xor) turns "lane = point, registers = pairs" into "lane = pair, registers = 8 points", after which
// the 8 lanes of a group store 128 contiguous bytes of one row -- and read the 128 bytes of the input row
// next to them for the noise term.  No transposing pass, no scratch copy of the result.
// ---------------------------------------------------------------------------
template <int GP, int TB>
__global__ void __launch_bounds__(256, 2) from_grid_2d_rows_kernel(const InterpArgs a, int passes_per_cta) {
    static_assert(GP == 8, "one pair per lane of an 8-lane group");
    typedef Gather3Smem<GP, TB> Smem;
    constexpr int W = Smem::W;
    extern __shared__ __align__(16) unsigned char smem_raw5[];
    Smem& s = *reinterpret_cast<Smem*>(smem_raw5);
    const int tid = threadIdx.x;
    const int tile = blockIdx.x % a.tiles;
    const int d = blockIdx.x / a.tiles;
    const int mx = a.m0, my = a.m1;
    const int BX0 = (tile / a.tiles1) * TB, BY0 = (tile % a.tiles1) * TB;
    const int X0 = BX0 - 3, Y0 = BY0 - 3;
    const int* bs = a.bin_start + (long)d * a.NB;
    const int npairs_tot = (a.ncols + 1) >> 1;
    const int pair_lo = blockIdx.y * passes_per_cta * GP;
    const int pair_hi = min(npairs_tot, pair_lo + passes_per_cta * GP);
    if (tid < TB) {
        const int bx = BX0 + tid;
        int beg = 0, end = 0;
        if (bx < a.nb0) {
            beg = bs[(long)bx * a.nb1 + BY0];
            end = bs[(long)bx * a.nb1 + min(BY0 + TB, a.nb1)];
        }
        s.row_beg[tid] = beg;
        s.row_off[tid + 1] = end - beg;
    }
    __syncthreads();
    if (tid == 0) {
        s.row_off[0] = 0;
        for (int r = 0; r < TB; ++r) s.row_off[r + 1] += s.row_off[r];
    }
    __syncthreads();
    const int npts = s.row_off[TB];
    if (npts == 0) return;

    auto stage = [&](int pbase, int buf) {
        for (int p = 0; p < GP; ++p) {
            const int pair = pbase + p;
            if (pair >= pair_hi) break;
            const cplx* gsl = a.Gc + ((long)pair * a.D + d) * a.grid_pitch;
            for (int c = tid; c < W * W; c += 256) {
                const int x = c / W, y = c - x * W;
                const int gx = clampi(X0 + x, 0, mx - 1), gy = clampi(Y0 + y, 0, my - 1);
                cp_async16(&s.cell[buf][p][c], gsl + (long)gx * my + gy);
            }
        }
        cp_async_commit();
    };
    stage(pair_lo, 0);

    double wx[2][4], wy[2][4];
    int o0[2], so[2];
    bool have[2];
#pragma unroll
    for (int q = 0; q < 2; ++q) {
        const int li = tid + 256 * q;
        have[q] = li < npts;
        so[q] = 0;
        if (have[q]) {
            int r = 0;
            while (li >= s.row_off[r + 1]) ++r;
            const long gi = s.row_beg[r] + (li - s.row_off[r]);
            keys_weights(a.u0[gi], wx[q]);
            keys_weights(a.u1[gi], wy[q]);
            const int ix0 = a.i00[gi] - 1, iy0 = a.i01[gi] - 1;
            o0[q] = (ix0 - X0) * W + (iy0 - Y0);
            so[q] = a.perm_out ? a.perm_out[gi] : (int)gi;
        } else {
#pragma unroll
            for (int k = 0; k < 4; ++k) wx[q][k] = wy[q][k] = 0.0;
            o0[q] = 0;
        }
    }
    const double nz = a.noise ? a.noise[d] : 0.0;
    const int lane = tid & 31, gb = lane & ~7, g = lane & 7;
    const double* __restrict__ xin = a.in_rows;
    double* __restrict__ yout = a.out;
    const bool no_prefetch = a.extra != 0;             // A/B switch (LMC_NO_ROWPREFETCH)

    int buf = 0;
    for (int pbase = pair_lo; pbase < pair_hi; pbase += GP, buf ^= 1) {
        if (pbase + GP < pair_hi) { stage(pbase + GP, buf ^ 1); cp_async_wait_1(); }
        else cp_async_wait_all();
        __syncthreads();
        const int np = min(GP, pair_hi - pbase);       // uniform over the CTA
#pragma unroll
        for (int q = 0; q < 2; ++q) {
            const unsigned hv = __ballot_sync(0xffffffffu, have[q]);
            if (hv == 0) continue;                     // uniform over the warp
            if (a.noise && !no_prefetch && np > 1)
                prefetch_row_pieces8(so[q], hv, gb, g, 2 * (pbase + g), pbase + g < pair_hi, xin, a.ldr);
            double acc[GP][2];
#pragma unroll
            for (int p = 0; p < GP; ++p) {
                double r0s = 0.0, r1s = 0.0;
                if (p < np) {
                    const cplx* cq = s.cell[buf][p] + o0[q];
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        double r0 = 0.0, r1 = 0.0;
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            const cplx v = cq[k * W + j];
                            r0 = fma(wy[q][j], v.x, r0);
                            r1 = fma(wy[q][j], v.y, r1);
                        }
                        r0s = fma(wx[q][k], r0, r0s);
                        r1s = fma(wx[q][k], r1, r1s);
                    }
                }
                acc[p][0] = r0s;
                acc[p][1] = r1s;
            }
            if (np == 1) {
                // a pass with a single pair (the odd last pair of 8 k + 1): every thread stores its own point's
                // one or two values, no transpose (uniform over the CTA)
                if (have[q]) {
                    const int cA1 = 2 * pbase;
                    const bool okB1 = cA1 + 1 < a.ncols;
                    const long row = so[q];
                    double v0 = acc[0][0], v1 = acc[0][1];
                    if (a.noise) {
                        v0 = fma(nz, __ldcs(xin + row * a.ldr + cA1), v0);
                        if (okB1) v1 = fma(nz, __ldcs(xin + row * a.ldr + cA1 + 1), v1);
                    }
                    __stcs(yout + row * a.ldo + cA1, v0);
                    if (okB1) __stcs(yout + row * a.ldo + cA1 + 1, v1);
                }
                continue;
            }
            lane_group_transpose8(acc, g);     // acc[i] now belongs to point (gb + i), pair pbase + g
            const int pair = pbase + g;
            const int cA = 2 * pair;
            const bool okA = pair < pair_hi, okB = okA && cA + 1 < a.ncols;
            store_row_pieces8(acc, so[q], hv, gb, cA, okA, okB, a.noise != nullptr, nz, xin, a.ldr, yout, a.ldo);
        }
        __syncthreads();   // buffer `buf` is free for the pass after next
    }
}

// out[c][i] = in[c][perm[i]]: with perm = sorted -> caller it brings a block of caller-ordered columns
// into the operator's sorted point order ahead of the (coalesced) sorted-order kernels, with the
// inverse permutation it takes results back.  Writes are coalesced; the source column (8 n bytes)
// stays in L2 while it is gathered from.
template <int CPT, bool CG>
__global__ void __launch_bounds__(256) permute_cols_kernel(const double* __restrict__ in, long ld,
                                                           const int* __restrict__ perm, long n, int ncols,
                                                           double* __restrict__ out, long ldo) {
    // CPT columns per thread share one load of the indices; CG: gathers bypass L1 (every 8-byte element is
    // its own 32-byte sector, no reuse to cache)
    const int c0 = blockIdx.y * CPT;
    const long base = (long)blockIdx.x * 2048 + threadIdx.x;
    int idx[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) idx[k] = (base + 256 * k < n) ? perm[base + 256 * k] : 0;
#pragma unroll
    for (int j = 0; j < CPT; ++j) {
        const int c = c0 + j;
        if (c >= ncols) break;
        double v[8];
#pragma unroll
        for (int k = 0; k < 8; ++k)
            v[k] = CG ? __ldcg(in + (long)c * ld + idx[k]) : in[(long)c * ld + idx[k]];
#pragma unroll
        for (int k = 0; k < 8; ++k)
            if (base + 256 * k < n) out[(long)c * ldo + base + 256 * k] = v[k];
    }
}
int permute_cols(const PointSet& ps, bool to_sorted, const double* in, long ld, int ncols, double* out,
                 long ldo, cudaStream_t st) {
    if (ncols == 0) return 0;
    ProfScope prof(PROF_OTHER, st);
    // 4 columns per thread (one load of the indices), gathers through L2 only: 1.230 -> 1.204 ms at config E.
    // The pass stays bound by L2 -> SM sector traffic: 32 bytes fetched per 8 gathered.
    const dim3 grid((unsigned)ceil_div(ps.n, 2048), (unsigned)ceil_div(ncols, 4));
    permute_cols_kernel<4, true><<<grid, 256, 0, st>>>(in, ld, to_sorted ? ps.perm : ps.iperm, ps.n, ncols, out, ldo);
    count_launch();
    LMC_CHECK(cudaGetLastError());
    return 0;
}

static int env_int(const char* name, int dflt) {
    const char* v = getenv(name);
    return v ? atoi(v) : dflt;
}

// ---------------------------------------------------------------------------
// Point-major blocks (rows[i][c], numpy's C order for an [n, P] array, caller's point order) <-> the
// operator's sorted column-major layout.  A CTA owns 32 consecutive sorted points and up to kRowChunk
// columns: whole row pieces on the point-major side, 256-byte column pieces on the other, transposed
// through shared memory -- both sides move full sectors, unlike a permutation of column-major data
// (one 32-byte sector per 8-byte element).
// ---------------------------------------------------------------------------
static const int kRowChunk = 136, kRowPitch = kRowChunk + 1;

// cols[c][j] = rows[perm[j]][c]
__global__ void __launch_bounds__(256) rows_to_sorted_cols_kernel(const double* __restrict__ rows, long ldr,
                                                                  const int* __restrict__ perm, long n, int ncols,
                                                                  double* __restrict__ cols, long ldc) {
    __shared__ double tile[32 * kRowPitch];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const long j0 = (long)blockIdx.x * 32;
    const int c0 = blockIdx.y * kRowChunk;
    const int cw = min(kRowChunk, ncols - c0);
    for (int p = warp; p < 32; p += 8) {
        const long j = j0 + p;
        if (j >= n) break;
        const double* row = rows + (long)(perm ? perm[j] : j) * ldr + c0;
        for (int c = lane; c < cw; c += 32) tile[p * kRowPitch + c] = __ldcs(row + c);
    }
    __syncthreads();
    if (j0 + lane < n)
        for (int c = warp; c < cw; c += 8) cols[(long)(c0 + c) * ldc + j0 + lane] = tile[lane * kRowPitch + c];
}

// rows_out[perm[j]][c] = cols[c][j] + noise[d(j)] * rows_in[perm[j]][c]   (the D v term of K~ v, same
// fused multiply-add as the gather kernels' epilogue).  All loads of a phase are issued before their first
// use: the pass is pure data movement (3 x 8 n P bytes) and lives off the bytes it keeps in flight.
__global__ void __launch_bounds__(256, 4) sorted_cols_to_rows_kernel(const double* __restrict__ cols, long ldc,
                                                                  const int* __restrict__ perm, long n, int ncols,
                                                                  const double* __restrict__ noise,
                                                                  const long* __restrict__ out_start, int D,
                                                                  const double* __restrict__ rows_in, long ldi,
                                                                  double* __restrict__ rows_out, long ldo) {
    __shared__ double tile[32 * kRowPitch];
    constexpr int CI = (kRowChunk + 31) / 32;    // columns per lane on the row side
    constexpr int CJ = (kRowChunk + 7) / 8;      // columns per warp on the column side
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const long j0 = (long)blockIdx.x * 32;
    const int c0 = blockIdx.y * kRowChunk;
    const int cw = min(kRowChunk, ncols - c0);
    // this warp's 4 points: row index and noise, one lane each, handed round by shuffle
    long my_row = 0;
    double my_nz = 0.0;
    {
        const long j = j0 + 4 * warp + (lane & 3);
        if (j < n) {
            my_row = perm ? perm[j] : j;
            if (noise) {
                int d = 0;
                while (d + 1 < D && j >= out_start[d + 1]) ++d;
                my_nz = noise[d];
            }
        }
    }
    {
        double v[CJ];
        const bool have = j0 + lane < n;
#pragma unroll
        for (int k = 0; k < CJ; ++k) {
            const int c = warp + 8 * k;
            v[k] = (have && c < cw) ? __ldcs(cols + (long)(c0 + c) * ldc + j0 + lane) : 0.0;
        }
#pragma unroll
        for (int k = 0; k < CJ; ++k) {
            const int c = warp + 8 * k;
            if (c < cw) tile[lane * kRowPitch + c] = v[k];
        }
    }
    // the input rows of the noise term: in flight across the barrier
    double x[4][CI];
#pragma unroll
    for (int p = 0; p < 4; ++p) {
        const long r = __shfl_sync(0xffffffffu, my_row, p);
#pragma unroll
        for (int k = 0; k < CI; ++k) {
            const int c = lane + 32 * k;
            x[p][k] = (noise && c < cw && j0 + 4 * warp + p < n) ? __ldcs(rows_in + r * ldi + c0 + c) : 0.0;
        }
    }
    __syncthreads();
#pragma unroll
    for (int p = 0; p < 4; ++p) {
        const long r = __shfl_sync(0xffffffffu, my_row, p);
        const double nz = __shfl_sync(0xffffffffu, my_nz, p);
        if (j0 + 4 * warp + p >= n) continue;
        double* yout = rows_out + r * ldo + c0;
#pragma unroll
        for (int k = 0; k < CI; ++k) {
            const int c = lane + 32 * k;
            if (c < cw) {
                double v = tile[(4 * warp + p) * kRowPitch + c];
                if (noise) v = fma(nz, x[p][k], v);
                __stcs(yout + c, v);
            }
        }
    }
}

int rows_to_sorted_cols(const PointSet& ps, const double* rows, long ldr, int ncols, double* cols, long ldc,
                        cudaStream_t st) {
    if (ncols == 0) return 0;
    ProfScope prof(PROF_OTHER, st);
    const dim3 grid((unsigned)ceil_div(ps.n, 32), (unsigned)ceil_div(ncols, kRowChunk));
    rows_to_sorted_cols_kernel<<<grid, 256, 0, st>>>(rows, ldr, ps.identity ? nullptr : ps.perm, ps.n, ncols, cols, ldc);
    count_launch();
    LMC_CHECK(cudaGetLastError());
    return 0;
}

int sorted_cols_to_rows(const PointSet& ps, const double* cols, long ldc, int ncols, const double* noise,
                        const double* rows_in, long ldi, double* rows_out, long ldo, cudaStream_t st) {
    if (ncols == 0) return 0;
    ProfScope prof(PROF_OTHER, st);
    const dim3 grid((unsigned)ceil_div(ps.n, 32), (unsigned)ceil_div(ncols, kRowChunk));
    sorted_cols_to_rows_kernel<<<grid, 256, 0, st>>>(cols, ldc, ps.identity ? nullptr : ps.perm, ps.n, ncols, noise,
                                                     ps.out_start_dev, ps.D, rows_in, ldi, rows_out, ldo);
    count_launch();
    LMC_CHECK(cudaGetLastError());
    return 0;
}

// ---------------------------------------------------------------------------
// host launchers
// ---------------------------------------------------------------------------
static InterpArgs make_args(const PointSet& ps, const ColumnView& cv) {
    InterpArgs a = {};
    a.u0 = ps.u[0]; a.u1 = ps.u[1];
    a.i00 = ps.i0[0]; a.i01 = ps.i0[1];
    a.perm_in = (cv.sorted_in || ps.identity) ? nullptr : ps.perm;
    a.perm_out = (cv.sorted_out || ps.identity) ? nullptr : ps.perm;
    a.bin_start = ps.bin_start;
    a.out_start = ps.out_start_dev;
    a.in = cv.in; a.out = cv.out; a.ld = cv.ld; a.ldo = cv.ld_out ? cv.ld_out : cv.ld;
    if (cv.rows_in) { a.in_rows = cv.in; a.ldr = cv.ld; a.in = nullptr; }
    if (cv.rows_out) a.perm_out = ps.identity ? nullptr : ps.perm;
    a.ncols = cv.ncols;
    a.in_scale = cv.in_scale; a.active = cv.active;
    a.grid_pitch = ps.grid_pitch;
    a.D = ps.D; a.m0 = ps.m[0]; a.m1 = ps.m[1];
    a.NB = ps.NB; a.nb0 = ps.nb[0]; a.nb1 = ps.nb[1];
    return a;
}

template <class K>
static int set_smem(K kernel, size_t bytes) {
    LMC_CHECK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
    return 0;
}

static const int kCap16 = 304, kCap8 = 448;   // staged points per CTA: 8x8 tile (16 pair lanes), 16x8 tile (8 lanes)

// which 2-D scatter kernel a segment of RHS pairs goes through
enum ScatterKind { SCATTER_AUTO, SCATTER_G16, SCATTER_G8, SCATTER_PAIR, SCATTER_G16X, SCATTER_G8X };

static int to_grid_launch(const PointSet& ps, const ColumnView& cv, cplx* G, ScatterKind kind, cudaStream_t st);

template <int G, int TX, int TY, int CAP>
static int launch_strips(const PointSet& ps, InterpArgs a, int npairs, cudaStream_t st) {
    typedef Tile3Smem<G, TX, TY, CAP> Smem;
    static bool attr3 = false;
    if (!attr3) {
        LMC_TRY(set_smem(to_grid_2d_v3_kernel<G, TX, TY, CAP, false>, sizeof(Smem)));
        LMC_TRY(set_smem(to_grid_2d_v3_kernel<G, TX, TY, CAP, true>, sizeof(Smem)));
        attr3 = true;
    }
    a.tiles1 = ceil_div(ps.m[1], TY);
    a.tiles = ceil_div(ps.m[0], TX) * a.tiles1;
    if (a.extra) npairs -= 1;                     // the odd last column rides with the last full group
    const int ngroups = ceil_div(npairs, G);
    const long ctas1 = (long)a.tiles * ps.D;
    // all groups in one CTA (weights computed once) unless that leaves the machine underfilled
    int gpc = std::min(ngroups, 64);
    while (gpc > 1 && ctas1 * ceil_div(ngroups, gpc) < 148L * 2 * 4) gpc = (gpc + 1) / 2;
    dim3 grid3((unsigned)ctas1, (unsigned)ceil_div(ngroups, gpc));
    if (a.in_rows) to_grid_2d_v3_kernel<G, TX, TY, CAP, true><<<grid3, 256, sizeof(Smem), st>>>(a, gpc);
    else to_grid_2d_v3_kernel<G, TX, TY, CAP, false><<<grid3, 256, sizeof(Smem), st>>>(a, gpc);
    return 0;
}

struct ScatterSeg { int first, count; ScatterKind kind; };

// 2-D: the strip kernel's lanes are RHS pairs, 16 per group.  A block whose pair count is not a
// multiple of 16 (y plus an even number of probes; the 8 or 9 pairs a rank holds when 128 probes
// are sharded over 8 GPUs) would idle the lanes of its last group, so the remainder is split off:
// up to 8 pairs go through the 8-lane variant (16x8-cell tiles), one or two left-over pairs through
// the one-pair-per-CTA kernel.  Returns the number of segments, 0 when the strip kernels do not apply.
static int scatter_plan(const PointSet& ps, int ncols, ScatterSeg (&segs)[3]) {
    static const int variant = env_int("LMC_TOGRID2D", 3);
    if (ps.ndim != 2 || variant < 3 || ps.max_tile_pts_8x8 > kCap16) return 0;
    const int npairs = (ncols + 1) / 2;
    const bool g8_ok = ps.max_tile_pts_16x8 <= kCap8;
    int nseg = 0;
    int done = npairs / 16 * 16, rem = npairs - done;
    static const bool no_extra = getenv("LMC_NO_EXTRACOL") != nullptr;
    if ((ncols & 1) && !no_extra && (rem == 1 || (rem == 9 && g8_ok)) && npairs > 1) {
        // 16 k (+ 8) pairs and one odd column: the column rides along as a third accumulator of the last
        // group (129 columns on one GPU, the 17 of a rank when 128 probes are sharded over 8)
        if (rem == 1) {
            segs[nseg++] = {0, npairs, SCATTER_G16X};
        } else {
            if (done) segs[nseg++] = {0, done, SCATTER_G16};
            segs[nseg++] = {done, rem, SCATTER_G8X};
        }
        done = npairs; rem = 0;
    } else if (done) {
        segs[nseg++] = {0, done, SCATTER_G16};
    }
    if (rem >= 3 && rem <= 10 && g8_ok) {
        // 9 or 10 pairs: the second 8-lane group (1 or 2 live lanes) reuses the CTA's staged weights,
        // which is cheaper than a separate launch of the one-pair kernel
        segs[nseg++] = {done, rem, SCATTER_G8};
        done += rem; rem = 0;
    }
    if (rem >= 1 && rem <= 2 && done > 0) segs[nseg++] = {done, rem, SCATTER_PAIR};
    else if (rem) segs[nseg++] = {done, rem, SCATTER_G16};
    return nseg;
}

bool to_grid_takes_rows(const PointSet& ps, int ncols) {
    static const bool off = getenv("LMC_NO_ROWS_SCATTER") != nullptr;
    if (off || ps.identity) return false;
    if (ps.ndim == 1) return true;
    ScatterSeg segs[3];
    const int nseg = scatter_plan(ps, ncols, segs);
    for (int i = 0; i < nseg; ++i)
        if (segs[i].kind == SCATTER_PAIR) return false;
    return nseg > 0;
}

int to_grid(const PointSet& ps, const ColumnView& cv, cplx* G, cudaStream_t st) {
    if (cv.ncols == 0) return 0;
    ProfScope prof(PROF_TO_GRID, st);
    ScatterSeg segs[3];
    const int nseg = scatter_plan(ps, cv.ncols, segs);
    if (cv.rows_in) LMC_REQUIRE(to_grid_takes_rows(ps, cv.ncols), "point-major block not supported by this scatter");
    if (nseg == 0) return to_grid_launch(ps, cv, G, SCATTER_AUTO, st);
    for (int i = 0; i < nseg; ++i) {
        ColumnView part = cv;
        const int c0 = 2 * segs[i].first;
        part.ncols = std::min(2 * segs[i].count, cv.ncols - c0);
        part.in = cv.rows_in ? cv.in + c0 : cv.in + (long)c0 * cv.ld;
        part.in_scale = cv.in_scale ? cv.in_scale + c0 : nullptr;
        part.active = cv.active ? cv.active + c0 : nullptr;
        LMC_TRY(to_grid_launch(ps, part, G + (size_t)segs[i].first * ps.D * ps.grid_pitch, segs[i].kind, st));
    }
    return 0;
}

static int to_grid_launch(const PointSet& ps, const ColumnView& cv, cplx* G, ScatterKind kind, cudaStream_t st) {
    InterpArgs a = make_args(ps, cv);
    a.G = G;
    const int npairs = (cv.ncols + 1) / 2;
    if (ps.ndim == 1) {
        constexpr int G1 = 8, CAP1 = 512, CAP1S = 256;
        typedef Scatter1Smem<G1, CAP1> Smem;
        typedef Scatter1Smem<G1, CAP1S> SmemS;
        static bool attr = false;
        static const int cap_sel = env_int("LMC_TG1_CAP", CAP1S);   // 256: four CTAs per SM (0.125 -> 0.101 ms at config D)
        if (!attr) {
            LMC_TRY(set_smem(to_grid_1d_v3_kernel<G1, CAP1, false>, sizeof(Smem)));
            LMC_TRY(set_smem(to_grid_1d_v3_kernel<G1, CAP1S, false>, sizeof(SmemS)));
            LMC_TRY(set_smem(to_grid_1d_v3_kernel<G1, CAP1, true>, sizeof(Smem)));
            LMC_TRY(set_smem(to_grid_1d_v3_kernel<G1, CAP1S, true>, sizeof(SmemS)));
            attr = true;
        }
        const int TC = 256 / G1 - 3;
        a.tiles = ceil_div(ps.m[0], TC);
        dim3 grid((unsigned)(a.tiles * ps.D), (unsigned)ceil_div(npairs, G1));
        if (a.in_rows) {
            const dim3 grid_r(grid.y, grid.x);      // group index fastest
            if (cap_sel == CAP1S) to_grid_1d_v3_kernel<G1, CAP1S, true><<<grid_r, 256, sizeof(SmemS), st>>>(a);
            else to_grid_1d_v3_kernel<G1, CAP1, true><<<grid_r, 256, sizeof(Smem), st>>>(a);
        } else {
            if (cap_sel == CAP1S) to_grid_1d_v3_kernel<G1, CAP1S, false><<<grid, 256, sizeof(SmemS), st>>>(a);
            else to_grid_1d_v3_kernel<G1, CAP1, false><<<grid, 256, sizeof(Smem), st>>>(a);
        }
    } else if (kind == SCATTER_G16 || kind == SCATTER_G16X) {
        a.extra = kind == SCATTER_G16X;
        LMC_TRY((launch_strips<16, 8, 8, kCap16>(ps, a, npairs, st)));
    } else if (kind == SCATTER_G8 || kind == SCATTER_G8X) {
        a.extra = kind == SCATTER_G8X;
        LMC_TRY((launch_strips<8, 16, 8, kCap8>(ps, a, npairs, st)));
    } else {
        a.tiles1 = ceil_div(ps.m[1], kTY);
        a.tiles = ceil_div(ps.m[0], kTX) * a.tiles1;
        dim3 grid((unsigned)(a.tiles * ps.D), (unsigned)npairs);
        const size_t smem = sizeof(double) * 32 * kBX * kBY;
        static bool attr = false;
        if (!attr) { LMC_TRY(set_smem(to_grid_2d_kernel, smem)); attr = true; }
        to_grid_2d_kernel<<<grid, 384, smem, st>>>(a);
    }
    count_launch();
    LMC_CHECK(cudaGetLastError());
    return 0;
}

bool from_grid_writes_rows(const PointSet& ps) {
    static const bool off = getenv("LMC_NO_ROWS_GATHER") != nullptr;
    static const int variant = env_int("LMC_FROMGRID2D", 3);
    if (off) return false;
    return ps.ndim == 1 || (ps.ndim == 2 && variant >= 3 && ps.max_gather_tile_pts <= 512);
}

int from_grid(const PointSet& ps, const ColumnView& cv, const cplx* G, const double* noise,
              cudaStream_t st) {
    if (cv.ncols == 0) return 0;
    if (cv.rows_out) LMC_REQUIRE(from_grid_writes_rows(ps) && cv.rows_in, "point-major output not supported by this gather");
    InterpArgs a = make_args(ps, cv);
    a.Gc = G;
    ProfScope prof(PROF_FROM_GRID, st);
    a.noise = noise;
    long maxlen = 0;
    for (int d = 0; d < ps.D; ++d) maxlen = std::max(maxlen, ps.out_start[d + 1] - ps.out_start[d]);
    if (maxlen == 0) return 0;
    const int npairs = (cv.ncols + 1) / 2;
    if (ps.ndim == 1) {
        const long ctas1 = (long)ceil_div(maxlen, 256) * ps.D;
        int ppc = npairs;   // all pairs per thread (weights once) unless that underfills the machine
        while (ppc > 1 && ctas1 * ceil_div(npairs, ppc) < 148L * 8) ppc = (ppc + 1) / 2;
        if (cv.rows_out) ppc = ceil_div(ppc, 8) * 8;      // whole groups of 8 pairs per CTA
        dim3 grid((unsigned)ceil_div(maxlen, 256), (unsigned)ps.D, (unsigned)ceil_div(npairs, ppc));
        if (cv.rows_out) from_grid_1d_rows_kernel<<<grid, 256, 0, st>>>(a, ppc);
        else from_grid_1d_v3_kernel<<<grid, 256, 0, st>>>(a, ppc);
    } else {
        static const int variant = env_int("LMC_FROMGRID2D", 3);
        constexpr int GP = 8, TB = 16;
        if (cv.rows_out) {
            typedef Gather3Smem<GP, TB> Smem;
            static bool attr_r = false;
            if (!attr_r) { LMC_TRY(set_smem(from_grid_2d_rows_kernel<GP, TB>, sizeof(Smem))); attr_r = true; }
            a.tiles1 = ceil_div(ps.nb[1], TB);
            a.tiles = ceil_div(ps.nb[0], TB) * a.tiles1;
            const long ctas1 = (long)a.tiles * ps.D;
            const int npass = ceil_div(npairs, GP);
            int ppc = npass;
            while (ppc > 1 && ctas1 * ceil_div(npass, ppc) < 148L * 2 * 4) ppc = (ppc + 1) / 2;
            dim3 grid((unsigned)ctas1, (unsigned)ceil_div(npass, ppc));
            static const int no_pf = env_int("LMC_NO_ROWPREFETCH", 0);
            a.extra = no_pf;
            from_grid_2d_rows_kernel<GP, TB><<<grid, 256, sizeof(Smem), st>>>(a, ppc);
        } else if (variant >= 3 && ps.max_gather_tile_pts <= 512) {
            typedef Gather3Smem<GP, TB> Smem;
            static bool attr = false;
            if (!attr) { LMC_TRY(set_smem(from_grid_2d_v3_kernel<GP, TB>, sizeof(Smem))); attr = true; }
            a.tiles1 = ceil_div(ps.nb[1], TB);
            a.tiles = ceil_div(ps.nb[0], TB) * a.tiles1;
            const long ctas1 = (long)a.tiles * ps.D;
            const int npass = ceil_div(npairs, GP);
            int ppc = std::min(npass, 272 / (2 * GP));
            while (ppc > 1 && ctas1 * ceil_div(npass, ppc) < 148L * 2 * 4) ppc = (ppc + 1) / 2;
            dim3 grid((unsigned)ctas1, (unsigned)ceil_div(npass, ppc));
            from_grid_2d_v3_kernel<GP, TB><<<grid, 256, sizeof(Smem), st>>>(a, ppc);
        } else {
            const long ctas1 = (long)ceil_div(maxlen, 128) * ps.D;
            int ppc = (int)std::max<long>(1, std::min<long>(8, (ctas1 * npairs) / (148L * 16 * 4)));
            ppc = std::min(ppc, npairs);
            dim3 grid((unsigned)ceil_div(maxlen, 128), (unsigned)ps.D, (unsigned)ceil_div(npairs, ppc));
            from_grid_2d_v2_kernel<<<grid, 128, 0, st>>>(a, ppc);
        }
    }
    count_launch();
    LMC_CHECK(cudaGetLastError());
    return 0;
}

}  // namespace lmc
