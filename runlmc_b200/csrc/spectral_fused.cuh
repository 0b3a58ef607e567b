// Fused column-transform + coregionalisation-mix + inverse kernel and its launcher.
// Included by spectral.cu (declarations, extern templates) and by spectral_fused_*.cu, which hold the
// explicit instantiations for four values of D each so the build compiles them in parallel.
#pragma once
#include "spectral.cuh"
#include "fft.cuh"
#include "spectral_col512.cuh"

#include <algorithm>
#include <cstdlib>

namespace lmc {

static const size_t kFusedSmemMax = 200 * 1024;

// Forward transform of the D lines of one RHS pair, per-bin coregionalisation mix, inverse
// transform -- all in shared memory, in place on global memory.  The mixing matrices travel as
// kernel parameters so they are constant-bank operands of the DFMAs.
// Mixing operators, passed as kernel parameters (constant-bank operands).
// Dense:    y = (sum_q f_q B_q) x                      (Q + 2) D^2 DFMA per bin and RHS pair
// Low rank: B_q = A_q^T A_q + diag(kappa_q) (the LMC parameterisation, reference
//           functional_kernel.py:280-287):  y = sum_q f_q A_q^T (A_q x) + (sum_q f_q kappa_q) .* x
//           sum_q R_q (4D + 2) + (Q + 2) D DFMA
template <int D>
struct MixB {
    static const int NQ = 8;
    double b[8][D][D];
    // in place on the D values of one bin
    __device__ __forceinline__ void apply(const double* f, int Q, cplx* x) const {
        cplx y[D];
#pragma unroll
        for (int dp = 0; dp < D; ++dp) {
            double m[D];
#pragma unroll
            for (int d = 0; d < D; ++d) m[d] = 0.0;
#pragma unroll
            for (int q = 0; q < 8; ++q) {
                if (q < Q) {   // warp-uniform
#pragma unroll
                    for (int d = 0; d < D; ++d) m[d] = fma(f[q], b[q][dp][d], m[d]);
                }
            }
            double yr = 0.0, yi = 0.0;
#pragma unroll
            for (int d = 0; d < D; ++d) {
                yr = fma(m[d], x[d].x, yr);
                yi = fma(m[d], x[d].y, yi);
            }
            y[dp] = make_double2(yr, yi);
        }
#pragma unroll
        for (int d = 0; d < D; ++d) x[d] = y[d];
    }
    // Same product with every output written to dst[dp * stride] as soon as it is complete (x is in
    // registers, so dst may be where x came from): no second register array.
    __device__ __forceinline__ void apply_store(const double* f, int Q, const cplx* x, cplx* dst, int stride) const {
#pragma unroll
        for (int dp = 0; dp < D; ++dp) {
            double yr = 0.0, yi = 0.0;
#pragma unroll
            for (int d = 0; d < D; ++d) {
                double m = 0.0;
#pragma unroll
                for (int q = 0; q < 8; ++q)
                    if (q < Q) m = fma(f[q], b[q][dp][d], m);   // warp-uniform
                yr = fma(m, x[d].x, yr);
                yi = fma(m, x[d].y, yi);
            }
            dst[dp * stride] = make_double2(yr, yi);
        }
    }
};

static const int kMaxRankPerKernel = 2;   // low-rank path: every B_q has rank <= 2 (+ diagonal)
// NQ = number of kernels the struct holds: Q itself for Q <= 4 (loops fully unrolled, no per-q
// branches, Q spectrum loads per bin), or 8 with run-time Q.  RK = largest rank among the kernels
// (compile time; lower-rank kernels are zero padded), so no rank test is left in the inner loops.
template <int D, int NQ_, int RK_>
struct MixLR {
    static const int NQ = NQ_;
    static const int RK = RK_;
    double a[NQ][RK][D];
    double kappa[NQ][D];
    // In place.  The mix is a real matrix, so the real and the imaginary parts of the bin are two
    // independent real products: they are done one after the other with one set of D accumulators,
    // which keeps the live registers at ~3D doubles (x complex + y) instead of 5D.
    __device__ __forceinline__ void apply(const double* f, int Q, cplx* x) const {
#pragma unroll
        for (int part = 0; part < 2; ++part) {
            double y[D];
#pragma unroll
            for (int d = 0; d < D; ++d) y[d] = 0.0;
#pragma unroll
            for (int q = 0; q < NQ; ++q) {
                if (NQ <= 4 || q < Q) {   // warp-uniform
#pragma unroll
                    for (int r = 0; r < RK; ++r) {
                        double t = 0.0;
#pragma unroll
                        for (int d = 0; d < D; ++d) t = fma(a[q][r][d], part ? x[d].y : x[d].x, t);
                        t *= f[q];
#pragma unroll
                        for (int d = 0; d < D; ++d) y[d] = fma(a[q][r][d], t, y[d]);
                    }
                }
            }
            // diagonal part (sum_q f_q kappa_q[d]) x[d], recomputed per part rather than held in D
            // more registers
#pragma unroll
            for (int d = 0; d < D; ++d) {
                double ks = 0.0;
#pragma unroll
                for (int q = 0; q < NQ; ++q)
                    if (NQ <= 4 || q < Q) ks = fma(f[q], kappa[q][d], ks);
                if (part) x[d].y = fma(ks, x[d].y, y[d]);
                else x[d].x = fma(ks, x[d].x, y[d]);
            }
        }
    }
    // Both parts at once without an accumulator array: first the projections t_{q,r} = f_q (a_{q,r} . x)
    // (complex, 2 Q RK doubles), then every output y_d = (sum_q f_q kappa_q[d]) x_d + sum a_{q,r}[d] t_{q,r}
    // is written to dst[d * stride] as soon as it is complete (x is in registers, so dst may be where x came
    // from).  Live values: x (2 D doubles) + t; Q RK (4 D + 2) + (Q + 2) D fp64 operations per bin.
    __device__ __forceinline__ void apply_store(const double* f, int Q, const cplx* x, cplx* dst, int stride) const {
        cplx t[NQ][RK];
#pragma unroll
        for (int q = 0; q < NQ; ++q) {
#pragma unroll
            for (int r = 0; r < RK; ++r) {
                double tr = 0.0, ti = 0.0;
                if (NQ <= 4 || q < Q) {   // warp-uniform
#pragma unroll
                    for (int d = 0; d < D; ++d) {
                        tr = fma(a[q][r][d], x[d].x, tr);
                        ti = fma(a[q][r][d], x[d].y, ti);
                    }
                    tr *= f[q];
                    ti *= f[q];
                }
                t[q][r] = make_double2(tr, ti);
            }
        }
#pragma unroll
        for (int d = 0; d < D; ++d) {
            double ks = 0.0;
#pragma unroll
            for (int q = 0; q < NQ; ++q)
                if (NQ <= 4 || q < Q) ks = fma(f[q], kappa[q][d], ks);
            double yr = ks * x[d].x, yi = ks * x[d].y;
#pragma unroll
            for (int q = 0; q < NQ; ++q) {
                if (NQ <= 4 || q < Q) {
#pragma unroll
                    for (int r = 0; r < RK; ++r) {
                        yr = fma(a[q][r][d], t[q][r].x, yr);
                        yi = fma(a[q][r][d], t[q][r].y, yi);
                    }
                }
            }
            dst[d * stride] = make_double2(yr, yi);
        }
    }
};

struct FusedArgs {
    cplx* data;
    long slab_stride, line_stride;
    int n_lines, L, valid, lpc;   // lines per slab, line length, valid prefix, line-sets per CTA
    int Q;
    const double* specL;          // [Q][n_lines][L]
    const cplx* stage_tw;         // per-stage twiddle tables (global), layout `lay`
    int tw_total;
    StageTw lay;
    FftPlan plan;
    int pitch, half;
    const double* specP;          // spectra in the col512 mix layout (or null)
    const cplx* tw512;            // col512 first-stage twiddles
    int force_col512;             // the row passes already committed to the register transform's order
    const cplx* tw_plain;         // exp(-2 pi i k / tw_n), k < tw_n, L | tw_n (block-cooperative variant)
    int tw_n;
    int block_variant;            // few long lines: one CTA per line set instead of one warp per line
};

template <int D, class MIX>
__global__ void __launch_bounds__(320, 2) fused_lines_kernel(const FusedArgs a, const MIX mb) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    cplx* tws = reinterpret_cast<cplx*>(smem_raw);            // per-stage twiddle tables
    cplx* tile = tws + a.tw_total;                            // [lpc*D][pitch]
    const int L = a.L, pitch = a.pitch, lpc = a.lpc;
    const int lL = 31 - __clz(L);
    const int line0 = blockIdx.x * lpc;
    const long pair = blockIdx.y;
    cplx* base = a.data + pair * D * a.slab_stride;
    const int nl = lpc * D;
    const int warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
    for (int i = threadIdx.x; i < a.tw_total; i += blockDim.x) tws[i] = a.stage_tw[i];
    __syncthreads();
    for (int li = warp; li < nl; li += nwarps) {
        const int ll = li / D, d = li - ll * D;
        if (line0 + ll >= a.n_lines) continue;
        const cplx* g = base + d * a.slab_stride + (long)(line0 + ll) * a.line_stride;
        warp_fft_forward(g, a.valid, tile + li * pitch, L, a.plan, a.lay, tws, a.half != 0);
    }
    __syncthreads();
    // ---- mix:  y[dp] = sum_d (sum_q f_q B_q[dp][d]) x[d]  at every bin of every line-set ----
    for (int w = threadIdx.x; w < lpc * L; w += blockDim.x) {
        const int ll = w >> lL, p = w & (L - 1);
        if (line0 + ll >= a.n_lines) continue;
        cplx* col = tile + (ll * D) * pitch + pad_idx(p);
        cplx x[D];
#pragma unroll
        for (int d = 0; d < D; ++d) x[d] = col[d * pitch];
        double f[MIX::NQ];
        const double* sp = a.specL + (long)(line0 + ll) * L + p;
        const long qstride = (long)a.n_lines * L;
#pragma unroll
        for (int q = 0; q < MIX::NQ; ++q)
            f[q] = (MIX::NQ <= 4 || q < a.Q) ? __ldg(sp + q * qstride) : 0.0;
        mb.apply(f, a.Q, x);
#pragma unroll
        for (int d = 0; d < D; ++d) col[d * pitch] = x[d];
    }
    __syncthreads();
    for (int li = warp; li < nl; li += nwarps) {
        const int ll = li / D, d = li - ll * D;
        if (line0 + ll >= a.n_lines) continue;
        cplx* g = base + d * a.slab_stride + (long)(line0 + ll) * a.line_stride;
        warp_fft_inverse(tile + li * pitch, L, a.plan, a.lay, tws, a.half != 0, g, a.valid);
    }
}


// Few, long lines (the launch-bound configs A / B / C: 16-120 lines in the whole launch): one warp per line
// leaves the GPU with a few dozen warps on a chain of dependent shared-memory stages.  This variant gives every
// line set (the D lines of one frequency row / 1-D grid of one RHS pair) a whole CTA: block-wide radix stages
// (fft_tile_forward / fft_tile_inverse, fft.cuh), same plan and therefore the same frequency order as the
// warp-per-line kernel, same mix.
template <int D, class MIX>
__global__ void __launch_bounds__(256) fused_lines_block_kernel(const FusedArgs a, const MIX mb) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    cplx* tile = reinterpret_cast<cplx*>(smem_raw);           // [D][pitch]
    const int L = a.L, pitch = a.pitch;
    const int line = blockIdx.x;
    const long pair = blockIdx.y;
    cplx* base = a.data + pair * D * a.slab_stride + (long)line * a.line_stride;
    const int cnt = a.half ? (L >> 1) : L;                    // the upper half is known zero / not wanted
    for (int idx = threadIdx.x; idx < cnt * D; idx += blockDim.x) {
        const int d = idx / cnt, e = idx - d * cnt;
        tile[d * pitch + pad_idx(e)] = e < a.valid ? base[d * a.slab_stride + e] : make_double2(0.0, 0.0);
    }
    __syncthreads();
    fft_tile_forward(tile, pitch, D, L, a.plan, a.half != 0, a.tw_plain, a.tw_n);
    for (int p = threadIdx.x; p < L; p += blockDim.x) {
        cplx* col = tile + pad_idx(p);
        cplx x[D];
#pragma unroll
        for (int d = 0; d < D; ++d) x[d] = col[d * pitch];
        double f[MIX::NQ];
        const double* sp = a.specL + (long)line * L + p;
        const long qstride = (long)a.n_lines * L;
#pragma unroll
        for (int q = 0; q < MIX::NQ; ++q) f[q] = (MIX::NQ <= 4 || q < a.Q) ? __ldg(sp + q * qstride) : 0.0;
        mb.apply(f, a.Q, x);
#pragma unroll
        for (int d = 0; d < D; ++d) col[d * pitch] = x[d];
    }
    __syncthreads();
    fft_tile_inverse(tile, pitch, D, L, a.plan, a.half != 0, a.tw_plain, a.tw_n);
    for (int idx = threadIdx.x; idx < cnt * D; idx += blockDim.x) {
        const int d = idx / cnt, e = idx - d * cnt;
        if (e < a.valid) base[d * a.slab_stride + e] = tile[d * pitch + pad_idx(e)];
    }
}

// the register/shuffle column kernel for 512-point pruned lines (spectral_col512.cuh)
static inline bool use_col512(const FusedArgs& a) {
    static const bool off = getenv("LMC_NO_COL512") != nullptr;
    return (a.force_col512 || !off) && a.L == 512 && a.half && a.specP && a.tw512 && a.valid <= 256;
}

template <int D, class MIX>
static int launch_col512_kernel(const FusedArgs& a, const MIX& m, int npairs, cudaStream_t st) {
    static const int ppc_env = getenv("LMC_COL512_PPC") ? atoi(getenv("LMC_COL512_PPC")) : 8;   // 4 -> 8: 1.435 -> 1.412 ms at config E
    const size_t smem = sizeof(cplx) * (size_t)(D * kC512Line);
    constexpr int MINB = D <= 10 ? 2 : 1;
    static bool attr = false;
    if (!attr) {
        LMC_CHECK(cudaFuncSetAttribute(fused_col512_kernel<D, MIX, MINB>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       (int)kFusedSmemMax));
        attr = true;
    }
    Col512Args c = {};
    c.data = a.data; c.slab_stride = a.slab_stride; c.line_stride = a.line_stride;
    // pairs per CTA: amortises the twiddle table and lets a CTA prefetch its next lines; keep >= 4 waves of CTAs
    c.n_lines = a.n_lines; c.valid = a.valid; c.npairs = npairs;
    c.ppc = std::max(1, std::min(ppc_env, (int)((long)a.n_lines * npairs / (4 * 296))));
    c.Q = a.Q; c.specP = a.specP; c.tw1 = a.tw512;
    dim3 grid((unsigned)a.n_lines, (unsigned)ceil_div(npairs, c.ppc));
    fused_col512_kernel<D, MIX, MINB><<<grid, 32 * D, smem, st>>>(c, m);
    return 0;
}

template <int D, class MIX>
static int launch_fused_kernel(const FusedArgs& a, const MIX& m, dim3 grid, int threads, size_t smem,
                               cudaStream_t st) {
    if (use_col512(a)) return launch_col512_kernel<D, MIX>(a, m, (int)grid.y, st);
    if (a.block_variant) {
        static bool attr_b = false;
        if (!attr_b) {
            LMC_CHECK(cudaFuncSetAttribute(fused_lines_block_kernel<D, MIX>,
                                           cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kFusedSmemMax));
            attr_b = true;
        }
        const size_t smem_b = (size_t)D * a.pitch * sizeof(cplx);
        fused_lines_block_kernel<D, MIX><<<dim3((unsigned)a.n_lines, grid.y), 256, smem_b, st>>>(a, m);
        return 0;
    }
    static bool attr = false;
    if (!attr) {
        LMC_CHECK(cudaFuncSetAttribute(fused_lines_kernel<D, MIX>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       (int)kFusedSmemMax));
        attr = true;
    }
    fused_lines_kernel<D, MIX><<<grid, threads, smem, st>>>(a, m);
    return 0;
}

template <int D, int NQ, int RK>
static int launch_fused_lowrank(const FusedArgs& a, const MixSpec& mix, dim3 grid, int threads, size_t smem,
                                cudaStream_t st) {
    MixLR<D, NQ, RK> ml = {};     // kernel parameter, copied at launch; kernels of lower rank stay zero padded
    int r = 0;
    for (int q = 0; q < a.Q; ++q) {
        for (int k = 0; k < mix.ranks[q]; ++k, ++r)
            for (int d = 0; d < D; ++d) ml.a[q][k][d] = mix.A[(size_t)r * D + d];
        for (int d = 0; d < D; ++d) ml.kappa[q][d] = mix.kappa[(size_t)q * D + d];
    }
    return launch_fused_kernel<D>(a, ml, grid, threads, smem, st);
}

template <int D, int RK>
static int launch_fused_lowrank_q(const FusedArgs& a, const MixSpec& mix, dim3 grid, int threads, size_t smem,
                                  cudaStream_t st) {
    switch (a.Q) {
        case 1: return launch_fused_lowrank<D, 1, RK>(a, mix, grid, threads, smem, st);
        case 2: return launch_fused_lowrank<D, 2, RK>(a, mix, grid, threads, smem, st);
        case 3: return launch_fused_lowrank<D, 3, RK>(a, mix, grid, threads, smem, st);
        case 4: return launch_fused_lowrank<D, 4, RK>(a, mix, grid, threads, smem, st);
        default: return launch_fused_lowrank<D, 8, RK>(a, mix, grid, threads, smem, st);
    }
}

template <int D>
int launch_fused_lines(FusedArgs a, const cplx* stage_tw, const MixSpec& mix, int npairs,
                       cudaStream_t st) {
    int total_rank = 0, max_rank = 0;
    if (mix.ranks)
        for (int q = 0; q < a.Q; ++q) {
            total_rank += mix.ranks[q];
            max_rank = std::max(max_rank, mix.ranks[q]);
        }
    // cost of the zero-padded low-rank form against the dense one
    const bool lowrank = mix.ranks && max_rank <= kMaxRankPerKernel &&
                         a.Q * std::max(max_rank, 1) * (4 * D + 2) + (a.Q + 2) * D < (a.Q + 2) * D * D;
    a.plan = make_plan(a.L);
    a.lay = stage_tw_layout(a.L, a.plan);
    a.tw_total = a.lay.total;
    a.stage_tw = stage_tw;
    a.pitch = line_pitch(a.L);
    a.half = (a.L >= 2 && a.valid <= a.L / 2) ? 1 : 0;
    const size_t per_set = (size_t)D * a.pitch * sizeof(cplx);
    static const bool no_block = getenv("LMC_NO_BLOCK_LINES") != nullptr;
    // measured (B200, MINRES iteration): L = 2048 (config B) 0.078 -> 0.067 ms; L = 512 with 13 outputs (config C)
    // and L = 256 (config A) gain nothing, so only long lines take it
    a.block_variant = (!no_block && a.tw_plain && a.L >= 1024 && per_set <= kFusedSmemMax &&
                       (long)a.n_lines * npairs * D <= 148L * 4) ? 1 : 0;
    int lpc = (int)std::max<size_t>(1, std::min<size_t>(48 * 1024 / per_set, (size_t)(4096 / (D * a.L) + 1)));
    lpc = std::max(1, std::min(lpc, a.n_lines));
    a.lpc = lpc;
    const size_t smem = per_set * lpc + sizeof(cplx) * (size_t)a.tw_total;
    LMC_REQUIRE(smem <= kFusedSmemMax, "fused spectral tile does not fit shared memory");
    const int threads = 32 * std::max(2, std::min(10, lpc * D));   // one warp per line, up to 10 warps
    dim3 grid((unsigned)ceil_div(a.n_lines, lpc), (unsigned)npairs);
    ProfScope prof(PROF_MIX, st);
    int rc;
    (void)total_rank;
    if (lowrank) {
        rc = max_rank <= 1 ? launch_fused_lowrank_q<D, 1>(a, mix, grid, threads, smem, st)
                           : launch_fused_lowrank_q<D, 2>(a, mix, grid, threads, smem, st);
    } else {
        MixB<D> mb = {};     // kernel parameter; only the first Q blocks are read
        for (int q = 0; q < a.Q; ++q)
            for (int i = 0; i < D; ++i)
                for (int j = 0; j < D; ++j) mb.b[q][i][j] = mix.B[((size_t)q * D + i) * D + j];
        rc = launch_fused_kernel<D>(a, mb, grid, threads, smem, st);
    }
    LMC_TRY(rc);
    count_launch();
    LMC_CHECK(cudaGetLastError());
    return 0;
}


#define LMC_FUSED_EXTERN(DD) \
    extern template int launch_fused_lines<DD>(FusedArgs, const cplx*, const MixSpec&, int, cudaStream_t);
#define LMC_FUSED_INSTANTIATE(DD) \
    template int launch_fused_lines<DD>(FusedArgs, const cplx*, const MixSpec&, int, cudaStream_t);

}  // namespace lmc
