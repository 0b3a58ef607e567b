// Register-resident 512-point column transform + coregionalisation mix + inverse for the 2-D fused path
// (grids with 128 < m_x <= 256, i.e. embedding 512: config E).  Replaces, for those geometries, the
// shared-memory radix-8 pipeline of fused_lines_kernel: numpy rfftn/irfftn inside BTTB.matvec (reference
// runlmc/linalg/bttb.py:144-148) and the Kronecker/SumMatrix glue (kronecker.py:39-46,
// sum_matrix.py:31-32) between them.
//
// One warp owns one line (512 complex points of one output, one frequency row, one RHS pair), 16 points
// per lane, all butterfly indices compile-time constants:
//
//   forward   lane m holds x[m + 32 a], a < 16 (coalesced loads; a >= 8 is the zero padding)
//             DFT16 over a in registers, twiddle W512^{m c}                                  -> y_c[m]
//             exchange through shared memory: lane (c, h) takes y_c[h + 2 j], j < 16
//             DFT16 over j in registers, twiddle W32^{k2} on the odd lanes                   -> Z_h[k2]
//             radix-2 butterfly ACROSS the lane pair with __shfl_xor: lane h = 0 keeps
//             X[c + 16 k2], lane h = 1 keeps -X[c + 16 (k2 + 16)]  (one DFMA per component)
//   mix       all D lines of the CTA meet in shared memory; thread <-> frequency bin, the sign of the
//             odd lanes passes through the (linear) mix unchanged
//   inverse   exact mirror, the last DFT16 only produces the 8 outputs below m_x.
//
// Position p = k2 * 32 + lane of the mix layout holds frequency k = c + 16 (k2 + 16 h), c = lane >> 1,
// h = lane & 1; the spectra are stored in that order once per parameter update (col512_spec_kernel).
// Per element and direction the data crosses shared memory twice (4 round trips per product against 6
// for the radix-8 pipeline) and no index arithmetic is left in the butterflies.
#pragma once
#include "common.cuh"
#include "fft.cuh"

namespace lmc {

static const int kC512Line = 16 * 34;     // complex elements of shared memory per line (pitch 34: conflict free)
static const int kC512Tw = 15 * 32;       // W512^{m c}, c = 1..15, m < 32

// a * W16^e (forward, exp(-2 pi i e / 16)) or its conjugate (inverse); e is a compile-time constant
// after unrolling
template <bool INV>
__device__ __forceinline__ cplx mul_w16(cplx a, int e) {
    const double C1 = 0.92387953251128673848, S1 = 0.38268343236508978178;
    double wr, wi;   // forward twiddle (wr, -wi)
    switch (e) {
        case 0: return a;
        case 2: return rot45<INV>(a);
        case 4: return rot90<INV>(a);
        case 6: return rot135<INV>(a);
        case 1: wr = C1; wi = S1; break;
        case 3: wr = S1; wi = C1; break;
        default: wr = -C1; wi = -S1; break;   // e == 9
    }
    if (INV) wi = -wi;
    return make_double2(fma(a.y, wi, a.x * wr), fma(a.y, wr, -(a.x * wi)));
}

// a * W32^k (forward) or its conjugate, k = 0..15 compile-time
template <bool INV>
__device__ __forceinline__ cplx mul_w32(cplx a, int k) {
    double wr, wi;
    switch (k) {
        case 0: return a;
        case 4: return rot45<INV>(a);
        case 8: return rot90<INV>(a);
        case 12: return rot135<INV>(a);
        case 1: wr = 9.80785280403230430579e-01; wi = 1.95090322016128248084e-01; break;
        case 2: wr = 9.23879532511286738483e-01; wi = 3.82683432365089781779e-01; break;
        case 3: wr = 8.31469612302545235671e-01; wi = 5.55570233019602177649e-01; break;
        case 5: wr = 5.55570233019602288671e-01; wi = 8.31469612302545235671e-01; break;
        case 6: wr = 3.82683432365089837290e-01; wi = 9.23879532511286738483e-01; break;
        case 7: wr = 1.95090322016128331351e-01; wi = 9.80785280403230430579e-01; break;
        case 9: wr = -1.95090322016128192573e-01; wi = 9.80785280403230430579e-01; break;
        case 10: wr = -3.82683432365089726268e-01; wi = 9.23879532511286738483e-01; break;
        case 11: wr = -5.55570233019601955604e-01; wi = 8.31469612302545457716e-01; break;
        case 13: wr = -8.31469612302545346694e-01; wi = 5.55570233019602177649e-01; break;
        case 14: wr = -9.23879532511286738483e-01; wi = 3.82683432365089892802e-01; break;
        default: wr = -9.80785280403230430579e-01; wi = 1.95090322016128608906e-01; break;   // 15
    }
    if (INV) wi = -wi;
    return make_double2(fma(a.y, wi, a.x * wr), fma(a.y, wr, -(a.x * wi)));
}

// 16-point DFT in registers, natural order in and out (radix 4 x 4, decimation in frequency).
// HIN: inputs 8..15 are zero and not read.  HOUT: only outputs 0..7 are produced.  INV: conjugate
// transform, unscaled.
template <bool INV, bool HIN, bool HOUT>
__device__ __forceinline__ void dft16(cplx* x) {
    cplx t[16];   // t[4 n0 + q] = sum_n1 x[n0 + 4 n1] W4^{n1 q}
#pragma unroll
    for (int n0 = 0; n0 < 4; ++n0) {
        if (HIN) {
            const cplx a = x[n0], b = x[n0 + 4], sb = rot90<INV>(b);
            t[4 * n0 + 0] = cadd(a, b);
            t[4 * n0 + 1] = cadd(a, sb);
            t[4 * n0 + 2] = csub(a, b);
            t[4 * n0 + 3] = csub(a, sb);
        } else {
            const cplx a = x[n0], b = x[n0 + 4], c = x[n0 + 8], d = x[n0 + 12];
            const cplx ac = cadd(a, c), amc = csub(a, c), bd = cadd(b, d), bmd = rot90<INV>(csub(b, d));
            t[4 * n0 + 0] = cadd(ac, bd);
            t[4 * n0 + 1] = cadd(amc, bmd);
            t[4 * n0 + 2] = csub(ac, bd);
            t[4 * n0 + 3] = csub(amc, bmd);
        }
    }
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        const cplx a = t[q], b = mul_w16<INV>(t[4 + q], q), c = mul_w16<INV>(t[8 + q], 2 * q),
                   d = mul_w16<INV>(t[12 + q], 3 * q);
        const cplx ac = cadd(a, c), amc = csub(a, c), bd = cadd(b, d), bmd = rot90<INV>(csub(b, d));
        x[q] = cadd(ac, bd);
        x[q + 4] = cadd(amc, bmd);
        if (!HOUT) {
            x[q + 8] = csub(ac, bd);
            x[q + 12] = csub(amc, bmd);
        }
    }
}

// x[c] *= w^c (INV: conj(w)^c), c = 1..15, with the powers built on the fly from w = W512^m (the lane's own
// root, kept in two registers for the whole kernel): two interleaved chains (odd / even exponents, step w^2),
// so no power is more than 8 multiplications deep.  Trades 15 shared-memory loads per transform side for 15
// complex multiplications: the kernel is bound by the LSU data pipe (80 % busy), not by the fp64 pipe (45 %).
template <bool INV>
__device__ __forceinline__ void twiddle_powers(cplx* x, cplx w) {
    const cplx w2 = cmul(w, w);
    cplx po = w, pe = w2;      // w^1, w^2
#pragma unroll
    for (int c = 1; c < 16; c += 2) {
        x[c] = INV ? cmulc(x[c], po) : cmul(x[c], po);
        if (c + 1 < 16) x[c + 1] = INV ? cmulc(x[c + 1], pe) : cmul(x[c + 1], pe);
        if (c + 2 < 16) { po = cmul(po, w2); pe = cmul(pe, w2); }
    }
}

struct Col512Args {
    cplx* data;               // S_T[pair][d][line][x]
    long slab_stride;         // elements per (pair, output) slab
    long line_stride;         // elements between consecutive lines (x pitch)
    int n_lines, valid;       // frequency rows, m_x (<= 256)
    int npairs, ppc;          // RHS pairs, pairs per CTA
    int Q;
    const double* specP;      // [Q][n_lines][512] spectra in the mix layout
    const cplx* tw1;          // [32]  W512^m (the first row of the [15][32] table W512^{m c})
};

template <int D, class MIX, int MINB>
__global__ void __launch_bounds__(32 * D, MINB) fused_col512_kernel(const Col512Args a, const MIX mb) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    cplx* bufs = reinterpret_cast<cplx*>(smem_raw);   // [D][kC512Line]
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int line = blockIdx.x;
    const int c2 = lane >> 1, h = lane & 1;
    const double sgn = h ? -1.0 : 1.0;
    const cplx w1 = a.tw1[lane];                  // W512^lane
    cplx* buf = bufs + warp * kC512Line;
    cplx* bx = buf + c2 * 34 + h;                 // second-stage view: y_c[h + 2 j] at bx[2 j]
    for (int pp = 0; pp < a.ppc; ++pp) {
        const long pair = (long)blockIdx.y * a.ppc + pp;
        if (pair >= a.npairs) break;              // uniform over the CTA
        cplx* g = a.data + (pair * D + warp) * a.slab_stride + (long)line * a.line_stride + lane;
        if (pp + 1 < a.ppc && pair + 1 < a.npairs && 8 * lane < a.valid) {
            // the next pair's line of this warp: start it on its way from HBM to L2 now (128 B per lane)
            const cplx* nx = g - lane + (long)D * a.slab_stride + 8 * lane;
            asm volatile("prefetch.global.L2 [%0];" ::"l"(nx));
        }
        cplx x[16];
        // ---- forward ----
#pragma unroll
        for (int r = 0; r < 8; ++r) x[r] = (lane + 32 * r < a.valid) ? g[32 * r] : make_double2(0.0, 0.0);
        dft16<false, true, false>(x);
        twiddle_powers<false>(x, w1);
#pragma unroll
        for (int c = 0; c < 16; ++c) buf[c * 34 + lane] = x[c];
        __syncwarp();
#pragma unroll
        for (int j = 0; j < 16; ++j) x[j] = bx[2 * j];
        __syncwarp();
        dft16<false, false, false>(x);
        if (h) {
#pragma unroll
            for (int k = 1; k < 16; ++k) x[k] = mul_w32<false>(x[k], k);
        }
#pragma unroll
        for (int k = 0; k < 16; ++k) {            // radix-2 butterfly across the lane pair
            const double rx = __shfl_xor_sync(0xffffffffu, x[k].x, 1);
            const double ry = __shfl_xor_sync(0xffffffffu, x[k].y, 1);
            x[k].x = fma(sgn, rx, x[k].x);
            x[k].y = fma(sgn, ry, x[k].y);
        }
#pragma unroll
        for (int k = 0; k < 16; ++k) buf[k * 32 + lane] = x[k];
        __syncthreads();
        // ---- mix: y[dp] = sum_d (sum_q f_q B_q[dp][d]) x[d] at every bin ----
        for (int p = threadIdx.x; p < 512; p += 32 * D) {
            cplx v[D];
#pragma unroll
            for (int d = 0; d < D; ++d) v[d] = bufs[d * kC512Line + p];
            double f[MIX::NQ];
            const double* sp = a.specP + (long)line * 512 + p;
            const long qstride = (long)a.n_lines * 512;
#pragma unroll
            for (int q = 0; q < MIX::NQ; ++q) f[q] = (MIX::NQ <= 4 || q < a.Q) ? __ldg(sp + q * qstride) : 0.0;
            mb.apply_store(f, a.Q, v, bufs + p, kC512Line);
        }
        __syncthreads();
        // ---- inverse ----
#pragma unroll
        for (int k = 0; k < 16; ++k) x[k] = buf[k * 32 + lane];
        __syncwarp();
#pragma unroll
        for (int k = 0; k < 16; ++k) {
            const double rx = __shfl_xor_sync(0xffffffffu, x[k].x, 1);
            const double ry = __shfl_xor_sync(0xffffffffu, x[k].y, 1);
            x[k].x = fma(-sgn, rx, x[k].x);
            x[k].y = fma(-sgn, ry, x[k].y);
        }
        if (h) {
#pragma unroll
            for (int k = 1; k < 16; ++k) x[k] = mul_w32<true>(x[k], k);
        }
        dft16<true, false, false>(x);
#pragma unroll
        for (int j = 0; j < 16; ++j) bx[2 * j] = x[j];
        __syncwarp();
#pragma unroll
        for (int c = 0; c < 16; ++c) x[c] = buf[c * 34 + lane];
        twiddle_powers<true>(x, w1);
        dft16<true, false, true>(x);
#pragma unroll
        for (int r = 0; r < 8; ++r)
            if (lane + 32 * r < a.valid) g[32 * r] = x[r];
        __syncwarp();
    }
}

// specP[q][line][p] = specL[q][line][old position of frequency k(p)], p = k2 * 32 + lane  (file header);
// specL holds the spectrum in the digit-reversed order of the radix-8 plan: frequency q0 + 8 q1 + 64 q2
// sits at position 64 q0 + 8 q1 + q2.
static __global__ void col512_spec_kernel(const double* __restrict__ specL, double* specP, long total) {
    const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    const int p = (int)(i & 511);
    const int lane = p & 31, k2 = p >> 5;
    const int k = (lane >> 1) + 16 * (k2 + 16 * (lane & 1));
    const int old = (k & 7) * 64 + ((k >> 3) & 7) * 8 + (k >> 6);
    specP[i] = specL[i - p + old];
}

static __global__ void col512_twiddle_kernel(cplx* tw1) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= kC512Tw) return;
    const int c = i / 32 + 1, m = i % 32;
    double s, co;
    sincospi(-2.0 * (double)(m * c) / 512.0, &s, &co);
    tw1[i] = make_double2(co, s);
}

}  // namespace lmc
