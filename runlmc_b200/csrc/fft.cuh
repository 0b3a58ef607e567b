// Hand-written fp64 complex FFT building blocks for sm_100a.
//
// Convolution-only design: the forward transform is an in-place
// decimation-in-frequency (natural order in, digit-reversed order out) and the
// inverse is its exact mirror (decimation-in-time, digit-reversed in, natural
// out), so no reordering pass is ever needed: spectra of the circulant
// embeddings are produced by the same forward kernels and therefore live in the
// same digit-reversed layout as the data they multiply.
//
// Radix plan for a line of length L = 2^k: an optional leading radix-2/4 stage
// (largest span) followed by radix-8 stages.  With the element padding
// pad_idx(i) = i + (i >> 3) every stage's 16-byte shared-memory accesses are
// bank-conflict free (spans are multiples of 8, or 1).
#pragma once
#include "common.cuh"

namespace lmc {

struct FftPlan {
    int nst;
    int radix[8];
};

static inline FftPlan make_plan(int L) {
    FftPlan p;
    p.nst = 0;
    int k = ilog2((unsigned)L);
    int r = k % 3;
    if (r) p.radix[p.nst++] = 1 << r;
    for (int i = 0; i < k / 3; ++i) p.radix[p.nst++] = 8;
    return p;
}

__host__ __device__ __forceinline__ int pad_idx(int i) { return i + (i >> 3); }
// pitch (in elements) of one padded line inside a shared-memory tile; the +1
// skews consecutive lines onto different 16-byte bank groups.
static inline int line_pitch(int L) { return L + (L >> 3) + 1; }

// position in the DIF output -> logical frequency index
__device__ __forceinline__ int digit_reverse(int p, int L, const FftPlan& pl) {
    int k = 0, mult = 1, rem = L;
    for (int s = 0; s < pl.nst; ++s) {
        rem /= pl.radix[s];
        int q = p / rem;
        p -= q * rem;
        k += q * mult;
        mult *= pl.radix[s];
    }
    return k;
}

#define LMC_SQRT1_2 0.70710678118654752440

// multiply by -i (forward) / +i (inverse)
template <bool INV>
__device__ __forceinline__ cplx rot90(cplx a) {
    return INV ? make_double2(-a.y, a.x) : make_double2(a.y, -a.x);
}
// multiply by W8^1 = (1-i)/sqrt2 (forward) / conj (inverse)
template <bool INV>
__device__ __forceinline__ cplx rot45(cplx a) {
    return INV ? make_double2((a.x - a.y) * LMC_SQRT1_2, (a.x + a.y) * LMC_SQRT1_2)
               : make_double2((a.x + a.y) * LMC_SQRT1_2, (a.y - a.x) * LMC_SQRT1_2);
}
// multiply by W8^3 = (-1-i)/sqrt2 (forward) / conj (inverse)
template <bool INV>
__device__ __forceinline__ cplx rot135(cplx a) {
    return INV ? make_double2((-a.x - a.y) * LMC_SQRT1_2, (a.x - a.y) * LMC_SQRT1_2)
               : make_double2((a.y - a.x) * LMC_SQRT1_2, (-a.x - a.y) * LMC_SQRT1_2);
}

// In-register R-point DFTs.  HIN: inputs r >= R/2 are known zero (pruned
// forward first stage).  HOUT: only outputs r < R/2 are needed (pruned
// inverse last stage).
template <int R, bool INV, bool HIN, bool HOUT>
struct Dft;

template <bool INV, bool HIN, bool HOUT>
struct Dft<2, INV, HIN, HOUT> {
    static __device__ __forceinline__ void run(cplx* x) {
        if (HIN) {
            x[1] = x[0];
        } else {
            cplx a = x[0], b = x[1];
            x[0] = cadd(a, b);
            if (!HOUT) x[1] = csub(a, b);
        }
    }
};

template <bool INV, bool HIN, bool HOUT>
struct Dft<4, INV, HIN, HOUT> {
    static __device__ __forceinline__ void run(cplx* x) {
        cplx a0, a1, a2, a3;
        if (HIN) {
            a0 = x[0]; a1 = x[0]; a2 = x[1]; a3 = rot90<INV>(x[1]);
        } else {
            a0 = cadd(x[0], x[2]); a1 = csub(x[0], x[2]);
            a2 = cadd(x[1], x[3]); a3 = rot90<INV>(csub(x[1], x[3]));
        }
        x[0] = cadd(a0, a2);
        x[1] = cadd(a1, a3);
        if (!HOUT) {
            x[2] = csub(a0, a2);
            x[3] = csub(a1, a3);
        }
    }
};

template <bool INV, bool HIN, bool HOUT>
struct Dft<8, INV, HIN, HOUT> {
    static __device__ __forceinline__ void run(cplx* x) {
        cplx a0, a1, a2, a3, a4, a5, a6, a7;
        if (HIN) {
            a0 = x[0]; a1 = x[0]; a2 = x[2]; a3 = rot90<INV>(x[2]);
            a4 = x[1]; a5 = x[1]; a6 = x[3]; a7 = rot90<INV>(x[3]);
        } else {
            a0 = cadd(x[0], x[4]); a1 = csub(x[0], x[4]);
            a2 = cadd(x[2], x[6]); a3 = rot90<INV>(csub(x[2], x[6]));
            a4 = cadd(x[1], x[5]); a5 = csub(x[1], x[5]);
            a6 = cadd(x[3], x[7]); a7 = rot90<INV>(csub(x[3], x[7]));
        }
        cplx b0 = cadd(a0, a2), b2 = csub(a0, a2);
        cplx b1 = cadd(a1, a3), b3 = csub(a1, a3);
        cplx b4 = cadd(a4, a6), b6 = rot90<INV>(csub(a4, a6));
        cplx b5 = rot45<INV>(cadd(a5, a7)), b7 = rot135<INV>(csub(a5, a7));
        x[0] = cadd(b0, b4);
        x[1] = cadd(b1, b5);
        x[2] = cadd(b2, b6);
        x[3] = cadd(b3, b7);
        if (!HOUT) {
            x[4] = csub(b0, b4);
            x[5] = csub(b1, b5);
            x[6] = csub(b2, b6);
            x[7] = csub(b3, b7);
        }
    }
};

// One radix-R stage over a shared-memory tile of `nl` padded lines of length L.
// Forward (DIF):  X_q = DFT_R(x)_q * W_Ns^{o q}, in place.
// Inverse (DIT):  x_r = IDFT_R(y_q * conj(W_Ns^{o q}))_r, in place (unscaled).
// `tw` holds exp(-2 pi i k / tw_n), k < tw_n, with Ns | tw_n.
template <int R, bool INV, bool HIN, bool HOUT>
__device__ __forceinline__ void fft_stage(cplx* tile, int pitch, int nl, int L, int Ns,
                                          const cplx* __restrict__ tw, int tw_n) {
    const int span = Ns / R;
    const int lspan = 31 - __clz(span);
    const int per_line = L / R;
    const int lper = 31 - __clz(per_line);
    const int total = per_line * nl;
    const int twmul = tw_n / Ns;
    for (int w = threadIdx.x; w < total; w += blockDim.x) {
        const int line = w >> lper;
        const int b = w & (per_line - 1);
        const int blk = b >> lspan;
        const int o = b & (span - 1);
        cplx* a = tile + line * pitch;
        const int base = blk * Ns + o;
        cplx x[R];
        if (!INV) {
#pragma unroll
            for (int r = 0; r < R; ++r)
                if (!HIN || r < R / 2) x[r] = a[pad_idx(base + r * span)];
            Dft<R, false, HIN, false>::run(x);
            if (span > 1) {
#pragma unroll
                for (int q = 1; q < R; ++q) x[q] = cmul(x[q], __ldg(&tw[o * q * twmul]));
            }
#pragma unroll
            for (int q = 0; q < R; ++q) a[pad_idx(base + q * span)] = x[q];
        } else {
#pragma unroll
            for (int q = 0; q < R; ++q) x[q] = a[pad_idx(base + q * span)];
            if (span > 1) {
#pragma unroll
                for (int q = 1; q < R; ++q) x[q] = cmulc(x[q], __ldg(&tw[o * q * twmul]));
            }
            Dft<R, true, false, HOUT>::run(x);
#pragma unroll
            for (int r = 0; r < R; ++r)
                if (!HOUT || r < R / 2) a[pad_idx(base + r * span)] = x[r];
        }
    }
}

template <bool INV, bool HIN, bool HOUT>
__device__ __forceinline__ void fft_stage_dispatch(int R, cplx* tile, int pitch, int nl, int L,
                                                   int Ns, const cplx* __restrict__ tw, int tw_n) {
    if (R == 8) fft_stage<8, INV, HIN, HOUT>(tile, pitch, nl, L, Ns, tw, tw_n);
    else if (R == 4) fft_stage<4, INV, HIN, HOUT>(tile, pitch, nl, L, Ns, tw, tw_n);
    else fft_stage<2, INV, HIN, HOUT>(tile, pitch, nl, L, Ns, tw, tw_n);
}

// Full forward transform of a tile resident in shared memory.  `half_in`:
// elements >= L/2 of every line are zero and were not written to the tile.
__device__ __forceinline__ void fft_tile_forward(cplx* tile, int pitch, int nl, int L,
                                                 const FftPlan& pl, bool half_in,
                                                 const cplx* __restrict__ tw, int tw_n) {
    int Ns = L;
    for (int s = 0; s < pl.nst; ++s) {
        const int R = pl.radix[s];
        if (s == 0 && half_in) fft_stage_dispatch<false, true, false>(R, tile, pitch, nl, L, Ns, tw, tw_n);
        else fft_stage_dispatch<false, false, false>(R, tile, pitch, nl, L, Ns, tw, tw_n);
        Ns /= R;
        __syncthreads();
    }
}

// Mirror of fft_tile_forward.  `half_out`: only elements < L/2 are produced.
__device__ __forceinline__ void fft_tile_inverse(cplx* tile, int pitch, int nl, int L,
                                                 const FftPlan& pl, bool half_out,
                                                 const cplx* __restrict__ tw, int tw_n) {
    int Ns = 1;
    for (int s = pl.nst - 1; s >= 0; --s) {
        const int R = pl.radix[s];
        Ns *= R;
        if (s == 0 && half_out) fft_stage_dispatch<true, false, true>(R, tile, pitch, nl, L, Ns, tw, tw_n);
        else fft_stage_dispatch<true, false, false>(R, tile, pitch, nl, L, Ns, tw, tw_n);
        __syncthreads();
    }
}

// ---------------------------------------------------------------------------
// Warp-per-line variant.  One warp owns one padded line in shared memory, so the
// radix stages only need __syncwarp() between them and different lines of a CTA
// drift apart freely (no block-wide barrier per stage).  The first forward stage
// reads its inputs straight from global memory into registers and the last
// inverse stage writes its outputs straight to global memory, saving one
// shared-memory round trip each way.  Twiddles come from a per-stage table
// tws[stage_off + (q-1)*span + o] = exp(-2 pi i o q / Ns) staged in shared memory
// (lane-contiguous, conflict free).
// ---------------------------------------------------------------------------
struct StageTw {
    int off[8];   // offset of each stage's block inside the table (stages with span == 1 have none)
    int total;
};

static inline StageTw stage_tw_layout(int L, const FftPlan& pl) {
    StageTw t;
    int Ns = L, off = 0;
    for (int s = 0; s < pl.nst; ++s) {
        const int R = pl.radix[s], span = Ns / R;
        t.off[s] = off;
        if (span > 1) off += (R - 1) * span;
        Ns /= R;
    }
    for (int s = pl.nst; s < 8; ++s) t.off[s] = off;
    t.total = off;
    return t;
}

// LSPAN >= 0: log2 of the butterfly span is a compile-time constant, which turns every padded
// shared-memory index into `pad_idx(base) + constant` and every twiddle index into `o + constant`
// (about 10 integer instructions per radix-8 butterfly instead of ~100).  LSPAN < 0: run-time span.
template <int R, int LSPAN, bool INV, bool HIN, bool HOUT, bool GLOBAL_IO>
__device__ __forceinline__ void warp_stage(cplx* line, int L, int Ns, const cplx* tws,
                                           const cplx* __restrict__ gsrc, cplx* gdst, int valid) {
    constexpr int lR = R == 8 ? 3 : (R == 4 ? 2 : 1);
    const int lspan = LSPAN >= 0 ? LSPAN : (31 - __clz(Ns)) - lR;   // powers of two: shifts only
    const int span = 1 << lspan;
    const int lNs = lspan + lR;
    // padded distance between the R elements of a butterfly: exact because base + r*span never
    // carries across a multiple of 8 differently from base (span % 8 == 0, or span == 1 with
    // base % 8 == 0 and R <= 8); spans 2 and 4 only occur for L < 64 and take the general form
    const bool lin = span >= 8 || (span == 1 && R == 8);
    const int pstep = span >= 8 ? span + (span >> 3) : 1;
    const int nb = L >> lR;
    const int lane = threadIdx.x & 31;
    for (int b = lane; b < nb; b += 32) {
        const int blk = b >> lspan;
        const int o = b & (span - 1);
        const int base = (blk << lNs) + o;
        const int pbase = pad_idx(base);
        cplx x[R];
        if (!INV) {
#pragma unroll
            for (int r = 0; r < R; ++r) {
                if (HIN && r >= R / 2) continue;
                const int e = base + r * span;
                if (GLOBAL_IO) x[r] = (e < valid) ? gsrc[e] : make_double2(0.0, 0.0);
                else x[r] = line[lin ? pbase + r * pstep : pad_idx(e)];
            }
            Dft<R, false, HIN, false>::run(x);
            if (span > 1) {
#pragma unroll
                for (int q = 1; q < R; ++q) x[q] = cmul(x[q], tws[(q - 1) * span + o]);
            }
#pragma unroll
            for (int q = 0; q < R; ++q) line[lin ? pbase + q * pstep : pad_idx(base + q * span)] = x[q];
        } else {
#pragma unroll
            for (int q = 0; q < R; ++q) x[q] = line[lin ? pbase + q * pstep : pad_idx(base + q * span)];
            if (span > 1) {
#pragma unroll
                for (int q = 1; q < R; ++q) x[q] = cmulc(x[q], tws[(q - 1) * span + o]);
            }
            Dft<R, true, false, HOUT>::run(x);
#pragma unroll
            for (int r = 0; r < R; ++r) {
                if (HOUT && r >= R / 2) continue;
                const int e = base + r * span;
                if (GLOBAL_IO) { if (e < valid) gdst[e] = x[r]; }
                else line[lin ? pbase + r * pstep : pad_idx(e)] = x[r];
            }
        }
    }
}

template <bool INV, bool HALF, bool GLOBAL_IO>
__device__ __forceinline__ void warp_stage_dispatch(int R, cplx* line, int L, int Ns, const cplx* tws,
                                                    const cplx* gsrc, cplx* gdst, int valid) {
    if (R == 8) {
        // the spans a radix-8 stage can have in a plan (2|4)? 8 8 ... 8 are 8^k, k = 0..3
        switch (Ns) {
            case 8: warp_stage<8, 0, INV, !INV && HALF, INV && HALF, GLOBAL_IO>(line, L, Ns, tws, gsrc, gdst, valid); break;
            case 64: warp_stage<8, 3, INV, !INV && HALF, INV && HALF, GLOBAL_IO>(line, L, Ns, tws, gsrc, gdst, valid); break;
            case 512: warp_stage<8, 6, INV, !INV && HALF, INV && HALF, GLOBAL_IO>(line, L, Ns, tws, gsrc, gdst, valid); break;
            default: warp_stage<8, 9, INV, !INV && HALF, INV && HALF, GLOBAL_IO>(line, L, Ns, tws, gsrc, gdst, valid); break;
        }
    } else if (R == 4) {
        warp_stage<4, -1, INV, !INV && HALF, INV && HALF, GLOBAL_IO>(line, L, Ns, tws, gsrc, gdst, valid);
    } else {
        warp_stage<2, -1, INV, !INV && HALF, INV && HALF, GLOBAL_IO>(line, L, Ns, tws, gsrc, gdst, valid);
    }
}

// First forward stage split in two halves so a kernel can keep the loads of the NEXT line in flight
// while it transforms the current one: inputs of the (at most NBL) butterflies of a lane in registers.
// Only the pruned form (inputs >= L/2 are zero) is provided: it is the one every circulant embedding
// uses.
template <int R, int NBL>
struct FirstStageRegs {
    cplx x[NBL][R / 2];
};

template <int R, int NBL>
__device__ __forceinline__ void first_stage_load(FirstStageRegs<R, NBL>& f, const cplx* __restrict__ gsrc,
                                                 int valid, int L) {
    constexpr int lR = R == 8 ? 3 : (R == 4 ? 2 : 1);
    const int span = L >> lR;
    const int lane = threadIdx.x & 31;
#pragma unroll
    for (int u = 0; u < NBL; ++u) {
        const int b = lane + 32 * u;
#pragma unroll
        for (int r = 0; r < R / 2; ++r) {
            const int e = b + r * span;
            f.x[u][r] = (b < span && e < valid) ? gsrc[e] : make_double2(0.0, 0.0);
        }
    }
}

template <int R, int NBL>
__device__ __forceinline__ void first_stage_compute(const FirstStageRegs<R, NBL>& f, cplx* line, int L,
                                                    const cplx* tws) {
    constexpr int lR = R == 8 ? 3 : (R == 4 ? 2 : 1);
    const int span = L >> lR;
    const int lane = threadIdx.x & 31;
#pragma unroll
    for (int u = 0; u < NBL; ++u) {
        const int b = lane + 32 * u;
        if (b >= span) continue;
        cplx x[R];
#pragma unroll
        for (int r = 0; r < R / 2; ++r) x[r] = f.x[u][r];
        Dft<R, false, true, false>::run(x);
        if (span > 1) {
#pragma unroll
            for (int q = 1; q < R; ++q) x[q] = cmul(x[q], tws[(q - 1) * span + b]);
        }
#pragma unroll
        for (int q = 0; q < R; ++q) line[pad_idx(b + q * span)] = x[q];
    }
}

// stages 1.. of the forward transform of one line already holding the first stage's output
__device__ __forceinline__ void warp_fft_forward_rest(cplx* line, int L, const FftPlan& pl, const StageTw& lay,
                                                      const cplx* tws) {
    int Ns = L / pl.radix[0];
    __syncwarp();
    for (int s = 1; s < pl.nst; ++s) {
        const int R = pl.radix[s];
        warp_stage_dispatch<false, false, false>(R, line, L, Ns, tws + lay.off[s], nullptr, nullptr, 0);
        Ns /= R;
        __syncwarp();
    }
}

// forward transform of one line by the calling warp: global (valid prefix, rest zero) -> shared
__device__ __forceinline__ void warp_fft_forward(const cplx* gsrc, int valid, cplx* line, int L,
                                                 const FftPlan& pl, const StageTw& lay, const cplx* tws,
                                                 bool half) {
    int Ns = L;
    for (int s = 0; s < pl.nst; ++s) {
        const int R = pl.radix[s];
        const cplx* t = tws + lay.off[s];
        if (s == 0) {
            if (half) warp_stage_dispatch<false, true, true>(R, line, L, Ns, t, gsrc, nullptr, valid);
            else warp_stage_dispatch<false, false, true>(R, line, L, Ns, t, gsrc, nullptr, valid);
        } else {
            warp_stage_dispatch<false, false, false>(R, line, L, Ns, t, nullptr, nullptr, 0);
        }
        Ns /= R;
        __syncwarp();
    }
}

// inverse transform of one line by the calling warp: shared -> global (valid prefix only)
__device__ __forceinline__ void warp_fft_inverse(cplx* line, int L, const FftPlan& pl, const StageTw& lay,
                                                 const cplx* tws, bool half, cplx* gdst, int valid) {
    int Ns = 1;
    for (int s = pl.nst - 1; s >= 0; --s) {
        const int R = pl.radix[s];
        Ns *= R;
        const cplx* t = tws + lay.off[s];
        if (s == 0) {
            if (half) warp_stage_dispatch<true, true, true>(R, line, L, Ns, t, nullptr, gdst, valid);
            else warp_stage_dispatch<true, false, true>(R, line, L, Ns, t, nullptr, gdst, valid);
        } else {
            warp_stage_dispatch<true, false, false>(R, line, L, Ns, t, nullptr, nullptr, 0);
            __syncwarp();
        }
    }
}

}  // namespace lmc
