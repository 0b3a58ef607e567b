// Circulant-embedding spectral engine: batched pruned FFT passes + coregionalisation mix.
#pragma once
#include "common.cuh"
#include "fft.cuh"
#include <vector>

namespace lmc {

// Geometry of one symmetric block-Toeplitz (BTTB) embedding, ndim <= 3.
// sizes m_p, embedding mt_p = 2^ceil(log2(2 m_p)) (reference bttb.py:16-19).
struct Embedding {
    int ndim = 0;
    int m[3] = {1, 1, 1};
    int mt[3] = {1, 1, 1};
    long cells = 0;       // prod m
    long bins = 0;        // prod mt  (full complex spectrum; RHS are processed in complex pairs)
    // 1-D long lines are split four-step style: mt[0] = L1 * L2
    int L1 = 0, L2 = 0;
    long grid_pitch = 0;  // allocated cells per (pair, output) grid slab (>= cells)
};

int embedding_init(Embedding* e, int ndim, const int* sizes);

// Host description of the coregionalisation mix: dense B [Q][D][D], optionally with the factors of
// B_q = A_q^T A_q + diag(kappa_q)  (ranks[Q], A [sum ranks][D], kappa [Q][D]).
struct MixSpec {
    const double* B = nullptr;
    const int* ranks = nullptr;
    const double* A = nullptr;
    const double* kappa = nullptr;
};

class SpectralEngine {
  public:
    ~SpectralEngine();
    int init(const Embedding& emb);
    const Embedding& emb() const { return emb_; }

    // spec[bins] (device, real, scaled by 1/bins, digit-reversed layout) from a
    // top row top[cells] (device).  Work buffer: bins cplx.
    int spectrum(const double* top_dev, double* spec_dev, cplx* work, cudaStream_t st);

    // In place on grid slabs G[nslab][grid_pitch] (complex pairs): forward
    // pruned transform into S[nslab][bins].
    int forward(const cplx* G, cplx* S, int nslab, cudaStream_t st);
    // inverse pruned transform S -> G (cropped)
    int inverse(cplx* S, cplx* G, int nslab, cudaStream_t st);

    // S[pair][D][bins] <- (sum_q F_q[bin] B_q) S[pair][:][bin]   (in place)
    int mix(cplx* S, int npairs, int D, int Q, const double* spec /*[Q][bins]*/,
            const double* B /*[Q][D][D] device*/, cudaStream_t st);

    size_t work_elems(int nslab) const { return (size_t)nslab * emb_.bins; }

    // ---- fused path: G <- (sum_q B_q (x) T_q) G in place on the grid slabs ----
    // 2-D: transposing row pass -> [column FFT + mix + inverse column FFT in one kernel] -> row pass;
    // 1-D: one kernel (lines fit a CTA) or the four-step outer passes around the fused inner kernel.
    bool fused_supported(int D, int Q) const;
    // elements of scratch S needed per RHS pair on the fused path
    size_t fused_elems_per_pair(int D) const;
    // line-major spectra for the fused kernel: specL[q][line][pos]  (2-D: transposed copy of spec)
    // specP (same size, may be null): the copy in the mix layout of the 512-point column kernel
    // (spectral_col512.cuh), filled when the geometry uses it
    int spectrum_lines(const double* spec, double* specL, double* specP, int Q, cudaStream_t st);
    bool col512() const { return emb_.ndim == 2 && emb_.mt[0] == 512; }
    int apply_fused(cplx* G, cplx* S, int npairs, int D, int Q, const double* specL, const double* specP,
                    const MixSpec& mix, cudaStream_t st);

  private:
    Embedding emb_;
    cplx* tw_[3] = {nullptr, nullptr, nullptr};  // per-axis twiddle tables exp(-2 pi i k / mt_p)
    int tw_n_[3] = {0, 0, 0};
    FftPlan plan_[3];
    FftPlan plan1_, plan2_;                       // four-step sub-plans
    cplx* stage_tw_rows_ = nullptr;               // same for the transposing row pass (2-D)
    cplx* stage_tw_ = nullptr;                    // per-stage twiddles of the fused kernel's line length
    cplx* tw512_ = nullptr;                       // first-stage twiddles of the 512-point column kernel
    bool rows512_ = false;                        // both axes 512: register/bulk-copy row passes, specP ordered for them
};

// real grid vectors X[k][D*m] <-> complex pair slabs Z[ceil(k/2)][D][gpitch]
int pack_pairs(const double* X, int k, int D, long m, cplx* Z, long gpitch, cudaStream_t st);
int unpack_pairs(const cplx* Z, long gpitch, double* Y, int k, int D, long m, cudaStream_t st);

}  // namespace lmc
