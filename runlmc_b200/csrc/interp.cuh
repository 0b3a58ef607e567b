// Cubic-convolution interpolation between scattered points and the inducing grid.
#pragma once
#include "common.cuh"
#include <vector>

namespace lmc {

// Device-resident description of the (fixed) inputs X, sorted by
// (output, grid bin).  Built once per operator by build_points() on the host.
struct PointSet {
    int D = 0, ndim = 0;
    long n = 0;
    int m[2] = {1, 1};    // grid sizes
    int nb[2] = {1, 1};   // bins per axis = m + 3  (clamped base index i0 in [-2, m])
    long NB = 0;          // bins per output
    long grid_pitch = 0;  // cells allocated per (pair, output) slab
    bool identity = false;  // sorted order == caller's order
    int max_tile_pts_8x8 = 0;     // 2-D: most points in the 11x11 bins around an 8x8-cell scatter tile
    int max_tile_pts_16x8 = 0;    // 2-D: same for the 19x11 bins around a 16x8-cell tile
    int max_gather_tile_pts = 0;  // 2-D: most points in a 16x16-bin gather tile
    int* perm = nullptr;     // [n] sorted position -> caller's index
    int* iperm = nullptr;    // [n] caller's index -> sorted position (nullptr when identity)
    double* u[2] = {nullptr, nullptr};  // [n] fractional offsets (sorted order)
    int* i0[2] = {nullptr, nullptr};    // [n] clamped base index (sorted order)
    int* bin_start = nullptr;           // [D*NB + 1] offsets into the sorted arrays
    long out_start[17] = {0};           // host copy, [D+1] point offsets per output
    long* out_start_dev = nullptr;
};

// Host-side build (counting sort by bin); reproduces f=(s-g0)/delta, i0=floor(f),
// u=f-i0 of reference approx/interpolation.py:98-101 bit for bit.
int build_points(PointSet* ps, int D, int ndim, const int* grid_sizes, const double* origin,
                 const double* delta, const int* lens, const double* X_host, long grid_pitch);
// Same point set from coordinates already on the device: keys, a stable radix sort and the bin CSR
// are computed there (setup.cu); identical result, bit for bit.
int build_points_dev(PointSet* ps, int D, int ndim, const int* grid_sizes, const double* origin,
                     const double* delta, const int* lens, const double* X_dev, long grid_pitch,
                     cudaStream_t st);
void tile_populations(PointSet* ps, const std::vector<int>& start);
void free_points(PointSet* ps);

struct ColumnView {
    const double* in = nullptr;  // [ncols][ld] input vectors (caller's order unless sorted_in)
    double* out = nullptr;       // [ncols][ld_out] outputs (caller's order unless sorted_out)
    long ld = 0;
    long ld_out = 0;             // 0: same as ld
    int ncols = 0;
    const double* in_scale = nullptr;  // optional per-column factor applied to `in`
    const int* active = nullptr;       // optional per-column flag; pairs with no active column are skipped
    bool sorted_in = false;            // `in` already is in the operator's sorted point order
    bool sorted_out = false;           // `out` is wanted in sorted order
    bool rows_in = false;              // `in` is point-major: in[point][column], row stride ld, caller's order
                                       // (to_grid where to_grid_takes_rows() says so; from_grid with rows_out)
    bool rows_out = false;             // `out` is point-major too (from_grid where from_grid_writes_rows())
};

// block of columns between the caller's point order and the sorted order (out[c][i] = in[c][p[i]],
// p = perm when to_sorted, else its inverse)
int permute_cols(const PointSet& ps, bool to_sorted, const double* in, long ld, int ncols, double* out,
                 long ldo, cudaStream_t st);
// point-major blocks (rows[i][c], caller's order) <-> sorted column-major; the way back adds noise_d * rows_in
int rows_to_sorted_cols(const PointSet& ps, const double* rows, long ldr, int ncols, double* cols, long ldc,
                        cudaStream_t st);
int sorted_cols_to_rows(const PointSet& ps, const double* cols, long ldc, int ncols, const double* noise,
                        const double* rows_in, long ldi, double* rows_out, long ldo, cudaStream_t st);
bool to_grid_takes_rows(const PointSet& ps, int ncols);
bool from_grid_writes_rows(const PointSet& ps);
// G[pair][d][cell] (complex pairs of columns 2p, 2p+1)  <-  W^T in     (deterministic, no atomics)
int to_grid(const PointSet& ps, const ColumnView& cv, cplx* G, cudaStream_t st);
// out = W G (+ noise_d * in when noise != nullptr)
int from_grid(const PointSet& ps, const ColumnView& cv, const cplx* G, const double* noise,
              cudaStream_t st);

}  // namespace lmc
