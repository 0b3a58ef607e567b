// Operator lifetime, parameter upload and the fused MVM pipeline
//   K~ V = W ( sum_q B_q (x) T_q ) W^T V + diag(noise) V
// (reference: SumMatrix.matvec over the tree assembled by gen_grid_kernel,
// runlmc/lmc/grid_kernel.py:49-74 and 126-136, call stack in SURVEY.md 3.2).
// All three reference representations (sum / bt / slfm) are the same matrix;
// here the sum over q is applied per frequency bin between ONE forward and ONE
// inverse transform per output instead of Q*D transform pairs.
#include "op.cuh"

#include <algorithm>
#include <cstdlib>
#include <vector>

namespace lmc {

static thread_local std::string g_err;   // per host thread, like errno
void set_error(const std::string& s) { g_err = s; }
const char* get_error() { return g_err.c_str(); }
std::atomic<unsigned long long> g_launches{0};

bool g_prof_on = false;
namespace {
struct ProfRec { int cat; cudaEvent_t a, b; };
std::vector<ProfRec> g_prof_recs;
cudaEvent_t g_prof_open[PROF_NCAT];
}  // namespace
void prof_begin(int cat, cudaStream_t st) {
    cudaEvent_t e;
    cudaEventCreate(&e);
    cudaEventRecord(e, st);
    g_prof_open[cat] = e;
}
void prof_end(int cat, cudaStream_t st) {
    cudaEvent_t e;
    cudaEventCreate(&e);
    cudaEventRecord(e, st);
    g_prof_recs.push_back({cat, g_prof_open[cat], e});
}

static size_t spectrum_tile_bytes() {   // pairs per spectral launch: large launches beat L2 residency (measured)
    static const char* e = getenv("LMC_SPECTRUM_TILE_MB");
    return (e ? (size_t)atoi(e) : 1536) << 20;   // config E: all 65 pairs of a 129-column block in one launch per pass
}

int op_ensure_workspace(lmc_op* op) {
    if (op->G) return 0;
    const size_t per_pair = sizeof(cplx) * (size_t)op->D * op->emb.bins;
    int tile = (int)std::max<size_t>(1, spectrum_tile_bytes() / per_pair);
    tile = std::min(tile, 256);
    op->tile_pairs = tile;
    const size_t fpp = sizeof(cplx) * op->eng.fused_elems_per_pair(op->D);
    op->fused_tile_pairs = fpp ? (int)std::max<size_t>(1, (per_pair * tile) / fpp) : 256;
    // interpolation runs over a wider block of pairs than the spectral stage: the grid slabs are
    // small (cells << bins) and the scatter/gather kernels amortise their per-point weights over pairs
    const size_t per_pair_g = sizeof(cplx) * (size_t)op->D * op->emb.grid_pitch;
    int gp = (int)std::max<size_t>(tile, std::min<size_t>(128, (3ull << 29) / per_pair_g));
    op->g_pairs = gp;
    const size_t gbytes = per_pair_g * gp;
    LMC_CHECK(cudaMalloc(&op->G, gbytes));
    LMC_CHECK(cudaMemset(op->G, 0, gbytes));
    LMC_CHECK(cudaMalloc(&op->S, per_pair * tile));
    LMC_CHECK(cudaMemset(op->S, 0, per_pair * tile));
    return 0;
}

int op_grid_apply(lmc_op* op, cplx* G, int npairs, int Q, const double* spec, const double* B,
                  cudaStream_t st) {
    LMC_TRY(op->eng.forward(G, op->S, npairs * op->D, st));
    LMC_TRY(op->eng.mix(op->S, npairs, op->D, Q, spec, B, st));
    LMC_TRY(op->eng.inverse(op->S, G, npairs * op->D, st));
    return 0;
}

// (sum_q B_q (x) T_q) applied in place to `cnt` RHS pairs of grid slabs, in L2-sized sub-tiles
int op_grid_block(lmc_op* op, cplx* G, int cnt, cudaStream_t st) {
    const size_t slab = (size_t)op->D * op->emb.grid_pitch;
    const int step = op->fused ? op->fused_tile_pairs : op->tile_pairs;
    for (int q0 = 0; q0 < cnt; q0 += step) {
        const int qc = std::min(step, cnt - q0);
        cplx* g = G + (size_t)q0 * slab;
        if (op->fused)
            LMC_TRY(op->eng.apply_fused(g, op->S, qc, op->D, op->Q, op->specL, op->specP, op->mix_spec(), st));
        else
            LMC_TRY(op_grid_apply(op, g, qc, op->Q, op->spec, op->B, st));
    }
    return 0;
}

int op_mvm(lmc_op* op, const ColumnView& cv, cudaStream_t st) {
    LMC_REQUIRE(op->Q > 0, "operator parameters not set (call lmc_op_set_params)");
    if (cv.ncols == 0) return 0;   // an empty block is a valid product (Matrix.matmat of an [n, 0] array)
    LMC_TRY(op_ensure_workspace(op));
    const int npairs = (cv.ncols + 1) / 2;
    const long n = op->ps.n;
    // Large blocks in the caller's point order are first brought into sorted order (one L2-friendly
    // gather per column), run through the coalesced sorted-order kernels in place, and taken back
    // with the inverse permutation.  Small blocks skip the two extra launches and let the kernels
    // index through the permutation.
    const bool big = !op->ps.identity && n * (long)cv.ncols >= (1L << 20);
    const bool presort = big && !cv.sorted_in, postsort = presort && !cv.sorted_out;
    if (presort && !op->Vs) LMC_CHECK(cudaMalloc(&op->Vs, sizeof(double) * (size_t)2 * op->g_pairs * n));
    const long ldo = cv.ld_out ? cv.ld_out : cv.ld;
    // equal blocks of pairs (a 129-column block is one block of 65 pairs or 33 + 32, never 64 + 1)
    const int nblocks = ceil_div(npairs, op->g_pairs);
    const int per_block = ceil_div(npairs, nblocks);
    for (int p0 = 0; p0 < npairs; p0 += per_block) {
        const int cnt = std::min(per_block, npairs - p0);
        ColumnView t = cv;
        const int c0 = 2 * p0;
        t.in = cv.in + (long)c0 * cv.ld;
        t.out = cv.out + (long)c0 * ldo;
        t.ld_out = ldo;
        t.ncols = std::min(2 * cnt, cv.ncols - c0);
        t.in_scale = cv.in_scale ? cv.in_scale + c0 : nullptr;
        t.active = cv.active ? cv.active + c0 : nullptr;
        if (presort) {
            LMC_TRY(permute_cols(op->ps, true, t.in, cv.ld, t.ncols, op->Vs, n, st));
            t.in = op->Vs;
            t.ld = n;
            t.sorted_in = true;
        }
        if (postsort) {   // the gather kernel reads in[c][i] and writes out[c][i] from the same thread
            t.out = op->Vs;
            t.ld_out = n;
            t.sorted_out = true;
        }
        LMC_TRY(to_grid(op->ps, t, op->G, st));
        LMC_TRY(op_grid_block(op, op->G, cnt, st));
        LMC_TRY(from_grid(op->ps, t, op->G, op->noise, st));
        if (postsort) LMC_TRY(permute_cols(op->ps, false, op->Vs, n, t.ncols, cv.out + (long)c0 * ldo, ldo, st));
    }
    return 0;
}

// K~ applied to a point-major block: X[i][c], Y[i][c] (numpy's C order for the [n, P] argument of the
// reference's Matrix.matmat, linalg/matrix.py:27-41), points in the caller's order.  The scatter stages its
// points straight from the rows of X (a point's columns are contiguous there, so the permutation into
// sorted order costs no pass of its own), the gather leaves its result in sorted column-major scratch and
// one transposing pass writes the rows of Y, adding the noise term D v on the way.
int op_mvm_rows(lmc_op* op, const double* X, long ldx, int P, double* Y, long ldy, cudaStream_t st) {
    LMC_REQUIRE(op->Q > 0, "operator parameters not set (call lmc_op_set_params)");
    if (P == 0) return 0;
    LMC_TRY(op_ensure_workspace(op));
    const int npairs = (P + 1) / 2;
    const long n = op->ps.n;
    if (!op->Vs) LMC_CHECK(cudaMalloc(&op->Vs, sizeof(double) * (size_t)2 * op->g_pairs * n));
    const int nblocks = ceil_div(npairs, op->g_pairs);
    const int per_block = ceil_div(npairs, nblocks);
    for (int p0 = 0; p0 < npairs; p0 += per_block) {
        const int cnt = std::min(per_block, npairs - p0);
        const int c0 = 2 * p0;
        ColumnView t;
        t.ncols = std::min(2 * cnt, P - c0);
        if (to_grid_takes_rows(op->ps, t.ncols)) {
            t.in = X + c0; t.ld = ldx; t.rows_in = true;
        } else {
            LMC_TRY(rows_to_sorted_cols(op->ps, X + c0, ldx, t.ncols, op->Vs, n, st));
            t.in = op->Vs; t.ld = n; t.sorted_in = true;
        }
        LMC_TRY(to_grid(op->ps, t, op->G, st));
        LMC_TRY(op_grid_block(op, op->G, cnt, st));
        ColumnView u;
        u.ncols = t.ncols;
        if (from_grid_writes_rows(op->ps)) {
            // the gather writes the caller's rows itself and reads X's rows beside them for the noise term
            u.in = X + c0; u.ld = ldx; u.rows_in = true;
            u.out = Y + c0; u.ld_out = ldy; u.rows_out = true;
            LMC_TRY(from_grid(op->ps, u, op->G, op->noise, st));
        } else {
            u.out = op->Vs; u.ld = u.ld_out = n; u.sorted_in = u.sorted_out = true;
            LMC_TRY(from_grid(op->ps, u, op->G, nullptr, st));
            LMC_TRY(sorted_cols_to_rows(op->ps, op->Vs, n, t.ncols, op->noise, X + c0, ldx, Y + c0, ldy, st));
        }
    }
    return 0;
}

}  // namespace lmc

lmc_op::~lmc_op() {
    lmc::free_points(&ps);
    for (int i = 0; i < 2; ++i) {
        cudaFree(stage_in[i]);
        cudaFree(stage_out[i]);
        if (ev_in[i]) cudaEventDestroy(ev_in[i]);
        if (ev_cmp[i]) cudaEventDestroy(ev_cmp[i]);
        if (ev_out[i]) cudaEventDestroy(ev_out[i]);
    }
    for (int i = 0; i < 3; ++i)
        if (hs[i]) cudaStreamDestroy(hs[i]);
    cudaFree(spec);
    cudaFree(specL);
    cudaFree(specP);
    cudaFree(B);
    cudaFree(noise);
    cudaFree(G);
    cudaFree(S);
    cudaFree(Vs);
    cudaFree(solver_ws);
    cudaFree(grad_ws);
    cudaFree(jacobi);
}

lmc_bttb::~lmc_bttb() {
    cudaFree(spec);
    cudaFree(one);
    cudaFree(G);
    cudaFree(S);
}

extern "C" {
static const char* kProfNames[lmc::PROF_NCAT] = {
    "to_grid", "fft_fwd_contig", "fft_fwd_strided", "mix", "fft_inv_strided", "fft_inv_contig",
    "from_grid", "minres_vec", "minres_scalar", "grad", "other"};
int lmc_profile_ncat(void) { return lmc::PROF_NCAT; }
const char* lmc_profile_name(int cat) { return (cat >= 0 && cat < lmc::PROF_NCAT) ? kProfNames[cat] : ""; }
int lmc_profile_begin(void) {
    for (auto& r : lmc::g_prof_recs) { cudaEventDestroy(r.a); cudaEventDestroy(r.b); }
    lmc::g_prof_recs.clear();
    lmc::g_prof_on = true;
    return 0;
}
// ms[cat] = total device time, counts[cat] = launches, since lmc_profile_begin()
int lmc_profile_end(double* ms, int* counts) {
    lmc::g_prof_on = false;
    LMC_CHECK(cudaDeviceSynchronize());
    for (int c = 0; c < lmc::PROF_NCAT; ++c) { ms[c] = 0.0; counts[c] = 0; }
    for (auto& r : lmc::g_prof_recs) {
        float t = 0.f;
        cudaEventElapsedTime(&t, r.a, r.b);
        ms[r.cat] += t;
        counts[r.cat] += 1;
        cudaEventDestroy(r.a);
        cudaEventDestroy(r.b);
    }
    lmc::g_prof_recs.clear();
    return 0;
}
}
