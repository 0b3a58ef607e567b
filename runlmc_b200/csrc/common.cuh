// Shared declarations for the sm_100a SKI-LMC kernels.
#pragma once
#include <cuda_runtime.h>
#include <atomic>
#include <cstdint>
#include <cstdio>
#include <string>

typedef double2 cplx;

namespace lmc {

// last error text for lmc_last_error()
void set_error(const std::string& s);
const char* get_error();

#define LMC_CHECK(call)                                                       \
    do {                                                                      \
        cudaError_t e_ = (call);                                              \
        if (e_ != cudaSuccess) {                                              \
            lmc::set_error(std::string(#call) + ": " + cudaGetErrorString(e_) \
                           + " at " + __FILE__ + ":" + std::to_string(__LINE__)); \
            return 2;                                                         \
        }                                                                     \
    } while (0)

#define LMC_REQUIRE(cond, msg)                                                \
    do {                                                                      \
        if (!(cond)) {                                                        \
            lmc::set_error(std::string("invalid argument: ") + (msg));        \
            return 1;                                                         \
        }                                                                     \
    } while (0)

#define LMC_TRY(expr)                                                         \
    do {                                                                      \
        int rc_ = (expr);                                                     \
        if (rc_ != 0) return rc_;                                             \
    } while (0)

// kernel launch counter (bench.py reports gpu_launches from it)
extern std::atomic<unsigned long long> g_launches;
inline void count_launch(int k = 1) { g_launches.fetch_add((unsigned long long)k, std::memory_order_relaxed); }

// Optional per-kernel-family timing with CUDA events on the launching stream
// (bench.py's roofline figures come from here).  Disabled => zero overhead.
enum ProfCat {
    PROF_TO_GRID = 0, PROF_FFT_FWD_CONTIG, PROF_FFT_FWD_STRIDED, PROF_MIX, PROF_FFT_INV_STRIDED,
    PROF_FFT_INV_CONTIG, PROF_FROM_GRID, PROF_MINRES_VEC, PROF_MINRES_SCALAR, PROF_GRAD, PROF_OTHER,
    PROF_NCAT
};
extern bool g_prof_on;
void prof_begin(int cat, cudaStream_t st);
void prof_end(int cat, cudaStream_t st);
struct ProfScope {
    int cat; cudaStream_t st;
    ProfScope(int c, cudaStream_t s) : cat(c), st(s) { if (g_prof_on) prof_begin(cat, st); }
    ~ProfScope() { if (g_prof_on) prof_end(cat, st); }
};

static inline int ceil_div(long a, long b) { return (int)((a + b - 1) / b); }
static inline int ilog2(unsigned v) { int k = 0; while ((1u << k) < v) ++k; return k; }

__device__ __forceinline__ cplx cadd(cplx a, cplx b) { return make_double2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ cplx csub(cplx a, cplx b) { return make_double2(a.x - b.x, a.y - b.y); }
__device__ __forceinline__ cplx cmul(cplx a, cplx b) {
    return make_double2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}
// a * conj(b)
__device__ __forceinline__ cplx cmulc(cplx a, cplx b) {
    return make_double2(a.x * b.x + a.y * b.y, a.y * b.x - a.x * b.y);
}

}  // namespace lmc
