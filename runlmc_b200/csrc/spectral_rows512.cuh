// Transposing 2-D row passes for 512-point rows (128 < m_y <= 256), register-resident FFT + TMA on both
// sides: the rows of G by 1-D bulk copies, the transposed tiles of S_T by 2-D tensor copies
// (rows512_*_tma_kernel below: cp.async.bulk.tensor.2d with the 128-byte swizzle, two copies per 64 KB tile;
// the *_kernel variants without tensor maps move the same tiles as 512 bulk copies of 128 bytes and remain as
// the fall-back when the driver's tensor-map encoder cannot be found).  With spectral_col512.cuh they replace numpy's
// rfftn / irfftn inside BTTB.matvec (reference runlmc/linalg/bttb.py:144-148) for config-E geometries:
//
//   forward   G[slab][x][y]  (y contiguous)  ->  S_T[slab][pos][x]   (x contiguous)
//   inverse   S_T[slab][pos][x]              ->  G[slab][x][y]       (cropped to m_y)
//
// One CTA owns 8 consecutive rows x0..x0+7 (one warp per row) of `spc` consecutive slabs.
//
// Data movement: the 8 rows of a slab (4 KB each, contiguous) arrive in shared memory by
// cp.async.bulk + mbarrier (complete_tx); the next slab's rows are requested as soon as the current ones
// sit in registers, so the copy overlaps the transform.  The transposed side moves as 512 chunks of
// 8 complex = 128 B (one frequency position of the CTA's 8 rows): the forward pass assembles them in a
// padded shared-memory tile [pos][9] (conflict free for the per-position writes of a warp) and hands every
// chunk to cp.async.bulk.global.shared::cta; the inverse pass fetches the 512 chunks with cp.async.bulk into
// the same padded tile and the warps pick their row out of it.  No thread-issued global load or store
// touches the transposed array.
//
// Transform: the 512-point scheme of spectral_col512.cuh (DFT16 in registers, one exchange through shared
// memory, DFT16, radix-2 butterfly across the lane pair with __shfl_xor).  Position pos = k2 * 32 + lane
// holds frequency c + 16 (k2 + 16 h) (c = lane >> 1, h = lane & 1) with its true sign; the spectra the
// column kernel multiplies with are stored in the same order along both axes (col512_spec2_kernel).
#pragma once
#include <cuda.h>
#include "spectral_col512.cuh"

namespace lmc {

static const int kR512TilePitch = 9;                         // complex elements per position in the transposed tile
static const int kR512Tile = 512 * kR512TilePitch;           // >= 8 * kC512Line: also the 8 warps' exchange buffers
static const int kR512In = 8 * 256;                          // staged rows (forward)
static const size_t kR512SmemFwd = sizeof(cplx) * (size_t)(kR512Tile + kR512In) + 16;
static const size_t kR512SmemInv = sizeof(cplx) * (size_t)kR512Tile + 16;
static_assert(kR512Tile >= 8 * kC512Line, "tile must hold the exchange buffers");

__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long* bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(count));
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(unsigned long long* bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, unsigned parity) {
    asm volatile(
        "{\n"
        ".reg .pred P1;\n"
        "LMC_WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
        "@P1 bra LMC_DONE_%=;\n"
        "bra LMC_WAIT_%=;\n"
        "LMC_DONE_%=:\n"
        "}\n" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
// global -> shared, completion counted in bytes on the mbarrier
__device__ __forceinline__ void bulk_g2s(void* dst_smem, const void* src, unsigned bytes, unsigned long long* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n" ::"r"(
                     smem_u32(dst_smem)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
// shared -> global, tracked by the issuing thread's bulk async-group
__device__ __forceinline__ void bulk_s2g(void* dst, const void* src_smem, unsigned bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;\n" ::"l"(dst), "r"(smem_u32(src_smem)),
                 "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;\n" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;\n" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;\n" ::: "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory"); }

// x[a] = row[lane + 32 a], a < 8 (zero padded to 16) -> x[k2] = X[c + 16 (k2 + 16 h)], lane = 2 c + h
__device__ __forceinline__ void warp_fft512_fwd(cplx* x, cplx* buf, cplx w1, int lane) {
    const int c2 = lane >> 1, h = lane & 1;
    const double sgn = h ? -1.0 : 1.0;
    cplx* bx = buf + c2 * 34 + h;
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
    for (int k = 0; k < 16; ++k) {   // h = 0: Z0 + W Z1 = X32[k];  h = 1: Z0 - W Z1 = X32[k + 16]
        const double rx = __shfl_xor_sync(0xffffffffu, x[k].x, 1);
        const double ry = __shfl_xor_sync(0xffffffffu, x[k].y, 1);
        x[k].x = fma(sgn, x[k].x, rx);
        x[k].y = fma(sgn, x[k].y, ry);
    }
}

// mirror: x[k2] in the position layout -> x[a] = row[lane + 32 a], a < 8 (unscaled inverse)
__device__ __forceinline__ void warp_fft512_inv(cplx* x, cplx* buf, cplx w1, int lane) {
    const int c2 = lane >> 1, h = lane & 1;
    const double sgn = h ? -1.0 : 1.0;
    cplx* bx = buf + c2 * 34 + h;
#pragma unroll
    for (int k = 0; k < 16; ++k) {   // h = 0: S + Dd = Z0;  h = 1: S - Dd = W Z1
        const double rx = __shfl_xor_sync(0xffffffffu, x[k].x, 1);
        const double ry = __shfl_xor_sync(0xffffffffu, x[k].y, 1);
        x[k].x = fma(sgn, x[k].x, rx);
        x[k].y = fma(sgn, x[k].y, ry);
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
    __syncwarp();
}

struct Rows512Args {
    const cplx* G_in;     // forward source  [slab][mx][my]
    cplx* G_out;          // inverse destination
    cplx* ST;             // [slab][512][xpitch]
    long g_slab, st_slab; // elements per slab
    int mx, my, xpitch, nslab, spc;
    const cplx* tw1;      // [15][32] W512^{m c}
};

__global__ void __launch_bounds__(256, 2) rows512_fwd_kernel(const Rows512Args a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    cplx* tile = reinterpret_cast<cplx*>(smem_raw);    // [512][9], first the warps' exchange buffers
    cplx* in = tile + kR512Tile;                       // [8][256]
    unsigned long long* bar = reinterpret_cast<unsigned long long*>(in + kR512In);
    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    const int x0 = blockIdx.x * 8;
    const long slab0 = (long)blockIdx.y * a.spc;
    const int cnt = (int)min((long)a.spc, (long)a.nslab - slab0);
    const int nrows = min(8, a.mx - x0);
    const unsigned row_bytes = (unsigned)a.my * (unsigned)sizeof(cplx);
    const bool row_ok = w < nrows;
    if (tid == 0) mbar_init(bar, 1);
    const cplx w1 = a.tw1[lane];                       // W512^lane
    __syncthreads();
    if (tid == 0) {
        mbar_arrive_expect_tx(bar, row_bytes * nrows);
        for (int r = 0; r < nrows; ++r)
            bulk_g2s(in + r * 256, a.G_in + slab0 * a.g_slab + (long)(x0 + r) * a.my, row_bytes, bar);
    }
    cplx* buf = tile + w * kC512Line;
    for (int t = 0; t < cnt; ++t) {
        const long slab = slab0 + t;
        mbar_wait(bar, t & 1);
        cplx x[16];
#pragma unroll
        for (int r = 0; r < 8; ++r)
            x[r] = (row_ok && lane + 32 * r < a.my) ? in[w * 256 + lane + 32 * r] : make_double2(0.0, 0.0);
        bulk_wait_read();                              // this thread's chunk stores of the previous slab have left the tile
        __syncthreads();                               // rows are in registers, tile is free
        if (tid == 0 && t + 1 < cnt) {
            mbar_arrive_expect_tx(bar, row_bytes * nrows);
            for (int r = 0; r < nrows; ++r)
                bulk_g2s(in + r * 256, a.G_in + (slab + 1) * a.g_slab + (long)(x0 + r) * a.my, row_bytes, bar);
        }
        warp_fft512_fwd(x, buf, w1, lane);
        __syncthreads();                               // every warp is done with its exchange buffer
#pragma unroll
        for (int k = 0; k < 16; ++k) tile[(k * 32 + lane) * kR512TilePitch + w] = x[k];
        fence_async_smem();                            // make the tile visible to the bulk-copy engine
        __syncthreads();
        cplx* dst = a.ST + slab * a.st_slab + x0;
#pragma unroll
        for (int j = 0; j < 2; ++j) {
            const int pos = tid + 256 * j;
            bulk_s2g(dst + (long)pos * a.xpitch, tile + pos * kR512TilePitch, 8 * sizeof(cplx));
        }
        bulk_commit();
    }
    bulk_wait_all();
}

__global__ void __launch_bounds__(256, 2) rows512_inv_kernel(const Rows512Args a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    cplx* tile = reinterpret_cast<cplx*>(smem_raw);
    unsigned long long* bar = reinterpret_cast<unsigned long long*>(tile + kR512Tile);
    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    const int x0 = blockIdx.x * 8;
    const long slab0 = (long)blockIdx.y * a.spc;
    const int cnt = (int)min((long)a.spc, (long)a.nslab - slab0);
    const bool row_ok = x0 + w < a.mx;
    if (tid == 0) mbar_init(bar, 1);
    const cplx w1 = a.tw1[lane];                       // W512^lane
    __syncthreads();                                   // barrier object visible to everyone
    cplx* buf = tile + w * kC512Line;
    for (int t = 0; t < cnt; ++t) {
        const long slab = slab0 + t;
        if (tid == 0) mbar_arrive_expect_tx(bar, 512u * 8u * (unsigned)sizeof(cplx));
        __syncthreads();                               // barrier armed (and, from the second slab on, tile free)
        const cplx* src = a.ST + slab * a.st_slab + x0;
#pragma unroll
        for (int j = 0; j < 2; ++j) {
            const int pos = tid + 256 * j;
            bulk_g2s(tile + pos * kR512TilePitch, src + (long)pos * a.xpitch, 8 * sizeof(cplx), bar);
        }
        mbar_wait(bar, t & 1);
        cplx x[16];
#pragma unroll
        for (int k = 0; k < 16; ++k) x[k] = tile[(k * 32 + lane) * kR512TilePitch + w];
        __syncthreads();                               // every warp has its row: the tile becomes the exchange buffers
        warp_fft512_inv(x, buf, w1, lane);
        if (row_ok) {
            cplx* g = a.G_out + slab * a.g_slab + (long)(x0 + w) * a.my + lane;
#pragma unroll
            for (int r = 0; r < 8; ++r)
                if (lane + 32 * r < a.my) g[32 * r] = x[r];
        }
        __syncthreads();                               // exchange buffers done before the next chunks land
    }
}

// ---------------------------------------------------------------------------
// Inverse pass with the transposed tile fetched by two TENSOR copies (cp.async.bulk.tensor.2d, a 3-D view
// would add nothing: S_T is [slab * 512 + pos][x] with a constant row pitch): box = 8 complex (128 B) x 256
// positions, 128-byte swizzle, so the 64 KB tile lands dense in shared memory and a warp still picks its row
// out of it without bank conflicts (chunk index XOR (pos & 7)).  One elected thread issues two copies per
// slab instead of 512 bulk copies issued by all threads.
// ---------------------------------------------------------------------------
static const size_t kR512SmemInvTma = sizeof(cplx) * (size_t)(8 * kC512Line) + 16 + 1024;
static_assert(8 * kC512Line >= 512 * 8, "exchange buffers cover the dense tile");

__device__ __forceinline__ void tma_load_2d(void* dst_smem, const CUtensorMap* tmap, int c0, int c1,
                                            unsigned long long* bar) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];\n" ::"r"(
            smem_u32(dst_smem)), "l"(reinterpret_cast<unsigned long long>(tmap)), "r"(c0), "r"(c1), "r"(smem_u32(bar))
        : "memory");
}

__global__ void __launch_bounds__(256, 2) rows512_inv_tma_kernel(const Rows512Args a,
                                                                 const __grid_constant__ CUtensorMap tmap) {
    extern __shared__ unsigned char smem_raw_t[];
    unsigned char* base = reinterpret_cast<unsigned char*>(
        (reinterpret_cast<unsigned long long>(smem_raw_t) + 1023ull) & ~1023ull);   // swizzle atoms are 1 KB
    cplx* tile = reinterpret_cast<cplx*>(base);                   // [512][8] swizzled, then the exchange buffers
    unsigned long long* bar = reinterpret_cast<unsigned long long*>(tile + 8 * kC512Line);
    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    const int x0 = blockIdx.x * 8;
    const long slab0 = (long)blockIdx.y * a.spc;
    const int cnt = (int)min((long)a.spc, (long)a.nslab - slab0);
    const bool row_ok = x0 + w < a.mx;
    if (tid == 0) mbar_init(bar, 1);
    const cplx w1 = a.tw1[lane];
    __syncthreads();
    cplx* buf = tile + w * kC512Line;
    for (int t = 0; t < cnt; ++t) {
        const long slab = slab0 + t;
        if (tid == 0) {
            fence_async_smem();                        // the warps' generic-proxy use of the tile precedes the copies
            mbar_arrive_expect_tx(bar, 512u * 8u * (unsigned)sizeof(cplx));
            tma_load_2d(tile, &tmap, 2 * x0, (int)(slab * 512), bar);
            tma_load_2d(tile + 256 * 8, &tmap, 2 * x0, (int)(slab * 512 + 256), bar);
        }
        mbar_wait(bar, t & 1);
        cplx x[16];
#pragma unroll
        for (int k = 0; k < 16; ++k) {
            const int pos = k * 32 + lane;
            x[k] = tile[pos * 8 + (w ^ (pos & 7))];
        }
        __syncthreads();                               // every warp has its row: the tile becomes the exchange buffers
        warp_fft512_inv(x, buf, w1, lane);
        if (row_ok) {
            cplx* g = a.G_out + slab * a.g_slab + (long)(x0 + w) * a.my + lane;
#pragma unroll
            for (int r = 0; r < 8; ++r)
                if (lane + 32 * r < a.my) g[32 * r] = x[r];
        }
        __syncthreads();                               // exchange buffers done before the next tile lands
    }
}

// Forward pass, transposed side stored by two tensor copies from the swizzled dense tile.
static const size_t kR512SmemFwdTma = sizeof(cplx) * (size_t)(8 * kC512Line + kR512In) + 16 + 1024;

__device__ __forceinline__ void tma_store_2d(const CUtensorMap* tmap, int c0, int c1, const void* src_smem) {
    asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%1, %2}], [%3];\n" ::"l"(
                     reinterpret_cast<unsigned long long>(tmap)), "r"(c0), "r"(c1), "r"(smem_u32(src_smem))
                 : "memory");
}

__global__ void __launch_bounds__(256, 2) rows512_fwd_tma_kernel(const Rows512Args a,
                                                                 const __grid_constant__ CUtensorMap tmap) {
    extern __shared__ unsigned char smem_raw_t[];
    unsigned char* base = reinterpret_cast<unsigned char*>(
        (reinterpret_cast<unsigned long long>(smem_raw_t) + 1023ull) & ~1023ull);
    cplx* tile = reinterpret_cast<cplx*>(base);        // the warps' exchange buffers, then [512][8] swizzled
    cplx* in = tile + 8 * kC512Line;                   // [8][256]
    unsigned long long* bar = reinterpret_cast<unsigned long long*>(in + kR512In);
    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    const int x0 = blockIdx.x * 8;
    const long slab0 = (long)blockIdx.y * a.spc;
    const int cnt = (int)min((long)a.spc, (long)a.nslab - slab0);
    const int nrows = min(8, a.mx - x0);
    const unsigned row_bytes = (unsigned)a.my * (unsigned)sizeof(cplx);
    const bool row_ok = w < nrows;
    if (tid == 0) mbar_init(bar, 1);
    const cplx w1 = a.tw1[lane];
    __syncthreads();
    if (tid == 0) {
        mbar_arrive_expect_tx(bar, row_bytes * nrows);
        for (int r = 0; r < nrows; ++r)
            bulk_g2s(in + r * 256, a.G_in + slab0 * a.g_slab + (long)(x0 + r) * a.my, row_bytes, bar);
    }
    cplx* buf = tile + w * kC512Line;
    for (int t = 0; t < cnt; ++t) {
        const long slab = slab0 + t;
        mbar_wait(bar, t & 1);
        cplx x[16];
#pragma unroll
        for (int r = 0; r < 8; ++r)
            x[r] = (row_ok && lane + 32 * r < a.my) ? in[w * 256 + lane + 32 * r] : make_double2(0.0, 0.0);
        if (tid == 0) bulk_wait_read();                // the previous slab's tensor stores have read the tile
        __syncthreads();                               // rows are in registers, tile is free
        if (tid == 0 && t + 1 < cnt) {
            mbar_arrive_expect_tx(bar, row_bytes * nrows);
            for (int r = 0; r < nrows; ++r)
                bulk_g2s(in + r * 256, a.G_in + (slab + 1) * a.g_slab + (long)(x0 + r) * a.my, row_bytes, bar);
        }
        warp_fft512_fwd(x, buf, w1, lane);
        __syncthreads();                               // every warp is done with its exchange buffer
#pragma unroll
        for (int k = 0; k < 16; ++k) {
            const int pos = k * 32 + lane;
            tile[pos * 8 + (w ^ (pos & 7))] = x[k];
        }
        fence_async_smem();                            // make the tile visible to the copy engine
        __syncthreads();
        if (tid == 0) {
            tma_store_2d(&tmap, 2 * x0, (int)(slab * 512), tile);
            tma_store_2d(&tmap, 2 * x0, (int)(slab * 512 + 256), tile + 256 * 8);
            bulk_commit();
        }
    }
    if (tid == 0) bulk_wait_all();
}

// specP2[q][l][p] = specL[q][old(k(l))][old(k(p))]: both axes in the position order of the register
// transform (k(.) as in the file header; old(.) the digit-reversed position of the radix-8 plan)
static __global__ void col512_spec2_kernel(const double* __restrict__ specL, double* specP, long total) {
    const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    const int p = (int)(i & 511), l = (int)((i >> 9) & 511);
    const long q = i >> 18;
    auto old_of = [](int pos) {
        const int lane = pos & 31, k2 = pos >> 5;
        const int k = (lane >> 1) + 16 * (k2 + 16 * (lane & 1));
        return (k & 7) * 64 + ((k >> 3) & 7) * 8 + (k >> 6);
    };
    specP[i] = specL[(q << 18) + ((long)old_of(l) << 9) + old_of(p)];
}

}  // namespace lmc
