// common.cuh -- shared device/host helpers of libkfacb (sm_100a).
//
// The fp32 SIMT GEMM core used by every non-tensor-core kernel of the hot
// path: 128x128 output tile per CTA, 256 threads, 8x8 register micro-tile per
// thread, BK = 32 k-slab staged through shared memory with register prefetch.
// Operands are described by small loader functors so one core serves the
// factor SYRKs (im2col / NCHW / row-major views), the Jacobi Gram + update
// products, the preconditioning GEMM chain and the blocked inverse.
#pragma once
#include <cuda_runtime.h>
#include <type_traits>
#include <stdint.h>
#include <stdio.h>
#include "kfacb.h"

namespace kfacb {

void set_error(const char* fmt, ...);

#define KFACB_CHECK_CUDA(expr)                                                     \
    do {                                                                           \
        cudaError_t _e = (expr);                                                   \
        if (_e != cudaSuccess) {                                                   \
            ::kfacb::set_error("%s failed: %s (%s:%d)", #expr,                     \
                               cudaGetErrorString(_e), __FILE__, __LINE__);        \
            return KFACB_ERR_CUDA;                                                 \
        }                                                                          \
    } while (0)

void count_launch(int n = 1);   // kernels launched by this library (reported by bench.py)

#define KFACB_CHECK_LAUNCH()                                                       \
    do { ::kfacb::count_launch(); KFACB_CHECK_CUDA(cudaGetLastError()); } while (0)

#define KFACB_REQUIRE(cond, ...)                                                   \
    do {                                                                           \
        if (!(cond)) { ::kfacb::set_error(__VA_ARGS__); return KFACB_ERR_ARG; }    \
    } while (0)

static inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }
static inline int cdiv(long a, long b) { return (int)((a + b - 1) / b); }

int sm_count();   // cached SM count of the current device

// -------------------------------------------------------------------------
// GEMM core
// -------------------------------------------------------------------------
constexpr int TILE = 128;         // output tile edge
constexpr int BK = 32;            // k-slab
constexpr int GEMM_THREADS = 256;
constexpr int PITCH = TILE + 4;   // smem row pitch (floats): keeps float4 alignment, spreads banks
constexpr int FRAG = 16;          // elements each thread stages per operand per slab

struct GemmSmem {
    float a[BK][PITCH];
    float b[BK][PITCH];
};

// Each operand tile is TILE (m or n) x BK (k).  Thread -> element mapping of the
// global fetch depends on which index is contiguous in memory:
//   K_FAST: lane walks k (32 consecutive k = one 128 B line), warps walk mn
//   else  : lane walks mn (128 consecutive mn), 2 k per pass
template <class L>
__device__ __forceinline__ void fetch_tile(const L& ld, int mn0, int k0, float (&r)[FRAG]) {
    const int tid = threadIdx.x;
    if (L::K_FAST) {
        const int kk = tid & 31, mm = tid >> 5;
#pragma unroll
        for (int i = 0; i < FRAG; ++i) r[i] = ld(mn0 + mm + 8 * i, k0 + kk);
    } else {
        const int mm = tid & 127, kk = tid >> 7;
#pragma unroll
        for (int i = 0; i < FRAG; ++i) r[i] = ld(mn0 + mm, k0 + kk + 2 * i);
    }
}

template <class L>
__device__ __forceinline__ void stage_tile(float (*s)[PITCH], const float (&r)[FRAG]) {
    const int tid = threadIdx.x;
    if (L::K_FAST) {
        const int kk = tid & 31, mm = tid >> 5;
#pragma unroll
        for (int i = 0; i < FRAG; ++i) s[kk][mm + 8 * i] = r[i];
    } else {
        const int mm = tid & 127, kk = tid >> 7;
#pragma unroll
        for (int i = 0; i < FRAG; ++i) s[kk + 2 * i][mm] = r[i];
    }
}

__device__ __forceinline__ void tile_mac(const GemmSmem& sm, float (&acc)[8][8]) {
    const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
#pragma unroll
    for (int k = 0; k < BK; ++k) {
        const float4 a0 = *reinterpret_cast<const float4*>(&sm.a[k][ty * 4]);
        const float4 a1 = *reinterpret_cast<const float4*>(&sm.a[k][64 + ty * 4]);
        const float4 b0 = *reinterpret_cast<const float4*>(&sm.b[k][tx * 4]);
        const float4 b1 = *reinterpret_cast<const float4*>(&sm.b[k][64 + tx * 4]);
        const float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
        const float b[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
        for (int i = 0; i < 8; ++i)
#pragma unroll
            for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
}

// row / col of micro-tile element (i, j) inside the 128x128 tile
__device__ __forceinline__ int frag_row(int i) { const int ty = threadIdx.x >> 4; return (i < 4 ? ty * 4 + i : 64 + ty * 4 + (i - 4)); }
__device__ __forceinline__ int frag_col(int j) { const int tx = threadIdx.x & 15; return (j < 4 ? tx * 4 + j : 64 + tx * 4 + (j - 4)); }

// acc += sum_{k in [kbeg,kend)} A(m0+., k) * B(n0+., k).  All threads of the CTA.
template <class LA, class LB>
__device__ __forceinline__ void gemm_mainloop(const LA& la, const LB& lb, int m0, int n0,
                                              int kbeg, int kend, GemmSmem& sm,
                                              float (&acc)[8][8]) {
    float ra[FRAG], rb[FRAG];
    if (kbeg < kend) {
        fetch_tile(la, m0, kbeg, ra);
        fetch_tile(lb, n0, kbeg, rb);
    }
    for (int k0 = kbeg; k0 < kend; k0 += BK) {
        __syncthreads();                    // previous slab fully consumed
        stage_tile<LA>(sm.a, ra);
        stage_tile<LB>(sm.b, rb);
        __syncthreads();
        if (k0 + BK < kend) {               // prefetch next slab into registers
            fetch_tile(la, m0, k0 + BK, ra);
            fetch_tile(lb, n0, k0 + BK, rb);
        }
        tile_mac(sm, acc);
    }
}

// -------------------------------------------------------------------------
// operand loaders: element (mn, k) of a TILE x K operand, 0 outside bounds
// -------------------------------------------------------------------------
// p[mn * s_mn + k]  (k contiguous)
struct LdKContig {
    static constexpr bool K_FAST = true;
    const float* p; long s_mn; int MN, K;
    __device__ __forceinline__ float operator()(int mn, int k) const {
        return (mn < MN && k < K) ? __ldg(p + (long)mn * s_mn + k) : 0.f;
    }
};
// p[k * s_k + mn]  (mn contiguous)
struct LdMnContig {
    static constexpr bool K_FAST = false;
    const float* p; long s_k; int MN, K;
    __device__ __forceinline__ float operator()(int mn, int k) const {
        return (mn < MN && k < K) ? __ldg(p + (long)k * s_k + mn) : 0.f;
    }
    __device__ __forceinline__ float4 quad(int mn, int k) const {
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (mn < MN) {
            const float* q = p + (long)k * s_k + mn;
            if (k < K) v.x = __ldg(q);
            if (k + 1 < K) v.y = __ldg(q + s_k);
            if (k + 2 < K) v.z = __ldg(q + 2 * s_k);
            if (k + 3 < K) v.w = __ldg(q + 3 * s_k);
        }
        return v;
    }
};

// -------------------------------------------------------------------------
// generic GEMM kernel:  C(m, n) = epi(m, n, sum_k A(m,k) B(n,k))
// grid = (tiles_n, tiles_m, batch); the functors may use blockIdx.z.
// -------------------------------------------------------------------------
// An epilogue may declare `static constexpr bool UPPER_ONLY = true` for a symmetric product: only the tiles on
// or above the diagonal are computed and the epilogue mirrors the off-diagonal ones itself.
template <class E, class = void>
struct epi_upper_only : std::false_type {};
template <class E>
struct epi_upper_only<E, std::void_t<decltype(E::UPPER_ONLY)>> : std::integral_constant<bool, E::UPPER_ONLY> {};

template <class LA, class LB, class Epi>
__global__ void __launch_bounds__(GEMM_THREADS, 2)
gemm_kernel(LA la, LB lb, Epi epi, int M, int N, int K) {
    __shared__ GemmSmem sm;
    if (epi_upper_only<Epi>::value && blockIdx.x < blockIdx.y) return;
    const int m0 = blockIdx.y * TILE, n0 = blockIdx.x * TILE;
    float acc[8][8];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;
    gemm_mainloop(la, lb, m0, n0, 0, K, sm, acc);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int m = m0 + frag_row(i);
        if (m >= M) continue;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int n = n0 + frag_col(j);
            if (n < N) epi(m, n, acc[i][j]);
        }
    }
}

// C[m*ldc + n] = alpha * v + beta * C[m*ldc + n]
struct EpiStore {
    float* C; long ldc; float alpha; float beta;
    __device__ __forceinline__ void operator()(int m, int n, float v) const {
        float* c = C + (long)m * ldc + n;
        *c = (beta == 0.f) ? alpha * v : fmaf(beta, *c, alpha * v);
    }
};

template <class LA, class LB, class Epi>
static inline cudaError_t launch_gemm(const LA& la, const LB& lb, const Epi& epi,
                                      int M, int N, int K, cudaStream_t st, int batch = 1) {
    dim3 grid(cdiv(N, TILE), cdiv(M, TILE), batch);
    gemm_kernel<LA, LB, Epi><<<grid, GEMM_THREADS, 0, st>>>(la, lb, epi, M, N, K);
    count_launch();
    return cudaGetLastError();
}

// -------------------------------------------------------------------------
// small reductions
// -------------------------------------------------------------------------
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

}  // namespace kfacb
