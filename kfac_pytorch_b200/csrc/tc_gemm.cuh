// tc_gemm.cuh -- generic tcgen05 3xTF32 GEMM core:  D(128x128, TMEM) = sum_k A(m,k) B(n,k)
// with operands described by the same loader functors as the SIMT core (common.cuh).
//
// CTA = 256 threads.  Per 32-wide k-slab every thread gathers 4 k-quads of A and 4 of B
// (coalesced: lanes walk k for K_FAST loaders, rows otherwise), splits them into tf32 hi/lo and
// stores 16-byte units into the no-swizzle K-major core-matrix layout
//     (row, k) -> (row & 7) * 16 + (row >> 3) * 128 + (k >> 2) * LBO + (k & 3) * 4
// One thread issues 3 passes x 4 UMMAs (M = N = 128, K = 8) per slab into a 128-column fp32 TMEM
// accumulator and commits to an mbarrier; the gathers of slab s+1 are in flight meanwhile.
// The accumulator is drained through shared memory so that epilogues run with lanes along n
// (coalesced global access, conflict-free transposed reads for mirrored stores).
#pragma once
#include <type_traits>
#include <utility>
#include "common.cuh"
#include "tc.cuh"

namespace kfacb {
namespace tcg {

constexpr int THREADS = 256;
constexpr int BKS = 32;                                  // k-slab
constexpr int SBO = 128;
constexpr int LBO = (TILE / 8) * 128 + 16;               // 2064 (+16: conflict-free staging stores)
constexpr int PART = (BKS / 4) * LBO;                    // 16512 bytes per operand part
constexpr int SMEM = 4 * PART;                           // A_hi, A_lo, B_hi, B_lo = 66048 = 128 * 129 * 4
constexpr int DPITCH = TILE + 1;                         // drained accumulator pitch (floats)
static_assert(SMEM >= TILE * DPITCH * 4, "drain buffer must fit in the operand buffers");

// The tensor core accumulates into TMEM with round-toward-zero: a long chain of accumulations picks up a
// systematic bias of ~3e-8 per UMMA K-step (measured: -1.4e-5 on the trace of a 2304-term product).  The
// accumulator is therefore FLUSHED into fp32 registers (round-to-nearest adds) every FLUSH_SLABS k-slabs, which
// bounds the in-TMEM chain to 32 K-steps (~1e-6) and makes the cross-chunk summation unbiased.
constexpr int FLUSH_SLABS = 8;

struct Ctx {
    unsigned char* dsm;
    uint64_t* mbar;
    uint32_t tmem;
    int slabs_done;       // slabs committed so far (mbarrier phase bookkeeping)
    int flushed;          // 1 once acc[] holds flushed partial sums
    float acc[64];        // this thread's share of the 128x128 tile (row 32*(warp&3)+lane, 64 columns)
};

// call at kernel start (all threads); `tmem_slot` and `mbar` live in static shared memory
__device__ __forceinline__ void setup(Ctx& c, unsigned char* dsm, uint64_t* mbar, uint32_t* tmem_slot) {
    if (threadIdx.x == 0) { tc::mbar_init(mbar, 1); tc::fence_mbar_init(); }
    if ((threadIdx.x >> 5) == 0) tc::tmem_alloc<128>(tmem_slot);
    tc::tc_fence_before();
    __syncthreads();
    tc::tc_fence_after();
    c.dsm = dsm; c.mbar = mbar; c.tmem = *tmem_slot; c.slabs_done = 0; c.flushed = 0;
}
__device__ __forceinline__ void teardown(Ctx& c) {
    tc::tc_fence_before();
    __syncthreads();
    if ((threadIdx.x >> 5) == 0) tc::tmem_dealloc<128>(c.tmem);
}

// Loaders may provide `float4 quad(mn, k)` = elements (mn, k..k+3) with k % 4 == 0 (one address computation,
// vector load where the memory layout allows); the generic path calls operator() four times.
template <class L, class = void>
struct has_quad : std::false_type {};
template <class L>
struct has_quad<L, std::void_t<decltype(std::declval<const L&>().quad(0, 0))>> : std::true_type {};

template <class L>
__device__ __forceinline__ float4 load_quad(const L& ld, int mn, int k) {
    if constexpr (has_quad<L>::value) {
        return ld.quad(mn, k);
    } else {
        return make_float4(ld(mn, k), ld(mn, k + 1), ld(mn, k + 2), ld(mn, k + 3));
    }
}

template <class L>
__device__ __forceinline__ void fetch(const L& ld, int mn0, int k0, float4 (&r)[4]) {
    const int tid = threadIdx.x;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int idx = tid + THREADS * i;
        int row, kq;
        if (L::K_FAST) { kq = idx & 7; row = idx >> 3; } else { row = idx & 127; kq = idx >> 7; }
        r[i] = load_quad(ld, mn0 + row, k0 + 4 * kq);
    }
}
template <class L>
__device__ __forceinline__ void stage(unsigned char* hi_base, unsigned char* lo_base, const float4 (&r)[4]) {
    const int tid = threadIdx.x;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int idx = tid + THREADS * i;
        int row, kq;
        if (L::K_FAST) { kq = idx & 7; row = idx >> 3; } else { row = idx & 127; kq = idx >> 7; }
        const int off = (row & 7) * 16 + (row >> 3) * SBO + kq * LBO;
        float4 hi, lo;
        tc::split_tf32(r[i], hi, lo);
        *reinterpret_cast<float4*>(hi_base + off) = hi;
        *reinterpret_cast<float4*>(lo_base + off) = lo;
    }
}

// add the TMEM accumulator into the register accumulator (all threads; the MMAs must have completed)
__device__ __forceinline__ void flush_add(Ctx& c) {
    const int warp = threadIdx.x >> 5;
    const uint32_t taddr = c.tmem + ((uint32_t)(32 * (warp & 3)) << 16) + (warp >> 2) * 64;
    float v[32];
#pragma unroll
    for (int h = 0; h < 2; ++h) {
        tc::tmem_ld32(taddr + 32 * h, v);
#pragma unroll
        for (int i = 0; i < 32; ++i) c.acc[32 * h + i] = c.flushed ? c.acc[32 * h + i] + v[i] : v[i];
    }
    c.flushed = 1;
    tc::tc_fence_before();        // ordered before the next __syncthreads / UMMA that overwrites the accumulator
}

// accumulate over k in [kbeg, kend) into TMEM columns [0,128) (+ the flushed registers); returns when complete
// FLUSH = false: the caller bounds the chain itself (split-K chunks of <= 512 with an fp32 second stage) and
// saves the 64 accumulator registers (3 CTAs / SM instead of 1).
template <bool FLUSH, class LA, class LB>
__device__ __forceinline__ void mainloop(const LA& la, const LB& lb, int m0, int n0, int kbeg, int kend,
                                         bool same_operand, Ctx& c) {
    unsigned char* a_hi = c.dsm;
    unsigned char* a_lo = c.dsm + PART;
    unsigned char* b_hi = same_operand ? a_hi : c.dsm + 2 * PART;
    unsigned char* b_lo = same_operand ? a_lo : c.dsm + 3 * PART;
    float4 ra[4], rb[4];
    const int nslab = (kend - kbeg + BKS - 1) / BKS;
    if (nslab > 0) { fetch(la, m0, kbeg, ra); if (!same_operand) fetch(lb, n0, kbeg, rb); }
    constexpr uint32_t idesc = tc::idesc_tf32(128, 128, 0, 0);
    int since = 0;                                           // slabs accumulated in TMEM since the last flush
    for (int s = 0; s < nslab; ++s) {
        if (c.slabs_done > 0) tc::mbar_wait(c.mbar, (uint32_t)((c.slabs_done - 1) & 1));   // smem free again
        if (FLUSH && since == FLUSH_SLABS) { tc::tc_fence_after(); flush_add(c); since = 0; }
        stage<LA>(a_hi, a_lo, ra);
        if (!same_operand) stage<LB>(b_hi, b_lo, rb);
        if (s + 1 < nslab) {
            fetch(la, m0, kbeg + (s + 1) * BKS, ra);
            if (!same_operand) fetch(lb, n0, kbeg + (s + 1) * BKS, rb);
        }
        tc::fence_async_smem();
        __syncthreads();
        if (threadIdx.x == 0) {
            tc::tc_fence_after();
            const uint32_t ah = tc::smem_u32(a_hi), al = tc::smem_u32(a_lo);
            const uint32_t bh = tc::smem_u32(b_hi), bl = tc::smem_u32(b_lo);
#pragma unroll
            for (int pass = 0; pass < 3; ++pass) {
                const uint32_t ab = pass == 1 ? al : ah;
                const uint32_t bb = pass == 2 ? bl : bh;
#pragma unroll
                for (int j = 0; j < BKS / 8; ++j) {
                    const uint64_t ad = tc::smem_desc(ab + 2 * j * LBO, LBO, SBO);
                    const uint64_t bd = tc::smem_desc(bb + 2 * j * LBO, LBO, SBO);
                    tc::umma_tf32(c.tmem, ad, bd, idesc, !(since == 0 && pass == 0 && j == 0));
                }
            }
            tc::umma_commit(c.mbar);
        }
        c.slabs_done += 1;
        since += 1;
    }
    if (c.slabs_done > 0) tc::mbar_wait(c.mbar, (uint32_t)((c.slabs_done - 1) & 1));
    tc::tc_fence_after();
}

// TMEM accumulator (lane = m, column = n) -> shared memory float[128][DPITCH] (overwrites the operand buffers)
template <bool FLUSH>
__device__ __forceinline__ float* drain(Ctx& c, bool has_data) {
    float* d = reinterpret_cast<float*>(c.dsm);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int m = 32 * (warp & 3) + lane;
    const int cbase = (warp >> 2) * 64;                  // warps 0-3: columns 0-63, warps 4-7: 64-127
    float v[32];
#pragma unroll
    for (int h = 0; h < 2; ++h) {
        if (has_data) {
            tc::tmem_ld32(c.tmem + ((uint32_t)(32 * (warp & 3)) << 16) + cbase + 32 * h, v);
            if (FLUSH && c.flushed) {
#pragma unroll
                for (int i = 0; i < 32; ++i) v[i] += c.acc[32 * h + i];
            }
        } else {
#pragma unroll
            for (int i = 0; i < 32; ++i) v[i] = 0.f;
        }
#pragma unroll
        for (int i = 0; i < 32; ++i) d[m * DPITCH + cbase + 32 * h + i] = v[i];
    }
    tc::tc_fence_before();
    __syncthreads();
    return d;
}

// ---------------------------------------------------------------------------
// High-precision variant ("hp") for products whose error is amplified downstream (the orthogonal transformations
// of the eigensolver).  Measured on B200: every tcgen05.mma accumulation into TMEM TRUNCATES at the accumulator's
// ulp (~1 ulp of systematic loss per MMA), and a tiny addend (the lo*hi / hi*lo correction terms, 2^-11 of the
// main term) is truncated at the BIG accumulator's ulp too.  A 60-MMA chain lost 4e-6 of orthogonality per GEMM.
//   * two accumulators: TMEM columns [0,128) take hi*hi only, [128,256) take the two cross terms, so the cross
//     terms are aligned to their own (small) magnitude; they are added once, in fp32 RN, at the drain;
//   * the main accumulator is flushed into fp32 registers (RN adds) every HP_FLUSH_SLABS slab(s) = 4 MMAs, which
//     also restarts the truncation at the magnitude of a 64-term partial sum instead of the running total.
// Same MMA count and staging as the standard core; ~5e-7 per GEMM like an fp32 FMA chain.
// ---------------------------------------------------------------------------
constexpr int HP_FLUSH_SLABS = 1;

__device__ __forceinline__ void setup_hp(Ctx& c, unsigned char* dsm, uint64_t* mbar, uint32_t* tmem_slot) {
    if (threadIdx.x == 0) { tc::mbar_init(mbar, 1); tc::fence_mbar_init(); }
    if ((threadIdx.x >> 5) == 0) tc::tmem_alloc<256>(tmem_slot);
    tc::tc_fence_before();
    __syncthreads();
    tc::tc_fence_after();
    c.dsm = dsm; c.mbar = mbar; c.tmem = *tmem_slot; c.slabs_done = 0; c.flushed = 0;
}
__device__ __forceinline__ void teardown_hp(Ctx& c) {
    tc::tc_fence_before();
    __syncthreads();
    if ((threadIdx.x >> 5) == 0) tc::tmem_dealloc<256>(c.tmem);
}

template <class LA, class LB>
__device__ __forceinline__ void mainloop_hp(const LA& la, const LB& lb, int m0, int n0, int kbeg, int kend,
                                            bool same_operand, Ctx& c) {
    unsigned char* a_hi = c.dsm;
    unsigned char* a_lo = c.dsm + PART;
    unsigned char* b_hi = same_operand ? a_hi : c.dsm + 2 * PART;
    unsigned char* b_lo = same_operand ? a_lo : c.dsm + 3 * PART;
    float4 ra[4], rb[4];
    const int nslab = (kend - kbeg + BKS - 1) / BKS;
    if (nslab > 0) { fetch(la, m0, kbeg, ra); if (!same_operand) fetch(lb, n0, kbeg, rb); }
    constexpr uint32_t idesc = tc::idesc_tf32(128, 128, 0, 0);
    int since = 0;                                           // slabs in the main accumulator since the last flush
    for (int s = 0; s < nslab; ++s) {
        if (c.slabs_done > 0) tc::mbar_wait(c.mbar, (uint32_t)((c.slabs_done - 1) & 1));   // smem free, MMAs complete
        if (since == HP_FLUSH_SLABS) { tc::tc_fence_after(); flush_add(c); since = 0; }
        stage<LA>(a_hi, a_lo, ra);
        if (!same_operand) stage<LB>(b_hi, b_lo, rb);
        if (s + 1 < nslab) {
            fetch(la, m0, kbeg + (s + 1) * BKS, ra);
            if (!same_operand) fetch(lb, n0, kbeg + (s + 1) * BKS, rb);
        }
        tc::fence_async_smem();
        __syncthreads();
        if (threadIdx.x == 0) {
            tc::tc_fence_after();
            const uint32_t ah = tc::smem_u32(a_hi), al = tc::smem_u32(a_lo);
            const uint32_t bh = tc::smem_u32(b_hi), bl = tc::smem_u32(b_lo);
            const bool first_cross = c.slabs_done == 0;
#pragma unroll
            for (int j = 0; j < BKS / 8; ++j)                 // hi * hi -> main accumulator
                tc::umma_tf32(c.tmem, tc::smem_desc(ah + 2 * j * LBO, LBO, SBO), tc::smem_desc(bh + 2 * j * LBO, LBO, SBO),
                              idesc, !(since == 0 && j == 0));
#pragma unroll
            for (int j = 0; j < BKS / 8; ++j)                 // lo * hi -> cross accumulator
                tc::umma_tf32(c.tmem + 128, tc::smem_desc(al + 2 * j * LBO, LBO, SBO), tc::smem_desc(bh + 2 * j * LBO, LBO, SBO),
                              idesc, !(first_cross && j == 0));
#pragma unroll
            for (int j = 0; j < BKS / 8; ++j)                 // hi * lo -> cross accumulator
                tc::umma_tf32(c.tmem + 128, tc::smem_desc(ah + 2 * j * LBO, LBO, SBO), tc::smem_desc(bl + 2 * j * LBO, LBO, SBO),
                              idesc, true);
            tc::umma_commit(c.mbar);
        }
        c.slabs_done += 1;
        since += 1;
    }
    if (c.slabs_done > 0) tc::mbar_wait(c.mbar, (uint32_t)((c.slabs_done - 1) & 1));
    tc::tc_fence_after();
}

// main (+ flushed registers) + cross -> shared memory float[128][DPITCH]
__device__ __forceinline__ float* drain_hp(Ctx& c, bool has_data) {
    float* d = reinterpret_cast<float*>(c.dsm);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int m = 32 * (warp & 3) + lane;
    const int cbase = (warp >> 2) * 64;
    float v[32], x[32];
#pragma unroll
    for (int h = 0; h < 2; ++h) {
        if (has_data) {
            const uint32_t ta = c.tmem + ((uint32_t)(32 * (warp & 3)) << 16) + cbase + 32 * h;
            tc::tmem_ld32(ta, v);
            tc::tmem_ld32(ta + 128, x);
#pragma unroll
            for (int i = 0; i < 32; ++i) v[i] = (c.flushed ? v[i] + c.acc[32 * h + i] : v[i]) + x[i];
        } else {
#pragma unroll
            for (int i = 0; i < 32; ++i) v[i] = 0.f;
        }
#pragma unroll
        for (int i = 0; i < 32; ++i) d[m * DPITCH + cbase + 32 * h + i] = v[i];
    }
    tc::tc_fence_before();
    __syncthreads();
    return d;
}

// Epilogues may declare themselves read-modify-write by providing `float* addr(m, n)` (nullptr = skip the element)
// and `float combine(old, value)` next to operator(): the tensor-core kernel then batches the loads ahead of the stores.
template <class E, class = void>
struct is_rmw_epilogue : std::false_type {};
template <class E>
struct is_rmw_epilogue<E, std::void_t<decltype(std::declval<const E&>().addr(0, 0))>> : std::true_type {};

// generic GEMM: C(m, n) = epi(m, n, sum_k A(m,k) B(n,k)); grid (tiles_n, tiles_m)
// FLUSH is chosen by the host from K: products with K <= 512 keep the chain short anyway and run the
// register-lean variant (3 CTAs / SM).
template <class LA, class LB, class Epi, bool FLUSH>
__global__ void __launch_bounds__(THREADS)
tc_gemm_kernel(LA la, LB lb, Epi epi, int M, int N, int K) {
    extern __shared__ __align__(128) unsigned char dsm_raw[];
    __shared__ __align__(8) uint64_t mbar;
    __shared__ uint32_t tmem_slot;
    Ctx c;
    setup(c, dsm_raw, &mbar, &tmem_slot);
    const int m0 = blockIdx.y * TILE, n0 = blockIdx.x * TILE;
    mainloop<FLUSH>(la, lb, m0, n0, 0, K, false, c);
    const float* d = drain<FLUSH>(c, K > 0);
    if constexpr (is_rmw_epilogue<Epi>::value) {
        // read-modify-write epilogue: the old values of a batch are requested BEFORE its stores.  Written element by
        // element (`W[i] += v`) every store has to wait for its own load and the loop pays one L2 round trip per
        // element (measured: ~40 us per 128x128 tile)
        constexpr int BATCH = 16;
        for (int e0 = threadIdx.x; e0 < TILE * TILE; e0 += THREADS * BATCH) {
            float* p[BATCH]; float old[BATCH];
#pragma unroll
            for (int b = 0; b < BATCH; ++b) {
                const int e = e0 + b * THREADS, m = e >> 7, n = e & 127;
                p[b] = (m0 + m < M && n0 + n < N) ? epi.addr(m0 + m, n0 + n) : nullptr;
                old[b] = p[b] ? *p[b] : 0.f;
            }
#pragma unroll
            for (int b = 0; b < BATCH; ++b) {
                const int e = e0 + b * THREADS;
                if (p[b]) *p[b] = epi.combine(old[b], d[(e >> 7) * DPITCH + (e & 127)]);
            }
        }
    } else {
        for (int e = threadIdx.x; e < TILE * TILE; e += THREADS) {
            const int m = e >> 7, n = e & 127;
            if (m0 + m < M && n0 + n < N) epi(m0 + m, n0 + n, d[m * DPITCH + n]);
        }
    }
    teardown(c);
}

template <class LA, class LB, class Epi, bool FLUSH>
static inline cudaError_t launch_variant(const LA& la, const LB& lb, const Epi& epi, int M, int N, int K, cudaStream_t st) {
    {   // per call: the attribute is per device, and one process may drive several GPUs
        cudaError_t e = cudaFuncSetAttribute(tc_gemm_kernel<LA, LB, Epi, FLUSH>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM);
        if (e != cudaSuccess) return e;
    }
    dim3 grid(cdiv(N, TILE), cdiv(M, TILE));
    tc_gemm_kernel<LA, LB, Epi, FLUSH><<<grid, THREADS, SMEM, st>>>(la, lb, epi, M, N, K);
    count_launch();
    return cudaGetLastError();
}
template <class LA, class LB, class Epi>
static inline cudaError_t launch(const LA& la, const LB& lb, const Epi& epi, int M, int N, int K, cudaStream_t st) {
    if (K > FLUSH_SLABS * BKS * 2) return launch_variant<LA, LB, Epi, true>(la, lb, epi, M, N, K, st);
    return launch_variant<LA, LB, Epi, false>(la, lb, epi, M, N, K, st);
}

}  // namespace tcg

bool use_tc_gemm();   // false when KFACB_GEMM_SIMT=1 (debug / A-B comparisons)

// dispatch: tensor cores unless the problem is tiny or the SIMT path is forced
template <class LA, class LB, class Epi>
static inline cudaError_t launch_gemm_auto(const LA& la, const LB& lb, const Epi& epi, int M, int N, int K,
                                           cudaStream_t st) {
    if (use_tc_gemm() && (long)M * N >= 64 * 64 && K >= 32) return tcg::launch(la, lb, epi, M, N, K, st);
    return launch_gemm(la, lb, epi, M, N, K, st);
}

}  // namespace kfacb
