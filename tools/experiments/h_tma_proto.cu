// PROTOTYPE for round 2 (standalone; not part of the library): the projected-block kernel of the Jacobi solver
// (H = V_R W_R^T for a pair of 64-row blocks, split-K) fed by TMA instead of LDG -> split -> STS.
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I kfac_pytorch_b200/csrc -I include \
//        -o tools/_bin/h_tma_proto tools/experiments/h_tma_proto.cu && tools/_bin/h_tma_proto [kchunk] [stages] [persistent]
//
// Why: bj_h_tc_kernel (eigh.cu) reaches 2.5-3.0 TB/s (57-69 us for one round of a 4608-dim factor, 170 MB);
// tools/hbm_pattern_bench.cu shows the memory system gives > 5.5 TB/s on row-strided segments.  The kernel is bound
// by its per-slab chain (wait MMA -> stage -> barrier -> issue) on a single operand buffer.  Facts this design
// uses (measured with tools/tc_selftest2.cu, profiles/r01_tc_selftest2.txt):
//   * rows of X are K-contiguous: a TMA box of 32 floats x 64 rows with SWIZZLE_128B lands exactly in the
//     K-major SW128 operand layout (descriptor: SBO 1024, layout 2, +32 B per UMMA K step);
//   * the tensor core TRUNCATES a raw fp32 word to tf32, so the raw tile IS the `hi` operand; the `lo` operand
//     is lo = rna_tf32(x - trunc(x)), one elementwise pass by 4 converter warps into a second buffer with the
//     same offsets (3xTF32: hi*hi + lo*hi + hi*lo; representation error 2^-21 instead of 2^-23 -- fine for H, which
//     only steers the rotations).
// Roles (192 threads): warp 0 = TMA producer (one thread), warp 1 = UMMA issuer (one thread), warps 2-5 =
// converters + epilogue.  STAGES-deep ring, mbarriers: full (TMA bytes landed) -> conv (128 converter arrivals)
// -> empty (tcgen05.commit) -> producer.
// The program checks two pairs against a double-precision CPU product and times one "round" (all 36 pairs).
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cmath>
#include <vector>
#include <cuda.h>
#include <cuda_runtime.h>
#include "tc.cuh"

using namespace kfacb;

constexpr int BK = 32;                       // floats per TMA box row = 128 B (the SWIZZLE_128B span)
constexpr int OPER = 128 * BK * 4;           // 16384 B: one 128-row operand slab (two 64-row boxes)
constexpr int STAGE = 4 * OPER;              // [V raw | W raw | V lo | W lo]
constexpr int MAX_STAGES = 3;
constexpr int THREADS = 192;

__device__ __forceinline__ uint64_t desc_sw128(uint32_t saddr) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3FFF);
    d |= (uint64_t)(1024 >> 4) << 32;             // SBO: 8-row groups are 1024 B apart (LBO is ignored)
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;                       // SWIZZLE_128B
    return d;
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(tc::smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tma_load_2d(void* smem_dst, const CUtensorMap* tm, int c_inner, int c_outer, uint64_t* bar) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(
            tc::smem_u32(smem_dst)),
        "l"(reinterpret_cast<uint64_t>(tm)), "r"(c_inner), "r"(c_outer), "r"(tc::smem_u32(bar))
        : "memory");
}

// lo = rna_tf32(x - trunc_tf32(x))
__device__ __forceinline__ float lo_part(float x) {
    const float t = __uint_as_float(__float_as_uint(x) & 0xFFFFE000u);
    uint32_t r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x - t));
    return __uint_as_float(r);
}

// grid (ksplit, pairs); pair p works on row blocks 2p and 2p+1 (the prototype uses a fixed pairing)
__global__ void __launch_bounds__(THREADS)
h_tma_kernel(const __grid_constant__ CUtensorMap tm, float* __restrict__ partial, int n_pad, int kchunk, int stages) {
    extern __shared__ __align__(1024) unsigned char dsm_raw[];
    unsigned char* dsm = dsm_raw + ((1024u - (tc::smem_u32(dsm_raw) & 1023u)) & 1023u);    // SWIZZLE_128B: 1024-B aligned boxes
    __shared__ __align__(8) uint64_t bar_full[MAX_STAGES], bar_conv[MAX_STAGES], bar_empty[MAX_STAGES], bar_acc;
    __shared__ uint32_t tmem_slot;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int pair = blockIdx.y, z = blockIdx.x;
    const int kbeg = z * kchunk;
    if (kbeg >= n_pad) return;
    const int nst = (min(kchunk, n_pad - kbeg) + BK - 1) / BK;
    if (tid == 0) {
        for (int s = 0; s < stages; ++s) { tc::mbar_init(&bar_full[s], 1); tc::mbar_init(&bar_conv[s], 128); tc::mbar_init(&bar_empty[s], 1); }
        tc::mbar_init(&bar_acc, 1);
        tc::fence_mbar_init();
    }
    if (warp == 2) tc::tmem_alloc<128>(&tmem_slot);
    tc::tc_fence_before();
    __syncthreads();
    tc::tc_fence_after();
    const uint32_t tmem = tmem_slot;
    const int row_a = (2 * pair) * 64, row_b = (2 * pair + 1) * 64;

    if (warp == 0) {
        if (lane == 0) {
            for (int i = 0; i < nst; ++i) {
                const int s = i % stages, u = i / stages;
                if (u > 0) tc::mbar_wait(&bar_empty[s], (uint32_t)((u - 1) & 1));
                unsigned char* st = dsm + (size_t)s * STAGE;
                mbar_expect_tx(&bar_full[s], 2 * OPER);
                const int k = kbeg + i * BK;
                tma_load_2d(st, &tm, n_pad + k, row_a, &bar_full[s]);                 // V rows of block a
                tma_load_2d(st + OPER / 2, &tm, n_pad + k, row_b, &bar_full[s]);      // V rows of block b
                tma_load_2d(st + OPER, &tm, k, row_a, &bar_full[s]);                  // W rows of block a
                tma_load_2d(st + OPER + OPER / 2, &tm, k, row_b, &bar_full[s]);       // W rows of block b
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            constexpr uint32_t idesc = tc::idesc_tf32(128, 128, 0, 0);
            for (int i = 0; i < nst; ++i) {
                const int s = i % stages, u = i / stages;
                tc::mbar_wait(&bar_conv[s], (uint32_t)(u & 1));
                tc::tc_fence_after();
                const uint32_t base = tc::smem_u32(dsm + (size_t)s * STAGE);
                const uint32_t vh = base, wh = base + OPER, vl = base + 2 * OPER, wl = base + 3 * OPER;
#pragma unroll
                for (int pass = 0; pass < 3; ++pass) {
                    const uint32_t ab = pass == 1 ? vl : vh;
                    const uint32_t bb = pass == 2 ? wl : wh;
#pragma unroll
                    for (int j = 0; j < BK / 8; ++j)
                        tc::umma_tf32(tmem, desc_sw128(ab + 32 * j), desc_sw128(bb + 32 * j), idesc, !(i == 0 && pass == 0 && j == 0));
                }
                tc::umma_commit(&bar_empty[s]);
            }
            tc::umma_commit(&bar_acc);
        }
    } else {
        // converters: lo = rna_tf32(x - trunc_tf32(x)) for both operands, same (swizzled) offsets
        const int ct = tid - 64;                                  // 0..127
        for (int i = 0; i < nst; ++i) {
            const int s = i % stages, u = i / stages;
            tc::mbar_wait(&bar_full[s], (uint32_t)(u & 1));
            const float4* raw = reinterpret_cast<const float4*>(dsm + (size_t)s * STAGE);
            float4* lo = reinterpret_cast<float4*>(dsm + (size_t)s * STAGE + 2 * OPER);
#pragma unroll 4
            for (int c = ct; c < 2 * OPER / 16; c += 128) {
                const float4 x = raw[c];
                float4 l;
                l.x = lo_part(x.x); l.y = lo_part(x.y); l.z = lo_part(x.z); l.w = lo_part(x.w);
                lo[c] = l;
            }
            tc::fence_async_smem();
            tc::mbar_arrive(&bar_conv[s]);
        }
        // epilogue: accumulator -> partial[pair][z][128][128]
        tc::mbar_wait(&bar_acc, 0);
        tc::tc_fence_after();
        const int w4 = warp & 3;                                  // TMEM lane quarter this warp may read
        const int m = 32 * w4 + lane;
        float* dst = partial + ((size_t)pair * gridDim.x + z) * (128 * 128) + (size_t)m * 128;
        float v[32];
        for (int c0 = 0; c0 < 128; c0 += 32) {
            tc::tmem_ld32(tmem + ((uint32_t)(32 * w4) << 16) + c0, v);
#pragma unroll
            for (int q = 0; q < 32; q += 4) *reinterpret_cast<float4*>(dst + c0 + q) = make_float4(v[q], v[q + 1], v[q + 2], v[q + 3]);
        }
    }
    tc::tc_fence_before();
    __syncthreads();
    if (warp == 2) tc::tmem_dealloc<128>(tmem);
}

// ---------------------------------------------------------------------------------------------------------------
// PERSISTENT variant: one CTA per SM walks the (pair, k-chunk) work items round-robin.  The operand ring runs
// across items, the accumulator is double buffered in TMEM (2 x 128 columns) and a separate epilogue warp group
// drains item n while the mainloop of item n+1 is running: set-up, pipeline fill and the 64 KB store are paid once
// per CTA instead of once per item, and there is no wave quantisation (432 items on 148 SMs = 2.9 waves above).
// Roles (320 threads): warp 0 TMA producer, warp 1 UMMA issuer, warps 2-5 converters, warps 6-9 epilogue.
// ---------------------------------------------------------------------------------------------------------------
constexpr int P_THREADS = 320;
__global__ void __launch_bounds__(P_THREADS)
h_tma_persistent_kernel(const __grid_constant__ CUtensorMap tm, float* __restrict__ partial, int n_pad, int kchunk, int stages,
                        int ksplit, int nitems) {
    extern __shared__ __align__(1024) unsigned char dsm_raw[];
    unsigned char* dsm = dsm_raw + ((1024u - (tc::smem_u32(dsm_raw) & 1023u)) & 1023u);
    __shared__ __align__(8) uint64_t bar_full[MAX_STAGES], bar_conv[MAX_STAGES], bar_empty[MAX_STAGES], acc_full[2], acc_empty[2];
    __shared__ uint32_t tmem_slot;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    if (tid == 0) {
        for (int s = 0; s < stages; ++s) { tc::mbar_init(&bar_full[s], 1); tc::mbar_init(&bar_conv[s], 128); tc::mbar_init(&bar_empty[s], 1); }
        for (int b = 0; b < 2; ++b) { tc::mbar_init(&acc_full[b], 1); tc::mbar_init(&acc_empty[b], 128); }
        tc::fence_mbar_init();
    }
    if (warp == 2) tc::tmem_alloc<256>(&tmem_slot);
    tc::tc_fence_before();
    __syncthreads();
    tc::tc_fence_after();
    const uint32_t tmem = tmem_slot;
    auto item_stages = [&](int it) { const int z = it % ksplit; return (min(kchunk, n_pad - z * kchunk) + BK - 1) / BK; };

    if (warp == 0) {
        if (lane == 0) {
            int g = 0;                                                   // ring position, continuous across items
            for (int it = blockIdx.x; it < nitems; it += gridDim.x) {
                const int pair = it / ksplit, z = it % ksplit, nst = item_stages(it);
                const int row_a = (2 * pair) * 64, row_b = (2 * pair + 1) * 64;
                for (int i = 0; i < nst; ++i, ++g) {
                    const int s = g % stages, u = g / stages;
                    if (u > 0) tc::mbar_wait(&bar_empty[s], (uint32_t)((u - 1) & 1));
                    unsigned char* st = dsm + (size_t)s * STAGE;
                    mbar_expect_tx(&bar_full[s], 2 * OPER);
                    const int k = z * kchunk + i * BK;
                    tma_load_2d(st, &tm, n_pad + k, row_a, &bar_full[s]);
                    tma_load_2d(st + OPER / 2, &tm, n_pad + k, row_b, &bar_full[s]);
                    tma_load_2d(st + OPER, &tm, k, row_a, &bar_full[s]);
                    tma_load_2d(st + OPER + OPER / 2, &tm, k, row_b, &bar_full[s]);
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            constexpr uint32_t idesc = tc::idesc_tf32(128, 128, 0, 0);
            int g = 0, n = 0;
            for (int it = blockIdx.x; it < nitems; it += gridDim.x, ++n) {
                const int nst = item_stages(it), ab = n & 1;
                if (n >= 2) tc::mbar_wait(&acc_empty[ab], (uint32_t)(((n >> 1) - 1) & 1));    // epilogue drained this buffer
                for (int i = 0; i < nst; ++i, ++g) {
                    const int s = g % stages, u = g / stages;
                    tc::mbar_wait(&bar_conv[s], (uint32_t)(u & 1));
                    tc::tc_fence_after();
                    const uint32_t base = tc::smem_u32(dsm + (size_t)s * STAGE);
                    const uint32_t vh = base, wh = base + OPER, vl = base + 2 * OPER, wl = base + 3 * OPER;
#pragma unroll
                    for (int pass = 0; pass < 3; ++pass) {
                        const uint32_t a = pass == 1 ? vl : vh;
                        const uint32_t b = pass == 2 ? wl : wh;
#pragma unroll
                        for (int j = 0; j < BK / 8; ++j)
                            tc::umma_tf32(tmem + 128 * ab, desc_sw128(a + 32 * j), desc_sw128(b + 32 * j), idesc,
                                          !(i == 0 && pass == 0 && j == 0));
                    }
                    tc::umma_commit(&bar_empty[s]);
                }
                tc::umma_commit(&acc_full[ab]);
            }
        }
    } else if (warp < 6) {
        const int ct = tid - 64;
        int g = 0;
        for (int it = blockIdx.x; it < nitems; it += gridDim.x) {
            const int nst = item_stages(it);
            for (int i = 0; i < nst; ++i, ++g) {
                const int s = g % stages, u = g / stages;
                tc::mbar_wait(&bar_full[s], (uint32_t)(u & 1));
                const float4* raw = reinterpret_cast<const float4*>(dsm + (size_t)s * STAGE);
                float4* lo = reinterpret_cast<float4*>(dsm + (size_t)s * STAGE + 2 * OPER);
#pragma unroll 4
                for (int c = ct; c < 2 * OPER / 16; c += 128) {
                    const float4 x = raw[c];
                    float4 l;
                    l.x = lo_part(x.x); l.y = lo_part(x.y); l.z = lo_part(x.z); l.w = lo_part(x.w);
                    lo[c] = l;
                }
                tc::fence_async_smem();
                tc::mbar_arrive(&bar_conv[s]);
            }
        }
    } else {
        const int w4 = warp & 3, m = 32 * w4 + lane;
        int n = 0;
        for (int it = blockIdx.x; it < nitems; it += gridDim.x, ++n) {
            const int ab = n & 1;
            tc::mbar_wait(&acc_full[ab], (uint32_t)((n >> 1) & 1));
            tc::tc_fence_after();
            float* dst = partial + (size_t)it * (128 * 128) + (size_t)m * 128;      // it = pair * ksplit + z
            float v[32];
            for (int c0 = 0; c0 < 128; c0 += 32) {
                tc::tmem_ld32(tmem + ((uint32_t)(32 * w4) << 16) + 128 * ab + c0, v);
#pragma unroll
                for (int q = 0; q < 32; q += 4) *reinterpret_cast<float4*>(dst + c0 + q) = make_float4(v[q], v[q + 1], v[q + 2], v[q + 3]);
            }
            tc::tc_fence_before();
            tc::mbar_arrive(&acc_empty[ab]);
        }
    }
    tc::tc_fence_before();
    __syncthreads();
    if (warp == 2) tc::tmem_dealloc<256>(tmem);
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main(int argc, char** argv) {
    const int n_pad = 4608;
    const int kchunk = argc > 1 ? atoi(argv[1]) : 512;
    const int stages = argc > 2 ? atoi(argv[2]) : 3;
    const int persistent = argc > 3 ? atoi(argv[3]) : 0;       // 1: one CTA per SM walking the work items
    const long ld = 2L * n_pad;
    const int pairs = n_pad / 128, ksplit = (n_pad + kchunk - 1) / kchunk;
    std::vector<float> X((size_t)n_pad * ld);
    srand(3);
    for (auto& x : X) x = (rand() / (float)RAND_MAX) * 2.f - 1.f;
    float *dX, *dP;
    cudaMalloc(&dX, X.size() * 4);
    cudaMalloc(&dP, (size_t)pairs * ksplit * 128 * 128 * 4);
    cudaMemcpy(dX, X.data(), X.size() * 4, cudaMemcpyHostToDevice);

    EncodeTiledFn encode = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", (void**)&encode, cudaEnableDefault, &qres) != cudaSuccess || !encode) {
        printf("cuTensorMapEncodeTiled not available\n");
        return 1;
    }
    CUtensorMap tm;
    const cuuint64_t dims[2] = {(cuuint64_t)ld, (cuuint64_t)n_pad};
    const cuuint64_t strides[1] = {(cuuint64_t)ld * 4};
    const cuuint32_t box[2] = {BK, 64};
    const cuuint32_t estr[2] = {1, 1};
    const CUresult cr = encode(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, dX, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                               CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (cr != CUDA_SUCCESS) { printf("cuTensorMapEncodeTiled failed: %d\n", (int)cr); return 1; }

    const int smem = stages * STAGE + 1024;
    cudaFuncSetAttribute(h_tma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    cudaFuncSetAttribute(h_tma_persistent_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    const dim3 grid(ksplit, pairs);
    int sms = 148;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    const int nitems = pairs * ksplit;
    auto launch = [&] {
        if (persistent) h_tma_persistent_kernel<<<nitems < sms ? nitems : sms, P_THREADS, smem>>>(tm, dP, n_pad, kchunk, stages, ksplit, nitems);
        else h_tma_kernel<<<grid, THREADS, smem>>>(tm, dP, n_pad, kchunk, stages);
    };
    launch();
    cudaError_t e = cudaDeviceSynchronize();
    printf("%s kchunk=%d stages=%d smem=%d items=%d: %s\n", persistent ? "PERSISTENT" : "one CTA per item", kchunk, stages, smem,
           nitems, cudaGetErrorString(e));
    if (e != cudaSuccess) return 1;

    std::vector<float> P((size_t)pairs * ksplit * 128 * 128);
    cudaMemcpy(P.data(), dP, P.size() * 4, cudaMemcpyDeviceToHost);
    for (int pair : {0, 17}) {
        double maxerr = 0, maxref = 0;
        for (int m = 0; m < 128; m += 5)
            for (int n = 0; n < 128; n += 3) {
                const long rm = (m < 64 ? 2L * pair * 64 + m : (2L * pair + 1) * 64 + (m - 64));
                const long rn = (n < 64 ? 2L * pair * 64 + n : (2L * pair + 1) * 64 + (n - 64));
                double r = 0;
                for (int k = 0; k < n_pad; ++k) r += (double)X[rm * ld + n_pad + k] * X[rn * ld + k];
                double d = 0;
                for (int z = 0; z < ksplit; ++z) d += P[((size_t)pair * ksplit + z) * 128 * 128 + (size_t)m * 128 + n];
                maxerr = fmax(maxerr, fabs(d - r)); maxref = fmax(maxref, fabs(r));
            }
        printf("pair %2d: max|H - ref| = %.3e  (max|ref| = %.2f)  %s\n", pair, maxerr, maxref, maxerr < 2e-3 ? "OK" : "MISMATCH");
    }
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int i = 0; i < 3; ++i) launch();
    cudaEventRecord(e0);
    const int reps = 20;
    for (int i = 0; i < reps; ++i) launch();
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    const double bytes = 2.0 * n_pad * (double)n_pad * 4;       // V and W halves of every row, once
    printf("one round (36 pairs): %.1f us  = %.0f GB/s algorithmic   (bj_h_tc_kernel: 57-69 us)\n", ms / reps * 1e3,
           bytes / (ms / reps * 1e-3) / 1e9);
    return 0;
}
