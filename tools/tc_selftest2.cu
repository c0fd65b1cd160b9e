// Standalone tcgen05 experiments for the round-2 redesign of the Jacobi GEMM kernels (debug tool, not part of
// the library; every experiment is its own process run so that a faulting descriptor cannot poison the next).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -I kfac_pytorch_b200/csrc -o /tmp/tc_selftest2 tools/tc_selftest2.cu
//   for e in 1 2 3 4; do /tmp/tc_selftest2 $e; done
// Questions (D = A * B^T, A: 128xK, B: 128xK, kind::tf32, fp32 accumulate):
//   1  raw fp32 bits as a tf32 operand: does the tensor core TRUNCATE or ROUND the low 13 mantissa bits?
//      (decides whether a TMA-loaded raw tile can serve as the "hi" part of the 3xTF32 split, lo = x - trunc(x))
//   2  K-major SWIZZLE_128B operands (the layout a TMA box of 32 floats x rows produces): descriptor fields
//   3  A operand from TENSOR MEMORY (tcgen05.st 32x32b -> tcgen05.mma [d], [a], bdesc): would let the 128x128
//      rotation block U live in TMEM (256 columns hi+lo) instead of 128 KB of shared memory
//   4  MN-major SWIZZLE_128B operands (rows of X as they lie in memory, no register transposition)
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cmath>
#include <vector>
#include "tc.cuh"

using namespace kfacb;

constexpr int M = 128, N = 128;

__device__ __forceinline__ uint64_t desc_sw(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes, uint32_t layout) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3FFF);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)(layout & 7) << 61;            // 0 none, 1 128B(base 32B), 2 128B, 4 64B, 6 32B
    return d;
}
__device__ __forceinline__ void umma_tf32_tmemA(uint32_t d_tmem, uint32_t a_tmem, uint64_t bdesc, uint32_t idesc,
                                                bool accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t"
        "}\n" ::"r"(d_tmem), "r"(a_tmem), "l"(bdesc), "r"(idesc), "r"((uint32_t)accumulate)
        : "memory");
}
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const float (&v)[8]) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(taddr),
                 "r"(__float_as_uint(v[0])), "r"(__float_as_uint(v[1])), "r"(__float_as_uint(v[2])),
                 "r"(__float_as_uint(v[3])), "r"(__float_as_uint(v[4])), "r"(__float_as_uint(v[5])),
                 "r"(__float_as_uint(v[6])), "r"(__float_as_uint(v[7]))
                 : "memory");
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
}

// byte offset of element (mn, k) for the layouts under test
//   0: K-major no swizzle (library layout)    (mn&7)*16 + (mn>>3)*128 + (k>>2)*LBO0 + (k&3)*4
//   1: K-major SW128: 8-row x 128 B atoms      (mn>>3)*1024 + (mn&7)*128 + (((k>>2)&7) ^ (mn&7))*16 + (k&3)*4 + (k>>5)*kblk
//   2: MN-major SW128: 8 k-rows x 128 B atoms  (mn>>5)*mnblk + (k>>3)*1024 + (k&7)*128 + ((((mn>>2)&7) ^ (k&7))*16) + (mn&3)*4
__device__ __forceinline__ int off_of(int layout, int mn, int k, int K) {
    if (layout == 0) return (mn & 7) * 16 + (mn >> 3) * 128 + (k >> 2) * ((128 / 8) * 128 + 16) + (k & 3) * 4;
    if (layout == 1) return (mn >> 3) * 1024 + (mn & 7) * 128 + ((((k >> 2) & 7) ^ (mn & 7)) * 16) + (k & 3) * 4 + (k >> 5) * (128 * 128);
    return (mn >> 5) * (K * 128) + (k >> 3) * 1024 + (k & 7) * 128 + ((((mn >> 2) & 7) ^ (k & 7)) * 16) + (mn & 3) * 4;
}

// exp: 1 raw operands (layout 0) | 2 K-major SW128 | 3 A from TMEM (B layout 0) | 4 MN-major SW128
// var: descriptor variant of the experiment
__global__ void test_kernel(const float* A, const float* B, float* D, int K, int exp, int var) {
    extern __shared__ __align__(1024) unsigned char dsm[];
    __shared__ __align__(8) uint64_t mbar;
    __shared__ uint32_t tmem_slot;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    if (tid == 0) { tc::mbar_init(&mbar, 1); tc::fence_mbar_init(); }
    if (warp == 0) tc::tmem_alloc<256>(&tmem_slot);
    tc::tc_fence_before();
    __syncthreads();
    tc::tc_fence_after();
    const uint32_t tmem = tmem_slot;
    const int layout = exp == 2 ? 1 : exp == 4 ? 2 : 0;
    const int part = 128 * K * 4 + 4096;                    // generous: one operand
    unsigned char* a_s = dsm;
    unsigned char* b_s = dsm + ((part + 1023) & ~1023);
    for (int e = tid; e < M * K; e += blockDim.x) {
        const int m = e / K, k = e % K;
        *reinterpret_cast<float*>(a_s + off_of(layout, m, k, K)) = A[e];
        *reinterpret_cast<float*>(b_s + off_of(layout, m, k, K)) = B[e];
    }
    if (exp == 3) {                                          // A rows -> TMEM columns [128, 128 + K)
        for (int k0 = 0; k0 < K; k0 += 8) {
            float v[8];
            for (int i = 0; i < 8; ++i) v[i] = A[(32 * warp + lane) * K + k0 + i];
            tmem_st8(tmem + ((uint32_t)(32 * warp) << 16) + 128 + k0, v);
        }
        tc::tc_fence_before();
    }
    tc::fence_async_smem();
    __syncthreads();
    if (tid == 0) {
        tc::tc_fence_after();
        const int mn_major = exp == 4 ? 1 : 0;
        const uint32_t idesc = tc::idesc_tf32(M, N, mn_major, mn_major);
        const uint32_t ab = tc::smem_u32(a_s), bb = tc::smem_u32(b_s);
        constexpr int LBO0 = (128 / 8) * 128 + 16;
        for (int j = 0; j < K / 8; ++j) {
            uint64_t ad, bd;
            if (layout == 0) {
                ad = tc::smem_desc(ab + 2 * j * LBO0, LBO0, 128);
                bd = tc::smem_desc(bb + 2 * j * LBO0, LBO0, 128);
            } else if (layout == 1) {
                // K step of 8 floats = 32 B inside the 128-B swizzle row; a new 32-float block every 4 steps
                const uint32_t step = (j & 3) * 32 + (j >> 2) * (128 * 128);
                const uint32_t lbo = var == 0 ? 0 : var == 1 ? 16 : 1024;
                ad = desc_sw(ab + step, lbo, 1024, 2);
                bd = desc_sw(bb + step, lbo, 1024, 2);
            } else {
                // one UMMA = one 8-k-row group (1024 B atoms); MN blocks of 32 elements are K*128 B apart
                const uint32_t step = j * 1024;
                const uint32_t mnblk = K * 128;
                const uint32_t lbo = (var & 1) ? 1024 : mnblk, sbo = (var & 1) ? mnblk : 1024;
                ad = desc_sw(ab + step, lbo, sbo, 2);
                bd = desc_sw(bb + step, lbo, sbo, 2);
            }
            if (exp == 3) umma_tf32_tmemA(tmem, tmem + 128 + 8 * j, bd, idesc, j != 0);
            else tc::umma_tf32(tmem, ad, bd, idesc, j != 0);
        }
        tc::umma_commit(&mbar);
    }
    tc::mbar_wait(&mbar, 0);
    tc::tc_fence_after();
    float v[32];
    for (int c0 = 0; c0 < N; c0 += 32) {
        tc::tmem_ld32(tmem + ((uint32_t)(32 * warp) << 16) + c0, v);
        for (int i = 0; i < 32; ++i) D[(32 * warp + lane) * N + c0 + i] = v[i];
    }
    tc::tc_fence_before();
    __syncthreads();
    if (warp == 0) tc::tmem_dealloc<256>(tmem);
}

static float trunc_tf32(float x) { uint32_t u; memcpy(&u, &x, 4); u &= 0xFFFFE000u; memcpy(&x, &u, 4); return x; }
static float round_tf32(float x) { uint32_t u; memcpy(&u, &x, 4); u = (u + 0x1000u) & 0xFFFFE000u; memcpy(&x, &u, 4); return x; }

static void run(int K, int exp, int var) {
    std::vector<float> A(M * K), B(N * K), D(128 * N);
    srand(1);
    for (auto& x : A) x = (rand() / (float)RAND_MAX) * 2.f - 1.f;
    for (auto& x : B) x = (rand() / (float)RAND_MAX) * 2.f - 1.f;
    if (exp != 1) {                       // experiments 2-4 check layouts: make the operands exact in tf32
        for (auto& x : A) x = trunc_tf32(x);
        for (auto& x : B) x = trunc_tf32(x);
    }
    float *dA, *dB, *dD;
    cudaMalloc(&dA, A.size() * 4); cudaMalloc(&dB, B.size() * 4); cudaMalloc(&dD, D.size() * 4);
    cudaMemcpy(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice);
    cudaMemcpy(dB, B.data(), B.size() * 4, cudaMemcpyHostToDevice);
    cudaMemset(dD, 0, D.size() * 4);
    const int smem = 2 * (((128 * K * 4 + 4096) + 1023) & ~1023) + 1024;
    cudaFuncSetAttribute(test_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    test_kernel<<<1, 128, smem>>>(dA, dB, dD, K, exp, var);
    cudaError_t e = cudaDeviceSynchronize();
    cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost);
    double err_exact = 0, err_trunc = 0, err_round = 0, maxref = 0;
    for (int m = 0; m < M; ++m)
        for (int n = 0; n < N; ++n) {
            double r = 0, rt = 0, rr = 0;
            for (int k = 0; k < K; ++k) {
                r += (double)A[m * K + k] * B[n * K + k];
                rt += (double)trunc_tf32(A[m * K + k]) * trunc_tf32(B[n * K + k]);
                rr += (double)round_tf32(A[m * K + k]) * round_tf32(B[n * K + k]);
            }
            const double d = D[m * N + n];
            err_exact = fmax(err_exact, fabs(d - r)); err_trunc = fmax(err_trunc, fabs(d - rt));
            err_round = fmax(err_round, fabs(d - rr)); maxref = fmax(maxref, fabs(r));
        }
    printf("exp=%d var=%d K=%d: cuda=%s  max|D-exact|=%.3e  max|D-trunc(tf32)|=%.3e  max|D-round(tf32)|=%.3e  (max|ref|=%.2f)  D[0][0..3]=%g %g %g %g\n",
           exp, var, K, cudaGetErrorString(e), err_exact, err_trunc, err_round, maxref, D[0], D[1], D[2], D[3]);
    cudaFree(dA); cudaFree(dB); cudaFree(dD);
}

int main(int argc, char** argv) {
    const int exp = argc > 1 ? atoi(argv[1]) : 1;
    const int var = argc > 2 ? atoi(argv[2]) : 0;
    const int K = argc > 3 ? atoi(argv[3]) : 32;
    run(K, exp, var);
    return 0;
}
