// Standalone tcgen05 descriptor self-test (debug tool, not part of the library).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -I kfac_pytorch_b200/csrc -o /tmp/tc_selftest tools/tc_selftest.cu
// Runs D = A * B^T (A: MxK, B: NxK) through kind::tf32 UMMAs with operands staged in the
// no-swizzle core-matrix layouts used by eigh.cu, for K-major and MN-major operands and
// with the LBO/SBO roles straight or swapped, and prints the max error vs a CPU reference.
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <vector>
#include "tc.cuh"

using namespace kfacb;

// layout: mode 0 = K-major  : (mn&7)*16 + (mn>>3)*S1 + (k>>2)*S2 + (k&3)*4
//         mode 1 = MN-major : (mn&3)*4 + (k&7)*16 + (mn>>2)*S1 + (k>>3)*S2
__device__ __forceinline__ int elem_off(int mode, int mn, int k, int s1, int s2) {
    return mode == 0 ? (mn & 7) * 16 + (mn >> 3) * s1 + (k >> 2) * s2 + (k & 3) * 4
                     : (mn & 3) * 4 + (k & 7) * 16 + (mn >> 2) * s1 + (k >> 3) * s2;
}

template <int M, int N>
__global__ void test_kernel(const float* A, const float* B, float* D, int K, int mode, int swap,
                            int a_s1, int a_s2, int b_s1, int b_s2, int passes) {
    extern __shared__ __align__(128) unsigned char dsm[];
    __shared__ __align__(8) uint64_t mbar;
    __shared__ uint32_t tmem_slot;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    if (tid == 0) { tc::mbar_init(&mbar, 1); tc::fence_mbar_init(); }
    if (warp == 0) tc::tmem_alloc<128>(&tmem_slot);
    tc::tc_fence_before();
    __syncthreads();
    tc::tc_fence_after();
    const uint32_t tmem = tmem_slot;
    const int a_part = mode == 0 ? (K / 4) * a_s2 : (K / 8) * a_s2;
    const int b_part = mode == 0 ? (K / 4) * b_s2 : (K / 8) * b_s2;
    unsigned char* a_hi = dsm; unsigned char* a_lo = dsm + a_part;
    unsigned char* b_hi = dsm + 2 * a_part; unsigned char* b_lo = dsm + 2 * a_part + b_part;
    for (int e = tid; e < M * K; e += blockDim.x) {
        const int m = e / K, k = e % K;
        float hi, lo; tc::split_tf32(A[e], hi, lo);
        *reinterpret_cast<float*>(a_hi + elem_off(mode, m, k, a_s1, a_s2)) = hi;
        *reinterpret_cast<float*>(a_lo + elem_off(mode, m, k, a_s1, a_s2)) = lo;
    }
    for (int e = tid; e < N * K; e += blockDim.x) {
        const int n = e / K, k = e % K;
        float hi, lo; tc::split_tf32(B[e], hi, lo);
        *reinterpret_cast<float*>(b_hi + elem_off(mode, n, k, b_s1, b_s2)) = hi;
        *reinterpret_cast<float*>(b_lo + elem_off(mode, n, k, b_s1, b_s2)) = lo;
    }
    tc::fence_async_smem();
    __syncthreads();
    if (tid == 0) {
        tc::tc_fence_after();
        const uint32_t idesc = tc::idesc_tf32(M, N, mode, mode);
        for (int pass = 0; pass < passes; ++pass) {
            const uint32_t ab = tc::smem_u32(pass == 1 ? a_lo : a_hi);
            const uint32_t bb = tc::smem_u32(pass == 2 ? b_lo : b_hi);
            for (int j = 0; j < K / 8; ++j) {
                // K-major: s1 = MN-group stride (SBO), s2 = K-quad stride (LBO); step = 2 quads
                // MN-major: s1 = MN-group stride (SBO), s2 = K-group stride (LBO); step = 1 group
                const int astep = mode == 0 ? 2 * j * a_s2 : j * a_s2;
                const int bstep = mode == 0 ? 2 * j * b_s2 : j * b_s2;
                const uint64_t ad = swap ? tc::smem_desc(ab + astep, a_s1, a_s2) : tc::smem_desc(ab + astep, a_s2, a_s1);
                const uint64_t bd = swap ? tc::smem_desc(bb + bstep, b_s1, b_s2) : tc::smem_desc(bb + bstep, b_s2, b_s1);
                tc::umma_tf32(tmem, ad, bd, idesc, !(pass == 0 && j == 0));
            }
        }
        tc::umma_commit(&mbar);
    }
    tc::mbar_wait(&mbar, 0);
    tc::tc_fence_after();
    // dump all 128 lanes x N columns raw: D[lane][col]
    float v[32];
    for (int c0 = 0; c0 < N; c0 += 32) {
        tc::tmem_ld32(tmem + ((uint32_t)(32 * warp) << 16) + c0, v);
        for (int i = 0; i < 32; ++i) D[(32 * warp + lane) * N + c0 + i] = v[i];
    }
    tc::tc_fence_before();
    __syncthreads();
    if (warp == 0) tc::tmem_dealloc<128>(tmem);
}

template <int M, int N>
void run(int K, int mode, int swap, int passes) {
    std::vector<float> A(M * K), B(N * K), D(128 * N);
    srand(1);
    for (auto& x : A) x = (rand() / (float)RAND_MAX) * 2.f - 1.f;
    for (auto& x : B) x = (rand() / (float)RAND_MAX) * 2.f - 1.f;
    float *dA, *dB, *dD;
    cudaMalloc(&dA, A.size() * 4); cudaMalloc(&dB, B.size() * 4); cudaMalloc(&dD, D.size() * 4);
    cudaMemcpy(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice);
    cudaMemcpy(dB, B.data(), B.size() * 4, cudaMemcpyHostToDevice);
    cudaMemset(dD, 0, D.size() * 4);
    int a_s1, a_s2, b_s1, b_s2;
    if (mode == 0) { a_s1 = 128; a_s2 = (M / 8) * 128 + 16; b_s1 = 128; b_s2 = (N / 8) * 128 + 16; }
    else { a_s1 = 144; a_s2 = (M / 4) * 144; b_s1 = 144; b_s2 = (N / 4) * 144; }
    const int a_part = mode == 0 ? (K / 4) * a_s2 : (K / 8) * a_s2;
    const int b_part = mode == 0 ? (K / 4) * b_s2 : (K / 8) * b_s2;
    const int smem = 2 * a_part + 2 * b_part;
    cudaFuncSetAttribute(test_kernel<M, N>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    test_kernel<M, N><<<1, 128, smem>>>(dA, dB, dD, K, mode, swap, a_s1, a_s2, b_s1, b_s2, passes);
    cudaError_t e = cudaDeviceSynchronize();
    cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost);
    double maxerr = 0, maxref = 0;
    int bad = 0;
    for (int m = 0; m < M; ++m) {
        const int lane = (M == 64) ? (m & 15) + 32 * (m >> 4) : m;
        for (int n = 0; n < N; ++n) {
            double r = 0;
            for (int k = 0; k < K; ++k) r += (double)A[m * K + k] * B[n * K + k];
            const double d = D[lane * N + n];
            if (!(fabs(d - r) < 1e-2)) ++bad;
            maxerr = fmax(maxerr, fabs(d - r)); maxref = fmax(maxref, fabs(r));
        }
    }
    printf("M=%d N=%d K=%d mode=%s swap=%d passes=%d smem=%d: cuda=%s maxerr=%.3e (max|ref|=%.2f) bad=%d  D[0][0..3]=%g %g %g %g\n",
           M, N, K, mode ? "MN" : "K", swap, passes, smem, cudaGetErrorString(e), maxerr, maxref, bad,
           D[0], D[1], D[2], D[3]);
    cudaFree(dA); cudaFree(dB); cudaFree(dD);
    if (e != cudaSuccess) { cudaDeviceReset(); }
}

int main(int argc, char** argv) {
    const int v = argc > 1 ? atoi(argv[1]) : 0;
    switch (v) {
        case 0: run<64, 64>(64, 0, 0, 3); run<128, 64>(64, 0, 0, 1); run<128, 64>(64, 0, 0, 3); run<128, 128>(32, 0, 0, 3); break;
        case 1: run<128, 64>(8, 1, 0, 1); run<128, 64>(64, 1, 0, 3); break;
        case 2: run<128, 64>(8, 1, 1, 1); break;
        case 3: run<128, 64>(64, 1, 1, 3); break;
        case 4: run<64, 64>(8, 1, 0, 1); break;
    }
    return 0;
}
