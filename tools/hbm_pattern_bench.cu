// What can HBM deliver for the access pattern of the Jacobi row-update kernel?  (debug tool, not part of the library)
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/hbm_pattern_bench tools/hbm_pattern_bench.cu && /tmp/hbm_pattern_bench
// One Jacobi round of a 4608-dim factor reads and rewrites, in place, every row of the 4608 x 9216 fp32 work
// array: a CTA owns the 128 rows of a block pair and a range of 128-column tiles, i.e. 128 segments of 512 B that
// lie 36 KB apart.  bj_update_tc_kernel reaches ~2.3-3.1 TB/s on it.  This program measures plain SIMT kernels with
// the SAME pattern but no math, at several segment widths / residencies, plus a contiguous copy, to separate
// "pattern limit" from "kernel limit".
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

constexpr int NP = 4608;
constexpr long LD = 2L * NP;

// each thread owns VEC consecutive columns of a SEG-byte row segment; a CTA sweeps the 128 rows of its pair,
// UNROLL rows in flight per thread, for TILES consecutive segments
template <int VEC, int UNROLL>
__global__ void pattern_kernel(float* X, int seg_floats, int tiles) {
    const int pair = blockIdx.y;
    const int threads_per_row = seg_floats / VEC;
    const int rows_per_pass = blockDim.x / threads_per_row;
    const int tr = threadIdx.x / threads_per_row, tc = threadIdx.x % threads_per_row;
    for (int t = 0; t < tiles; ++t) {
        const long col = ((long)blockIdx.x * tiles + t) * seg_floats + tc * VEC;
        if (col >= LD) return;
        for (int r0 = tr; r0 < 128; r0 += rows_per_pass * UNROLL) {
            float v[UNROLL][VEC];
#pragma unroll
            for (int u = 0; u < UNROLL; ++u) {
                const int r = r0 + u * rows_per_pass;
                const int k = r < 128 ? r : 127;
                const long grow = (k < 64 ? 2L * pair * 64 + k : (2L * pair + 1) * 64 + (k - 64));
                const float* p = X + grow * LD + col;
                if (VEC == 4) { const float4 q = *reinterpret_cast<const float4*>(p); v[u][0] = q.x; v[u][1 % VEC] = q.y; v[u][2 % VEC] = q.z; v[u][3 % VEC] = q.w; }
                else v[u][0] = *p;
            }
#pragma unroll
            for (int u = 0; u < UNROLL; ++u) {
                const int r = r0 + u * rows_per_pass;
                if (r >= 128) continue;
                const long grow = (r < 64 ? 2L * pair * 64 + r : (2L * pair + 1) * 64 + (r - 64));
                float* p = X + grow * LD + col;
                if (VEC == 4) *reinterpret_cast<float4*>(p) = make_float4(v[u][0] + 1.f, v[u][1 % VEC] + 1.f, v[u][2 % VEC] + 1.f, v[u][3 % VEC] + 1.f);
                else *p = v[u][0] + 1.f;
            }
        }
    }
}

__global__ void copy_kernel(const float4* __restrict__ src, float4* __restrict__ dst, long n4) {
    for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < n4; i += (long)gridDim.x * blockDim.x) {
        float4 v = src[i];
        v.x += 1.f;
        dst[i] = v;
    }
}

template <class F>
static float time_ms(F f, int reps) {
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    f(); f();
    cudaDeviceSynchronize();
    cudaEventRecord(e0);
    for (int i = 0; i < reps; ++i) f();
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    return ms / reps;
}

int main() {
    const size_t bytes = (size_t)NP * LD * sizeof(float);       // 170 MB (> 126 MB L2)
    float *X, *Y;
    cudaMalloc(&X, bytes); cudaMalloc(&Y, bytes);
    cudaMemset(X, 0, bytes); cudaMemset(Y, 0, bytes);
    const double gb = 2.0 * bytes / 1e9;                        // read + write
    {
        const float ms = time_ms([&] { copy_kernel<<<148 * 16, 256>>>((const float4*)X, (float4*)Y, (long)(bytes / 16)); }, 10);
        printf("contiguous copy X -> Y                              : %7.1f us  %6.0f GB/s\n", ms * 1e3, gb / ms * 1e3);
    }
    {
        const float ms = time_ms([&] { copy_kernel<<<148 * 16, 256>>>((const float4*)X, (float4*)X, (long)(bytes / 16)); }, 10);
        printf("contiguous in-place X -> X                          : %7.1f us  %6.0f GB/s\n", ms * 1e3, gb / ms * 1e3);
    }
    const int pairs = NP / 128;
    struct Cfg { int seg, tiles, threads; } cfgs[] = {
        {128, 1, 256}, {128, 8, 256}, {128, 9, 256}, {256, 4, 256}, {512, 2, 256}, {1024, 1, 256}, {128, 1, 1024}, {128, 72, 256},
    };
    for (const Cfg& c : cfgs) {
        const dim3 grid((unsigned)((LD / c.seg + c.tiles - 1) / c.tiles), pairs);
        float ms = 0;
        if (c.seg <= c.threads) {
            ms = time_ms([&] { pattern_kernel<1, 8><<<grid, c.threads>>>(X, c.seg, c.tiles); }, 10);
            printf("pattern seg=%4d floats tiles/CTA=%2d thr=%4d scalar x8 : %7.1f us  %6.0f GB/s  (%u CTAs)\n", c.seg, c.tiles,
                   c.threads, ms * 1e3, gb / ms * 1e3, grid.x * grid.y);
        }
        if (c.seg / 4 > c.threads) continue;
        ms = time_ms([&] { pattern_kernel<4, 4><<<grid, c.threads>>>(X, c.seg, c.tiles); }, 10);
        printf("pattern seg=%4d floats tiles/CTA=%2d thr=%4d float4 x4 : %7.1f us  %6.0f GB/s\n", c.seg, c.tiles, c.threads,
               ms * 1e3, gb / ms * 1e3);
    }
    printf("cuda: %s\n", cudaGetErrorString(cudaDeviceSynchronize()));
    return 0;
}
