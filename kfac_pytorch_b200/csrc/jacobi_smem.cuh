// jacobi_smem.cuh -- two-sided cyclic Jacobi on a small symmetric matrix resident in shared memory.
// Shared by the direct eigensolver path (n <= 128, eigh.cu) and the leaves of the divide-and-conquer tree
// (eigh_dc.cu).
#pragma once
#include "common.cuh"

namespace kfacb {

constexpr int DIRECT_MAX_N = 128;

// ---------------------------------------------------------------------------
// two-sided cyclic Jacobi on an m x m symmetric matrix in shared memory
// (m even, pitch P odd; M_CT = m at compile time or 0).  On exit S is diagonal (to tolerance) and U holds the
// accumulated rotations (columns = eigenvectors of the input).  Returns the
// number of sweeps executed; *sig_out = 1 if any rotated element exceeded sig_abs.
// All threads of the CTA must call it.
// ---------------------------------------------------------------------------
template <int M_CT>
__device__ int jacobi_smem(float* S, float* U, int m_rt, int P, float tol_rel, float abs_thr,
                           float sig_abs, int max_sweeps, int* sig_out) {
    __shared__ float s_c[DIRECT_MAX_N / 2], s_s[DIRECT_MAX_N / 2];
    __shared__ int s_p[DIRECT_MAX_N / 2], s_q[DIRECT_MAX_N / 2];
    __shared__ int s_flag, s_sig;
    const int m = M_CT ? M_CT : m_rt;            // compile-time size for the 64x64 inner problem
    const int tid = threadIdx.x, nt = blockDim.x;
    const int half = m >> 1, mm1 = m - 1;
    if (tid == 0) s_sig = 0;
    int sweeps = 0;
    int prev_p = -1, prev_q = -1;
    for (int sw = 0; sw < max_sweeps; ++sw) {
        __syncthreads();
        if (tid == 0) s_flag = 0;
        __syncthreads();
        for (int r = 0; r < mm1; ++r) {
            if (tid < half) {
                if (prev_p >= 0) { S[prev_p * P + prev_q] = 0.f; S[prev_q * P + prev_p] = 0.f; prev_p = -1; }
                int p, q;
                if (tid == 0) { p = mm1; q = r; }
                else { p = r + tid; if (p >= mm1) p -= mm1; q = r - tid; if (q < 0) q += mm1; }
                if (p > q) { const int t = p; p = q; q = t; }
                const float app = S[p * P + p], aqq = S[q * P + q], apq = S[p * P + q];
                float c = 1.f, s = 0.f;
                const float mag = fabsf(apq);
                const float ref = sqrtf(fabsf(app) * fabsf(aqq));
                if (mag > tol_rel * ref && mag > abs_thr) {
                    // Rutishauser's stable formulas; correctly rounded sqrt/div keep c^2+s^2 = 1 to ~1 ulp
                    const float tau = (aqq - app) / (2.f * apq);
                    const float t = (tau >= 0.f ? 1.f : -1.f) / (fabsf(tau) + sqrtf(1.f + tau * tau));
                    c = 1.f / sqrtf(1.f + t * t);
                    s = t * c;
                    s_flag = 1;
                    if (mag > sig_abs) s_sig = 1;
                    prev_p = p; prev_q = q;
                }
                s_c[tid] = c; s_s[tid] = s; s_p[tid] = p; s_q[tid] = q;
            }
            __syncthreads();
            // column rotations: S <- S J, U <- U J   (lanes walk the row index: conflict-free, P odd)
            for (int e = tid; e < half * m; e += nt) {
                const int k = e / m, i = e - k * m;
                const float s = s_s[k];
                if (s != 0.f) {
                    const float c = s_c[k];
                    const int p = s_p[k], q = s_q[k];
                    float x = S[i * P + p], y = S[i * P + q];
                    S[i * P + p] = c * x - s * y;
                    S[i * P + q] = s * x + c * y;
                    x = U[i * P + p]; y = U[i * P + q];
                    U[i * P + p] = c * x - s * y;
                    U[i * P + q] = s * x + c * y;
                }
            }
            __syncthreads();
            // row rotations: S <- J^T S
            for (int e = tid; e < half * m; e += nt) {
                const int k = e / m, j = e - k * m;
                const float s = s_s[k];
                if (s != 0.f) {
                    const float c = s_c[k];
                    const int p = s_p[k], q = s_q[k];
                    const float x = S[p * P + j], y = S[q * P + j];
                    S[p * P + j] = c * x - s * y;
                    S[q * P + j] = s * x + c * y;
                }
            }
            __syncthreads();
        }
        ++sweeps;
        if (!s_flag) break;
    }
    __syncthreads();
    if (tid < half && prev_p >= 0) { S[prev_p * P + prev_q] = 0.f; S[prev_q * P + prev_p] = 0.f; }
    __syncthreads();
    if (sig_out) *sig_out = s_sig;
    return sweeps;
}

}  // namespace kfacb
