// inverse.cu -- pi-damped SPD inverse (C-ABI: kfacb_pi_damped_inverse).
// Replaces kfac/kfac_preconditioner_inv_dp.py:98-123, _add_value_to_diagonal
// (kfac_preconditioner_inv.py:106-107) and mat_inv (kfac/utils.py:11-16:
// cholesky + cholesky_inverse) of the reference.
//
// The inverse is computed by an in-place BLOCK GAUSS-JORDAN sweep without
// pivoting (all pivots are Schur complements of an SPD matrix, hence SPD):
// for each 64-wide diagonal block k
//     D      = W_kk^-1                       (scalar Gauss-Jordan in shared memory)
//     C      = -W[:,k] D                     (n x 64 panel, SIMT GEMM)
//     W_ij  += C_i W_kj      (i,j != k)      (rank-64 update of the whole matrix, SIMT GEMM)
//     W_kj   = D W_kj, W_ik = C_i, W_kk = D
// After the last block W = M^-1; it is symmetrised on the way out (the
// reference's cholesky_inverse result is exactly symmetric).
#include "common.cuh"
#include "tc_gemm.cuh"

namespace kfacb {

constexpr int GJ_NB = 64;

struct DampJob { float* A; float* G; int nA, nG; };

// pi = sqrt((tr A / dA) / (tr G / dG));  A += sqrt(damping)*pi I;  G += sqrt(damping)/pi I
__global__ void pi_damp_kernel(const DampJob* jobs, float sqrt_damping) {
    const DampJob jb = jobs[blockIdx.x];
    __shared__ float red[2][8];
    float ta = 0.f, tg = 0.f;
    for (int i = threadIdx.x; i < jb.nA; i += blockDim.x) ta += jb.A[(long)i * jb.nA + i];
    for (int i = threadIdx.x; i < jb.nG; i += blockDim.x) tg += jb.G[(long)i * jb.nG + i];
    ta = warp_sum(ta); tg = warp_sum(tg);
    if ((threadIdx.x & 31) == 0) { red[0][threadIdx.x >> 5] = ta; red[1][threadIdx.x >> 5] = tg; }
    __syncthreads();
    ta = 0.f; tg = 0.f;
    for (int i = 0; i < (int)(blockDim.x >> 5); ++i) { ta += red[0][i]; tg += red[1][i]; }
    const float pi = sqrtf((ta / (float)jb.nA) / (tg / (float)jb.nG));
    const float da = sqrt_damping * pi, dg = sqrt_damping / pi;
    for (int i = threadIdx.x; i < jb.nA; i += blockDim.x) jb.A[(long)i * jb.nA + i] += da;
    for (int i = threadIdx.x; i < jb.nG; i += blockDim.x) jb.G[(long)i * jb.nG + i] += dg;
}

// D = inverse of the nb x nb diagonal block at (k0, k0) of W (n x n).  one CTA, 1024 threads: thread (i, jq) owns the
// 4 elements a[i][4 jq .. 4 jq + 3] in REGISTERS for the whole elimination; a pivot step publishes row s and column s
// through shared memory (one barrier per pivot).  This kernel is the serial link of every block step.
constexpr int GJ_DIAG_THREADS = GJ_NB * GJ_NB / 4;
__global__ void __launch_bounds__(GJ_DIAG_THREADS)
gj_diag_kernel(const float* __restrict__ W, int n, int k0, int nb, float* __restrict__ D) {
    __shared__ float rowv[2][GJ_NB], colv[2][GJ_NB];
    const int tid = threadIdx.x, i = tid >> 4, j0 = (tid & 15) * 4;
    float a[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
        const int j = j0 + u;
        a[u] = (i < nb && j < nb) ? W[(long)(k0 + i) * n + k0 + j] : (i == j ? 1.f : 0.f);
    }
    for (int s = 0; s < nb; ++s) {
        const int b = s & 1;                                   // double-buffered: one barrier per pivot
        if (i == s) {
#pragma unroll
            for (int u = 0; u < 4; ++u) rowv[b][j0 + u] = a[u];
        }
        if ((s >> 2) == (tid & 15)) colv[b][i] = a[s & 3];
        __syncthreads();
        const float piv = 1.f / rowv[b][s];
        const float ci = colv[b][i] * piv;
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int j = j0 + u;
            float v;
            if (i == s && j == s) v = piv;
            else if (i == s) v = a[u] * piv;
            else if (j == s) v = -ci;
            else v = fmaf(-ci, rowv[b][j], a[u]);
            a[u] = v;
        }
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) D[i * GJ_NB + j0 + u] = a[u];
}

// whole-matrix path for n <= 64: out = M^-1 directly.  one CTA per matrix
struct SmallInvJob { const float* M; float* out; int n; };
__global__ void __launch_bounds__(256)
gj_small_kernel(const SmallInvJob* jobs) {
    const SmallInvJob jb = jobs[blockIdx.x];
    __shared__ float a[GJ_NB][GJ_NB + 1];
    __shared__ float colv[GJ_NB], rowv[GJ_NB];
    const int tid = threadIdx.x, n = jb.n;
    for (int e = tid; e < GJ_NB * GJ_NB; e += 256) {
        const int i = e / GJ_NB, j = e - i * GJ_NB;
        a[i][j] = (i < n && j < n) ? jb.M[(long)i * n + j] : (i == j ? 1.f : 0.f);
    }
    __syncthreads();
    for (int s = 0; s < n; ++s) {
        const float piv = 1.f / a[s][s];
        if (tid < GJ_NB) { colv[tid] = a[tid][s]; rowv[tid] = a[s][tid]; }
        __syncthreads();
        for (int e = tid; e < GJ_NB * GJ_NB; e += 256) {
            const int i = e / GJ_NB, j = e - i * GJ_NB;
            float v;
            if (i == s && j == s) v = piv;
            else if (i == s) v = rowv[j] * piv;
            else if (j == s) v = -colv[i] * piv;
            else v = a[i][j] - colv[i] * piv * rowv[j];
            a[i][j] = v;
        }
        __syncthreads();
    }
    for (int e = tid; e < n * n; e += 256) {
        const int i = e / n, j = e - i * n;
        jb.out[e] = 0.5f * (a[i][j] + a[j][i]);
    }
}

// C[i][:] = -W[i][k0:k0+nb] D   for all rows i (rows inside block k are computed but unused)
struct LdPanel {       // A operand (m = row i, kk): W[i*n + k0 + kk]
    static constexpr bool K_FAST = true;
    const float* W; int n, k0, nb;
    __device__ __forceinline__ float operator()(int mn, int k) const {
        return (mn < n && k < nb) ? W[(long)mn * n + k0 + k] : 0.f;
    }
};
struct LdDcol {        // B operand (n = col c of D, kk): D[kk*64 + c]
    static constexpr bool K_FAST = false;
    const float* D; int nb;
    __device__ __forceinline__ float operator()(int mn, int k) const {
        return (mn < nb && k < nb) ? D[k * GJ_NB + mn] : 0.f;
    }
};
struct EpiNegStore {
    float* C;
    __device__ __forceinline__ void operator()(int m, int n, float v) const { C[(long)m * GJ_NB + n] = -v; }
};

// rank-nb update: W[i][j] += sum_kk C[i][kk] W[k0+kk][j]   for i, j outside block k
struct LdC {
    static constexpr bool K_FAST = true;
    const float* C; int n, nb;
    __device__ __forceinline__ float operator()(int mn, int k) const {
        return (mn < n && k < nb) ? C[(long)mn * GJ_NB + k] : 0.f;
    }
};
struct LdRowBlk {
    static constexpr bool K_FAST = false;
    const float* W; int n, k0, nb;
    __device__ __forceinline__ float operator()(int mn, int k) const {
        return (mn < n && k < nb) ? W[(long)(k0 + k) * n + mn] : 0.f;
    }
};
struct EpiRankUpdate {
    float* W; int n, k0, nb;
    __device__ __forceinline__ void operator()(int m, int c, float v) const {
        if ((unsigned)(m - k0) < (unsigned)nb || (unsigned)(c - k0) < (unsigned)nb) return;
        W[(long)m * n + c] += v;
    }
    // read-modify-write form for the tensor-core kernel (loads batched ahead of the stores, tc_gemm.cuh)
    __device__ __forceinline__ float* addr(int m, int c) const {
        if ((unsigned)(m - k0) < (unsigned)nb || (unsigned)(c - k0) < (unsigned)nb) return nullptr;
        return W + (long)m * n + c;
    }
    __device__ __forceinline__ float combine(float old, float v) const { return old + v; }
};

// W[k rows][j] = D W[k rows][j] (j outside k);  W[i][k cols] = C[i] (i outside k);  W_kk = D
// Two jobs per CTA of 128 threads: (1) the thread of column j updates W[k0 .. k0+nb)[j] (coalesced along j; D is read
// through registers 4 at a time, the 64 old values stay in registers); (2) the same 128 rows get their column block
// from C with lanes walking the 64 contiguous columns of a row (coalesced 256-byte row pieces; thread-per-row copies
// touched 64 separate sectors per instruction and were the slowest kernel of the inverse).
__global__ void __launch_bounds__(128)
gj_fix_kernel(float* __restrict__ W, int n, int k0, int nb,
              const float* __restrict__ D, const float* __restrict__ C) {
    __shared__ float ds[GJ_NB][GJ_NB + 1];
    for (int e = threadIdx.x; e < GJ_NB * GJ_NB; e += blockDim.x) ds[e / GJ_NB][e % GJ_NB] = D[e];
    __syncthreads();
    const int j = blockIdx.x * blockDim.x + threadIdx.x;     // column of the row-block fix
    if (j < n && (unsigned)(j - k0) >= (unsigned)nb) {
        float old[GJ_NB];
#pragma unroll 8
        for (int s = 0; s < GJ_NB; ++s) old[s] = s < nb ? W[(long)(k0 + s) * n + j] : 0.f;
        for (int r = 0; r < nb; ++r) {
            float acc = 0.f;
#pragma unroll 8
            for (int s = 0; s < GJ_NB; ++s) acc = fmaf(ds[r][s], old[s], acc);
            W[(long)(k0 + r) * n + j] = acc;
        }
    }
    // column block (and the diagonal block itself) of the rows of this CTA: warp w takes rows w, w + 4, ..
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int rr = warp; rr < (int)blockDim.x; rr += 4) {
        const int i = blockIdx.x * blockDim.x + rr;
        if (i >= n) break;
        const bool inside = (unsigned)(i - k0) < (unsigned)nb;
        for (int c = lane; c < nb; c += 32)
            W[(long)i * n + k0 + c] = inside ? ds[i - k0][c] : C[(long)i * GJ_NB + c];
    }
}

__global__ void symmetrize_kernel(float* __restrict__ W, int n) {
    const long idx = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (long)n * n) return;
    const int i = (int)(idx / n), j = (int)(idx - (long)i * n);
    if (i < j) {
        const float v = 0.5f * (W[idx] + W[(long)j * n + i]);
        W[idx] = v;
        W[(long)j * n + i] = v;
    }
}

constexpr int INV_STREAMS = 8;     // independent matrices are inverted concurrently on internal streams (VGG-16 has five 4609-dim chains of 72 dependent block steps each)

static size_t inv_scratch_bytes(int nmax) {
    // D (64x64) + C (nmax x 64)
    return align_up(GJ_NB * GJ_NB * sizeof(float), 256) + align_up((size_t)nmax * GJ_NB * sizeof(float), 256);
}

static int invert_large(const float* M, float* out, int n, float* D, float* C, cudaStream_t st) {
    KFACB_CHECK_CUDA(cudaMemcpyAsync(out, M, (size_t)n * n * sizeof(float), cudaMemcpyDeviceToDevice, st));
    for (int k0 = 0; k0 < n; k0 += GJ_NB) {
        const int nb = (n - k0 < GJ_NB) ? n - k0 : GJ_NB;
        gj_diag_kernel<<<1, GJ_DIAG_THREADS, 0, st>>>(out, n, k0, nb, D);
        KFACB_CHECK_LAUNCH();
        KFACB_CHECK_CUDA(launch_gemm_auto(LdPanel{out, n, k0, nb}, LdDcol{D, nb}, EpiNegStore{C}, n, nb, nb, st));
        KFACB_CHECK_CUDA(launch_gemm_auto(LdC{C, n, nb}, LdRowBlk{out, n, k0, nb},
                                     EpiRankUpdate{out, n, k0, nb}, n, n, nb, st));
        gj_fix_kernel<<<cdiv(n, 128), 128, 0, st>>>(out, n, k0, nb, D, C);
        KFACB_CHECK_LAUNCH();
    }
    symmetrize_kernel<<<cdiv((long)n * n, 256), 256, 0, st>>>(out, n);
    KFACB_CHECK_LAUNCH();
    return KFACB_OK;
}

}  // namespace kfacb

using namespace kfacb;

extern "C" size_t kfacb_spd_inverse_workspace_bytes(const int32_t* nA, const int32_t* nG, int32_t count) {
    int nmax = 1;
    for (int i = 0; i < count; ++i) {
        if (nA[i] > nmax) nmax = nA[i];
        if (nG[i] > nmax) nmax = nG[i];
    }
    return align_up(sizeof(DampJob) * (size_t)(count + 1), 256) +
           align_up(sizeof(SmallInvJob) * (size_t)(2 * count + 1), 256) + INV_STREAMS * inv_scratch_bytes(nmax) + 256;
}

extern "C" int kfacb_pi_damped_inverse(const int32_t* nA, const int32_t* nG, int32_t count,
                                       float* const* A, float* const* G,
                                       float* const* invA, float* const* invG,
                                       float damping,
                                       void* workspace, size_t workspace_bytes,
                                       kfacb_stream_t stream) {
    cudaStream_t st = (cudaStream_t)stream;
    KFACB_REQUIRE(nA && nG && A && G && invA && invG && count >= 0, "null argument");
    if (workspace_bytes < kfacb_spd_inverse_workspace_bytes(nA, nG, count) || ((uintptr_t)workspace & 255)) {
        set_error("inverse workspace too small or misaligned");
        return KFACB_ERR_WORKSPACE;
    }
    char* ws = (char*)workspace;
    DampJob* djobs = (DampJob*)ws;
    size_t off = align_up(sizeof(DampJob) * (size_t)(count + 1), 256);
    SmallInvJob* sjobs = (SmallInvJob*)(ws + off);
    off += align_up(sizeof(SmallInvJob) * (size_t)(2 * count + 1), 256);
    float* D = (float*)(ws + off);
    float* C = (float*)(ws + off + align_up(GJ_NB * GJ_NB * sizeof(float), 256));

    DampJob* hd = new DampJob[count + 1];
    SmallInvJob* hs = new SmallInvJob[2 * count + 1];
    int nd = 0, ns = 0;
    for (int i = 0; i < count; ++i) {
        if (!A[i]) continue;
        if (!G[i] || !invA[i] || !invG[i] || nA[i] <= 0 || nG[i] <= 0) {
            delete[] hd; delete[] hs;
            set_error("layer %d: null pointer or bad size", i);
            return KFACB_ERR_ARG;
        }
        hd[nd++] = DampJob{A[i], G[i], nA[i], nG[i]};
        if (nA[i] <= GJ_NB) hs[ns++] = SmallInvJob{A[i], invA[i], nA[i]};
        if (nG[i] <= GJ_NB) hs[ns++] = SmallInvJob{G[i], invG[i], nG[i]};
    }
    cudaError_t e1 = cudaSuccess, e2 = cudaSuccess;
    if (nd) e1 = cudaMemcpyAsync(djobs, hd, sizeof(DampJob) * nd, cudaMemcpyHostToDevice, st);
    if (ns) e2 = cudaMemcpyAsync(sjobs, hs, sizeof(SmallInvJob) * ns, cudaMemcpyHostToDevice, st);
    delete[] hd; delete[] hs;
    KFACB_CHECK_CUDA(e1); KFACB_CHECK_CUDA(e2);
    if (nd == 0) return KFACB_OK;
    pi_damp_kernel<<<nd, 256, 0, st>>>(djobs, (float)sqrt((double)damping));
    KFACB_CHECK_LAUNCH();
    if (ns) {
        gj_small_kernel<<<ns, 256, 0, st>>>(sjobs);
        KFACB_CHECK_LAUNCH();
    }
    // The ~n/64 dependent steps of one block Gauss-Jordan sweep use few SMs each: run independent matrices
    // concurrently on INV_STREAMS internal streams (fork from / join into the caller's stream with events).
    static cudaStream_t xs[INV_STREAMS] = {nullptr};
    static cudaEvent_t ev_fork = nullptr, ev_join[INV_STREAMS] = {nullptr};
    if (!ev_fork) {
        KFACB_CHECK_CUDA(cudaEventCreateWithFlags(&ev_fork, cudaEventDisableTiming));
        for (int k = 0; k < INV_STREAMS; ++k) {
            KFACB_CHECK_CUDA(cudaStreamCreateWithFlags(&xs[k], cudaStreamNonBlocking));
            KFACB_CHECK_CUDA(cudaEventCreateWithFlags(&ev_join[k], cudaEventDisableTiming));
        }
    }
    int nmax = 1;
    for (int i = 0; i < count; ++i) { if (nA[i] > nmax) nmax = nA[i]; if (nG[i] > nmax) nmax = nG[i]; }
    const size_t scratch = inv_scratch_bytes(nmax);
    KFACB_CHECK_CUDA(cudaEventRecord(ev_fork, st));
    for (int k = 0; k < INV_STREAMS; ++k) KFACB_CHECK_CUDA(cudaStreamWaitEvent(xs[k], ev_fork, 0));
    int next = 0;
    auto run = [&](const float* Msrc, float* out, int n) -> int {
        const int k = next++ % INV_STREAMS;
        float* Dk = (float*)((char*)D + k * scratch);
        float* Ck = (float*)((char*)Dk + align_up(GJ_NB * GJ_NB * sizeof(float), 256));
        return invert_large(Msrc, out, n, Dk, Ck, xs[k]);
    };
    // largest first: better packing of the streams
    int rc_all = KFACB_OK;
    for (int pass = 0; pass < 2 && !rc_all; ++pass) {
        for (int i = 0; i < count && !rc_all; ++i) {
            if (!A[i]) continue;
            const bool bigA = nA[i] > 1024, bigG = nG[i] > 1024;
            if (nA[i] > GJ_NB && (pass == 0) == bigA) rc_all = run(A[i], invA[i], nA[i]);
            if (!rc_all && nG[i] > GJ_NB && (pass == 0) == bigG) rc_all = run(G[i], invG[i], nG[i]);
        }
    }
    for (int k = 0; k < INV_STREAMS; ++k) {
        KFACB_CHECK_CUDA(cudaEventRecord(ev_join[k], xs[k]));
        KFACB_CHECK_CUDA(cudaStreamWaitEvent(st, ev_join[k], 0));
    }
    if (rc_all) return rc_all;
    return KFACB_OK;
}
