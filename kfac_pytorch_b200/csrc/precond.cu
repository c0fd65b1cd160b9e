// precond.cu -- two-sided preconditioned gradient and the KL-clip rewrite of
// the gradients (C-ABI: kfacb_precond_eigen / kfacb_precond_inverse /
// kfacb_finalize_grads).  Replaces kfac/kfac_preconditioner_eigen_dp.py:78-93,
// kfac/kfac_preconditioner_inv_dp.py:126-138 and
// kfac/kfac_preconditioner_inv.py:145-154,178-217 of the reference.
#include "common.cuh"
#include "tc_gemm.cuh"

namespace kfacb {

// gradient matrix [W.view(out,-1) | b] read in place: (m, k) with k < dW from W, k == dW from b
struct LdGrad {
    static constexpr bool K_FAST = true;
    const float* W; const float* b; int dG, dW, dA;
    __device__ __forceinline__ float operator()(int mn, int k) const {
        if (mn >= dG || k >= dA) return 0.f;
        return k < dW ? __ldg(W + (long)mn * dW + k) : __ldg(b + mn);
    }
};

// V2 = V1 / (dG_i dA_j + damping)
struct EpiEigDivide {
    float* C; long ldc; const float* dG; const float* dA; float damping;
    __device__ __forceinline__ void operator()(int m, int n, float v) const {
        C[(long)m * ldc + n] = v / (dG[m] * dA[n] + damping);
    }
};

// out = v, stored into the local slab AND into the same offset of every peer's slab (peer memory over NVLink):
// the exchange of kfac_preconditioner_inv.py:164-175 fused into the producing GEMM (exchange.cu)
constexpr int MAX_PEERS = 15;
struct PeerDeltas { int n; long long d[MAX_PEERS]; };
struct EpiStoreBcast {
    float* C; long ldc; PeerDeltas peers;
    __device__ __forceinline__ void operator()(int m, int n, float v) const {
        float* p = C + (long)m * ldc + n;
        *p = v;
        for (int q = 0; q < peers.n; ++q) *reinterpret_cast<float*>(reinterpret_cast<char*>(p) + peers.d[q]) = v;
    }
};

static size_t precond_buf_bytes(const kfacb_layer_t* layers, int count) {
    size_t mx = 0;
    for (int i = 0; i < count; ++i) {
        const size_t e = (size_t)kfacb_layer_dA(&layers[i]) * kfacb_layer_dG(&layers[i]);
        if (e > mx) mx = e;
    }
    return align_up(mx * sizeof(float), 256);
}

// ---------------------------------------------------------------------------
// finalize
// ---------------------------------------------------------------------------
constexpr int FIN_CHUNK = 8192;
struct FinJob { const float* v; float* w; float* b; int dG, dA, dW; int chunk_begin; };

__device__ __forceinline__ int fin_find_job(const FinJob* jobs, int njobs, int chunk) {
    int lo = 0, hi = njobs - 1;
    while (lo < hi) {
        const int mid = (lo + hi + 1) >> 1;
        if (jobs[mid].chunk_begin <= chunk) lo = mid; else hi = mid - 1;
    }
    return lo;
}

// partial[chunk] = sum over the chunk of v * grad   (double accumulation, fixed order)
__global__ void __launch_bounds__(256)
fin_dot_kernel(const FinJob* jobs, int njobs, double* __restrict__ partial) {
    const int ji = fin_find_job(jobs, njobs, blockIdx.x);
    const FinJob jb = jobs[ji];
    const long tot = (long)jb.dG * jb.dA;
    const long beg = (long)(blockIdx.x - jb.chunk_begin) * FIN_CHUNK;
    const long end = min(tot, beg + FIN_CHUNK);
    double s = 0.0;
    for (long e = beg + threadIdx.x; e < end; e += 256) {
        const int m = (int)(e / jb.dA), k = (int)(e - (long)m * jb.dA);
        const float g = k < jb.dW ? jb.w[(long)m * jb.dW + k] : jb.b[m];
        s += (double)(jb.v[e] * g);
    }
    __shared__ double red[8];
    s = warp_sum(s);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0.0;
        for (int i = 0; i < 8; ++i) t += red[i];
        partial[blockIdx.x] = t;
    }
}

// nu = min(1, sqrt(kl_clip / |vg_sum * lr^2|))
__global__ void fin_nu_kernel(const double* __restrict__ partial, int nchunks, float lr,
                              float kl_clip, int force_one, float* __restrict__ nu_out) {
    __shared__ double red[32];
    double s = 0.0;
    for (int i = threadIdx.x; i < nchunks; i += blockDim.x) s += partial[i];
    s = warp_sum(s);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0.0;
        for (int i = 0; i < (int)(blockDim.x >> 5); ++i) t += red[i];
        float nu = 1.f;
        if (kl_clip > 0.f && !force_one) {
            const double vg = t * (double)lr * (double)lr;
            const double r = sqrt((double)kl_clip / fabs(vg));
            nu = (float)(r < 1.0 ? r : 1.0);
        }
        *nu_out = nu;
    }
}

// wgrad <- nu * v[:, :dW], bgrad <- nu * v[:, dW]
__global__ void __launch_bounds__(256)
fin_write_kernel(const FinJob* jobs, int njobs, const float* __restrict__ nu_ptr) {
    const int ji = fin_find_job(jobs, njobs, blockIdx.x);
    const FinJob jb = jobs[ji];
    const float nu = *nu_ptr;
    const long tot = (long)jb.dG * jb.dA;
    const long beg = (long)(blockIdx.x - jb.chunk_begin) * FIN_CHUNK;
    const long end = min(tot, beg + FIN_CHUNK);
    for (long e = beg + threadIdx.x; e < end; e += 256) {
        const int m = (int)(e / jb.dA), k = (int)(e - (long)m * jb.dA);
        const float v = jb.v[e] * nu;
        if (k < jb.dW) jb.w[(long)m * jb.dW + k] = v; else jb.b[m] = v;
    }
}

// dst[0 .. dG*dW) <- wgrad, dst[dG*dW .. dG*dA) <- bgrad   (FinJob.v doubles as the destination)
__global__ void __launch_bounds__(256)
pack_grads_kernel(const FinJob* jobs, int njobs) {
    const int ji = fin_find_job(jobs, njobs, blockIdx.x);
    const FinJob jb = jobs[ji];
    const long tot = (long)jb.dG * jb.dA, nw = (long)jb.dG * jb.dW;
    const long beg = (long)(blockIdx.x - jb.chunk_begin) * FIN_CHUNK;
    const long end = min(tot, beg + FIN_CHUNK);
    float* dst = const_cast<float*>(jb.v);
    for (long e = beg + threadIdx.x; e < end; e += 256) dst[e] = e < nw ? jb.w[e] : jb.b[e - nw];
}

// out[0] = sum of the chunk partials (fixed order)
__global__ void fin_sum_kernel(const double* __restrict__ partial, int nchunks, double* __restrict__ out) {
    __shared__ double red[32];
    double s = 0.0;
    for (int i = threadIdx.x; i < nchunks; i += blockDim.x) s += partial[i];
    s = warp_sum(s);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0.0;
        for (int i = 0; i < (int)(blockDim.x >> 5); ++i) t += red[i];
        *out = t;
    }
}

// uploads the job table of a finalize-family call; returns the number of chunks (or -1 with the error set)
static int fin_upload_jobs(const kfacb_layer_t* layers, int count, const float* const* precon, float* const* wgrad,
                           float* const* bgrad, void* workspace, cudaStream_t st) {
    FinJob* hj = new FinJob[count > 0 ? count : 1];
    int chunk = 0;
    for (int i = 0; i < count; ++i) {
        const kfacb_layer_t& l = layers[i];
        const int nA = kfacb_layer_dA(&l), nG = kfacb_layer_dG(&l);
        if (!precon[i] || !wgrad[i] || (l.has_bias && !bgrad[i])) {
            delete[] hj;
            set_error("layer %d: null pointer", i);
            return -1;
        }
        hj[i] = FinJob{precon[i], wgrad[i], bgrad[i], nG, nA, nA - (l.has_bias ? 1 : 0), chunk};
        chunk += cdiv((long)nA * nG, FIN_CHUNK);
    }
    const cudaError_t e = cudaMemcpyAsync(workspace, hj, sizeof(FinJob) * count, cudaMemcpyHostToDevice, st);
    delete[] hj;
    if (e != cudaSuccess) { set_error("job table upload failed: %s", cudaGetErrorString(e)); return -1; }
    return chunk;
}

}  // namespace kfacb

using namespace kfacb;

extern "C" size_t kfacb_precond_workspace_bytes(const kfacb_layer_t* layers, int32_t count) {
    return 2 * precond_buf_bytes(layers, count) + 256;
}

static int check_precond_args(const kfacb_layer_t* layers, int32_t count, void* workspace,
                              size_t workspace_bytes) {
    KFACB_REQUIRE(layers && count >= 0, "null argument");
    if (workspace_bytes < kfacb_precond_workspace_bytes(layers, count) || ((uintptr_t)workspace & 255)) {
        set_error("precond workspace too small or misaligned");
        return KFACB_ERR_WORKSPACE;
    }
    return KFACB_OK;
}

static int make_peers(int32_t npeers, const int64_t* deltas, PeerDeltas& pd) {
    KFACB_REQUIRE(npeers >= 0 && npeers <= MAX_PEERS && (npeers == 0 || deltas), "bad peer list");
    pd.n = npeers;
    for (int q = 0; q < npeers; ++q) pd.d[q] = (long long)deltas[q];
    return KFACB_OK;
}

static int precond_eigen_impl(const kfacb_layer_t* layers, int32_t count,
                              const float* const* QA, const float* const* QG,
                              const float* const* dA, const float* const* dG,
                              const float* const* wgrad, const float* const* bgrad,
                              float* const* out, float damping, const PeerDeltas& pd,
                              void* workspace, size_t workspace_bytes,
                              kfacb_stream_t stream) {
    cudaStream_t st = (cudaStream_t)stream;
    int rc = check_precond_args(layers, count, workspace, workspace_bytes);
    if (rc) return rc;
    KFACB_REQUIRE(QA && QG && dA && dG && wgrad && bgrad && out, "null argument");
    float* t0 = (float*)workspace;
    float* t1 = (float*)((char*)workspace + precond_buf_bytes(layers, count));
    for (int i = 0; i < count; ++i) {
        if (!wgrad[i]) continue;
        const kfacb_layer_t& l = layers[i];
        const int nA = kfacb_layer_dA(&l), nG = kfacb_layer_dG(&l);
        KFACB_REQUIRE(QA[i] && QG[i] && dA[i] && dG[i] && out[i] && (!l.has_bias || bgrad[i]),
                      "layer %d: null pointer", i);
        const int dW = nA - (l.has_bias ? 1 : 0);
        // T1 = grad QA
        KFACB_CHECK_CUDA(launch_gemm_auto(LdGrad{wgrad[i], bgrad[i], nG, dW, nA},
                                     LdMnContig{QA[i], (long)nA, nA, nA},
                                     EpiStore{t0, (long)nA, 1.f, 0.f}, nG, nA, nA, st));
        // V2 = (QG^T T1) / (dG dA^T + damping)
        KFACB_CHECK_CUDA(launch_gemm_auto(LdMnContig{QG[i], (long)nG, nG, nG},
                                     LdMnContig{t0, (long)nA, nA, nG},
                                     EpiEigDivide{t1, (long)nA, dG[i], dA[i], damping}, nG, nA, nG, st));
        // T2 = V2 QA^T
        KFACB_CHECK_CUDA(launch_gemm_auto(LdKContig{t1, (long)nA, nG, nA},
                                     LdKContig{QA[i], (long)nA, nA, nA},
                                     EpiStore{t0, (long)nA, 1.f, 0.f}, nG, nA, nA, st));
        // out = QG T2  (+ the same elements into every peer's slab when an exchange is attached)
        if (pd.n > 0)
            KFACB_CHECK_CUDA(launch_gemm_auto(LdKContig{QG[i], (long)nG, nG, nG}, LdMnContig{t0, (long)nA, nA, nG},
                                              EpiStoreBcast{out[i], (long)nA, pd}, nG, nA, nG, st));
        else
            KFACB_CHECK_CUDA(launch_gemm_auto(LdKContig{QG[i], (long)nG, nG, nG}, LdMnContig{t0, (long)nA, nA, nG},
                                              EpiStore{out[i], (long)nA, 1.f, 0.f}, nG, nA, nG, st));
    }
    return KFACB_OK;
}

extern "C" int kfacb_precond_eigen(const kfacb_layer_t* layers, int32_t count,
                                   const float* const* QA, const float* const* QG,
                                   const float* const* dA, const float* const* dG,
                                   const float* const* wgrad, const float* const* bgrad,
                                   float* const* out, float damping,
                                   void* workspace, size_t workspace_bytes,
                                   kfacb_stream_t stream) {
    PeerDeltas pd; pd.n = 0;
    return precond_eigen_impl(layers, count, QA, QG, dA, dG, wgrad, bgrad, out, damping, pd, workspace, workspace_bytes, stream);
}

extern "C" int kfacb_precond_eigen_bcast(const kfacb_layer_t* layers, int32_t count,
                                         const float* const* QA, const float* const* QG,
                                         const float* const* dA, const float* const* dG,
                                         const float* const* wgrad, const float* const* bgrad,
                                         float* const* out, float damping,
                                         int32_t npeers, const int64_t* peer_deltas,
                                         void* workspace, size_t workspace_bytes,
                                         kfacb_stream_t stream) {
    PeerDeltas pd;
    int rc = make_peers(npeers, peer_deltas, pd);
    if (rc) return rc;
    return precond_eigen_impl(layers, count, QA, QG, dA, dG, wgrad, bgrad, out, damping, pd, workspace, workspace_bytes, stream);
}

static int precond_inverse_impl(const kfacb_layer_t* layers, int32_t count,
                                const float* const* invA, const float* const* invG,
                                const float* const* wgrad, const float* const* bgrad,
                                float* const* out, const PeerDeltas& pd,
                                void* workspace, size_t workspace_bytes,
                                kfacb_stream_t stream) {
    cudaStream_t st = (cudaStream_t)stream;
    int rc = check_precond_args(layers, count, workspace, workspace_bytes);
    if (rc) return rc;
    KFACB_REQUIRE(invA && invG && wgrad && bgrad && out, "null argument");
    float* t0 = (float*)workspace;
    for (int i = 0; i < count; ++i) {
        if (!wgrad[i]) continue;
        const kfacb_layer_t& l = layers[i];
        const int nA = kfacb_layer_dA(&l), nG = kfacb_layer_dG(&l);
        KFACB_REQUIRE(invA[i] && invG[i] && out[i] && (!l.has_bias || bgrad[i]), "layer %d: null pointer", i);
        const int dW = nA - (l.has_bias ? 1 : 0);
        // T = grad invA
        KFACB_CHECK_CUDA(launch_gemm_auto(LdGrad{wgrad[i], bgrad[i], nG, dW, nA},
                                     LdMnContig{invA[i], (long)nA, nA, nA},
                                     EpiStore{t0, (long)nA, 1.f, 0.f}, nG, nA, nA, st));
        // out = invG T  (+ peers)
        if (pd.n > 0)
            KFACB_CHECK_CUDA(launch_gemm_auto(LdKContig{invG[i], (long)nG, nG, nG}, LdMnContig{t0, (long)nA, nA, nG},
                                              EpiStoreBcast{out[i], (long)nA, pd}, nG, nA, nG, st));
        else
            KFACB_CHECK_CUDA(launch_gemm_auto(LdKContig{invG[i], (long)nG, nG, nG}, LdMnContig{t0, (long)nA, nA, nG},
                                              EpiStore{out[i], (long)nA, 1.f, 0.f}, nG, nA, nG, st));
    }
    return KFACB_OK;
}

extern "C" int kfacb_precond_inverse(const kfacb_layer_t* layers, int32_t count,
                                     const float* const* invA, const float* const* invG,
                                     const float* const* wgrad, const float* const* bgrad,
                                     float* const* out,
                                     void* workspace, size_t workspace_bytes,
                                     kfacb_stream_t stream) {
    PeerDeltas pd; pd.n = 0;
    return precond_inverse_impl(layers, count, invA, invG, wgrad, bgrad, out, pd, workspace, workspace_bytes, stream);
}

extern "C" int kfacb_precond_inverse_bcast(const kfacb_layer_t* layers, int32_t count,
                                           const float* const* invA, const float* const* invG,
                                           const float* const* wgrad, const float* const* bgrad,
                                           float* const* out,
                                           int32_t npeers, const int64_t* peer_deltas,
                                           void* workspace, size_t workspace_bytes,
                                           kfacb_stream_t stream) {
    PeerDeltas pd;
    int rc = make_peers(npeers, peer_deltas, pd);
    if (rc) return rc;
    return precond_inverse_impl(layers, count, invA, invG, wgrad, bgrad, out, pd, workspace, workspace_bytes, stream);
}

static int fin_total_chunks(const kfacb_layer_t* layers, int count) {
    long c = 0;
    for (int i = 0; i < count; ++i)
        c += cdiv((long)kfacb_layer_dA(&layers[i]) * kfacb_layer_dG(&layers[i]), FIN_CHUNK);
    return (int)c;
}

extern "C" size_t kfacb_finalize_workspace_bytes(const kfacb_layer_t* layers, int32_t count) {
    return align_up(sizeof(FinJob) * (size_t)(count + 1), 256) +
           align_up(sizeof(double) * (size_t)(fin_total_chunks(layers, count) + 1), 256) + 256;
}

extern "C" int kfacb_finalize_grads(const kfacb_layer_t* layers, int32_t count,
                                    const float* const* precon,
                                    float* const* wgrad, float* const* bgrad,
                                    float lr, float kl_clip, int32_t force_nu_one,
                                    float* nu_out,
                                    void* workspace, size_t workspace_bytes,
                                    kfacb_stream_t stream) {
    cudaStream_t st = (cudaStream_t)stream;
    KFACB_REQUIRE(layers && precon && wgrad && bgrad && nu_out && count >= 0, "null argument");
    if (workspace_bytes < kfacb_finalize_workspace_bytes(layers, count) || ((uintptr_t)workspace & 255)) {
        set_error("finalize workspace too small or misaligned");
        return KFACB_ERR_WORKSPACE;
    }
    if (count == 0) return KFACB_OK;
    FinJob* hj = new FinJob[count];
    int chunk = 0;
    for (int i = 0; i < count; ++i) {
        const kfacb_layer_t& l = layers[i];
        const int nA = kfacb_layer_dA(&l), nG = kfacb_layer_dG(&l);
        if (!precon[i] || !wgrad[i] || (l.has_bias && !bgrad[i])) {
            delete[] hj;
            set_error("layer %d: null pointer (finalize needs every layer)", i);
            return KFACB_ERR_ARG;
        }
        hj[i] = FinJob{precon[i], wgrad[i], bgrad[i], nG, nA, nA - (l.has_bias ? 1 : 0), chunk};
        chunk += cdiv((long)nA * nG, FIN_CHUNK);
    }
    FinJob* dj = (FinJob*)workspace;
    double* partial = (double*)((char*)workspace + align_up(sizeof(FinJob) * (size_t)(count + 1), 256));
    cudaError_t e = cudaMemcpyAsync(dj, hj, sizeof(FinJob) * count, cudaMemcpyHostToDevice, st);
    delete[] hj;
    KFACB_CHECK_CUDA(e);
    if (kl_clip > 0.f && !force_nu_one) {
        fin_dot_kernel<<<chunk, 256, 0, st>>>(dj, count, partial);
        KFACB_CHECK_LAUNCH();
    }
    fin_nu_kernel<<<1, 1024, 0, st>>>(partial, chunk, lr, kl_clip, force_nu_one, nu_out);
    KFACB_CHECK_LAUNCH();
    fin_write_kernel<<<chunk, 256, 0, st>>>(dj, count, nu_out);
    KFACB_CHECK_LAUNCH();
    return KFACB_OK;
}

// ---------------------------------------------------------------------------
// split finalize + gradient packing for the fused gradient-reduce path (replaces the caller's
// all-reduce, examples/pytorch_imagenet_resnet.py:371-383, with reduce-to-owner -> precondition -> exchange)
// ---------------------------------------------------------------------------
extern "C" int kfacb_pack_grads(const kfacb_layer_t* layers, int32_t count,
                                const float* const* wgrad, const float* const* bgrad, float* const* dst,
                                void* workspace, size_t workspace_bytes, kfacb_stream_t stream) {
    cudaStream_t st = (cudaStream_t)stream;
    KFACB_REQUIRE(layers && wgrad && bgrad && dst && count >= 0, "null argument");
    if (workspace_bytes < kfacb_finalize_workspace_bytes(layers, count) || ((uintptr_t)workspace & 255)) {
        set_error("pack workspace too small or misaligned");
        return KFACB_ERR_WORKSPACE;
    }
    if (count == 0) return KFACB_OK;
    const int chunk = fin_upload_jobs(layers, count, const_cast<const float* const*>(dst), const_cast<float* const*>(wgrad),
                                      const_cast<float* const*>(bgrad), workspace, st);
    if (chunk < 0) return KFACB_ERR_ARG;
    pack_grads_kernel<<<chunk, 256, 0, st>>>((const FinJob*)workspace, count);
    KFACB_CHECK_LAUNCH();
    return KFACB_OK;
}

extern "C" int kfacb_finalize_dot(const kfacb_layer_t* layers, int32_t count,
                                  const float* const* precon, const float* const* wgrad, const float* const* bgrad,
                                  double* vg_sum_out, void* workspace, size_t workspace_bytes, kfacb_stream_t stream) {
    cudaStream_t st = (cudaStream_t)stream;
    KFACB_REQUIRE(layers && precon && wgrad && bgrad && vg_sum_out && count >= 0, "null argument");
    if (workspace_bytes < kfacb_finalize_workspace_bytes(layers, count) || ((uintptr_t)workspace & 255)) {
        set_error("finalize workspace too small or misaligned");
        return KFACB_ERR_WORKSPACE;
    }
    if (count == 0) { KFACB_CHECK_CUDA(cudaMemsetAsync(vg_sum_out, 0, sizeof(double), st)); return KFACB_OK; }
    const int chunk = fin_upload_jobs(layers, count, precon, const_cast<float* const*>(wgrad), const_cast<float* const*>(bgrad),
                                      workspace, st);
    if (chunk < 0) return KFACB_ERR_ARG;
    double* partial = (double*)((char*)workspace + align_up(sizeof(FinJob) * (size_t)(count + 1), 256));
    fin_dot_kernel<<<chunk, 256, 0, st>>>((const FinJob*)workspace, count, partial);
    KFACB_CHECK_LAUNCH();
    fin_sum_kernel<<<1, 1024, 0, st>>>(partial, chunk, vg_sum_out);
    KFACB_CHECK_LAUNCH();
    return KFACB_OK;
}

extern "C" int kfacb_finalize_write(const kfacb_layer_t* layers, int32_t count,
                                    const float* const* precon, float* const* wgrad, float* const* bgrad,
                                    const double* vg_sum, float lr, float kl_clip, int32_t force_nu_one, float* nu_out,
                                    void* workspace, size_t workspace_bytes, kfacb_stream_t stream) {
    cudaStream_t st = (cudaStream_t)stream;
    KFACB_REQUIRE(layers && precon && wgrad && bgrad && nu_out && count >= 0, "null argument");
    KFACB_REQUIRE(vg_sum || !(kl_clip > 0.f) || force_nu_one, "vg_sum is required when the KL clip is active");
    if (workspace_bytes < kfacb_finalize_workspace_bytes(layers, count) || ((uintptr_t)workspace & 255)) {
        set_error("finalize workspace too small or misaligned");
        return KFACB_ERR_WORKSPACE;
    }
    if (count == 0) return KFACB_OK;
    const int chunk = fin_upload_jobs(layers, count, precon, wgrad, bgrad, workspace, st);
    if (chunk < 0) return KFACB_ERR_ARG;
    fin_nu_kernel<<<1, 32, 0, st>>>(vg_sum, vg_sum ? 1 : 0, lr, kl_clip, force_nu_one || !vg_sum, nu_out);
    KFACB_CHECK_LAUNCH();
    fin_write_kernel<<<chunk, 256, 0, st>>>((const FinJob*)workspace, count, nu_out);
    KFACB_CHECK_LAUNCH();
    return KFACB_OK;
}
