// factor.cu -- Kronecker-factor accumulation fused with the running average
// (C-ABI: kfacb_factor_update).  Replaces kfac/utils.py:33-140 and
// kfac/kfac_preconditioner_inv_dp.py:75-90 of the reference.
//
// One SYRK kernel:  F <- (1-alpha) F + alpha * scale * X^T X  over the upper
// triangle of 128x128 tiles (mirrored on store), X never materialised:
//   * conv A : implicit im2col gather from the NCHW input (zero padding,
//              stride; feature order (c,kh,kw)); bias "ones" column synthesised
//   * conv G : NCHW grad-output viewed as (B*S) x C_out rows
//   * linear : row-major (B x D) view (after the sequence mean for 3-D inputs)
// Tall-skinny problems (rows >> n) are split along the row dimension over
// gridDim.y with a deterministic second-stage reduction fused with the EMA.
#include "common.cuh"
#include "tc_gemm.cuh"

namespace kfacb {

// exact n / d for 0 <= n < 2^31 (Granlund-Montgomery, 33-bit magic folded into add+shift)
struct FastDiv {
    uint32_t mul, shr;
    __host__ void init(uint32_t d) {
        uint32_t l = 0;
        while ((1ull << l) < d) ++l;
        shr = l;
        mul = (uint32_t)(((1ull << 32) * ((1ull << l) - d)) / d + 1);
    }
    __device__ __forceinline__ uint32_t div(uint32_t n) const { return (__umulhi(mul, n) + n) >> shr; }
};

// linear layers: X(k, m) = p[k*D + m]; optional ones column at index D
struct LdRows {
    static constexpr bool K_FAST = false;
    const float* p; int D; int ones_col; int K;
    __device__ __forceinline__ float operator()(int mn, int k) const {
        if (k >= K) return 0.f;
        if (mn < D) return __ldg(p + (long)k * D + mn);
        return mn == ones_col ? 1.f : 0.f;
    }
};

// NCHW tensor viewed as rows (b, s) x channels: X(k=(b,s), m=c) = p[(b*C + c)*S + s]
struct LdNCS {
    static constexpr bool K_FAST = true;
    const float* p; int C; int S; int ones_col; int K; FastDiv divS;
    __device__ __forceinline__ float operator()(int mn, int k) const {
        if (k >= K) return 0.f;
        if (mn < C) {
            const int b = divS.div(k), s = k - b * S;
            return __ldg(p + ((long)b * C + mn) * S + s);
        }
        return mn == ones_col ? 1.f : 0.f;
    }
    __device__ __forceinline__ float4 quad(int mn, int k) const {       // 4 consecutive rows (b, s..s+3)
        if (mn < C && k + 3 < K) {
            const int b = divS.div(k), s = k - b * S;
            if (s + 3 < S) {
                const float* q = p + ((long)b * C + mn) * S + s;
                if ((S & 3) == 0) return __ldg(reinterpret_cast<const float4*>(q));     // k % 4 == 0 => 16 B aligned
                return make_float4(__ldg(q), __ldg(q + 1), __ldg(q + 2), __ldg(q + 3));
            }
        }
        return make_float4((*this)(mn, k), (*this)(mn, k + 1), (*this)(mn, k + 2), (*this)(mn, k + 3));
    }
};

// implicit im2col: X(k=(b,oh,ow), m=(c,i,j)) = x[b, c, oh*sh+i-ph, ow*sw+j-pw] (0 outside)
struct LdIm2col {
    static constexpr bool K_FAST = true;
    const float* p;
    int C, H, W, KW, KHW, SH, SW, PH, PW, OW, S, D, ones_col, K;
    FastDiv divS, divOW, divKHW, divKW;
    __device__ __forceinline__ float operator()(int mn, int k) const {
        if (k >= K) return 0.f;
        if (mn >= D) return mn == ones_col ? 1.f : 0.f;
        const int c = divKHW.div(mn), r = mn - c * KHW;
        const int i = divKW.div(r), j = r - i * KW;
        const int b = divS.div(k), s = k - b * S;
        const int oh = divOW.div(s), ow = s - oh * OW;
        const int ih = oh * SH + i - PH, iw = ow * SW + j - PW;
        if ((unsigned)ih >= (unsigned)H || (unsigned)iw >= (unsigned)W) return 0.f;
        return __ldg(p + (((long)b * C + c) * H + ih) * W + iw);
    }
    __device__ __forceinline__ float4 quad(int mn, int k) const {       // 4 consecutive output positions, one feature
        if (mn < D && k + 3 < K) {
            const int b = divS.div(k), s = k - b * S;
            const int oh = divOW.div(s), ow = s - oh * OW;
            if (ow + 3 < OW) {                                   // same output row: decode once, walk iw by SW
                const int c = divKHW.div(mn), r = mn - c * KHW;
                const int i = divKW.div(r), j = r - i * KW;
                const int ih = oh * SH + i - PH;
                if ((unsigned)ih >= (unsigned)H) return make_float4(0.f, 0.f, 0.f, 0.f);
                const int iw = ow * SW + j - PW;
                const float* q = p + (((long)b * C + c) * H + ih) * W;
                float4 v;
                v.x = (unsigned)iw < (unsigned)W ? __ldg(q + iw) : 0.f;
                v.y = (unsigned)(iw + SW) < (unsigned)W ? __ldg(q + iw + SW) : 0.f;
                v.z = (unsigned)(iw + 2 * SW) < (unsigned)W ? __ldg(q + iw + 2 * SW) : 0.f;
                v.w = (unsigned)(iw + 3 * SW) < (unsigned)W ? __ldg(q + iw + 3 * SW) : 0.f;
                return v;
            }
        }
        return make_float4((*this)(mn, k), (*this)(mn, k + 1), (*this)(mn, k + 2), (*this)(mn, k + 3));
    }
};

// upper-triangular tile index -> (ti <= tj)
__device__ __forceinline__ void tri_decode(int t, int T, int& ti, int& tj) {
    ti = 0;
    int rowlen = T;
    while (t >= rowlen) { t -= rowlen; --rowlen; ++ti; }
    tj = ti + t;
}

template <class L>
__global__ void __launch_bounds__(GEMM_THREADS, 2)
syrk_ema_kernel(L ld, int n, int K, int kchunk, float scale, float alpha, int first_step,
                float* __restrict__ factor, float* __restrict__ partial) {
    __shared__ GemmSmem sm;
    const int T = (n + TILE - 1) / TILE;
    int ti, tj;
    tri_decode(blockIdx.x, T, ti, tj);
    const int kbeg = blockIdx.y * kchunk;
    const int kend = min(K, kbeg + kchunk);
    float acc[8][8];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;
    gemm_mainloop(ld, ld, ti * TILE, tj * TILE, kbeg, kend, sm, acc);

    const float keep = 1.f - alpha;
    if (gridDim.y == 1) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const int m = ti * TILE + frag_row(i);
            if (m >= n) continue;
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const int c = tj * TILE + frag_col(j);
                if (c >= n) continue;
                const float old = first_step ? (m == c ? 1.f : 0.f) : factor[(long)m * n + c];
                const float v = fmaf(alpha, scale * acc[i][j], keep * old);
                factor[(long)m * n + c] = v;
                if (ti != tj) factor[(long)c * n + m] = v;
            }
        }
    } else {
        float* dst = partial + (long)blockIdx.y * n * n;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const int m = ti * TILE + frag_row(i);
            if (m >= n) continue;
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const int c = tj * TILE + frag_col(j);
                if (c < n) dst[(long)m * n + c] = acc[i][j];
            }
        }
    }
}

// tensor-core version of syrk_ema_kernel: same tiling / split-K / epilogue semantics, the 128x128
// tile product runs on tcgen05 (3xTF32, fp32 accumulation in TMEM) through tcg::mainloop.  Diagonal
// tiles stage the operand once and use it for both sides.  The drained accumulator is read row-wise
// for the (m, c) store and column-wise (conflict-free) for the mirrored (c, m) store: both coalesced.
template <class L>
__global__ void __launch_bounds__(tcg::THREADS)
tc_syrk_ema_kernel(L ld, int n, int K, int kchunk, float scale, float alpha, int first_step,
                   float* __restrict__ factor, float* __restrict__ partial) {
    extern __shared__ __align__(128) unsigned char dsm_raw[];
    __shared__ __align__(8) uint64_t mbar;
    __shared__ uint32_t tmem_slot;
    tcg::Ctx c;
    tcg::setup(c, dsm_raw, &mbar, &tmem_slot);
    const int T = (n + TILE - 1) / TILE;
    int ti, tj;
    tri_decode(blockIdx.x, T, ti, tj);
    const int kbeg = blockIdx.y * kchunk;
    const int kend = min(K, kbeg + kchunk);
    tcg::mainloop<false>(ld, ld, ti * TILE, tj * TILE, kbeg, kend, ti == tj, c);    // chain bounded by kchunk (<= ~512)
    const float* d = tcg::drain<false>(c, kend > kbeg);
    const float keep = 1.f - alpha;
    const int m0 = ti * TILE, c0 = tj * TILE;
    if (gridDim.y == 1) {
        for (int e = threadIdx.x; e < TILE * TILE; e += tcg::THREADS) {
            const int m = m0 + (e >> 7), cc = c0 + (e & 127);
            if (m < n && cc < n) {
                const float old = first_step ? (m == cc ? 1.f : 0.f) : factor[(long)m * n + cc];
                // diagonal tiles: take both (m, cc) and (cc, m) from the upper triangle => exactly symmetric factor
                const int r = e >> 7, q = e & 127;
                const float v = (ti == tj && r > q) ? d[q * tcg::DPITCH + r] : d[r * tcg::DPITCH + q];
                factor[(long)m * n + cc] = fmaf(alpha, scale * v, keep * old);
            }
        }
        if (ti != tj) {          // mirrored tile: element (cc, m) takes the value computed for (m, cc)
            for (int e = threadIdx.x; e < TILE * TILE; e += tcg::THREADS) {
                const int cc = c0 + (e >> 7), m = m0 + (e & 127);
                if (m < n && cc < n) {
                    const float old = first_step ? 0.f : factor[(long)cc * n + m];
                    factor[(long)cc * n + m] = fmaf(alpha, scale * d[(e & 127) * tcg::DPITCH + (e >> 7)], keep * old);
                }
            }
        }
    } else {
        float* dst = partial + (long)blockIdx.y * n * n;
        for (int e = threadIdx.x; e < TILE * TILE; e += tcg::THREADS) {
            const int m = m0 + (e >> 7), cc = c0 + (e & 127);
            if (m < n && cc < n) dst[(long)m * n + cc] = d[(e >> 7) * tcg::DPITCH + (e & 127)];
        }
    }
    tcg::teardown(c);
}

// second stage of the split-K SYRK: fixed-order sum of the partials + scale + EMA + mirror
__global__ void syrk_reduce_ema_kernel(const float* __restrict__ partial, int ksplit, int n,
                                       float scale, float alpha, int first_step,
                                       float* __restrict__ factor) {
    const long idx = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (long)n * n) return;
    const int i = (int)(idx / n), j = (int)(idx - (long)i * n);
    // read the upper-triangular element (tile-wise, and element-wise inside diagonal tiles): exact symmetry
    const bool upper = (i / TILE < j / TILE) || (i / TILE == j / TILE && i <= j);
    const int r = upper ? i : j;
    const int c = upper ? j : i;
    // up to ~800 partials (conv1: 401 408 rows in chunks of 512): summed in double -- a sequential fp32 sum loses
    // ~sqrt(ksplit) ulp, which the conditioning of the preconditioner amplifies (the kernel is memory bound anyway)
    double s = 0.0;
    const float* p = partial + (long)r * n + c;
    const long stride = (long)n * n;
    int z = 0;
    for (; z + 4 <= ksplit; z += 4) {                              // 4 loads in flight
        const float a0 = p[(long)z * stride], a1 = p[(long)(z + 1) * stride], a2 = p[(long)(z + 2) * stride], a3 = p[(long)(z + 3) * stride];
        s += ((double)a0 + (double)a1) + ((double)a2 + (double)a3);
    }
    for (; z < ksplit; ++z) s += (double)p[(long)z * stride];
    const float old = first_step ? (i == j ? 1.f : 0.f) : factor[idx];
    factor[idx] = fmaf(alpha, scale * (float)s, (1.f - alpha) * old);
}

// mean over the sequence axis: in (B, T, D) -> out (B, D)   (kfac/utils.py:98-99,133-134)
// grid (ceil(D/64), B); 256 threads = 8 warps striding over t, lanes own 2 consecutive d
__global__ void seq_mean_kernel(const float* __restrict__ in, float* __restrict__ out,
                                int T, int D) {
    __shared__ float red[8][64];
    const int b = blockIdx.y;
    const int d0 = blockIdx.x * 64 + (threadIdx.x & 31) * 2;
    const int w = threadIdx.x >> 5;
    float s0 = 0.f, s1 = 0.f;
    const float* base = in + (long)b * T * D;
    if (d0 + 1 < D && (D % 2 == 0)) {
        for (int t = w; t < T; t += 8) {
            const float2 v = __ldg(reinterpret_cast<const float2*>(base + (long)t * D + d0));
            s0 += v.x; s1 += v.y;
        }
    } else {
        for (int t = w; t < T; t += 8) {
            if (d0 < D) s0 += __ldg(base + (long)t * D + d0);
            if (d0 + 1 < D) s1 += __ldg(base + (long)t * D + d0 + 1);
        }
    }
    red[w][(threadIdx.x & 31) * 2] = s0;
    red[w][(threadIdx.x & 31) * 2 + 1] = s1;
    __syncthreads();
    if (threadIdx.x < 64) {
        float s = 0.f;
#pragma unroll
        for (int i = 0; i < 8; ++i) s += red[i][threadIdx.x];
        const int d = blockIdx.x * 64 + threadIdx.x;
        if (d < D) out[(long)b * D + d] = s / (float)T;
    }
}

// ---------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------
static constexpr size_t kMaxPartialBytes = 512ull << 20;

constexpr long TC_MAX_KCHUNK = 512;      // tensor-core path: bounds the round-toward-zero TMEM accumulation chain

static int choose_ksplit(int n, long K) {
    const int T = cdiv(n, TILE);
    const long tiles = (long)T * (T + 1) / 2;
    const long target = 3L * sm_count();            // ~3 waves of 2-CTA/SM
    long ks = target / tiles;
    if (use_tc_gemm() && K >= 64) { const long need = (K + TC_MAX_KCHUNK - 1) / TC_MAX_KCHUNK; if (ks < need) ks = need; }
    const long max_by_k = (K + 4 * BK - 1) / (4 * BK);   // at least 4 slabs per split
    if (ks > max_by_k) ks = max_by_k;
    const long max_by_ws = (long)(kMaxPartialBytes / ((size_t)n * n * sizeof(float)));
    if (ks > max_by_ws) ks = max_by_ws;
    if (ks > 1024) ks = 1024;
    if (ks < 1) ks = 1;
    return (int)ks;
}

static size_t syrk_partial_bytes(int n, long K) {
    const int ks = choose_ksplit(n, K);
    return ks > 1 ? align_up((size_t)ks * n * n * sizeof(float), 256) : 0;
}

template <class L>
static int run_syrk(const L& ld, int n, long K, float scale, float alpha, int first_step,
                    float* factor, float* partial, cudaStream_t st) {
    const int ks = choose_ksplit(n, K);
    int kchunk = (int)((K + ks - 1) / ks);
    kchunk = (kchunk + BK - 1) / BK * BK;
    const int ksplit = (int)((K + kchunk - 1) / kchunk);
    const int T = cdiv(n, TILE);
    dim3 grid(T * (T + 1) / 2, ksplit);
    // the tensor-core path bounds its in-TMEM accumulation chain by the k-chunk (round-toward-zero accumulation, tc_gemm.cuh);
    // if the workspace / split caps of choose_ksplit push the chunk beyond twice that bound (batches of millions of rows),
    // the exact-fp32 SIMT kernel takes over instead of silently accumulating a biased chain
    if (use_tc_gemm() && K >= 64 && kchunk <= 2 * TC_MAX_KCHUNK) {
        // (set per call: the attribute is per device, and one process may drive several GPUs)
        KFACB_CHECK_CUDA(cudaFuncSetAttribute(tc_syrk_ema_kernel<L>, cudaFuncAttributeMaxDynamicSharedMemorySize, tcg::SMEM));
        tc_syrk_ema_kernel<L><<<grid, tcg::THREADS, tcg::SMEM, st>>>(ld, n, (int)K, kchunk, scale, alpha,
                                                                   first_step, factor, partial);
    } else {
        syrk_ema_kernel<L><<<grid, GEMM_THREADS, 0, st>>>(ld, n, (int)K, kchunk, scale, alpha,
                                                          first_step, factor, partial);
    }
    KFACB_CHECK_LAUNCH();
    if (ksplit > 1) {
        const long tot = (long)n * n;
        syrk_reduce_ema_kernel<<<cdiv(tot, 256), 256, 0, st>>>(partial, ksplit, n, scale, alpha,
                                                               first_step, factor);
        KFACB_CHECK_LAUNCH();
    }
    return KFACB_OK;
}

static int validate_layer(const kfacb_layer_t& l, int idx) {
    KFACB_REQUIRE(l.kind == KFACB_LINEAR || l.kind == KFACB_CONV2D,
                  "layer %d: unsupported layer kind %d (only Linear and Conv2d)", idx, l.kind);
    KFACB_REQUIRE(l.batch > 0 && l.in_c > 0 && l.out_c > 0, "layer %d: bad dims", idx);
    if (l.kind == KFACB_CONV2D) {
        KFACB_REQUIRE(l.kh > 0 && l.kw > 0 && l.sh > 0 && l.sw > 0 && l.ph >= 0 && l.pw >= 0 &&
                      l.in_h > 0 && l.in_w > 0, "layer %d: bad conv geometry", idx);
        KFACB_REQUIRE(l.out_h == (l.in_h + 2 * l.ph - l.kh) / l.sh + 1 &&
                      l.out_w == (l.in_w + 2 * l.pw - l.kw) / l.sw + 1,
                      "layer %d: out_h/out_w inconsistent with conv geometry", idx);
        KFACB_REQUIRE((long)l.batch * l.out_h * l.out_w < (1L << 31) &&
                      (long)l.batch * l.in_h * l.in_w < (1L << 31), "layer %d: too many rows", idx);
    } else {
        KFACB_REQUIRE(l.seq >= 1, "layer %d: seq must be >= 1", idx);
    }
    return KFACB_OK;
}

// workspace layout per call: [seq-mean buffer][split-K partials]  (layers run one after another)
static void layer_ws_needs(const kfacb_layer_t& l, size_t& mean_bytes, size_t& part_bytes) {
    mean_bytes = 0; part_bytes = 0;
    const int dA = kfacb_layer_dA(&l), dG = kfacb_layer_dG(&l);
    if (l.kind == KFACB_LINEAR) {
        if (l.seq > 1)
            mean_bytes = align_up((size_t)l.batch * (l.in_c > l.out_c ? l.in_c : l.out_c) * sizeof(float), 256);
        part_bytes = syrk_partial_bytes(dA, l.batch);
        const size_t pg = syrk_partial_bytes(dG, l.batch);
        if (pg > part_bytes) part_bytes = pg;
    } else {
        const long rows = (long)l.batch * l.out_h * l.out_w;
        part_bytes = syrk_partial_bytes(dA, rows);
        const size_t pg = syrk_partial_bytes(dG, rows);
        if (pg > part_bytes) part_bytes = pg;
    }
}

}  // namespace kfacb

using namespace kfacb;

extern "C" int32_t kfacb_layer_dA(const kfacb_layer_t* l) {
    const int k = (l->kind == KFACB_CONV2D) ? l->kh * l->kw : 1;
    return l->in_c * k + (l->has_bias ? 1 : 0);
}
extern "C" int32_t kfacb_layer_dG(const kfacb_layer_t* l) { return l->out_c; }

extern "C" size_t kfacb_factor_workspace_bytes(const kfacb_layer_t* layers, int32_t count) {
    size_t mean_max = 0, part_max = 0;
    for (int i = 0; i < count; ++i) {
        size_t m, p;
        layer_ws_needs(layers[i], m, p);
        if (m > mean_max) mean_max = m;
        if (p > part_max) part_max = p;
    }
    return mean_max + part_max + 256;
}

extern "C" int kfacb_factor_update(const kfacb_layer_t* layers, int32_t count,
                                   const float* const* a, const float* const* g,
                                   float* const* A, float* const* G,
                                   float alpha, int32_t first_step,
                                   void* workspace, size_t workspace_bytes,
                                   kfacb_stream_t stream) {
    cudaStream_t st = (cudaStream_t)stream;
    KFACB_REQUIRE(layers && a && g && A && G && count >= 0, "null argument");
    if (workspace_bytes < kfacb_factor_workspace_bytes(layers, count) ||
        ((uintptr_t)workspace & 255)) {
        set_error("factor workspace too small or misaligned");
        return KFACB_ERR_WORKSPACE;
    }
    size_t mean_max = 0;
    for (int i = 0; i < count; ++i) {
        size_t m, p;
        layer_ws_needs(layers[i], m, p);
        if (m > mean_max) mean_max = m;
    }
    float* mean_buf = (float*)workspace;
    float* partial = (float*)((char*)workspace + mean_max);

    for (int i = 0; i < count; ++i) {
        const kfacb_layer_t& l = layers[i];
        int rc = validate_layer(l, i);
        if (rc) return rc;
        const int dA = kfacb_layer_dA(&l), dG = kfacb_layer_dG(&l);
        if (l.kind == KFACB_LINEAR) {
            const int B = l.batch;
            if (a[i]) {
                KFACB_REQUIRE(A[i], "layer %d: A is null", i);
                const float* src = a[i];
                if (l.seq > 1) {
                    dim3 grid(cdiv(l.in_c, 64), B);
                    seq_mean_kernel<<<grid, 256, 0, st>>>(a[i], mean_buf, l.seq, l.in_c);
                    KFACB_CHECK_LAUNCH();
                    src = mean_buf;
                }
                LdRows ld{src, l.in_c, l.has_bias ? l.in_c : -1, B};
                rc = run_syrk(ld, dA, B, (float)(1.0 / B), alpha, first_step, A[i], partial, st);
                if (rc) return rc;
            }
            if (g[i]) {
                KFACB_REQUIRE(G[i], "layer %d: G is null", i);
                const float* src = g[i];
                if (l.seq > 1) {
                    dim3 grid(cdiv(l.out_c, 64), B);
                    seq_mean_kernel<<<grid, 256, 0, st>>>(g[i], mean_buf, l.seq, l.out_c);
                    KFACB_CHECK_LAUNCH();
                    src = mean_buf;
                }
                LdRows ld{src, l.out_c, -1, B};
                rc = run_syrk(ld, dG, B, (float)B, alpha, first_step, G[i], partial, st);
                if (rc) return rc;
            }
        } else {
            const int S = l.out_h * l.out_w;
            const long rows = (long)l.batch * S;
            if (a[i]) {
                KFACB_REQUIRE(A[i], "layer %d: A is null", i);
                const float scale = (float)(1.0 / ((double)l.batch * (double)S * (double)S));
                const int ones = l.has_bias ? dA - 1 : -1;
                const bool pointwise = l.kh == 1 && l.kw == 1 && l.sh == 1 && l.sw == 1 &&
                                       l.ph == 0 && l.pw == 0;
                if (pointwise) {
                    LdNCS ld{a[i], l.in_c, S, ones, (int)rows, {}};
                    ld.divS.init(S);
                    rc = run_syrk(ld, dA, rows, scale, alpha, first_step, A[i], partial, st);
                } else {
                    LdIm2col ld;
                    ld.p = a[i]; ld.C = l.in_c; ld.H = l.in_h; ld.W = l.in_w; ld.KW = l.kw;
                    ld.KHW = l.kh * l.kw; ld.SH = l.sh; ld.SW = l.sw; ld.PH = l.ph; ld.PW = l.pw;
                    ld.OW = l.out_w; ld.S = S; ld.D = l.in_c * l.kh * l.kw; ld.ones_col = ones;
                    ld.K = (int)rows;
                    ld.divS.init(S); ld.divOW.init(l.out_w); ld.divKHW.init(l.kh * l.kw);
                    ld.divKW.init(l.kw);
                    rc = run_syrk(ld, dA, rows, scale, alpha, first_step, A[i], partial, st);
                }
                if (rc) return rc;
            }
            if (g[i]) {
                KFACB_REQUIRE(G[i], "layer %d: G is null", i);
                LdNCS ld{g[i], l.out_c, S, -1, (int)rows, {}};
                ld.divS.init(S);
                rc = run_syrk(ld, dG, rows, (float)((double)l.batch * (double)S), alpha,
                              first_step, G[i], partial, st);
                if (rc) return rc;
            }
        }
    }
    return KFACB_OK;
}
