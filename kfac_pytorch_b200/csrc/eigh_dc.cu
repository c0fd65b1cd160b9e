// eigh_dc.cu -- eigendecomposition of the n > 128 factors of kfacb_eigh_batched, stages 2-4:
//   (1) A = Q_H T Q_H^T          persistent team-scheduled Householder tridiagonalisation (eigh_trd.cu)
//   (2) T = Z diag(lam) Z^T      divide and conquer (Cuppen / Gu-Eisenstat) on the tridiagonal matrix, all merges of
//                                a tree level of ALL matrices of the batch in the same launches
//   (3) Q = Q_H Z                block-reflector back-transformation on the tensor cores (tcgen05 3xTF32)
//   (4) floor + write             d <- lam * (lam > eps), Q column-wise (kfac/utils.py:22-25, eigen_dp.py:62-75)
// Replaces the round-1 block-Jacobi path as the default: ~10x fewer flops, a FIXED launch sequence (no convergence
// test, no host synchronisation: the call only enqueues work) and LAPACK-class
// accuracy (reconstruction ~1e-6).  tools/eigh_trd_dc_model.py is the same algorithm in numpy.
//
// D&C, per merge of two solved children (d1, Z1), (d2, Z2) torn apart at sub-diagonal element e (rho = 2|e|,
// z = [last column of Z1^T ; sign(e) first column of Z2^T] / sqrt 2):
//   prep     sort the poles, LAPACK-style deflation scan (tiny z_i; nearly equal poles -> Givens rotation log)
//   rotate   apply the (rare) rotations to the eigenvector rows
//   secular  one WARP per root: the root of 1 + rho sum z_i^2 / (d_i - lam) is bracketed between two poles and
//            found by bisection ON THE FLOAT BIT PATTERN of its offset from the nearer pole: <= 32 steps give the
//            offset to the last bit whatever its magnitude (no convergence heuristics, branch-free inner loop)
//   zhat/vec Gu-Eisenstat: recompute z from the roots (Loewner), eigenvectors zhat_i / (d_i - lam_j) -- orthogonal
//            to working precision even for clustered roots
//   rank     final ascending position of every root / deflated value
//   gemm     new eigenvector rows = U^T x (gathered old rows): tcgen05 3xTF32, output rows scattered to their rank
//   copy     deflated rows move unchanged
// Eigenvector storage is ROW-wise (row = eigenvector, Zt) with ping-pong buffers, so gathers, rotations and the
// back-transformation's operands are all contiguous.
#include <stdlib.h>
#include <string.h>
#include <algorithm>
#include <vector>
#include "eigh_trd.cuh"
#include "jacobi_smem.cuh"
#include "tc_gemm.cuh"

namespace kfacb {
namespace dc {

constexpr int LEAF = 32;
constexpr int PB = 128;                  // reflectors per block reflector of the back-transformation
constexpr int YCH = 512;                 // split-K chunk of Y = Zt V^T (bounds the in-TMEM accumulation chain)
constexpr float FEPS = 5.9604645e-8f;    // 2^-24

struct DcMat {
    float* T; float* Z[2]; float* U; float* Ypart; float* Y2; float* Tb;
    float* d; float* e; float* tau;
    float* dcur[2]; float* dl; float* w; float* zh; float* lam; float* vals;
    float* rotcs;
    int* src; int* rnk; int* rotab;
    float* Q; float* dout;
    int n, ldz, npanels, fin;
};
struct DcLeaf { int mat, off, nl, pad; };
struct DcMerge { int mat, off, n1, n2; };
struct DcState { int K, nrot; float rho, sgn; };
struct Tile { int a, b, c, d; };

// ---------------------------------------------------------------------------
// leaves: dense Jacobi in shared memory on the torn tridiagonal block
// ---------------------------------------------------------------------------
constexpr int LEAF_MAX = LEAF + 8;
__global__ void __launch_bounds__(256)
dc_leaf_kernel(const DcMat* mats, const DcLeaf* leaves) {
    extern __shared__ float smem[];
    __shared__ float s_d[LEAF_MAX], s_e[LEAF_MAX], s_lam[LEAF_MAX], s_nrm[LEAF_MAX], s_scale;
    __shared__ int s_rank[LEAF_MAX];
    const DcLeaf lf = leaves[blockIdx.x];
    const DcMat mt = mats[lf.mat];
    const int nl = lf.nl, off = lf.off, m = (nl + 1) & ~1, P = m + 1;
    const int tid = threadIdx.x, nt = blockDim.x;
    float* S = smem;
    float* U = smem + (size_t)m * P;
    for (int i = tid; i < nl; i += nt) {
        float dd = mt.d[off + i];
        if (i == 0 && off > 0) dd -= fabsf(mt.e[off - 1]);                 // rank-one tears at both ends
        if (i == nl - 1 && off + nl < mt.n) dd -= fabsf(mt.e[off + nl - 1]);
        s_d[i] = dd;
        s_e[i] = (i < nl - 1) ? mt.e[off + i] : 0.f;
    }
    __syncthreads();
    for (int e = tid; e < m * m; e += nt) {
        const int i = e / m, j = e - i * m;
        float v = 0.f;
        if (i < nl && j < nl) {
            if (i == j) v = s_d[i];
            else if (j == i + 1) v = s_e[i];
            else if (i == j + 1) v = s_e[j];
        }
        S[i * P + j] = v;
        U[i * P + j] = (i == j) ? 1.f : 0.f;
    }
    if (tid == 0) {
        float sc = 0.f;
        for (int i = 0; i < nl; ++i) sc = fmaxf(sc, fabsf(s_d[i]));
        s_scale = sc;
    }
    __syncthreads();
    jacobi_smem<0>(S, U, m, P, 1.2e-7f, 1.5e-8f * s_scale, 1e30f, 30, nullptr);
    __syncthreads();
    // ~400 fp32 rotations per column leave U orthogonal only to ~4e-6; every merge above inherits that.  One
    // Newton-Schulz step  U <- U (I - (U^T U - I) / 2)  in shared memory (2 m^3 flops) brings it to ~2e-7.
    for (int e = tid; e < m * m; e += nt) {
        const int i = e / m, j = e - i * m;
        float acc = 0.f;
        for (int k = 0; k < m; ++k) acc = fmaf(U[k * P + i], U[k * P + j], acc);
        S[i * P + j] = acc - (i == j ? 1.f : 0.f);
    }
    __syncthreads();
    {
        constexpr int PER = (LEAF_MAX * LEAF_MAX + 255) / 256;
        float un[PER];
#pragma unroll
        for (int q = 0; q < PER; ++q) {
            const int e = tid + q * 256;
            un[q] = 0.f;
            if (e < m * m) {
                const int i = e / m, j = e - i * m;
                float acc = 0.f;
                for (int k = 0; k < m; ++k) acc = fmaf(U[i * P + k], S[k * P + j], acc);
                un[q] = fmaf(-0.5f, acc, U[i * P + j]);
            }
        }
        __syncthreads();
#pragma unroll
        for (int q = 0; q < PER; ++q) {
            const int e = tid + q * 256;
            if (e < m * m) U[(e / m) * P + (e % m)] = un[q];
        }
    }
    __syncthreads();
    // eigenvalues as Rayleigh quotients of the tridiagonal block (second-order accurate in the vector error)
    if (tid < nl) {
        double num = 0.0, den = 0.0;
        for (int i = 0; i < nl; ++i) {
            const float u = U[i * P + tid];
            float y = s_d[i] * u;
            if (i > 0) y = fmaf(s_e[i - 1], U[(i - 1) * P + tid], y);
            if (i < nl - 1) y = fmaf(s_e[i], U[(i + 1) * P + tid], y);
            num += (double)u * (double)y;
            den += (double)u * (double)u;
        }
        s_lam[tid] = (float)(num / den);
        s_nrm[tid] = (float)(1.0 / sqrt(den));
    }
    __syncthreads();
    if (tid < nl) {
        const float v = s_lam[tid];
        int rk = 0;
        for (int j = 0; j < nl; ++j) {
            const float w = s_lam[j];
            rk += (w < v) || (w == v && j < tid);
        }
        s_rank[tid] = rk;
        mt.dcur[0][off + rk] = v;
    }
    __syncthreads();
    float* Z = mt.Z[0];
    for (int e = tid; e < nl * nl; e += nt) {
        const int j = e / nl, i = e - j * nl;                                // eigenvector j, component i
        Z[(size_t)(off + s_rank[j]) * mt.ldz + off + i] = U[i * P + j] * s_nrm[j];
    }
}

// ---------------------------------------------------------------------------
// per-level kernels
// ---------------------------------------------------------------------------
// the children's blocks must be block-diagonal inside the merged node (stale data of two levels ago otherwise)
__global__ void dc_zero_kernel(const DcMat* mats, const DcMerge* merges, int in) {
    const DcMerge mg = merges[blockIdx.y];
    if (mg.n2 == 0) return;
    const int nm = mg.n1 + mg.n2;
    const int r = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (r >= nm) return;
    const DcMat mt = mats[mg.mat];
    float* row = mt.Z[in] + (size_t)(mg.off + r) * mt.ldz + mg.off;
    const int c0 = r < mg.n1 ? mg.n1 : 0, c1 = r < mg.n1 ? nm : mg.n1;
    for (int c = c0 + (threadIdx.x & 31); c < c1; c += 32) row[c] = 0.f;
}

__global__ void __launch_bounds__(1024)
dc_prep_kernel(const DcMat* mats, const DcMerge* merges, DcState* states, int in, int nmax) {
    extern __shared__ float smem[];
    __shared__ float s_red[3][32];
    const DcMerge mg = merges[blockIdx.x];
    const DcMat mt = mats[mg.mat];
    const int n1 = mg.n1, n2 = mg.n2, nm = n1 + n2, off = mg.off, tid = threadIdx.x;
    float* sd = smem;
    float* sz = smem + nmax;
    int* sord = reinterpret_cast<int*>(smem + 2 * nmax);
    const float* din = mt.dcur[in] + off;
    float* vals = mt.vals + off;
    int* src = mt.src + off;
    if (n2 == 0) {                                   // pass-through node: everything "deflated"
        for (int q = tid; q < nm; q += blockDim.x) { vals[q] = din[q]; src[q] = off + q; }
        if (tid == 0) states[blockIdx.x] = DcState{0, 0, 0.f, 1.f};
        return;
    }
    const float es = mt.e[off + n1 - 1];
    const float sgn = es >= 0.f ? 1.f : -1.f, rho = 2.f * fabsf(es);
    const float* Zin = mt.Z[in];
    float dmax = 0.f, zmax = 0.f, tmax = 0.f;
    // max-norm of the whole tridiagonal matrix: LAPACK (sstedc) scales the problem to unit norm before the merges, so
    // slaed2's tolerance 8 eps max(|d|max, |z|max) compares the dimensionless z with d ~ 1.  Unscaled, z has to carry
    // the scale of the matrix -- K-FAC's conv A factors are ~0.0025 I + .., and an absolute 8 eps |z|max deflated them
    // far too eagerly (measured: reconstruction 6e-5 instead of 1e-6 on ResNet-50 conv1 at step 1)
    for (int i = tid; i < mt.n; i += blockDim.x)
        tmax = fmaxf(tmax, fmaxf(fabsf(mt.d[i]), i + 1 < mt.n ? fabsf(mt.e[i]) : 0.f));
    for (int i = tid; i < nm; i += blockDim.x) {
        const float dv = din[i];
        const float zv = (i < n1 ? Zin[(size_t)(off + i) * mt.ldz + off + n1 - 1]
                                 : sgn * Zin[(size_t)(off + i) * mt.ldz + off + n1]) * 0.70710678118654752f;
        sd[i] = dv; sz[i] = zv;
        dmax = fmaxf(dmax, fabsf(dv)); zmax = fmaxf(zmax, fabsf(zv));
    }
    for (int o = 16; o > 0; o >>= 1) {
        dmax = fmaxf(dmax, __shfl_xor_sync(0xffffffffu, dmax, o));
        zmax = fmaxf(zmax, __shfl_xor_sync(0xffffffffu, zmax, o));
        tmax = fmaxf(tmax, __shfl_xor_sync(0xffffffffu, tmax, o));
    }
    if ((tid & 31) == 0) { s_red[0][tid >> 5] = dmax; s_red[1][tid >> 5] = zmax; s_red[2][tid >> 5] = tmax; }
    __syncthreads();
    // merge of the two sorted lists (stable: first child first on ties)
    for (int i = tid; i < nm; i += blockDim.x) {
        const float v = sd[i];
        int lo, hi;
        if (i < n1) { lo = n1; hi = nm; while (lo < hi) { const int mid = (lo + hi) >> 1; if (sd[mid] < v) lo = mid + 1; else hi = mid; } sord[i + (lo - n1)] = i; }
        else { lo = 0; hi = n1; while (lo < hi) { const int mid = (lo + hi) >> 1; if (sd[mid] <= v) lo = mid + 1; else hi = mid; } sord[(i - n1) + lo] = i; }
    }
    __syncthreads();
    if (tid != 0) return;
    dmax = 0.f; zmax = 0.f; tmax = 0.f;
    for (int i = 0; i < 32; ++i) { dmax = fmaxf(dmax, s_red[0][i]); zmax = fmaxf(zmax, s_red[1][i]); tmax = fmaxf(tmax, s_red[2][i]); }
    const float tol = 8.f * FEPS * fmaxf(dmax, tmax * zmax);
    float* dl = mt.dl + off;
    float* w = mt.w + off;
    int* rotab = mt.rotab + 2 * off;
    float* rotcs = mt.rotcs + 2 * off;
    int K = 0, nd = 0, nrot = 0, pj = -1;
    if (rho * zmax <= tol) {
        for (int p = 0; p < nm; ++p) { const int jj = sord[p]; vals[nm - 1 - nd] = sd[jj]; src[nm - 1 - nd] = off + jj; ++nd; }
    } else {
        for (int p = 0; p < nm; ++p) {
            const int jj = sord[p];
            if (rho * fabsf(sz[jj]) <= tol) { vals[nm - 1 - nd] = sd[jj]; src[nm - 1 - nd] = off + jj; ++nd; continue; }
            if (pj < 0) { pj = jj; continue; }
            float s = sz[pj], c = sz[jj];
            const float tau = sqrtf(fmaf(c, c, s * s));
            const float t = sd[jj] - sd[pj];
            c /= tau; s = -s / tau;
            if (fabsf(t * c * s) <= tol) {           // the two poles are numerically one: rotate z_pj into z_jj
                sz[jj] = tau; sz[pj] = 0.f;
                rotab[2 * nrot] = off + pj; rotab[2 * nrot + 1] = off + jj;
                rotcs[2 * nrot] = c; rotcs[2 * nrot + 1] = s; ++nrot;
                const float tn = sd[pj] * c * c + sd[jj] * s * s;
                sd[jj] = sd[pj] * s * s + sd[jj] * c * c;
                sd[pj] = tn;
                vals[nm - 1 - nd] = tn; src[nm - 1 - nd] = off + pj; ++nd;
                pj = jj;
            } else {
                dl[K] = sd[pj]; w[K] = sz[pj]; src[K] = off + pj; ++K;
                pj = jj;
            }
        }
        if (pj >= 0) { dl[K] = sd[pj]; w[K] = sz[pj]; src[K] = off + pj; ++K; }
    }
    states[blockIdx.x] = DcState{K, nrot, rho, sgn};
}

// rows (a, b) <- (c a + s b, c b - s a), in log order; one thread per column
__global__ void dc_rotate_kernel(const DcMat* mats, const DcMerge* merges, const DcState* states, int in) {
    const DcState st = states[blockIdx.y];
    if (st.nrot == 0) return;
    const DcMerge mg = merges[blockIdx.y];
    const int nm = mg.n1 + mg.n2;
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= nm) return;
    const DcMat mt = mats[mg.mat];
    float* Z = mt.Z[in] + mg.off + c;
    const int* rotab = mt.rotab + 2 * mg.off;
    const float* rotcs = mt.rotcs + 2 * mg.off;
    for (int r = 0; r < st.nrot; ++r) {
        const int ra = rotab[2 * r], rb = rotab[2 * r + 1];
        const float cc = rotcs[2 * r], ss = rotcs[2 * r + 1];
        const float a = Z[(size_t)ra * mt.ldz], b = Z[(size_t)rb * mt.ldz];
        Z[(size_t)ra * mt.ldz] = cc * a + ss * b;
        Z[(size_t)rb * mt.ldz] = cc * b - ss * a;
    }
}

// one warp per root
__global__ void __launch_bounds__(256)
dc_secular_kernel(const DcMat* mats, const DcMerge* merges, const DcState* states, int nmax) {
    extern __shared__ float smem[];
    const DcState st = states[blockIdx.y];
    const int K = st.K;
    if ((int)blockIdx.x * 8 >= K) return;
    const DcMerge mg = merges[blockIdx.y];
    const DcMat mt = mats[mg.mat];
    const int off = mg.off, tid = threadIdx.x, lane = tid & 31;
    float* sdl = smem;
    float* sw2 = smem + nmax;
    for (int i = tid; i < K; i += blockDim.x) { sdl[i] = mt.dl[off + i]; const float wv = mt.w[off + i]; sw2[i] = wv * wv; }
    __syncthreads();
    const int j = blockIdx.x * 8 + (tid >> 5);
    if (j >= K) return;
    const float rho = st.rho;
    int o; float sg, hi;
    if (j < K - 1) {
        const float dj = sdl[j];
        const float half = 0.5f * (sdl[j + 1] - dj);
        float acc = 0.f;
        for (int i = lane; i < K; i += 32) acc += sw2[i] / ((sdl[i] - dj) - half);
        acc = warp_sum(acc);
        const float g = fmaf(rho, acc, 1.f);
        if (g > 0.f) { o = j; sg = 1.f; } else { o = j + 1; sg = -1.f; }
        hi = half;
    } else {
        float acc = 0.f;
        for (int i = lane; i < K; i += 32) acc += sw2[i];
        acc = warp_sum(acc);
        o = j; sg = 1.f; hi = rho * acc;
    }
    const float dorg = sdl[o];
    unsigned lo_b = 1u, hi_b = __float_as_uint(hi);
    for (int it = 0; it < 34 && hi_b - lo_b > 1u; ++it) {
        const unsigned mid_b = lo_b + ((hi_b - lo_b) >> 1);
        const float x = sg * __uint_as_float(mid_b);
        float acc = 0.f;
        for (int i = lane; i < K; i += 32) acc += sw2[i] / ((sdl[i] - dorg) - x);
        acc = warp_sum(acc);
        const float g = fmaf(rho, acc, 1.f);
        const bool too_small = sg > 0.f ? (g < 0.f) : (g > 0.f);
        if (too_small) lo_b = mid_b; else hi_b = mid_b;
    }
    const float mu = sg * __uint_as_float(hi_b);
    float* urow = mt.U + (size_t)(off + j) * mt.ldz + off;
    for (int i = lane; i < K; i += 32) urow[i] = (sdl[i] - dorg) - mu;
    if (lane == 0) mt.lam[off + j] = dorg + mu;
}

// zhat_i = sign(w_i) sqrt( prod_j |d_i - lam_j| / prod_{j != i} |d_i - d_j| ); block = 32 i x 32 j-lanes
__global__ void __launch_bounds__(1024)
dc_zhat_kernel(const DcMat* mats, const DcMerge* merges, const DcState* states) {
    __shared__ float red[32][33];
    const int K = states[blockIdx.y].K;
    if ((int)blockIdx.x * 32 >= K) return;
    const DcMerge mg = merges[blockIdx.y];
    const DcMat mt = mats[mg.mat];
    const int off = mg.off;
    const int il = threadIdx.x & 31, jl = threadIdx.x >> 5;
    const int i = blockIdx.x * 32 + il;
    float pr = 1.f;
    if (i < K) {
        const float di = mt.dl[off + i];
        const float* ucol = mt.U + (size_t)off * mt.ldz + off + i;
        for (int j = jl; j < K; j += 32) {
            const float num = fabsf(ucol[(size_t)j * mt.ldz]);
            pr *= (j == i) ? num : num / fabsf(di - mt.dl[off + j]);
        }
    }
    red[jl][il] = pr;
    __syncthreads();
    if (jl == 0 && i < K) {
        float p = 1.f;
        for (int q = 0; q < 32; ++q) p *= red[q][il];
        mt.zh[off + i] = copysignf(sqrtf(p), mt.w[off + i]);
    }
}

// U[j][i] <- zhat_i / delta[j][i], normalised; one warp per root
__global__ void __launch_bounds__(256)
dc_vec_kernel(const DcMat* mats, const DcMerge* merges, const DcState* states) {
    const int K = states[blockIdx.y].K;
    const int j = blockIdx.x * 8 + (threadIdx.x >> 5);
    if (j >= K) return;
    const DcMerge mg = merges[blockIdx.y];
    const DcMat mt = mats[mg.mat];
    const int off = mg.off, lane = threadIdx.x & 31;
    float* urow = mt.U + (size_t)(off + j) * mt.ldz + off;
    const float* zh = mt.zh + off;
    float ss = 0.f;
    for (int i = lane; i < K; i += 32) { const float v = zh[i] / urow[i]; ss = fmaf(v, v, ss); }
    ss = warp_sum(ss);
    const float inv = 1.f / sqrtf(ss);
    for (int i = lane; i < K; i += 32) urow[i] = (zh[i] / urow[i]) * inv;
}

// final ascending position of the K roots (q < K) and of the deflated values (q >= K); roots first on ties.
// grid (ceil(nm / 256), merges)
__global__ void __launch_bounds__(256)
dc_rank_kernel(const DcMat* mats, const DcMerge* merges, const DcState* states, int out, int nmax) {
    extern __shared__ float smem[];
    const DcMerge mg = merges[blockIdx.y];
    const int nm = mg.n1 + mg.n2;
    if ((int)blockIdx.x * 256 >= nm) return;
    const DcMat mt = mats[mg.mat];
    const int off = mg.off, K = states[blockIdx.y].K, tid = threadIdx.x;
    float* sv = smem;                                      // [0,K): roots, [K,nm): deflated values
    for (int q = tid; q < nm; q += blockDim.x) sv[q] = q < K ? mt.lam[off + q] : mt.vals[off + q];
    __syncthreads();
    const int q = blockIdx.x * 256 + tid;
    if (q >= nm) return;
    const float v = sv[q];
    int rk;
    if (q < K) {
        rk = q;
        for (int p = K; p < nm; ++p) rk += sv[p] < v;
    } else {
        int lo = 0, hi = K;
        while (lo < hi) { const int mid = (lo + hi) >> 1; if (sv[mid] <= v) lo = mid + 1; else hi = mid; }
        rk = lo;
        for (int p = K; p < nm; ++p) rk += (sv[p] < v) || (sv[p] == v && p < q);
    }
    mt.rnk[off + q] = rk;
    mt.dcur[out][off + rk] = v;
}

// One 128x128 output tile  sum_{k in [kbeg,kend)} A(m0+., k) B(n0+., k), handed element-wise to epi(m, n, value)
// (local indices; lanes walk n).  tc: the high-precision tcgen05 3xTF32 core (separate accumulators for the main and
// the correction terms, main accumulator flushed to fp32 registers every k-slab pair -- tc_gemm.cuh);
// !tc: exact-fp32 SIMT core (small matrices; everything with KFACB_GEMM_SIMT=1 for A/B accuracy comparisons).
template <class LA, class LB, class Epi>
__device__ __forceinline__ void tile_product(bool tc, const LA& la, const LB& lb, int m0, int n0, int kbeg, int kend,
                                             bool same_operand, const Epi& epi) {
    extern __shared__ __align__(128) unsigned char dsm_raw[];
    if (tc) {
        __shared__ __align__(8) uint64_t mbar;
        __shared__ uint32_t tmem_slot;
        tcg::Ctx c;
        tcg::setup_hp(c, dsm_raw, &mbar, &tmem_slot);           // two-accumulator, frequently flushed core (tc_gemm.cuh)
        tcg::mainloop_hp(la, lb, m0, n0, kbeg, kend, same_operand, c);
        const float* d = tcg::drain_hp(c, kend > kbeg);
        for (int e = threadIdx.x; e < TILE * TILE; e += tcg::THREADS) epi(e >> 7, e & 127, d[(e >> 7) * tcg::DPITCH + (e & 127)]);
        tcg::teardown_hp(c);
    } else {
        GemmSmem& sm = *reinterpret_cast<GemmSmem*>(dsm_raw);
        float acc[8][8];
#pragma unroll
        for (int i = 0; i < 8; ++i)
#pragma unroll
            for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;
        gemm_mainloop(la, lb, m0, n0, kbeg, kend, sm, acc);
#pragma unroll
        for (int i = 0; i < 8; ++i)
#pragma unroll
            for (int j = 0; j < 8; ++j) epi(frag_row(i), frag_col(j), acc[i][j]);
    }
}
// Same, for a read-modify-write epilogue  *addr(m, n) -= value  (addr returns nullptr outside the matrix): the old
// values of a batch are loaded BEFORE its stores -- written as `Z[i] -= v` per element the compiler must assume that a
// store aliases the next load and serialises one L2 round trip per element (measured: 40 us per 128x128 tile).
template <class LA, class LB, class Addr>
__device__ __forceinline__ void tile_product_sub(bool tc, const LA& la, const LB& lb, int m0, int n0, int kbeg, int kend,
                                                 const Addr& addr) {
    extern __shared__ __align__(128) unsigned char dsm_raw[];
    if (tc) {
        __shared__ __align__(8) uint64_t mbar;
        __shared__ uint32_t tmem_slot;
        tcg::Ctx c;
        tcg::setup_hp(c, dsm_raw, &mbar, &tmem_slot);
        tcg::mainloop_hp(la, lb, m0, n0, kbeg, kend, false, c);
        const float* d = tcg::drain_hp(c, kend > kbeg);
        constexpr int BATCH = 16;
        for (int e0 = threadIdx.x; e0 < TILE * TILE; e0 += tcg::THREADS * BATCH) {
            float* p[BATCH]; float old[BATCH];
#pragma unroll
            for (int b = 0; b < BATCH; ++b) {
                const int e = e0 + b * tcg::THREADS;
                p[b] = addr(e >> 7, e & 127);
                old[b] = p[b] ? *p[b] : 0.f;
            }
#pragma unroll
            for (int b = 0; b < BATCH; ++b) {
                const int e = e0 + b * tcg::THREADS;
                if (p[b]) *p[b] = old[b] - d[(e >> 7) * tcg::DPITCH + (e & 127)];
            }
        }
        tcg::teardown_hp(c);
    } else {
        GemmSmem& sm = *reinterpret_cast<GemmSmem*>(dsm_raw);
        float acc[8][8];
#pragma unroll
        for (int i = 0; i < 8; ++i)
#pragma unroll
            for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;
        gemm_mainloop(la, lb, m0, n0, kbeg, kend, sm, acc);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            float* p[8]; float old[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) { p[j] = addr(frag_row(i), frag_col(j)); old[j] = p[j] ? *p[j] : 0.f; }
#pragma unroll
            for (int j = 0; j < 8; ++j) if (p[j]) *p[j] = old[j] - acc[i][j];
        }
    }
}
static_assert(sizeof(GemmSmem) <= tcg::SMEM, "the SIMT core runs in the tensor-core core's dynamic shared memory");
// Matrices up to this size run their GEMM-shaped stages on the exact-fp32 SIMT core: their flops are negligible
// (sum n^3 over them is < 3 % of a ResNet-50 step) and the small, tightly clustered factors (conv1: n = 147 with
// 4e5 samples) are where the preconditioned gradient is most sensitive to the orthogonality of Q.
constexpr int SIMT_MAX_N = 640;
__device__ __forceinline__ bool use_tc_for(int n, int force) { return force == 0 ? n > SIMT_MAX_N : force > 0; }
static_assert(tcg::THREADS == GEMM_THREADS, "the tile kernels run either core with the same block size");

// ---- GEMM-shaped kernels ----
struct LdUt {                                              // A(m = new vector j, k = old vector i) = U[j][i]
    static constexpr bool K_FAST = true;
    const float* U; long ld; int K;
    __device__ __forceinline__ float operator()(int mn, int k) const { return (mn < K && k < K) ? __ldg(U + (long)mn * ld + k) : 0.f; }
    __device__ __forceinline__ float4 quad(int mn, int k) const {
        if (mn >= K || k >= K) return make_float4(0.f, 0.f, 0.f, 0.f);
        const float* p = U + (long)mn * ld + k;
        if (k + 3 < K) return __ldg(reinterpret_cast<const float4*>(p));
        return make_float4(__ldg(p), k + 1 < K ? __ldg(p + 1) : 0.f, k + 2 < K ? __ldg(p + 2) : 0.f, 0.f);
    }
};
struct LdZsrc {                                            // B(n = component c, k = old vector i) = Zin[src[i]][off + c]
    static constexpr bool K_FAST = false;
    const float* Z; long ld; const int* src; int nm, K;   // Z already offset by the node's column origin
    __device__ __forceinline__ float operator()(int mn, int k) const { return (mn < nm && k < K) ? __ldg(Z + (long)__ldg(src + k) * ld + mn) : 0.f; }
    __device__ __forceinline__ float4 quad(int mn, int k) const {
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (mn < nm) {
            if (k < K) v.x = __ldg(Z + (long)__ldg(src + k) * ld + mn);
            if (k + 1 < K) v.y = __ldg(Z + (long)__ldg(src + k + 1) * ld + mn);
            if (k + 2 < K) v.z = __ldg(Z + (long)__ldg(src + k + 2) * ld + mn);
            if (k + 3 < K) v.w = __ldg(Z + (long)__ldg(src + k + 3) * ld + mn);
        }
        return v;
    }
};

__global__ void __launch_bounds__(tcg::THREADS)
dc_gemm_kernel(const DcMat* mats, const DcMerge* merges, const DcState* states, const Tile* tiles, int in, int force) {
    const Tile tl = tiles[blockIdx.x];                     // a = merge (within the level), b = tile m, c = tile n
    const int K = states[tl.a].K;
    if (tl.b * TILE >= K) return;
    const DcMerge mg = merges[tl.a];
    const DcMat mt = mats[mg.mat];
    const int nm = mg.n1 + mg.n2, off = mg.off;
    LdUt la{mt.U + (size_t)off * mt.ldz + off, mt.ldz, K};
    LdZsrc lb{mt.Z[in] + off, mt.ldz, mt.src + off, nm, K};
    const int m0 = tl.b * TILE, n0 = tl.c * TILE;
    float* Zout = mt.Z[in ^ 1];
    const int* rnk = mt.rnk + off;
    const long ldz = mt.ldz;
    tile_product(use_tc_for(mt.n, force), la, lb, m0, n0, 0, K, false, [=](int m, int n, float v) {
        if (m0 + m < K && n0 + n < nm) Zout[(size_t)(off + rnk[m0 + m]) * ldz + off + n0 + n] = v;
    });
}

// deflated rows (and whole pass-through nodes) move to their rank unchanged; one warp per row
__global__ void dc_copy_kernel(const DcMat* mats, const DcMerge* merges, const DcState* states, int in) {
    const DcMerge mg = merges[blockIdx.y];
    const int nm = mg.n1 + mg.n2, K = states[blockIdx.y].K;
    const int q = K + blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (q >= nm) return;
    const DcMat mt = mats[mg.mat];
    const float* srow = mt.Z[in] + (size_t)mt.src[mg.off + q] * mt.ldz + mg.off;
    float* drow = mt.Z[in ^ 1] + (size_t)(mg.off + mt.rnk[mg.off + q]) * mt.ldz + mg.off;
    for (int c = threadIdx.x & 31; c < nm; c += 32) drow[c] = srow[c];
}

// ---------------------------------------------------------------------------
// back-transformation  Zt <- Zt B_p^T,  B_p = I - V_p T_p V_p^T,  p = last .. first
// ---------------------------------------------------------------------------
struct LdVt {                                              // (r, c) -> component c of reflector jb + r  (k = c contiguous)
    static constexpr bool K_FAST = true;
    const float* T; long ld; int jb, kp, n;               // T already offset to row jb
    __device__ __forceinline__ float operator()(int mn, int k) const {
        return (mn < kp && k > jb + mn && k < n) ? __ldg(T + (long)mn * ld + k) : 0.f;
    }
    __device__ __forceinline__ float4 quad(int mn, int k) const {
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (mn >= kp || k + 3 <= jb + mn || k >= n) return v;
        const float* p = T + (long)mn * ld + k;
        if (k > jb + mn && k + 3 < n) return __ldg(reinterpret_cast<const float4*>(p));
        const int lo = jb + mn;
        if (k > lo && k < n) v.x = __ldg(p);
        if (k + 1 > lo && k + 1 < n) v.y = __ldg(p + 1);
        if (k + 2 > lo && k + 2 < n) v.z = __ldg(p + 2);
        if (k + 3 > lo && k + 3 < n) v.w = __ldg(p + 3);
        return v;
    }
};
struct LdVtMn {                                            // (c - c_base, r): same data, c contiguous across lanes
    static constexpr bool K_FAST = false;
    const float* T; long ld; int jb, kp, n, cbase;
    __device__ __forceinline__ float operator()(int mn, int k) const {
        const int c = cbase + mn;
        return (k < kp && c > jb + k && c < n) ? __ldg(T + (long)k * ld + c) : 0.f;
    }
    __device__ __forceinline__ float4 quad(int mn, int k) const {
        return make_float4((*this)(mn, k), (*this)(mn, k + 1), (*this)(mn, k + 2), (*this)(mn, k + 3));
    }
};
struct LdZt {                                              // (row m of Zt, component c)
    static constexpr bool K_FAST = true;
    const float* Z; long ld; int n;
    __device__ __forceinline__ float operator()(int mn, int k) const { return (mn < n && k < n) ? __ldg(Z + (long)mn * ld + k) : 0.f; }
    __device__ __forceinline__ float4 quad(int mn, int k) const {
        if (mn >= n || k >= n) return make_float4(0.f, 0.f, 0.f, 0.f);
        const float* p = Z + (long)mn * ld + k;
        if (k + 3 < n) return __ldg(reinterpret_cast<const float4*>(p));
        return make_float4(__ldg(p), k + 1 < n ? __ldg(p + 1) : 0.f, k + 2 < n ? __ldg(p + 2) : 0.f, 0.f);
    }
};
struct LdYsum {                                            // (m, r) -> sum over the split-K chunks of Y
    static constexpr bool K_FAST = true;
    const float* Y; int n, nks;
    __device__ __forceinline__ float operator()(int mn, int k) const {
        if (mn >= n) return 0.f;
        float s = 0.f;
        for (int q = 0; q < nks; ++q) s += __ldg(Y + ((size_t)q * n + mn) * PB + k);
        return s;
    }
    __device__ __forceinline__ float4 quad(int mn, int k) const {
        float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
        if (mn >= n) return s;
        for (int q = 0; q < nks; ++q) {
            const float4 t = __ldg(reinterpret_cast<const float4*>(Y + ((size_t)q * n + mn) * PB + k));
            s.x += t.x; s.y += t.y; s.z += t.z; s.w += t.w;
        }
        return s;
    }
};

// G = V_p V_p^T (128 x 128) of every panel of every matrix; tiles: a = matrix, b = panel.
// EXACT fp32 (SIMT core): T_p is the inverse of a matrix built from G, and an error dG shows up in the
// orthogonality of the block reflector amplified by ||V_p||^2 ~ 1e1..1e2 -- 3xTF32 (1e-6) is not enough here.
__global__ void __launch_bounds__(GEMM_THREADS, 2)
bt_gram_kernel(const DcMat* mats, const Tile* tiles) {
    __shared__ GemmSmem sm;
    const Tile tl = tiles[blockIdx.x];
    const DcMat mt = mats[tl.a];
    const int jb = tl.b * PB, kp = min(PB, mt.n - 1 - jb);
    LdVt lv{mt.T + (size_t)jb * mt.ldz, mt.ldz, jb, kp, mt.n};
    float acc[8][8];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;
    gemm_mainloop(lv, lv, 0, 0, jb, mt.n, sm, acc);
    float* G = mt.Tb + (size_t)tl.b * PB * PB;
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) G[frag_row(i) * PB + frag_col(j)] = acc[i][j];
}

// T_p = (diag(1/tau) + striu(G))^-1 in place (upper triangular; dead reflectors, tau = 0, drop out)
constexpr int TP = PB + 1;
__global__ void __launch_bounds__(PB)
bt_tinv_kernel(const DcMat* mats, const Tile* tiles) {
    extern __shared__ float smem[];
    const Tile tl = tiles[blockIdx.x];
    const DcMat mt = mats[tl.a];
    const int jb = tl.b * PB, kp = min(PB, mt.n - 1 - jb), c = threadIdx.x;
    float* sT = smem;
    float* sX = smem + PB * TP;
    __shared__ float s_tau[PB];
    float* G = mt.Tb + (size_t)tl.b * PB * PB;
    s_tau[c] = c < kp ? mt.tau[jb + c] : 0.f;
    __syncthreads();
    for (int r = 0; r < PB; ++r) {
        const bool live = s_tau[r] != 0.f && s_tau[c] != 0.f;
        float v = 0.f;
        if (c > r) v = live ? G[r * PB + c] : 0.f;
        else if (c == r) v = s_tau[r] != 0.f ? 1.f / s_tau[r] : 1.f;
        sT[r * TP + c] = v;
        sX[r * TP + c] = 0.f;
    }
    __syncthreads();
    sX[c * TP + c] = 1.f / sT[c * TP + c];
    for (int r = PB - 2; r >= 0; --r) {
        if (c > r) {
            float acc = 0.f;
            for (int k = r + 1; k <= c; ++k) acc = fmaf(sT[r * TP + k], sX[k * TP + c], acc);
            sX[r * TP + c] = -acc / sT[r * TP + r];
        }
    }
    __syncthreads();
    for (int r = 0; r < PB; ++r) {
        const bool live = s_tau[r] != 0.f && s_tau[c] != 0.f;
        G[r * PB + c] = (c >= r && live) ? sX[r * TP + c] : 0.f;
    }
}

// Ypart[ks] = Zt[:, chunk ks] V_p[:, chunk ks]^T; tiles: a = matrix, b = tile m, c = chunk, d = panel
__global__ void __launch_bounds__(tcg::THREADS)
bt_y_kernel(const DcMat* mats, const Tile* tiles, int force) {
    const Tile tl = tiles[blockIdx.x];
    const DcMat mt = mats[tl.a];
    const int jb = tl.d * PB, kp = min(PB, mt.n - 1 - jb);
    const int kbeg = jb + tl.c * YCH, kend = min(mt.ldz, kbeg + YCH);
    LdZt la{mt.Z[mt.fin], mt.ldz, mt.n};
    LdVt lb{mt.T + (size_t)jb * mt.ldz, mt.ldz, jb, kp, mt.n};
    const int m0 = tl.b * TILE, n = mt.n;
    float* Y = mt.Ypart + (size_t)tl.c * mt.n * PB;
    tile_product(use_tc_for(mt.n, force), la, lb, m0, 0, kbeg, kend, false, [=](int m, int r, float v) {
        if (m0 + m < n) Y[(size_t)(m0 + m) * PB + r] = v;
    });
}

// Y2 = (sum_ks Ypart) T_p^T; tiles: a = matrix, b = tile m, c = number of chunks, d = panel
struct LdTb {
    static constexpr bool K_FAST = true;
    const float* Tb;
    __device__ __forceinline__ float operator()(int mn, int k) const { return __ldg(Tb + mn * PB + k); }
    __device__ __forceinline__ float4 quad(int mn, int k) const { return __ldg(reinterpret_cast<const float4*>(Tb + mn * PB + k)); }
};
__global__ void __launch_bounds__(tcg::THREADS)
bt_yt_kernel(const DcMat* mats, const Tile* tiles, int force) {
    const Tile tl = tiles[blockIdx.x];
    const DcMat mt = mats[tl.a];
    LdYsum la{mt.Ypart, mt.n, tl.c};
    LdTb lb{mt.Tb + (size_t)tl.d * PB * PB};
    const int m0 = tl.b * TILE, n = mt.n;
    float* Y2 = mt.Y2;
    tile_product(use_tc_for(mt.n, force), la, lb, m0, 0, 0, PB, false, [=](int m, int r, float v) {
        if (m0 + m < n) Y2[(size_t)(m0 + m) * PB + r] = v;
    });
}

// Zt[:, jb + 128 tn ..] -= Y2 V_p; tiles: a = matrix, b = tile m, c = tile n, d = panel
__global__ void __launch_bounds__(tcg::THREADS)
bt_upd_kernel(const DcMat* mats, const Tile* tiles, int force) {
    const Tile tl = tiles[blockIdx.x];
    const DcMat mt = mats[tl.a];
    const int jb = tl.d * PB, kp = min(PB, mt.n - 1 - jb);
    LdKContig la{mt.Y2, PB, mt.n, PB};
    LdVtMn lb{mt.T + (size_t)jb * mt.ldz, mt.ldz, jb, kp, mt.n, jb};
    const int m0 = tl.b * TILE, n0 = tl.c * TILE, n = mt.n;
    float* Z = mt.Z[mt.fin];
    const long ldz = mt.ldz;
    tile_product_sub(use_tc_for(mt.n, force), la, lb, m0, n0, 0, PB, [=](int m, int c) -> float* {
        const int cc = jb + n0 + c;
        return (m0 + m < n && cc < n) ? Z + (size_t)(m0 + m) * ldz + cc : nullptr;
    });
}

// Q[r][c] = Zt[c][r], d = lam * (lam > eps).  One CTA per 32x32 tile; tile_prefix[a] = tiles of the matrices before a
// (the CTA finds its matrix by bisection: no per-tile table, no empty CTAs).
__global__ void eig_write_kernel(const DcMat* mats, const int* tile_prefix, int nb, float eps_floor, int32_t* sweeps_out,
                                 const int* out_index) {
    __shared__ float t[32][33];
    int lo = 0, hi = nb - 1;
    while (lo < hi) { const int mid = (lo + hi + 1) >> 1; if (tile_prefix[mid] <= (int)blockIdx.x) lo = mid; else hi = mid - 1; }
    const DcMat mt = mats[lo];
    const int local = blockIdx.x - tile_prefix[lo];
    const int nt = (mt.n + 31) >> 5;
    const int by = local / nt, bx = local - by * nt;
    const int c0 = bx * 32, r0 = by * 32;
    const float* Zt = mt.Z[mt.fin];
    const int lx = threadIdx.x & 31, ly = threadIdx.x >> 5;
    for (int cc = ly; cc < 32; cc += 8) {
        const int c = c0 + cc, r = r0 + lx;
        t[cc][lx] = (c < mt.n && r < mt.n) ? Zt[(size_t)c * mt.ldz + r] : 0.f;
    }
    __syncthreads();
    for (int rr = ly; rr < 32; rr += 8) {
        const int r = r0 + rr, c = c0 + lx;
        if (r < mt.n && c < mt.n) mt.Q[(size_t)r * mt.n + c] = t[lx][rr];
    }
    if (by == 0 && ly == 0) {
        const int c = c0 + lx;
        if (c < mt.n) { const float v = mt.dcur[mt.fin][c]; mt.dout[c] = v > eps_floor ? v : 0.f; }
        if (bx == 0 && lx == 0 && sweeps_out) sweeps_out[out_index[lo]] = 1;
    }
}

// ---------------------------------------------------------------------------
// host: layout of the workspace and of the device tables
// ---------------------------------------------------------------------------
struct MatLayout {
    int n, ldz, npanels, nks, nleaf, depth;
    size_t T, Z0, Z1, U, Ypart, Y2, Tb, vecs, Vp, Wp, part, zpart, ints, rotcs;
};
struct Plan {
    int nb = 0, sms = 148, grid = 0, nlevels = 0, nmax = 0, max_panels = 0;
    std::vector<MatLayout> ml;
    std::vector<int> team;
    // host images of the device tables (uploaded in one copy)
    std::vector<int> cta_begin, entries;
    std::vector<DcLeaf> leaves;
    std::vector<DcMerge> merges;                 // all levels, level l at [merge_off[l], merge_off[l+1])
    std::vector<int> merge_off, level_nmax;
    std::vector<Tile> gemm_tiles; std::vector<int> gemm_off;
    std::vector<Tile> panel_tiles;               // (matrix, panel) for gram / tinv
    std::vector<Tile> y_tiles, yt_tiles, upd_tiles; std::vector<int> y_off, yt_off, upd_off;   // per iteration
    // offsets inside the workspace
    size_t off_tables = 0, off_zero = 0, zero_bytes = 0, off_mats = 0, total = 0;
    size_t t_jobs = 0, t_mats = 0, t_src = 0, t_oidx = 0, t_wpre = 0, t_cta = 0, t_ent = 0, t_leaves = 0, t_merges = 0, t_gemm = 0,
           t_panels = 0, t_y = 0, t_yt = 0, t_upd = 0, t_tmaps = 0, tables_bytes = 0;
    size_t z_bar = 0, z_states = 0, z_prof = 0;
};

static int round8(int n) { return (n + 7) & ~7; }     // row pitch: 8-row blocks of the tridiagonalisation read 32-byte slices

static void leaf_bounds(int n, std::vector<int>& b) {
    const int L = (n + LEAF - 1) / LEAF;
    b.resize(L + 1);
    for (int i = 0; i <= L; ++i) b[i] = (int)(((long)n * i) / L) & ~3;
    b[L] = n;
}

static void build_plan(const int32_t* n, const int* idx, int nb, Plan& pl) {
    pl.nb = nb;
    pl.sms = sm_count();
    pl.ml.resize(nb);
    std::vector<int> ns(nb);
    size_t off = 0;
    auto take = [&](size_t bytes) { const size_t o = off; off += align_up(bytes, 256); return o; };
    // ---- tables first (sizes depend only on the n's) ----
    pl.team.assign(nb, 1);
    pl.cta_begin.assign(pl.sms + 1, 0);
    pl.entries.assign((size_t)2 * nb * pl.sms + 2, 0);
    for (int a = 0; a < nb; ++a) ns[a] = n[idx[a]];
    pl.grid = trd::plan_teams(ns.data(), nb, pl.sms, pl.team.data(), pl.cta_begin.data(), pl.entries.data());
    int maxdepth = 0;
    std::vector<std::vector<int>> bounds(nb);
    for (int a = 0; a < nb; ++a) {
        MatLayout& m = pl.ml[a];
        m.n = ns[a]; m.ldz = round8(m.n);
        m.npanels = (m.n - 1 + PB - 1) / PB;
        m.nks = (m.ldz + YCH - 1) / YCH;
        leaf_bounds(m.n, bounds[a]);
        m.nleaf = (int)bounds[a].size() - 1;
        m.depth = 0;
        while ((1 << m.depth) < m.nleaf) ++m.depth;
        if (m.depth > maxdepth) maxdepth = m.depth;
        if (m.n > pl.nmax) pl.nmax = m.n;
        if (m.npanels > pl.max_panels) pl.max_panels = m.npanels;
        for (int i = 0; i < m.nleaf; ++i) pl.leaves.push_back(DcLeaf{a, bounds[a][i], bounds[a][i + 1] - bounds[a][i], 0});
    }
    pl.nlevels = maxdepth;
    pl.merge_off.assign(maxdepth + 2, 0);
    pl.level_nmax.assign(maxdepth + 2, 0);
    pl.gemm_off.assign(maxdepth + 2, 0);
    for (int l = 1; l <= maxdepth; ++l) {
        pl.merge_off[l] = (int)pl.merges.size();
        pl.gemm_off[l] = (int)pl.gemm_tiles.size();
        for (int a = 0; a < nb; ++a) {
            const MatLayout& m = pl.ml[a];
            if (l > m.depth) continue;
            const int span = 1 << l, half = span >> 1;
            for (int k = 0; k * span < m.nleaf; ++k) {
                const int l0 = k * span, l1 = std::min(m.nleaf, l0 + half), l2 = std::min(m.nleaf, l0 + span);
                const int o = bounds[a][l0], n1 = bounds[a][l1] - o, n2 = bounds[a][l2] - bounds[a][l1];
                const int mi = (int)pl.merges.size() - pl.merge_off[l];
                pl.merges.push_back(DcMerge{a, o, n1, n2});
                const int nm = n1 + n2;
                if (nm > pl.level_nmax[l]) pl.level_nmax[l] = nm;
                if (n2 > 0) {
                    const int t = (nm + TILE - 1) / TILE;
                    for (int tm = 0; tm < t; ++tm)
                        for (int tn = 0; tn < t; ++tn) pl.gemm_tiles.push_back(Tile{mi, tm, tn, 0});
                }
            }
        }
    }
    pl.merge_off[maxdepth + 1] = (int)pl.merges.size();
    pl.gemm_off[maxdepth + 1] = (int)pl.gemm_tiles.size();
    for (int a = 0; a < nb; ++a)
        for (int p = 0; p < pl.ml[a].npanels; ++p) pl.panel_tiles.push_back(Tile{a, p, 0, 0});
    pl.y_off.assign(pl.max_panels + 1, 0); pl.yt_off.assign(pl.max_panels + 1, 0); pl.upd_off.assign(pl.max_panels + 1, 0);
    for (int it = 0; it < pl.max_panels; ++it) {
        pl.y_off[it] = (int)pl.y_tiles.size(); pl.yt_off[it] = (int)pl.yt_tiles.size(); pl.upd_off[it] = (int)pl.upd_tiles.size();
        for (int a = 0; a < nb; ++a) {
            const MatLayout& m = pl.ml[a];
            if (it >= m.npanels) continue;
            const int p = m.npanels - 1 - it, jb = p * PB;
            const int tmn = (m.n + TILE - 1) / TILE;
            const int nks = (m.ldz - jb + YCH - 1) / YCH;
            const int tnn = (m.n - jb + TILE - 1) / TILE;
            for (int tm = 0; tm < tmn; ++tm) {
                for (int ks = 0; ks < nks; ++ks) pl.y_tiles.push_back(Tile{a, tm, ks, p});
                pl.yt_tiles.push_back(Tile{a, tm, nks, p});
                for (int tn = 0; tn < tnn; ++tn) pl.upd_tiles.push_back(Tile{a, tm, tn, p});
            }
        }
    }
    pl.y_off[pl.max_panels] = (int)pl.y_tiles.size(); pl.yt_off[pl.max_panels] = (int)pl.yt_tiles.size();
    pl.upd_off[pl.max_panels] = (int)pl.upd_tiles.size();
    // ---- table block ----
    size_t t = 0;
    auto ttake = [&](size_t bytes) { const size_t o = t; t += align_up(bytes, 256); return o; };
    pl.t_jobs = ttake(sizeof(trd::Job) * (nb + 1));
    pl.t_mats = ttake(sizeof(DcMat) * (nb + 1));
    pl.t_src = ttake(sizeof(void*) * (nb + 1));
    pl.t_oidx = ttake(sizeof(int) * (nb + 1));
    pl.t_wpre = ttake(sizeof(int) * (nb + 2));
    pl.t_cta = ttake(sizeof(int) * pl.cta_begin.size());
    pl.t_ent = ttake(sizeof(int) * pl.entries.size());
    pl.t_leaves = ttake(sizeof(DcLeaf) * (pl.leaves.size() + 1));
    pl.t_merges = ttake(sizeof(DcMerge) * (pl.merges.size() + 1));
    pl.t_gemm = ttake(sizeof(Tile) * (pl.gemm_tiles.size() + 1));
    pl.t_panels = ttake(sizeof(Tile) * (pl.panel_tiles.size() + 1));
    pl.t_y = ttake(sizeof(Tile) * (pl.y_tiles.size() + 1));
    pl.t_yt = ttake(sizeof(Tile) * (pl.yt_tiles.size() + 1));
    pl.t_upd = ttake(sizeof(Tile) * (pl.upd_tiles.size() + 1));
    pl.t_tmaps = ttake((size_t)128 * (nb + 1));      // CUtensorMap of every work array (TMA-fed symmetric product)
    pl.tables_bytes = t;
    pl.off_tables = take(t);
    // ---- zero-initialised block: barrier counters (one 128-byte line each), merge states ----
    size_t z = 0;
    auto ztake = [&](size_t bytes) { const size_t o = z; z += align_up(bytes, 256); return o; };
    pl.z_bar = ztake((size_t)128 * (nb + 1));
    pl.z_states = ztake(sizeof(DcState) * (pl.merges.size() + 1));
    pl.z_prof = ztake(sizeof(long long) * 16 * (nb + 1));
    pl.zero_bytes = z;
    pl.off_zero = take(z);
    // ---- per-matrix buffers ----
    for (int a = 0; a < nb; ++a) {
        MatLayout& m = pl.ml[a];
        const size_t nn = (size_t)m.n * m.ldz * sizeof(float);
        m.T = take(nn); m.Z0 = take(nn); m.Z1 = take(nn); m.U = take(nn);
        m.Ypart = take((size_t)m.nks * m.n * PB * sizeof(float));
        m.Y2 = take((size_t)m.n * PB * sizeof(float));
        m.Tb = take((size_t)m.npanels * PB * PB * sizeof(float));
        m.vecs = take((size_t)13 * m.ldz * sizeof(float));
        m.Vp = take((size_t)m.n * trd::NB * sizeof(float));
        m.Wp = take((size_t)m.n * trd::NB * sizeof(float));
        m.part = take((size_t)trd::NPART * trd::TS * sizeof(float));
        m.zpart = take((size_t)pl.team[a] * m.ldz * sizeof(float));
        m.ints = take((size_t)4 * m.ldz * sizeof(int));
        m.rotcs = take((size_t)2 * m.ldz * sizeof(float));
    }
    pl.total = off + 256;
}

}  // namespace dc

using namespace dc;

// CUDA events around the stages of the last call (per device): recorded on the caller's stream (asynchronous, no
// host synchronisation), read by kfacb_eigh_last_profile after the caller has synchronised
struct StageEvents { cudaEvent_t ev[4] = {nullptr, nullptr, nullptr, nullptr}; bool ok = false; double bytes_trd = 0, n3 = 0; int nb = 0, grid = 0; };
static StageEvents g_stage[64];
static StageEvents* stage_events() {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return nullptr;
    StageEvents& s = g_stage[dev];
    if (!s.ok) {
        for (int i = 0; i < 4; ++i) if (cudaEventCreate(&s.ev[i]) != cudaSuccess) return nullptr;
        s.ok = true;
    }
    return &s;
}
int eigh_tdc_last_profile(double* out8) {
    StageEvents* s = stage_events();
    for (int i = 0; i < 8; ++i) out8[i] = 0.0;
    if (!s || s->nb == 0) return KFACB_OK;
    if (cudaEventSynchronize(s->ev[3]) != cudaSuccess) { (void)cudaGetLastError(); return KFACB_OK; }
    float t = 0.f;
    if (cudaEventElapsedTime(&t, s->ev[0], s->ev[1]) == cudaSuccess) out8[0] = t;      // tridiagonalisation (copy + persistent kernel)
    if (cudaEventElapsedTime(&t, s->ev[1], s->ev[2]) == cudaSuccess) out8[1] = t;      // divide and conquer
    if (cudaEventElapsedTime(&t, s->ev[2], s->ev[3]) == cudaSuccess) out8[2] = t;      // back-transformation + write
    out8[3] = s->bytes_trd;     // algorithmic bytes of the tridiagonalisation: sum (2/3 + 1/24) n^3 -- the lower triangle of the trailing
                                // matrix read once per column (symmetric product) + one read-modify-write of it per 32-column panel
    out8[4] = s->n3;            // sum n^3 of the blocked matrices
    out8[5] = s->nb; out8[6] = s->grid;
    (void)cudaGetLastError();
    return KFACB_OK;
}

}  // namespace kfacb

// Introspection of the scheduler (tests): team size per matrix (0 for the n <= 128 matrices that
// take the direct path) for a device with `sms` SMs; returns the number of CTAs of the persistent kernel.
extern "C" int kfacb_eigh_plan_teams(const int32_t* n, int32_t count, int32_t sms, int32_t* team_out) {
    if (!n || !team_out || count < 0 || sms <= 0) return -1;
    std::vector<int> ns, idx, team, cta(sms + 1, 0);
    for (int i = 0; i < count; ++i) { team_out[i] = 0; if (n[i] > 64) { ns.push_back(n[i]); idx.push_back(i); } }
    if (ns.empty()) return 0;
    team.assign(ns.size(), 1);
    std::vector<int> entries((size_t)2 * ns.size() * sms + 2, 0);
    const int grid = kfacb::trd::plan_teams(ns.data(), (int)ns.size(), sms, team.data(), cta.data(), entries.data());
    for (size_t a = 0; a < ns.size(); ++a) team_out[idx[a]] = team[a];
    return grid;
}

namespace kfacb {

size_t eigh_tdc_workspace_bytes(const int32_t* n, int count, const int* idx, int nb) {
    (void)count;
    if (nb <= 0) return 0;
    Plan pl;
    build_plan(n, idx, nb, pl);
    return pl.total;
}

int eigh_tdc_run(const int32_t* n, const int* idx, int nb, const float* const* M, float* const* Q, float* const* d,
                 float eps_floor, int32_t* sweeps_out, char* ws, size_t ws_bytes, cudaStream_t st) {
    if (nb <= 0) return KFACB_OK;
    Plan pl;
    build_plan(n, idx, nb, pl);
    if (ws_bytes < pl.total) { set_error("eigh workspace too small (%zu < %zu)", ws_bytes, pl.total); return KFACB_ERR_WORKSPACE; }
    for (int a = 0; a < nb; ++a)
        KFACB_REQUIRE(pl.ml[a].n <= trd::MAXN, "matrix %d: n = %d > %d is not supported", idx[a], pl.ml[a].n, trd::MAXN);
    static int trd_prof = -1;                   // KFACB_TRD_PROF=1: per-phase clock totals of the tridiagonalisation (debug, synchronises)
    if (trd_prof < 0) { const char* e = getenv("KFACB_TRD_PROF"); trd_prof = (e && *e && atoi(e) != 0) ? 1 : 0; }
    // ---- host image of the tables ----
    std::vector<char> img(pl.tables_bytes, 0);
    trd::Job* jobs = reinterpret_cast<trd::Job*>(img.data() + pl.t_jobs);
    DcMat* mats = reinterpret_cast<DcMat*>(img.data() + pl.t_mats);
    const float** srcp = reinterpret_cast<const float**>(img.data() + pl.t_src);
    int* oidx = reinterpret_cast<int*>(img.data() + pl.t_oidx);
    int* wpre = reinterpret_cast<int*>(img.data() + pl.t_wpre);
    int write_tiles = 0;
    char* zero = ws + pl.off_zero;
    for (int a = 0; a < nb; ++a) {
        const MatLayout& m = pl.ml[a];
        float* vec = reinterpret_cast<float*>(ws + m.vecs);
        int* ints = reinterpret_cast<int*>(ws + m.ints);
        auto V = [&](int k) { return vec + (size_t)k * m.ldz; };
        trd::Job& jb = jobs[a];
        jb.T = reinterpret_cast<float*>(ws + m.T);
        jb.d = V(0); jb.e = V(1); jb.tau = V(2); jb.x = V(3); jb.y = V(4);
        jb.Vp = reinterpret_cast<float*>(ws + m.Vp); jb.Wp = reinterpret_cast<float*>(ws + m.Wp);
        jb.part = reinterpret_cast<float*>(ws + m.part);
        jb.zpart = reinterpret_cast<float*>(ws + m.zpart);
        jb.bar = reinterpret_cast<unsigned*>(zero + pl.z_bar + (size_t)128 * a);
        jb.n = m.n; jb.ldt = m.ldz; jb.team = pl.team[a];
        jb.prof = trd_prof ? reinterpret_cast<long long*>(zero + pl.z_prof) + 16 * a : nullptr;
        jb.tmap = ws + pl.off_tables + pl.t_tmaps + (size_t)128 * a;
        {
            const int er = trd::encode_tmap(img.data() + pl.t_tmaps + (size_t)128 * a, jb.T, m.n, m.ldz);
            KFACB_REQUIRE(er == 0, "cuTensorMapEncodeTiled failed (%d) for matrix %d (n = %d)", er, idx[a], m.n);
        }
        DcMat& mt = mats[a];
        mt.T = jb.T; mt.Z[0] = reinterpret_cast<float*>(ws + m.Z0); mt.Z[1] = reinterpret_cast<float*>(ws + m.Z1);
        mt.U = reinterpret_cast<float*>(ws + m.U); mt.Ypart = reinterpret_cast<float*>(ws + m.Ypart);
        mt.Y2 = reinterpret_cast<float*>(ws + m.Y2); mt.Tb = reinterpret_cast<float*>(ws + m.Tb);
        mt.d = jb.d; mt.e = jb.e; mt.tau = jb.tau;
        mt.dcur[0] = V(5); mt.dcur[1] = V(6); mt.dl = V(7); mt.w = V(8); mt.zh = V(9); mt.lam = V(10); mt.vals = V(11);
        mt.rotcs = reinterpret_cast<float*>(ws + m.rotcs);
        mt.src = ints; mt.rnk = ints + m.ldz; mt.rotab = ints + 2 * m.ldz;
        mt.Q = Q[idx[a]]; mt.dout = d[idx[a]];
        mt.n = m.n; mt.ldz = m.ldz; mt.npanels = m.npanels; mt.fin = m.depth & 1;
        srcp[a] = M[idx[a]];
        oidx[a] = idx[a];
        wpre[a] = write_tiles;
        write_tiles += ((m.n + 31) / 32) * ((m.n + 31) / 32);
    }
    memcpy(img.data() + pl.t_cta, pl.cta_begin.data(), sizeof(int) * pl.cta_begin.size());
    memcpy(img.data() + pl.t_ent, pl.entries.data(), sizeof(int) * pl.entries.size());
    auto put = [&](size_t at, const void* p, size_t bytes) { if (bytes) memcpy(img.data() + at, p, bytes); };
    put(pl.t_leaves, pl.leaves.data(), sizeof(DcLeaf) * pl.leaves.size());
    put(pl.t_merges, pl.merges.data(), sizeof(DcMerge) * pl.merges.size());
    put(pl.t_gemm, pl.gemm_tiles.data(), sizeof(Tile) * pl.gemm_tiles.size());
    put(pl.t_panels, pl.panel_tiles.data(), sizeof(Tile) * pl.panel_tiles.size());
    put(pl.t_y, pl.y_tiles.data(), sizeof(Tile) * pl.y_tiles.size());
    put(pl.t_yt, pl.yt_tiles.data(), sizeof(Tile) * pl.yt_tiles.size());
    put(pl.t_upd, pl.upd_tiles.data(), sizeof(Tile) * pl.upd_tiles.size());
    char* tb = ws + pl.off_tables;
    KFACB_CHECK_CUDA(cudaMemcpyAsync(tb, img.data(), pl.tables_bytes, cudaMemcpyHostToDevice, st));
    KFACB_CHECK_CUDA(cudaMemsetAsync(zero, 0, pl.zero_bytes, st));
    const trd::Job* d_jobs = reinterpret_cast<const trd::Job*>(tb + pl.t_jobs);
    const DcMat* d_mats = reinterpret_cast<const DcMat*>(tb + pl.t_mats);
    const DcLeaf* d_leaves = reinterpret_cast<const DcLeaf*>(tb + pl.t_leaves);
    const DcMerge* d_merges = reinterpret_cast<const DcMerge*>(tb + pl.t_merges);
    DcState* d_states = reinterpret_cast<DcState*>(zero + pl.z_states);
    const Tile* d_gemm = reinterpret_cast<const Tile*>(tb + pl.t_gemm);
    const Tile* d_panels = reinterpret_cast<const Tile*>(tb + pl.t_panels);
    const Tile* d_y = reinterpret_cast<const Tile*>(tb + pl.t_y);
    const Tile* d_yt = reinterpret_cast<const Tile*>(tb + pl.t_yt);
    const Tile* d_upd = reinterpret_cast<const Tile*>(tb + pl.t_upd);

    // GEMM core per matrix: 0 = by size (SIMT up to SIMT_MAX_N), -1 = SIMT everywhere (KFACB_GEMM_SIMT=1), +1 = tensor cores everywhere
    static int force = -2;
    if (force == -2) { const char* e = getenv("KFACB_TDC_TC"); force = !use_tc_gemm() ? -1 : ((e && *e) ? (atoi(e) > 0 ? 1 : -1) : 0); }
    static int dbg = -1;                        // KFACB_TDC_DEBUG=1: skip the back-transformation (Q = eigenvectors of T)
    if (dbg < 0) { const char* e = getenv("KFACB_TDC_DEBUG"); dbg = (e && *e) ? atoi(e) : 0; }
    StageEvents* sev = stage_events();
    if (sev) {
        sev->nb = nb; sev->grid = pl.grid; sev->bytes_trd = 0; sev->n3 = 0;
        for (int a = 0; a < nb; ++a) { const double n3 = (double)pl.ml[a].n * pl.ml[a].n * pl.ml[a].n; sev->n3 += n3; sev->bytes_trd += (2.0 / 3.0 + 1.0 / 24.0) * n3; }
        cudaEventRecord(sev->ev[0], st);
    }
    // ---- (1) tridiagonalisation ----
    KFACB_CHECK_CUDA(trd::launch_copy(d_jobs, reinterpret_cast<const float* const*>(tb + pl.t_src), nb, pl.nmax, st));
    size_t trd_smem = 0;
    for (int a = 0; a < nb; ++a) trd_smem = std::max(trd_smem, trd::smem_bytes(pl.ml[a].n, pl.team[a]));
    KFACB_CHECK_CUDA(trd::launch(d_jobs, reinterpret_cast<const int*>(tb + pl.t_cta),
                                 reinterpret_cast<const int*>(tb + pl.t_ent), pl.grid, trd_smem, st));
    if (sev) cudaEventRecord(sev->ev[1], st);
    if (trd_prof) {
        std::vector<long long> hp((size_t)16 * nb);
        if (cudaStreamSynchronize(st) == cudaSuccess &&
            cudaMemcpy(hp.data(), zero + pl.z_prof, hp.size() * sizeof(long long), cudaMemcpyDeviceToHost) == cudaSuccess)
        {
            long long t0 = 0;
            for (int a = 0; a < nb; ++a) { const long long s0 = hp[(size_t)16 * a + 9]; if (s0 && (!t0 || s0 < t0)) t0 = s0; }
            for (int a = 0; a < nb; ++a) {
                const long long* q = hp.data() + 16 * a;
                fprintf(stderr, "[kfacb trd profile] n=%d team=%d  start %.2f end %.2f ms  kcycles: A_pro %lld A_rows %lld bar1 %lld B_setup %lld B_symv %lld "
                        "B_red %lld bar2 %lld pan_fin %lld rank2k %lld | warp 0: tma_wait %lld units %lld\n", pl.ml[a].n, pl.team[a], (q[9] - t0) * 1e-6, (q[10] - t0) * 1e-6,
                        q[0] / 1000, q[1] / 1000, q[2] / 1000, q[3] / 1000, q[4] / 1000, q[5] / 1000, q[6] / 1000, q[7] / 1000, q[8] / 1000,
                        q[11] / 1000, q[12]);
            }
        }
    }
    // ---- (2) divide and conquer ----
    static bool attr_done[64] = {false};
    int dev = 0;
    KFACB_CHECK_CUDA(cudaGetDevice(&dev));
    if (dev < 0 || dev >= 64 || !attr_done[dev]) {
        KFACB_CHECK_CUDA(cudaFuncSetAttribute(dc_leaf_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 2 * LEAF_MAX * (LEAF_MAX + 1) * 4));
        KFACB_CHECK_CUDA(cudaFuncSetAttribute(dc_prep_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 3 * trd::MAXN * 4));
        KFACB_CHECK_CUDA(cudaFuncSetAttribute(dc_secular_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 2 * trd::MAXN * 4));
        KFACB_CHECK_CUDA(cudaFuncSetAttribute(dc_rank_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, trd::MAXN * 4));
        KFACB_CHECK_CUDA(cudaFuncSetAttribute(dc_gemm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, tcg::SMEM));
        KFACB_CHECK_CUDA(cudaFuncSetAttribute(bt_y_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, tcg::SMEM));
        KFACB_CHECK_CUDA(cudaFuncSetAttribute(bt_yt_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, tcg::SMEM));
        KFACB_CHECK_CUDA(cudaFuncSetAttribute(bt_upd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, tcg::SMEM));
        KFACB_CHECK_CUDA(cudaFuncSetAttribute(bt_tinv_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 2 * PB * TP * 4));
        if (dev >= 0 && dev < 64) attr_done[dev] = true;
    }
    {
        int lmax = 0;
        for (const DcLeaf& lf : pl.leaves) lmax = std::max(lmax, lf.nl);
        const int m = (lmax + 1) & ~1;
        dc_leaf_kernel<<<(unsigned)pl.leaves.size(), 256, 2 * (size_t)m * (m + 1) * sizeof(float), st>>>(d_mats, d_leaves);
        KFACB_CHECK_LAUNCH();
    }
    for (int l = 1; l <= pl.nlevels; ++l) {
        const int in = (l - 1) & 1, out = l & 1;
        const int nmg = pl.merge_off[l + 1] - pl.merge_off[l];
        if (nmg == 0) continue;
        const DcMerge* mg = d_merges + pl.merge_off[l];
        DcState* sts = d_states + pl.merge_off[l];
        const int nmx = pl.level_nmax[l];
        dc_zero_kernel<<<dim3(cdiv(nmx, 8), nmg), 256, 0, st>>>(d_mats, mg, in);
        dc_prep_kernel<<<nmg, 1024, 3 * (size_t)nmx * sizeof(float), st>>>(d_mats, mg, sts, in, nmx);
        dc_rotate_kernel<<<dim3(cdiv(nmx, 256), nmg), 256, 0, st>>>(d_mats, mg, sts, in);
        dc_secular_kernel<<<dim3(cdiv(nmx, 8), nmg), 256, 2 * (size_t)nmx * sizeof(float), st>>>(d_mats, mg, sts, nmx);
        dc_zhat_kernel<<<dim3(cdiv(nmx, 32), nmg), 1024, 0, st>>>(d_mats, mg, sts);
        dc_vec_kernel<<<dim3(cdiv(nmx, 8), nmg), 256, 0, st>>>(d_mats, mg, sts);
        dc_rank_kernel<<<dim3(cdiv(nmx, 256), nmg), 256, (size_t)nmx * sizeof(float), st>>>(d_mats, mg, sts, out, nmx);
        const int ngt = pl.gemm_off[l + 1] - pl.gemm_off[l];
        if (ngt > 0)
            dc_gemm_kernel<<<ngt, tcg::THREADS, tcg::SMEM, st>>>(d_mats, mg, sts, d_gemm + pl.gemm_off[l], in, force);
        dc_copy_kernel<<<dim3(cdiv(nmx, 8), nmg), 256, 0, st>>>(d_mats, mg, sts, in);
        count_launch(8);
        KFACB_CHECK_LAUNCH();
    }
    if (sev) cudaEventRecord(sev->ev[2], st);
    // ---- (3) back-transformation ----
    if (dbg != 1) {
        const int np = (int)pl.panel_tiles.size();
        bt_gram_kernel<<<np, GEMM_THREADS, 0, st>>>(d_mats, d_panels);
        bt_tinv_kernel<<<np, PB, 2 * PB * TP * sizeof(float), st>>>(d_mats, d_panels);
        count_launch(1);
        KFACB_CHECK_LAUNCH();
        for (int it = 0; it < pl.max_panels; ++it) {
            const int ny = pl.y_off[it + 1] - pl.y_off[it], nyt = pl.yt_off[it + 1] - pl.yt_off[it], nu = pl.upd_off[it + 1] - pl.upd_off[it];
            if (ny == 0) continue;
            bt_y_kernel<<<ny, tcg::THREADS, tcg::SMEM, st>>>(d_mats, d_y + pl.y_off[it], force);
            bt_yt_kernel<<<nyt, tcg::THREADS, tcg::SMEM, st>>>(d_mats, d_yt + pl.yt_off[it], force);
            bt_upd_kernel<<<nu, tcg::THREADS, tcg::SMEM, st>>>(d_mats, d_upd + pl.upd_off[it], force);
            count_launch(2);
            KFACB_CHECK_LAUNCH();
        }
    }
    // ---- (4) floor + write ----
    eig_write_kernel<<<write_tiles, 256, 0, st>>>(d_mats, reinterpret_cast<const int*>(tb + pl.t_wpre), nb, eps_floor, sweeps_out,
                                                  reinterpret_cast<const int*>(tb + pl.t_oidx));
    KFACB_CHECK_LAUNCH();
    if (sev) cudaEventRecord(sev->ev[3], st);
    return KFACB_OK;
}

}  // namespace kfacb
