// eigh.cu -- batched symmetric eigendecomposition (C-ABI: kfacb_eigh_batched).
// Replaces torch.linalg.eigh + the eigenvalue floor of the reference
// (kfac/utils.py:22-25, kfac/kfac_preconditioner_eigen_dp.py:62-75).
//
// Two paths, both Jacobi (high relative accuracy, embarrassingly parallel,
// no library dependency):
//   n <= 128 : one CTA per matrix, two-sided cyclic Jacobi with the round-robin
//              parallel ordering, matrix and eigenvector accumulator resident
//              in shared memory.
//   n  > 128 : implicit two-sided BLOCK Jacobi.  The work array
//              X = [W^T | V^T] (n_pad x 2 n_pad, row i = column i of W = A V and
//              of V) starts as [A | I].  A round pairs up 64-row blocks with the
//              round-robin schedule; per pair: the 128x128 projected block
//              H = V_R^T A V_R = V_R W_R^T (tcgen05 3xTF32, split-K), one cyclic
//              Jacobi sweep of H in shared memory -> U, rows <- U^T rows (W and V
//              together, tcgen05 3xTF32).  All pairs of all matrices of the batch
//              run in the same launches.  After convergence V is re-orthogonalised
//              (one Newton-Schulz step), eigenvalues are recomputed as Rayleigh
//              quotients v^T A v, sorted ascending, floored, and V is written out
//              column-wise.  Working on H (not on the Gram W_R W_R^T of one-sided
//              Hestenes Jacobi) avoids squaring the condition number.
#include <stdlib.h>
#include "common.cuh"
#include "tc_gemm.cuh"
#include "tc.cuh"
#include "jacobi_smem.cuh"
#include "eigh_trd.cuh"

namespace kfacb {

constexpr int BJ_B = 64;              // rows per block
constexpr int BJ_M = 2 * BJ_B;        // rows per pair = size of the inner eigenproblem (= TILE)
constexpr int BJ_KCHUNK = 512;        // largest split-K chunk of the projected-block GEMM (the host picks 256/384/512)
static_assert(BJ_M == TILE, "a pair is one 128-row GEMM tile");
constexpr int BJ_MAX_SWEEPS = 16;     // cap of outer sweeps (KFACB_MAX_SWEEPS)

struct BjMat {
    float* X;             // n_pad x 2 n_pad work array
    const float* A;       // input matrix n x n
    float* Q;             // output eigenvectors (columns) n x n
    float* d;             // output eigenvalues n
    float* lam;           // workspace n_pad (rayleigh quotients)
    float* nrm;           // workspace n_pad (||v_i||^2)
    int* perm;            // workspace n
    int n, n_pad, p;      // p = even number of blocks (incl. a dummy when odd)
    int nblk;             // real number of blocks
    int pair_begin;       // first global pair index of this matrix
    int ksplit;           // projected-block split-K factor = ceil(n_pad / kchunk)
    int kchunk;           // k-chunk of one projected-block CTA
    float tol_conv;       // sweep-time off-norm / ||A||_F below which the matrix is converged
};

struct BjState {          // per matrix, device
    int scale_bits;       // float bits of max |diag(A)| (absolute scale for the thresholds)
    int converged;
    int rot;              // unused (kept for layout stability)
    int sweeps;
    float off2;           // sum over the pairs visited this sweep of ||cross block of H||_F^2
    float norm2;          // ||A||_F^2
};

// ---------------------------------------------------------------------------
// direct path: n <= 128, one CTA per matrix
// ---------------------------------------------------------------------------
struct DirectMat { const float* A; float* Q; float* d; int n; };

__global__ void __launch_bounds__(512, 1)
jacobi_direct_kernel(const DirectMat* mats, float eps_floor, int* sweeps_out, const int* out_index) {
    extern __shared__ float smem[];
    const DirectMat mt = mats[blockIdx.x];
    const int n = mt.n, m = (n + 1) & ~1, P = m + 1;
    float* S = smem;
    float* U = smem + (size_t)m * P;
    __shared__ float s_lam[DIRECT_MAX_N];
    __shared__ int s_rank[DIRECT_MAX_N];
    __shared__ float s_red[16];
    const int tid = threadIdx.x, nt = blockDim.x;
    float dmax = 0.f;
    for (int e = tid; e < m * m; e += nt) {
        const int i = e / m, j = e - i * m;
        float v = 0.f;
        if (i < n && j < n) v = 0.5f * (mt.A[(long)i * n + j] + mt.A[(long)j * n + i]);
        S[i * P + j] = v;
        U[i * P + j] = (i == j) ? 1.f : 0.f;
        if (i == j) dmax = fmaxf(dmax, fabsf(v));
    }
    // max |diag| as the absolute scale
    for (int o = 16; o > 0; o >>= 1) dmax = fmaxf(dmax, __shfl_xor_sync(0xffffffffu, dmax, o));
    if ((tid & 31) == 0) s_red[tid >> 5] = dmax;
    __syncthreads();
    if (tid == 0) {
        float v = 0.f;
        for (int i = 0; i < (nt >> 5); ++i) v = fmaxf(v, s_red[i]);
        s_red[0] = v;
    }
    __syncthreads();
    const float scale = s_red[0];
    __syncthreads();
    const int sw = jacobi_smem<0>(S, U, m, P, 1.2e-7f, 1.5e-8f * scale, 1e30f, 30, nullptr);
    // Eigenvalues as Rayleigh quotients of the ORIGINAL matrix: Y = A U (into S), lam_j = u_j.y_j / u_j.u_j.
    // This is second-order accurate in the eigenvector error and immune to drift of diag(S).
    __syncthreads();
    for (int e = tid; e < n * n; e += nt) {
        const int i = e / n, j = e - i * n;
        float acc = 0.f;
        for (int k = 0; k < n; ++k) acc = fmaf(__ldg(mt.A + (long)i * n + k), U[k * P + j], acc);
        S[i * P + j] = acc;
    }
    __syncthreads();
    __shared__ float s_nrm[DIRECT_MAX_N];
    if (tid < n) {
        double num = 0.0, den = 0.0;
        for (int i = 0; i < n; ++i) {
            const double u = U[i * P + tid];
            num += u * (double)S[i * P + tid];
            den += u * u;
        }
        s_lam[tid] = (float)(num / den);
        s_nrm[tid] = (float)(1.0 / sqrt(den));
    }
    __syncthreads();
    if (tid < n) {
        const float v = s_lam[tid];
        int rk = 0;
        for (int j = 0; j < n; ++j) {
            const float w = s_lam[j];
            rk += (w < v) || (w == v && j < tid);
        }
        s_rank[tid] = rk;
        mt.d[rk] = v > eps_floor ? v : 0.f;
    }
    __syncthreads();
    for (int e = tid; e < n * n; e += nt) {
        const int i = e / n, j = e - i * n;      // eigenvector j, component i
        mt.Q[(long)i * n + s_rank[j]] = U[i * P + j] * s_nrm[j];
    }
    if (tid == 0 && sweeps_out) sweeps_out[out_index[blockIdx.x]] = sw;
}

// ---------------------------------------------------------------------------
// blocked one-sided path
// ---------------------------------------------------------------------------
__device__ __forceinline__ bool bj_pair_blocks(const BjMat& mt, int q, int round, int& bi, int& bj) {
    const int pm1 = mt.p - 1;
    const int r = round % pm1;
    int a, b;
    if (q == 0) { a = pm1; b = r; }
    else { a = (r + q) % pm1; b = (r - q + pm1) % pm1; }
    bi = min(a, b); bj = max(a, b);
    return bj < mt.nblk;           // pairs with the dummy block idle this round
}

__global__ void bj_init_kernel(const BjMat* mats, BjState* state) {
    const BjMat mt = mats[blockIdx.y];
    float dmax = 0.f;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < mt.n; i += gridDim.x * blockDim.x)
        dmax = fmaxf(dmax, fabsf(mt.A[(long)i * mt.n + i]));
    if (dmax > 0.f) atomicMax(&state[blockIdx.y].scale_bits, __float_as_int(dmax));
    float ss = 0.f;
    const long ld = 2L * mt.n_pad;
    const long tot = (long)mt.n_pad * ld;
    for (long e = (long)blockIdx.x * blockDim.x + threadIdx.x; e < tot;
         e += (long)gridDim.x * blockDim.x) {
        const int i = (int)(e / ld), c = (int)(e - (long)i * ld);
        float v;
        if (c < mt.n_pad) { v = (i < mt.n && c < mt.n) ? 0.5f * (mt.A[(long)i * mt.n + c] + mt.A[(long)c * mt.n + i]) : 0.f; ss = fmaf(v, v, ss); }
        else v = (c - mt.n_pad == i) ? 1.f : 0.f;
        mt.X[e] = v;
    }
    ss = warp_sum(ss);
    if ((threadIdx.x & 31) == 0 && ss > 0.f) atomicAdd(&state[blockIdx.y].norm2, ss);
}

// ---------------------------------------------------------------------------
// The three kernels of a Jacobi round.  Blocks are 64 rows, a pair is 128 rows: the two GEMMs are
// 128x128-tile tcgen05 3xTF32 products on the shared core of tc_gemm.cuh, the 128x128 projected
// problem is solved in shared memory.
// ---------------------------------------------------------------------------
// rows of a pair viewed as a 128 x K operand (k contiguous): rows 0-63 = block bi, 64-127 = block bj
struct LdPairRows {
    static constexpr bool K_FAST = true;
    const float* base; long ld; int row_a, row_b; int kend;
    __device__ __forceinline__ float operator()(int mn, int k) const {
        if (k >= kend) return 0.f;
        const int grow = mn < BJ_B ? row_a + mn : row_b + (mn - BJ_B);
        return __ldg(base + (long)grow * ld + k);
    }
    __device__ __forceinline__ float4 quad(int mn, int k) const {       // rows are 256-byte aligned, k % 4 == 0
        if (k >= kend) return make_float4(0.f, 0.f, 0.f, 0.f);
        const int grow = mn < BJ_B ? row_a + mn : row_b + (mn - BJ_B);
        return __ldg(reinterpret_cast<const float4*>(base + (long)grow * ld + k));
    }
};
// columns of the pair's rows viewed as a (128 columns) x (128 rows-as-k) operand (column index contiguous)
struct LdPairCols {
    static constexpr bool K_FAST = false;
    const float* X; long ld; int row_a, row_b; long c0;
    __device__ __forceinline__ float operator()(int mn, int k) const {
        const long c = c0 + mn;
        if (c >= ld) return 0.f;
        const int grow = k < BJ_B ? row_a + k : row_b + (k - BJ_B);
        return X[(long)grow * ld + c];          // plain load: X is rewritten in place by this kernel
    }
    __device__ __forceinline__ float4 quad(int mn, int k) const {       // a quad never straddles the two blocks
        const long c = c0 + mn;
        if (c >= ld) return make_float4(0.f, 0.f, 0.f, 0.f);
        const int grow = k < BJ_B ? row_a + k : row_b + (k - BJ_B);
        const float* p = X + (long)grow * ld + c;
        return make_float4(p[0], p[ld], p[2 * ld], p[3 * ld]);
    }
};
// U (128x128 row-major, U[k][i]) as operand B(n = i, k)
struct LdUmat {
    static constexpr bool K_FAST = false;
    const float* U;
    __device__ __forceinline__ float operator()(int mn, int k) const { return __ldg(U + k * BJ_M + mn); }
    __device__ __forceinline__ float4 quad(int mn, int k) const {
        const float* p = U + k * BJ_M + mn;
        return make_float4(__ldg(p), __ldg(p + BJ_M), __ldg(p + 2 * BJ_M), __ldg(p + 3 * BJ_M));
    }
};

// (1) H = V_R^T A V_R = V_R W_R^T (128x128) over one k-chunk.  grid (max ksplit, total_pairs)
__global__ void __launch_bounds__(tcg::THREADS)
bj_h_tc_kernel(const BjMat* mats, const int* pair_mat, const BjState* state, int round,
               float* __restrict__ partial, const long* partial_off) {
    extern __shared__ __align__(128) unsigned char dsm_raw[];
    __shared__ __align__(8) uint64_t mbar;
    __shared__ uint32_t tmem_slot;
    // grid (max ksplit, total_pairs): the k-chunks of one pair are neighbours in launch order, so the CTAs
    // resident together stream adjacent 2 KB segments of the same 128 rows (DRAM page locality)
    const int pair_idx = blockIdx.y, zchunk = blockIdx.x;
    const int mi = pair_mat[pair_idx];
    const BjMat mt = mats[mi];
    if (state[mi].converged || zchunk >= mt.ksplit) return;
    int bi, bj;
    if (!bj_pair_blocks(mt, pair_idx - mt.pair_begin, round, bi, bj)) return;
    tcg::Ctx c;
    tcg::setup(c, dsm_raw, &mbar, &tmem_slot);
    const long ld = 2L * mt.n_pad;
    const int kbeg = zchunk * mt.kchunk, kend = min(mt.n_pad, kbeg + mt.kchunk);
    LdPairRows lv{mt.X + mt.n_pad, ld, bi * BJ_B, bj * BJ_B, kend};      // V rows
    LdPairRows lw{mt.X, ld, bi * BJ_B, bj * BJ_B, kend};                 // W rows
    tcg::mainloop<false>(lv, lw, 0, 0, kbeg, kend, false, c);     // 64 K-steps per chunk; H only steers the rotations
    const float* d = tcg::drain<false>(c, true);
    float* dst = partial + partial_off[mi] +
                 ((long)(pair_idx - mt.pair_begin) * mt.ksplit + zchunk) * (BJ_M * BJ_M);
    for (int e = threadIdx.x; e < BJ_M * BJ_M; e += tcg::THREADS)
        dst[e] = d[(e >> 7) * tcg::DPITCH + (e & 127)];
    tcg::teardown(c);
}

// (3) rows <- U^T rows for the 128 rows of a pair (W and V halves together):
//     D[c][i] = sum_k X[r_k][c] U[k][i]   (M = 128 columns of X per tile, N = 128, K = 128).
// Warp-specialised, persistent over `tiles_per_cta` consecutive 128-column tiles of one pair:
//   * U (128x128) is split into tf32 hi/lo and staged ONCE (128 KB, K-major core-matrix layout);
//   * warps 0-7 (producers): thread (c, h) owns column c / row half h of a 32-row k-slab; its 16 loads of
//     slabs q+1..q+4 are in flight (64 KB per SM) while slab q is split + stored into one of two 32 KB
//     operand buffers and its 12 UMMAs (3xTF32, 128x128x8) are issued; buffers are recycled through
//     tcgen05.commit -> mbarrier;
//   * warps 8-11 (epilogue): wait for the tile's accumulator (2 x 128 TMEM columns, double buffered),
//     tcgen05.ld it and store the rows back in place (lanes = consecutive columns: coalesced).
// (A 13th warp dedicated to the UMMA issue was measured slower: 416 threads cap the kernel at 128 registers
// and the producers spill.)
// grid (ceil(col_tiles / tiles_per_cta), total_pairs); tiles_per_cta is chosen per group by the host so that
// the CTA count fills whole waves of the 148 SMs (1 CTA / SM).
constexpr int UPD_PRODUCERS = 256;                       // 8 producer warps
constexpr int UPD_THREADS = UPD_PRODUCERS + 128;         // + 4 epilogue warps
constexpr int UPD_LBO = (TILE / 8) * 128;                 // 2048: stride between k-quads
constexpr int UPD_U_PART = (BJ_M / 4) * UPD_LBO;          // 65536 bytes: U hi (or lo)
constexpr int UPD_A_PART = (32 / 4) * UPD_LBO;            // 16384 bytes: one k-slab of X hi (or lo)
constexpr int UPD_SMEM = 2 * UPD_U_PART + 4 * UPD_A_PART; // 196608
__global__ void __launch_bounds__(UPD_THREADS)
bj_update_tc_kernel(const BjMat* mats, const int* pair_mat, const BjState* state, int round,
                    const float* __restrict__ ubuf, int tiles_per_cta) {
    extern __shared__ __align__(128) unsigned char dsm_raw[];
    __shared__ __align__(8) uint64_t bar_smem[2], bar_full[2], bar_empty[2];
    __shared__ uint32_t tmem_slot;
    const int pair_idx = blockIdx.y;
    const int mi = pair_mat[pair_idx];
    const BjMat mt = mats[mi];
    const long ld = 2L * mt.n_pad;
    const int tile0 = blockIdx.x * tiles_per_cta;
    if (state[mi].converged || (long)tile0 * TILE >= ld) return;
    int bi, bj;
    if (!bj_pair_blocks(mt, pair_idx - mt.pair_begin, round, bi, bj)) return;
    const int ntiles = min(tiles_per_cta, (int)((ld - (long)tile0 * TILE + TILE - 1) / TILE));
    const int tid = threadIdx.x, warp = tid >> 5;
    if (tid == 0) {
        tc::mbar_init(&bar_smem[0], 1); tc::mbar_init(&bar_smem[1], 1);
        tc::mbar_init(&bar_full[0], 1); tc::mbar_init(&bar_full[1], 1);
        tc::mbar_init(&bar_empty[0], 128); tc::mbar_init(&bar_empty[1], 128);
        tc::fence_mbar_init();
    }
    if (warp == 0) tc::tmem_alloc<256>(&tmem_slot);
    unsigned char* u_hi = dsm_raw;
    unsigned char* u_lo = dsm_raw + UPD_U_PART;
    unsigned char* a_buf = dsm_raw + 2 * UPD_U_PART;          // [buffer][hi|lo] x UPD_A_PART
    // stage U once: operand B(n = i, k) = U[k][i]; thread (i, kq) gathers 4 rows of column i (coalesced along i)
    {
        const float* usrc = ubuf + (long)pair_idx * (BJ_M * BJ_M);
        const int i = tid & 127;
        const int boff = (i & 7) * 16 + (i >> 3) * 128;
#pragma unroll 4
        for (int j = 0; j < (BJ_M / 4) / 2 && tid < 256; ++j) {
            const int kq = (tid >> 7) + 2 * j;
            const float* src = usrc + (4 * kq) * BJ_M + i;
            float4 u = make_float4(__ldg(src), __ldg(src + BJ_M), __ldg(src + 2 * BJ_M), __ldg(src + 3 * BJ_M));
            float4 hi, lo;
            tc::split_tf32(u, hi, lo);
            *reinterpret_cast<float4*>(u_hi + boff + kq * UPD_LBO) = hi;
            *reinterpret_cast<float4*>(u_lo + boff + kq * UPD_LBO) = lo;
        }
    }
    tc::fence_async_smem();
    tc::tc_fence_before();
    __syncthreads();
    tc::tc_fence_after();
    const uint32_t tmem = tmem_slot;
    const int row_a = bi * BJ_B, row_b = bj * BJ_B;
    // the 12 UMMAs of k-slab q (3xTF32: hi*hi + lo*hi + hi*lo), by ONE thread
    auto issue_slab = [&](int q) {
        const int b = q & 1;
        const int t = q >> 2, sl = q & 3, tb = t & 1, ut = t >> 1;
        if (sl == 0 && ut > 0) tc::mbar_wait(&bar_empty[tb], (uint32_t)((ut - 1) & 1));   // accumulator drained
        tc::tc_fence_after();
        constexpr uint32_t idesc = tc::idesc_tf32(128, 128, 0, 0);
        const uint32_t ah = tc::smem_u32(a_buf) + (2 * b) * UPD_A_PART, al = ah + UPD_A_PART;
        const uint32_t bh = tc::smem_u32(u_hi) + 8 * sl * UPD_LBO, bl = tc::smem_u32(u_lo) + 8 * sl * UPD_LBO;
        const uint32_t d = tmem + 128 * tb;
#pragma unroll
        for (int pass = 0; pass < 3; ++pass) {
            const uint32_t ab = pass == 1 ? al : ah;
            const uint32_t bb = pass == 2 ? bl : bh;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const uint64_t ad = tc::smem_desc(ab + 2 * j * UPD_LBO, UPD_LBO, 128);
                const uint64_t bd = tc::smem_desc(bb + 2 * j * UPD_LBO, UPD_LBO, 128);
                tc::umma_tf32(d, ad, bd, idesc, !(sl == 0 && pass == 0 && j == 0));
            }
        }
        tc::umma_commit(&bar_smem[b]);
        if (sl == 3) tc::umma_commit(&bar_full[tb]);
    };

    if (warp < UPD_PRODUCERS / 32) {
        // ===================== producers: load / split / stage / issue =====================
        const int c = tid & 127, hrow = tid >> 7;            // column of the tile, half (16 rows) of each slab
        const int aoff = (c & 7) * 16 + (c >> 3) * 128;
        const int nslab = ntiles * 4;
        float x[4][16];                                      // four k-slabs (this thread's half) in flight
        auto load_slab = [&](int q, float (&dst)[16]) {
            const int t = q >> 2, sl = q & 3;
            const long col = (long)(tile0 + t) * TILE + c;
            if (q < nslab && col < ld) {
#pragma unroll
                for (int kk = 0; kk < 16; ++kk) {
                    const int k = 32 * sl + 16 * hrow + kk;
                    const int grow = k < BJ_B ? row_a + k : row_b + (k - BJ_B);
                    dst[kk] = mt.X[(long)grow * ld + col];
                }
            } else {
#pragma unroll
                for (int kk = 0; kk < 16; ++kk) dst[kk] = 0.f;
            }
        };
        auto stage_slab = [&](int q, float (&src)[16]) {
            const int b = q & 1, u = q >> 1;
            if (u > 0) tc::mbar_wait(&bar_smem[b], (uint32_t)((u - 1) & 1));        // buffer b free again
            unsigned char* hi_p = a_buf + (2 * b) * UPD_A_PART;
            unsigned char* lo_p = hi_p + UPD_A_PART;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int kq = 4 * hrow + j;
                float4 hi, lo;
                tc::split_tf32(make_float4(src[4 * j], src[4 * j + 1], src[4 * j + 2], src[4 * j + 3]), hi, lo);
                *reinterpret_cast<float4*>(hi_p + aoff + kq * UPD_LBO) = hi;
                *reinterpret_cast<float4*>(lo_p + aoff + kq * UPD_LBO) = lo;
            }
            load_slab(q + 4, src);                           // refill the registers: in flight for four slabs
            tc::fence_async_smem();
            tc::named_bar_sync(1, UPD_PRODUCERS);
            if (tid == 0) issue_slab(q);
        };
        load_slab(0, x[0]); load_slab(1, x[1]); load_slab(2, x[2]); load_slab(3, x[3]);
        for (int q = 0; q < nslab; q += 4) {                 // nslab is a multiple of 4
            stage_slab(q, x[0]);
            stage_slab(q + 1, x[1]);
            stage_slab(q + 2, x[2]);
            stage_slab(q + 3, x[3]);
        }
    } else {
        // ===================== epilogue: TMEM -> rows of X, in place =====================
        const int w4 = warp & 3, lane = tid & 31;
        const int c = 32 * w4 + lane;                        // TMEM lane = column of the tile
        for (int t = 0; t < ntiles; ++t) {
            const int tb = t & 1, ut = t >> 1;
            tc::mbar_wait(&bar_full[tb], (uint32_t)(ut & 1));
            tc::tc_fence_after();
            const long col = (long)(tile0 + t) * TILE + c;
            const uint32_t taddr = tmem + ((uint32_t)(32 * w4) << 16) + 128 * tb;
            float v[32];
#pragma unroll
            for (int h = 0; h < 4; ++h) {
                tc::tmem_ld32(taddr + 32 * h, v);
                if (col < ld) {
#pragma unroll
                    for (int i = 0; i < 32; ++i) {
                        const int r = 32 * h + i;
                        const int grow = r < BJ_B ? row_a + r : row_b + (r - BJ_B);
                        mt.X[(long)grow * ld + col] = v[i];
                    }
                }
            }
            tc::tc_fence_before();
            tc::mbar_arrive(&bar_empty[tb]);
        }
    }
    tc::tc_fence_before();
    __syncthreads();
    if (warp == 0) tc::tmem_dealloc<256>(tmem);
}

// (2) inner eigenproblem: fixed-order sum of the H partials, symmetrise, one cyclic two-sided Jacobi
// sweep of the 128x128 block in shared memory, emit the accumulated rotations U.
// Element pairs inside one 64-row block meet in EVERY round of the outer sweep, pairs across the two
// blocks meet only here: bipartite 64-round tournament {(i, 64 + (i + r) % 64)} normally, the full
// 127-round tournament in the first round(s) of each outer sweep.
// In bipartite rounds S' = J^T S J decomposes into independent 2x2 blocks (pair a x pair b): one thread loads the
// 4 entries, applies the right (pair b) and left (pair a) rotation in registers and stores them -- each element of
// S is read and written once per round.  Rotation application is branch-free (skipped pairs carry c = 1, s = 0).
//
// Two layouts of the same arithmetic (bit-identical results), chosen by the host per group:
//   SPLIT = false  one CTA of 1024 threads per pair holds S and U (132 KB): best THROUGHPUT when there are
//                  more pairs than SMs (the sweep is bound by shared-memory bandwidth either way);
//   SPLIT = true   a CLUSTER OF TWO CTAs (512 threads, 66 KB each) per pair: best LATENCY when the pairs cannot
//                  fill the SMs (one large factor per GPU).
//     * CTA 0 ("S") owns the projected block S, computes the rotations of every round and applies them to S;
//       each round's 64 rotations {c, s, p, q} are published to a global log + a release flag;
//     * CTA 1 ("U") owns the accumulator U, follows the log (acquire on the flag) and applies the same
//       rotations to U.  U never feeds back into S, so the two run concurrently, U lagging by < 1 round.
//     The cluster guarantees co-scheduling (S never waits on U, so the spin in U cannot deadlock).
constexpr int SOLVE_P = BJ_M + 1;
constexpr int SOLVE_MAX_ROUNDS = BJ_M - 1;                   // 127 (full tournament)
constexpr int SOLVE_LOG_STRIDE = SOLVE_MAX_ROUNDS * (BJ_M / 2);   // float4 entries per pair
template <bool SPLIT>
struct SolveCfg {
    static constexpr int THREADS = SPLIT ? 512 : 1024;
    static constexpr int SMEM = (SPLIT ? 1 : 2) * BJ_M * SOLVE_P * (int)sizeof(float);
};

__device__ __forceinline__ int ld_acquire(const int* p) {
    int v;
    asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release(int* p, int v) {
    asm volatile("st.release.gpu.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

// X <- X J for the rotations rb[0..64) of one round; this thread: row `row`, pairs grp, grp + GROUPS, ...
// (the pairs of a round are disjoint: loads of a chunk first, then its stores)
template <int GROUPS, int CH>
__device__ __forceinline__ void solve_rotate_columns(float* X, const float4* rb, int row, int grp) {
    constexpr int P = SOLVE_P, PPT = (BJ_M / 2) / GROUPS;
#pragma unroll
    for (int u0 = 0; u0 < PPT; u0 += CH) {
        float4 rp[CH];
        float xp[CH], xq[CH];
#pragma unroll
        for (int u = 0; u < CH; ++u) {
            rp[u] = rb[grp + GROUPS * (u0 + u)];
            xp[u] = X[row * P + __float_as_int(rp[u].z)];
            xq[u] = X[row * P + __float_as_int(rp[u].w)];
        }
#pragma unroll
        for (int u = 0; u < CH; ++u) {
            X[row * P + __float_as_int(rp[u].z)] = rp[u].x * xp[u] - rp[u].y * xq[u];
            X[row * P + __float_as_int(rp[u].w)] = rp[u].y * xp[u] + rp[u].x * xq[u];
        }
    }
}

template <bool SPLIT>
__global__ void __launch_bounds__(SolveCfg<SPLIT>::THREADS, 1)
bj_solve_kernel(const BjMat* mats, const int* pair_mat, BjState* state, int round,
                const float* __restrict__ partial, const long* partial_off,
                float* __restrict__ ubuf, float4* __restrict__ rotlog, int* __restrict__ rotflag,
                int epoch, int full_inner) {
    constexpr int M = BJ_M, P = SOLVE_P, HALF = BJ_M / 2, MM1 = BJ_M - 1;
    constexpr int THREADS = SolveCfg<SPLIT>::THREADS;
    constexpr int GROUPS = THREADS / M;                      // 4 | 8 row groups
    constexpr int RCH = SPLIT ? 16 : 8;                      // rotations per load/store batch (the split layout has registers to spare)
    extern __shared__ float sm_f[];
    float* S = sm_f;                                         // SPLIT: the matrix of this CTA's role (S or U)
    float* U = SPLIT ? sm_f : sm_f + M * P;
    __shared__ float4 rot[2][HALF];                          // S side uses rot[0]; the U follower double-buffers
    __shared__ float red[THREADS / 32];
    __shared__ int s_avail[2];
    const int pair_idx = SPLIT ? (int)(blockIdx.x >> 1) : (int)blockIdx.x;
    const int role = SPLIT ? (int)(blockIdx.x & 1) : 0;      // 0 = S (+ U when not split), 1 = U follower
    const int mi = pair_mat[pair_idx];
    const BjMat mt = mats[mi];
    if (state[mi].converged) return;                         // (both CTAs of a cluster take the same exit)
    int bi, bj;
    if (!bj_pair_blocks(mt, pair_idx - mt.pair_begin, round, bi, bj)) return;
    const int tid = threadIdx.x;
    // (with an odd block count one block idles per round: two leading full rounds cover every block)
    const bool full = full_inner || (round % (mt.p - 1)) < ((mt.nblk & 1) ? 2 : 1);
    const int nrounds = full ? MM1 : BJ_B;
    float4* log = rotlog + (long)pair_idx * SOLVE_LOG_STRIDE;
    int* flag = rotflag + pair_idx;
    const int flag_base = epoch << 8;                        // rounds < 256: (epoch, round) packs into one int
    const int row = tid & (M - 1), grp = tid / M;
    float* dst = ubuf + (long)pair_idx * (M * M);

    if (SPLIT && role == 1) {
        // ============================ U: follow the rotation log ============================
        for (int e = tid; e < M * M; e += THREADS) {
            const int i = e / M, j = e - i * M;
            U[i * P + j] = (i == j) ? 1.f : 0.f;
        }
        // `avail` = rounds known to be published.  The flag is polled only when U has caught up with S; while
        // it lags, the next round's rotations are prefetched into shared memory (cp.async, L2 only) during the
        // current round, so the L2 latency of flag + log is off the critical path.
        int avail = 0;
        bool pref = false;
        for (int r = 0; r < nrounds; ++r) {
            float4* rb = rot[r & 1];
            if (pref) {
                if (tid < HALF) asm volatile("cp.async.wait_all;" ::: "memory");
                __syncthreads();
            } else {                                         // (r >= avail here)
                if (tid == 0) {
                    int v;
                    while ((v = ld_acquire(flag)) < flag_base + r + 1) { }
                    s_avail[r & 1] = v - flag_base;
                }
                __syncthreads();                             // round r is published; round r-1 of U is complete
                avail = s_avail[r & 1];
                if (tid < HALF) rb[tid] = __ldcg(log + r * HALF + tid);
                __syncthreads();
            }
            pref = r + 1 < avail;
            if (pref && tid < HALF) {
                const unsigned d = (unsigned)__cvta_generic_to_shared(&rot[(r + 1) & 1][tid]);
                asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(log + (r + 1) * HALF + tid)
                             : "memory");
                asm volatile("cp.async.commit_group;" ::: "memory");
            }
            solve_rotate_columns<GROUPS, RCH>(U, rb, row, grp);
        }
        __syncthreads();
        for (int e = tid; e < M * M; e += THREADS) {
            const int i = e / M, j = e - i * M;
            dst[e] = U[i * P + j];
        }
        return;
    }

    // ============================ S: rotations (+ U when not split) ============================
    const float* src = partial + partial_off[mi] +
                       (long)(pair_idx - mt.pair_begin) * mt.ksplit * (M * M);
    // sum the split-K partials in chunk order (fixed => deterministic): float4 loads, 4 independent elements in flight
    {
        const float4* src4 = reinterpret_cast<const float4*>(src);
        constexpr int NH = (M * M / 4) / THREADS;            // float4 elements per thread: 4 | 8
#pragma unroll
        for (int h0 = 0; h0 < NH; h0 += 4) {
            float4 acc[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) acc[u] = make_float4(0.f, 0.f, 0.f, 0.f);
            for (int z = 0; z < mt.ksplit; ++z) {
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const float4 t = __ldg(src4 + (long)z * (M * M / 4) + tid + (h0 + u) * THREADS);
                    acc[u].x += t.x; acc[u].y += t.y; acc[u].z += t.z; acc[u].w += t.w;
                }
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int e4 = tid + (h0 + u) * THREADS;
                const int i = (e4 * 4) / M, j = (e4 * 4) - i * M;
                S[i * P + j] = acc[u].x; S[i * P + j + 1] = acc[u].y;
                S[i * P + j + 2] = acc[u].z; S[i * P + j + 3] = acc[u].w;
            }
        }
    }
    __syncthreads();
    // H is symmetric only up to rounding: S = (H + H^T)/2, in place (each unordered pair by one thread)
    float off2 = 0.f;
    for (int e = tid; e < M * M; e += THREADS) {
        const int i = e / M, j = e - i * M;
        if (i < j) {
            const float v = 0.5f * (S[i * P + j] + S[j * P + i]);
            S[i * P + j] = v; S[j * P + i] = v;
            if (i < BJ_B && j >= BJ_B) off2 = fmaf(v, v, off2);     // cross block (counted once)
        }
        if (!SPLIT) U[i * P + j] = (i == j) ? 1.f : 0.f;
    }
    off2 = warp_sum(off2);
    if ((tid & 31) == 0) red[tid >> 5] = off2;
    __syncthreads();
    if (tid == 0) {
        float t = 0.f;
        for (int i = 0; i < THREADS / 32; ++i) t += red[i];
        if (t > 0.f) atomicAdd(&state[mi].off2, 2.f * t);
    }
    const float abs_thr = 5e-8f * __int_as_float(state[mi].scale_bits);
    for (int r = 0; r < nrounds; ++r) {
        if (tid < HALF) {
            int p, q;
            if (full) {
                if (tid == 0) { p = MM1; q = r; }
                else { p = r + tid; if (p >= MM1) p -= MM1; q = r - tid; if (q < 0) q += MM1; }
                if (p > q) { const int t = p; p = q; q = t; }
            } else {
                p = tid; q = BJ_B + ((tid + r) & (BJ_B - 1));
            }
            const float app = S[p * P + p], aqq = S[q * P + q], apq = S[p * P + q];
            float c = 1.f, sn = 0.f;
            const float mag = fabsf(apq);
            if (mag > 1.0e-7f * sqrtf(fabsf(app) * fabsf(aqq)) && mag > abs_thr) {
                // angle from fast divide / sqrt (an angle error only leaves a residual a_pq);
                // (c, s) is then re-normalised so that c^2 + s^2 = 1 to fp32 rounding.
                const float tau = __fdividef(aqq - app, 2.f * apq);
                const float t = __fdividef(tau >= 0.f ? 1.f : -1.f, fabsf(tau) + __fsqrt_rn(fmaf(tau, tau, 1.f)));
                c = rsqrtf(fmaf(t, t, 1.f));
                sn = t * c;
                const float f = fmaf(-0.5f, fmaf(c, c, sn * sn), 1.5f);
                c *= f; sn *= f;
            }
            const float4 rv = make_float4(c, sn, __int_as_float(p), __int_as_float(q));
            rot[0][tid] = rv;
            if (SPLIT) log[r * HALF + tid] = rv;             // for the U CTA
        }
        __syncthreads();
        if (SPLIT && tid == 0) st_release(flag, flag_base + r + 1);   // publishes the log entries written before the barrier
        if (!full) {
            // 2x2 blocks (pair a, pair b); lanes walk b: conflict-free (pitch 129)
            const int b = tid & (HALF - 1);
            const float4 rb = rot[0][b];
            const int qb = __float_as_int(rb.w);
            constexpr int AG = THREADS / HALF;                          // 8 | 16 groups of pair a
            constexpr int NIT = HALF / AG;                              // 8 | 4 2x2 blocks per thread ...
            constexpr int CH = SPLIT ? 8 : 4;                           // ... in chunks: 16 | 32 loads in flight
#pragma unroll
            for (int i0 = 0; i0 < NIT; i0 += CH) {
                float4 ra[CH];
                float x11[CH], x12[CH], x21[CH], x22[CH];
#pragma unroll
                for (int i = 0; i < CH; ++i) {
                    const int a = (tid >> 6) + AG * (i0 + i);
                    ra[i] = rot[0][a];
                    const int qa = __float_as_int(ra[i].w);
                    x11[i] = S[a * P + b]; x12[i] = S[a * P + qb]; x21[i] = S[qa * P + b]; x22[i] = S[qa * P + qb];
                }
#pragma unroll
                for (int i = 0; i < CH; ++i) {
                    const int a = (tid >> 6) + AG * (i0 + i);
                    const int qa = __float_as_int(ra[i].w);
                    const float y11 = rb.x * x11[i] - rb.y * x12[i], y12 = rb.y * x11[i] + rb.x * x12[i];
                    const float y21 = rb.x * x21[i] - rb.y * x22[i], y22 = rb.y * x21[i] + rb.x * x22[i];
                    float z11 = ra[i].x * y11 - ra[i].y * y21, z21 = ra[i].y * y11 + ra[i].x * y21;
                    float z12 = ra[i].x * y12 - ra[i].y * y22, z22 = ra[i].y * y12 + ra[i].x * y22;
                    if (a == b && ra[i].y != 0.f) { z12 = 0.f; z21 = 0.f; }     // the annihilated element
                    S[a * P + b] = z11; S[a * P + qb] = z12; S[qa * P + b] = z21; S[qa * P + qb] = z22;
                }
            }
            if (!SPLIT) solve_rotate_columns<GROUPS, RCH>(U, rot[0], row, grp);     // U <- U J (no barrier needed before)
            __syncthreads();
            continue;
        }
        // generic (full tournament) round: column pass (S and U), barrier, row pass
        solve_rotate_columns<GROUPS, RCH>(S, rot[0], row, grp);
        if (!SPLIT) solve_rotate_columns<GROUPS, RCH>(U, rot[0], row, grp);
        __syncthreads();
        {
            constexpr int PPT = HALF / GROUPS, CH = 8;
#pragma unroll
            for (int u0 = 0; u0 < PPT; u0 += CH) {
                float4 rp[CH];
                float sp[CH], sq[CH];
#pragma unroll
                for (int u = 0; u < CH; ++u) {
                    rp[u] = rot[0][grp + GROUPS * (u0 + u)];
                    sp[u] = S[__float_as_int(rp[u].z) * P + row];
                    sq[u] = S[__float_as_int(rp[u].w) * P + row];
                }
#pragma unroll
                for (int u = 0; u < CH; ++u) {
                    const int pp = __float_as_int(rp[u].z), qq = __float_as_int(rp[u].w);
                    float np = rp[u].x * sp[u] - rp[u].y * sq[u];
                    float nq = rp[u].y * sp[u] + rp[u].x * sq[u];
                    if (rp[u].y != 0.f) { if (row == qq) np = 0.f; if (row == pp) nq = 0.f; }   // the rotated pair element
                    S[pp * P + row] = np;
                    S[qq * P + row] = nq;
                }
            }
        }
        __syncthreads();
    }
    if (!SPLIT) {
        for (int e = tid; e < M * M; e += THREADS) {
            const int i = e / M, j = e - i * M;
            dst[e] = U[i * P + j];
        }
    }
}

// end-of-round bookkeeping: sweep boundaries, convergence flags, active count for the host
__global__ void bj_check_kernel(const BjMat* mats, BjState* state, int nmat, int round,
                                int max_sweeps, volatile int* host_active, int one_sweep) {
    __shared__ int active;
    if (threadIdx.x == 0) active = 0;
    __syncthreads();
    for (int mi = threadIdx.x; mi < nmat; mi += blockDim.x) {
        BjState st = state[mi];
        if (!st.converged) {
            const int pm1 = mats[mi].p - 1;
            if ((round + 1) % pm1 == 0) {
                st.sweeps += 1;
                // off-block-diagonal norm seen DURING this sweep; Jacobi converges quadratically, so once
                // it is below tol * ||A||_F the sweep just finished has pushed it to fp32 noise level.
                const float tol = mats[mi].tol_conv;
                if (one_sweep || st.off2 <= tol * tol * st.norm2 || st.sweeps >= max_sweeps) st.converged = 1;
                st.off2 = 0.f;
                state[mi] = st;
            }
            if (!st.converged) atomicAdd(&active, 1);
        }
    }
    __syncthreads();
    if (threadIdx.x == 0 && host_active) *host_active = active;
}

// arms the host-visible progress flag of a group IN STREAM ORDER (a host store would race with the still queued
// kernels of the previous call, which write the same flag)
__global__ void bj_arm_kernel(volatile int* host_active) { *host_active = 1; }

// lam_i = (v_i . y_i) / (v_i . v_i) with y_i = A v_i stored in the W half.  one warp per row
__global__ void bj_rayleigh_kernel(const BjMat* mats) {
    const BjMat mt = mats[blockIdx.y];
    const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (row >= mt.n) return;
    const long ld = 2L * mt.n_pad;
    const float* y = mt.X + (long)row * ld;
    const float* v = y + mt.n_pad;
    double num = 0.0, den = 0.0;
    for (int c = threadIdx.x & 31; c < mt.n; c += 32) {
        const float vv = v[c];
        num += (double)vv * (double)y[c];
        den += (double)vv * (double)vv;
    }
    num = warp_sum(num); den = warp_sum(den);
    if ((threadIdx.x & 31) == 0) {
        mt.lam[row] = (float)(num / den);
        mt.nrm[row] = (float)den;
    }
}

// ascending sort of the eigenvalues of one matrix (bitonic in shared memory), emits perm + d
__global__ void __launch_bounds__(1024)
bj_sort_kernel(const BjMat* mats, float eps_floor) {
    extern __shared__ unsigned char raw[];
    const BjMat mt = mats[blockIdx.x];
    int np2 = 1;
    while (np2 < mt.n) np2 <<= 1;
    float* key = reinterpret_cast<float*>(raw);
    int* idx = reinterpret_cast<int*>(raw + (size_t)np2 * sizeof(float));
    for (int i = threadIdx.x; i < np2; i += blockDim.x) {
        key[i] = i < mt.n ? mt.lam[i] : INFINITY;
        idx[i] = i;
    }
    __syncthreads();
    for (int k = 2; k <= np2; k <<= 1) {
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int i = threadIdx.x; i < np2; i += blockDim.x) {
                const int ixj = i ^ j;
                if (ixj > i) {
                    const bool up = (i & k) == 0;
                    const float a = key[i], b = key[ixj];
                    const int ia = idx[i], ib = idx[ixj];
                    const bool gt = (a > b) || (a == b && ia > ib);
                    if (gt == up) { key[i] = b; key[ixj] = a; idx[i] = ib; idx[ixj] = ia; }
                }
            }
            __syncthreads();
        }
    }
    for (int i = threadIdx.x; i < mt.n; i += blockDim.x) {
        mt.perm[i] = idx[i];
        const float v = key[i];
        mt.d[i] = v > eps_floor ? v : 0.f;
    }
}

// Q[r][c] = V[r, perm[c]] / ||v||  = X[perm[c]][n_pad + r] * rsqrt(nrm).  32x32 transpose tiles
__global__ void bj_write_kernel(const BjMat* mats) {
    __shared__ float t[32][33];
    const BjMat mt = mats[blockIdx.z];
    const int c0 = blockIdx.x * 32, r0 = blockIdx.y * 32;
    if (c0 >= mt.n || r0 >= mt.n) return;
    const long ld = 2L * mt.n_pad;
    const int lx = threadIdx.x & 31, ly = threadIdx.x >> 5;    // 256 threads: ly in [0,8)
    for (int cc = ly; cc < 32; cc += 8) {
        const int c = c0 + cc, r = r0 + lx;
        float v = 0.f;
        if (c < mt.n && r < mt.n) {
            const int src = mt.perm[c];
            v = mt.X[(long)src * ld + mt.n_pad + r] * rsqrtf(mt.nrm[src]);
        }
        t[cc][lx] = v;
    }
    __syncthreads();
    for (int rr = ly; rr < 32; rr += 8) {
        const int r = r0 + rr, c = c0 + lx;
        if (r < mt.n && c < mt.n) mt.Q[(long)r * mt.n + c] = t[lx][rr];
    }
}

// re-arm every matrix of a group for the single clean-up sweep
__global__ void bj_rearm_kernel(BjState* state, int nmat) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < nmat) { state[i].converged = 0; state[i].off2 = 0.f; }
}

__global__ void bj_store_sweeps_kernel(const BjState* state, const int* out_index, int nmat, int* sweeps_out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < nmat) sweeps_out[out_index[i]] = state[i].sweeps;
}

// Newton-Schulz re-orthogonalisation of the accumulated eigenvector rows:
//   E = V V^T - I  (into the W half of X),  V <- V - 0.5 E V  (via the Q buffer)
struct EpiGramMinusI {
    static constexpr bool UPPER_ONLY = true;      // V V^T is symmetric: tiles above the diagonal are mirrored
    float* X; long ld;
    __device__ __forceinline__ void operator()(int m, int n, float v) const {
        X[(long)m * ld + n] = v - (m == n ? 1.f : 0.f);
        if (m / TILE != n / TILE) X[(long)n * ld + m] = v;
    }
};
struct EpiNsUpdate {
    float* Q; const float* X; long ld; int n, n_pad;
    __device__ __forceinline__ void operator()(int m, int c, float v) const {
        Q[(long)m * n + c] = fmaf(-0.5f, v, X[(long)m * ld + n_pad + c]);
    }
};

// epilogue of the Rayleigh GEMM: Y -> W half of X (row stride 2 n_pad)
struct EpiToX {
    float* X; long ld;
    __device__ __forceinline__ void operator()(int m, int n, float v) const { X[(long)m * ld + n] = v; }
};

// ---------------------------------------------------------------------------
// workspace plan
// ---------------------------------------------------------------------------
// ---------------------------------------------------------------------------
// workspace plan.  The blocked-path matrices are split into up to MAX_GROUPS groups (longest-processing-time
// on n^3); every group runs its own round loop on its own internal stream, so the latency-bound solve kernel of
// one group overlaps with the bandwidth-bound GEMM kernels of the others.
// ---------------------------------------------------------------------------
constexpr int MAX_GROUPS = 4;

struct GroupPlan {
    int nb = 0, total_pairs = 0, max_ksplit = 1, max_rounds = 0, nmax = 0;
    int kchunk = BJ_KCHUNK;       // k-chunk of the projected-block CTAs (wave-quantisation heuristic below)
    int upd_tiles = 8;            // 128-column tiles per update CTA (same)
    int solve_split = 0;          // 1 = two-CTA cluster per pair in the inner solve (latency layout)
    long max_ld = 0;
    int idx[512];                 // indices (into the caller's arrays) of this group's matrices
    size_t off_bmats = 0, off_bindex = 0, off_pairmat = 0, off_partoff = 0, off_state = 0, off_ubuf = 0,
           off_partial = 0, off_perm = 0, off_lam = 0, off_x = 0, off_rotlog = 0, off_rotflag = 0;
};
struct EighPlan {
    int n_direct = 0, n_block = 0, ngroups = 0;
    size_t off_dmats = 0, off_dindex = 0, total = 0;
    GroupPlan g[MAX_GROUPS];
};

static double g_profile[8] = {0};      // last KFACB_PROFILE=1 run: rounds, ms H/solve/update, bytes H/update, matrices, pairs

static int env_int(const char* name, int dflt) {
    const char* v = getenv(name);
    return (v && *v) ? atoi(v) : dflt;
}

static int bj_npad(int n) { return cdiv(n, BJ_B) * BJ_B; }

// SM count of the current device (148 on B200; also the fallback when no device is visible, e.g. a
// workspace-size query in a CPU-only process)
static int sm_count_cached() {
    static int sms = 0;
    if (sms == 0) {
        int dev = 0, v = 0;
        if (cudaGetDevice(&dev) == cudaSuccess &&
            cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess && v > 0) sms = v;
        else { sms = 148; (void)cudaGetLastError(); }
    }
    return sms;
}

static bool make_plan(const int32_t* n, int count, EighPlan& pl) {
    // block-path matrices, largest first (stable), dealt to the least-loaded group
    int order[512 * MAX_GROUPS];
    int nblock = 0;
    for (int i = 0; i < count; ++i) {
        if (n[i] <= 0) continue;
        if (n[i] <= DIRECT_MAX_N) { pl.n_direct++; continue; }
        if (nblock >= 512 * MAX_GROUPS) return false;
        order[nblock++] = i;
    }
    pl.n_block = nblock;
    for (int a = 1; a < nblock; ++a) {           // insertion sort by n descending (stable)
        const int v = order[a];
        int b = a - 1;
        while (b >= 0 && n[order[b]] < n[v]) { order[b + 1] = order[b]; --b; }
        order[b + 1] = v;
    }
    int want = env_int("KFACB_EIGH_GROUPS", 3);
    if (env_int("KFACB_PROFILE", 0)) want = 1;   // the per-round profiler serialises one group
    if (want < 1) want = 1;
    if (want > MAX_GROUPS) want = MAX_GROUPS;
    pl.ngroups = nblock < want ? nblock : want;
    double load[MAX_GROUPS] = {0};
    for (int a = 0; a < nblock; ++a) {
        int best = 0;
        for (int k = 1; k < pl.ngroups; ++k) if (load[k] < load[best]) best = k;
        GroupPlan& g = pl.g[best];
        if (g.nb >= 512) return false;
        g.idx[g.nb++] = order[a];
        load[best] += (double)n[order[a]] * n[order[a]] * n[order[a]];
    }
    size_t off = 0;
    auto take = [&](size_t bytes) { size_t o = off; off += align_up(bytes, 256); return o; };
    pl.off_dmats = take(sizeof(DirectMat) * (size_t)(pl.n_direct + 1));
    pl.off_dindex = take(sizeof(int) * (size_t)(pl.n_direct + 1));
    for (int k = 0; k < pl.ngroups; ++k) {
        GroupPlan& g = pl.g[k];
        size_t x_bytes = 0, part_floats = 0, lam_floats = 0, perm_ints = 0;
        // Launch shapes that fill whole waves.  Both kernels are latency-bound per CTA, so the time of a launch is
        // (CTAs an SM has to run one after the other) x (length of one CTA):
        //   projected block: 3 CTAs / SM resident, one CTA = kchunk / 32 k-slabs + ~3 slabs of set-up / drain;
        //   update         : 1 CTA / SM, one CTA = upd_tiles tiles + ~0.8 tile of set-up (U staging, pipeline fill).
        // The heuristics apply when this group has the GPU (nearly) to itself: a single group, or one whose n^3 load
        // dominates the call (one big factor per rank in multi-GPU runs).  With several balanced groups in flight
        // the kernels of the others fill the gaps and the defaults (largest chunks, least set-up work) are best.
        {
            const int sms = sm_count_cached();
            const int kc_env = env_int("KFACB_H_KCHUNK", 0), ut_env = env_int("KFACB_UPD_TILES", 0);
            const int ss_env = env_int("KFACB_SOLVE_SPLIT", -1);
            double total_load = 0;
            for (int q = 0; q < pl.ngroups; ++q) total_load += load[q];
            const bool alone = pl.ngroups == 1 || load[k] >= 0.6 * total_load;
            long pairs = 0;
            for (int a = 0; a < g.nb; ++a) pairs += (bj_npad(n[g.idx[a]]) / BJ_B + 1) / 2;
            g.solve_split = alone && 2 * pairs <= sms;       // the pairs cannot fill the SMs: two SMs per pair
            if (ss_env >= 0) g.solve_split = ss_env != 0;
            double best = 1e300;
            for (int c = BJ_KCHUNK; alone && c >= 256; c -= 128) {
                long ctas = 0;
                for (int a = 0; a < g.nb; ++a) {
                    const int np = bj_npad(n[g.idx[a]]);
                    ctas += (long)((np / BJ_B + 1) / 2) * cdiv(np, c);
                }
                const double cost = (double)cdiv(ctas, 3L * sms) * (c / 32 + 3);
                if (cost < best) { best = cost; g.kchunk = c; }
            }
            if (kc_env >= 32 && kc_env <= BJ_KCHUNK && kc_env % 32 == 0) g.kchunk = kc_env;
            best = 1e300;
            for (int t = 16; alone && t >= 4; --t) {
                long ctas = 0;
                for (int a = 0; a < g.nb; ++a) {
                    const int np = bj_npad(n[g.idx[a]]);
                    ctas += (long)((np / BJ_B + 1) / 2) * cdiv(2 * np / TILE, t);
                }
                const double cost = (double)cdiv(ctas, (long)sms) * (t + 0.8);
                if (cost < best) { best = cost; g.upd_tiles = t; }
            }
            if (ut_env > 0) g.upd_tiles = ut_env;
        }
        for (int a = 0; a < g.nb; ++a) {
            const int nn = n[g.idx[a]];
            const int np = bj_npad(nn);
            const int nblk = np / BJ_B;
            const int p = (nblk + 1) & ~1;
            const int ks = cdiv(np, g.kchunk);
            g.total_pairs += p / 2;
            if (ks > g.max_ksplit) g.max_ksplit = ks;
            if ((p - 1) * 64 > g.max_rounds) g.max_rounds = (p - 1) * 64;   // hard cap; the device-side sweep limit ends earlier
            if (2L * np > g.max_ld) g.max_ld = 2L * np;
            if (nn > g.nmax) g.nmax = nn;
            x_bytes += align_up((size_t)np * 2 * np * sizeof(float), 256);
            part_floats += (size_t)(p / 2) * ks * BJ_M * BJ_M;
            lam_floats += 2 * (size_t)np;
            perm_ints += (size_t)np;
        }
        g.off_bmats = take(sizeof(BjMat) * (size_t)(g.nb + 1));
        g.off_bindex = take(sizeof(int) * (size_t)(g.nb + 1));
        g.off_pairmat = take(sizeof(int) * (size_t)(g.total_pairs + 1));
        g.off_partoff = take(sizeof(long) * (size_t)(g.nb + 1));
        g.off_state = take(sizeof(BjState) * (size_t)(g.nb + 1));
        g.off_ubuf = take(sizeof(float) * (size_t)g.total_pairs * BJ_M * BJ_M);
        g.off_rotlog = take(sizeof(float4) * (size_t)g.total_pairs * SOLVE_LOG_STRIDE);
        g.off_rotflag = take(sizeof(int) * (size_t)(g.total_pairs + 1));
        g.off_partial = take(sizeof(float) * part_floats);
        g.off_perm = take(sizeof(int) * perm_ints);
        g.off_lam = take(sizeof(float) * lam_floats);
        g.off_x = take(x_bytes);
    }
    pl.total = off + 256;
    return true;
}

// pinned, device-mapped progress flags (one per group) + the internal streams / events of the groups
struct GroupRuntime {
    int* h_active = nullptr;      // [MAX_GROUPS] host view
    int* d_active = nullptr;      // device view
    cudaStream_t stream[MAX_GROUPS] = {nullptr};
    cudaEvent_t fork = nullptr, join[MAX_GROUPS] = {nullptr}, prog[MAX_GROUPS][2] = {{nullptr}};
    bool ok = false;
};
static GroupRuntime* group_runtime() {
    static GroupRuntime rt;
    if (!rt.ok) {
        if (cudaHostAlloc((void**)&rt.h_active, sizeof(int) * MAX_GROUPS, cudaHostAllocMapped) != cudaSuccess) return nullptr;
        if (cudaHostGetDevicePointer((void**)&rt.d_active, rt.h_active, 0) != cudaSuccess) return nullptr;
        if (cudaEventCreateWithFlags(&rt.fork, cudaEventDisableTiming) != cudaSuccess) return nullptr;
        for (int k = 0; k < MAX_GROUPS; ++k) {
            if (cudaStreamCreateWithFlags(&rt.stream[k], cudaStreamNonBlocking) != cudaSuccess) return nullptr;
            if (cudaEventCreateWithFlags(&rt.join[k], cudaEventDisableTiming) != cudaSuccess) return nullptr;
            for (int e = 0; e < 2; ++e)
                if (cudaEventCreateWithFlags(&rt.prog[k][e], cudaEventDisableTiming) != cudaSuccess) return nullptr;
        }
        rt.ok = true;
    }
    return &rt;
}

// host-side state of one group while its rounds are being enqueued
struct GroupRun {
    const GroupPlan* gp = nullptr;
    BjMat* hm = nullptr;
    const BjMat* dm = nullptr;
    const int* dpm = nullptr;
    const long* dpoff = nullptr;
    BjState* dstate = nullptr;
    float* ubuf = nullptr;
    float* partial = nullptr;
    float4* rotlog = nullptr;
    int* rotflag = nullptr;
    int epoch = 0;                // solve launches so far (the rotation-log flags are monotonic in it)
    cudaStream_t st = nullptr;
    bool done = false;
};

// one round of inner eigenproblems for every pair of the group (layout chosen in make_plan)
static cudaError_t launch_solve(GroupRun& gr, const GroupPlan& gp, int round, int full_inner, cudaStream_t st) {
    const int epoch = ++gr.epoch;
    if (!gp.solve_split) {
        bj_solve_kernel<false><<<gp.total_pairs, SolveCfg<false>::THREADS, SolveCfg<false>::SMEM, st>>>(
            gr.dm, gr.dpm, gr.dstate, round, gr.partial, gr.dpoff, gr.ubuf, gr.rotlog, gr.rotflag, epoch, full_inner);
        return cudaGetLastError();
    }
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(2 * gp.total_pairs);
    cfg.blockDim = dim3(SolveCfg<true>::THREADS);
    cfg.dynamicSmemBytes = SolveCfg<true>::SMEM;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, bj_solve_kernel<true>, gr.dm, gr.dpm, (BjState*)gr.dstate, round,
                              (const float*)gr.partial, gr.dpoff, gr.ubuf, gr.rotlog, gr.rotflag, epoch, full_inner);
}

}  // namespace kfacb

using namespace kfacb;

// n > 128: tridiagonalisation + divide and conquer (eigh_trd.cu / eigh_dc.cu) unless KFACB_EIGH_ALGO=jacobi asks
// for the round-1 block-Jacobi path below (kept for A/B comparisons)
static bool use_block_jacobi() {
    static int v = -1;
    if (v < 0) { const char* e = getenv("KFACB_EIGH_ALGO"); v = (e && (e[0] == 'j' || e[0] == 'J')) ? 1 : 0; }
    return v == 1;
}
extern "C" int kfacb_eigh_last_profile(double* out8) {
    if (!out8) return KFACB_ERR_ARG;
    if (!use_block_jacobi()) return eigh_tdc_last_profile(out8);
    for (int i = 0; i < 8; ++i) out8[i] = g_profile[i];
    return KFACB_OK;
}

static size_t direct_tables_bytes(int n_direct) {
    return align_up(sizeof(DirectMat) * (size_t)(n_direct + 1), 256) + align_up(sizeof(int) * (size_t)(n_direct + 1), 256);
}
// In-shared-memory Jacobi up to this size; above, tridiagonalisation + divide and conquer.  (The Jacobi kernel handles
// up to DIRECT_MAX_N = 128, but one 128-dim matrix keeps its single CTA busy for ~5 ms -- a serial stage in front
// of every step -- while the other path does it in ~1 ms.)
constexpr int TDC_DIRECT_N = 64;
static int collect_blocked(const int32_t* n, int count, int* idx) {
    int nb = 0;
    for (int i = 0; i < count; ++i) if (n[i] > TDC_DIRECT_N) idx[nb++] = i;
    return nb;
}

extern "C" size_t kfacb_eigh_workspace_bytes(const int32_t* n, int32_t count) {
    if (!n || count <= 0) return 256;
    size_t t = 0;
    if (use_block_jacobi()) {
        EighPlan* pl = new EighPlan();
        const bool ok = make_plan(n, count, *pl);
        t = ok ? pl->total : 0;
        delete pl;
        return t;
    }
    int* idx = new int[count];
    const int nb = collect_blocked(n, count, idx);
    t = direct_tables_bytes(count - nb) + eigh_tdc_workspace_bytes(n, count, idx, nb) + 256;
    delete[] idx;
    return t;
}

extern "C" int kfacb_eigh_batched(const int32_t* n, int32_t count,
                                  const float* const* M, float* const* Q, float* const* d,
                                  float eps_floor, int32_t* sweeps_out,
                                  void* workspace, size_t workspace_bytes,
                                  kfacb_stream_t stream) {
    cudaStream_t st = (cudaStream_t)stream;
    KFACB_REQUIRE(n && M && Q && d && count >= 0, "null argument");
    if (!use_block_jacobi()) {
        // direct Jacobi for n <= 64, tridiagonalisation + divide and conquer above; only enqueues work
        KFACB_REQUIRE(workspace && !((uintptr_t)workspace & 255), "eigh workspace missing or misaligned");
        for (int i = 0; i < count; ++i)
            if (n[i] > 0) KFACB_REQUIRE(M[i] && Q[i] && d[i], "matrix %d: null pointer", i);
        int* idx = new int[count > 0 ? count : 1];
        struct IdxGuard { int* p; ~IdxGuard() { delete[] p; } } idx_guard{idx};
        const int nb = collect_blocked(n, count, idx);
        int n_direct = 0;
        for (int i = 0; i < count; ++i) n_direct += (n[i] > 0 && n[i] <= TDC_DIRECT_N);
        const size_t dbytes = direct_tables_bytes(count - nb);
        if (workspace_bytes < dbytes) { set_error("eigh workspace too small"); return KFACB_ERR_WORKSPACE; }
        char* ws = (char*)workspace;
        if (n_direct > 0) {
            DirectMat* hm = new DirectMat[n_direct];
            int* hidx = new int[n_direct];
            int k = 0, nmax = 0;
            for (int i = 0; i < count; ++i) {
                if (n[i] <= 0 || n[i] > TDC_DIRECT_N) continue;
                hm[k] = DirectMat{M[i], Q[i], d[i], n[i]};
                hidx[k] = i;
                if (n[i] > nmax) nmax = n[i];
                ++k;
            }
            char* dmats = ws;
            char* dindex = ws + align_up(sizeof(DirectMat) * (size_t)(count - nb + 1), 256);
            cudaError_t e1 = cudaMemcpyAsync(dmats, hm, sizeof(DirectMat) * k, cudaMemcpyHostToDevice, st);
            cudaError_t e2 = cudaMemcpyAsync(dindex, hidx, sizeof(int) * k, cudaMemcpyHostToDevice, st);
            delete[] hm; delete[] hidx;
            KFACB_CHECK_CUDA(e1); KFACB_CHECK_CUDA(e2);
            const int m = (nmax + 1) & ~1;
            const size_t smem = 2 * (size_t)m * (m + 1) * sizeof(float);
            KFACB_CHECK_CUDA(cudaFuncSetAttribute(jacobi_direct_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            const int threads = nmax <= 32 ? 128 : (nmax <= 64 ? 256 : 512);
            jacobi_direct_kernel<<<k, threads, smem, st>>>((const DirectMat*)dmats, eps_floor, sweeps_out, (const int*)dindex);
            KFACB_CHECK_LAUNCH();
        }
        return eigh_tdc_run(n, idx, nb, M, Q, d, eps_floor, sweeps_out, ws + dbytes, workspace_bytes - dbytes, st);
    }
    EighPlan* plp = new EighPlan();
    struct PlanGuard { EighPlan* p; ~PlanGuard() { delete p; } } plan_guard{plp};
    KFACB_REQUIRE(make_plan(n, count, *plp), "too many matrices in one eigh batch");
    const EighPlan& pl = *plp;
    if (workspace_bytes < pl.total || ((uintptr_t)workspace & 255)) {
        set_error("eigh workspace too small or misaligned (%zu < %zu)", workspace_bytes, pl.total);
        return KFACB_ERR_WORKSPACE;
    }
    char* ws = (char*)workspace;
    for (int i = 0; i < count; ++i)
        KFACB_REQUIRE(n[i] <= 8192, "matrix %d: n = %d > 8192 is not supported by the in-kernel eigenvalue sort", i, n[i]);
    GroupRuntime* rt = nullptr;
    const bool single = (pl.ngroups <= 1);       // one group runs directly on the caller's stream
    if (pl.n_block > 0) {
        rt = group_runtime();
        if (!rt) { set_error("cannot create the eigh streams / events / progress flags"); return KFACB_ERR_CUDA; }
        if (!single) {                           // fork now: the groups need not wait for the small-matrix kernel below
            KFACB_CHECK_CUDA(cudaEventRecord(rt->fork, st));
            for (int k = 0; k < pl.ngroups; ++k) KFACB_CHECK_CUDA(cudaStreamWaitEvent(rt->stream[k], rt->fork, 0));
        }
    }

    // ---- direct path ----
    if (pl.n_direct > 0) {
        DirectMat* hm = new DirectMat[pl.n_direct];
        int* hidx = new int[pl.n_direct];
        int k = 0, nmax = 0;
        for (int i = 0; i < count; ++i) {
            if (n[i] <= 0 || n[i] > DIRECT_MAX_N) continue;
            if (!M[i] || !Q[i] || !d[i]) { delete[] hm; delete[] hidx; set_error("matrix %d: null pointer", i); return KFACB_ERR_ARG; }
            hm[k] = DirectMat{M[i], Q[i], d[i], n[i]};
            hidx[k] = i;
            if (n[i] > nmax) nmax = n[i];
            ++k;
        }
        cudaError_t e1 = cudaMemcpyAsync(ws + pl.off_dmats, hm, sizeof(DirectMat) * k, cudaMemcpyHostToDevice, st);
        cudaError_t e2 = cudaMemcpyAsync(ws + pl.off_dindex, hidx, sizeof(int) * k, cudaMemcpyHostToDevice, st);
        delete[] hm; delete[] hidx;
        KFACB_CHECK_CUDA(e1); KFACB_CHECK_CUDA(e2);
        const int m = (nmax + 1) & ~1;
        const size_t smem = 2 * (size_t)m * (m + 1) * sizeof(float);
        KFACB_CHECK_CUDA(cudaFuncSetAttribute(jacobi_direct_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        const int threads = nmax <= 32 ? 128 : (nmax <= 64 ? 256 : 512);
        jacobi_direct_kernel<<<k, threads, smem, st>>>((const DirectMat*)(ws + pl.off_dmats), eps_floor,
                                                       sweeps_out, (const int*)(ws + pl.off_dindex));
        KFACB_CHECK_LAUNCH();
    }
    if (pl.n_block == 0) return KFACB_OK;

    // ---- blocked path ----
    for (int i = 0; i < count; ++i)
        if (n[i] > DIRECT_MAX_N) KFACB_REQUIRE(M[i] && Q[i] && d[i], "matrix %d: null pointer", i);
    const bool prof = env_int("KFACB_PROFILE", 0) != 0;
    const int full_inner = env_int("KFACB_FULL_INNER", 0);                     // 1 = full inner tournament every visit
    const int max_sweeps = env_int("KFACB_MAX_SWEEPS", BJ_MAX_SWEEPS);
    KFACB_CHECK_CUDA(cudaFuncSetAttribute(bj_h_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, tcg::SMEM));
    KFACB_CHECK_CUDA(cudaFuncSetAttribute(bj_update_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, UPD_SMEM));
    KFACB_CHECK_CUDA(cudaFuncSetAttribute(bj_solve_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, SolveCfg<false>::SMEM));
    KFACB_CHECK_CUDA(cudaFuncSetAttribute(bj_solve_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, SolveCfg<true>::SMEM));

    const int G = pl.ngroups;
    GroupRun run[MAX_GROUPS];
    struct RunGuard { GroupRun* r; int g; ~RunGuard() { for (int k = 0; k < g; ++k) delete[] r[k].hm; } } run_guard{run, G};
    for (int k = 0; k < G; ++k) {
        const GroupPlan& gp = pl.g[k];
        GroupRun& gr = run[k];
        gr.gp = &gp;
        gr.st = single ? st : rt->stream[k];
        gr.hm = new BjMat[gp.nb];
        int* hidx = new int[gp.nb];
        long* hpoff = new long[gp.nb];
        int* hpm = new int[gp.total_pairs];
        size_t xo = gp.off_x, lo = gp.off_lam, po = gp.off_perm;
        long partial_off = 0;
        int pair = 0;
        for (int a = 0; a < gp.nb; ++a) {
            const int i = gp.idx[a];
            BjMat& b = gr.hm[a];
            b.n = n[i]; b.n_pad = bj_npad(n[i]); b.nblk = b.n_pad / BJ_B; b.p = (b.nblk + 1) & ~1;
            b.A = M[i]; b.Q = Q[i]; b.d = d[i];
            b.X = (float*)(ws + xo); xo += align_up((size_t)b.n_pad * 2 * b.n_pad * sizeof(float), 256);
            b.lam = (float*)(ws + lo); b.nrm = b.lam + b.n_pad; lo += 2 * (size_t)b.n_pad * sizeof(float);
            b.perm = (int*)(ws + po); po += (size_t)b.n_pad * sizeof(int);
            b.pair_begin = pair; b.kchunk = gp.kchunk; b.ksplit = cdiv(b.n_pad, gp.kchunk);
            // few blocks = short sweeps and little quadratic tail: ask for more before stopping (these are cheap)
            b.tol_conv = 1e-6f * (float)env_int("KFACB_CONV_TOL_E6", b.p >= 16 ? 100 : 20);
            hpoff[a] = partial_off;
            partial_off += (long)(b.p / 2) * b.ksplit * BJ_M * BJ_M;
            for (int q = 0; q < b.p / 2; ++q) hpm[pair++] = a;
            hidx[a] = i;
        }
        cudaError_t e1 = cudaMemcpyAsync(ws + gp.off_bmats, gr.hm, sizeof(BjMat) * gp.nb, cudaMemcpyHostToDevice, gr.st);
        cudaError_t e2 = cudaMemcpyAsync(ws + gp.off_bindex, hidx, sizeof(int) * gp.nb, cudaMemcpyHostToDevice, gr.st);
        cudaError_t e3 = cudaMemcpyAsync(ws + gp.off_partoff, hpoff, sizeof(long) * gp.nb, cudaMemcpyHostToDevice, gr.st);
        cudaError_t e4 = cudaMemcpyAsync(ws + gp.off_pairmat, hpm, sizeof(int) * gp.total_pairs, cudaMemcpyHostToDevice, gr.st);
        cudaError_t e5 = cudaMemsetAsync(ws + gp.off_state, 0, sizeof(BjState) * gp.nb, gr.st);
        cudaError_t e6 = cudaMemsetAsync(ws + gp.off_rotflag, 0, sizeof(int) * (gp.total_pairs + 1), gr.st);
        delete[] hidx; delete[] hpoff; delete[] hpm;
        KFACB_CHECK_CUDA(e1); KFACB_CHECK_CUDA(e2); KFACB_CHECK_CUDA(e3); KFACB_CHECK_CUDA(e4); KFACB_CHECK_CUDA(e5);
        KFACB_CHECK_CUDA(e6);
        gr.rotlog = (float4*)(ws + gp.off_rotlog);
        gr.rotflag = (int*)(ws + gp.off_rotflag);
        gr.dm = (const BjMat*)(ws + gp.off_bmats);
        gr.dpm = (const int*)(ws + gp.off_pairmat);
        gr.dpoff = (const long*)(ws + gp.off_partoff);
        gr.dstate = (BjState*)(ws + gp.off_state);
        gr.ubuf = (float*)(ws + gp.off_ubuf);
        gr.partial = (float*)(ws + gp.off_partial);
        bj_arm_kernel<<<1, 1, 0, gr.st>>>(rt->d_active + k);
        bj_init_kernel<<<dim3(4 * sm_count(), gp.nb), 256, 0, gr.st>>>(gr.dm, gr.dstate);
        KFACB_CHECK_LAUNCH();
    }

    // ---- round loop: the groups' rounds are enqueued interleaved, each on its own stream ----
    constexpr int CH = 24;            // rounds per progress chunk (host look-ahead = 2 chunks)
    int rc = KFACB_OK;
    cudaEvent_t pev[4];
    double ptime[3] = {0, 0, 0}, pbytes[2] = {0, 0};
    int prounds = 0;
    BjState* hstate = prof ? new BjState[pl.g[0].nb] : nullptr;
    if (prof) for (int k = 0; k < 4; ++k) cudaEventCreate(&pev[k]);
    int max_rounds = 0;
    for (int k = 0; k < G; ++k) if (pl.g[k].max_rounds > max_rounds) max_rounds = pl.g[k].max_rounds;
    for (int r = 0; r < max_rounds && rc == KFACB_OK; ++r) {
        bool any_active = false;
        for (int k = 0; k < G; ++k) {
            GroupRun& gr = run[k];
            const GroupPlan& gp = *gr.gp;
            if (gr.done || r >= gp.max_rounds) continue;
            if (r % CH == 0 && r >= 2 * CH) {
                if (cudaEventSynchronize(rt->prog[k][(r / CH) & 1]) != cudaSuccess) { rc = KFACB_ERR_CUDA; set_error("event sync failed"); break; }
                if (((volatile int*)rt->h_active)[k] == 0) { gr.done = true; continue; }
            }
            any_active = true;
            const int col_tiles = cdiv(gp.max_ld, TILE);
            if (prof) cudaEventRecord(pev[0], gr.st);
            bj_h_tc_kernel<<<dim3(gp.max_ksplit, gp.total_pairs), tcg::THREADS, tcg::SMEM, gr.st>>>(gr.dm, gr.dpm, gr.dstate, r, gr.partial, gr.dpoff);
            if (prof) cudaEventRecord(pev[1], gr.st);
            KFACB_CHECK_CUDA(launch_solve(gr, gp, r, full_inner, gr.st));
            if (prof) cudaEventRecord(pev[2], gr.st);
            bj_update_tc_kernel<<<dim3(cdiv(col_tiles, gp.upd_tiles), gp.total_pairs), UPD_THREADS, UPD_SMEM, gr.st>>>(gr.dm, gr.dpm, gr.dstate, r, gr.ubuf, gp.upd_tiles);
            if (prof) {          // KFACB_PROFILE=1: per-kernel time of every round (debug only: synchronises each round)
                cudaEventRecord(pev[3], gr.st);
                cudaEventSynchronize(pev[3]);
                float t;
                for (int e = 0; e < 3; ++e) { cudaEventElapsedTime(&t, pev[e], pev[e + 1]); ptime[e] += t; }
                ++prounds;
                // algorithmic HBM bytes of this round: every ACTIVE matrix has its W and V rows read once by the
                // projected-block kernel (8 n_pad^2) and read + written once by the update kernel (16 n_pad^2)
                if (cudaMemcpy(hstate, gr.dstate, sizeof(BjState) * gp.nb, cudaMemcpyDeviceToHost) == cudaSuccess)
                    for (int a = 0; a < gp.nb; ++a)
                        if (!hstate[a].converged) {
                            const double n2 = (double)gr.hm[a].n_pad * gr.hm[a].n_pad;
                            pbytes[0] += 8.0 * n2; pbytes[1] += 16.0 * n2;
                        }
            }
            bool boundary = false;
            for (int a = 0; a < gp.nb && !boundary; ++a) boundary = ((r + 1) % (gr.hm[a].p - 1)) == 0;
            if (boundary) {
                bj_check_kernel<<<1, 256, 0, gr.st>>>(gr.dm, gr.dstate, gp.nb, r, max_sweeps, rt->d_active + k, 0);
                count_launch(1);
            }
            count_launch(3);
            if (cudaGetLastError() != cudaSuccess) { rc = KFACB_ERR_CUDA; set_error("jacobi round launch failed"); break; }
            if (r % CH == CH - 1) cudaEventRecord(rt->prog[k][(r / CH) & 1], gr.st);
        }
        if (!any_active) break;
    }
    if (prof) {
        fprintf(stderr, "[kfacb eigh profile] matrices=%d pairs=%d rounds=%d  H %.1f ms  solve %.1f ms  update %.1f ms"
                "  (kchunk %d, tiles/CTA %d, split solve %d)\n",
                pl.g[0].nb, pl.g[0].total_pairs, prounds, ptime[0], ptime[1], ptime[2], pl.g[0].kchunk, pl.g[0].upd_tiles, pl.g[0].solve_split);
        for (int k = 0; k < 4; ++k) cudaEventDestroy(pev[k]);
        g_profile[0] = prounds; g_profile[1] = ptime[0]; g_profile[2] = ptime[1]; g_profile[3] = ptime[2];
        g_profile[4] = pbytes[0]; g_profile[5] = pbytes[1]; g_profile[6] = pl.g[0].nb; g_profile[7] = pl.g[0].total_pairs;
        delete[] hstate;
    }

    // ---- per group: re-orthogonalise, Rayleigh quotients, sort, write; then join the caller's stream ----
    for (int k = 0; k < G && rc == KFACB_OK; ++k) {
        GroupRun& gr = run[k];
        const GroupPlan& gp = *gr.gp;
        cudaStream_t gs = gr.st;
        // Clean-up.  ~1e3 rounds of 3xTF32 row updates leave V slightly non-orthogonal and W = A V slightly
        // inconsistent (the tensor core accumulates round-toward-zero: a biased drift of ~4e-4 in the row norms),
        // which would cap the accuracy at ~1e-4.  So: (a) one Newton-Schulz step re-orthogonalises V,
        // (b) W is recomputed from scratch as A V, (c) ONE more Jacobi sweep on the now consistent pair removes
        // what is left (quadratically), with only one sweep's worth of fresh rounding.
        if (env_int("KFACB_NS", 1)) {
            for (int a = 0; a < gp.nb; ++a) {
                const BjMat& b = gr.hm[a];
                const long ld = 2L * b.n_pad;
                LdKContig lv{b.X + b.n_pad, ld, b.n, b.n};
                KFACB_CHECK_CUDA(launch_gemm(lv, lv, EpiGramMinusI{b.X, ld}, b.n, b.n, b.n, gs));   // V V^T - I cancels to ~1e-6: exact fp32 FMAs, not 3xTF32
                LdKContig le{b.X, ld, b.n, b.n};
                LdMnContig lvt{b.X + b.n_pad, ld, b.n, b.n};
                KFACB_CHECK_CUDA(launch_gemm_auto(le, lvt, EpiNsUpdate{b.Q, b.X, ld, b.n, b.n_pad}, b.n, b.n, b.n, gs));   // E ~ 1e-5: 3xTF32 is ample
                KFACB_CHECK_CUDA(cudaMemcpy2DAsync(b.X + b.n_pad, ld * sizeof(float), b.Q, (size_t)b.n * sizeof(float),
                                                   (size_t)b.n * sizeof(float), b.n, cudaMemcpyDeviceToDevice, gs));
            }
        }
        if (env_int("KFACB_CLEANUP_SWEEP", 1) && !prof) {
            int pmax = 0;
            for (int a = 0; a < gp.nb; ++a) {
                const BjMat& b = gr.hm[a];
                if (b.p > pmax) pmax = b.p;
                LdKContig la{b.X + b.n_pad, 2L * b.n_pad, b.n, b.n};       // W^T rows = (A v_i)^T = v_i^T A
                LdMnContig lb{b.A, (long)b.n, b.n, b.n};
                KFACB_CHECK_CUDA(launch_gemm_auto(la, lb, EpiToX{b.X, 2L * b.n_pad}, b.n, b.n, b.n, gs));
            }
            bj_rearm_kernel<<<cdiv(gp.nb, 128), 128, 0, gs>>>(gr.dstate, gp.nb);
            KFACB_CHECK_LAUNCH();
            const int col_tiles = cdiv(gp.max_ld, TILE);
            for (int r = 0; r < pmax - 1; ++r) {
                bj_h_tc_kernel<<<dim3(gp.max_ksplit, gp.total_pairs), tcg::THREADS, tcg::SMEM, gs>>>(gr.dm, gr.dpm, gr.dstate, r, gr.partial, gr.dpoff);
                KFACB_CHECK_CUDA(launch_solve(gr, gp, r, full_inner, gs));
                bj_update_tc_kernel<<<dim3(cdiv(col_tiles, gp.upd_tiles), gp.total_pairs), UPD_THREADS, UPD_SMEM, gs>>>(gr.dm, gr.dpm, gr.dstate, r, gr.ubuf, gp.upd_tiles);
                bool boundary = false;
                for (int a = 0; a < gp.nb && !boundary; ++a) boundary = ((r + 1) % (gr.hm[a].p - 1)) == 0;
                if (boundary) { bj_check_kernel<<<1, 256, 0, gs>>>(gr.dm, gr.dstate, gp.nb, r, max_sweeps, nullptr, 1); count_launch(1); }
                count_launch(3);
            }
            KFACB_CHECK_LAUNCH();
        }
        // Rayleigh quotients: Y = V^T A  (row i of Y = A v_i), written over the W half
        for (int a = 0; a < gp.nb; ++a) {
            const BjMat& b = gr.hm[a];
            LdKContig la{b.X + b.n_pad, 2L * b.n_pad, b.n, b.n};
            LdMnContig lb{b.A, (long)b.n, b.n, b.n};
            EpiToX epi{b.X, 2L * b.n_pad};
            KFACB_CHECK_CUDA(launch_gemm(la, lb, epi, b.n, b.n, b.n, gs));     // eigenvalues: exact fp32 FMAs (3xTF32 costs ~4e-6)
        }
        bj_rayleigh_kernel<<<dim3(cdiv(gp.nmax, 8), gp.nb), 256, 0, gs>>>(gr.dm);
        KFACB_CHECK_LAUNCH();
        int np2 = 1;
        while (np2 < gp.nmax) np2 <<= 1;
        const size_t sort_smem = (size_t)np2 * 8;
        KFACB_CHECK_CUDA(cudaFuncSetAttribute(bj_sort_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 8 * 8192));
        bj_sort_kernel<<<gp.nb, 1024, sort_smem, gs>>>(gr.dm, eps_floor);
        KFACB_CHECK_LAUNCH();
        bj_write_kernel<<<dim3(cdiv(gp.nmax, 32), cdiv(gp.nmax, 32), gp.nb), 256, 0, gs>>>(gr.dm);
        KFACB_CHECK_LAUNCH();
        if (sweeps_out) {
            bj_store_sweeps_kernel<<<cdiv(gp.nb, 128), 128, 0, gs>>>(gr.dstate, (const int*)(ws + gp.off_bindex), gp.nb, sweeps_out);
            KFACB_CHECK_LAUNCH();
        }
    }
    if (!single) {                               // always join, also on the error path
        for (int k = 0; k < G; ++k) {
            cudaEventRecord(rt->join[k], rt->stream[k]);
            cudaStreamWaitEvent(st, rt->join[k], 0);
        }
    }
    return rc;
}
