// eigh_trd.cu -- Householder tridiagonalisation A = Q_H T Q_H^T of a BATCH of symmetric fp32 matrices in ONE
// persistent kernel (first stage of the n > 64 path of kfacb_eigh_batched; replaces the LAPACK ssytrd inside
// torch.linalg.eigh of the reference, kfac/utils.py:22-25 / kfac_preconditioner_eigen_dp.py:62-75).
//
// Why this shape on B200.  The reduction is a chain of n dependent columns; each column needs one symmetric
// matrix-vector product with the trailing matrix (BLAS-2: 2 m^2 bytes of its lower triangle at column j,
// m = n - j) -- HBM/L2 bandwidth and LATENCY bound, nothing for the tensor cores.  So:
//   * the batch is scheduled on the SMs as TEAMS of CTAs (1 CTA / SM, cooperative launch): a big factor gets a
//     team sized by a cost model fitted to measurements (job_time below), small factors are packed several to a
//     single-CTA team; all matrices advance concurrently and each chain only pays its own latency;
//   * inside a team, rows are dealt in 8-row blocks (block b -> CTA b mod team).  The symmetric product reads the
//     LOWER triangle only: a work unit is (owned 8-row block, 128-column chunk) = one 4 KB box that a warp fetches
//     with ONE TMA request (cp.async.bulk.tensor.2d through the tensor map of the work array) into its private
//     2-slot shared-memory ring, completion on an mbarrier; the warp then forms the 8 row sums (9 shuffles) and the
//     128 transposed sums of the box; units are numbered chunk-major and dealt to the 16 warps in equal ranges;
//   * a column costs exactly TWO team barriers (release / acquire on a counter in L2): column + |x|^2 + the raw
//     panel inner products | barrier | reflector + product | barrier.  Everything else that LAPACK's slatrd does
//     with extra synchronisation (w = p - ..v, the panel corrections V(W^T v) + W(V^T v), v^T p) is folded
//     algebraically so that it needs only the sums made in the same pass; the team-wide partial sums go through
//     per-CTA slots in global memory and are added in a fixed order (no float atomics in global memory; inside a
//     CTA the product accumulates with shared-memory float atomics: runs agree to rounding, not bitwise);
//   * updates are delayed LAPACK-style: within a panel of 32 columns the trailing matrix is read only, the
//     rank-64 update A -= V W^T + W V^T runs once per panel on the lower triangle of 128x128 tiles as an
//     exact-fp32 register-tiled SIMT GEMM (K = 64: too skinny to be worth a TMEM round trip; 2/3 n^3 flops);
//   * reflector j is left in the dead row j of the work array (T[j][j+1..n)), K-contiguous, which is exactly the
//     operand layout the tensor-core back-transformation (eigh_dc.cu) wants.
// The formulas are pinned by tools/eigh_trd_dc_model.py (same phases in numpy).
#include <math.h>
#include <cuda.h>
#include <stdlib.h>
#include <string.h>
#include <type_traits>
#include "eigh_trd.cuh"

namespace kfacb {
namespace trd {

constexpr int RBLK8 = 8;                                         // rows per block (see RBLK below)
constexpr int OP_PITCH = 132;                                   // 128 + 4: float4-aligned, spreads banks
constexpr int NSLOT = 2;                                        // product ring: slots per warp
constexpr int SLOT_BYTES = RBLK8 * 512;                          // one unit = 8 rows x 128 columns
constexpr int RING_BYTES = NWARP * NSLOT * SLOT_BYTES;          // 128 KB

// Fixed part of the dynamic shared memory; followed by the per-job vectors  v[ldt] | z[ldt] | ysum[ycap]:
//   v     reflector of the current column (absolute index)
//   z     this CTA's transposed contributions  z_c = sum_{own rows i > c} T[i][c] v_i
//   ysum  row sums of the CTA's own blocks: [8 k + u] = row u of its k-th owned block
struct __align__(128) Smem {
    unsigned char ring[RING_BYTES];                             // operands of the symmetric product, filled by TMA bulk copies
                                                                // (per warp NSLOT x 4 KB); the rank-2k operands alias it
    float red[3][NWARP][33];
    float qraw[72];                                             // team sums of the raw (unscaled) panel inner products, [64] = |x|^2
    float q[72];                                                // panel inner products of the current reflector: [0,32) W^T v, [32,64) V^T v
    float rowV[32], rowW[32];
    float cw, tau, scale;
    unsigned long long mbar[NWARP][NSLOT];                      // completion barriers of the ring slots
};
static_assert(sizeof(float) * 4 * NB * OP_PITCH <= RING_BYTES, "the rank-2k operands alias the product ring");
constexpr int SMEM_LIMIT = 232448 - 1024;                       // 227 KB per CTA on sm_100, minus the static shared memory (the job descriptor)

#ifdef TRD_EXPERIMENT_NO_ATOMICS
#define SMEM_ADD(p, v) (*(p) += (v))
#else
#define SMEM_ADD(p, v) atomicAdd((p), (v))
#endif
// ---- TMA bulk copy (cp.async.bulk, SASS UBLKCP) + mbarrier transaction counting ----
__device__ __forceinline__ uint32_t smem_addr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* bar, uint32_t bytes) {
    asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.expect_tx.shared::cta.b64 st, [%0], %1;\n\t}" ::"r"(smem_addr(bar)), "r"(bytes) : "memory");
}
// one request = one 8-row x 128-column box of T (4 KB) -> row-major in the slot; out-of-range rows / columns arrive as zeros
__device__ __forceinline__ void tma_box_g2s(void* dst, const void* tmap, int col, int row, unsigned long long* bar) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                 ::"r"(smem_addr(dst)), "l"(tmap), "r"(col), "r"(row), "r"(smem_addr(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait_parity(unsigned long long* bar, uint32_t parity) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tWAIT_%=:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t@p bra DONE_%=;\n\tbra WAIT_%=;\n\tDONE_%=:\n\t}\n"
        ::"r"(smem_addr(bar)), "r"(parity) : "memory");
}
// generic-proxy writes to global memory (rank-2k update) before later async-proxy (TMA) reads of the same locations
__device__ __forceinline__ void fence_proxy_async_global() { asm volatile("fence.proxy.async.global;" ::: "memory"); }

// Team barrier: one release-arrival + acquire-poll on a counter in L2 by thread 0, bracketed by CTA barriers (the CTA's
// writes are ordered before the arrival through bar.sync cumulativity; every global load of the kernel is .cg, so
// no stale L1 line can be read after the acquire).  `tma_fence`: the barriers around the rank-2k update, the only
// place where generic-proxy writes (the updated trailing matrix) are later read through the async proxy (TMA).
__device__ __forceinline__ void team_barrier(unsigned* ctr, unsigned& target, int team, bool tma_fence = false) {
    if (tma_fence) fence_proxy_async_global();
    __syncthreads();
    if (team > 1) {
        target += (unsigned)team;
        if (threadIdx.x == 0) {
            asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(ctr) : "memory");
            unsigned v;
            do {
                asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(ctr) : "memory");
            } while (v < target);
        }
        __syncthreads();
    }
}

// Rows are dealt to the team in BLOCKS of 8 consecutive rows (block b -> CTA b mod team): the 8 rows of a block have
// the same length in the lower triangle (one warp streams them together, 8 loads in flight per lane) and their
// slice of any per-row vector is one 32-byte sector.  First block owned by `rank` that contains a row >= r0:
constexpr int RBLK = RBLK8;
__device__ __forceinline__ int first_block(int r0, int rank, int team) {
    const int b0 = r0 >> 3;
    int o = rank - b0 % team;
    if (o < 0) o += team;
    return b0 + o;
}

// Sum of 8 per-lane values over the 32 lanes of a warp, all 8 at once: each of the exchanges over lane bits 4, 3, 2
// halves the number of rows a lane carries (it sends the half it gives up), two more butterflies finish the one row
// left.  9 shuffles instead of 40.  On return the lanes with (lane & 3) == 0 hold the total of row (lane >> 2).
__device__ __forceinline__ float row_reduce8(const float (&a)[8]) {
    const int lane = threadIdx.x & 31;
    const bool h16 = lane & 16, h8 = lane & 8, h4 = lane & 4;
    float b[4], c[2];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
        const float send = h16 ? a[u] : a[u + 4], keep = h16 ? a[u + 4] : a[u];
        b[u] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
    }
#pragma unroll
    for (int u = 0; u < 2; ++u) {
        const float send = h8 ? b[u] : b[u + 2], keep = h8 ? b[u + 2] : b[u];
        c[u] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
    }
    const float send = h4 ? c[0] : c[1], keep = h4 ? c[1] : c[0];
    float r = keep + __shfl_xor_sync(0xffffffffu, send, 4);
    r += __shfl_xor_sync(0xffffffffu, r, 2);
    r += __shfl_xor_sync(0xffffffffu, r, 1);
    return r;
}

// Start of a column (after the product's barrier): cw = tau^2 (v.y - 2 s.t) / 2 from the team-summed v.y and the
// panel inner products s = W^T v (sm.q[0..)), t = V^T v (sm.q[32..)) that B_setup already reduced; with want_row,
// warp 0 also finishes W[row_j][kc] redundantly (every CTA needs row j of the panels) and stages rows j of V and W.
// The loads are issued by `prologue_issue` BEFORE the caller issues its own row loads, so both round trips overlap.
struct ProRegs { float vy, vj, wj, zz, yj; };
__device__ __forceinline__ void prologue_issue(const Job& jb, int kc, int row_j, bool want_row, ProRegs& r) {
    const int lane = threadIdx.x & 31;
    r.vy = 0.f; r.vj = 0.f; r.wj = 0.f; r.zz = 0.f; r.yj = 0.f;
#pragma unroll
    for (int it = 0; it < TS / 32; ++it) {
        const int q = lane + 32 * it;
        if (q < jb.team) {
            r.vy += __ldcg(jb.part + 64 * TS + q);
            if (want_row) r.zz += __ldcg(jb.zpart + (size_t)q * jb.ldt + row_j);
        }
    }
    if (want_row) {
        r.vj = lane <= kc ? __ldcg(jb.Vp + (size_t)row_j * NB + lane) : 0.f;
        r.wj = lane < kc ? __ldcg(jb.Wp + (size_t)row_j * NB + lane) : 0.f;
        r.yj = __ldcg(jb.y + row_j);
    }
}
// warp 0 only; the caller synchronises the CTA afterwards
__device__ __forceinline__ void prologue_finish(Smem& sm, int kc, bool want_row, const ProRegs& r) {
    const int lane = threadIdx.x & 31;
    const float vy = warp_sum(r.vy);
    const float s = sm.q[lane], t = sm.q[32 + lane];
    const float st = warp_sum(lane < kc ? s * t : 0.f);
    const float tau = sm.tau;
    const float cw = 0.5f * tau * tau * (vy - 2.f * st);
    if (lane == 0) sm.cw = cw;
    if (want_row) {
        const float yj = r.yj + warp_sum(r.zz);
        const float dot = warp_sum(lane < kc ? r.vj * s + r.wj * t : 0.f);
        const float vlast = __shfl_sync(0xffffffffu, r.vj, kc);
        const float wnew = tau * (yj - dot) - cw * vlast;
        sm.rowV[lane] = r.vj;
        sm.rowW[lane] = lane == kc ? wnew : r.wj;
    }
}

// A[R0.., R0..] -= V W^T + W V^T (lower triangle of 128x128 tiles) over the first kc panel columns; tiles dealt to the team.
// 16 warps = 4 x 4, lanes = 4 (row groups of 8) x 8 (column groups of 4): 8x4 outputs per thread; per k a warp
// reads 6 shared-memory wavefronts (broadcast within a group) for 64 FMAs -> FMA bound.
__device__ __noinline__ void rank2k_update(const Job& jb, Smem& sm, int R0, int kc, int rank) {
    const int n = jb.n, ldt = jb.ldt;
    const int mt = n - R0;
    if (mt <= 0) return;
    const int nt = (mt + 127) >> 7;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int row8 = 32 * (warp >> 2) + 8 * (lane >> 3);
    const int col4 = 32 * (warp & 3) + 4 * (lane & 7);
    const bool aligned = (R0 & 3) == 0;
    float (*op)[NB][OP_PITCH] = reinterpret_cast<float (*)[NB][OP_PITCH]>(sm.ring);
    // only the LOWER triangle of tiles (diagonal tiles in full): nothing reads the trailing matrix above the diagonal
    // any more (the product reads c <= i, the column of phase A comes from T[i][j]) -- half the update work
    for (int q = rank; q < nt * (nt + 1) / 2; q += jb.team) {
        int ti = (int)((sqrtf(8.f * (float)q + 1.f) - 1.f) * 0.5f);
        while (ti * (ti + 1) / 2 > q) --ti;
        while ((ti + 1) * (ti + 2) / 2 <= q) ++ti;
        const int tj = q - ti * (ti + 1) / 2;
        __syncthreads();
#pragma unroll
        for (int it = 0; it < (128 * 8) / THREADS; ++it) {
            const int item = tid + it * THREADS;
            const int r = item >> 3, kq = item & 7;
            const int gi = R0 + 128 * ti + r, gj = R0 + 128 * tj + r;
            float4 vi = make_float4(0.f, 0.f, 0.f, 0.f), wi = vi, vj = vi, wj = vi;
            if (gi < n) {
                vi = __ldcg(reinterpret_cast<const float4*>(jb.Vp + (size_t)gi * NB) + kq);
                wi = __ldcg(reinterpret_cast<const float4*>(jb.Wp + (size_t)gi * NB) + kq);
            }
            if (gj < n) {
                vj = __ldcg(reinterpret_cast<const float4*>(jb.Vp + (size_t)gj * NB) + kq);
                wj = __ldcg(reinterpret_cast<const float4*>(jb.Wp + (size_t)gj * NB) + kq);
            }
            const float a[4][4] = {{vi.x, vi.y, vi.z, vi.w}, {wi.x, wi.y, wi.z, wi.w},
                                   {wj.x, wj.y, wj.z, wj.w}, {vj.x, vj.y, vj.z, vj.w}};
#pragma unroll
            for (int o = 0; o < 4; ++o)
#pragma unroll
                for (int u = 0; u < 4; ++u) op[o][4 * kq + u][r] = (4 * kq + u < kc) ? a[o][u] : 0.f;
        }
        __syncthreads();
        // the old values of the tile are requested BEFORE the k loop: their L2 / HBM round trip hides behind the FMAs
        const int gc = R0 + 128 * tj + col4;
        const bool vec = aligned && gc + 3 < n;
        float4 c[8];
#pragma unroll
        for (int a = 0; a < 8; ++a) {
            const int gi = R0 + 128 * ti + row8 + a;
            c[a] = (vec && gi < n) ? __ldcg(reinterpret_cast<const float4*>(jb.T + (size_t)gi * ldt + gc)) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
        float acc[8][4];
#pragma unroll
        for (int a = 0; a < 8; ++a)
#pragma unroll
            for (int b = 0; b < 4; ++b) acc[a][b] = 0.f;
        for (int k = 0; k < kc; ++k) {
            const float4 av0 = *reinterpret_cast<const float4*>(&op[0][k][row8]);
            const float4 av1 = *reinterpret_cast<const float4*>(&op[0][k][row8 + 4]);
            const float4 aw0 = *reinterpret_cast<const float4*>(&op[1][k][row8]);
            const float4 aw1 = *reinterpret_cast<const float4*>(&op[1][k][row8 + 4]);
            const float4 bw = *reinterpret_cast<const float4*>(&op[2][k][col4]);
            const float4 bv = *reinterpret_cast<const float4*>(&op[3][k][col4]);
            const float va[8] = {av0.x, av0.y, av0.z, av0.w, av1.x, av1.y, av1.z, av1.w};
            const float wa[8] = {aw0.x, aw0.y, aw0.z, aw0.w, aw1.x, aw1.y, aw1.z, aw1.w};
            const float wb[4] = {bw.x, bw.y, bw.z, bw.w}, vb[4] = {bv.x, bv.y, bv.z, bv.w};
#pragma unroll
            for (int a = 0; a < 8; ++a)
#pragma unroll
                for (int b = 0; b < 4; ++b) acc[a][b] = fmaf(va[a], wb[b], fmaf(wa[a], vb[b], acc[a][b]));
        }
        if (vec) {
#pragma unroll
            for (int a = 0; a < 8; ++a) {
                const int gi = R0 + 128 * ti + row8 + a;
                if (gi < n) {
                    c[a].x -= acc[a][0]; c[a].y -= acc[a][1]; c[a].z -= acc[a][2]; c[a].w -= acc[a][3];
                    *reinterpret_cast<float4*>(jb.T + (size_t)gi * ldt + gc) = c[a];
                }
            }
        } else {
#pragma unroll
            for (int a = 0; a < 8; ++a) {
                const int gi = R0 + 128 * ti + row8 + a;
                if (gi >= n) continue;
                float* p = jb.T + (size_t)gi * ldt + gc;
#pragma unroll
                for (int b = 0; b < 4; ++b)
                    if (gc + b < n) p[b] = __ldcg(p + b) - acc[a][b];
            }
        }
    }
}

// y_i of the 8 rows of block `blk` (rows 8 blk .. 8 blk + 7): the own lower-triangle part jb.y plus the team's
// transposed contributions zpart[r][i].  Lane l sums CTAs r = l, l + 32, ..; the 8 reductions are interleaved.
__device__ __forceinline__ void assemble_y(const Job& jb, int blk, float (&y8)[RBLK]) {
    const int lane = threadIdx.x & 31;
    float4 a0 = make_float4(0.f, 0.f, 0.f, 0.f), a1 = a0;
    const int i0 = blk * RBLK;
#pragma unroll
    for (int it = 0; it < TS / 32; ++it) {
        const int r = lane + 32 * it;
        if (r < jb.team) {
            const float4* z = reinterpret_cast<const float4*>(jb.zpart + (size_t)r * jb.ldt + i0);
            const float4 z0 = __ldcg(z), z1 = __ldcg(z + 1);
            a0.x += z0.x; a0.y += z0.y; a0.z += z0.z; a0.w += z0.w;
            a1.x += z1.x; a1.y += z1.y; a1.z += z1.z; a1.w += z1.w;
        }
    }
    float s[RBLK] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
#pragma unroll
    for (int o = 16; o > 0; o >>= 1)
#pragma unroll
        for (int u = 0; u < RBLK; ++u) s[u] += __shfl_xor_sync(0xffffffffu, s[u], o);
    const float4* yl = reinterpret_cast<const float4*>(jb.y + i0);
    const float4 y0 = __ldcg(yl), y1 = __ldcg(yl + 1);
    const float yo[RBLK] = {y0.x, y0.y, y0.z, y0.w, y1.x, y1.y, y1.z, y1.w};
#pragma unroll
    for (int u = 0; u < RBLK; ++u) y8[u] = yo[u] + s[u];
}

// Per-lane partial of y_i for the 8 rows of block `blk` (no cross-lane reduction): lane l carries the transposed
// contributions of CTAs r = l, l + 32, ..; lane 0 also carries the own lower-triangle part jb.y.
__device__ __forceinline__ void partial_y(const Job& jb, int blk, float (&s)[RBLK]) {
    const int lane = threadIdx.x & 31;
    float4 a0 = make_float4(0.f, 0.f, 0.f, 0.f), a1 = a0;
    const int i0 = blk * RBLK;
#pragma unroll
    for (int it = 0; it < TS / 32; ++it) {
        const int r = lane + 32 * it;
        if (r < jb.team) {
            const float4* z = reinterpret_cast<const float4*>(jb.zpart + (size_t)r * jb.ldt + i0);
            const float4 z0 = __ldcg(z), z1 = __ldcg(z + 1);
            a0.x += z0.x; a0.y += z0.y; a0.z += z0.z; a0.w += z0.w;
            a1.x += z1.x; a1.y += z1.y; a1.z += z1.z; a1.w += z1.w;
        }
    }
    if (lane == 0) {
        const float4* yl = reinterpret_cast<const float4*>(jb.y + i0);
        const float4 y0 = __ldcg(yl), y1 = __ldcg(yl + 1);
        a0.x += y0.x; a0.y += y0.y; a0.z += y0.z; a0.w += y0.w;
        a1.x += y1.x; a1.y += y1.y; a1.z += y1.z; a1.w += y1.w;
    }
    s[0] = a0.x; s[1] = a0.y; s[2] = a0.z; s[3] = a0.w; s[4] = a1.x; s[5] = a1.y; s[6] = a1.z; s[7] = a1.w;
}

__device__ void run_job(const Job& jb, int rank, Smem& sm, float* dynf, uint32_t& ring_par) {
    const int n = jb.n, ldt = jb.ldt, team = jb.team;
    float* const smv = dynf;
    float* const smz = dynf + ldt;
    float* const smy = dynf + 2 * ldt;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    unsigned target = 0;
    int j0 = 0;
    if (tid == 0) {
        sm.tau = 0.f;
        asm volatile("prefetch.tensormap [%0];" ::"l"(jb.tmap) : "memory");   // descriptor into the TMA unit's cache
    }
    __syncthreads();
    // optional phase profile (KFACB_TRD_PROF=1): clock64 deltas of thread 0 of the team's first CTA
    const bool prof_on = jb.prof != nullptr && rank == 0 && tid == 0;
    long long prof_acc[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
    long long prof_wait = 0, prof_units = 0;                     // warp 0: cycles spent waiting for TMA data / units consumed
    long long t_last = prof_on ? clock64() : 0;
    if (prof_on) { unsigned long long g; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(g)); jb.prof[9] = (long long)g; }
#define TRD_TICK(slot) do { if (prof_on) { const long long now_ = clock64(); prof_acc[slot] += now_ - t_last; t_last = now_; } } while (0)
    for (int j = 0; j < n - 1; ++j) {
        const int jp = j - j0;
        // ---------------- phase A: finish w_{j-1} (rows >= j), form column j, |x|^2 and the raw panel inner products ----
        // s = W^T v and t = V^T v of THIS column are linear in v = (x with x_{j+1} -> 1) * scale, and the panel rows
        // are in registers here anyway: the sums over rows >= j+2 of W[i][k] x_i, V[i][k] x_i ride on the |x|^2 barrier
        // (B_setup applies `scale` and adds row j+1), so the product phase needs no second pass over the panels.
        {
            ProRegs pr;
            if (jp > 0 && warp == 0) prologue_issue(jb, jp - 1, j, true, pr);
            const int blk0 = first_block(j, rank, team) + team * warp;
            float vk[RBLK], wk[RBLK], tji[RBLK], yi[RBLK];
            auto load_block = [&](int blk) {
#pragma unroll
                for (int u = 0; u < RBLK; ++u) {
                    const int i = blk * RBLK + u;
                    const bool ok = i >= j && i < n;
                    vk[u] = (ok && lane < jp) ? __ldcg(jb.Vp + (size_t)i * NB + lane) : 0.f;
                    wk[u] = (ok && lane < jp - 1) ? __ldcg(jb.Wp + (size_t)i * NB + lane) : 0.f;
                    tji[u] = ok ? __ldcg(jb.T + (size_t)i * ldt + j) : 0.f;      // column j from the lower triangle
                }
                if (jp > 0) partial_y(jb, blk, yi);                  // per-lane partials; reduced together with the panel dot
            };
            if (blk0 * RBLK < n) load_block(blk0);
            if (jp > 0) {
                if (warp == 0) prologue_finish(sm, jp - 1, true, pr);
                __syncthreads();
            }
            TRD_TICK(0);
            const float tau_p = sm.tau, cw = sm.cw;
            float nrm = 0.f, sraw = 0.f, traw = 0.f;
            for (int blk = blk0; blk * RBLK < n; blk += team * NWARP) {
                if (blk != blk0) load_block(blk);
                if (jp > 0) {
                    float e[RBLK];                                     // per-lane partial of y_i - (V s + W t)_i
#pragma unroll
                    for (int u = 0; u < RBLK; ++u) e[u] = yi[u] - (lane < jp - 1 ? vk[u] * sm.q[lane] + wk[u] * sm.q[32 + lane] : 0.f);
                    const float r = row_reduce8(e);                   // lane 4 u: row u
#pragma unroll
                    for (int u = 0; u < RBLK; ++u) {
                        const int i = blk * RBLK + u;
                        const float ydu = __shfl_sync(0xffffffffu, r, 4 * u);
                        const float vlast = __shfl_sync(0xffffffffu, vk[u], jp - 1);
                        const float wnew = tau_p * ydu - cw * vlast;
                        if (lane == jp - 1 && i >= j && i < n) { wk[u] = wnew; jb.Wp[(size_t)i * NB + lane] = wnew; }
                    }
                }
                float corr[RBLK];
#pragma unroll
                for (int u = 0; u < RBLK; ++u) corr[u] = lane < jp ? vk[u] * sm.rowW[lane] + wk[u] * sm.rowV[lane] : 0.f;
                const float rc = row_reduce8(corr);                   // lane 4 u: row u
#pragma unroll
                for (int u = 0; u < RBLK; ++u) {
                    const int i = blk * RBLK + u;
                    const float xi = tji[u] - __shfl_sync(0xffffffffu, rc, 4 * u);
                    if (i >= j + 2 && i < n) {                          // lane k: column k of the panels
                        sraw = fmaf(wk[u], xi, sraw);
                        traw = fmaf(vk[u], xi, traw);
                    }
                    if (lane == 0 && i >= j && i < n) {
                        if (i == j) jb.d[j] = xi;
                        else { jb.x[i] = xi; if (i >= j + 2) nrm = fmaf(xi, xi, nrm); }
                    }
                }
            }
            sm.red[0][warp][lane] = sraw;
            sm.red[1][warp][lane] = traw;
            if (lane == 0) sm.red[2][warp][0] = nrm;
            __syncthreads();
            if (warp == 0) {
                float ss = 0.f, tt = 0.f, tot = 0.f;
                for (int w = 0; w < NWARP; ++w) { ss += sm.red[0][w][lane]; tt += sm.red[1][w][lane]; tot += sm.red[2][w][0]; }
                if (lane < jp) { jb.part[lane * TS + rank] = ss; jb.part[(32 + lane) * TS + rank] = tt; }
                if (lane == 0) jb.part[65 * TS + rank] = tot;
            }
        }
        TRD_TICK(1);
        team_barrier(jb.bar, target, team);
        TRD_TICK(2);
        // ---------------- phase B: reflector, y = T v (lower triangle only), panel inner products ----------------
        const int c_al = (j + 1) & ~3;
        constexpr int XR = MAXN / THREADS;
        float xr[XR];                                            // raw column, in flight while |x|^2 is being reduced
#pragma unroll
        for (int u = 0; u < XR; ++u) {
            const int c = c_al + tid + u * THREADS;
            xr[u] = (c > j + 1 && c < n) ? __ldcg(jb.x + c) : 0.f;
        }
        if (warp == 0) {
            float s2 = 0.f;
#pragma unroll
            for (int it = 0; it < TS / 32; ++it) {
                const int r = lane + 32 * it;
                if (r < team) s2 += __ldcg(jb.part + 65 * TS + r);
            }
            const float alpha = __ldcg(jb.x + j + 1);
            s2 = warp_sum(s2);
            float beta, tau, scale;
            if (s2 == 0.f) { tau = 0.f; beta = alpha; scale = 0.f; }
            else {
                beta = -copysignf(sqrtf(fmaf(alpha, alpha, s2)), alpha);
                tau = (beta - alpha) / beta;
                scale = 1.f / (alpha - beta);
            }
            if (lane == 0) {
                sm.tau = tau; sm.scale = scale;
                if (rank == 0) { jb.e[j] = beta; jb.tau[j] = tau; }
            }
        }
        __syncthreads();
        {
            const float scale = sm.scale;
#pragma unroll
            for (int u = 0; u < XR; ++u) {
                const int c = c_al + tid + u * THREADS;
                if (c < ldt) { smv[c] = (c == j + 1) ? 1.f : xr[u] * scale; smz[c] = 0.f; }
            }
        }
        __syncthreads();
        TRD_TICK(3);
        {
            // Symmetric product from the LOWER triangle (half the bytes).  Element T[i][c], c <= i, feeds the row sum
            // y_i += T[i][c] v_c and, for c < i, the transposed half z_c += T[i][c] v_i.  Work unit = (owned 8-row
            // block k, chunk of 32 quads = 128 columns): a warp loads the 8 rows of the chunk together (8 x 16 bytes
            // in flight per lane), sums the transposed half over the 8 rows in registers (one shared-memory atomic
            // per column) and reduces the 8 row partials over its lanes into shared-memory row accumulators.
            // Blocks have very different lengths in the lower triangle, so the units are dealt chunk-major and
            // zig-zag (chunk c -> warp c mod 16, reversed every other round): every warp gets long and short blocks.
            const int q_beg = c_al >> 2;
            const int b_first = first_block(j + 1, rank, team);
            const int K = b_first * RBLK < n ? ((n - 1) / RBLK - b_first) / team + 1 : 0;      // owned blocks this column
            for (int e = tid; e < K * RBLK; e += THREADS) smy[e] = 0.f;
            __syncthreads();
            const int q_hi = K > 0 ? 2 * (b_first + (K - 1) * team) + 1 : -1;                   // last quad of the longest block
            const int nchunk = q_hi >= q_beg ? (q_hi - q_beg) / 32 + 1 : 0;
            // Balanced dealing: the units (chunk c, block k >= kmin(c)) are numbered chunk-major and every warp takes a
            // contiguous range of total / NWARP units -- equal work whatever the team size and n (a single-CTA team
            // with n = 512 has 4 chunks but 128 units), and a warp's consecutive units mostly share the chunk, so
            // the transposed half stays in registers and is flushed once per chunk (4 shared-memory atomics / lane).
            // Lane l holds the unit counts of chunks l and l + 32 (nchunk <= MAXN / 128 = 64); warp-scan for the prefix.
            const int qb2 = q_beg >> 1;
            int km0, km1;
            {
                const int c0_ = lane, c1_ = lane + 32;
                int t0 = (qb2 + 16 * c0_ - b_first + team - 1) / team; if (t0 < 0) t0 = 0;
                int t1 = (qb2 + 16 * c1_ - b_first + team - 1) / team; if (t1 < 0) t1 = 0;
                km0 = t0; km1 = t1;
            }
            const int w0 = (lane < nchunk && km0 < K) ? K - km0 : 0;
            const int w1 = (lane + 32 < nchunk && km1 < K) ? K - km1 : 0;
            int p0 = w0, p1 = w1;                                   // inclusive scans
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int a0 = __shfl_up_sync(0xffffffffu, p0, o), a1 = __shfl_up_sync(0xffffffffu, p1, o);
                if (lane >= o) { p0 += a0; p1 += a1; }
            }
            const int tot0 = __shfl_sync(0xffffffffu, p0, 31);
            p1 += tot0;
            const int total = __shfl_sync(0xffffffffu, p1, 31);
            const int u_lo = (int)(((long long)total * warp) / NWARP), u_hi = (int)(((long long)total * (warp + 1)) / NWARP);
            if (u_hi > u_lo) {
                // Unit iterators over the warp's range (warp-uniform): `pi` runs NSLOT units ahead of `ci` and issues
                // the TMA copy of a unit (ONE tensor-map request: the 8 x 128 box of T) into ring slot (unit & 1);
                // the warp then waits on the slot's mbarrier, reads its quads from shared memory and, once the values
                // are consumed, refills the slot.  Bytes in flight are bounded by the ring (8 KB / warp), not registers.
                struct It { int c, k, k_end, u; };
                const unsigned b0m = __ballot_sync(0xffffffffu, p0 > u_lo), b1m = __ballot_sync(0xffffffffu, p1 > u_lo);
                const int c_first = b0m ? __ffs(b0m) - 1 : 32 + __ffs(b1m) - 1;
                auto seek = [&](It& it, int c) {                       // position `it` on its unit it.u inside chunk c
                    const int src = c & 31;
                    const int pin = c < 32 ? __shfl_sync(0xffffffffu, p0, src) : __shfl_sync(0xffffffffu, p1, src);
                    const int wc = c < 32 ? __shfl_sync(0xffffffffu, w0, src) : __shfl_sync(0xffffffffu, w1, src);
                    const int kmc = c < 32 ? __shfl_sync(0xffffffffu, km0, src) : __shfl_sync(0xffffffffu, km1, src);
                    const int pst = pin - wc;
                    const int u_end = pin < u_hi ? pin : u_hi;
                    it.c = c; it.k = kmc + (it.u - pst); it.k_end = kmc + (u_end - pst);
                };
                auto advance = [&](It& it) {
                    ++it.u; ++it.k;
                    while (it.k >= it.k_end && it.u < u_hi) seek(it, it.c + 1);
                };
                unsigned char* const ring = sm.ring + warp * (NSLOT * SLOT_BYTES);
                // rows of a unit that exist and reach the chunk (lane u < 8 <-> row u); the same on both sides
                auto unit_rows = [&](const It& it, int& i0, int& col0) -> unsigned {
                    i0 = (b_first + it.k * team) * RBLK;
                    col0 = 4 * (q_beg + 32 * it.c);
                    const int i = i0 + lane;
                    return __ballot_sync(0xffffffffu, lane < RBLK && i > j && i < n && col0 <= i);
                };
                auto issue = [&](const It& it) {
                    int i0, col0;
                    const unsigned rows = unit_rows(it, i0, col0);
                    if (rows == 0u) return;
                    const int slot = (it.u - u_lo) & (NSLOT - 1);
                    if (lane == 0) {
                        unsigned long long* bar = &sm.mbar[warp][slot];
                        mbar_expect_tx(bar, SLOT_BYTES);            // a tiled box always delivers all its bytes (zeros outside the tensor)
                        tma_box_g2s(ring + slot * SLOT_BYTES, jb.tmap, col0, i0, bar);
                    }
                };
                It pi, ci;
                pi.u = u_lo; seek(pi, c_first);
                while (pi.k >= pi.k_end && pi.u < u_hi) seek(pi, pi.c + 1);
                ci = pi;
#pragma unroll
                for (int s_ = 0; s_ < NSLOT; ++s_)
                    if (pi.u < u_hi) { issue(pi); advance(pi); }
                int cur_c = -1, q = 0;
                float4 vv = make_float4(0.f, 0.f, 0.f, 0.f), zc = vv;
                while (ci.u < u_hi) {
                    if (ci.c != cur_c) {
                        if (cur_c >= 0 && q < (ldt >> 2)) {
                            SMEM_ADD(&smz[4 * q], zc.x); SMEM_ADD(&smz[4 * q + 1], zc.y);
                            SMEM_ADD(&smz[4 * q + 2], zc.z); SMEM_ADD(&smz[4 * q + 3], zc.w);
                        }
                        cur_c = ci.c;
                        q = q_beg + 32 * cur_c + lane;
                        vv = *reinterpret_cast<const float4*>(&smv[4 * (q < (ldt >> 2) ? q : q_beg)]);
                        zc = make_float4(0.f, 0.f, 0.f, 0.f);
                    }
                    int i0, col0;
                    const unsigned rows = unit_rows(ci, i0, col0);
                    const int slot = (ci.u - u_lo) & (NSLOT - 1);
                    const int k = ci.k;
                    if (rows != 0u) {
                        const long long tw0 = prof_on ? clock64() : 0;
                        mbar_wait_parity(&sm.mbar[warp][slot], (ring_par >> slot) & 1u);
                        ring_par ^= 1u << slot;
                        if (prof_on) { prof_wait += clock64() - tw0; ++prof_units; }
                    }
                    const unsigned char* sl = ring + slot * SLOT_BYTES + lane * 16;
                    float acc[RBLK];
                    // The product is instruction-issue bound (4 warps per scheduler), so the common case -- all 8 rows
                    // live and the whole 128-column chunk strictly below the diagonal -- runs without a single
                    // predicate: 10 LDS.128 + 64 FFMA per lane.  Warp-uniform test.
                    if (i0 > j && i0 + RBLK <= n && col0 + 128 <= i0) {
                        const float4 va = *reinterpret_cast<const float4*>(&smv[i0]);
                        const float4 vb = *reinterpret_cast<const float4*>(&smv[i0 + 4]);
                        const float vi8[RBLK] = {va.x, va.y, va.z, va.w, vb.x, vb.y, vb.z, vb.w};
#pragma unroll
                        for (int uu = 0; uu < RBLK; ++uu) {
                            const float4 t = *reinterpret_cast<const float4*>(sl + uu * 512);
                            acc[uu] = fmaf(t.x, vv.x, fmaf(t.y, vv.y, fmaf(t.z, vv.z, t.w * vv.w)));
                            zc.x = fmaf(t.x, vi8[uu], zc.x); zc.y = fmaf(t.y, vi8[uu], zc.y);
                            zc.z = fmaf(t.z, vi8[uu], zc.z); zc.w = fmaf(t.w, vi8[uu], zc.w);
                        }
                    } else {
                        float4 m[RBLK];
#pragma unroll
                        for (int uu = 0; uu < RBLK; ++uu) {
                            const int i = i0 + uu;
                            m[uu] = (((rows >> uu) & 1u) && 4 * q <= i) ? *reinterpret_cast<const float4*>(sl + uu * 512)
                                                                         : make_float4(0.f, 0.f, 0.f, 0.f);
                        }
                        const bool edge = 4 * q + 3 >= i0;            // this quad touches the diagonal of some row of the block
#pragma unroll
                        for (int uu = 0; uu < RBLK; ++uu) {
                            float4 t = m[uu];
                            const int i = i0 + uu;
                            const float vi = (i > j && i < n) ? smv[i] : 0.f;
                            if (edge) {                               // c <= i for the row sum, c < i for the transposed half
                                const int cc = 4 * q;
                                if (cc > i) t.x = 0.f;
                                if (cc + 1 > i) t.y = 0.f;
                                if (cc + 2 > i) t.z = 0.f;
                                if (cc + 3 > i) t.w = 0.f;
                                acc[uu] = fmaf(t.x, vv.x, fmaf(t.y, vv.y, fmaf(t.z, vv.z, t.w * vv.w)));
                                if (cc == i) t.x = 0.f;
                                if (cc + 1 == i) t.y = 0.f;
                                if (cc + 2 == i) t.z = 0.f;
                                if (cc + 3 == i) t.w = 0.f;
                            } else {
                                acc[uu] = fmaf(t.x, vv.x, fmaf(t.y, vv.y, fmaf(t.z, vv.z, t.w * vv.w)));
                            }
                            zc.x = fmaf(t.x, vi, zc.x); zc.y = fmaf(t.y, vi, zc.y);
                            zc.z = fmaf(t.z, vi, zc.z); zc.w = fmaf(t.w, vi, zc.w);
                        }
                    }
                    // 8 row sums over 32 lanes with 9 shuffles: every exchange halves the rows a lane carries
                    const float tot = row_reduce8(acc);               // lanes with (lane & 3) == 0 hold row lane >> 2
                    if ((lane & 3) == 0) SMEM_ADD(&smy[k * RBLK + (lane >> 2)], tot);
                    // every lane's reads of the slot have been consumed by the reductions above: refill it
                    if (pi.u < u_hi) { issue(pi); advance(pi); }
                    advance(ci);
                }
                if (cur_c >= 0 && q < (ldt >> 2)) {
                    SMEM_ADD(&smz[4 * q], zc.x); SMEM_ADD(&smz[4 * q + 1], zc.y);
                    SMEM_ADD(&smz[4 * q + 2], zc.z); SMEM_ADD(&smz[4 * q + 3], zc.w);
                }
            }
            TRD_TICK(4);
            __syncthreads();                                       // row sums and z complete
            // team sums of the raw panel inner products written before the |x|^2 barrier (quantities k < jp and 32 + k):
            // only the NEXT column needs them, so their round trip hides behind the stores below.  One warp per
            // quantity, all loads of a warp issued before its first reduction, fixed order (same bits on every CTA).
            constexpr int PER = 64 / NWARP;
            float qa[PER];
#pragma unroll
            for (int u = 0; u < PER; ++u) {
                const int k = warp + NWARP * u;
                qa[u] = 0.f;
                if ((k & 31) < jp) {
#pragma unroll
                    for (int it = 0; it < TS / 32; ++it) {
                        const int r = lane + 32 * it;
                        if (r < team) qa[u] += __ldcg(jb.part + k * TS + r);
                    }
                }
            }
            float wj1 = 0.f, vj1 = 0.f;                             // row j+1 of the panels: its v is exactly 1
            if (warp == 1) {
                wj1 = lane < jp ? __ldcg(jb.Wp + (size_t)(j + 1) * NB + lane) : 0.f;
                vj1 = lane < jp ? __ldcg(jb.Vp + (size_t)(j + 1) * NB + lane) : 0.f;
            }
            // per owned row: publish the own part of y and the reflector (lane u < 8 takes row u of the block)
            float acc_vy = 0.f;
            for (int k = warp; k < K; k += NWARP) {
                const int i = (b_first + k * team) * RBLK + lane;
                if (lane < RBLK && i > j && i < n) {
                    const float vi = smv[i], yl = smy[k * RBLK + lane];
                    jb.y[i] = yl;                                      // own (lower-triangle) part of y_i
                    acc_vy = fmaf(vi, yl, acc_vy);
                    jb.Vp[(size_t)i * NB + jp] = vi;
                    jb.T[(size_t)j * ldt + i] = vi;
                }
            }
            // hand z to the team (coalesced) and fold its share of v.y:  v.y = sum_i v_i ylow_i + sum_c v_c z_c
            for (int c = c_al + tid; c < ldt; c += THREADS) {
                const float zc = smz[c];
                jb.zpart[(size_t)rank * ldt + c] = zc;
                acc_vy = fmaf(smv[c], zc, acc_vy);
            }
            acc_vy = warp_sum(acc_vy);
            if (lane == 0) sm.red[2][warp][0] = acc_vy;
#pragma unroll
            for (int u = 0; u < PER; ++u) {
                const float t = warp_sum(qa[u]);
                if (lane == 0) sm.qraw[warp + NWARP * u] = t;
            }
            __syncthreads();
            if (tid == 0) {
                float vy = 0.f;
                for (int w = 0; w < NWARP; ++w) vy += sm.red[2][w][0];
                jb.part[64 * TS + rank] = vy;
            }
            if (warp == 1) {                                        // s = scale (W^T x) + W[j+1], t = scale (V^T x) + V[j+1]
                const float scale = sm.scale;
                sm.q[lane] = lane < jp ? fmaf(scale, sm.qraw[lane], wj1) : 0.f;
                sm.q[32 + lane] = lane < jp ? fmaf(scale, sm.qraw[32 + lane], vj1) : 0.f;
            }
        }
        TRD_TICK(5);
        team_barrier(jb.bar, target, team);
        TRD_TICK(6);
        // ---------------- panel end: finish w_j (rows > j), rank-2k update of the trailing square ----------------
        if (jp == NB - 1 || j == n - 2) {
            {
                ProRegs pr;
                if (warp == 0) { prologue_issue(jb, jp, 0, false, pr); prologue_finish(sm, jp, false, pr); }
                __syncthreads();
            }
            const float tau_p = sm.tau, cw = sm.cw;
            for (int blk = first_block(j + 1, rank, team) + team * warp; blk * RBLK < n; blk += team * NWARP) {
                float vk[RBLK], wk[RBLK], yi[RBLK], dot[RBLK];
#pragma unroll
                for (int u = 0; u < RBLK; ++u) {
                    const int i = blk * RBLK + u;
                    const bool ok = i > j && i < n;
                    vk[u] = (ok && lane <= jp) ? __ldcg(jb.Vp + (size_t)i * NB + lane) : 0.f;
                    wk[u] = (ok && lane < jp) ? __ldcg(jb.Wp + (size_t)i * NB + lane) : 0.f;
                }
                assemble_y(jb, blk, yi);
#pragma unroll
                for (int u = 0; u < RBLK; ++u) dot[u] = lane < jp ? vk[u] * sm.q[lane] + wk[u] * sm.q[32 + lane] : 0.f;
#pragma unroll
                for (int o = 16; o > 0; o >>= 1)
#pragma unroll
                    for (int u = 0; u < RBLK; ++u) dot[u] += __shfl_xor_sync(0xffffffffu, dot[u], o);
#pragma unroll
                for (int u = 0; u < RBLK; ++u) {
                    const int i = blk * RBLK + u;
                    const float vlast = __shfl_sync(0xffffffffu, vk[u], jp);
                    if (lane == jp && i > j && i < n) jb.Wp[(size_t)i * NB + lane] = tau_p * (yi[u] - dot[u]) - cw * vlast;
                }
            }
            team_barrier(jb.bar, target, team);
            TRD_TICK(7);
            rank2k_update(jb, sm, j + 1, jp + 1, rank);
            team_barrier(jb.bar, target, team, true);
            TRD_TICK(8);
            j0 = j + 1;
        }
    }
    if (rank == 0 && tid == 0) jb.d[n - 1] = __ldcg(jb.T + (size_t)(n - 1) * ldt + (n - 1));
    if (prof_on) {
        for (int q = 0; q < 9; ++q) jb.prof[q] = prof_acc[q];
        unsigned long long g; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(g)); jb.prof[10] = (long long)g;
        jb.prof[11] = prof_wait; jb.prof[12] = prof_units;
    }
#undef TRD_TICK
    __syncthreads();
}

__global__ void __launch_bounds__(THREADS, 1)
trd_kernel(const Job* jobs, const int* cta_begin, const int* entries) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    Smem& sm = *reinterpret_cast<Smem*>(smem_raw);
    float* dynf = reinterpret_cast<float*>(smem_raw + sizeof(Smem));
    __shared__ Job jb;
    if (threadIdx.x == 0) {
        for (int w = 0; w < NWARP; ++w)
            for (int s_ = 0; s_ < NSLOT; ++s_)
                asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_addr(&sm.mbar[w][s_])));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    uint32_t ring_par = 0;                                       // per warp: phase parity of its ring slots (persists across jobs)
    for (int e = cta_begin[blockIdx.x]; e < cta_begin[blockIdx.x + 1]; ++e) {
        __syncthreads();
        if (threadIdx.x == 0) jb = jobs[entries[2 * e]];
        __syncthreads();
        run_job(jb, entries[2 * e + 1], sm, dynf, ring_par);
    }
}

// lower triangle of src mirrored into the n x ldt work array (the reference's eigh reads the lower triangle only).
// grid (row tiles of the largest matrix, jobs); a block walks the tiles (ti, 0..ti) of its row of 32x32 tiles.
__global__ void trd_copy_kernel(const Job* jobs, const float* const* src) {
    __shared__ float tile[32][33];
    const Job jb = jobs[blockIdx.y];
    const int n = jb.n, ldt = jb.ldt;
    const int nt = (n + 31) >> 5;
    const int ti = blockIdx.x;
    if (ti >= nt) return;
    const float* A = src[blockIdx.y];
    const int lx = threadIdx.x & 31, ly = threadIdx.x >> 5;          // 256 threads
    for (int tj = 0; tj <= ti; ++tj) {
        __syncthreads();
        for (int r = ly; r < 32; r += 8) {
            const int gi = 32 * ti + r, gc = 32 * tj + lx;
            tile[r][lx] = (gi < n && gc < n) ? A[(size_t)gi * n + gc] : 0.f;
        }
        __syncthreads();
        for (int r = ly; r < 32; r += 8) {
            const int gi = 32 * ti + r, gc = 32 * tj + lx;
            if (gi < n && gc < ldt) {                                    // (pad columns [n, ldt) get zeros)
                float v = tile[r][lx];
                if (ti == tj && lx > r) v = tile[lx][r];
                jb.T[(size_t)gi * ldt + gc] = gc < n ? v : 0.f;
            }
            if (ti != tj) {
                const int ui = 32 * tj + r, uc = 32 * ti + lx;           // mirrored tile
                if (ui < n && uc < ldt) jb.T[(size_t)ui * ldt + uc] = uc < n ? tile[lx][r] : 0.f;
            }
        }
    }
}

// ---------------------------------------------------------------------------
// host: team sizes and the per-CTA schedule
// ---------------------------------------------------------------------------
// Cost model of one job on a team (seconds), fitted to B200 measurements of single matrices on forced team sizes
// (tools/trd_calib.py, profiles/r02_trd_calibration.txt; n = 256 .. 4608, teams 1 .. 147, all points within 6 %):
//   per column: the dependent chain of short phases + two team barriers -- 7.4 us on a single CTA, 11.8 us on teams
//               up to 16, + 1.5 us per doubling beyond (barrier fan-in, slowest-member wait);
//   row phases (finish w, form the column, publish y / the reflector): 25 ns per (row, column) pair / team;
//   symmetric product + rank-2k updates: 2.05e-11 s n^3 per SM (the product is instruction-issue bound at ~40 - 60 GB/s
//               per SM, so the HBM ceiling is not reached even by 147 CTAs).  A matrix that is L2-resident while it runs
//               ALONE measures 1.7e-11, but inside a batch nothing stays resident: the in-batch figure is used.
static double job_time(int n, int team) {
    double f = team > 1 ? 11.8e-6 : 7.4e-6;
    if (team > 16) f += 1.5e-6 * log2((double)team / 16.0);
    const double n2 = (double)n * n, n3 = n2 * n;
    return n * f + (0.0246e-6 * n2 + 2.05e-11 * n3) / team;
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
int encode_tmap(void* slot128, const float* T, int n, int ldt) {
    static EncodeTiledFn encode = nullptr;
    if (!encode) {
        cudaDriverEntryPointQueryResult qres;
        void* fn = nullptr;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres) != cudaSuccess || !fn) return -1;
        encode = (EncodeTiledFn)fn;
    }
    static_assert(sizeof(CUtensorMap) == 128, "tensor map slot");
    alignas(64) CUtensorMap tm;
    const cuuint64_t dims[2] = {(cuuint64_t)ldt, (cuuint64_t)n};
    const cuuint64_t strides[1] = {(cuuint64_t)ldt * sizeof(float)};
    const cuuint32_t box[2] = {128, (cuuint32_t)RBLK};
    const cuuint32_t estr[2] = {1, 1};
    const CUresult cr = encode(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(T), dims, strides, box, estr,
                               CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                               CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (cr != CUDA_SUCCESS) return (int)cr;
    memcpy(slot128, &tm, 128);
    return 0;
}

// dynamic shared memory of a CTA working on an n x n job in a team of `team`: fixed part + v | z | ysum
size_t smem_bytes(int n, int team) {
    const int ldt = (n + 7) & ~7;
    const int blocks = (n + RBLK - 1) / RBLK;
    const int ycap = RBLK * ((blocks + team - 1) / team + 1);
    return sizeof(Smem) + sizeof(float) * (size_t)(2 * ldt + ycap);
}
static bool fits(int n, int team) { return smem_bytes(n, team) <= (size_t)SMEM_LIMIT; }

int plan_teams(const int* n, int njobs, int sms, int* team, int* cta_begin, int* entries) {
    static int forced = -2;                                           // KFACB_TRD_TEAM=k: every job on a team of k (tuning aid)
    if (forced == -2) { const char* e = getenv("KFACB_TRD_TEAM"); forced = (e && *e) ? atoi(e) : -1; }
    struct TeamRec { int size; double load; int first_job, njobs; };
    int* order = new int[njobs];
    for (int i = 0; i < njobs; ++i) order[i] = i;
    for (int a = 1; a < njobs; ++a) {                                 // by n descending (stable)
        const int v = order[a];
        int b = a - 1;
        while (b >= 0 && n[order[b]] < n[v]) { order[b + 1] = order[b]; --b; }
        order[b + 1] = v;
    }
    int* t_of = new int[njobs];
    int* team_of_job = new int[njobs];
    TeamRec* teams = new TeamRec[njobs > 0 ? njobs : 1];
    int nteams = 0;
    auto try_pack = [&](double limit) -> bool {
        nteams = 0;
        int used = 0;
        for (int a = 0; a < njobs; ++a) {
            const int jn = n[order[a]];
            int tsel = -1;
            if (forced > 0) { tsel = forced > sms ? sms : forced; while (tsel < sms && !fits(jn, tsel)) ++tsel; }
            else for (int t = 1; t <= sms; ++t)                      // smallest team that meets the limit (and holds its vectors)
                if (fits(jn, t) && job_time(jn, t) <= limit) { tsel = t; break; }
            if (tsel < 0) return false;
            const double jt = job_time(jn, tsel);
            int k = -1;
            for (int q = 0; q < nteams; ++q)
                if (teams[q].size == tsel && teams[q].load + jt <= limit) { k = q; break; }
            if (k < 0) {
                if (used + tsel > sms) return false;
                k = nteams++;
                teams[k] = TeamRec{tsel, 0.0, 0, 0};
                used += tsel;
            }
            teams[k].load += jt;
            t_of[a] = tsel;
            team_of_job[a] = k;
        }
        return true;
    };
    double lo = 0.0, hi = 0.0;
    for (int i = 0; i < njobs; ++i) { hi += job_time(n[i], 1); const double t = job_time(n[i], sms); if (t > lo) lo = t; }
    if (hi < lo) hi = lo;
    while (!try_pack(hi)) hi *= 2.0;
    for (int it = 0; it < 40; ++it) {
        const double mid = 0.5 * (lo + hi);
        if (try_pack(mid)) hi = mid; else lo = mid;
    }
    try_pack(hi);
    // lay the teams out on consecutive CTAs; CTA c of team k runs the team's jobs in order
    int* team_base = new int[nteams > 0 ? nteams : 1];
    int used = 0;
    for (int k = 0; k < nteams; ++k) { team_base[k] = used; used += teams[k].size; }
    int ne = 0;
    for (int c = 0; c <= sms; ++c) cta_begin[c] = 0;
    for (int k = 0; k < nteams; ++k)
        for (int r = 0; r < teams[k].size; ++r) {
            const int c = team_base[k] + r;
            cta_begin[c] = ne;
            for (int a = 0; a < njobs; ++a)
                if (team_of_job[a] == k) { entries[2 * ne] = order[a]; entries[2 * ne + 1] = r; ++ne; }
        }
    for (int c = used; c <= sms; ++c) cta_begin[c] = ne;
    for (int a = 0; a < njobs; ++a) team[order[a]] = t_of[a];
    delete[] order; delete[] t_of; delete[] team_of_job; delete[] teams; delete[] team_base;
    return used > 0 ? used : 0;        // CTAs in use (the grid); entries written: sum over CTAs
}

cudaError_t launch(const Job* d_jobs, const int* d_cta_begin, const int* d_entries, int grid, size_t smem, cudaStream_t st) {
    if (grid <= 0) return cudaSuccess;
    if (smem > (size_t)SMEM_LIMIT) return cudaErrorInvalidValue;
    cudaError_t e = cudaFuncSetAttribute(trd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_LIMIT);
    if (e != cudaSuccess) return e;
    void* args[] = {(void*)&d_jobs, (void*)&d_cta_begin, (void*)&d_entries};
    e = cudaLaunchCooperativeKernel((const void*)trd_kernel, dim3(grid), dim3(THREADS), args, smem, st);
    count_launch();
    return e;
}

cudaError_t launch_copy(const Job* d_jobs, const float* const* d_src, int njobs, int nmax, cudaStream_t st) {
    if (njobs <= 0) return cudaSuccess;
    trd_copy_kernel<<<dim3((nmax + 31) / 32, njobs), 256, 0, st>>>(d_jobs, d_src);
    count_launch();
    return cudaGetLastError();
}

}  // namespace trd
}  // namespace kfacb
