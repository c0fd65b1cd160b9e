/*
 * kfacb.h -- C ABI of libkfacb.so, the sm_100a CUDA implementation of the
 * DP-KFAC `preconditioner.step()` hot path.
 *
 * The reference (lzhangbv/kfac_pytorch @ c45acb8) has NO FFI on this path: its
 * boundary is the Python API of package `kfac`.  This
 * library is what a native replacement of the reference's per-layer math
 * exposes; every entry point cites the reference code it replaces (paths
 * relative to the reference repo root).  The Python package
 * `kfac_pytorch_b200` (and the `kfac` drop-in shim) binds it with ctypes; the
 * binding a reference maintainer would add is shown in INTEGRATION.md.
 *
 * Conventions
 *  - plain C: raw device pointers, sizes, a CUDA stream handle; no torch types.
 *  - every function returns a kfacb_status (0 = ok).  No C++ exceptions cross
 *    the boundary, nothing aborts.  kfacb_last_error() gives a message for the
 *    calling thread's last non-zero status.
 *  - the library never allocates or frees device memory: the caller owns all
 *    buffers, including workspaces whose sizes come from the *_workspace_bytes
 *    queries.  Workspaces must be 256-byte aligned.
 *  - all work is enqueued on `stream`; no call synchronises the device or the host
 *    (exceptions, both opt-in debug paths: KFACB_EIGH_ALGO=jacobi, KFACB_TRD_PROF=1).
 *  - host arrays (descriptors, pointer tables) are read before the call
 *    returns and may be reused immediately.
 *  - all matrices are dense row-major fp32.
 *  - a NULL entry in a per-layer pointer table means "this rank does not own
 *    that layer/factor": the layer is skipped (reference: the
 *    `backend.comm.rank() == rank_a` guards, kfac_preconditioner_inv_dp.py:79,86).
 */
#ifndef KFACB_H_
#define KFACB_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef void* kfacb_stream_t;              /* cudaStream_t */

typedef enum {
    KFACB_OK = 0,
    KFACB_ERR_ARG = 1,          /* bad shape / null pointer / unsupported layer kind
                                   (reference: NotImplementedError, kfac/utils.py:82,115) */
    KFACB_ERR_CUDA = 2,         /* a CUDA runtime call failed */
    KFACB_ERR_WORKSPACE = 3,    /* workspace too small or misaligned */
    KFACB_ERR_ARCH = 4          /* device is not sm_100 */
} kfacb_status;

/* kinds of preconditioned layers (reference: class-name match on
 * {'Linear','Conv2d'}, kfac/kfac_preconditioner_base.py:132-147) */
enum { KFACB_LINEAR = 0, KFACB_CONV2D = 1 };

/* Shape of one layer.  Built once at hook-registration time from the module
 * (what the reference reads off `layer` in kfac/utils.py:85-103,118-140). */
typedef struct {
    int32_t kind;       /* KFACB_LINEAR | KFACB_CONV2D */
    int32_t has_bias;   /* appends the homogeneous "ones" column to A and the bias column to the gradient */
    int32_t batch;      /* B: leading dimension of a and g */
    int32_t seq;        /* linear: T when a/g are (B,T,D) -- they are mean-reduced over T
                           (kfac/utils.py:98-99,133-134); 1 for (B,D) inputs; conv: 1 */
    int32_t in_c;       /* linear: in_features; conv: C_in */
    int32_t out_c;      /* linear: out_features; conv: C_out */
    int32_t in_h, in_w; /* conv input spatial size (unused for linear) */
    int32_t kh, kw;     /* conv kernel */
    int32_t sh, sw;     /* conv stride */
    int32_t ph, pw;     /* conv zero padding (dilation/groups are ignored, as in kfac/utils.py:33-54) */
    int32_t out_h, out_w;
} kfacb_layer_t;

/* dA = in_c*kh*kw (+1 if bias)  [linear: kh = kw = 1];  dG = out_c */
int32_t kfacb_layer_dA(const kfacb_layer_t* l);
int32_t kfacb_layer_dG(const kfacb_layer_t* l);

int         kfacb_version(void);
/* sha256 of the sources + flags this library was built from (kfac_pytorch_b200/build.py) */
const char* kfacb_build_fingerprint(void);
/* number of CUDA kernels this library has launched in this process (monotonic; for benchmarking) */
uint64_t    kfacb_launch_count(void);
const char* kfacb_last_error(void);
/* fills SM count and compute capability of the current device */
int         kfacb_device_info(int32_t* sm_count, int32_t* cc_major, int32_t* cc_minor);

/* ---------------------------------------------------------------------------
 * (1) Kronecker-factor accumulation fused with the running average.
 * Replaces ComputeA / ComputeG / update_running_avg and the identity
 * initialisation: kfac/utils.py:33-54,66-140 and
 * kfac/kfac_preconditioner_inv_dp.py:75-90.
 *
 *   A_new = [P 1]^T [P 1] / (B*S^2)  (conv; P = im2col rows, never materialised)
 *         =  [a 1]^T [a 1] / B       (linear; a mean-reduced over T first)
 *   G_new = B*S * g_rows^T g_rows    (conv)      = B * g^T g   (linear)
 *   first_step:  factor <- (1-alpha)*I + alpha*new     (identity init, inv_dp.py:82-84)
 *   otherwise :  factor <- (1-alpha)*factor + alpha*new   (alpha = factor_decay)
 *
 * a[i]: layer input  (B,C,H,W) | (B,D) | (B,T,D);  g[i]: grad wrt the layer
 * output (B,C_out,oh,ow) | (B,D') | (B,T,D'), of the batch-AVERAGED loss.
 * A[i] is dA x dA, G[i] is dG x dG.  a[i]==NULL skips A, g[i]==NULL skips G.
 * ------------------------------------------------------------------------- */
size_t kfacb_factor_workspace_bytes(const kfacb_layer_t* layers, int32_t count);
int    kfacb_factor_update(const kfacb_layer_t* layers, int32_t count,
                           const float* const* a, const float* const* g,
                           float* const* A, float* const* G,
                           float alpha, int32_t first_step,
                           void* workspace, size_t workspace_bytes,
                           kfacb_stream_t stream);

/* ---------------------------------------------------------------------------
 * (2a) Batched symmetric eigendecomposition with eigenvalue floor.
 * Replaces mat_eig + the floor mask: kfac/utils.py:22-25 and
 * kfac/kfac_preconditioner_eigen_dp.py:62-75 (and the unused cuSOLVER wrapper
 * packages/tcmm/src/tcmm_kernel.cu:56-116).
 *
 * For each i: M[i] (n[i] x n[i], symmetric -- the lower triangle is read, like
 * torch.linalg.eigh) = Q diag(d) Q^T with d ascending, eigenvectors in the COLUMNS
 * of Q[i] (row-major n x n), then d <- d * (d > eps_floor).  Eigenvectors are
 * defined up to sign / rotation inside degenerate clusters.
 *   n <= 64 : two-sided cyclic Jacobi, one CTA per matrix, in shared memory.
 *   n  > 64 : Householder tridiagonalisation of the whole batch in ONE persistent
 *             kernel (teams of CTAs per matrix), divide and conquer on the
 *             tridiagonal matrices, block-reflector back-transformation on the
 *             tensor cores.  Fixed launch sequence: the call only ENQUEUES work (no
 *             convergence test, no host synchronisation).  Not graph-capturable as is: the per-call
 *             job tables are uploaded from a transient host image.
 *             n <= 8192.  (KFACB_EIGH_ALGO=jacobi selects the round-1 block-Jacobi
 *             path, which does synchronise the host on its progress events.)
 * sweeps_out (device int32[count], may be NULL) receives the number of Jacobi
 * sweeps (n <= 64) or 1.
 * ------------------------------------------------------------------------- */
size_t kfacb_eigh_workspace_bytes(const int32_t* n, int32_t count);
int    kfacb_eigh_batched(const int32_t* n, int32_t count,
                          const float* const* M, float* const* Q, float* const* d,
                          float eps_floor, int32_t* sweeps_out,
                          void* workspace, size_t workspace_bytes,
                          kfacb_stream_t stream);

/* Scheduler introspection (no GPU needed): team_out[i] = CTAs that tridiagonalise matrix i on a device with
 * `sms` SMs (0 for n[i] <= 64); returns the number of CTAs of the persistent kernel, < 0 on bad arguments. */
int    kfacb_eigh_plan_teams(const int32_t* n, int32_t count, int32_t sms, int32_t* team_out);

/* Benchmark aid.  kfacb_eigh_batched records CUDA events on the caller's stream around its stages (asynchronous);
 * this call waits for the last one and returns, for the LAST call on the current device:
 *   out8 = { ms tridiagonalisation, ms divide and conquer, ms back-transformation + write,
 *            algorithmic bytes of the tridiagonalisation (sum 4/3 n^3: one read of the trailing matrix per column),
 *            sum n^3 of the n > 128 matrices, their count, CTAs of the persistent kernel, 0 }
 * (With KFACB_EIGH_ALGO=jacobi and KFACB_PROFILE=1: the per-round kernel totals of the block-Jacobi path,
 *  { rounds, ms projected-block, ms solve, ms update, bytes H, bytes update, matrices, block pairs }.) */
int    kfacb_eigh_last_profile(double* out8);

/* ---------------------------------------------------------------------------
 * (2b) pi-damped SPD inverse.  Replaces
 * kfac/kfac_preconditioner_inv_dp.py:98-123, _add_value_to_diagonal
 * (kfac_preconditioner_inv.py:106-107) and mat_inv (kfac/utils.py:11-16).
 *
 *   pi = sqrt((tr A / dA) / (tr G / dG))
 *   A += sqrt(damping)*pi * I   IN PLACE (persists in the running factor)
 *   G += sqrt(damping)/pi * I   IN PLACE
 *   invA = A^-1, invG = G^-1    (full symmetric inverses)
 * A[i]==NULL skips layer i.
 * ------------------------------------------------------------------------- */
size_t kfacb_spd_inverse_workspace_bytes(const int32_t* nA, const int32_t* nG, int32_t count);
int    kfacb_pi_damped_inverse(const int32_t* nA, const int32_t* nG, int32_t count,
                               float* const* A, float* const* G,
                               float* const* invA, float* const* invG,
                               float damping,
                               void* workspace, size_t workspace_bytes,
                               kfacb_stream_t stream);

/* ---------------------------------------------------------------------------
 * (3) Two-sided preconditioned gradient.  The gradient matrix
 * [W.view(out,-1) | b] (kfac_preconditioner_inv.py:145-154) is read straight
 * from wgrad/bgrad (no concatenated copy).
 *
 * eigen   (kfac_preconditioner_eigen_dp.py:87-89):
 *     V1 = QG^T grad QA;  V2 = V1 / (dG dA^T + damping);  out = QG V2 QA^T
 * inverse (kfac_preconditioner_inv_dp.py:135):
 *     out = invG grad invA
 * out[i] is the dG x dA preconditioned gradient (`m_precon_grad`), the buffer
 * that is exchanged between ranks.  wgrad[i]==NULL skips layer i.
 * ------------------------------------------------------------------------- */
size_t kfacb_precond_workspace_bytes(const kfacb_layer_t* layers, int32_t count);
int    kfacb_precond_eigen(const kfacb_layer_t* layers, int32_t count,
                           const float* const* QA, const float* const* QG,
                           const float* const* dA, const float* const* dG,
                           const float* const* wgrad, const float* const* bgrad,
                           float* const* out, float damping,
                           void* workspace, size_t workspace_bytes,
                           kfacb_stream_t stream);
int    kfacb_precond_inverse(const kfacb_layer_t* layers, int32_t count,
                             const float* const* invA, const float* const* invG,
                             const float* const* wgrad, const float* const* bgrad,
                             float* const* out,
                             void* workspace, size_t workspace_bytes,
                             kfacb_stream_t stream);

/* Same as kfacb_precond_eigen / _inverse, but the epilogue of the LAST GEMM also stores every output element at
 * byte offset peer_deltas[q] from its local address, for q < npeers: with the deltas of an attached exchange
 * (kfacb_exchange_peer_deltas) the preconditioned gradient lands in every peer's slab over NVLink while it is
 * being produced -- the exchange of kfac_preconditioner_inv.py:164-175 fused into the producing kernel.
 * `out[i]` must point into the local exchange slab.  Bracket with kfacb_exchange_begin / _signal. */
int    kfacb_precond_eigen_bcast(const kfacb_layer_t* layers, int32_t count,
                                 const float* const* QA, const float* const* QG,
                                 const float* const* dA, const float* const* dG,
                                 const float* const* wgrad, const float* const* bgrad,
                                 float* const* out, float damping,
                                 int32_t npeers, const int64_t* peer_deltas,
                                 void* workspace, size_t workspace_bytes,
                                 kfacb_stream_t stream);
int    kfacb_precond_inverse_bcast(const kfacb_layer_t* layers, int32_t count,
                                   const float* const* invA, const float* const* invG,
                                   const float* const* wgrad, const float* const* bgrad,
                                   float* const* out,
                                   int32_t npeers, const int64_t* peer_deltas,
                                   void* workspace, size_t workspace_bytes,
                                   kfacb_stream_t stream);

/* ---------------------------------------------------------------------------
 * (3b) Exchange of the preconditioned gradients between the ranks of ONE node.
 * Replaces `_communicate_pred`, kfac/kfac_preconditioner_inv.py:164-175 (one hvd/torch broadcast per layer).
 *
 * Every rank creates an exchange with the same slab size; the slab (device memory OWNED BY THE EXCHANGE -- the one
 * exception to "the library never allocates": it has to be IPC-exportable) holds all layers' preconditioned
 * gradients.  The 128-byte handles are all-gathered by the caller (any host transport) and connected: the slabs
 * are then mapped into each other (CUDA IPC, peer memory over NVLink / NVSwitch).  Per step, on the caller's
 * stream, in this order on every rank (all calls only enqueue tiny kernels):
 *     kfacb_exchange_begin     wait until every peer has released the previous step's slab
 *     ... produce the owned segment: kfacb_precond_*_bcast (fused: stored into every peer while computed), or
 *         plain kfacb_precond_* followed by kfacb_exchange_push(offset, bytes) (one copy kernel, then signals) ...
 *     kfacb_exchange_signal    (fused variant only) tell every peer that this rank's segment has landed
 *     kfacb_exchange_wait      wait until every peer's segment has landed in the local slab
 *     ... consume the whole slab (kfacb_finalize_grads) ...
 *     kfacb_exchange_release   tell every peer that the local slab may be overwritten
 * A rank that owns no layer still calls begin / signal (or push with 0 bytes) / wait / release.
 * ------------------------------------------------------------------------- */
typedef struct kfacb_exchange kfacb_exchange_t;
#define KFACB_EXCHANGE_HANDLE_BYTES 128
int    kfacb_exchange_create(int32_t rank, int32_t size, size_t slab_bytes, kfacb_exchange_t** out);
int    kfacb_exchange_handle(kfacb_exchange_t* ex, void* handle_out /* KFACB_EXCHANGE_HANDLE_BYTES */);
int    kfacb_exchange_connect(kfacb_exchange_t* ex, const void* all_handles /* size x KFACB_EXCHANGE_HANDLE_BYTES, by rank */);
void*  kfacb_exchange_slab(kfacb_exchange_t* ex);
int    kfacb_exchange_peer_deltas(kfacb_exchange_t* ex, int64_t* deltas_out /* size - 1 */, int32_t* npeers_out);
int    kfacb_exchange_begin(kfacb_exchange_t* ex, kfacb_stream_t stream);
int    kfacb_exchange_signal(kfacb_exchange_t* ex, kfacb_stream_t stream);
int    kfacb_exchange_push(kfacb_exchange_t* ex, size_t offset_bytes, size_t nbytes, kfacb_stream_t stream);
int    kfacb_exchange_wait(kfacb_exchange_t* ex, kfacb_stream_t stream);
int    kfacb_exchange_release(kfacb_exchange_t* ex, kfacb_stream_t stream);
int    kfacb_exchange_destroy(kfacb_exchange_t* ex);

/* ---------------------------------------------------------------------------
 * (4) KL-clip and in-place gradient rewrite.  Replaces
 * _reshape_preconditioned_grad + _update_grad_in_place,
 * kfac_preconditioner_inv.py:178-217, without the per-tensor `.item()` syncs:
 *
 *   vg_sum = sum_layers sum(v * grad) * lr^2        (grad = the pre-overwrite grads)
 *   nu     = kl_clip > 0 ? min(1, sqrt(kl_clip / |vg_sum|)) : 1     (force_nu_one -> 1)
 *   wgrad  <- nu * v[:, :-1],  bgrad <- nu * v[:, -1]
 *
 * precon[i] must be non-NULL for EVERY layer (after the exchange all ranks hold
 * all preconditioned gradients).  nu_out: device float[1] receiving nu.
 * workspace: kfacb_finalize_workspace_bytes(count) bytes.
 * ------------------------------------------------------------------------- */
size_t kfacb_finalize_workspace_bytes(const kfacb_layer_t* layers, int32_t count);
int    kfacb_finalize_grads(const kfacb_layer_t* layers, int32_t count,
                            const float* const* precon,
                            float* const* wgrad, float* const* bgrad,
                            float lr, float kl_clip, int32_t force_nu_one,
                            float* nu_out,
                            void* workspace, size_t workspace_bytes,
                            kfacb_stream_t stream);

/* ---------------------------------------------------------------------------
 * (4b) Split finalize + gradient packing: the pieces of the FUSED gradient-reduce
 * path (reduce-to-owner -> precondition -> exchange instead of the caller's
 * all-reduce followed by step(); examples/pytorch_imagenet_resnet.py:371-383 and
 * kfac_preconditioner_inv.py:188-217).
 *
 *   kfacb_pack_grads     dst[i] <- [wgrad[i] (dG x dW) | bgrad[i] (dG)], contiguous: the
 *                        send buffer of the reduce-to-owner; a segment of that buffer is
 *                        afterwards handed to kfacb_precond_* as wgrad / bgrad pointers.
 *   kfacb_finalize_dot   vg_sum_out[0] = sum over the GIVEN layers of sum(v * grad)
 *                        (double, fixed order, no lr^2): an owner's partial of the KL-clip
 *                        sum, to be summed across ranks by the caller (one 8-byte
 *                        all-reduce).
 *   kfacb_finalize_write nu from the GLOBAL vg_sum (device double[1]) exactly like
 *                        kfacb_finalize_grads, then the in-place rewrite of every layer.
 * workspace: kfacb_finalize_workspace_bytes(layers, count) for all three.
 * ------------------------------------------------------------------------- */
int    kfacb_pack_grads(const kfacb_layer_t* layers, int32_t count,
                        const float* const* wgrad, const float* const* bgrad, float* const* dst,
                        void* workspace, size_t workspace_bytes, kfacb_stream_t stream);
int    kfacb_finalize_dot(const kfacb_layer_t* layers, int32_t count,
                          const float* const* precon,
                          const float* const* wgrad, const float* const* bgrad,
                          double* vg_sum_out,
                          void* workspace, size_t workspace_bytes, kfacb_stream_t stream);
int    kfacb_finalize_write(const kfacb_layer_t* layers, int32_t count,
                            const float* const* precon,
                            float* const* wgrad, float* const* bgrad,
                            const double* vg_sum, float lr, float kl_clip, int32_t force_nu_one,
                            float* nu_out,
                            void* workspace, size_t workspace_bytes, kfacb_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* KFACB_H_ */
