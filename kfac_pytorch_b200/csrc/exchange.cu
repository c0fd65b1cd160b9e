// exchange.cu -- the owner -> everyone exchange of the preconditioned gradients as a library component
// (C-ABI: kfacb_exchange_*).  Replaces the per-layer broadcasts of the reference,
// kfac/kfac_preconditioner_inv.py:164-175 (`_communicate_pred`), for the ranks of ONE node.
//
// Design (B200 / NVSwitch): every rank owns one SLAB of the same layout (all layers' preconditioned gradients,
// grouped by owner rank) allocated by this component and mapped into every other rank of the node through CUDA
// IPC -- peer memory over NVLink.  A rank never "sends" afterwards: the epilogue of its last preconditioning GEMM
// (precond.cu, EpiStoreBcast) stores every output element into its own slab AND into the same offset of every
// peer's slab, so the transfer is overlapped with the math tile by tile and there is no separate collective.
// What remains is synchronisation, four tiny kernels per step on the caller's stream:
//   begin    wait until every peer has RELEASED the previous epoch (it finished reading its slab)   [WAR]
//   signal   publish "my segment of epoch e is in your slab" to every peer (st.release.sys)          [RAW]
//   wait     spin (ld.acquire.sys on local flags) until every peer's segment of epoch e has arrived
//   release  tell every peer that this rank is done reading epoch e
// kfacb_exchange_push is the unfused variant (one kernel copies the caller's segment to all peers), used when the
// producer is not one of this library's GEMMs (the MPD variants) and as the A/B baseline of the fused epilogue.
#include <string.h>
#include "common.cuh"

namespace kfacb {

constexpr int EX_MAX_RANKS = 16;

struct Exchange {
    int rank = 0, size = 1, dev = 0;
    size_t slab_bytes = 0;
    float* slab = nullptr;                 // local
    int* flags = nullptr;                  // local: arrived[EX_MAX_RANKS], released[EX_MAX_RANKS], counter
    float* peer_slab[EX_MAX_RANKS] = {nullptr};
    int* peer_flags[EX_MAX_RANKS] = {nullptr};
    bool opened[EX_MAX_RANKS] = {false};
    int epoch = 0;
    bool connected = false;
};

struct ExPeers { float* slab[EX_MAX_RANKS]; int* flags[EX_MAX_RANKS]; };

__device__ __forceinline__ int ld_acquire_sys(const int* p) {
    int v;
    asm volatile("ld.acquire.sys.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_sys(int* p, int v) {
    asm volatile("st.release.sys.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

// one thread per peer: spin until flags[base + peer] >= epoch
__global__ void ex_wait_kernel(const int* flags, int base, int size, int rank, int epoch) {
    const int p = threadIdx.x;
    if (p < size && p != rank)
        while (ld_acquire_sys(flags + base + p) < epoch) { }
}
// one thread per peer: flags_of_peer[base + rank] = epoch
__global__ void ex_post_kernel(ExPeers peers, int base, int size, int rank, int epoch) {
    const int p = threadIdx.x;
    if (p < size && p != rank) {
        __threadfence_system();
        st_release_sys(peers.flags[p] + base + rank, epoch);
    }
}
// unfused push: own segment -> every peer, then arrived[rank] = epoch on every peer (last block)
__global__ void __launch_bounds__(256)
ex_push_kernel(ExPeers peers, const float* local, size_t off_f4, size_t n_f4, int size, int rank, int epoch, int* counter) {
    const float4* src = reinterpret_cast<const float4*>(local) + off_f4;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_f4; i += (size_t)gridDim.x * blockDim.x) {
        const float4 v = src[i];
        for (int p = 0; p < size; ++p)
            if (p != rank) (reinterpret_cast<float4*>(peers.slab[p]) + off_f4)[i] = v;
    }
    __threadfence_system();
    __syncthreads();
    __shared__ int last;
    if (threadIdx.x == 0) last = (atomicAdd(counter, 1) == (int)gridDim.x - 1);
    __syncthreads();
    if (last) {
        if (threadIdx.x == 0) *counter = 0;
        if ((int)threadIdx.x < size && (int)threadIdx.x != rank) {
            __threadfence_system();
            st_release_sys(peers.flags[threadIdx.x] + rank, epoch);
        }
    }
}

static ExPeers peers_of(const Exchange* ex) {
    ExPeers p;
    for (int i = 0; i < EX_MAX_RANKS; ++i) { p.slab[i] = ex->peer_slab[i]; p.flags[i] = ex->peer_flags[i]; }
    return p;
}

}  // namespace kfacb

using namespace kfacb;

extern "C" int kfacb_exchange_create(int32_t rank, int32_t size, size_t slab_bytes, kfacb_exchange_t** out) {
    KFACB_REQUIRE(out && size >= 1 && size <= EX_MAX_RANKS && rank >= 0 && rank < size, "bad rank / size (<= %d ranks of one node)", EX_MAX_RANKS);
    Exchange* ex = new Exchange();
    ex->rank = rank; ex->size = size;
    ex->slab_bytes = align_up(slab_bytes > 0 ? slab_bytes : 16, 256);
    if (cudaGetDevice(&ex->dev) != cudaSuccess ||
        cudaMalloc((void**)&ex->slab, ex->slab_bytes) != cudaSuccess ||
        cudaMalloc((void**)&ex->flags, 256) != cudaSuccess ||
        cudaMemset(ex->slab, 0, ex->slab_bytes) != cudaSuccess || cudaMemset(ex->flags, 0, 256) != cudaSuccess) {
        set_error("exchange: cannot allocate the slab (%zu bytes): %s", ex->slab_bytes, cudaGetErrorString(cudaGetLastError()));
        if (ex->slab) cudaFree(ex->slab);
        if (ex->flags) cudaFree(ex->flags);
        delete ex;
        return KFACB_ERR_CUDA;
    }
    ex->peer_slab[rank] = ex->slab;
    ex->peer_flags[rank] = ex->flags;
    *out = reinterpret_cast<kfacb_exchange_t*>(ex);
    return KFACB_OK;
}

extern "C" int kfacb_exchange_handle(kfacb_exchange_t* h, void* handle_out) {
    Exchange* ex = reinterpret_cast<Exchange*>(h);
    KFACB_REQUIRE(ex && handle_out, "null argument");
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
    cudaIpcMemHandle_t hs, hf;
    KFACB_CHECK_CUDA(cudaIpcGetMemHandle(&hs, ex->slab));
    KFACB_CHECK_CUDA(cudaIpcGetMemHandle(&hf, ex->flags));
    memcpy(handle_out, &hs, 64);
    memcpy((char*)handle_out + 64, &hf, 64);
    return KFACB_OK;
}

extern "C" int kfacb_exchange_connect(kfacb_exchange_t* h, const void* all_handles) {
    Exchange* ex = reinterpret_cast<Exchange*>(h);
    KFACB_REQUIRE(ex && all_handles, "null argument");
    for (int p = 0; p < ex->size; ++p) {
        if (p == ex->rank || ex->opened[p]) continue;
        cudaIpcMemHandle_t hs, hf;
        memcpy(&hs, (const char*)all_handles + (size_t)p * KFACB_EXCHANGE_HANDLE_BYTES, 64);
        memcpy(&hf, (const char*)all_handles + (size_t)p * KFACB_EXCHANGE_HANDLE_BYTES + 64, 64);
        KFACB_CHECK_CUDA(cudaIpcOpenMemHandle((void**)&ex->peer_slab[p], hs, cudaIpcMemLazyEnablePeerAccess));
        KFACB_CHECK_CUDA(cudaIpcOpenMemHandle((void**)&ex->peer_flags[p], hf, cudaIpcMemLazyEnablePeerAccess));
        ex->opened[p] = true;
    }
    ex->connected = true;
    return KFACB_OK;
}

extern "C" void* kfacb_exchange_slab(kfacb_exchange_t* h) {
    Exchange* ex = reinterpret_cast<Exchange*>(h);
    return ex ? ex->slab : nullptr;
}

extern "C" int kfacb_exchange_peer_deltas(kfacb_exchange_t* h, int64_t* deltas_out, int32_t* npeers_out) {
    Exchange* ex = reinterpret_cast<Exchange*>(h);
    KFACB_REQUIRE(ex && deltas_out && npeers_out && ex->connected, "exchange not connected");
    int k = 0;
    for (int p = 0; p < ex->size; ++p)
        if (p != ex->rank) deltas_out[k++] = (int64_t)((intptr_t)ex->peer_slab[p] - (intptr_t)ex->slab);
    *npeers_out = k;
    return KFACB_OK;
}

extern "C" int kfacb_exchange_begin(kfacb_exchange_t* h, kfacb_stream_t stream) {
    Exchange* ex = reinterpret_cast<Exchange*>(h);
    KFACB_REQUIRE(ex && ex->connected, "exchange not connected");
    ex->epoch += 1;
    if (ex->size > 1) {
        ex_wait_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(ex->flags, EX_MAX_RANKS, ex->size, ex->rank, ex->epoch - 1);
        KFACB_CHECK_LAUNCH();
    }
    return KFACB_OK;
}

extern "C" int kfacb_exchange_signal(kfacb_exchange_t* h, kfacb_stream_t stream) {
    Exchange* ex = reinterpret_cast<Exchange*>(h);
    KFACB_REQUIRE(ex && ex->connected && ex->epoch > 0, "exchange: begin first");
    if (ex->size > 1) {
        ex_post_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(peers_of(ex), 0, ex->size, ex->rank, ex->epoch);
        KFACB_CHECK_LAUNCH();
    }
    return KFACB_OK;
}

extern "C" int kfacb_exchange_push(kfacb_exchange_t* h, size_t offset_bytes, size_t nbytes, kfacb_stream_t stream) {
    Exchange* ex = reinterpret_cast<Exchange*>(h);
    KFACB_REQUIRE(ex && ex->connected && ex->epoch > 0, "exchange: begin first");
    KFACB_REQUIRE((offset_bytes & 15) == 0 && (nbytes & 15) == 0 && offset_bytes + nbytes <= ex->slab_bytes,
                  "exchange: segment must be 16-byte aligned and inside the slab");
    if (ex->size > 1) {
        const size_t n4 = nbytes / 16;
        int grid = (int)((n4 + 255) / 256);
        if (grid > 4 * sm_count()) grid = 4 * sm_count();
        if (grid < 1) grid = 1;
        ex_push_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(peers_of(ex), ex->slab, offset_bytes / 16, n4, ex->size, ex->rank,
                                                               ex->epoch, ex->flags + 2 * EX_MAX_RANKS);
        KFACB_CHECK_LAUNCH();
    }
    return KFACB_OK;
}

extern "C" int kfacb_exchange_wait(kfacb_exchange_t* h, kfacb_stream_t stream) {
    Exchange* ex = reinterpret_cast<Exchange*>(h);
    KFACB_REQUIRE(ex && ex->connected && ex->epoch > 0, "exchange: begin first");
    if (ex->size > 1) {
        ex_wait_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(ex->flags, 0, ex->size, ex->rank, ex->epoch);
        KFACB_CHECK_LAUNCH();
    }
    return KFACB_OK;
}

extern "C" int kfacb_exchange_release(kfacb_exchange_t* h, kfacb_stream_t stream) {
    Exchange* ex = reinterpret_cast<Exchange*>(h);
    KFACB_REQUIRE(ex && ex->connected && ex->epoch > 0, "exchange: begin first");
    if (ex->size > 1) {
        ex_post_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(peers_of(ex), EX_MAX_RANKS, ex->size, ex->rank, ex->epoch);
        KFACB_CHECK_LAUNCH();
    }
    return KFACB_OK;
}

extern "C" int kfacb_exchange_destroy(kfacb_exchange_t* h) {
    Exchange* ex = reinterpret_cast<Exchange*>(h);
    if (!ex) return KFACB_OK;
    cudaDeviceSynchronize();
    for (int p = 0; p < ex->size; ++p)
        if (p != ex->rank && ex->opened[p]) { cudaIpcCloseMemHandle(ex->peer_slab[p]); cudaIpcCloseMemHandle(ex->peer_flags[p]); }
    cudaFree(ex->slab);
    cudaFree(ex->flags);
    delete ex;
    (void)cudaGetLastError();
    return KFACB_OK;
}
