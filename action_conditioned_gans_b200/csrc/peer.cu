// Data-parallel exchange over NVLink peer memory.
//
// The reference normalises over the WHOLE batch (slim.batch_norm inside one TF graph, models.py:11,32,81); with the
// batch sharded over GPUs every batch-norm layer therefore needs the sum of a tiny fp64 vector ([2C] moments forward,
// [2C] reduction terms backward, <= 8 KB) over all ranks -- ~66 times per training iteration.  A library all-reduce
// costs a launch plus a multi-hop protocol each time.  Here every rank owns a MAILBOX segment that all peers map
// (cudaIpc handles, exchanged by the host once); one single-CTA kernel per exchange
//   1. PUSHES its vector into slot[parity][my rank] of every peer's mailbox with plain NVLink stores,
//   2. publishes a monotonically increasing epoch flag with a system-scope release store,
//   3. spins (bounded) on its OWN mailbox until every peer's flag reached the epoch (local memory polling),
//   4. sums the `world` vectors in rank order (every rank gets bit-identical sums, so replicas never drift),
//   5. optionally finalises the batch-norm coefficients (mean / rstd / scale / shift) in the same launch.
// The epoch counter lives in device memory and is advanced by the kernel itself, so the launch can be captured in a
// CUDA graph and replayed.  Slots are double-buffered by epoch parity: a rank can be at most one exchange ahead of a
// peer on the same slot (it needs the peer's next flag to get further), so it never overwrites data still being read.
#include "common.cuh"

namespace acg {

struct PeerPtrs {
    unsigned char* p[ACG_MAX_PEERS];
};

__device__ __forceinline__ void st_release_sys(unsigned long long* p, unsigned long long v) {
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long* p) {
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ unsigned long long global_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}

// slot layout inside a mailbox, at byte offset slot_off (256-byte aligned):
//   flags[ACG_MAX_PEERS] u64 | pad to 128 B | data[2 parities][world][cap] fp64
__global__ void __launch_bounds__(256)
peer_allreduce_kernel(double* __restrict__ vec, int n, int cap, long long slot_off, int rank, int world, PeerPtrs peers,
                      unsigned long long* __restrict__ epoch_ptr, long long timeout_ns,
                      // optional batch-norm finalisation of vec = [sum | sum of squares][C]
                      int bn_C, const float* __restrict__ beta, double inv_rows, float eps, float* __restrict__ mean,
                      float* __restrict__ rstd, float* __restrict__ scale, float* __restrict__ shift) {
    // PDL: wait for the producer of `vec`, but do NOT let the dependents of this kernel be scheduled while it spins on
    // its peers -- their CTAs would sit on the SMs (griddepcontrol.wait) and keep the kernels of the step's other
    // branches, whose pushes the PEER is waiting for, from running: a cross-rank deadlock (seen at 2 GPUs).  The
    // trigger comes after the flag wait instead.
    pdl_wait();
    const int tid = threadIdx.x;
    const unsigned long long epoch = *epoch_ptr + 1ull;
    const size_t data_off = (size_t)slot_off + 128 + (size_t)(epoch & 1ull) * world * cap * sizeof(double);
    // 1. push
    for (int p = 0; p < world; ++p) {
        double* dst = reinterpret_cast<double*>(peers.p[p] + data_off) + (size_t)rank * cap;
        for (int i = tid; i < n; i += blockDim.x) dst[i] = vec[i];
    }
    __threadfence_system();
    __syncthreads();
    // 2. publish
    if (tid < world) {
        unsigned long long* flag = reinterpret_cast<unsigned long long*>(peers.p[tid] + slot_off) + rank;
        st_release_sys(flag, epoch);
    }
    // 3. wait for every peer (flags in MY mailbox)
    if (tid < world) {
        const unsigned long long* flag = reinterpret_cast<const unsigned long long*>(peers.p[rank] + slot_off) + tid;
        const unsigned long long t0 = global_ns();
        while (ld_acquire_sys(flag) < epoch) {
            if ((long long)(global_ns() - t0) > timeout_ns) {
                printf("acg: peer exchange timed out (rank %d waits for rank %d, slot offset %lld, epoch %llu)\n", rank,
                       tid, slot_off, epoch);
                __trap();
            }
        }
    }
    __syncthreads();
    pdl_launch_dependents();
    // 4. sum in rank order
    const double* mine = reinterpret_cast<const double*>(peers.p[rank] + data_off);
    for (int i = tid; i < n; i += blockDim.x) {
        double s = 0.0;
        for (int r = 0; r < world; ++r) s += __ldcg(mine + (size_t)r * cap + i);
        vec[i] = s;
    }
    // 5. batch-norm coefficients (same arithmetic as bn_finalize_kernel)
    if (bn_C > 0) {
        __syncthreads();
        for (int c = tid; c < bn_C; c += blockDim.x) {
            const double mu = vec[c] * inv_rows;
            double var = vec[bn_C + c] * inv_rows - mu * mu;
            if (var < 0.0) var = 0.0;
            const float rs = (float)(1.0 / sqrt(var + (double)eps));
            const float b = beta ? beta[c] : 0.f;
            mean[c] = (float)mu;
            rstd[c] = rs;
            scale[c] = rs;
            shift[c] = b - (float)mu * rs;
        }
    }
    if (tid == 0) *epoch_ptr = epoch;
}

}  // namespace acg

extern "C" {

int acg_peer_alloc(long long bytes, void** out_ptr) {
    using namespace acg;
    ACG_REQUIRE(bytes > 0 && out_ptr, ACG_ERR_INVALID, "acg_peer_alloc: bad arguments");
    void* p = nullptr;
    cudaError_t e = cudaMalloc(&p, (size_t)bytes);
    if (e == cudaSuccess) e = cudaMemset(p, 0, (size_t)bytes);
    if (e == cudaSuccess) e = cudaDeviceSynchronize();
    if (e != cudaSuccess) {
        set_error("acg_peer_alloc: CUDA error %d (%s)", (int)e, cudaGetErrorString(e));
        return ACG_ERR_CUDA;
    }
    *out_ptr = p;
    return ACG_OK;
}

int acg_peer_free(void* ptr) {
    using namespace acg;
    cudaError_t e = cudaFree(ptr);
    if (e != cudaSuccess) {
        set_error("acg_peer_free: CUDA error %d (%s)", (int)e, cudaGetErrorString(e));
        return ACG_ERR_CUDA;
    }
    return ACG_OK;
}

int acg_peer_export(void* ptr, void* host_handle64) {
    using namespace acg;
    static_assert(sizeof(cudaIpcMemHandle_t) == ACG_PEER_HANDLE_BYTES, "handle size");
    ACG_REQUIRE(ptr && host_handle64, ACG_ERR_INVALID, "acg_peer_export: null pointer");
    cudaIpcMemHandle_t h;
    cudaError_t e = cudaIpcGetMemHandle(&h, ptr);
    if (e != cudaSuccess) {
        set_error("acg_peer_export: CUDA error %d (%s)", (int)e, cudaGetErrorString(e));
        return ACG_ERR_CUDA;
    }
    memcpy(host_handle64, &h, sizeof(h));
    return ACG_OK;
}

int acg_peer_open(const void* host_handle64, void** out_ptr) {
    using namespace acg;
    ACG_REQUIRE(host_handle64 && out_ptr, ACG_ERR_INVALID, "acg_peer_open: null pointer");
    cudaIpcMemHandle_t h;
    memcpy(&h, host_handle64, sizeof(h));
    void* p = nullptr;
    cudaError_t e = cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess);
    if (e != cudaSuccess) {
        set_error("acg_peer_open: CUDA error %d (%s)", (int)e, cudaGetErrorString(e));
        return ACG_ERR_CUDA;
    }
    *out_ptr = p;
    return ACG_OK;
}

int acg_peer_close(void* ptr) {
    using namespace acg;
    cudaError_t e = cudaIpcCloseMemHandle(ptr);
    if (e != cudaSuccess) {
        set_error("acg_peer_close: CUDA error %d (%s)", (int)e, cudaGetErrorString(e));
        return ACG_ERR_CUDA;
    }
    return ACG_OK;
}

long long acg_peer_slot_bytes(int cap, int world) {
    if (cap <= 0 || world <= 0 || world > ACG_MAX_PEERS) return -1;
    long long b = 128 + 2ll * world * cap * (long long)sizeof(double);
    return (b + 255) / 256 * 256;
}

int acg_peer_allreduce_f64(double* vec, int n, int cap, long long slot_off, int rank, int world,
                           void* const* host_mailboxes, unsigned long long* epoch, float timeout_s, int bn_C,
                           const float* beta, long long bn_rows, float eps, float* mean, float* rstd, float* scale,
                           float* shift, void* stream) {
    using namespace acg;
    ACG_REQUIRE(vec && host_mailboxes && epoch, ACG_ERR_INVALID, "acg_peer_allreduce_f64: null pointer");
    ACG_REQUIRE(world >= 1 && world <= ACG_MAX_PEERS && rank >= 0 && rank < world, ACG_ERR_INVALID,
                "acg_peer_allreduce_f64: rank %d / world %d (at most %d peers)", rank, world, ACG_MAX_PEERS);
    ACG_REQUIRE(n > 0 && n <= cap && slot_off >= 0 && slot_off % 256 == 0, ACG_ERR_INVALID,
                "acg_peer_allreduce_f64: n %d cap %d slot offset %lld", n, cap, slot_off);
    ACG_REQUIRE(bn_C == 0 || (2 * bn_C == n && bn_rows > 0 && mean && rstd && scale && shift), ACG_ERR_INVALID,
                "acg_peer_allreduce_f64: batch-norm finalisation needs n == 2*C and the four outputs");
    PeerPtrs pp;
    for (int i = 0; i < ACG_MAX_PEERS; ++i) pp.p[i] = nullptr;
    for (int i = 0; i < world; ++i) {
        ACG_REQUIRE(host_mailboxes[i], ACG_ERR_INVALID, "acg_peer_allreduce_f64: mailbox %d is NULL", i);
        pp.p[i] = static_cast<unsigned char*>(host_mailboxes[i]);
    }
    const long long timeout_ns = (long long)((timeout_s > 0.f ? timeout_s : 30.f) * 1e9);
    launch_pdl(peer_allreduce_kernel, 1, 256, 0, static_cast<cudaStream_t>(stream), vec, n, cap, slot_off, rank, world, pp, epoch, timeout_ns, bn_C, beta, bn_C ? 1.0 / (double)bn_rows : 0.0, eps,
        mean, rstd, scale, shift);
    return check_launch("acg_peer_allreduce_f64");
}

}  // extern "C"
