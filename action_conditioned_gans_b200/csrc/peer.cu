// Data-parallel exchange over NVLink peer memory.
//
// The reference normalises over the WHOLE batch (slim.batch_norm inside one TF graph, models.py:11,32,81); with the
// batch sharded over GPUs every batch-norm layer therefore needs the sum of a tiny fp64 vector ([2C] moments forward,
// [2C] reduction terms backward, <= 8 KB) over all ranks -- ~63 times per training iteration, each one a point where
// every rank waits for the slowest.  A library all-reduce costs a launch plus a multi-hop protocol each time.  Here
// every rank owns a MAILBOX segment that all peers map (cudaIpc handles, exchanged by the host once) and an exchange is
//   1. PUSH: every value goes into the sender's row of every peer's mailbox as a self-validating 16-byte cell
//      (value halves + epoch tags, plain NVLink stores: no fence, no flag, one one-way latency),
//   2. PULL: poll the cells of the OWN mailbox (local memory) until they carry the epoch, add them in rank order (every
//      rank gets bit-identical sums, so replicas never drift),
//   3. optionally finalise the batch-norm coefficients (mean / rstd / scale / shift).
// It runs either as the single-CTA kernel below or inside the LAST CTA of the kernel that produced the vector (conv
// epilogue moments, backward reduction pass: no extra launch).  The epoch counter lives in device memory and is
// advanced by the exchange itself, so launches can be captured in a CUDA graph and replayed.  The device code and the
// double-buffering argument are in peer.cuh.
#include "peer.cuh"

namespace acg {

__global__ void __launch_bounds__(256)
peer_allreduce_kernel(double* __restrict__ vec, int n, PeerExchange x,
                      // optional batch-norm finalisation of vec = [sum | sum of squares][C]
                      int bn_C, const float* __restrict__ beta, double inv_rows, float eps, float* __restrict__ mean,
                      float* __restrict__ rstd, float* __restrict__ scale, float* __restrict__ shift) {
    // PDL: wait for the producer of `vec`, but do NOT let the dependents of this kernel be scheduled while it spins on
    // its peers -- their CTAs would sit on the SMs (griddepcontrol.wait) and keep the kernels of the step's other
    // branches, whose pushes the PEER is waiting for, from running: a cross-rank deadlock (seen at 2 GPUs).  The
    // trigger comes after the flag wait instead.
    pdl_wait();
    const int tid = threadIdx.x;
    peer_exchange(x, vec, n, tid, (int)blockDim.x, [] { pdl_launch_dependents(); });
    // batch-norm coefficients (same arithmetic as bn_finalize_kernel)
    if (bn_C > 0) {
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
}

int fill_peer_exchange(PeerExchange* x, const acg_peer_exchange* d, int n, const char* who) {
    ACG_REQUIRE(d && d->mailboxes && d->epoch, ACG_ERR_INVALID, "%s: peer exchange: null pointer", who);
    ACG_REQUIRE(d->world >= 1 && d->world <= ACG_MAX_PEERS && d->rank >= 0 && d->rank < d->world, ACG_ERR_INVALID,
                "%s: peer exchange: rank %d / world %d (at most %d peers)", who, d->rank, d->world, ACG_MAX_PEERS);
    ACG_REQUIRE(n > 0 && n <= d->cap && d->slot_off >= 0 && d->slot_off % 256 == 0, ACG_ERR_INVALID,
                "%s: peer exchange: n %d cap %d slot offset %lld", who, n, d->cap, d->slot_off);
    for (int i = 0; i < ACG_MAX_PEERS; ++i) x->mbox[i] = nullptr;
    for (int i = 0; i < d->world; ++i) {
        ACG_REQUIRE(d->mailboxes[i], ACG_ERR_INVALID, "%s: peer exchange: mailbox %d is NULL", who, i);
        x->mbox[i] = static_cast<unsigned char*>(d->mailboxes[i]);
    }
    x->epoch = d->epoch;
    x->slot_off = d->slot_off;
    x->timeout_ns = (long long)((d->timeout_s > 0.f ? d->timeout_s : 30.f) * 1e9);
    x->rank = d->rank; x->world = d->world; x->cap = d->cap;
    return ACG_OK;
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
    long long b = 128 + 2ll * world * cap * 16;      // 16-byte cells: value + tags (peer.cuh)
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
    acg_peer_exchange d{host_mailboxes, epoch, slot_off, rank, world, cap, timeout_s};
    PeerExchange x;
    int rc = fill_peer_exchange(&x, &d, n, "acg_peer_allreduce_f64");
    if (rc) return rc;
    launch_pdl(peer_allreduce_kernel, 1, 256, 0, static_cast<cudaStream_t>(stream), vec, n, x, bn_C, beta,
               bn_C ? 1.0 / (double)bn_rows : 0.0, eps, mean, rstd, scale, shift);
    return check_launch("acg_peer_allreduce_f64");
}

}  // extern "C"
