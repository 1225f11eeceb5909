// Shared helpers for the acg_b200 kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <utility>

#include "../../include/acg_b200.h"

namespace acg {

void set_error(const char* fmt, ...);
void count_launch(int n = 1);
int num_sms();

// Checks the launch that was just enqueued; returns ACG_OK or ACG_ERR_CUDA.
int check_launch(const char* what);

#define ACG_REQUIRE(cond, code, ...)          \
    do {                                      \
        if (!(cond)) {                        \
            acg::set_error(__VA_ARGS__);      \
            return (code);                    \
        }                                     \
    } while (0)

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

template <typename T> __device__ __forceinline__ float ld_as_float(const T* p, size_t i);
template <> __device__ __forceinline__ float ld_as_float<float>(const float* p, size_t i) { return p[i]; }
template <> __device__ __forceinline__ float ld_as_float<__nv_bfloat16>(const __nv_bfloat16* p, size_t i) {
    return __bfloat162float(p[i]);
}
template <typename T> __device__ __forceinline__ void st_from_float(T* p, size_t i, float v);
template <> __device__ __forceinline__ void st_from_float<float>(float* p, size_t i, float v) { p[i] = v; }
template <> __device__ __forceinline__ void st_from_float<__nv_bfloat16>(__nv_bfloat16* p, size_t i, float v) {
    p[i] = __float2bfloat16_rn(v);
}

// activation applied to the pre-activation value u
__device__ __forceinline__ float act_fwd(float u, int act) {
    switch (act) {
        case ACG_ACT_RELU: return fmaxf(u, 0.f);
        case ACG_ACT_LRELU: return 0.6f * u + 0.4f * fabsf(u);   // ops.py:22-26 with leak 0.2
        case ACG_ACT_TANH: return tanhf(u);
        default: return u;
    }
}
// derivative of the activation w.r.t. the pre-activation u (TF: relu' = u>0, abs' = sign)
__device__ __forceinline__ float act_bwd(float u, int act) {
    switch (act) {
        case ACG_ACT_RELU: return u > 0.f ? 1.f : 0.f;
        case ACG_ACT_LRELU: return 0.6f + 0.4f * (u > 0.f ? 1.f : (u < 0.f ? -1.f : 0.f));
        case ACG_ACT_TANH: { float t = tanhf(u); return 1.f - t * t; }
        default: return 1.f;
    }
}

// ---- programmatic dependent launch (PDL) ------------------------------------------------------------------------
// A training iteration is ~200 dependent launches inside a CUDA graph; with a plain kernel -> kernel edge the next grid
// is only scheduled once the previous one has fully drained.  Every kernel here is launched with the programmatic
// stream-serialization attribute and (a) allows its dependents to be scheduled as soon as all of its own CTAs are
// resident (griddepcontrol.launch_dependents at the top), (b) executes griddepcontrol.wait -- which returns only when
// the prerequisite grid has COMPLETED and its memory is visible -- before its first global-memory access.  So launch
// latency, CTA rasterisation and the per-CTA prologue (barrier init, TMEM allocation, descriptor prefetch) of kernel
// N+1 hide under the tail of kernel N, with unchanged memory semantics.  ACG_PDL=0 launches without the attribute.
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_prologue() {
    pdl_launch_dependents();
    pdl_wait();
}
// slim.batch_norm (scale=False, eps inside the sqrt, biased variance) from the per-channel totals; ONE definition so that
// the in-kernel finalize of the convolutions and the finalize inside the activation pass give the same bits.
__device__ __forceinline__ void bn_finalize_channel(double sum, double sumsq, double inv_rows, float eps, float beta,
                                                    float* mean, float* rstd, float* shift) {
    const double mu = sum * inv_rows;
    double var = sumsq * inv_rows - mu * mu;
    if (var < 0.0) var = 0.0;
    const float rs = (float)(1.0 / sqrt(var + (double)eps));
    *mean = (float)mu;
    *rstd = rs;
    *shift = beta - (float)mu * rs;
}
// value of one integer-limb accumulator triple (conv_tc.cuh fix_add): H * 2^24 + M * 2^-8 + L * 2^-40
__device__ __forceinline__ double limbs_to_double(unsigned long long h, unsigned long long m, unsigned long long l) {
    return (double)(long long)h * 0x1p24 + ((double)m * 0x1p-8 + (double)l * 0x1p-40);
}
bool pdl_enabled();

template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream,
                              Args&&... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = pdl_enabled() ? 1 : 0;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

// ---- mbarrier / bulk-copy (TMA) PTX wrappers ---------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async() {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.b32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
// Bounded wait: a lost arrive traps (kernel error) instead of hanging the GPU box.
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    uint32_t spins = 0;
    while (!mbar_try_wait(bar, parity)) {
        if (++spins > 200000000u) {
            printf("acg: mbarrier wait timed out (block %d thread %d)\n", blockIdx.x, threadIdx.x);
            __trap();
        }
    }
}
// 1-D bulk copy global -> shared (TMA engine, SASS UBLKCP); bytes % 16 == 0, 16 B aligned both sides.
__device__ __forceinline__ void bulk_g2s(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
            smem_u32(smem_dst)),
        "l"(gsrc), "r"(bytes), "r"(smem_u32(bar))
        : "memory");
}
// 1-D bulk copy shared -> global
__device__ __forceinline__ void bulk_s2g(void* gdst, const void* smem_src, uint32_t bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gdst),
                 "r"(smem_u32(smem_src)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void bulk_wait_read() {
    asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory");
}
template <int N> __device__ __forceinline__ void bulk_wait() {
    asm volatile("cp.async.bulk.wait_group %0;" ::"n"(N) : "memory");
}

}  // namespace acg
