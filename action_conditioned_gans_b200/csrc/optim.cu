// Fused optimizer updates over flat fp32 parameter buffers (one launch per network per step).
//
// Replaces tf.train.AdamOptimizer / tf.train.RMSPropOptimizer.minimize (train.py:91-102; one apply op
// per variable in TF) and the tf.clip_by_value assigns of train.py:89, with TF-1.0 numerics:
//   Adam   : m = b1 m + (1-b1) g; v = b2 v + (1-b2) g^2; p -= lr_t * m / (sqrt(v) + eps)
//            (lr_t = lr sqrt(1-b2^t)/(1-b1^t); eps is NOT bias corrected)
//   RMSProp: ms = d ms + (1-d) g^2 (ms starts at ONE); p -= lr * g / sqrt(ms + eps)
// then p = clip(p, lo, hi) when lo <= hi (update THEN clip; the reference leaves the order racy).
// HBM-bound: Adam 28 B/param, RMSProp 20 B/param, 128-bit accesses, grid sized to the SM count.
#include "common.cuh"

namespace acg {
namespace {

constexpr int kThreads = 256;

__device__ __forceinline__ float clipf(float p, float lo, float hi, bool do_clip) {
    return do_clip ? fminf(fmaxf(p, lo), hi) : p;
}

__global__ void __launch_bounds__(kThreads)
adam_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v,
            long long n, float lr_t, float b1, float b2, float eps, float lo, float hi, float gs,
            const float* __restrict__ lr_dev) {
    pdl_prologue();
    if (lr_dev) lr_t = *lr_dev;   // CUDA-graph replays: the bias-corrected rate changes every step
    const bool do_clip = lo <= hi;
    const long long n4 = n >> 2;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
        float4 pp = reinterpret_cast<float4*>(p)[i];
        float4 gg = reinterpret_cast<const float4*>(g)[i];
        float4 mm = reinterpret_cast<float4*>(m)[i];
        float4 vv = reinterpret_cast<float4*>(v)[i];
        float* pa = &pp.x; float* ga = &gg.x; float* ma = &mm.x; float* va = &vv.x;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const float gk = ga[k] * gs;
            ma[k] = b1 * ma[k] + (1.f - b1) * gk;
            va[k] = b2 * va[k] + (1.f - b2) * gk * gk;
            pa[k] = clipf(pa[k] - lr_t * ma[k] / (sqrtf(va[k]) + eps), lo, hi, do_clip);
        }
        reinterpret_cast<float4*>(p)[i] = pp;
        reinterpret_cast<float4*>(m)[i] = mm;
        reinterpret_cast<float4*>(v)[i] = vv;
    }
    // tail (n not a multiple of 4)
    for (long long i = (n4 << 2) + (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const float gk = g[i] * gs;
        const float mk = b1 * m[i] + (1.f - b1) * gk;
        const float vk = b2 * v[i] + (1.f - b2) * gk * gk;
        m[i] = mk;
        v[i] = vk;
        p[i] = clipf(p[i] - lr_t * mk / (sqrtf(vk) + eps), lo, hi, do_clip);
    }
}

__global__ void __launch_bounds__(kThreads)
rmsprop_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ ms, long long n, float lr,
               float decay, float eps, float lo, float hi, float gs, const float* __restrict__ lr_dev) {
    pdl_prologue();
    if (lr_dev) lr = *lr_dev;
    const bool do_clip = lo <= hi;
    const long long n4 = n >> 2;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
        float4 pp = reinterpret_cast<float4*>(p)[i];
        float4 gg = reinterpret_cast<const float4*>(g)[i];
        float4 ss = reinterpret_cast<float4*>(ms)[i];
        float* pa = &pp.x; float* ga = &gg.x; float* sa = &ss.x;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const float gk = ga[k] * gs;
            sa[k] = decay * sa[k] + (1.f - decay) * gk * gk;
            pa[k] = clipf(pa[k] - lr * gk / sqrtf(sa[k] + eps), lo, hi, do_clip);
        }
        reinterpret_cast<float4*>(p)[i] = pp;
        reinterpret_cast<float4*>(ms)[i] = ss;
    }
    for (long long i = (n4 << 2) + (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const float gk = g[i] * gs;
        const float sk = decay * ms[i] + (1.f - decay) * gk * gk;
        ms[i] = sk;
        p[i] = clipf(p[i] - lr * gk / sqrtf(sk + eps), lo, hi, do_clip);
    }
}

int grid_for(long long n) {
    long long blocks = (n / 4 + kThreads - 1) / kThreads;
    long long cap = (long long)num_sms() * 8;
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    return (int)blocks;
}

}  // namespace
}  // namespace acg

extern "C" {

int acg_adam_step(float* p, const float* g, float* m, float* v, long long n, float lr_t, float b1, float b2,
                  float eps, float clip_lo, float clip_hi, float grad_scale, const float* lr_t_dev, void* stream) {
    using namespace acg;
    ACG_REQUIRE(p && g && m && v, ACG_ERR_INVALID, "acg_adam_step: null pointer");
    ACG_REQUIRE(n > 0, ACG_ERR_INVALID, "acg_adam_step: n=%lld", n);
    ACG_REQUIRE((((uintptr_t)p | (uintptr_t)g | (uintptr_t)m | (uintptr_t)v) % 16) == 0, ACG_ERR_INVALID,
                "acg_adam_step: buffers must be 16-byte aligned");
    launch_pdl(adam_kernel, grid_for(n), kThreads, 0, static_cast<cudaStream_t>(stream), p, g, m, v, n, lr_t, b1, b2, eps,
                                                                                clip_lo, clip_hi, grad_scale, lr_t_dev);
    return check_launch("acg_adam_step");
}

int acg_rmsprop_step(float* p, const float* g, float* ms, long long n, float lr, float decay, float eps,
                     float clip_lo, float clip_hi, float grad_scale, const float* lr_dev, void* stream) {
    using namespace acg;
    ACG_REQUIRE(p && g && ms, ACG_ERR_INVALID, "acg_rmsprop_step: null pointer");
    ACG_REQUIRE(n > 0, ACG_ERR_INVALID, "acg_rmsprop_step: n=%lld", n);
    ACG_REQUIRE((((uintptr_t)p | (uintptr_t)g | (uintptr_t)ms) % 16) == 0, ACG_ERR_INVALID,
                "acg_rmsprop_step: buffers must be 16-byte aligned");
    launch_pdl(rmsprop_kernel, grid_for(n), kThreads, 0, static_cast<cudaStream_t>(stream), p, g, ms, n, lr, decay, eps,
                                                                                   clip_lo, clip_hi, grad_scale, lr_dev);
    return check_launch("acg_rmsprop_step");
}

}  // extern "C"
