// Loss kernels: forward value AND the gradient it implies, in one pass.
//
// Replaces ops.py:19-20 (build_psnr), ops.py:28-50 (build_g_adv_loss / build_d_loss),
// ops.py:100-120 (build_gdl: four 1x2 / 2x1 identity-channel convs + six elementwise passes) and the
// tf.norm terms of train.py:73,77 together with TF's autodiff of them.
#include "common.cuh"

namespace acg {
namespace {

constexpr int kThreads = 256;

__device__ __forceinline__ float sgn(float v) { return v > 0.f ? 1.f : (v < 0.f ? -1.f : 0.f); }

// q = d/d(dgen) | |dgt| - |dgen| |  (TF: abs' = sign, sign(0) = 0)
__device__ __forceinline__ float gdl_q(float dgt, float dgen) { return -sgn(fabsf(dgt) - fabsf(dgen)) * sgn(dgen); }

__global__ void __launch_bounds__(kThreads)
frame_losses_kernel(const float* __restrict__ g, const float* __restrict__ n, int B, int H, int W,
                    double* __restrict__ sums, float* __restrict__ dg, float w_l1, float w_gdl,
                    const float* __restrict__ dadv, int ld_adv, int adv_off) {
    pdl_prologue();
    const long long total = (long long)B * H * W * 3;
    const int rs = W * 3;  // row stride in floats
    float s_l1 = 0.f, s_sq = 0.f, s_gdl = 0.f;
    for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
         idx += (long long)gridDim.x * blockDim.x) {
        const int c = (int)(idx % 3);
        const long long px = idx / 3;
        const int j = (int)(px % W);
        const int i = (int)((px / W) % H);
        const float gc = g[idx], nc = n[idx];
        const float gr = (j + 1 < W) ? g[idx + 3] : 0.f, nr = (j + 1 < W) ? n[idx + 3] : 0.f;
        const float gd = (i + 1 < H) ? g[idx + rs] : 0.f, nd = (i + 1 < H) ? n[idx + rs] : 0.f;
        const float dxg = gr - gc, dxn = nr - nc;   // ops.py:112,114 (filter_x = [-1, +1], SAME pads after)
        const float dyg = gc - gd, dyn = nc - nd;   // ops.py:113,115 (filter_y = [+1; -1])
        const float diff = gc - nc;
        s_l1 += fabsf(diff);
        s_sq += diff * diff;
        s_gdl += fabsf(fabsf(dxn) - fabsf(dxg)) + fabsf(fabsf(dyn) - fabsf(dyg));
        if (dg) {
            float grad = w_l1 * sgn(diff);
            float q = -gdl_q(dxn, dxg) + gdl_q(dyn, dyg);
            if (j >= 1) {
                const float gl = g[idx - 3], nl = n[idx - 3];
                q += gdl_q(nc - nl, gc - gl);
            }
            if (i >= 1) {
                const float gu = g[idx - rs], nu = n[idx - rs];
                q -= gdl_q(nu - nc, gu - gc);
            }
            grad += w_gdl * q;
            if (dadv) grad += dadv[px * ld_adv + adv_off + c];
            dg[idx] = grad;
        }
    }
    __shared__ double red[3][kThreads / 32];
    double a = warp_sum((double)s_l1), b = warp_sum((double)s_sq), cc = warp_sum((double)s_gdl);
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    if (lane == 0) { red[0][wid] = a; red[1][wid] = b; red[2][wid] = cc; }
    __syncthreads();
    if (threadIdx.x < 3) {
        double t = 0.0;
        for (int w = 0; w < kThreads / 32; ++w) t += red[threadIdx.x][w];
        atomicAdd(&sums[threadIdx.x], t);
    }
}

__global__ void __launch_bounds__(kThreads)
dlogit_loss_kernel(const float* __restrict__ x, int n, int kind, float label_or_sign, float grad_scale,
                   float* __restrict__ loss_out, float* __restrict__ dlogits) {
    pdl_prologue();
    double acc = 0.0;
    const float inv_n = 1.f / (float)n;
    for (int i = threadIdx.x; i < n; i += kThreads) {
        const float v = x[i];
        float l, d;
        if (kind == ACG_LOSS_BCE) {
            // tf.losses.sigmoid_cross_entropy: max(x,0) - x*z + log1p(exp(-|x|))
            l = fmaxf(v, 0.f) - v * label_or_sign + log1pf(expf(-fabsf(v)));
            d = 1.f / (1.f + expf(-v)) - label_or_sign;
        } else {
            l = label_or_sign * v;
            d = label_or_sign;
        }
        acc += (double)l;
        if (dlogits) dlogits[i] = grad_scale * d * inv_n;
    }
    __shared__ double red[kThreads / 32];
    acc = warp_sum(acc);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0.0;
        for (int w = 0; w < kThreads / 32; ++w) t += red[w];
        loss_out[0] = (float)(t / (double)n);
    }
}

// sumsq_out != NULL: phase 1 of the data-parallel form -- only the LOCAL sum of squares is written (fp64), the caller
// sums it over the ranks.  sumsq_in != NULL: the norm is sqrt(*sumsq_in) (the GLOBAL sum) instead of the local one, so
// that loss and gradient are those of train.py:77's Frobenius norm over the whole batch.
__global__ void __launch_bounds__(kThreads)
state_loss_kernel(const float* __restrict__ s, const float* __restrict__ t, int n, float inv_batch,
                  float grad_scale, float* __restrict__ loss_out, float* __restrict__ ds,
                  double* __restrict__ sumsq_out, const double* __restrict__ sumsq_in) {
    pdl_prologue();
    __shared__ double red[kThreads / 32];
    __shared__ float norm_sh;
    if (!sumsq_in) {
        double acc = 0.0;
        for (int i = threadIdx.x; i < n; i += kThreads) {
            const double d = (double)s[i] - (double)t[i];
            acc += d * d;
        }
        acc = warp_sum(acc);
        if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        double tot = 0.0;
        if (sumsq_in) tot = *sumsq_in;
        else
            for (int w = 0; w < kThreads / 32; ++w) tot += red[w];
        if (sumsq_out) *sumsq_out = tot;
        const float norm = (float)sqrt(tot);
        norm_sh = norm;
        if (loss_out) loss_out[0] = norm * inv_batch;
    }
    __syncthreads();
    if (ds && !sumsq_out) {
        const float k = norm_sh > 0.f ? grad_scale * inv_batch / norm_sh : 0.f;
        for (int i = threadIdx.x; i < n; i += kThreads) ds[i] = k * (s[i] - t[i]);
    }
}

}  // namespace
}  // namespace acg

extern "C" {

int acg_frame_losses(const float* g, const float* n, int B, int H, int W, double* sums, float* dg, float w_l1,
                     float w_gdl, const float* dadv, int ld_adv, int adv_off, void* stream) {
    using namespace acg;
    ACG_REQUIRE(g && n && sums, ACG_ERR_INVALID, "acg_frame_losses: null pointer");
    ACG_REQUIRE(B > 0 && H > 0 && W > 0, ACG_ERR_INVALID, "acg_frame_losses: non-positive size");
    const long long total = (long long)B * H * W * 3;
    long long blocks = (total + kThreads - 1) / kThreads;
    const long long cap = (long long)num_sms() * 8;
    if (blocks > cap) blocks = cap;
    launch_pdl(frame_losses_kernel, (int)blocks, kThreads, 0, static_cast<cudaStream_t>(stream), g, n, B, H, W, sums, dg, w_l1, w_gdl, dadv, ld_adv, adv_off);
    return check_launch("acg_frame_losses");
}

int acg_dlogit_loss(const float* x, int n, int kind, float label_or_sign, float grad_scale, float* loss_out,
                    float* dlogits, void* stream) {
    using namespace acg;
    ACG_REQUIRE(x && loss_out, ACG_ERR_INVALID, "acg_dlogit_loss: null pointer");
    ACG_REQUIRE(n > 0, ACG_ERR_INVALID, "acg_dlogit_loss: n=%d", n);
    ACG_REQUIRE(kind == ACG_LOSS_BCE || kind == ACG_LOSS_WASS, ACG_ERR_INVALID, "unexpected loss argument");
    launch_pdl(dlogit_loss_kernel, 1, kThreads, 0, static_cast<cudaStream_t>(stream), x, n, kind, label_or_sign,
                                                                             grad_scale, loss_out, dlogits);
    return check_launch("acg_dlogit_loss");
}

int acg_state_loss(const float* s, const float* t, int n, float inv_batch, float grad_scale, float* loss_out,
                   float* dstate, double* sumsq_out, const double* sumsq_in, void* stream) {
    using namespace acg;
    ACG_REQUIRE(s && t && (loss_out || sumsq_out), ACG_ERR_INVALID, "acg_state_loss: null pointer");
    ACG_REQUIRE(!(sumsq_out && sumsq_in), ACG_ERR_INVALID, "acg_state_loss: sumsq_out and sumsq_in exclude each other");
    ACG_REQUIRE(n > 0, ACG_ERR_INVALID, "acg_state_loss: n=%d", n);
    launch_pdl(state_loss_kernel, 1, kThreads, 0, static_cast<cudaStream_t>(stream), s, t, n, inv_batch, grad_scale,
                                                                            loss_out, dstate, sumsq_out, sumsq_in);
    return check_launch("acg_state_loss");
}

}  // extern "C"
