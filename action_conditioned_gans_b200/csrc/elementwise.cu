// Batch-norm moments / apply / backward, activations, channel-slice copies and the action-tile concat.
//
// Replaces slim.batch_norm (arg_scope at models.py:11,32,81; slim defaults: batch statistics in train AND
// test, biased variance, epsilon 1e-3, beta only), tf.nn.relu / ops.lrelu (ops.py:22-26) / tf.tanh,
// tf.concat (models.py:16,38,84; train.py:64,68) and tf.tile (train.py:48-50), plus TF autodiff of them.
// All tensors are [rows, C] views of NHWC buffers with an explicit row stride so that producers can write
// straight into the wider concat buffers and consumers can read a channel slice.
#include "common.cuh"
#include "peer.cuh"

namespace acg {
namespace {

constexpr int kTX = 32, kTY = 8;

__device__ __forceinline__ float ldx(const void* p, int dt, size_t i) {
    return dt == ACG_F32 ? static_cast<const float*>(p)[i]
                         : __bfloat162float(static_cast<const __nv_bfloat16*>(p)[i]);
}
__device__ __forceinline__ void stx(void* p, int dt, size_t i, float v) {
    if (dt == ACG_F32) static_cast<float*>(p)[i] = v;
    else static_cast<__nv_bfloat16*>(p)[i] = __float2bfloat16_rn(v);
}

// column-wise two-moment reduction shared by bn_stats and bn_act_bwd_reduce.
// MODE 0: (z, z^2).  MODE 1: (dzh, dzh*xhat) with dzh = dA*act'(u).
template <int MODE>
__global__ void __launch_bounds__(kTX* kTY)
col_reduce_kernel(const void* __restrict__ z, int z_dt, int ld_z, const void* __restrict__ dA,
                  const void* __restrict__ dA2, int d_dt, int ld_d, long long rows_per_group, int C, const float* __restrict__ mean, const float* __restrict__ rstd,
                  const float* __restrict__ shift, int act, double* __restrict__ out) {
    pdl_prologue();
    const int c = blockIdx.x * kTX + threadIdx.x;
    const int g = blockIdx.z;
    const long long r_begin = (long long)g * rows_per_group;
    float s0 = 0.f, s1 = 0.f;
    double d0 = 0.0, d1 = 0.0;
    if (c < C) {
        float mu = 0.f, rs = 1.f, sh = 0.f;
        if (MODE == 1) {
            if (mean) mu = mean[g * C + c];
            if (rstd) rs = rstd[g * C + c];
            if (shift) sh = shift[g * C + c];
        }
        int k = 0;
        for (long long r = (long long)blockIdx.y * kTY + threadIdx.y; r < rows_per_group;
             r += (long long)gridDim.y * kTY) {
            const float zv = z ? ldx(z, z_dt, (size_t)(r_begin + r) * ld_z + c) : 0.f;
            if (MODE == 0) {
                s0 += zv;
                s1 += zv * zv;
            } else {
                const float u = zv * rs + sh;
                float dav = ldx(dA, d_dt, (size_t)(r_begin + r) * ld_d + c);
                if (dA2) dav += ldx(dA2, d_dt, (size_t)(r_begin + r) * ld_d + c);
                const float dzh = dav * act_bwd(u, act);
                s0 += dzh;
                s1 += dzh * ((zv - mu) * rs);
            }
            if (++k == 64) {  // flush fp32 partials into fp64 regularly
                d0 += s0; d1 += s1; s0 = s1 = 0.f; k = 0;
            }
        }
        d0 += s0;
        d1 += s1;
    }
    __shared__ double sh0[kTY][kTX], sh1[kTY][kTX];
    sh0[threadIdx.y][threadIdx.x] = d0;
    sh1[threadIdx.y][threadIdx.x] = d1;
    __syncthreads();
    if (threadIdx.y == 0 && c < C) {
        double t0 = 0.0, t1 = 0.0;
#pragma unroll
        for (int y = 0; y < kTY; ++y) { t0 += sh0[y][threadIdx.x]; t1 += sh1[y][threadIdx.x]; }
        atomicAdd(&out[(size_t)g * 2 * C + c], t0);
        atomicAdd(&out[(size_t)g * 2 * C + C + c], t1);
    }
}

__global__ void bn_finalize_kernel(const double* __restrict__ stats, const float* __restrict__ beta,
                                   long long rows_per_group, int C, int groups, float eps, float* __restrict__ mean,
                                   float* __restrict__ rstd, float* __restrict__ scale, float* __restrict__ shift) {
    pdl_prologue();
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= groups * C) return;
    const int g = i / C, c = i % C;
    const double inv = 1.0 / (double)rows_per_group;
    const double mu = stats[(size_t)g * 2 * C + c] * inv;
    double var = stats[(size_t)g * 2 * C + C + c] * inv - mu * mu;  // biased
    if (var < 0.0) var = 0.0;
    const float rs = (float)(1.0 / sqrt(var + (double)eps));
    const float b = beta ? beta[c] : 0.f;
    mean[i] = (float)mu;
    rstd[i] = rs;
    scale[i] = rs;
    shift[i] = b - (float)mu * rs;
}

__global__ void __launch_bounds__(256)
bn_act_fwd_kernel(const void* __restrict__ z, int z_dt, long long rows, int C, int ld_in, long long rows_per_group,
                  const float* __restrict__ scale, const float* __restrict__ shift, int act, void* __restrict__ out,
                  int o_dt, int ld_out) {
    pdl_prologue();
    const long long total = rows * C;
    for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
         idx += (long long)gridDim.x * blockDim.x) {
        const long long r = idx / C;
        const int c = (int)(idx - r * C);
        const int gc = (int)(r / rows_per_group) * C + c;
        float u = ldx(z, z_dt, (size_t)r * ld_in + c);
        if (scale) u *= scale[gc];
        if (shift) u += shift[gc];
        stx(out, o_dt, (size_t)r * ld_out + c, act_fwd(u, act));
    }
}

__global__ void __launch_bounds__(256)
bn_act_bwd_apply_kernel(const void* __restrict__ dA, const void* __restrict__ dA2, int d_dt, int ld_d, const void* __restrict__ z, int z_dt,
                        int ld_z, long long rows, int C, int groups, long long rows_per_group,
                        const float* __restrict__ mean, const float* __restrict__ rstd,
                        const float* __restrict__ shift, int act, int has_bn, const double* __restrict__ red,
                        void* __restrict__ dz, int dz_dt, int ld_dz, float* __restrict__ dbeta, long long norm_rows,
                        float dbeta_scale) {
    pdl_prologue();
    const long long total = rows * C;
    const float inv_r = 1.f / (float)norm_rows;
    for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
         idx += (long long)gridDim.x * blockDim.x) {
        const long long r = idx / C;
        const int c = (int)(idx - r * C);
        const int g = (int)(r / rows_per_group);
        const int gc = g * C + c;
        const float mu = mean ? mean[gc] : 0.f;
        const float rs = rstd ? rstd[gc] : 1.f;
        const float sh = shift ? shift[gc] : 0.f;
        const float zv = z ? ldx(z, z_dt, (size_t)r * ld_z + c) : 0.f;
        const float u = zv * rs + sh;
        float d = ldx(dA, d_dt, (size_t)r * ld_d + c);
        if (dA2) d += ldx(dA2, d_dt, (size_t)r * ld_d + c);
        d *= act_bwd(u, act);
        if (has_bn) {
            const float m0 = (float)red[(size_t)g * 2 * C + c] * inv_r;
            const float m1 = (float)red[(size_t)g * 2 * C + C + c] * inv_r;
            d = rs * (d - m0 - (zv - mu) * rs * m1);
        }
        stx(dz, dz_dt, (size_t)r * ld_dz + c, d);
    }
    if (dbeta && blockIdx.x == 0) {
        for (int c = threadIdx.x; c < C; c += blockDim.x) {
            double t = 0.0;
            for (int g = 0; g < groups; ++g) t += red[(size_t)g * 2 * C + c];
            atomicAdd(dbeta + c, dbeta_scale * (float)t);   // D(real) and D(generated) chains run concurrently
        }
    }
}

__global__ void __launch_bounds__(256)
copy_channels_kernel(const void* __restrict__ src, int s_dt, int ld_src, int off_src, void* __restrict__ dst,
                     int d_dt, int ld_dst, int off_dst, long long rows, int n) {
    pdl_prologue();
    const long long total = rows * n;
    for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
         idx += (long long)gridDim.x * blockDim.x) {
        const long long r = idx / n;
        const int c = (int)(idx - r * n);
        stx(dst, d_dt, (size_t)r * ld_dst + off_dst + c, ldx(src, s_dt, (size_t)r * ld_src + off_src + c));
    }
}

// One thread per pixel: dense fp32 frames a (and b) -> one bf16 row [a | b | zeros] of LD = 8 or 16 channels (one or
// two 16-byte stores; a warp writes 512 B / 1 KB contiguous); replaces two strided channel-slice copies per
// discriminator input.  LD = 8 keeps the first layers' implicit-GEMM K at 25 x 8 instead of 25 x 16.
template <int C, int LD>
__global__ void __launch_bounds__(256)
pack_frames_kernel(const float* __restrict__ a, const float* __restrict__ b, __nv_bfloat16* __restrict__ out,
                   long long rows) {
    pdl_prologue();
    for (long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x; r < rows;
         r += (long long)gridDim.x * blockDim.x) {
        float v[LD];
#pragma unroll
        for (int i = 0; i < LD; ++i) v[i] = 0.f;
#pragma unroll
        for (int c = 0; c < C; ++c) v[c] = a[r * C + c];
        if (b) {
#pragma unroll
            for (int c = 0; c < C; ++c) v[C + c] = b[r * C + c];
        }
        uint32_t w[LD / 2];
#pragma unroll
        for (int i = 0; i < LD / 2; ++i) {
            __nv_bfloat162 h = __floats2bfloat162_rn(v[2 * i], v[2 * i + 1]);
            w[i] = *reinterpret_cast<uint32_t*>(&h);
        }
        uint4* o = reinterpret_cast<uint4*>(out + r * LD);
#pragma unroll
        for (int i = 0; i < LD / 8; ++i) o[i] = make_uint4(w[4 * i], w[4 * i + 1], w[4 * i + 2], w[4 * i + 3]);
    }
}

__global__ void __launch_bounds__(256)
tile_actions_kernel(const float* __restrict__ actions, int B, int hw, int A, void* __restrict__ dst, int d_dt,
                    int ld_dst, int off) {
    pdl_prologue();
    const long long total = (long long)B * hw * A;
    for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
         idx += (long long)gridDim.x * blockDim.x) {
        const long long r = idx / A;
        const int a = (int)(idx - r * A);
        const int b = (int)(r / hw);
        stx(dst, d_dt, (size_t)r * ld_dst + off + a, actions[b * A + a]);
    }
}

__global__ void bias_grad_kernel(const double* __restrict__ red, int C, float scale, float* __restrict__ dbias) {
    pdl_prologue();
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c < C) atomicAdd(dbias + c, scale * (float)red[c]);
}

int ew_grid(long long total) {
    long long blocks = (total + 255) / 256;
    const long long cap = (long long)num_sms() * 16;
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    return (int)blocks;
}

dim3 reduce_grid(long long rows_per_group, int C, int groups) {
    const int gx = (C + kTX - 1) / kTX;
    long long gy = (rows_per_group + kTY * 8 - 1) / (kTY * 8);  // >= 8 rows per thread
    long long cap = ((long long)num_sms() * 8) / ((long long)gx * groups);
    if (cap < 1) cap = 1;
    if (gy > cap) gy = cap;
    if (gy < 1) gy = 1;
    return dim3(gx, (unsigned)gy, groups);
}

bool dt_ok(int dt) { return dt == ACG_F32 || dt == ACG_BF16; }

// ---- 8-channel vectorised variants (bf16 / fp32, 16-byte accesses) -----------------------------------------------
// Used whenever C % 8 == 0, every row stride is a multiple of 8 and the buffers are 16-byte aligned, i.e. for every
// layer of the tensor-core path; the scalar kernels above remain for ragged channel counts (3, 5, 25, 36, 138 ...).
struct F8 { float v[8]; };

__device__ __forceinline__ F8 load8(const void* p, int dt, size_t idx) {
    F8 r;
    if (dt == ACG_BF16) {
        const uint4 u = *reinterpret_cast<const uint4*>(static_cast<const __nv_bfloat16*>(p) + idx);
        const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            r.v[2 * i] = __uint_as_float(w[i] << 16);
            r.v[2 * i + 1] = __uint_as_float(w[i] & 0xffff0000u);
        }
    } else {
        const float4 a = reinterpret_cast<const float4*>(static_cast<const float*>(p) + idx)[0];
        const float4 b = reinterpret_cast<const float4*>(static_cast<const float*>(p) + idx)[1];
        r.v[0] = a.x; r.v[1] = a.y; r.v[2] = a.z; r.v[3] = a.w; r.v[4] = b.x; r.v[5] = b.y; r.v[6] = b.z; r.v[7] = b.w;
    }
    return r;
}
__device__ __forceinline__ void store8(void* p, int dt, size_t idx, const F8& f) {
    if (dt == ACG_BF16) {
        uint32_t w[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            __nv_bfloat162 h = __floats2bfloat162_rn(f.v[2 * i], f.v[2 * i + 1]);
            w[i] = *reinterpret_cast<uint32_t*>(&h);
        }
        *reinterpret_cast<uint4*>(static_cast<__nv_bfloat16*>(p) + idx) = make_uint4(w[0], w[1], w[2], w[3]);
    } else {
        float4* o = reinterpret_cast<float4*>(static_cast<float*>(p) + idx);
        o[0] = make_float4(f.v[0], f.v[1], f.v[2], f.v[3]);
        o[1] = make_float4(f.v[4], f.v[5], f.v[6], f.v[7]);
    }
}
__device__ __forceinline__ F8 load8f(const float* p, int c) {   // per-channel parameter vectors (may be NULL)
    F8 r;
    const float4 a = reinterpret_cast<const float4*>(p + c)[0], b = reinterpret_cast<const float4*>(p + c)[1];
    r.v[0] = a.x; r.v[1] = a.y; r.v[2] = a.z; r.v[3] = a.w; r.v[4] = b.x; r.v[5] = b.y; r.v[6] = b.z; r.v[7] = b.w;
    return r;
}

template <int MODE>
__global__ void __launch_bounds__(256, 3)
vec_col_reduce_kernel(const void* __restrict__ z, int z_dt, int ld_z, const void* __restrict__ dA,
                      const void* __restrict__ dA2, int d_dt, int ld_d, long long rows_per_group, int C,
                      const float* __restrict__ mean, const float* __restrict__ rstd, const float* __restrict__ shift,
                      int act, double* __restrict__ out) {
    pdl_prologue();
    __shared__ double sm[16][256];
    const int bx = blockDim.x, by = blockDim.y;
    const int tid = threadIdx.y * bx + threadIdx.x;
    const int nv = C >> 3;
    const int cv = blockIdx.x * bx + threadIdx.x;
    const int g = blockIdx.z;
    const long long r_begin = (long long)g * rows_per_group;
    // per-thread partial sums stay fp32 (a thread sees ~50 rows); everything across threads / blocks is fp64
    float s0[8], s1[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) { s0[i] = s1[i] = 0.f; }
    if (cv < nv) {
        const int c = cv * 8;
        F8 mu, rs, sh;
#pragma unroll
        for (int i = 0; i < 8; ++i) { mu.v[i] = 0.f; rs.v[i] = 1.f; sh.v[i] = 0.f; }
        if (MODE == 1) {
            if (mean) mu = load8f(mean + (size_t)g * C, c);
            if (rstd) rs = load8f(rstd + (size_t)g * C, c);
            if (shift) sh = load8f(shift + (size_t)g * C, c);
        }
        constexpr int U = MODE == 0 ? 4 : 2;     // independent row vectors in flight per thread
        const long long rstep = (long long)gridDim.y * by;
        for (long long r0 = (long long)blockIdx.y * by + threadIdx.y; r0 < rows_per_group; r0 += rstep * U) {
            F8 zv[U], da[U];
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const long long r = r0 + u * rstep;
                const bool ok = r < rows_per_group;
                if (z && ok) zv[u] = load8(z, z_dt, (size_t)(r_begin + r) * ld_z + c);
                else {
#pragma unroll
                    for (int i = 0; i < 8; ++i) zv[u].v[i] = 0.f;
                }
                if (MODE == 1) {
                    if (ok) {
                        da[u] = load8(dA, d_dt, (size_t)(r_begin + r) * ld_d + c);
                        if (dA2) {
                            const F8 db = load8(dA2, d_dt, (size_t)(r_begin + r) * ld_d + c);
#pragma unroll
                            for (int i = 0; i < 8; ++i) da[u].v[i] += db.v[i];
                        }
                    } else {
#pragma unroll
                        for (int i = 0; i < 8; ++i) da[u].v[i] = 0.f;
                    }
                }
            }
#pragma unroll
            for (int u = 0; u < U; ++u) {
                if (MODE == 0) {
#pragma unroll
                    for (int i = 0; i < 8; ++i) { s0[i] += zv[u].v[i]; s1[i] += zv[u].v[i] * zv[u].v[i]; }
                } else {
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        const float uu = zv[u].v[i] * rs.v[i] + sh.v[i];
                        const float dzh = da[u].v[i] * act_bwd(uu, act);
                        s0[i] += dzh;
                        s1[i] += dzh * ((zv[u].v[i] - mu.v[i]) * rs.v[i]);
                    }
                }
            }
        }
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) { sm[i][tid] = (double)s0[i]; sm[8 + i][tid] = (double)s1[i]; }
    __syncthreads();
    // thread (x, y) finalises values y, y+by, ... (of 16) of vector column x
    if (cv < nv) {
        for (int i = threadIdx.y; i < 16; i += by) {
            double t = 0.0;
            for (int y = 0; y < by; ++y) t += sm[i][y * bx + threadIdx.x];
            const int c = cv * 8 + (i & 7);
            atomicAdd(&out[(size_t)g * 2 * C + (i < 8 ? 0 : C) + c], t);
        }
    }
}

// Streaming kernels use the same 2-D thread map as the reductions: thread (x, y) owns the 8-channel vector column
// blockIdx.x*bx + x and walks rows blockIdx.y*by + y, + gridDim.y*by, ...  No index division in the loop, and the
// per-channel parameters live in registers for the whole kernel.
__global__ void __launch_bounds__(256)
vec_bn_act_fwd_kernel(const void* __restrict__ z, int z_dt, int C, int ld_in, long long rows_per_group,
                      const float* __restrict__ scale, const float* __restrict__ shift, int act, void* __restrict__ out,
                      int o_dt, int ld_out) {
    pdl_prologue();
    const int bx = blockDim.x, by = blockDim.y;
    const int cv = blockIdx.x * bx + threadIdx.x;
    if (cv >= (C >> 3)) return;
    const int c = cv * 8, g = blockIdx.z;
    F8 sc, sh;
#pragma unroll
    for (int i = 0; i < 8; ++i) { sc.v[i] = 1.f; sh.v[i] = 0.f; }
    if (scale) sc = load8f(scale + (size_t)g * C, c);
    if (shift) sh = load8f(shift + (size_t)g * C, c);
    const long long r_begin = (long long)g * rows_per_group;
    constexpr int U = 4;
    const long long rstep = (long long)gridDim.y * by;
    for (long long r0 = (long long)blockIdx.y * by + threadIdx.y; r0 < rows_per_group; r0 += rstep * U) {
        F8 u[U];
#pragma unroll
        for (int q = 0; q < U; ++q) {
            const long long r = r0 + q * rstep;
            if (r < rows_per_group) u[q] = load8(z, z_dt, (size_t)(r_begin + r) * ld_in + c);
        }
#pragma unroll
        for (int q = 0; q < U; ++q) {
            const long long r = r0 + q * rstep;
            if (r >= rows_per_group) break;
#pragma unroll
            for (int i = 0; i < 8; ++i) u[q].v[i] = act_fwd(fmaf(u[q].v[i], sc.v[i], sh.v[i]), act);
            store8(out, o_dt, (size_t)(r_begin + r) * ld_out + c, u[q]);
        }
    }
}

// dz = rs*dzh - rs*m0 - rs^2*m1*(z - mu) = rs*dzh + k1*z + k0 with per-channel k1 = -rs^2*m1, k0 = rs^2*m1*mu - rs*m0:
// the fp64 reduction results are folded into fp32 coefficients once per thread, the streaming loop is 3 FMAs/element.
__global__ void __launch_bounds__(256, 4)
vec_bn_act_bwd_apply_kernel(const void* __restrict__ dA, const void* __restrict__ dA2, int d_dt, int ld_d,
                            const void* __restrict__ z, int z_dt, int ld_z, int C, int groups,
                            long long rows_per_group, const float* __restrict__ mean, const float* __restrict__ rstd,
                            const float* __restrict__ shift, int act, int has_bn, const double* __restrict__ red,
                            void* __restrict__ dz, int dz_dt, int ld_dz, float* __restrict__ dbeta, long long norm_rows,
                            float dbeta_scale) {
    pdl_prologue();
    __shared__ float4 coef[256];               // this block's <= 32 vector columns x 8 channels: rs, sh, k1, k0
    const int bx = blockDim.x, by = blockDim.y;
    const int tid = threadIdx.y * bx + threadIdx.x;
    const int cv = blockIdx.x * bx + threadIdx.x;
    const int g = blockIdx.z;
    if (dbeta && blockIdx.x == 0 && blockIdx.y == 0 && g == 0) {
        for (int c = tid; c < C; c += bx * by) {
            double t = 0.0;
            for (int gg = 0; gg < groups; ++gg) t += red[(size_t)gg * 2 * C + c];
            atomicAdd(dbeta + c, dbeta_scale * (float)t);   // D(real) and D(generated) chains run concurrently
        }
    }
    if (tid < bx * 8) {
        const int c = blockIdx.x * bx * 8 + tid;
        float4 k = make_float4(1.f, 0.f, 0.f, 0.f);
        if (c < C) {
            const size_t gc = (size_t)g * C + c;
            const float r = rstd ? rstd[gc] : 1.f, mu = mean ? mean[gc] : 0.f;
            k.x = r;
            k.y = shift ? shift[gc] : 0.f;
            if (has_bn) {
                const double inv_r = 1.0 / (double)norm_rows;
                const double m0 = red[(size_t)g * 2 * C + c] * inv_r, m1 = red[(size_t)g * 2 * C + C + c] * inv_r;
                k.z = (float)(-(double)r * r * m1);
                k.w = (float)((double)r * r * m1 * mu - (double)r * m0);
            }
        }
        coef[tid] = k;
    }
    __syncthreads();
    if (cv >= (C >> 3)) return;
    const int c = cv * 8;
    const float4* cf = coef + threadIdx.x * 8;
    const long long r_begin = (long long)g * rows_per_group;
    constexpr int U = 2;
    const long long rstep = (long long)gridDim.y * by;
    for (long long r0 = (long long)blockIdx.y * by + threadIdx.y; r0 < rows_per_group; r0 += rstep * U) {
        F8 zv[U], d[U];
#pragma unroll
        for (int q = 0; q < U; ++q) {
            const long long r = r0 + q * rstep;
            const bool ok = r < rows_per_group;
            if (z && ok) zv[q] = load8(z, z_dt, (size_t)(r_begin + r) * ld_z + c);
            else {
#pragma unroll
                for (int i = 0; i < 8; ++i) zv[q].v[i] = 0.f;
            }
            if (ok) {
                d[q] = load8(dA, d_dt, (size_t)(r_begin + r) * ld_d + c);
                if (dA2) {
                    const F8 d2 = load8(dA2, d_dt, (size_t)(r_begin + r) * ld_d + c);
#pragma unroll
                    for (int i = 0; i < 8; ++i) d[q].v[i] += d2.v[i];
                }
            }
        }
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const float4 k = cf[i];                       // rs, sh, k1, k0
#pragma unroll
            for (int q = 0; q < U; ++q) {
                const float zz = zv[q].v[i];
                const float dzh = d[q].v[i] * act_bwd(fmaf(zz, k.x, k.y), act);
                d[q].v[i] = has_bn ? fmaf(k.x, dzh, fmaf(k.z, zz, k.w)) : dzh;
            }
        }
#pragma unroll
        for (int q = 0; q < U; ++q) {
            const long long r = r0 + q * rstep;
            if (r < rows_per_group) store8(dz, dz_dt, (size_t)(r_begin + r) * ld_dz + c, d[q]);
        }
    }
}

// ---- bf16 fast path of the backward batch-norm kernels ------------------------------------------------------------
// Every layer of the product path has bf16 z / dA / dz and relu, lrelu or no activation, so dtype and activation are
// template parameters here (the runtime-dispatched kernels above kept their F8 arrays in local memory: 136 B stack,
// 1.8-2.6 TB/s).  Loads are raw 16-byte vectors, FOUR rows in flight per thread before the first use, and the grid is
// sized so that a thread of a small layer walks its rows in ONE pass (the old ">= 16 rows per thread" rule serialised
// ~8 dependent memory latencies: 13 us for a 64 KB tensor).
template <int ACT> __device__ __forceinline__ float act_bwd_t(float u) {
    if (ACT == ACG_ACT_RELU) return u > 0.f ? 1.f : 0.f;
    if (ACT == ACG_ACT_LRELU) return 0.6f + 0.4f * (u > 0.f ? 1.f : (u < 0.f ? -1.f : 0.f));
    return 1.f;
}
__device__ __forceinline__ void unpack8(const uint4& u, float (&f)[8]) {
    const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        f[2 * i] = __uint_as_float(w[i] << 16);
        f[2 * i + 1] = __uint_as_float(w[i] & 0xffff0000u);
    }
}
__device__ __forceinline__ uint4 ldg16(const __nv_bfloat16* p) { return __ldg(reinterpret_cast<const uint4*>(p)); }

constexpr int kFastU = 4;

template <int ACT, bool HAS_Z, bool HAS_D2>
__global__ void __launch_bounds__(256)
fast_bwd_reduce_kernel(const __nv_bfloat16* __restrict__ z, int ld_z, const __nv_bfloat16* __restrict__ dA,
                       const __nv_bfloat16* __restrict__ dA2, int ld_d, long long rows, int C,
                       const float* __restrict__ mean, const float* __restrict__ rstd, const float* __restrict__ shift,
                       double* __restrict__ out, unsigned int* __restrict__ counter, PeerExchange px, int rev) {
    // counter != NULL (data parallel): the LAST block sums `out` over the ranks itself (peer.cuh) -- no exchange launch
    // between this pass and the apply pass.  Such a launch must not trigger its dependents early (see peer.cu).
    if (!counter) pdl_launch_dependents();
    pdl_wait();
    __shared__ float sm[16][257];
    __shared__ int last_block_sh;
    const int bx = blockDim.x, by = blockDim.y;
    const int tid = threadIdx.y * bx + threadIdx.x;
    const int nv = C >> 3;
    const int cv = blockIdx.x * bx + threadIdx.x;
    float s0[8], s1[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) { s0[i] = s1[i] = 0.f; }
    if (cv < nv) {
        const int c = cv * 8;
        float mu[8], rs[8], sh[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) { mu[i] = 0.f; rs[i] = 1.f; sh[i] = 0.f; }
        if (HAS_Z) {
            if (mean) { const F8 t = load8f(mean, c);
#pragma unroll
                for (int i = 0; i < 8; ++i) mu[i] = t.v[i]; }
            if (rstd) { const F8 t = load8f(rstd, c);
#pragma unroll
                for (int i = 0; i < 8; ++i) rs[i] = t.v[i]; }
            if (shift) { const F8 t = load8f(shift, c);
#pragma unroll
                for (int i = 0; i < 8; ++i) sh[i] = t.v[i]; }
        }
        const long long rstep = (long long)gridDim.y * by;
        for (long long r0 = (long long)blockIdx.y * by + threadIdx.y; r0 < rows; r0 += rstep * kFastU) {
            uint4 zr[kFastU], dr[kFastU], d2r[kFastU];
#pragma unroll
            for (int u = 0; u < kFastU; ++u) {
                const long long r = r0 + u * rstep;
                const long long ra = rev ? rows - 1 - r : r;      // traversal direction (see ew_rev())
                const bool ok = r < rows;
                zr[u] = make_uint4(0u, 0u, 0u, 0u); dr[u] = zr[u]; d2r[u] = zr[u];
                if (ok) {
                    if (HAS_Z) zr[u] = ldg16(z + (size_t)ra * ld_z + c);
                    dr[u] = ldg16(dA + (size_t)ra * ld_d + c);
                    if (HAS_D2) d2r[u] = ldg16(dA2 + (size_t)ra * ld_d + c);
                }
            }
#pragma unroll
            for (int u = 0; u < kFastU; ++u) {
                float zf[8], df[8];
                unpack8(dr[u], df);
                if (HAS_D2) {
                    float d2[8];
                    unpack8(d2r[u], d2);
#pragma unroll
                    for (int i = 0; i < 8; ++i) df[i] += d2[i];
                }
                if (HAS_Z) {
                    unpack8(zr[u], zf);
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        const float uu = fmaf(zf[i], rs[i], sh[i]);
                        const float dzh = df[i] * act_bwd_t<ACT>(uu);
                        s0[i] += dzh;
                        s1[i] += dzh * ((zf[i] - mu[i]) * rs[i]);
                    }
                } else {
#pragma unroll
                    for (int i = 0; i < 8; ++i) s0[i] += df[i];
                }
            }
        }
    }
    // block reduction over the by row slots (fp32: <= 128 partials of <= ~50 rows each), then fp64 across blocks
#pragma unroll
    for (int i = 0; i < 8; ++i) { sm[i][tid] = s0[i]; sm[8 + i][tid] = s1[i]; }
    __syncthreads();
    if (cv < nv) {
        for (int i = threadIdx.y; i < 16; i += by) {
            float t = 0.f;
            for (int y = 0; y < by; ++y) t += sm[i][y * bx + threadIdx.x];
            const int c = cv * 8 + (i & 7);
            atomicAdd(&out[(i < 8 ? 0 : C) + c], (double)t);
        }
    }
    if (counter) {
        __threadfence();
        __syncthreads();
        if (tid == 0) last_block_sh = (atomicAdd(counter, 1u) == gridDim.x * gridDim.y - 1u);
        __syncthreads();
        if (last_block_sh) {
            __threadfence();
            peer_exchange(px, out, 2 * C, tid, bx * by, [] {});
            if (tid == 0) *counter = 0u;          // ready for the next launch
        }
    }
}

template <int ACT, bool HAS_BN, bool HAS_D2>
__global__ void __launch_bounds__(256)
fast_bwd_apply_kernel(const __nv_bfloat16* __restrict__ dA, const __nv_bfloat16* __restrict__ dA2, int ld_d,
                      const __nv_bfloat16* __restrict__ z, int ld_z, int C, long long rows,
                      const float* __restrict__ mean, const float* __restrict__ rstd, const float* __restrict__ shift,
                      const double* __restrict__ red, __nv_bfloat16* __restrict__ dz, int ld_dz,
                      float* __restrict__ dbeta, long long norm_rows, float dbeta_scale, int rev) {
    pdl_prologue();
    __shared__ float4 coef[256];               // this block's <= 32 vector columns x 8 channels: rs, sh, k1, k0
    const int bx = blockDim.x, by = blockDim.y;
    const int tid = threadIdx.y * bx + threadIdx.x;
    const int cv = blockIdx.x * bx + threadIdx.x;
    if (dbeta && blockIdx.x == 0 && blockIdx.y == 0) {
        for (int c = tid; c < C; c += bx * by) atomicAdd(dbeta + c, dbeta_scale * (float)red[c]);
    }
    if (tid < bx * 8) {
        const int c = blockIdx.x * bx * 8 + tid;
        float4 k = make_float4(1.f, 0.f, 0.f, 0.f);
        if (c < C) {
            const float r = rstd ? rstd[c] : 1.f, mu = mean ? mean[c] : 0.f;
            k.x = r;
            k.y = shift ? shift[c] : 0.f;
            if (HAS_BN) {
                const double inv_r = 1.0 / (double)norm_rows;
                const double m0 = red[c] * inv_r, m1 = red[C + c] * inv_r;
                k.z = (float)(-(double)r * r * m1);
                k.w = (float)((double)r * r * m1 * mu - (double)r * m0);
            }
        }
        coef[tid] = k;
    }
    __syncthreads();
    if (cv >= (C >> 3)) return;
    const int c = cv * 8;
    float4 cf[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) cf[i] = coef[threadIdx.x * 8 + i];
    const long long rstep = (long long)gridDim.y * by;
    for (long long r0 = (long long)blockIdx.y * by + threadIdx.y; r0 < rows; r0 += rstep * kFastU) {
        uint4 zr[kFastU], dr[kFastU], d2r[kFastU];
#pragma unroll
        for (int u = 0; u < kFastU; ++u) {
            const long long r = r0 + u * rstep;
                const long long ra = rev ? rows - 1 - r : r;      // traversal direction (see ew_rev())
            zr[u] = make_uint4(0u, 0u, 0u, 0u); dr[u] = zr[u]; d2r[u] = zr[u];
            if (r < rows) {
                zr[u] = ldg16(z + (size_t)ra * ld_z + c);
                dr[u] = ldg16(dA + (size_t)ra * ld_d + c);
                if (HAS_D2) d2r[u] = ldg16(dA2 + (size_t)ra * ld_d + c);
            }
        }
#pragma unroll
        for (int u = 0; u < kFastU; ++u) {
            const long long r = r0 + u * rstep;
                const long long ra = rev ? rows - 1 - r : r;      // traversal direction (see ew_rev())
            if (r >= rows) break;
            float zf[8], df[8];
            unpack8(zr[u], zf);
            unpack8(dr[u], df);
            if (HAS_D2) {
                float d2[8];
                unpack8(d2r[u], d2);
#pragma unroll
                for (int i = 0; i < 8; ++i) df[i] += d2[i];
            }
            uint32_t w[4];
#pragma unroll
            for (int i = 0; i < 8; i += 2) {
                float o[2];
#pragma unroll
                for (int j = 0; j < 2; ++j) {
                    const float4 k = cf[i + j];
                    const float zz = zf[i + j];
                    const float dzh = df[i + j] * act_bwd_t<ACT>(fmaf(zz, k.x, k.y));
                    o[j] = HAS_BN ? fmaf(k.x, dzh, fmaf(k.z, zz, k.w)) : dzh;
                }
                __nv_bfloat162 h = __floats2bfloat162_rn(o[0], o[1]);
                w[i >> 1] = *reinterpret_cast<uint32_t*>(&h);
            }
            *reinterpret_cast<uint4*>(dz + (size_t)ra * ld_dz + c) = make_uint4(w[0], w[1], w[2], w[3]);
        }
    }
}

template <int ACT> __device__ __forceinline__ float act_fwd_t(float u) {
    if (ACT == ACG_ACT_RELU) return fmaxf(u, 0.f);
    if (ACT == ACG_ACT_LRELU) return 0.6f * u + 0.4f * fabsf(u);
    return u;
}

// RAW moments of a convolution launch (acg_bn_finalize_act_fwd): fp64 totals [2][C] (+ optional integer limb accumulators
// [3][2][C], conv_tc.cuh fix_add) that NO kernel has completed yet.  Every block of the activation pass completes the
// totals of ITS channels itself (<= 256 channels: one thread each) -- that replaces the convolution's tail of fence,
// ticket atomic, last-CTA conversion and finalize (~2-3 us at the end of every launch with fused moments), and the
// blocks of the first row block export mean / rstd / scale / shift for the backward pass.
struct RawMoments {
    const double* stats;                 // [2][C]; zero unless a non-finite value was added
    const unsigned long long* fix;       // [3][2][C] or NULL
    const float* beta;                   // [C] or NULL
    double inv_rows;
    float eps;
    float* mean; float* rstd; float* scale; float* shift;     // exported [C]
};

// a[r] = act(z[r]*scale + shift), bf16 -> bf16, same thread map / row pipelining as the backward kernels
template <int ACT, bool RAW>
__global__ void __launch_bounds__(256)
fast_bn_act_fwd_kernel(const __nv_bfloat16* __restrict__ z, int ld_in, int C, long long rows,
                       const float* __restrict__ scale, const float* __restrict__ shift,
                       __nv_bfloat16* __restrict__ out, int ld_out, int rev,
                       const float* __restrict__ acts, int n_act, int hw, int act_off, RawMoments rm) {
    pdl_prologue();
    const int bx = blockDim.x, by = blockDim.y;
    const int cv = blockIdx.x * bx + threadIdx.x;
    __shared__ float s_sc[RAW ? 256 : 1], s_sh[RAW ? 256 : 1];
    if (RAW) {
        const int tid = threadIdx.y * bx + threadIdx.x;
        const int ch = blockIdx.x * bx * 8 + tid;
        if (tid < bx * 8 && ch < C) {
            double s0 = __ldcg(rm.stats + ch), s1 = __ldcg(rm.stats + C + ch);
            if (rm.fix) {
                const unsigned long long* a = rm.fix + ch;
                s0 += limbs_to_double(__ldcg(a), __ldcg(a + 2 * C), __ldcg(a + 4 * C));
                s1 += limbs_to_double(__ldcg(a + C), __ldcg(a + 3 * C), __ldcg(a + 5 * C));
            }
            float mu, rs, sh_;
            bn_finalize_channel(s0, s1, rm.inv_rows, rm.eps, rm.beta ? rm.beta[ch] : 0.f, &mu, &rs, &sh_);
            s_sc[tid] = rs;
            s_sh[tid] = sh_;
            if (blockIdx.y == 0) { rm.mean[ch] = mu; rm.rstd[ch] = rs; rm.scale[ch] = rs; rm.shift[ch] = sh_; }
        }
        __syncthreads();
    }
    if (cv >= (C >> 3)) return;
    const int c = cv * 8;
    float sc[8], sh[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) { sc[i] = 1.f; sh[i] = 0.f; }
    if (RAW) {
#pragma unroll
        for (int i = 0; i < 8; ++i) { sc[i] = s_sc[threadIdx.x * 8 + i]; sh[i] = s_sh[threadIdx.x * 8 + i]; }
    } else {
    if (scale) { const F8 t = load8f(scale, c);
#pragma unroll
        for (int i = 0; i < 8; ++i) sc[i] = t.v[i]; }
    if (shift) { const F8 t = load8f(shift, c);
#pragma unroll
        for (int i = 0; i < 8; ++i) sh[i] = t.v[i]; }
    }
    const long long rstep = (long long)gridDim.y * by;
    for (long long r0 = (long long)blockIdx.y * by + threadIdx.y; r0 < rows; r0 += rstep * kFastU) {
        uint4 zr[kFastU];
#pragma unroll
        for (int u = 0; u < kFastU; ++u) {
            const long long r = r0 + u * rstep;
            const long long ra = rev ? rows - 1 - r : r;
            zr[u] = make_uint4(0u, 0u, 0u, 0u);
            if (r < rows) zr[u] = ldg16(z + (size_t)ra * ld_in + c);
        }
#pragma unroll
        for (int u = 0; u < kFastU; ++u) {
            const long long r = r0 + u * rstep;
            const long long ra = rev ? rows - 1 - r : r;
            if (r >= rows) break;
            float zf[8];
            unpack8(zr[u], zf);
            uint32_t w[4];
#pragma unroll
            for (int i = 0; i < 8; i += 2) {
                __nv_bfloat162 h = __floats2bfloat162_rn(act_fwd_t<ACT>(fmaf(zf[i], sc[i], sh[i])),
                                                         act_fwd_t<ACT>(fmaf(zf[i + 1], sc[i + 1], sh[i + 1])));
                w[i >> 1] = *reinterpret_cast<uint32_t*>(&h);
            }
            *reinterpret_cast<uint4*>(out + (size_t)ra * ld_out + c) = make_uint4(w[0], w[1], w[2], w[3]);
            // concat of the tiled action vector (models.py:16,38,84): the thread that owns the row's first channel
            // vector also writes the n_act action channels behind the features -- no acg_tile_actions launch
            if (acts && cv == 0) {
                const float* a = acts + (size_t)(ra / hw) * n_act;
                __nv_bfloat16* o = out + (size_t)ra * ld_out + act_off;
                for (int j = 0; j < n_act; ++j) o[j] = __float2bfloat16_rn(a[j]);
            }
        }
    }
}

// grid for the fast kernels: thread (x, y) = (8-channel vector column, row slot); at most `blocks_per_sm` resident
// blocks per SM in ONE wave, and never more blocks than one pass of kFastU rows per thread needs.  The reduction
// additionally pays 16*bx fp64 atomics per block (measured ~13 G atomics/s on the 2C hot addresses), so for it the
// row-block count minimises  passes x ~0.8 us (one dependent memory round trip each)  +  atomics / 13e3 us.
void fast_launch_dims(long long rows, int C, int blocks_per_sm, bool atomics, dim3* grid, dim3* block) {
    const int nv = C >> 3;
    int bx = 1;
    while (bx < nv && bx < 32) bx <<= 1;
    const int by = 256 / bx;
    const int gx = (nv + bx - 1) / bx;
    const long long per_pass = (long long)by * kFastU;
    long long gy = (rows + per_pass - 1) / per_pass;
    long long cap = ((long long)num_sms() * blocks_per_sm) / gx;
    if (cap < 1) cap = 1;
    if (gy > cap) gy = cap;
    if (gy < 1) gy = 1;
    if (atomics) {
        double best = 1e30;
        long long best_gy = gy;
        for (long long g = gy; g >= 1; g = (g + 1) / 2) {
            const double passes = (double)((rows + g * per_pass - 1) / (g * per_pass));
            const double t = passes * 0.8 + (double)(g * gx) * 16.0 * bx / 13000.0;
            if (t < best) { best = t; best_gy = g; }
            if (g == 1) break;
        }
        gy = best_gy;
    }
    *grid = dim3(gx, (unsigned)gy, 1);
    *block = dim3(bx, by);
}

// Row traversal direction of the streaming passes (bit 1: forward apply, 2: backward reduction, 4: backward apply run
// from the LAST row to the first).  A pass that follows a kernel which streamed the same tensors first-to-last finds the
// most recently touched rows still in the 126 MB L2 when it starts from the end.  ACG_EW_REV overrides the default.
int ew_rev() {
    static const int v = [] { const char* e = getenv("ACG_EW_REV"); return e ? atoi(e) : 0; }();
    return v;
}

template <int ACT, bool HAS_Z>
void launch_fast_reduce(dim3 grid, dim3 block, cudaStream_t st, const void* z, int ld_z, const void* dA, const void* dA2,
                        int ld_d, long long rows, int C, const float* mean, const float* rstd, const float* shift,
                        double* red, unsigned int* counter = nullptr, const PeerExchange* px = nullptr) {
    const __nv_bfloat16* zz = static_cast<const __nv_bfloat16*>(z);
    const __nv_bfloat16* d1 = static_cast<const __nv_bfloat16*>(dA);
    const __nv_bfloat16* d2 = static_cast<const __nv_bfloat16*>(dA2);
    PeerExchange x{};
    if (px) x = *px;
    if (dA2) launch_pdl(fast_bwd_reduce_kernel<ACT, HAS_Z, true>, grid, block, 0, st, zz, ld_z, d1, d2, ld_d, rows, C, mean, rstd, shift, red, counter, x, ew_rev() & 2);
    else launch_pdl(fast_bwd_reduce_kernel<ACT, HAS_Z, false>, grid, block, 0, st, zz, ld_z, d1, d2, ld_d, rows, C, mean, rstd, shift, red, counter, x, ew_rev() & 2);
}

template <int ACT, bool HAS_BN>
void launch_fast_apply(dim3 grid, dim3 block, cudaStream_t st, const void* dA, const void* dA2, int ld_d, const void* z,
                       int ld_z, int C, long long rows, const float* mean, const float* rstd, const float* shift,
                       const double* red, void* dz, int ld_dz, float* dbeta, long long norm_rows, float dbeta_scale) {
    const __nv_bfloat16* zz = static_cast<const __nv_bfloat16*>(z);
    const __nv_bfloat16* d1 = static_cast<const __nv_bfloat16*>(dA);
    const __nv_bfloat16* d2 = static_cast<const __nv_bfloat16*>(dA2);
    __nv_bfloat16* o = static_cast<__nv_bfloat16*>(dz);
    if (dA2) launch_pdl(fast_bwd_apply_kernel<ACT, HAS_BN, true>, grid, block, 0, st, d1, d2, ld_d, zz, ld_z, C, rows, mean, rstd, shift, red, o, ld_dz, dbeta, norm_rows, dbeta_scale, ew_rev() & 4);
    else launch_pdl(fast_bwd_apply_kernel<ACT, HAS_BN, false>, grid, block, 0, st, d1, d2, ld_d, zz, ld_z, C, rows, mean, rstd, shift, red, o, ld_dz, dbeta, norm_rows, dbeta_scale, ew_rev() & 4);
}

// 2-D launch shape shared by the streaming kernels: `waves` resident blocks per SM
void vec_stream_launch_dims(long long rows_per_group, int C, int groups, int blocks_per_sm, dim3* grid, dim3* block) {
    const int nv = C >> 3;
    int bx = 1;
    while (bx < nv && bx < 32) bx <<= 1;
    const int by = 256 / bx;
    const int gx = (nv + bx - 1) / bx;
    long long gy = (rows_per_group + (long long)by * 8 - 1) / ((long long)by * 8);
    long long cap = ((long long)num_sms() * blocks_per_sm) / ((long long)gx * groups);
    if (cap < 1) cap = 1;
    if (gy > cap) gy = cap;
    if (gy < 1) gy = 1;
    *grid = dim3(gx, (unsigned)gy, groups);
    *block = dim3(bx, by);
}

bool al16(const void* p) { return p == nullptr || ((uintptr_t)p & 15) == 0; }

void vec_reduce_launch_dims(long long rows_per_group, int C, int groups, dim3* grid, dim3* block) {
    const int nv = C >> 3;
    int bx = 1;
    while (bx < nv && bx < 32) bx <<= 1;
    const int by = 256 / bx;
    const int gx = (nv + bx - 1) / bx;
    long long gy = (rows_per_group + (long long)by * 16 - 1) / ((long long)by * 16);   // >= 16 rows per thread
    // two resident 256-thread blocks per SM (register limited): one wave, so the per-block tree reduction and the
    // fp64 atomics are amortised over ~50 rows per thread instead of ~14
    long long cap = ((long long)num_sms() * 3) / ((long long)gx * groups);
    if (cap < 1) cap = 1;
    if (gy > cap) gy = cap;
    if (gy < 1) gy = 1;
    *grid = dim3(gx, (unsigned)gy, groups);
    *block = dim3(bx, by);
}

int vec_ew_grid(long long total_vec) {
    long long blocks = (total_vec + 255) / 256;
    const long long cap = (long long)num_sms() * 16;
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    return (int)blocks;
}


}  // namespace
}  // namespace acg

extern "C" {

int acg_bn_stats(const void* z, int dtype, long long rows, int C, int ld, int groups, double* stats, void* stream) {
    using namespace acg;
    ACG_REQUIRE(z && stats, ACG_ERR_INVALID, "acg_bn_stats: null pointer");
    ACG_REQUIRE(rows > 0 && C > 0 && ld >= C && groups > 0 && rows % groups == 0, ACG_ERR_INVALID,
                "acg_bn_stats: rows=%lld C=%d ld=%d groups=%d", rows, C, ld, groups);
    ACG_REQUIRE(dt_ok(dtype), ACG_ERR_UNSUPPORTED, "acg_bn_stats: dtype %d", dtype);
    const long long rpg = rows / groups;
    if (C % 8 == 0 && ld % 8 == 0 && al16(z)) {
        dim3 grid, block;
        vec_reduce_launch_dims(rpg, C, groups, &grid, &block);
        launch_pdl(vec_col_reduce_kernel<0>, grid, block, 0, static_cast<cudaStream_t>(stream), z, dtype, ld, nullptr, nullptr, 0, 0, rpg, C, nullptr, nullptr, nullptr, 0, stats);
        return check_launch("acg_bn_stats");
    }
    launch_pdl(col_reduce_kernel<0>, reduce_grid(rpg, C, groups), dim3(kTX, kTY), 0, static_cast<cudaStream_t>(stream), z, dtype, ld, nullptr, nullptr, 0, 0, rpg, C, nullptr, nullptr, nullptr, 0, stats);
    return check_launch("acg_bn_stats");
}

int acg_bn_finalize(const double* stats, const float* beta, long long rows_per_group, int C, int groups, float eps,
                    float* mean, float* rstd, float* scale, float* shift, void* stream) {
    using namespace acg;
    ACG_REQUIRE(stats && mean && rstd && scale && shift, ACG_ERR_INVALID, "acg_bn_finalize: null pointer");
    ACG_REQUIRE(rows_per_group > 0 && C > 0 && groups > 0, ACG_ERR_INVALID, "acg_bn_finalize: bad size");
    const int n = C * groups;
    launch_pdl(bn_finalize_kernel, (n + 127) / 128, 128, 0, static_cast<cudaStream_t>(stream), stats, beta, rows_per_group, C, groups, eps, mean, rstd, scale, shift);
    return check_launch("acg_bn_finalize");
}

static int bn_act_fwd_impl(const void* z, int z_dtype, long long rows, int C, int ld_in, int groups, const float* scale,
                           const float* shift, int act, void* out, int out_dtype, int ld_out, const float* acts,
                           int n_act, int hw, int act_off, void* stream, const char* who) {
    using namespace acg;
    ACG_REQUIRE(z && out, ACG_ERR_INVALID, "%s: null pointer", who);
    ACG_REQUIRE(rows > 0 && C > 0 && ld_in >= C && ld_out >= C && groups > 0 && rows % groups == 0,
                ACG_ERR_INVALID, "%s: bad size", who);
    ACG_REQUIRE(dt_ok(z_dtype) && dt_ok(out_dtype), ACG_ERR_UNSUPPORTED, "%s: dtype", who);
    ACG_REQUIRE(!acts || (n_act > 0 && hw > 0 && rows % hw == 0 && act_off >= C && act_off + n_act <= ld_out),
                ACG_ERR_INVALID, "%s: action concat: %d channels at offset %d of %d, %d pixels per image", who, n_act,
                act_off, ld_out, hw);
    if (C % 8 == 0 && ld_in % 8 == 0 && ld_out % 8 == 0 && al16(z) && al16(out) && al16(scale) && al16(shift)) {
        dim3 grid, block;
        if (groups == 1 && z_dtype == ACG_BF16 && out_dtype == ACG_BF16 &&
            (act == ACG_ACT_NONE || act == ACG_ACT_RELU || act == ACG_ACT_LRELU) && !getenv("ACG_NO_FAST_EW")) {
            cudaStream_t st = static_cast<cudaStream_t>(stream);
            const __nv_bfloat16* zz = static_cast<const __nv_bfloat16*>(z);
            __nv_bfloat16* oo = static_cast<__nv_bfloat16*>(out);
            fast_launch_dims(rows, C, 4, false, &grid, &block);
            const RawMoments none{};
            if (act == ACG_ACT_RELU) launch_pdl(fast_bn_act_fwd_kernel<ACG_ACT_RELU, false>, grid, block, 0, st, zz, ld_in, C, rows, scale, shift, oo, ld_out, ew_rev() & 1, acts, n_act, hw, act_off, none);
            else if (act == ACG_ACT_LRELU) launch_pdl(fast_bn_act_fwd_kernel<ACG_ACT_LRELU, false>, grid, block, 0, st, zz, ld_in, C, rows, scale, shift, oo, ld_out, ew_rev() & 1, acts, n_act, hw, act_off, none);
            else launch_pdl(fast_bn_act_fwd_kernel<ACG_ACT_NONE, false>, grid, block, 0, st, zz, ld_in, C, rows, scale, shift, oo, ld_out, ew_rev() & 1, acts, n_act, hw, act_off, none);
            return check_launch(who);
        }
        vec_stream_launch_dims(rows / groups, C, groups, 8, &grid, &block);
        launch_pdl(vec_bn_act_fwd_kernel, grid, block, 0, static_cast<cudaStream_t>(stream), z, z_dtype, C, ld_in, rows / groups, scale, shift, act, out, out_dtype, ld_out);
    } else {
        launch_pdl(bn_act_fwd_kernel, ew_grid(rows * C), 256, 0, static_cast<cudaStream_t>(stream), z, z_dtype, rows, C, ld_in, rows / groups, scale, shift, act, out, out_dtype, ld_out);
    }
    int rc = check_launch(who);
    if (rc || !acts) return rc;
    // shapes outside the fast path: the concat as its own launch, same result
    return acg_tile_actions(acts, (int)(rows / hw), hw, n_act, out, out_dtype, ld_out, act_off, stream);
}

int acg_bn_act_fwd(const void* z, int z_dtype, long long rows, int C, int ld_in, int groups, const float* scale,
                   const float* shift, int act, void* out, int out_dtype, int ld_out, void* stream) {
    return bn_act_fwd_impl(z, z_dtype, rows, C, ld_in, groups, scale, shift, act, out, out_dtype, ld_out, nullptr, 0, 1, 0,
                           stream, "acg_bn_act_fwd");
}

int acg_bn_act_fwd_cat(const void* z, int z_dtype, long long rows, int C, int ld_in, const float* scale,
                       const float* shift, int act, void* out, int out_dtype, int ld_out, const float* actions,
                       int n_act, int hw, int act_off, void* stream) {
    ACG_REQUIRE(actions, ACG_ERR_INVALID, "acg_bn_act_fwd_cat: null actions");
    return bn_act_fwd_impl(z, z_dtype, rows, C, ld_in, 1, scale, shift, act, out, out_dtype, ld_out, actions, n_act, hw,
                           act_off, stream, "acg_bn_act_fwd_cat");
}

int acg_bn_finalize_act_fwd_ok(int C, int ld_in, int ld_out, int act) {
    return C > 0 && C % 8 == 0 && ld_in % 8 == 0 && ld_out % 8 == 0 && ld_in >= C && ld_out >= C &&
           (act == ACG_ACT_NONE || act == ACG_ACT_RELU || act == ACG_ACT_LRELU) && !getenv("ACG_NO_FAST_EW");
}

int acg_bn_finalize_act_fwd(const void* z, long long rows, int C, int ld_in, const double* stats,
                            const unsigned long long* stats_fix, const float* beta, long long norm_rows, float eps,
                            float* mean, float* rstd, float* scale, float* shift, int act, void* out, int ld_out,
                            const float* actions, int n_act, int hw, int act_off, void* stream) {
    using namespace acg;
    const char* who = "acg_bn_finalize_act_fwd";
    ACG_REQUIRE(z && out && stats && mean && rstd && scale && shift, ACG_ERR_INVALID, "%s: null pointer", who);
    ACG_REQUIRE(rows > 0 && norm_rows > 0, ACG_ERR_INVALID, "%s: bad size", who);
    ACG_REQUIRE(acg_bn_finalize_act_fwd_ok(C, ld_in, ld_out, act), ACG_ERR_UNSUPPORTED,
                "%s: C=%d ld_in=%d ld_out=%d act=%d (bf16 rows, channel counts in multiples of 8, relu / lrelu / none)", who, C,
                ld_in, ld_out, act);
    ACG_REQUIRE(al16(z) && al16(out) && ((uintptr_t)stats & 7) == 0 && ((uintptr_t)stats_fix & 7) == 0, ACG_ERR_INVALID,
                "%s: misaligned buffer", who);
    ACG_REQUIRE(!actions || (n_act > 0 && hw > 0 && rows % hw == 0 && act_off >= C && act_off + n_act <= ld_out),
                ACG_ERR_INVALID, "%s: action concat: %d channels at offset %d of %d, %d pixels per image", who, n_act,
                act_off, ld_out, hw);
    dim3 grid, block;
    fast_launch_dims(rows, C, 4, false, &grid, &block);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const __nv_bfloat16* zz = static_cast<const __nv_bfloat16*>(z);
    __nv_bfloat16* oo = static_cast<__nv_bfloat16*>(out);
    RawMoments rm{stats, stats_fix, beta, 1.0 / (double)norm_rows, eps, mean, rstd, scale, shift};
    const float* nof = nullptr;
    if (!actions) { n_act = 0; hw = 1; act_off = 0; }
    if (act == ACG_ACT_RELU) launch_pdl(fast_bn_act_fwd_kernel<ACG_ACT_RELU, true>, grid, block, 0, st, zz, ld_in, C, rows, nof, nof, oo, ld_out, ew_rev() & 1, actions, n_act, hw, act_off, rm);
    else if (act == ACG_ACT_LRELU) launch_pdl(fast_bn_act_fwd_kernel<ACG_ACT_LRELU, true>, grid, block, 0, st, zz, ld_in, C, rows, nof, nof, oo, ld_out, ew_rev() & 1, actions, n_act, hw, act_off, rm);
    else launch_pdl(fast_bn_act_fwd_kernel<ACG_ACT_NONE, true>, grid, block, 0, st, zz, ld_in, C, rows, nof, nof, oo, ld_out, ew_rev() & 1, actions, n_act, hw, act_off, rm);
    return check_launch(who);
}

int acg_bn_act_bwd_reduce(const void* dA, const void* dA2, int d_dtype, int ld_d, const void* z, int z_dtype, int ld_z,
                          long long rows, int C, int groups, const float* mean, const float* rstd,
                          const float* shift, int act, double* red, void* stream) {
    using namespace acg;
    ACG_REQUIRE(dA && red, ACG_ERR_INVALID, "acg_bn_act_bwd_reduce: null pointer");
    ACG_REQUIRE(z || (!mean && act == ACG_ACT_NONE), ACG_ERR_INVALID,
                "acg_bn_act_bwd_reduce: z may only be NULL for a layer without batch-norm and activation");
    ACG_REQUIRE(rows > 0 && C > 0 && ld_d >= C && (!z || ld_z >= C) && groups > 0 && rows % groups == 0,
                ACG_ERR_INVALID, "acg_bn_act_bwd_reduce: bad size");
    ACG_REQUIRE(dt_ok(d_dtype) && dt_ok(z_dtype), ACG_ERR_UNSUPPORTED, "acg_bn_act_bwd_reduce: dtype");
    const long long rpg = rows / groups;
    if (C % 8 == 0 && ld_d % 8 == 0 && (!z || ld_z % 8 == 0) && al16(z) && al16(dA) && al16(dA2) && al16(mean) &&
        al16(rstd) && al16(shift)) {
        dim3 grid, block;
        const bool fast = groups == 1 && d_dtype == ACG_BF16 && (!z || z_dtype == ACG_BF16) &&
                          (act == ACG_ACT_NONE || act == ACG_ACT_RELU || act == ACG_ACT_LRELU) && !getenv("ACG_NO_FAST_EW");
        if (fast) {
            cudaStream_t st = static_cast<cudaStream_t>(stream);
            fast_launch_dims(rows, C, 2, true, &grid, &block);   // 100-128 registers: two resident blocks per SM
            if (!z) launch_fast_reduce<ACG_ACT_NONE, false>(grid, block, st, z, ld_z, dA, dA2, ld_d, rows, C, mean, rstd, shift, red);
            else if (act == ACG_ACT_RELU) launch_fast_reduce<ACG_ACT_RELU, true>(grid, block, st, z, ld_z, dA, dA2, ld_d, rows, C, mean, rstd, shift, red);
            else if (act == ACG_ACT_LRELU) launch_fast_reduce<ACG_ACT_LRELU, true>(grid, block, st, z, ld_z, dA, dA2, ld_d, rows, C, mean, rstd, shift, red);
            else launch_fast_reduce<ACG_ACT_NONE, true>(grid, block, st, z, ld_z, dA, dA2, ld_d, rows, C, mean, rstd, shift, red);
            return check_launch("acg_bn_act_bwd_reduce");
        }
        vec_reduce_launch_dims(rpg, C, groups, &grid, &block);
        launch_pdl(vec_col_reduce_kernel<1>, grid, block, 0, static_cast<cudaStream_t>(stream), z, z_dtype, ld_z, dA, dA2, d_dtype, ld_d, rpg, C, mean, rstd, shift, act, red);
        return check_launch("acg_bn_act_bwd_reduce");
    }
    launch_pdl(col_reduce_kernel<1>, reduce_grid(rpg, C, groups), dim3(kTX, kTY), 0, static_cast<cudaStream_t>(stream), z, z_dtype, ld_z, dA, dA2, d_dtype, ld_d, rpg, C, mean, rstd, shift, act, red);
    return check_launch("acg_bn_act_bwd_reduce");
}

int acg_bn_act_bwd_reduce_sync(const void* dA, const void* dA2, int d_dtype, int ld_d, const void* z, int z_dtype,
                               int ld_z, long long rows, int C, const float* mean, const float* rstd,
                               const float* shift, int act, double* red, unsigned int* counter,
                               const acg_peer_exchange* peer, void* stream) {
    using namespace acg;
    ACG_REQUIRE(dA && red && counter && peer, ACG_ERR_INVALID, "acg_bn_act_bwd_reduce_sync: null pointer");
    ACG_REQUIRE(z || (!mean && act == ACG_ACT_NONE), ACG_ERR_INVALID,
                "acg_bn_act_bwd_reduce_sync: z may only be NULL for a layer without batch-norm and activation");
    ACG_REQUIRE(rows > 0 && C > 0 && ld_d >= C && (!z || ld_z >= C), ACG_ERR_INVALID, "acg_bn_act_bwd_reduce_sync: bad size");
    PeerExchange x;
    int rc = fill_peer_exchange(&x, peer, 2 * C, "acg_bn_act_bwd_reduce_sync");
    if (rc) return rc;
    const bool fast = C % 8 == 0 && ld_d % 8 == 0 && (!z || ld_z % 8 == 0) && al16(z) && al16(dA) && al16(dA2) &&
                      al16(mean) && al16(rstd) && al16(shift) && d_dtype == ACG_BF16 && (!z || z_dtype == ACG_BF16) &&
                      (act == ACG_ACT_NONE || act == ACG_ACT_RELU || act == ACG_ACT_LRELU) && !getenv("ACG_NO_FAST_EW") &&
                      peer->world > 1;
    if (!fast) {        // same result in two launches: the reduction pass, then the stand-alone exchange kernel
        rc = acg_bn_act_bwd_reduce(dA, dA2, d_dtype, ld_d, z, z_dtype, ld_z, rows, C, 1, mean, rstd, shift, act, red, stream);
        if (rc || peer->world <= 1) return rc;
        return acg_peer_allreduce_f64(red, 2 * C, peer->cap, peer->slot_off, peer->rank, peer->world, peer->mailboxes,
                                      peer->epoch, peer->timeout_s, 0, nullptr, 0, 0.f, nullptr, nullptr, nullptr, nullptr,
                                      stream);
    }
    dim3 grid, block;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    fast_launch_dims(rows, C, 2, true, &grid, &block);
    if (!z) launch_fast_reduce<ACG_ACT_NONE, false>(grid, block, st, z, ld_z, dA, dA2, ld_d, rows, C, mean, rstd, shift, red, counter, &x);
    else if (act == ACG_ACT_RELU) launch_fast_reduce<ACG_ACT_RELU, true>(grid, block, st, z, ld_z, dA, dA2, ld_d, rows, C, mean, rstd, shift, red, counter, &x);
    else if (act == ACG_ACT_LRELU) launch_fast_reduce<ACG_ACT_LRELU, true>(grid, block, st, z, ld_z, dA, dA2, ld_d, rows, C, mean, rstd, shift, red, counter, &x);
    else launch_fast_reduce<ACG_ACT_NONE, true>(grid, block, st, z, ld_z, dA, dA2, ld_d, rows, C, mean, rstd, shift, red, counter, &x);
    return check_launch("acg_bn_act_bwd_reduce_sync");
}

int acg_bn_act_bwd_apply(const void* dA, const void* dA2, int d_dtype, int ld_d, const void* z, int z_dtype, int ld_z,
                         long long rows, int C, int groups, const float* mean, const float* rstd,
                         const float* shift, int act, int has_bn, const double* red, void* dz, int dz_dtype,
                         int ld_dz, float* dbeta, long long norm_rows, float dbeta_scale, void* stream) {
    using namespace acg;
    ACG_REQUIRE(dA && red && dz, ACG_ERR_INVALID, "acg_bn_act_bwd_apply: null pointer");
    ACG_REQUIRE(z || (!has_bn && act == ACG_ACT_NONE), ACG_ERR_INVALID,
                "acg_bn_act_bwd_apply: z may only be NULL for a layer without batch-norm and activation");
    ACG_REQUIRE(rows > 0 && C > 0 && ld_d >= C && (!z || ld_z >= C) && ld_dz >= C && groups > 0 && rows % groups == 0,
                ACG_ERR_INVALID, "acg_bn_act_bwd_apply: bad size");
    ACG_REQUIRE(dt_ok(d_dtype) && dt_ok(z_dtype) && dt_ok(dz_dtype), ACG_ERR_UNSUPPORTED,
                "acg_bn_act_bwd_apply: dtype");
    if (C % 8 == 0 && ld_d % 8 == 0 && (!z || ld_z % 8 == 0) && ld_dz % 8 == 0 && al16(z) && al16(dA) && al16(dA2) &&
        al16(dz) && al16(mean) && al16(rstd) && al16(shift)) {
        dim3 grid, block;
        const bool fast = groups == 1 && z && d_dtype == ACG_BF16 && z_dtype == ACG_BF16 && dz_dtype == ACG_BF16 &&
                          (act == ACG_ACT_NONE || act == ACG_ACT_RELU || act == ACG_ACT_LRELU) && !getenv("ACG_NO_FAST_EW");
        if (fast) {
            cudaStream_t st = static_cast<cudaStream_t>(stream);
            const long long nr = norm_rows > 0 ? norm_rows : rows;
            fast_launch_dims(rows, C, 2, false, &grid, &block);   // 100-128 registers: two resident blocks per SM
#define ACG_FAST_APPLY(A, BN) launch_fast_apply<A, BN>(grid, block, st, dA, dA2, ld_d, z, ld_z, C, rows, mean, rstd, shift, red, dz, ld_dz, dbeta, nr, dbeta_scale)
            if (has_bn) {
                if (act == ACG_ACT_RELU) ACG_FAST_APPLY(ACG_ACT_RELU, true);
                else if (act == ACG_ACT_LRELU) ACG_FAST_APPLY(ACG_ACT_LRELU, true);
                else ACG_FAST_APPLY(ACG_ACT_NONE, true);
            } else {
                if (act == ACG_ACT_RELU) ACG_FAST_APPLY(ACG_ACT_RELU, false);
                else if (act == ACG_ACT_LRELU) ACG_FAST_APPLY(ACG_ACT_LRELU, false);
                else ACG_FAST_APPLY(ACG_ACT_NONE, false);
            }
#undef ACG_FAST_APPLY
            return check_launch("acg_bn_act_bwd_apply");
        }
        vec_stream_launch_dims(rows / groups, C, groups, 8, &grid, &block);   // 4 resident blocks/SM, 2 waves
        launch_pdl(vec_bn_act_bwd_apply_kernel, grid, block, 0, static_cast<cudaStream_t>(stream), dA, dA2, d_dtype, ld_d, z, z_dtype, ld_z, C, groups, rows / groups, mean, rstd, shift, act, has_bn, red,
            dz, dz_dtype, ld_dz, dbeta, norm_rows > 0 ? norm_rows : rows / groups, dbeta_scale);
        return check_launch("acg_bn_act_bwd_apply");
    }
    launch_pdl(bn_act_bwd_apply_kernel, ew_grid(rows * C), 256, 0, static_cast<cudaStream_t>(stream), dA, dA2, d_dtype, ld_d, z, z_dtype, ld_z, rows, C, groups, rows / groups, mean, rstd, shift, act, has_bn, red,
        dz, dz_dtype, ld_dz, dbeta, norm_rows > 0 ? norm_rows : rows / groups, dbeta_scale);
    return check_launch("acg_bn_act_bwd_apply");
}

int acg_bias_grad(const double* red, int C, float scale, float* dbias, void* stream) {
    using namespace acg;
    ACG_REQUIRE(red && dbias && C > 0, ACG_ERR_INVALID, "acg_bias_grad: bad argument");
    launch_pdl(bias_grad_kernel, (C + 127) / 128, 128, 0, static_cast<cudaStream_t>(stream), red, C, scale, dbias);
    return check_launch("acg_bias_grad");
}

int acg_copy_channels(const void* src, int src_dtype, int ld_src, int off_src, void* dst, int dst_dtype,
                      int ld_dst, int off_dst, long long rows, int n, void* stream) {
    using namespace acg;
    ACG_REQUIRE(src && dst, ACG_ERR_INVALID, "acg_copy_channels: null pointer");
    ACG_REQUIRE(rows > 0 && n > 0 && off_src >= 0 && off_dst >= 0 && off_src + n <= ld_src && off_dst + n <= ld_dst,
                ACG_ERR_INVALID, "acg_copy_channels: bad slice");
    ACG_REQUIRE(dt_ok(src_dtype) && dt_ok(dst_dtype), ACG_ERR_UNSUPPORTED, "acg_copy_channels: dtype");
    launch_pdl(copy_channels_kernel, ew_grid(rows * n), 256, 0, static_cast<cudaStream_t>(stream), src, src_dtype, ld_src, off_src, dst, dst_dtype, ld_dst, off_dst, rows, n);
    return check_launch("acg_copy_channels");
}

int acg_pack_frames(const float* a, const float* b, int C, void* out_bf16, int ld_out, long long rows, void* stream) {
    using namespace acg;
    ACG_REQUIRE(a && out_bf16 && rows > 0, ACG_ERR_INVALID, "acg_pack_frames: bad argument");
    ACG_REQUIRE(C == 3 && (ld_out == 8 || ld_out == 16) && al16(out_bf16), ACG_ERR_UNSUPPORTED,
                "acg_pack_frames: built for RGB frames into 8- or 16-channel rows (C=%d, ld_out=%d)", C, ld_out);
    long long blocks = (rows + 255) / 256;
    if (blocks > num_sms() * 8) blocks = num_sms() * 8;
    if (ld_out == 8)
        launch_pdl(pack_frames_kernel<3, 8>, (int)blocks, 256, 0, static_cast<cudaStream_t>(stream), a, b, static_cast<__nv_bfloat16*>(out_bf16), rows);
    else
        launch_pdl(pack_frames_kernel<3, 16>, (int)blocks, 256, 0, static_cast<cudaStream_t>(stream), a, b, static_cast<__nv_bfloat16*>(out_bf16), rows);
    return check_launch("acg_pack_frames");
}

int acg_tile_actions(const float* actions, int B, int hw, int A, void* dst, int dst_dtype, int ld_dst, int off,
                     void* stream) {
    using namespace acg;
    ACG_REQUIRE(actions && dst, ACG_ERR_INVALID, "acg_tile_actions: null pointer");
    ACG_REQUIRE(B > 0 && hw > 0 && A > 0 && off >= 0 && off + A <= ld_dst, ACG_ERR_INVALID,
                "acg_tile_actions: bad size");
    ACG_REQUIRE(dt_ok(dst_dtype), ACG_ERR_UNSUPPORTED, "acg_tile_actions: dtype");
    launch_pdl(tile_actions_kernel, ew_grid((long long)B * hw * A), 256, 0, static_cast<cudaStream_t>(stream), actions, B, hw, A, dst, dst_dtype, ld_dst, off);
    return check_launch("acg_tile_actions");
}

}  // extern "C"
