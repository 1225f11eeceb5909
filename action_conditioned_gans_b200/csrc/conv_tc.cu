// tcgen05 / TMEM implicit-GEMM convolution kernels for sm_100a (bf16 operands, fp32 accumulation in tensor memory).
//
// Replaces slim.conv2d / slim.conv2d_transpose (models.py:12-21,34-59,82-88) and TF autodiff's data gradients:
//   CONV gather  : y[m=(b,oh,ow)][n] = sum_{tap,ci} x[b, oh*s+a-pt, ow*s+c-pl, ci] * Wf[n][tap][ci]
//                  -> conv2d forward, and the data gradient of conv2d_transpose
//   ADJ  gather  : dx[m=(b,ih,iw)][n] = sum_{tap in class,co} dy[b,(ih+pt-a)/s,(iw+pl-c)/s,co] * Wb[class][n][tap][co]
//                  -> conv2d_transpose forward, and the data gradient of conv2d.  Stride-2 layers are split into the
//                     4 output-parity classes (grid.z), each with only the taps that hit it (3x3/3x2/2x3/2x2 of 5x5).
//
// One CTA = one 128 x N output tile (N <= 128, multiple of 16), 160 threads:
//   warps 0-3  producers: gather the A (activation) and B (packed weight) K-slices of 64 bf16 straight into the
//              128B-swizzled K-major shared-memory layout the tensor core reads (cp.async 16 B with zero fill for
//              the SAME padding / ragged edges), completion signalled on the stage's mbarrier
//              (cp.async.mbarrier.arrive.noinc); after the main loop the same warps are the epilogue.
//   warp 4     allocates tensor memory, one elected lane issues tcgen05.mma (M=128, N, K=16) x4 per stage and
//              tcgen05.commit's the stage's "empty" barrier, finally the accumulator barrier.
//   epilogue   tcgen05.ld 32 lanes x 16 columns per warp -> (+bias, tanh) -> bf16 / fp32 NHWC rows.
// Two CTAs are resident per SM (3 stages x 32 KB, 128 TMEM columns each) so one CTA's epilogue overlaps the
// other's main loop.
#include "common.cuh"

namespace acg {
namespace tc {

constexpr int BM = 128, BK = 64, BN = 128, STAGES = 3;
constexpr int kProducers = 128, kThreads = 160;
constexpr int kStageA = BM * BK * 2, kStageB = BN * BK * 2;
constexpr int kSmemBytes = STAGES * (kStageA + kStageB) + 1024;  // + alignment slack

struct Params {
    const __nv_bfloat16* a_src;
    const __nv_bfloat16* w_pack;
    void* out;
    const float* bias;
    int B, H, W, OH, OW, KH, KW, stride, pad_t, pad_l;
    int lda;         // channel stride of a_src == channels per tap in the packed K dimension (multiple of 8)
    int ldo;         // channel stride of the output rows (>= N)
    int N;           // GEMM N of this launch (multiple of 16)
    int n_bias;      // bias entries (real output channels)
    int n_store;     // output channels written per row: min(N, ldo)
    int out_dtype, out_act;
    long long w_class_off[4];   // ADJ: element offset of each parity class' [N][Kc] matrix inside w_pack
};

__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src, uint32_t src_bytes) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async_arrive_noinc(uint64_t* bar) {
    asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void tc_mma(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc)
        : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// shared-memory matrix descriptor with 128-byte swizzle (atoms of 8 rows x 128 B = 1024 B, 1024 B aligned).
//   K-major : a row is 64 consecutive K elements of one M/N index; SBO = stride between 8-row (M/N) groups;
//             LBO is unused.
//   MN-major: a row is 64 consecutive M/N elements of one K index; SBO = stride between 8-row (K) groups;
//             LBO = stride between 64-element M/N atoms.
__device__ __forceinline__ uint64_t sw128_desc(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr >> 4) & 0x3FFF);        // start address
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;  // leading byte offset
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;  // stride byte offset
    d |= (uint64_t)1 << 46;                            // descriptor version (Blackwell)
    d |= (uint64_t)2 << 61;                            // SWIZZLE_128B
    return d;
}
__device__ __forceinline__ uint64_t kmajor_sw128_desc(uint32_t smem_addr) { return sw128_desc(smem_addr, 16, 1024); }
// instruction descriptor: D=f32, A=B=bf16, both K-major, M=128, N runtime
__device__ __forceinline__ uint32_t make_idesc(int n, int a_mn_major, int b_mn_major) {
    return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)a_mn_major << 15) | ((uint32_t)b_mn_major << 16) |
           ((uint32_t)(n >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
}

enum { CONV = 0, ADJ = 1 };

template <int MODE>
__global__ void __launch_bounds__(kThreads, 2)
conv_tc_kernel(const Params p) {
    extern __shared__ unsigned char smem_raw[];
    __shared__ __align__(8) uint64_t full_bar[STAGES], empty_bar[STAGES], acc_bar;
    __shared__ uint32_t tmem_base_sh;

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t smemA = smem_base, smemB = smem_base + STAGES * kStageA;

    // ---- geometry of this CTA ---------------------------------------------------------------------------
    const int s = p.stride;
    int M, ntaps, nc = 1, a0 = 0, c0 = 0, ph = 0, pw = 0, Hp = 0, Wp = 0;
    const __nv_bfloat16* wmat = p.w_pack;
    if (MODE == CONV) {
        M = p.B * p.OH * p.OW;
        ntaps = p.KH * p.KW;
    } else {
        ph = blockIdx.z / s; pw = blockIdx.z % s;
        Hp = (p.H - ph + s - 1) / s; Wp = (p.W - pw + s - 1) / s;
        a0 = (ph + p.pad_t) % s; c0 = (pw + p.pad_l) % s;
        const int na = a0 < p.KH ? (p.KH - a0 + s - 1) / s : 0;
        nc = c0 < p.KW ? (p.KW - c0 + s - 1) / s : 0;
        ntaps = na * nc;
        M = p.B * Hp * Wp;
        wmat += p.w_class_off[blockIdx.z];
        if (nc == 0) nc = 1;
    }
    const int tile_m = blockIdx.x * BM;
    if (tile_m >= M) return;   // whole CTA exits together (smaller parity classes)
    const int n0 = blockIdx.y * BN;
    const int n_cta = min(BN, p.N - n0);
    const int Ktot = ntaps * p.lda;
    const int nkb = (Ktot + BK - 1) / BK;

    if (tid == 0) {
        for (int i = 0; i < STAGES; ++i) { mbar_init(&full_bar[i], kProducers); mbar_init(&empty_bar[i], 1); }
        mbar_init(&acc_bar, 1);
        fence_mbar_init();
    }
    if (warp == 4) {   // tensor-memory allocation is warp-collective; this warp also owns the dealloc
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_sh)),
                     "r"((uint32_t)BN) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = tmem_base_sh;

    if (warp < 4) {
        // ================================ producers ================================
        const int j = tid & 7, rslot = tid >> 3;
        int row_b[8], row_y[8], row_x[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const int m = tile_m + rslot + 16 * i;
            if (m < M) {
                if (MODE == CONV) {
                    const int b = m / (p.OH * p.OW), r = m - b * p.OH * p.OW;
                    const int oh = r / p.OW, ow = r - oh * p.OW;
                    row_b[i] = b * p.H * p.W; row_y[i] = oh * s - p.pad_t; row_x[i] = ow * s - p.pad_l;
                } else {
                    const int b = m / (Hp * Wp), r = m - b * Hp * Wp;
                    const int ih = (r / Wp) * s + ph, iw = (r % Wp) * s + pw;
                    row_b[i] = b * p.OH * p.OW; row_y[i] = ih + p.pad_t; row_x[i] = iw + p.pad_l;
                }
            } else {
                row_b[i] = 0; row_y[i] = -(1 << 28); row_x[i] = -(1 << 28);   // always out of range -> zero fill
            }
        }
        int tap = (j * 8) / p.lda, ci = (j * 8) % p.lda;
        for (int kb = 0; kb < nkb; ++kb) {
            const int stage = kb % STAGES;
            if (kb >= STAGES) mbar_wait(&empty_bar[stage], (uint32_t)(((kb / STAGES) - 1) & 1));
            const bool kvalid = tap < ntaps;
            int a, c;
            if (MODE == CONV) { a = tap / p.KW; c = tap - a * p.KW; }
            else { const int ta = tap / nc, tcc = tap - ta * nc; a = a0 + s * ta; c = c0 + s * tcc; }
            const uint32_t dstA = smemA + stage * kStageA;
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const int r = rslot + 16 * i;
                bool ok;
                long long off;
                if (MODE == CONV) {
                    const int ih = row_y[i] + a, iw = row_x[i] + c;
                    ok = kvalid && (unsigned)ih < (unsigned)p.H && (unsigned)iw < (unsigned)p.W;
                    off = ((long long)(row_b[i] + ih * p.W + iw)) * p.lda + ci;
                } else {
                    const int ny = row_y[i] - a, nx = row_x[i] - c;   // multiples of the stride inside a class
                    const int oh = ny / s, ow = nx / s;
                    ok = kvalid && ny >= 0 && nx >= 0 && oh < p.OH && ow < p.OW;
                    off = ((long long)(row_b[i] + oh * p.OW + ow)) * p.lda + ci;
                }
                cp_async16(dstA + r * 128 + ((j ^ (r & 7)) << 4), ok ? (const void*)(p.a_src + off) : (const void*)p.a_src,
                           ok ? 16u : 0u);
            }
            const uint32_t dstB = smemB + stage * kStageB;
            const long long kk = (long long)kb * BK + j * 8;
            for (int r = rslot; r < n_cta; r += 16) {
                const bool ok = kk < Ktot;
                cp_async16(dstB + r * 128 + ((j ^ (r & 7)) << 4),
                           ok ? (const void*)(wmat + (long long)(n0 + r) * Ktot + kk) : (const void*)wmat, ok ? 16u : 0u);
            }
            cp_async_arrive_noinc(&full_bar[stage]);
            ci += BK;
            while (ci >= p.lda) { ci -= p.lda; ++tap; }
        }
    } else if (lane == 0) {
        // ================================ MMA issuer ================================
        const uint32_t idesc = make_idesc(n_cta, 0, 0);
        for (int kb = 0; kb < nkb; ++kb) {
            const int stage = kb % STAGES;
            mbar_wait(&full_bar[stage], (uint32_t)((kb / STAGES) & 1));
            tc_fence_after();
            const uint32_t aaddr = smemA + stage * kStageA, baddr = smemB + stage * kStageB;
#pragma unroll
            for (int k = 0; k < BK / 16; ++k)
                tc_mma(tmem_base, kmajor_sw128_desc(aaddr + k * 32), kmajor_sw128_desc(baddr + k * 32), idesc,
                       (kb | k) != 0 ? 1u : 0u);
            tc_commit(&empty_bar[stage]);   // arrives when the MMAs above have finished reading this stage
        }
        tc_commit(&acc_bar);
    }

    // ================================ epilogue (warps 0-3) ================================
    if (warp < 4) {
        if (nkb > 0) {
            mbar_wait(&acc_bar, 0);
            tc_fence_after();
        }
        const int m = tile_m + warp * 32 + lane;
        size_t row_off = 0;
        if (m < M) {
            if (MODE == CONV) row_off = (size_t)m * p.ldo;
            else {
                const int b = m / (Hp * Wp), r = m - b * Hp * Wp;
                const int ih = (r / Wp) * s + ph, iw = (r % Wp) * s + pw;
                row_off = ((size_t)(b * p.H + ih) * p.W + iw) * p.ldo;
            }
        }
        for (int cb = 0; cb < n_cta; cb += 16) {
            uint32_t v[16];
            if (nkb > 0) tmem_ld16(tmem_base + ((uint32_t)(warp * 32) << 16) + cb, v);
            else {
#pragma unroll
                for (int i = 0; i < 16; ++i) v[i] = 0u;
            }
            if (m < M) {
                float f[16];
#pragma unroll
                for (int i = 0; i < 16; ++i) {
                    f[i] = __uint_as_float(v[i]);
                    const int n = n0 + cb + i;
                    if (p.bias && n < p.n_bias) f[i] += p.bias[n];
                    if (p.out_act == ACG_ACT_TANH) f[i] = tanhf(f[i]);
                }
                const bool full = n0 + cb + 16 <= p.n_store;
                if (p.out_dtype == ACG_BF16) {
                    __nv_bfloat16* o = static_cast<__nv_bfloat16*>(p.out) + row_off + n0 + cb;
                    if (full && (p.ldo & 7) == 0) {
                        uint32_t w[8];
#pragma unroll
                        for (int i = 0; i < 8; ++i) {
                            __nv_bfloat162 h = __floats2bfloat162_rn(f[2 * i], f[2 * i + 1]);
                            w[i] = *reinterpret_cast<uint32_t*>(&h);
                        }
                        reinterpret_cast<uint4*>(o)[0] = make_uint4(w[0], w[1], w[2], w[3]);
                        reinterpret_cast<uint4*>(o)[1] = make_uint4(w[4], w[5], w[6], w[7]);
                    } else {
#pragma unroll
                        for (int i = 0; i < 16; ++i)
                            if (n0 + cb + i < p.n_store) o[i] = __float2bfloat16_rn(f[i]);
                    }
                } else {
                    float* o = static_cast<float*>(p.out) + row_off + n0 + cb;
                    if (full && (p.ldo & 3) == 0) {
#pragma unroll
                        for (int i = 0; i < 4; ++i)
                            reinterpret_cast<float4*>(o)[i] = make_float4(f[4 * i], f[4 * i + 1], f[4 * i + 2], f[4 * i + 3]);
                    } else {
#pragma unroll
                        for (int i = 0; i < 16; ++i)
                            if (n0 + cb + i < p.n_store) o[i] = f[i];
                    }
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 4) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)BN) : "memory");
    }
}


// ---- wgrad ---------------------------------------------------------------------------------------------------
// dW[tap][ci][co] += sum_pix x[pix @ tap][ci] * dy[pix][co]   as   D[m = co][n = (tap, ci)] = sum_k A[m][k] B[n][k]
// with k = output pixel.  Both operands are channel-contiguous in NHWC, i.e. MN-major for this GEMM: a stage holds
// [64-channel atom][64 pixels][128 B] per operand and the instruction descriptor sets a_major = b_major = MN.
// grid = (N tiles over taps x ld_x, M tiles over Cout, pixel splits); partial sums are reduced with fp32 RED.
struct WgradParams {
    const __nv_bfloat16* x;
    const __nv_bfloat16* dy;
    float* dw;
    int B, H, W, OH, OW, KH, KW, stride, pad_t, pad_l;
    int Cin, Cout, ldx, ldy;
    int k_chunk;     // pixels per split (multiple of BK)
};

__global__ void __launch_bounds__(kThreads, 2)
conv_wgrad_tc_kernel(const WgradParams p) {
    extern __shared__ unsigned char smem_raw[];
    __shared__ __align__(8) uint64_t full_bar[STAGES], empty_bar[STAGES], acc_bar;
    __shared__ uint32_t tmem_base_sh;

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t smemA = smem_base, smemB = smem_base + STAGES * kStageA;

    const int Ntot = p.KH * p.KW * p.ldx;
    const int n0 = blockIdx.x * BN;
    const int n_cta = min(BN, Ntot - n0);
    const int co0 = blockIdx.y * BM;
    const int Kd = p.B * p.OH * p.OW;
    const int k_begin = blockIdx.z * p.k_chunk;
    const int k_end = min(Kd, k_begin + p.k_chunk);
    const int nkb = k_end > k_begin ? (k_end - k_begin + BK - 1) / BK : 0;
    if (nkb == 0) return;

    if (tid == 0) {
        for (int i = 0; i < STAGES; ++i) { mbar_init(&full_bar[i], kProducers); mbar_init(&empty_bar[i], 1); }
        mbar_init(&acc_bar, 1);
        fence_mbar_init();
    }
    if (warp == 4) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_sh)),
                     "r"((uint32_t)BN) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = tmem_base_sh;

    if (warp < 4) {
        // producers: thread = (16-byte channel chunk c16 of the 128-wide tile, pixel slot)
        const int c16 = tid & 15, pslot = tid >> 4;
        const int atom = c16 >> 3, jc = c16 & 7;
        // A: dy channels co0 + c16*8 .. +8
        const int a_ch = co0 + c16 * 8;
        const bool a_ch_ok = a_ch + 8 <= p.ldy;
        // B: n = n0 + c16*8 -> (tap, ci) fixed for the whole kernel
        const int nn = n0 + c16 * 8;
        const bool b_ch_ok = c16 * 8 < n_cta;
        const int tap = nn / p.ldx, ci = nn - tap * p.ldx;
        const int ta = tap / p.KW, tcc = tap - ta * p.KW;
        for (int kb = 0; kb < nkb; ++kb) {
            const int stage = kb % STAGES;
            if (kb >= STAGES) mbar_wait(&empty_bar[stage], (uint32_t)(((kb / STAGES) - 1) & 1));
            const uint32_t dstA = smemA + stage * kStageA + atom * (BK * 128);
            const uint32_t dstB = smemB + stage * kStageB + atom * (BK * 128);
#pragma unroll
            for (int i = 0; i < BK / 8; ++i) {
                const int k = pslot + 8 * i;
                const int pix = k_begin + kb * BK + k;
                const bool pv = pix < k_end;
                const uint32_t soff = k * 128 + ((jc ^ (k & 7)) << 4);
                const bool aok = pv && a_ch_ok;
                cp_async16(dstA + soff, aok ? (const void*)(p.dy + (long long)pix * p.ldy + a_ch) : (const void*)p.dy,
                           aok ? 16u : 0u);
                bool bok = false;
                long long boff = 0;
                if (pv && b_ch_ok) {
                    const int b = pix / (p.OH * p.OW), r = pix - b * p.OH * p.OW;
                    const int oh = r / p.OW, ow = r - oh * p.OW;
                    const int ih = oh * p.stride + ta - p.pad_t, iw = ow * p.stride + tcc - p.pad_l;
                    bok = (unsigned)ih < (unsigned)p.H && (unsigned)iw < (unsigned)p.W;
                    boff = ((long long)(b * p.H + ih) * p.W + iw) * p.ldx + ci;
                }
                cp_async16(dstB + soff, bok ? (const void*)(p.x + boff) : (const void*)p.x, bok ? 16u : 0u);
            }
            cp_async_arrive_noinc(&full_bar[stage]);
        }
    } else if (lane == 0) {
        const uint32_t idesc = make_idesc(n_cta, 1, 1);
        for (int kb = 0; kb < nkb; ++kb) {
            const int stage = kb % STAGES;
            mbar_wait(&full_bar[stage], (uint32_t)((kb / STAGES) & 1));
            tc_fence_after();
            const uint32_t aaddr = smemA + stage * kStageA, baddr = smemB + stage * kStageB;
#pragma unroll
            for (int k = 0; k < BK / 16; ++k)   // 16 pixels = two 8-row groups further down each atom
                tc_mma(tmem_base, sw128_desc(aaddr + k * 2048, BK * 128, 1024), sw128_desc(baddr + k * 2048, BK * 128, 1024),
                       idesc, (kb | k) != 0 ? 1u : 0u);
            tc_commit(&empty_bar[stage]);
        }
        tc_commit(&acc_bar);
    }

    if (warp < 4) {
        mbar_wait(&acc_bar, 0);
        tc_fence_after();
        const int co = co0 + warp * 32 + lane;
        for (int cb = 0; cb < n_cta; cb += 16) {
            uint32_t v[16];
            tmem_ld16(tmem_base + ((uint32_t)(warp * 32) << 16) + cb, v);
            if (co < p.Cout) {
                const int nn = n0 + cb;
                const int tap = nn / p.ldx, ci0 = nn - tap * p.ldx;   // 16-column groups never straddle a tap (ldx % 16 == 0)
#pragma unroll
                for (int i = 0; i < 16; ++i) {
                    const int ci = ci0 + i;
                    if (ci < p.Cin) atomicAdd(p.dw + ((size_t)tap * p.Cin + ci) * p.Cout + co, __uint_as_float(v[i]));
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 4) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)BN) : "memory");
    }
}

// ---- weight packing ------------------------------------------------------------------------------------------
// HWIO fp32 w[a][c][ci][co] -> CONV pack bf16 Wf[n = co (N rows, zero padded)][tap][cis (zero padded)]
__global__ void __launch_bounds__(256)
pack_conv_kernel(const float* __restrict__ w, int taps, int Cin, int Cout, int Cis, int N, __nv_bfloat16* __restrict__ out) {
    const long long total = (long long)N * taps * Cis;
    for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
         idx += (long long)gridDim.x * blockDim.x) {
        const int ci = (int)(idx % Cis);
        const long long r = idx / Cis;
        const int tap = (int)(r % taps);
        const int n = (int)(r / taps);
        float v = 0.f;
        if (ci < Cin && n < Cout) v = w[((size_t)tap * Cin + ci) * Cout + n];
        out[idx] = __float2bfloat16_rn(v);
    }
}
// HWIO fp32 -> ADJ pack: per parity class Wb[class][n = ci (N rows)][class tap][cos (zero padded)]
__global__ void __launch_bounds__(256)
pack_adj_kernel(const float* __restrict__ w, int KH, int KW, int Cin, int Cout, int Cos, int N, int stride, int pad_t,
                int pad_l, __nv_bfloat16* __restrict__ out) {
    const int cls = blockIdx.y;
    const int ph = cls / stride, pw = cls % stride;
    const int a0 = (ph + pad_t) % stride, c0 = (pw + pad_l) % stride;
    const int na = a0 < KH ? (KH - a0 + stride - 1) / stride : 0;
    const int nc = c0 < KW ? (KW - c0 + stride - 1) / stride : 0;
    // offset of this class = sum of the sizes of the classes before it
    long long off = 0;
    for (int q = 0; q < cls; ++q) {
        const int qa0 = (q / stride + pad_t) % stride, qc0 = (q % stride + pad_l) % stride;
        const int qna = qa0 < KH ? (KH - qa0 + stride - 1) / stride : 0;
        const int qnc = qc0 < KW ? (KW - qc0 + stride - 1) / stride : 0;
        off += (long long)N * qna * qnc * Cos;
    }
    const long long total = (long long)N * na * nc * Cos;
    for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
         idx += (long long)gridDim.x * blockDim.x) {
        const int co = (int)(idx % Cos);
        const long long r = idx / Cos;
        const int t = (int)(r % (na * nc));
        const int n = (int)(r / (na * nc));
        const int a = a0 + stride * (t / nc), c = c0 + stride * (t % nc);
        float v = 0.f;
        if (co < Cout && n < Cin) v = w[((size_t)(a * KW + c) * Cin + n) * Cout + co];
        out[off + idx] = __float2bfloat16_rn(v);
    }
}

int ru(int v, int m) { return (v + m - 1) / m * m; }

void class_taps(const acg_conv_shape* s, int cls, int* na, int* nc) {
    const int ph = cls / s->stride, pw = cls % s->stride;
    const int a0 = (ph + s->pad_t) % s->stride, c0 = (pw + s->pad_l) % s->stride;
    *na = a0 < s->KH ? (s->KH - a0 + s->stride - 1) / s->stride : 0;
    *nc = c0 < s->KW ? (s->KW - c0 + s->stride - 1) / s->stride : 0;
}

int check(const acg_conv_shape* s, const acg_tc_args* t, const char* who) {
    ACG_REQUIRE(s && t, ACG_ERR_INVALID, "%s: null shape/args", who);
    ACG_REQUIRE(s->stride == 1 || s->stride == 2, ACG_ERR_UNSUPPORTED, "%s: stride %d", who, s->stride);
    ACG_REQUIRE(t->ld_in % 8 == 0 && t->ld_in > 0, ACG_ERR_UNSUPPORTED, "%s: ld_in=%d must be a multiple of 8", who,
                t->ld_in);
    ACG_REQUIRE(t->ld_out > 0, ACG_ERR_INVALID, "%s: ld_out=%d", who, t->ld_out);
    ACG_REQUIRE(t->out_dtype == ACG_F32 || t->out_dtype == ACG_BF16, ACG_ERR_UNSUPPORTED, "%s: out dtype", who);
    ACG_REQUIRE((long long)s->B * s->H * s->W * (long long)t->ld_in < (1ll << 31) &&
                    (long long)s->B * s->OH * s->OW * (long long)t->ld_in < (1ll << 31) &&
                    (long long)s->B * s->H * s->W * (long long)t->ld_out < (1ll << 40),
                ACG_ERR_UNSUPPORTED, "%s: tensor too large", who);
    return ACG_OK;
}

int set_smem(const void* kern) {
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes) != cudaSuccess) {
        cudaError_t e = cudaGetLastError();
        set_error("conv_tc: cannot set dynamic smem: %s", cudaGetErrorString(e));
        return ACG_ERR_CUDA;
    }
    return ACG_OK;
}

}  // namespace tc
}  // namespace acg

extern "C" {

long long acg_pack_size(const acg_conv_shape* s, int which, int ld_k) {
    using namespace acg::tc;
    if (!s || ld_k <= 0) return -1;
    if (which == 0) return (long long)ru(s->Cout, 16) * s->KH * s->KW * ld_k;      // CONV pack
    long long tot = 0;
    for (int cls = 0; cls < s->stride * s->stride; ++cls) {
        int na, nc;
        class_taps(s, cls, &na, &nc);
        tot += (long long)ru(s->Cin, 16) * na * nc * ld_k;
    }
    return tot;
}

int acg_pack_weights(const acg_conv_shape* s, const float* w, int which, int ld_k, void* pack, void* stream) {
    using namespace acg;
    using namespace acg::tc;
    ACG_REQUIRE(s && w && pack, ACG_ERR_INVALID, "acg_pack_weights: null pointer");
    ACG_REQUIRE(ld_k % 8 == 0 && ld_k >= (which == 0 ? s->Cin : s->Cout), ACG_ERR_INVALID,
                "acg_pack_weights: ld_k=%d", ld_k);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (which == 0) {
        const int N = ru(s->Cout, 16);
        const long long total = (long long)N * s->KH * s->KW * ld_k;
        long long blocks = (total + 255) / 256;
        if (blocks > num_sms() * 8) blocks = num_sms() * 8;
        pack_conv_kernel<<<(int)blocks, 256, 0, st>>>(w, s->KH * s->KW, s->Cin, s->Cout, ld_k, N,
                                                     static_cast<__nv_bfloat16*>(pack));
    } else {
        const int N = ru(s->Cin, 16);
        dim3 grid(num_sms(), s->stride * s->stride);
        pack_adj_kernel<<<grid, 256, 0, st>>>(w, s->KH, s->KW, s->Cin, s->Cout, ld_k, N, s->stride, s->pad_t, s->pad_l,
                                             static_cast<__nv_bfloat16*>(pack));
    }
    return check_launch("acg_pack_weights");
}

int acg_conv_tc_supported(const acg_conv_shape* s, int which) {
    if (!s) return 0;
    if (s->stride != 1 && s->stride != 2) return 0;
    return 1;
}

int acg_conv_fprop_tc(const acg_conv_shape* s, const void* x_bf16, const void* w_pack, void* y, const acg_tc_args* t,
                      void* stream) {
    using namespace acg;
    using namespace acg::tc;
    int rc = check(s, t, "acg_conv_fprop_tc");
    if (rc) return rc;
    ACG_REQUIRE(x_bf16 && w_pack && y, ACG_ERR_INVALID, "acg_conv_fprop_tc: null pointer");
    ACG_REQUIRE(t->ld_in >= s->Cin, ACG_ERR_INVALID, "acg_conv_fprop_tc: ld_in < Cin");
    const int N = ru(s->Cout, 16);
    ACG_REQUIRE(t->ld_out >= s->Cout, ACG_ERR_INVALID, "acg_conv_fprop_tc: ld_out=%d < Cout=%d", t->ld_out, s->Cout);
    static bool ready = false;
    if (!ready) { rc = set_smem((const void*)conv_tc_kernel<CONV>); if (rc) return rc; ready = true; }
    Params p{};
    p.a_src = static_cast<const __nv_bfloat16*>(x_bf16); p.w_pack = static_cast<const __nv_bfloat16*>(w_pack);
    p.out = y; p.bias = t->bias;
    p.B = s->B; p.H = s->H; p.W = s->W; p.OH = s->OH; p.OW = s->OW; p.KH = s->KH; p.KW = s->KW;
    p.stride = s->stride; p.pad_t = s->pad_t; p.pad_l = s->pad_l;
    p.lda = t->ld_in; p.ldo = t->ld_out; p.N = N; p.n_bias = s->Cout; p.n_store = N < t->ld_out ? N : t->ld_out; p.out_dtype = t->out_dtype; p.out_act = t->out_act;
    const long long M = (long long)s->B * s->OH * s->OW;
    dim3 grid((unsigned)((M + BM - 1) / BM), (N + BN - 1) / BN, 1);
    conv_tc_kernel<CONV><<<grid, kThreads, kSmemBytes, static_cast<cudaStream_t>(stream)>>>(p);
    return check_launch("acg_conv_fprop_tc");
}

int acg_conv_dgrad_tc(const acg_conv_shape* s, const void* dy_bf16, const void* w_pack, void* dx, const acg_tc_args* t,
                      void* stream) {
    using namespace acg;
    using namespace acg::tc;
    int rc = check(s, t, "acg_conv_dgrad_tc");
    if (rc) return rc;
    ACG_REQUIRE(dy_bf16 && w_pack && dx, ACG_ERR_INVALID, "acg_conv_dgrad_tc: null pointer");
    ACG_REQUIRE(t->ld_in >= s->Cout, ACG_ERR_INVALID, "acg_conv_dgrad_tc: ld_in < Cout");
    const int N = ru(s->Cin, 16);
    ACG_REQUIRE(t->ld_out >= s->Cin, ACG_ERR_INVALID, "acg_conv_dgrad_tc: ld_out=%d < Cin=%d", t->ld_out, s->Cin);
    static bool ready = false;
    if (!ready) { rc = set_smem((const void*)conv_tc_kernel<ADJ>); if (rc) return rc; ready = true; }
    Params p{};
    p.a_src = static_cast<const __nv_bfloat16*>(dy_bf16); p.w_pack = static_cast<const __nv_bfloat16*>(w_pack);
    p.out = dx; p.bias = t->bias;
    p.B = s->B; p.H = s->H; p.W = s->W; p.OH = s->OH; p.OW = s->OW; p.KH = s->KH; p.KW = s->KW;
    p.stride = s->stride; p.pad_t = s->pad_t; p.pad_l = s->pad_l;
    p.lda = t->ld_in; p.ldo = t->ld_out; p.N = N; p.n_bias = s->Cin; p.n_store = N < t->ld_out ? N : t->ld_out; p.out_dtype = t->out_dtype; p.out_act = t->out_act;
    long long off = 0;
    const int ncls = s->stride * s->stride;
    for (int cls = 0; cls < ncls; ++cls) {
        int na, nc;
        class_taps(s, cls, &na, &nc);
        p.w_class_off[cls] = off;
        off += (long long)N * na * nc * t->ld_in;
    }
    const int Hp = (s->H + s->stride - 1) / s->stride, Wp = (s->W + s->stride - 1) / s->stride;
    const long long M = (long long)s->B * Hp * Wp;
    dim3 grid((unsigned)((M + BM - 1) / BM), (N + BN - 1) / BN, ncls);
    conv_tc_kernel<ADJ><<<grid, kThreads, kSmemBytes, static_cast<cudaStream_t>(stream)>>>(p);
    return check_launch("acg_conv_dgrad_tc");
}

int acg_conv_wgrad_tc(const acg_conv_shape* s, const void* x_bf16, const void* dy_bf16, float* dw, const acg_tc_args* t,
                      void* stream) {
    using namespace acg;
    using namespace acg::tc;
    ACG_REQUIRE(s && t && x_bf16 && dy_bf16 && dw, ACG_ERR_INVALID, "acg_conv_wgrad_tc: null pointer");
    ACG_REQUIRE(s->stride == 1 || s->stride == 2, ACG_ERR_UNSUPPORTED, "acg_conv_wgrad_tc: stride %d", s->stride);
    ACG_REQUIRE(t->ld_in % 16 == 0 && t->ld_in >= s->Cin, ACG_ERR_UNSUPPORTED,
                "acg_conv_wgrad_tc: ld_x=%d must be a multiple of 16 and >= Cin", t->ld_in);
    ACG_REQUIRE(t->ld_out % 8 == 0 && t->ld_out >= s->Cout, ACG_ERR_UNSUPPORTED,
                "acg_conv_wgrad_tc: ld_dy=%d must be a multiple of 8 and >= Cout", t->ld_out);
    ACG_REQUIRE((long long)s->B * s->H * s->W * (long long)t->ld_in < (1ll << 40) &&
                    (long long)s->B * s->OH * s->OW < (1ll << 31),
                ACG_ERR_UNSUPPORTED, "acg_conv_wgrad_tc: tensor too large");
    static bool ready = false;
    if (!ready) { int rc = set_smem((const void*)conv_wgrad_tc_kernel); if (rc) return rc; ready = true; }
    WgradParams p{};
    p.x = static_cast<const __nv_bfloat16*>(x_bf16); p.dy = static_cast<const __nv_bfloat16*>(dy_bf16); p.dw = dw;
    p.B = s->B; p.H = s->H; p.W = s->W; p.OH = s->OH; p.OW = s->OW; p.KH = s->KH; p.KW = s->KW;
    p.stride = s->stride; p.pad_t = s->pad_t; p.pad_l = s->pad_l;
    p.Cin = s->Cin; p.Cout = s->Cout; p.ldx = t->ld_in; p.ldy = t->ld_out;
    const int Ntot = s->KH * s->KW * t->ld_in;
    const int gx = (Ntot + BN - 1) / BN, gy = (s->Cout + BM - 1) / BM;
    const long long Kd = (long long)s->B * s->OH * s->OW;
    // split the pixel reduction so that ~2 waves of CTAs exist, at least 4 K blocks per split
    long long splits = ((long long)num_sms() * 4 + (long long)gx * gy - 1) / ((long long)gx * gy);
    const long long max_splits = (Kd + 4 * BK - 1) / (4 * BK);
    if (splits > max_splits) splits = max_splits;
    if (splits < 1) splits = 1;
    long long chunk = (Kd + splits - 1) / splits;
    chunk = (chunk + BK - 1) / BK * BK;
    splits = (Kd + chunk - 1) / chunk;
    p.k_chunk = (int)chunk;
    dim3 grid(gx, gy, (unsigned)splits);
    conv_wgrad_tc_kernel<<<grid, kThreads, kSmemBytes, static_cast<cudaStream_t>(stream)>>>(p);
    return check_launch("acg_conv_wgrad_tc");
}

}  // extern "C"
