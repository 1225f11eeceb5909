// fp32 SIMT implicit-GEMM convolution kernels: fprop, dgrad (== conv2d_transpose forward) and wgrad.
//
// Role in the framework: (1) the thin layers whose GEMM view cannot feed a tensor-core tile
// (g/conv1 with Cin=3, d/conv1 with Cin=6, g/sconv4-5, d/conv6 with Cout=1, the direct generator's
// 3-channel tconv4) and (2) the full-precision device reference the tcgen05 kernels are checked against at
// sizes the CPU oracle cannot finish.  Replaces slim.conv2d / slim.conv2d_transpose (models.py:12-21,34-59,
// 82-88) and their TF autodiff gradients with TF's SAME/VALID semantics carried in acg_conv_shape.
//
// One 64x64 output tile per CTA, K consumed in slices of 16 through shared memory, 4x4 outputs per thread.
// dgrad with stride 2 is decomposed into the 4 output-parity classes (grid.z) so that only the taps that
// actually hit a given pixel are multiplied (3x3 / 3x2 / 2x3 / 2x2 of the 5x5 filter) -- no zero MACs.
#include "common.cuh"

namespace acg {
namespace {

constexpr int TM = 64, TN = 64, TK = 16, kThreads = 256;
enum { FPROP = 0, DGRAD = 1, WGRAD = 2 };

struct Geo {
    acg_conv_shape s;
    int M, N, Kd;          // GEMM extents of this launch (per parity class for dgrad)
    int k_chunk;           // wgrad: K elements per grid.z slice
};

template <int MODE>
__global__ void __launch_bounds__(kThreads)
conv_simt_kernel(const float* __restrict__ A_src, const float* __restrict__ B_src, float* __restrict__ C_dst,
                 const acg_conv_shape s) {
    pdl_prologue();
    __shared__ __align__(16) float As[TK][TM + 4];
    __shared__ __align__(16) float Bs[TK][TN + 4];

    const int tid = threadIdx.x;
    const int tile_m = blockIdx.x * TM, tile_n = blockIdx.y * TN;

    // ---- per-launch geometry --------------------------------------------------------------
    int M, N, Kd, k_begin = 0, k_end;
    int ph = 0, pw = 0, Hp = 0, Wp = 0, a0 = 0, c0 = 0, na = 0, nc = 0;
    if (MODE == FPROP) {
        M = s.B * s.OH * s.OW; N = s.Cout; Kd = s.KH * s.KW * s.Cin; k_end = Kd;
    } else if (MODE == DGRAD) {
        ph = blockIdx.z / s.stride; pw = blockIdx.z % s.stride;
        Hp = (s.H - ph + s.stride - 1) / s.stride;
        Wp = (s.W - pw + s.stride - 1) / s.stride;
        a0 = (ph + s.pad_t) % s.stride; c0 = (pw + s.pad_l) % s.stride;
        na = a0 < s.KH ? (s.KH - a0 + s.stride - 1) / s.stride : 0;
        nc = c0 < s.KW ? (s.KW - c0 + s.stride - 1) / s.stride : 0;
        M = s.B * Hp * Wp; N = s.Cin; Kd = na * nc * s.Cout; k_end = Kd;
    } else {
        M = s.KH * s.KW * s.Cin; N = s.Cout; Kd = s.B * s.OH * s.OW;
        const int chunk = (Kd + gridDim.z - 1) / gridDim.z;
        k_begin = blockIdx.z * chunk;
        k_end = min(Kd, k_begin + chunk);
    }
    if (tile_m >= M) return;

    // ---- the A row this thread stages (fixed for the whole kernel) ---------------------------
    const int am = tile_m + (tid & (TM - 1));
    const bool am_ok = am < M;
    int m_b = 0, m_y = 0, m_x = 0, m_tap_a = 0, m_tap_c = 0, m_ci = 0;
    if (am_ok) {
        if (MODE == FPROP) {
            m_b = am / (s.OH * s.OW); int r = am - m_b * s.OH * s.OW; m_y = r / s.OW; m_x = r - m_y * s.OW;
        } else if (MODE == DGRAD) {
            m_b = am / (Hp * Wp); int r = am - m_b * Hp * Wp; m_y = (r / Wp) * s.stride + ph; m_x = (r % Wp) * s.stride + pw;
        } else {
            int tap = am / s.Cin; m_ci = am - tap * s.Cin; m_tap_a = tap / s.KW; m_tap_c = tap - m_tap_a * s.KW;
        }
    }
    const int bn = tile_n + (tid & (TN - 1));
    const bool bn_ok = bn < N;
    const int krow = tid >> 6;  // 0..3; this thread stages k = krow + 4*i

    const int tx = tid & 15, ty = tid >> 4;
    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

    for (int k0 = k_begin; k0 < k_end; k0 += TK) {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int kk = krow + 4 * i;
            const int k = k0 + kk;
            float av = 0.f, bv = 0.f;
            if (k < k_end) {
                if (MODE == FPROP) {
                    const int tap = k / s.Cin, ci = k - tap * s.Cin;
                    const int a = tap / s.KW, c = tap - a * s.KW;
                    if (am_ok) {
                        const int ih = m_y * s.stride + a - s.pad_t, iw = m_x * s.stride + c - s.pad_l;
                        if (ih >= 0 && ih < s.H && iw >= 0 && iw < s.W)
                            av = A_src[((size_t)(m_b * s.H + ih) * s.W + iw) * s.Cin + ci];
                    }
                    if (bn_ok) bv = B_src[(size_t)k * s.Cout + bn];
                } else if (MODE == DGRAD) {
                    const int t = k / s.Cout, co = k - t * s.Cout;
                    const int ta = t / nc, tc = t - ta * nc;
                    const int a = a0 + s.stride * ta, c = c0 + s.stride * tc;
                    if (am_ok) {
                        const int ny = m_y + s.pad_t - a, nx = m_x + s.pad_l - c;  // divisible by stride
                        if (ny >= 0 && nx >= 0) {
                            const int oh = ny / s.stride, ow = nx / s.stride;
                            if (oh < s.OH && ow < s.OW)
                                av = A_src[((size_t)(m_b * s.OH + oh) * s.OW + ow) * s.Cout + co];
                        }
                    }
                    if (bn_ok) bv = B_src[((size_t)(a * s.KW + c) * s.Cin + bn) * s.Cout + co];
                } else {
                    const int b = k / (s.OH * s.OW);
                    const int r = k - b * s.OH * s.OW;
                    const int oh = r / s.OW, ow = r - oh * s.OW;
                    if (am_ok) {
                        const int ih = oh * s.stride + m_tap_a - s.pad_t, iw = ow * s.stride + m_tap_c - s.pad_l;
                        if (ih >= 0 && ih < s.H && iw >= 0 && iw < s.W)
                            av = A_src[((size_t)(b * s.H + ih) * s.W + iw) * s.Cin + m_ci];
                    }
                    if (bn_ok) bv = B_src[(size_t)k * s.Cout + bn];
                }
            }
            As[kk][tid & (TM - 1)] = av;
            Bs[kk][tid & (TN - 1)] = bv;
        }
        __syncthreads();
#pragma unroll
        for (int kk = 0; kk < TK; ++kk) {
            const float4 a4 = *reinterpret_cast<const float4*>(&As[kk][ty * 4]);
            const float4 b4 = *reinterpret_cast<const float4*>(&Bs[kk][tx * 4]);
            const float a[4] = {a4.x, a4.y, a4.z, a4.w};
            const float b[4] = {b4.x, b4.y, b4.z, b4.w};
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
        }
        __syncthreads();
    }

    // ---- epilogue ---------------------------------------------------------------------------
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int m = tile_m + ty * 4 + i;
        if (m >= M) continue;
        size_t row_off;
        if (MODE == DGRAD) {
            const int b = m / (Hp * Wp); const int r = m - b * Hp * Wp;
            const int ih = (r / Wp) * s.stride + ph, iw = (r % Wp) * s.stride + pw;
            row_off = ((size_t)(b * s.H + ih) * s.W + iw) * s.Cin;
        } else {
            row_off = (size_t)m * N;
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int n = tile_n + tx * 4 + j;
            if (n >= N) continue;
            if (MODE == WGRAD) atomicAdd(&C_dst[row_off + n], acc[i][j]);
            else C_dst[row_off + n] = acc[i][j];
        }
    }
}

int check_shape(const acg_conv_shape* s, const char* who) {
    ACG_REQUIRE(s, ACG_ERR_INVALID, "%s: null shape", who);
    ACG_REQUIRE(s->B > 0 && s->H > 0 && s->W > 0 && s->Cin > 0 && s->OH > 0 && s->OW > 0 && s->Cout > 0 &&
                    s->KH > 0 && s->KW > 0 && s->pad_t >= 0 && s->pad_l >= 0,
                ACG_ERR_INVALID, "%s: non-positive extent", who);
    ACG_REQUIRE(s->stride == 1 || s->stride == 2, ACG_ERR_UNSUPPORTED, "%s: stride %d", who, s->stride);
    ACG_REQUIRE((long long)s->B * s->H * s->W * s->Cin < (1ll << 31) &&
                    (long long)s->B * s->OH * s->OW * s->Cout < (1ll << 31),
                ACG_ERR_UNSUPPORTED, "%s: tensor too large for 32-bit pixel indexing", who);
    // the last tap of the last output pixel may read at most into the padding, never past a row of zeros
    ACG_REQUIRE((s->OH - 1) * s->stride - s->pad_t < s->H && (s->OW - 1) * s->stride - s->pad_l < s->W,
                ACG_ERR_INVALID, "%s: output larger than the input allows", who);
    return ACG_OK;
}

}  // namespace
}  // namespace acg

extern "C" {

int acg_conv_fprop_f32(const acg_conv_shape* s, const float* x, const float* w, float* y, void* stream) {
    using namespace acg;
    int rc = check_shape(s, "acg_conv_fprop_f32");
    if (rc) return rc;
    ACG_REQUIRE(x && w && y, ACG_ERR_INVALID, "acg_conv_fprop_f32: null pointer");
    const int M = s->B * s->OH * s->OW;
    dim3 grid((M + TM - 1) / TM, (s->Cout + TN - 1) / TN, 1);
    launch_pdl(conv_simt_kernel<FPROP>, grid, kThreads, 0, static_cast<cudaStream_t>(stream), x, w, y, *s);
    return check_launch("acg_conv_fprop_f32");
}

int acg_conv_dgrad_f32(const acg_conv_shape* s, const float* dy, const float* w, float* dx, void* stream) {
    using namespace acg;
    int rc = check_shape(s, "acg_conv_dgrad_f32");
    if (rc) return rc;
    ACG_REQUIRE(dy && w && dx, ACG_ERR_INVALID, "acg_conv_dgrad_f32: null pointer");
    const int Hp = (s->H + s->stride - 1) / s->stride, Wp = (s->W + s->stride - 1) / s->stride;
    const int M = s->B * Hp * Wp;  // largest parity class
    dim3 grid((M + TM - 1) / TM, (s->Cin + TN - 1) / TN, s->stride * s->stride);
    launch_pdl(conv_simt_kernel<DGRAD>, grid, kThreads, 0, static_cast<cudaStream_t>(stream), dy, w, dx, *s);
    return check_launch("acg_conv_dgrad_f32");
}

int acg_conv_wgrad_f32(const acg_conv_shape* s, const float* x, const float* dy, float* dw, void* stream) {
    using namespace acg;
    int rc = check_shape(s, "acg_conv_wgrad_f32");
    if (rc) return rc;
    ACG_REQUIRE(x && dy && dw, ACG_ERR_INVALID, "acg_conv_wgrad_f32: null pointer");
    const int M = s->KH * s->KW * s->Cin;
    const long long Kd = (long long)s->B * s->OH * s->OW;
    const int gx = (M + TM - 1) / TM, gy = (s->Cout + TN - 1) / TN;
    // split the pixel reduction so that the grid fills the machine (>= 2 waves), >= 256 pixels per slice
    long long gz = ((long long)num_sms() * 4 + (long long)gx * gy - 1) / ((long long)gx * gy);
    const long long max_gz = (Kd + 255) / 256;
    if (gz > max_gz) gz = max_gz;
    if (gz < 1) gz = 1;
    if (gz > 65535) gz = 65535;
    dim3 grid(gx, gy, (unsigned)gz);
    launch_pdl(conv_simt_kernel<WGRAD>, grid, kThreads, 0, static_cast<cudaStream_t>(stream), x, dy, dw, *s);
    return check_launch("acg_conv_wgrad_f32");
}

}  // extern "C"
