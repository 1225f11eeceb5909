// Dynamic Neural Advection (DNA) transform, forward and backward, for sm_100a.
//
// Replaces models.py:60-72 of the reference (tf.nn.softmax over the K*K logits of every pixel,
// tf.extract_image_patches of the input frame with SAME zero padding, stack/multiply/reduce_sum)
// and TF's autodiff of that sub-graph.  The reference materialises the [B,64,64,K*K,3] patch tensor
// and a 3x stacked softmax in HBM (~15x the algorithmic traffic); here one persistent kernel streams
// each band of R image rows (R = 2; 4 behind ACG_DNA_ROWS) exactly once:
//
//   * warp 0 of the CTA is the producer: per band ONE bulk-copy (TMA engine, SASS UBLKCP) for the band's
//     contiguous logits and one per image row of the band + halo into a padded shared-memory tile, issued
//     by different lanes side by side, all completing on the stage's mbarrier (expect_tx byte counting);
//   * R x 64 consumer threads (one pixel each) read their K*K logits from shared memory (stride K*K words
//     -> conflict free for K=5), do max / exp / sum in registers, accumulate the K*K x 3 weighted
//     neighbourhood from the padded image tile (stride-3 words -> conflict free) and normalise once;
//   * the backward kernel recomputes the softmax, forms g_p = <dy, x_p>, writes
//     dz_p = s_p (g_p - sum_q s_q g_q) IN PLACE over the staged logits and ships the band back to HBM
//     with one shared->global bulk store.
//
// HBM traffic per pixel (fp32): forward (K*K + 3 + 3)*4 B, backward (2*K*K + 3 + 3)*4 B plus the
// halo rows (served from L2: the neighbouring band's CTA has just read them).
#include "common.cuh"

namespace acg {
namespace {

constexpr int kMaxW = 64;
// image rows per band R (template parameter): 4 for the training step's shapes; 2 for SMALL batches (B = 64 is only
// 1024 four-row bands = 3.5 per CTA slot: the launch is then ramp-up + drain around ~5 us of HBM time, and half-size
// bands halve the compute phase that trails the last load and the wait for the first one).  One pixel per thread.
constexpr int kPadCols = 4;              // zero columns left and right (4*12 B keeps rows 16 B aligned)
constexpr int kRowStride = (kMaxW + 2 * kPadCols) * 3;  // floats per staged image row

constexpr int pad16(int v) { return (v + 15) / 16 * 16; }

template <int K, typename LT, bool BWD, int R> struct StageLayout {
    static constexpr int KK = K * K;
    static constexpr int kThreads = R * kMaxW;
    static constexpr int NR = R + K - 1;  // band rows + halo
    static constexpr int logits_bytes = kThreads * KK * (int)sizeof(LT);
    static constexpr int img_bytes = NR * kRowStride * 4;
    static constexpr int dy_bytes = BWD ? kThreads * 3 * 4 : 0;
    static constexpr int off_img = logits_bytes;
    static constexpr int off_dy = off_img + img_bytes;
    static constexpr int bytes = ((off_dy + dy_bytes + 127) / 128) * 128;
};

// PADOUT (backward only): dlogits are written as bf16 with the channel count padded to 16 (zeros in the pad) into
// a separate double-buffered shared-memory tile -- the layout the tensor-core convolutions consume -- instead of
// in place over the staged logits.
template <int K, typename LT, bool BWD, int STAGES, bool PADOUT, int R>
__global__ void __launch_bounds__(R * kMaxW)
dna_kernel(const LT* __restrict__ logits, const float* __restrict__ img, const float* __restrict__ dy,
           float* __restrict__ out, void* __restrict__ dlogits_v, int B, int H, int W, int issue_lanes) {
    // PDL: dependents may be scheduled at once; the wait for the producer grid comes AFTER the shared-memory setup below
    // (barrier init, pad zeroing: no global memory involved), so that setup hides under the previous kernel's tail.  At
    // the small bench size (B=64, K=5: 10 us per launch against 5 us of HBM time) the prologue was 10 % of the kernel.
    pdl_launch_dependents();
    using L = StageLayout<K, LT, BWD, R>;
    constexpr int kRows = R;
    constexpr int kThreads = R * kMaxW;
    constexpr int KK = K * K;
    constexpr int LDO = pad16(KK);
    constexpr int PB = (K - 1) / 2;  // TF SAME: pad_before = (K-1)/2, the odd element goes after
    // forward keeps STAGES-1 bands in flight; the in-place backward one fewer so that the bulk store of the
    // previous band may still be reading its stage while the next load is issued.
    constexpr int PREFETCH = (BWD && !PADOUT) ? STAGES - 2 : STAGES - 1;
    static_assert(PREFETCH >= 1, "need a deeper ring");
    LT* dlogits = static_cast<LT*>(dlogits_v);

    extern __shared__ __align__(128) unsigned char smem[];
    __shared__ __align__(8) uint64_t full_bar[STAGES];

    const int tid = threadIdx.x;
    const int bands_per_img = H / kRows;
    const int nbands = B * bands_per_img;
    const int npx = kRows * W;
    const uint32_t lbytes = (uint32_t)(npx * KK * sizeof(LT));
    const uint32_t rbytes = (uint32_t)(W * 3 * 4);

    if (tid == 0) {
        for (int s = 0; s < STAGES; ++s) mbar_init(&full_bar[s], 1);
        fence_mbar_init();
    }
    // zero the pad columns of every staged image row once (bulk copies never touch them: the first loads, issued by
    // thread 0 below, run concurrently with this loop on disjoint bytes)
    auto zero_pads = [&]() {
    for (int idx = tid; idx < STAGES * L::NR * 2 * kPadCols * 3; idx += kThreads) {
        int s = idx / (L::NR * 2 * kPadCols * 3);
        int rem = idx % (L::NR * 2 * kPadCols * 3);
        int r = rem / (2 * kPadCols * 3);
        int c = rem % (2 * kPadCols * 3);
        float* row = reinterpret_cast<float*>(smem + (size_t)s * L::bytes + L::off_img) + r * kRowStride;
        int col = c < kPadCols * 3 ? c : (kPadCols + W) * 3 + (c - kPadCols * 3);
        row[col] = 0.f;
    }
    };

    // called by every lane of warp 0: lane 0 posts the byte count, then the band's copies are issued by different
    // lanes side by side (one thread issuing 7-11 bulk copies back to back was ~0.3 us of every CTA's ramp-up)
    auto issue = [&](int band, int stage) {
        unsigned char* st = smem + (size_t)stage * L::bytes;
        const int b = band / bands_per_img;
        const int r0 = (band % bands_per_img) * kRows;
        if (tid == 0) {
            int valid = 0;
            for (int t = 0; t < L::NR; ++t) {
                int row = r0 - PB + t;
                valid += (row >= 0 && row < H);
            }
            uint32_t total = lbytes + (uint32_t)valid * rbytes + (BWD ? (uint32_t)npx * 12u : 0u);
            mbar_expect_tx(&full_bar[stage], total);
        }
        if (issue_lanes == 1) {            // A/B: thread 0 issues every copy of the band itself
            bulk_g2s(st, logits + (size_t)band * npx * KK, lbytes, &full_bar[stage]);
            for (int t = 0; t < L::NR; ++t) {
                const int row = r0 - PB + t;
                if (row >= 0 && row < H)
                    bulk_g2s(reinterpret_cast<float*>(st + L::off_img) + t * kRowStride + kPadCols * 3,
                             img + ((size_t)(b * H + row) * W) * 3, rbytes, &full_bar[stage]);
            }
            if (BWD) bulk_g2s(st + L::off_dy, dy + (size_t)band * npx * 3, (uint32_t)npx * 12u, &full_bar[stage]);
            return;
        }
        __syncwarp();
        if (tid == 0) {
            bulk_g2s(st, logits + (size_t)band * npx * KK, lbytes, &full_bar[stage]);
        } else if (tid <= L::NR) {
            const int t = tid - 1;
            const int row = r0 - PB + t;
            if (row >= 0 && row < H)
                bulk_g2s(reinterpret_cast<float*>(st + L::off_img) + t * kRowStride + kPadCols * 3,
                         img + ((size_t)(b * H + row) * W) * 3, rbytes, &full_bar[stage]);
        } else if (BWD && tid == L::NR + 1) {
            bulk_g2s(st + L::off_dy, dy + (size_t)band * npx * 3, (uint32_t)npx * 12u, &full_bar[stage]);
        }
    };

    pdl_wait();                 // the producer grid has completed: global memory may be read from here on
    if (tid < issue_lanes) {
        for (int p = 0; p < PREFETCH; ++p) {
            int band = blockIdx.x + p * gridDim.x;
            if (band < nbands) issue(band, p % STAGES);
        }
    }
    zero_pads();
    __syncthreads();

    const int il = tid / W;   // row of this thread's pixel inside the band
    const int j = tid - il * W;
    const bool active = tid < npx;

    for (int it = 0;; ++it) {
        const int band = blockIdx.x + it * gridDim.x;
        if (band >= nbands) break;
        const int stage = it % STAGES;
        if (tid < issue_lanes) {
            int nb = band + PREFETCH * gridDim.x;
            if (nb < nbands) {
                // the store that last read this stage has drained (thread 0 committed it; the __syncwarp inside
                // issue() orders the other lanes' copies behind this wait)
                if (BWD && !PADOUT && tid == 0) bulk_wait_read<1>();
                issue(nb, (it + PREFETCH) % STAGES);
            }
        }
        mbar_wait(&full_bar[stage], (uint32_t)((it / STAGES) & 1));

        unsigned char* st = smem + (size_t)stage * L::bytes;
        LT* lg = reinterpret_cast<LT*>(st);
        const float* tile = reinterpret_cast<const float*>(st + L::off_img);
        const int r0 = (band % bands_per_img) * kRows;

        float e[KK];
        if (active) {
            float m = -INFINITY;
#pragma unroll
            for (int p = 0; p < KK; ++p) {
                e[p] = ld_as_float<LT>(lg, (size_t)tid * KK + p);
                m = fmaxf(m, e[p]);
            }
            float sum = 0.f;
#pragma unroll
            for (int p = 0; p < KK; ++p) {
                e[p] = __expf(e[p] - m);
                sum += e[p];
            }
            const float inv = 1.f / sum;
            if (!BWD) {
                float a0 = 0.f, a1 = 0.f, a2 = 0.f;
#pragma unroll
                for (int a = 0; a < K; ++a) {
                    const int row = r0 + il + a - PB;
                    if (row < 0 || row >= H) continue;  // uniform per warp pair: SAME zero rows
                    const float* rp = tile + (il + a) * kRowStride + (kPadCols + j - PB) * 3;
#pragma unroll
                    for (int c = 0; c < K; ++c) {
                        const float wgt = e[a * K + c];
                        a0 = fmaf(wgt, rp[c * 3 + 0], a0);
                        a1 = fmaf(wgt, rp[c * 3 + 1], a1);
                        a2 = fmaf(wgt, rp[c * 3 + 2], a2);
                    }
                }
                float* o = out + ((size_t)band * npx + tid) * 3;
                o[0] = a0 * inv;
                o[1] = a1 * inv;
                o[2] = a2 * inv;
            } else {
                const float* dyt = reinterpret_cast<const float*>(st + L::off_dy) + tid * 3;
                const float d0 = dyt[0], d1 = dyt[1], d2 = dyt[2];
                float g[KK];
                float dot = 0.f;
#pragma unroll
                for (int a = 0; a < K; ++a) {
                    const int row = r0 + il + a - PB;
                    const bool rv = (row >= 0 && row < H);
                    const float* rp = tile + (il + a) * kRowStride + (kPadCols + j - PB) * 3;
#pragma unroll
                    for (int c = 0; c < K; ++c) {
                        float gv = 0.f;
                        if (rv) gv = fmaf(d0, rp[c * 3 + 0], fmaf(d1, rp[c * 3 + 1], d2 * rp[c * 3 + 2]));
                        g[a * K + c] = gv;
                        e[a * K + c] *= inv;
                        dot = fmaf(e[a * K + c], gv, dot);
                    }
                }
                if (!PADOUT) {
#pragma unroll
                    for (int p = 0; p < KK; ++p) st_from_float<LT>(lg, (size_t)tid * KK + p, e[p] * (g[p] - dot));
                    fence_proxy_async();  // make the in-place dlogits visible to the bulk-copy engine
                } else {
#pragma unroll
                    for (int p = 0; p < KK; ++p) e[p] *= (g[p] - dot);
                }
            }
        }
        if (BWD && PADOUT) {
            // the bulk store that read this out tile two bands ago must have drained before it is overwritten
            if (tid == 0) bulk_wait_read<1>();
            __syncthreads();
            __nv_bfloat16* ot = reinterpret_cast<__nv_bfloat16*>(smem + (size_t)STAGES * L::bytes) +
                                (size_t)(it & 1) * kThreads * LDO;
            if (active) {
                __nv_bfloat16* row = ot + (size_t)tid * LDO;
#pragma unroll
                for (int p = 0; p < LDO; p += 2) {
                    const float v0 = p < KK ? e[p < KK ? p : 0] : 0.f;
                    const float v1 = p + 1 < KK ? e[p + 1 < KK ? p + 1 : 0] : 0.f;
                    *reinterpret_cast<__nv_bfloat162*>(row + p) = __floats2bfloat162_rn(v0, v1);
                }
                fence_proxy_async();
            }
            __syncthreads();
            if (tid == 0) {
                bulk_s2g(static_cast<__nv_bfloat16*>(dlogits_v) + (size_t)band * npx * LDO, ot, (uint32_t)(npx * LDO * 2));
                bulk_commit();
            }
            continue;
        }
        __syncthreads();  // every thread is done with this stage
        if (BWD && tid == 0) {
            bulk_s2g(dlogits + (size_t)band * npx * KK, lg, lbytes);
            bulk_commit();
        }
    }
    if (BWD && tid == 0) bulk_wait<0>();  // smem must stay alive until the last store has read it
}

template <int K, typename LT, bool BWD, int STAGES, bool PADOUT, int R>
int launch_r(const void* logits, const float* img, const float* dy, float* out, void* dlogits, int B, int H,
             int W, cudaStream_t stream) {
    using L = StageLayout<K, LT, BWD, R>;
    constexpr int kRows = R;
    constexpr int kThreads = R * kMaxW;
    auto kern = dna_kernel<K, LT, BWD, STAGES, PADOUT, R>;
    const int smem = STAGES * L::bytes + (PADOUT ? 2 * kThreads * pad16(K * K) * 2 : 0);
    static int occ_dev[64] = {0};  // per template instance and device (the smem attribute is per device)
    int dev = 0;
    cudaGetDevice(&dev);
    int& occ = occ_dev[dev & 63];
    if (occ == 0) {
        if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem) != cudaSuccess) {
            cudaError_t e = cudaGetLastError();
            set_error("acg_dna: cannot set %d B dynamic smem: %s", smem, cudaGetErrorString(e));
            return ACG_ERR_CUDA;
        }
        int o = 0;
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&o, kern, kThreads, smem) != cudaSuccess || o < 1) {
            cudaGetLastError();
            o = 1;
        }
        occ = o;
    }
    const int nbands = B * (H / kRows);
    int grid = num_sms() * occ;
    if (grid > nbands) grid = nbands;
    // every CTA walks the same number of bands: 1024 bands over 296 CTAs left 136 CTAs alone on the fourth pass (a
    // quarter of the kernel at half the memory parallelism); 256 CTAs x 4 bands finish together
    if (!getenv("ACG_DNA_UNBALANCED")) {
        const int passes = (nbands + grid - 1) / grid;
        grid = (nbands + passes - 1) / passes;
    }
#ifdef ACG_PROBES
    if (const char* e = getenv("ACG_DNA_GRID")) {
        const int g = atoi(e);
        if (g > 0) grid = g < nbands ? g : nbands;
    }
#endif
    const char* il = getenv("ACG_DNA_ISSUE_LANES");
    launch_pdl(kern, grid, kThreads, smem, stream, static_cast<const LT*>(logits), img, dy, out, dlogits, B, H, W,
               il && atoi(il) == 1 ? 1 : 32);
    return check_launch(BWD ? "acg_dna_bwd" : "acg_dna_fwd");
}

// Band height: 2 rows at every batch size (measured equal or better than 4 from B = 16 to 256: the half-size bands
// shorten ramp-up and drain, and 3-4 CTAs of 128 threads per SM overlap better than 2 of 256).  ACG_DNA_ROWS = 2 | 4
// forces one (A/B runs, tests of both variants).
int band_rows(int B, int H) {
    if (const char* e = getenv("ACG_DNA_ROWS")) {
        const int r = atoi(e);
        if (r == 2 || r == 4) return r;
    }
    (void)B; (void)H;
    return 2;
}

template <int K, typename LT, bool BWD, int STAGES, bool PADOUT = false>
int launch(const void* logits, const float* img, const float* dy, float* out, void* dlogits, int B, int H,
           int W, cudaStream_t stream) {
#ifdef ACG_PROBES
    // probe library only (scripts/dna_sweep.py): ring depth as a run-time choice
    if (const char* e = getenv("ACG_DNA_STAGES")) {
        const int st = atoi(e), r = band_rows(B, H);
        constexpr int kMin = (BWD && !PADOUT) ? 3 : 2;
#define ACG_DNA_TRY(S, R)                                                                                  \
        if constexpr ((S) >= kMin) {                                                                           \
            if (st == (S) && r == (R))                                                                         \
                return launch_r<K, LT, BWD, (S), PADOUT, (R)>(logits, img, dy, out, dlogits, B, H, W, stream); \
        }
        ACG_DNA_TRY(2, 2) ACG_DNA_TRY(3, 2) ACG_DNA_TRY(4, 2) ACG_DNA_TRY(5, 2)
        ACG_DNA_TRY(2, 4) ACG_DNA_TRY(3, 4) ACG_DNA_TRY(4, 4)
#undef ACG_DNA_TRY
    }
#endif
    if (band_rows(B, H) == 2) return launch_r<K, LT, BWD, STAGES, PADOUT, 2>(logits, img, dy, out, dlogits, B, H, W, stream);
    return launch_r<K, LT, BWD, STAGES, PADOUT, 4>(logits, img, dy, out, dlogits, B, H, W, stream);
}

int validate(const void* logits, const float* img, int B, int H, int W, int C, int K, int dtype) {
    ACG_REQUIRE(logits && img, ACG_ERR_INVALID, "acg_dna: null pointer");
    ACG_REQUIRE(B > 0 && H > 0 && W > 0, ACG_ERR_INVALID, "acg_dna: non-positive size");
    ACG_REQUIRE(C == 3, ACG_ERR_UNSUPPORTED, "acg_dna: C=%d (only 3 colour channels)", C);
    ACG_REQUIRE(K == 5 || K == 6, ACG_ERR_UNSUPPORTED, "acg_dna: K=%d (only 5 or 6)", K);
    ACG_REQUIRE(W <= kMaxW && W % 4 == 0 && H % 4 == 0, ACG_ERR_UNSUPPORTED,
                "acg_dna: H=%d W=%d (need H%%4==0, W%%4==0, W<=64)", H, W);
    ACG_REQUIRE(dtype == ACG_F32 || dtype == ACG_BF16, ACG_ERR_UNSUPPORTED, "acg_dna: logits dtype %d", dtype);
    ACG_REQUIRE(((uintptr_t)logits % 16) == 0 && ((uintptr_t)img % 16) == 0, ACG_ERR_INVALID,
                "acg_dna: buffers must be 16-byte aligned");
    return ACG_OK;
}

}  // namespace
}  // namespace acg

extern "C" {

int acg_dna_fwd(const void* logits, int logits_dtype, const float* img, float* out, int B, int H, int W, int C,
                int K, void* stream) {
    using namespace acg;
    int rc = validate(logits, img, B, H, W, C, K, logits_dtype);
    if (rc) return rc;
    ACG_REQUIRE(out, ACG_ERR_INVALID, "acg_dna_fwd: null out");
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    if (logits_dtype == ACG_F32) {
        if (K == 5) return launch<5, float, false, 3>(logits, img, nullptr, out, nullptr, B, H, W, s);
        return launch<6, float, false, 2>(logits, img, nullptr, out, nullptr, B, H, W, s);
    }
    if (K == 5) return launch<5, __nv_bfloat16, false, 3>(logits, img, nullptr, out, nullptr, B, H, W, s);
    return launch<6, __nv_bfloat16, false, 3>(logits, img, nullptr, out, nullptr, B, H, W, s);
}

int acg_dna_bwd(const void* logits, int logits_dtype, const float* img, const float* dy, void* dlogits,
                int dlogits_dtype, int ld_dlogits, int B, int H, int W, int C, int K, void* stream) {
    using namespace acg;
    int rc = validate(logits, img, B, H, W, C, K, logits_dtype);
    if (rc) return rc;
    const bool dense = dlogits_dtype == logits_dtype && ld_dlogits == K * K;
    const bool padded = logits_dtype == ACG_F32 && dlogits_dtype == ACG_BF16 && ld_dlogits == pad16(K * K);
    ACG_REQUIRE(dense || padded, ACG_ERR_UNSUPPORTED,
                "acg_dna_bwd: dlogits must be dense in the logits dtype, or bf16 with ld = K*K rounded up to 16");
    if (padded) {
        ACG_REQUIRE(dy && dlogits && ((uintptr_t)dy % 16) == 0 && ((uintptr_t)dlogits % 16) == 0, ACG_ERR_INVALID,
                    "acg_dna_bwd: null or misaligned pointer");
        cudaStream_t s2 = static_cast<cudaStream_t>(stream);
        if (K == 5) return launch<5, float, true, 3, true>(logits, img, dy, nullptr, dlogits, B, H, W, s2);
        return launch<6, float, true, 3, true>(logits, img, dy, nullptr, dlogits, B, H, W, s2);
    }
    ACG_REQUIRE(dy && dlogits, ACG_ERR_INVALID, "acg_dna_bwd: null pointer");
    ACG_REQUIRE(((uintptr_t)dy % 16) == 0 && ((uintptr_t)dlogits % 16) == 0, ACG_ERR_INVALID,
                "acg_dna_bwd: buffers must be 16-byte aligned");
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    if (logits_dtype == ACG_F32) {
        if (K == 5) return launch<5, float, true, 3>(logits, img, dy, nullptr, dlogits, B, H, W, s);
        return launch<6, float, true, 3>(logits, img, dy, nullptr, dlogits, B, H, W, s);
    }
    if (K == 5) return launch<5, __nv_bfloat16, true, 3>(logits, img, dy, nullptr, dlogits, B, H, W, s);
    return launch<6, __nv_bfloat16, true, 3>(logits, img, dy, nullptr, dlogits, B, H, W, s);
}

}  // extern "C"
