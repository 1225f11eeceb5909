// Pixel-major tcgen05 implicit GEMM for the layers with SMALL feature maps (4x4 / 2x2 / 8x8 grids: g/conv4, g/tconv1,
// d/conv4, d/conv5 and their data gradients; models.py:36-41,82-87 of the reference).
//
// The generic kernel (conv_tc.cu) makes a GEMM row of every (image, output pixel) and gathers the A operand row by row
// with 16-byte cp.async copies -- on these layers that gather is what the launch waits for (few tiles, K up to 6400,
// 200-300 TFLOP/s), and on a 2x2 / 4x4 grid about half of the gathered rows are zero padding that is still multiplied.
// Here a GEMM tile is ONE output pixel x 128 IMAGES:
//   * for a fixed output pixel and filter tap, the A rows of all images sit at the same (y, x) of the source tensor, so
//     the whole 128 x 64 operand slice is ONE 4-D TMA box {64 channels, 1, 1, 128 images} (hardware swizzle, no
//     per-row address arithmetic, zero fill past the batch);
//   * a tap that falls outside the image does so for ALL rows of the tile: it is skipped, not multiplied by zeros
//     (49 % of the MMAs of d/conv5, 72 % of d/conv4's remain);
//   * weights arrive as in the generic kernel (2-D TMA of the [N][tap * ld + channel] pack, K offset of the tap).
// Both gather forms: CONV (forward of conv2d / data gradient of conv2d_transpose) and ADJ (the adjoint: data gradient
// of conv2d / forward of conv2d_transpose, one weight matrix per output-parity class).
// Work item = (output pixel, batch tile of 128, N tile of 128[, K split]); warps 0-3 epilogue, warp 4 MMA issuer and
// TMEM owner, warp 5 TMA producer; 6-stage ring of (16 KB A + 16 KB B) slices; split-K through the generic kernel's
// workspace scheme (partial tiles parked in L2, last CTA of a tile adds them in split order).
#include "conv_tc.cuh"

namespace acg {
namespace tc {

constexpr int kPxStages = 6;
constexpr int kPxThreads = 192;
constexpr int kPxSmem = kPxStages * (kStageA + kStageB) + 4 * kEpiStageBytes + 1024;   // + 4 epilogue staging tiles
constexpr int kPxMaxTaps = 32;
constexpr int kPxMaxSplits = 4;

struct alignas(64) PxParams {
    Params p;
    CUtensorMap map_a;        // source activations as (C, X, Y, B); box {64, 1, 1, 128}, 128B swizzle
    CUtensorMap map_b[4];     // weights: CONV entry 0 / ADJ one per parity class; box {64, min(N,128)}
    int form;                 // 0 CONV, 1 ADJ
    int btiles;               // ceil(B / 128)
    int SX, SY;               // source grid (CONV: W, H of x; ADJ: OW, OH of dy)
    int DX, DY;               // destination grid (CONV: OW, OH; ADJ: W, H)
    int kchunks;              // 64-channel K slices per tap
    int nk16_last;            // K=16 MMAs of a tap's last slice
};

__global__ void __launch_bounds__(kPxThreads, 1) conv_px_kernel(const __grid_constant__ PxParams pp) {
    const Params& p = pp.p;
    extern __shared__ unsigned char smem_raw[];
    __shared__ __align__(8) uint64_t full_bar[kPxStages], empty_bar[kPxStages], acc_bar;
    __shared__ uint32_t tmem_base_sh;
    __shared__ float sm_stats[4][2][BN];       // per epilogue warp
    __shared__ int last_cta_sh;
    __shared__ int tap_x[kPxMaxTaps], tap_y[kPxMaxTaps], tap_k[kPxMaxTaps], ntap_sh;

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t smemA = smem_base, smemB = smem_base + kPxStages * kStageA;
    const uint32_t smemE = smemB + kPxStages * kStageB;        // per epilogue warp: 32 rows x 128 B staging tile
    for (int i = tid; i < 4 * 2 * BN; i += kPxThreads) (&sm_stats[0][0][0])[i] = 0.f;

    // ---- geometry of this CTA: output pixel (py, px), images [b0, b0 + 128), columns [n0, n0 + n_cta) ----
    const int s = p.stride;
    const int bt = (int)(blockIdx.x % (unsigned)pp.btiles), pix = (int)(blockIdx.x / (unsigned)pp.btiles);
    const int py = pix / pp.DX, px = pix - py * pp.DX;
    const int b0 = bt * BM;
    const int n0 = blockIdx.y * BN;
    const int n_cta = min(BN, p.N - n0);
    const int split = (int)blockIdx.z;
    const int cls = pp.form == 0 ? 0 : (py % s) * s + (px % s);
    if (warp == 0) {
        // valid taps of this pixel, compacted: source pixel and K offset inside the weight matrix
        int total, nc, by, bx, dir;
        if (pp.form == 0) {
            total = p.KH * p.KW; nc = p.KW; dir = 1;
            by = py * s - p.pad_t; bx = px * s - p.pad_l;
        } else {
            const int a0 = (py % s + p.pad_t) % s, c0 = (px % s + p.pad_l) % s;
            const int na = a0 < p.KH ? (p.KH - a0 + s - 1) / s : 0;
            nc = c0 < p.KW ? (p.KW - c0 + s - 1) / s : 0;
            total = na * nc; dir = -1;
            by = (py + p.pad_t - a0) / s; bx = (px + p.pad_l - c0) / s;     // source pixel of the class' tap (0,0)
            if (nc == 0) nc = 1;
        }
        const int ta = lane / nc, tcx = lane - ta * nc;
        const int sy = by + dir * ta, sx = bx + dir * tcx;
        const bool valid = lane < total && sy >= 0 && sy < pp.SY && sx >= 0 && sx < pp.SX;
        const uint32_t mask = __ballot_sync(0xffffffffu, valid);
        if (valid) {
            const int pos = __popc(mask & ((1u << lane) - 1u));
            tap_x[pos] = sx; tap_y[pos] = sy; tap_k[pos] = lane * p.lda;
        }
        if (lane == 0) ntap_sh = __popc(mask);
    }
    if (tid == 0) {
        for (int i = 0; i < kPxStages; ++i) { mbar_init(&full_bar[i], 1); mbar_init(&empty_bar[i], 1); }
        mbar_init(&acc_bar, 1);
        fence_mbar_init();
        tma_prefetch_desc(&pp.map_a);
        tma_prefetch_desc(&pp.map_b[cls]);
    }
    if (warp == 4) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_sh)),
                     "r"((uint32_t)BN) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    if (p.px.world <= 1) pdl_launch_dependents();     // a launch that exchanges with peers must not (peer.cu)
    pdl_wait();
    const uint32_t tmem_base = tmem_base_sh;
    const int nkb_all = ntap_sh * pp.kchunks;
    const int kb_per = p.splits > 1 ? (nkb_all + p.splits - 1) / p.splits : nkb_all;
    const int kb0 = split * kb_per;
    const int nkb = max(0, min(nkb_all - kb0, kb_per));

    if (warp == 5) {
        // ================================ TMA producer ================================
        const CUtensorMap* bmap = &pp.map_b[cls];
        const uint32_t tx = (uint32_t)kStageA + (uint32_t)min(p.N, BN) * 128u;
        int t = kb0 / pp.kchunks, ch = kb0 - t * pp.kchunks;
        for (int kb = 0; kb < nkb; ++kb) {
            const int stage = kb % kPxStages;
            if (kb >= kPxStages) mbar_wait(&empty_bar[stage], (uint32_t)(((kb / kPxStages) - 1) & 1));
            if (elect_one()) {
                // probe bits 1 / 2: no activation / no weight traffic
                const uint32_t txa = ACG_DBG(p, 1) ? 0u : (uint32_t)kStageA, txb = ACG_DBG(p, 2) ? 0u : tx - (uint32_t)kStageA;
                if (txa + txb == 0u) mbar_arrive(&full_bar[stage]);
                else mbar_expect_tx(&full_bar[stage], txa + txb);
                if (txa) tma_load_4d(smemA + stage * kStageA, &pp.map_a, ch * BK, tap_x[t], tap_y[t], b0, &full_bar[stage]);
                if (txb) tma_load_2d(smemB + stage * kStageB, bmap, tap_k[t] + ch * BK, n0, &full_bar[stage]);
            }
            __syncwarp();
            if (++ch == pp.kchunks) { ch = 0; ++t; }
        }
    } else if (warp == 4) {
        // ================================ MMA issuer ================================
        const uint32_t idesc = make_idesc(n_cta, 0, 0);
        const uint32_t hi = desc_hi(1024);
        const uint32_t alo0 = desc_lo(smemA, 16), blo0 = desc_lo(smemB, 16);
        const uint32_t tmem_u = __reduce_or_sync(0xffffffffu, tmem_base);
        int ch = kb0 % pp.kchunks;
        for (int kb = 0; kb < nkb; ++kb) {
            const int stage = kb % kPxStages;
            mbar_wait(&full_bar[stage], (uint32_t)((kb / kPxStages) & 1));
            tc_fence_after();
            const uint32_t alo = alo0 + stage * (kStageA >> 4), blo = blo0 + stage * (kStageB >> 4);
            const int nk = ch == pp.kchunks - 1 ? pp.nk16_last : BK / 16;
            if (elect_one()) {
                if (!ACG_DBG(p, 4))                                                // probe: no MMAs
                    for (int k = 0; k < nk; ++k)
                        tc_mma2(tmem_u, alo + 2 * k, hi, blo + 2 * k, hi, idesc, (kb | k) != 0 ? 1u : 0u);
                tc_commit(&empty_bar[stage]);
            }
            __syncwarp();
            if (++ch == pp.kchunks) ch = 0;
        }
        if (elect_one()) tc_commit(&acc_bar);
        __syncwarp();
    }

    // ================================ epilogue (warps 0-3) ================================
    if (warp < 4 && nkb > 0) {
        mbar_wait(&acc_bar, 0);
        tc_fence_after();
    }
    // split-K: park the partial tile, take a ticket; only the last CTA of the tile goes on to the epilogue
    bool final_cta = true;
    float* ws_tile = nullptr;
    if (p.splits > 1) {
        const unsigned int tile_id = blockIdx.y * gridDim.x + blockIdx.x;
        ws_tile = p.ws + (size_t)tile_id * p.splits * (BM * BN);
        if (warp < 4) {
            float* mine = ws_tile + (size_t)split * (BM * BN) + (warp * 32 + lane) * 16;
            for (int cb = 0; cb < n_cta; cb += 16) {
                uint32_t v[16];
                if (nkb > 0) tmem_ld16(tmem_base + ((uint32_t)(warp * 32) << 16) + cb, v);
                else {
#pragma unroll
                    for (int i = 0; i < 16; ++i) v[i] = 0u;
                }
                float4* o = reinterpret_cast<float4*>(mine + (size_t)(cb >> 4) * (BM * 16));   // [chunk][row][16]
#pragma unroll
                for (int i = 0; i < 4; ++i)
                    o[i] = make_float4(__uint_as_float(v[4 * i]), __uint_as_float(v[4 * i + 1]),
                                       __uint_as_float(v[4 * i + 2]), __uint_as_float(v[4 * i + 3]));
            }
            __threadfence();
        }
        __syncthreads();
        if (tid == 0) {
            const unsigned int tk = atomicAdd(&p.tickets[tile_id], 1u);
            last_cta_sh = (tk == (unsigned)p.splits - 1u);
            if (last_cta_sh) p.tickets[tile_id] = 0u;       // ready for the next launch
        }
        __syncthreads();
        final_cta = last_cta_sh != 0;
        if (final_cta) __threadfence();
    }
    if (warp < 4 && final_cta) {
        const int b = b0 + warp * 32 + lane;
        const bool row_ok = b < p.B;
        const size_t row_off = ((size_t)((size_t)b * pp.DY + py) * pp.DX + px) * p.ldo;
        // full 128-column bf16 tiles without bias / activation go out through the warp's staging tile: column-wise
        // moments instead of butterflies, full-line stores (conv_tc.cuh, epi_stage_*)
        const bool staged = p.out_dtype == ACG_BF16 && !p.bias && p.out_act == ACG_ACT_NONE && (p.ldo & 7) == 0 &&
                            n_cta == BN && n0 + BN <= p.n_store && (!p.stats || n0 + BN <= p.n_stat) && !p.direct_store &&
                            (reinterpret_cast<uintptr_t>(p.out) & 15) == 0;
        const uint32_t stile = smemE + (uint32_t)warp * kEpiStageBytes;
        unsigned char* out_tile = static_cast<unsigned char*>(p.out) + (size_t)n0 * 2;
        auto emit = [&](const uint32_t (&v)[16], int cb) {      // one 16-column chunk of this lane's row
            if (!staged) {
                epilogue_chunk(p, v, n0 + cb, row_ok, row_off, 0u, lane, &sm_stats[warp][0][cb], &sm_stats[warp][1][cb]);
                return;
            }
            const int ch = (cb >> 4) & 3;
            epi_stage_put_chunk<128>(stile, lane, ch, v, row_ok);
            if (ch == 3)
                epi_stage_moments_flush<128>(stile, lane, out_tile + (size_t)(cb - 48) * 2, (unsigned long long)row_off * 2ull,
                                             row_ok, p.stats != nullptr, &sm_stats[warp][0][cb - 48], &sm_stats[warp][1][cb - 48]);
        };
        if (p.splits > 1) {
            // all partial tiles in split order (this CTA's own one included: the sum does not depend on which CTA came
            // last).  The 16-column chunk after the current one is already in flight while this one goes through the
            // epilogue (one exposed L2 latency per tile instead of one per chunk); splits <= kPxMaxSplits.
            const float* mine = ws_tile + (warp * 32 + lane) * 16;
            float4 nxt[kPxMaxSplits][4];
#pragma unroll
            for (int j = 0; j < kPxMaxSplits; ++j)
                if (j < p.splits) {
                    const float4* o = reinterpret_cast<const float4*>(mine + (size_t)j * (BM * BN));
#pragma unroll
                    for (int i = 0; i < 4; ++i) nxt[j][i] = __ldcg(o + i);
                }
            for (int cb = 0; cb < n_cta; cb += 16) {
                float f[16];
#pragma unroll
                for (int i = 0; i < 16; ++i) f[i] = 0.f;
#pragma unroll
                for (int j = 0; j < kPxMaxSplits; ++j)
                    if (j < p.splits) {
#pragma unroll
                        for (int i = 0; i < 4; ++i) {
                            f[4 * i] += nxt[j][i].x; f[4 * i + 1] += nxt[j][i].y;
                            f[4 * i + 2] += nxt[j][i].z; f[4 * i + 3] += nxt[j][i].w;
                        }
                    }
                if (cb + 16 < n_cta) {
#pragma unroll
                    for (int j = 0; j < kPxMaxSplits; ++j)
                        if (j < p.splits) {
                            const float4* o = reinterpret_cast<const float4*>(mine + (size_t)j * (BM * BN) +
                                                                              (size_t)((cb + 16) >> 4) * (BM * 16));
#pragma unroll
                            for (int i = 0; i < 4; ++i) nxt[j][i] = __ldcg(o + i);
                        }
                }
                uint32_t v[16];
#pragma unroll
                for (int i = 0; i < 16; ++i) v[i] = __float_as_uint(f[i]);
                emit(v, cb);
            }
        } else {
            for (int cb = 0; cb < n_cta; cb += 16) {
                uint32_t v[16];
                if (nkb > 0) tmem_ld16(tmem_base + ((uint32_t)(warp * 32) << 16) + cb, v);
                else {
#pragma unroll
                    for (int i = 0; i < 16; ++i) v[i] = 0u;
                }
                emit(v, cb);
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 4) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)BN) : "memory");
    }
    if (p.stats && final_cta) {
        const bool has_col = tid < n_cta && n0 + tid < p.n_stat;
        double s0 = 0.0, s1 = 0.0;
        if (has_col) {
            float a = 0.f, b = 0.f;
#pragma unroll
            for (int w = 0; w < 4; ++w) { a += sm_stats[w][0][tid]; b += sm_stats[w][1][tid]; }
            s0 = (double)a;
            s1 = (double)b;
        }
        cta_stats_finish(p, tid, kPxThreads, has_col, n0 + tid, s0, s1, &last_cta_sh);
    }
}

// ---- host side -------------------------------------------------------------------------------------------------------

// form 0: CONV gather (destination grid OH x OW), 1: ADJ (destination grid H x W)
bool px_ok(const acg_conv_shape* s, const acg_tc_args* t, int form) {
    // environment switches are read per call (tests and probes flip them inside one process)
    const char* mp = getenv("ACG_PX_MAXPIX");
    const int max_pix = mp ? atoi(mp) : 64;
    if (getenv("ACG_NO_PX") || t->red_z) return false;
    const int DX = form == 0 ? s->OW : s->W, DY = form == 0 ? s->OH : s->H;
    if (DX * DY > max_pix || s->B < 64) return false;       // below 64 images a 128-row tile is mostly padding
    if (t->ld_in % 8 != 0 || s->KH * s->KW > kPxMaxTaps) return false;
    return true;
}

// work items (tiles) and the longest K loop of a launch
void px_items(const acg_conv_shape* s, int form, int ld_in, int N, long long* items, int* max_nkb) {
    const int DX = form == 0 ? s->OW : s->W, DY = form == 0 ? s->OH : s->H;
    const int btiles = (s->B + BM - 1) / BM;
    *items = (long long)DX * DY * btiles * ((N + BN - 1) / BN);
    int taps = s->KH * s->KW;
    if (form == 1) {
        taps = 0;
        for (int cls = 0; cls < s->stride * s->stride; ++cls) {
            int na, nc;
            class_taps(s, cls, &na, &nc);
            if (na * nc > taps) taps = na * nc;
        }
    }
    *max_nkb = taps * ((ld_in + BK - 1) / BK);
}

// split-K plan of a pixel-major launch: (splits, workspace bytes, tickets)
void px_split_plan(const acg_conv_shape* s, int form, int ld_in, int N, int* splits, long long* ws_bytes, int* tickets) {
    long long items;
    int max_nkb;
    px_items(s, form, ld_in, N, &items, &max_nkb);
    int sp = 1;
    if (!getenv("ACG_NO_SPLITK") && items * 2 <= num_sms() && max_nkb >= 16) {
        sp = (int)(num_sms() / items);
        if (sp > max_nkb / 8) sp = max_nkb / 8;         // at least 8 K slices per CTA on the longest pixel
        if (sp > kPxMaxSplits) sp = kPxMaxSplits;
        if (sp < 2) sp = 1;
    }
    *splits = sp;
    *ws_bytes = sp > 1 ? items * sp * (long long)(BM * BN) * 4 : 0;
    *tickets = sp > 1 ? (int)items : 0;
}

int launch_px(int form, const acg_conv_shape* s, const acg_tc_args* t, const Params& p_in, const void* src,
              const void* w_pack, int N, int Npack, cudaStream_t stream, const char* who) {
    int rc = set_smem((const void*)conv_px_kernel, kPxSmem);
    if (rc) return rc;
    PxParams pp;
    pp.p = p_in;
    pp.form = form;
    pp.btiles = (s->B + BM - 1) / BM;
    pp.SX = form == 0 ? s->W : s->OW;  pp.SY = form == 0 ? s->H : s->OH;
    pp.DX = form == 0 ? s->OW : s->W;  pp.DY = form == 0 ? s->OH : s->H;
    const int lda = t->ld_in;
    pp.kchunks = (lda + BK - 1) / BK;
    pp.nk16_last = ((lda - (pp.kchunks - 1) * BK) + 15) / 16;
    ACG_REQUIRE(((uintptr_t)src & 15) == 0, ACG_ERR_UNSUPPORTED, "%s: source not 16-byte aligned", who);
    {   // source activations [B][SY][SX][lda] bf16, innermost first; one box = 64 channels of one pixel of 128 images
        EncodeTiledFn enc = encode_tiled_fn();
        ACG_REQUIRE(enc, ACG_ERR_CUDA, "%s: cuTensorMapEncodeTiled is not available", who);
        const cuuint64_t dims[4] = {(cuuint64_t)lda, (cuuint64_t)pp.SX, (cuuint64_t)pp.SY, (cuuint64_t)s->B};
        const cuuint64_t strides[3] = {(cuuint64_t)lda * 2, (cuuint64_t)pp.SX * lda * 2, (cuuint64_t)pp.SY * pp.SX * lda * 2};
        const cuuint32_t box[4] = {64, 1, 1, (cuuint32_t)BM};
        const cuuint32_t estr[4] = {1, 1, 1, 1};
        CUresult r = enc(&pp.map_a, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(src), dims, strides, box, estr,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        ACG_REQUIRE(r == CUDA_SUCCESS, ACG_ERR_CUDA, "%s: activation tensor map failed (%d)", who, (int)r);
    }
    if (form == 0) {
        rc = encode_weight_map(&pp.map_b[0], w_pack, (long long)s->KH * s->KW * lda, Npack, who, N);
        if (rc) return rc;
        for (int c = 1; c < 4; ++c) pp.map_b[c] = pp.map_b[0];
    } else {
        for (int cls = 0; cls < s->stride * s->stride; ++cls) {
            int na, nc;
            class_taps(s, cls, &na, &nc);
            if (na * nc == 0) { pp.map_b[cls] = pp.map_b[0]; continue; }
            rc = encode_weight_map(&pp.map_b[cls], static_cast<const __nv_bfloat16*>(w_pack) + p_in.w_class_off[cls],
                                   (long long)na * nc * lda, Npack, who, N);
            if (rc) return rc;
        }
        if (s->stride == 1) for (int c = 1; c < 4; ++c) pp.map_b[c] = pp.map_b[0];
    }
    long long items;
    int max_nkb, splits, tickets;
    long long ws_bytes;
    px_items(s, form, lda, N, &items, &max_nkb);
    px_split_plan(s, form, lda, N, &splits, &ws_bytes, &tickets);
    pp.p.splits = 1;
    pp.p.kb_per_split = 0;
    pp.p.ws = nullptr;
    pp.p.tickets = nullptr;
    if (splits > 1 && t->splitk_ws && t->splitk_tickets && t->splitk_ws_bytes >= ws_bytes && t->splitk_n_tickets >= tickets) {
        pp.p.splits = splits;
        pp.p.ws = static_cast<float*>(t->splitk_ws);
        pp.p.tickets = t->splitk_tickets;
    }
    dim3 grid((unsigned)(pp.DX * pp.DY * pp.btiles), (unsigned)((N + BN - 1) / BN), (unsigned)pp.p.splits);
    rc = fill_bn(&pp.p, t, grid.x * grid.y, who);
    if (rc) return rc;
    set_stats_fix(&pp.p, t);
    launch_pdl(conv_px_kernel, grid, kPxThreads, kPxSmem, stream, pp);
    return check_launch(who);
}

}  // namespace tc
}  // namespace acg
