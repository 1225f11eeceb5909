// Persistent halo-tile tcgen05 kernel for the stride-2 layers with 16- or 32-pixel-wide tile grids, in BOTH gather forms:
//
//   ADJ  (conv2d_transpose forward, conv2d data gradient; models.py:17-21,39-40,53-59 and TF autodiff of :12-15,34-37,
//         82-86): every tap of an output-parity class is a pure 2-D shift of the small-grid input,
//             out_class[b][y][x] = sum_{ta,tc,k} in[b][y + ea - ta][x + ec - tc][k] * W[class][n][(ta,tc)][k].
//   CONV (conv2d forward, conv2d_transpose data gradient): the input is split into its 4 parity PLANES
//         x_p[b][i][j] = x[b][2i+pi][2j+pj]; tap (a,c) of a stride-2 filter reads plane ((a-pad_t)&1, (c-pad_l)&1) at the
//         shift (floor((a-pad_t)/2), floor((c-pad_l)/2)), so every tap is again a pure 2-D shift of a dense tile:
//             y[b][oh][ow] = sum_{planes} sum_{taps of plane,k} x_p[b][oh+di][ow+dj][k] * W[n][(a,c)][k].
//         A plane is addressed by a plain 4-D tensor map with doubled pixel strides and a shifted base pointer.
//
// One CTA per SM walks a tile list.  A tile is `NACC` accumulators of 128 pixels (8 columns x 16 rows each) = 16 rows x
// TW columns of TB images.  Per 64-channel block (and, CONV form, per plane) the (16+2) x (TW+2) halo patch is ONE 4-D TMA
// copy (hardware zero fill = SAME padding, hardware 128-byte swizzle); each tap's A operand is a shifted descriptor into
// it, every streamed weight slice feeds NACC accumulators.  Roles (384 threads):
//   warp 5 halo TMA producer | warp 6 weight-slice TMA producer | warp 4 tcgen05.mma issuer (+ TMEM owner)
//   warps 0-3 and 8-11: two epilogue groups (accumulators q even / odd), tcgen05.ld -> bias / moments -> global stores
// The accumulators are double buffered in tensor memory (2 x NACC x N <= 512 columns: NACC = 4 for N <= 64, NACC = 2 for
// N <= 128), so copies, MMAs and epilogue of consecutive tiles overlap.
//
// Round-2 measurements that shaped this version (scripts/probe_r2.py, stage knock-out on B200):
//   * the round-1 kernel issued its MMAs from `warp == 4 && lane == 0`: the compiler wrapped every UTCHMMA in a
//     uniformisation loop and the issue thread needed ~95-105 clk per MMA for ANY N (g/tconv4 forward: MMA phase alone
//     133 us of 143 us; math at N=48 is 24 clk).  Here a whole warp walks the loop and elect.sync picks the issuing lane.
//   * with N = 128 there was one accumulator set: MMA (43 us) and epilogue (40 us) of g/tconv3 forward ran back to back.
//   * 4 epilogue warps with one synchronous TMEM load per 16 columns took 83 us for g/tconv4's 151 MB of logits: now 8
//     warps, TMEM loads issued in batches of up to 64 columns before one wait.
#include "conv_tc.cuh"

namespace acg {
namespace tc {

constexpr int kH2Threads = 384;
constexpr int kH2Rows = 18;                         // staged halo rows per image: 16 output rows + 2
constexpr int kH2Data = 214 * 1024;                 // operand staging: 2 halo buffers + the weight-slice ring (the per-warp
                                                    // moment slots and the CTA totals take 10 KB of static shared memory)
constexpr int kH2Smem = kH2Data + 1024;             // + alignment slack (227 KB per CTA is the hardware limit)
constexpr int kH2MaxBStages = 8;
constexpr int kMaxTaps = 9;
constexpr int kEpiBatch = 3;                        // 16-column TMEM loads in flight per wait (48 columns = g/tconv4's N)

struct TapProg {
    int ntaps;
    int oy, ox;                   // halo origin relative to the tile origin (rows, columns of the staged grid)
    int a_off[kMaxTaps];          // tap -> offset of its shifted window inside the halo patch, in 16-byte units
    int w_off[kMaxTaps];          // tap -> K coordinate of its slice in the weight pack ([N][tap][lda]): tap index * lda
};

struct alignas(64) Halo2Params {
    Params p;
    CUtensorMap map_a[4];         // ADJ: [0];  CONV: one per input parity plane
    CUtensorMap map_b[4];         // ADJ: one per output parity class;  CONV: [0]
    TapProg prog[4];              // ADJ: per class;  CONV: per plane
    int form;                     // 0 ADJ, 1 CONV
    int Hs, Ws;                   // tile grid: ADJ class grid (= conv output grid), CONV output grid
    int TW, TB;                   // tile width (8 | 16 | 32) and images per tile
    int rows;                     // staged halo rows per image: 16 + 2, or 8 + 2 in the interleaved 8 x 8 mode
    int il;                       // 1: 8 x 8 grids -- the TB images of a tile are interleaved BY ROW in the halo patch
                                  // (patch row index = y * TB + b), so that the 16 row groups of an accumulator (8 pixels
                                  // each) keep ONE stride although they come from different images
    int nkc, nk16_last;           // 64-channel blocks; K=16 steps of the last block (lda % 64 != 0 -> fewer MMAs)
    int out_H, out_W;             // spatial size of the output tensor
    int npg;                      // CONV form: staged planes per tile (4 parity planes; 2 row planes in the pixel-pair mode)
    int nbs_cap;                  // weight-ring depth cap (kH2MaxBStages; the probe library reads ACG_H2_NBS)
    int staged_rb;                // staged epilogue: bytes per staging row (128 / 64), 0 = off
};

struct H2Tile {
    int b0, y0, x0, pg0, npg, cls;
};
__device__ __forceinline__ H2Tile h2_tile(const Halo2Params& hp, int t) {
    H2Tile h;
    int sp = t;
    if (hp.form == 0) {
        // the 4 parity classes of one spatial tile are adjacent in the list (they read the same input halo: neighbouring
        // CTAs find it in L2) and the class is rotated by the wave index so that every CTA gets 9-, 6- and 4-tap tiles
        sp = t >> 2;
        h.cls = ((t & 3) + t / (int)gridDim.x) & 3;
        h.pg0 = h.cls;
        h.npg = 1;
    } else {
        h.cls = 0;
        h.pg0 = 0;
        h.npg = hp.npg;
    }
    const int tiles_x = hp.il ? 1 : hp.Ws / hp.TW, tiles_y = hp.il ? 1 : hp.Hs >> 4;
    const int per_img = tiles_x * tiles_y;
    const int bi = sp / per_img, r = sp - bi * per_img;
    h.b0 = bi * hp.TB;
    h.y0 = (r / tiles_x) << 4;
    h.x0 = (r % tiles_x) * hp.TW;
    return h;
}

// MMAs of one tap: K step outer, accumulator inner, so that consecutive MMAs write different accumulators
template <int NACC, int NK>
__device__ __forceinline__ void issue_tap(uint32_t tacc, int N, uint32_t alo_t, const uint32_t (&acc_row8)[NACC],
                                          uint32_t ahi, uint32_t blo, uint32_t bhi, uint32_t idesc, uint32_t first) {
#pragma unroll
    for (int k = 0; k < NK; ++k)
#pragma unroll
        for (int q = 0; q < NACC; ++q)
            tc_mma2(tacc + q * N, alo_t + acc_row8[q] + 2 * k, ahi, blo + 2 * k, bhi, idesc, first | (uint32_t)k);
}

// Epilogue of one accumulator for the batch-norm layers (bf16 output, moments, no bias / activation) through the warp's
// staging tile: see epi_stage_put_chunk / epi_stage_moments_flush (conv_tc.cuh).
template <int RB>
__device__ __forceinline__ void epilogue_acc_staged(const Params& p, uint32_t tacc_q, int N, uint32_t stile, uint32_t rowtab,
                                                    int lane, size_t row_off, size_t pix, float* sm_sum, float* sm_sq) {
    constexpr int GC = RB / 2, NCH = GC / 16;           // columns / 16-column chunks per group
    unsigned char* out = static_cast<unsigned char*>(p.out);
    if (p.rz) epi_stage_rowtab(rowtab, lane, (unsigned long long)pix * (unsigned long long)p.rz_ld * 2ull);
    for (int g0 = 0; g0 < N; g0 += GC) {
#pragma unroll
        for (int i0 = 0; i0 < NCH; i0 += 2) {           // two 16-column TMEM loads in flight per wait
            uint32_t v[2][16];
            tmem_ld16_nowait(tacc_q + g0 + 16 * i0, v[0]);
            tmem_ld16_nowait(tacc_q + g0 + 16 * i0 + 16, v[1]);
            tmem_ld_wait();
            epi_stage_put_chunk<RB>(stile, lane, i0, v[0], true);
            epi_stage_put_chunk<RB>(stile, lane, i0 + 1, v[1], true);
        }
        if (ACG_DBG(p, 16)) continue;                                     // probe bit 16: TMEM loads and staging only
        if (p.rz)
            epi_stage_redux_flush<RB>(p, stile, rowtab, lane, g0, out + (size_t)g0 * 2, (unsigned long long)row_off * 2ull, true,
                                      sm_sum + g0, sm_sq + g0);
        else
            epi_stage_moments_flush<RB>(stile, lane, out + (size_t)g0 * 2, (unsigned long long)row_off * 2ull, true,
                                        p.stats != nullptr, sm_sum + g0, sm_sq + g0);
    }
}

template <int NACC>
__global__ void __launch_bounds__(kH2Threads, 1)
conv_halo2_kernel(const __grid_constant__ Halo2Params hp, int ntiles) {
    const Params& p = hp.p;
    extern __shared__ unsigned char smem_raw[];
    __shared__ __align__(8) uint64_t halo_full[2], halo_empty[2], b_full[kH2MaxBStages], b_empty[kH2MaxBStages];
    __shared__ __align__(8) uint64_t acc_full[2], acc_empty[2];
    __shared__ uint32_t tmem_base_sh;
    __shared__ float sm_stats[8][2][BN];       // per epilogue warp: no atomics, fixed summation order
    __shared__ double cta_acc[2][BN];          // this CTA's column totals over all of its tiles
    __shared__ int last_cta_sh;

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t smemH = smem_base;
    for (int i = tid; i < 8 * 2 * BN; i += kH2Threads) (&sm_stats[0][0][0])[i] = 0.f;
    if (tid < BN) { cta_acc[0][tid] = 0.0; cta_acc[1][tid] = 0.0; }

    const int N = p.N;
    const int XG = hp.TW >> 3;                               // 8-column accumulator groups per tile row
    const int WH = hp.TW + 2, HR = hp.rows * WH;             // halo: (16|8)+2 rows x (TW+2) pixels per image
    const uint32_t halo_bytes = (uint32_t)(hp.TB * HR) * 128u;
    const uint32_t halo_stride = (halo_bytes + 1023u) & ~1023u;        // buffers start on swizzle-atom boundaries
    constexpr int NH = 2;                                    // halo ring: a patch feeds >= 4 taps, two in flight are enough
    // The weight-slice ring takes the rest.  A stage is recycled by tcgen05.commit -> mbarrier -> producer -> TMA -> MMA
    // warp: ~3000 clk round trip (probe: 620 clk per tap with 3 stages and NO MMAs), against 384-512 clk of MMA work
    // per tap -- so the ring must hold ~8 taps or the tensor pipe idles (3 stages: 1032-1296 clk per tap measured).
    const uint32_t smemB = smemH + NH * halo_stride;
    const uint32_t b_stride = ((uint32_t)N * 128u + 1023u) & ~1023u;
    // staged epilogue (batch-norm layers): 8 staging tiles of 32 rows x 128 B (N = 128) or x 64 B at the END of the
    // operand area.  The ring gives them up for free: a sweep of its depth (scripts/halo_ring_sweep.py) showed no layer
    // slower at 4 stages than at 8 (g/tconv4's data gradient excepted: +1 us at 5).
    const uint32_t stage_rb = hp.staged_rb;                    // 0: every thread stores its own row
    const uint32_t stage_bytes = stage_rb ? 8u * 32u * stage_rb + 8u * 256u : 0u;     // tiles + per-warp row tables
    const uint32_t smemE = smem_base + (uint32_t)kH2Data - stage_bytes;
    int NBS = (int)(((uint32_t)kH2Data - stage_bytes - NH * halo_stride) / b_stride);
    NBS = NBS > hp.nbs_cap ? hp.nbs_cap : NBS;
    const int NB = (2 * NACC * N <= 512) ? 2 : 1;            // accumulator buffers in tensor memory
    uint32_t tmem_cols = 32;
    while (tmem_cols < (uint32_t)(NB * NACC * N)) tmem_cols <<= 1;

    if (tid == 0) {
        for (int i = 0; i < 2; ++i) {
            mbar_init(&halo_full[i], 1); mbar_init(&halo_empty[i], 1);
            mbar_init(&acc_full[i], 1); mbar_init(&acc_empty[i], 8);
        }
        for (int i = 0; i < kH2MaxBStages; ++i) { mbar_init(&b_full[i], 1); mbar_init(&b_empty[i], 1); }
        fence_mbar_init();
        for (int c = 0; c < 4; ++c) { tma_prefetch_desc(&hp.map_a[c]); tma_prefetch_desc(&hp.map_b[c]); }
    }
    if (warp == 4) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_sh)),
                     "r"(tmem_cols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    // PDL: dependents may be scheduled now that this CTA owns its tensor memory; nothing above touched global memory
    if (p.px.world <= 1) pdl_launch_dependents();     // a launch that exchanges with peers must not (peer.cu)
    pdl_wait();
    const uint32_t tmem_base = tmem_base_sh;

    if (warp == 5) {
        // ================================ halo producer (whole warp walks, one elected lane issues) ================
        int hb = 0;
        uint32_t eph = 1u;                     // first use of every buffer passes at once
        for (int t = blockIdx.x; t < ntiles; t += gridDim.x) {
            const H2Tile h = h2_tile(hp, t);
            for (int kc = 0; kc < hp.nkc; ++kc) {
                for (int pg = h.pg0; pg < h.pg0 + h.npg; ++pg) {
                    mbar_wait(&halo_empty[hb], eph);
                    if (elect_one()) {
                        if (ACG_DBG(p, 1)) {
                            mbar_arrive(&halo_full[hb]);                           // probe: no halo traffic
                        } else {
                            mbar_expect_tx(&halo_full[hb], halo_bytes);
                            if (hp.il)      // tensor map dimensions (C, X, B, Y): images interleaved by row
                                tma_load_4d(smemH + hb * halo_stride, &hp.map_a[hp.form ? pg : 0], kc * 64,
                                            hp.prog[pg].ox, h.b0, hp.prog[pg].oy, &halo_full[hb]);
                            else
                                tma_load_4d(smemH + hb * halo_stride, &hp.map_a[hp.form ? pg : 0], kc * 64,
                                            h.x0 + hp.prog[pg].ox, h.y0 + hp.prog[pg].oy, h.b0, &halo_full[hb]);
                        }
                    }
                    __syncwarp();
                    if (++hb == NH) { hb = 0; eph ^= 1u; }
                }
            }
        }
    } else if (warp == 6) {
        // ================================ weight producer ================================
        const uint32_t b_bytes = (uint32_t)N * 128u;
        int st = 0;
        uint32_t eph = 1u;                     // parity a wait on b_empty[st] must see: passes at once in the first round
        for (int t = blockIdx.x; t < ntiles; t += gridDim.x) {
            const H2Tile h = h2_tile(hp, t);
            for (int kc = 0; kc < hp.nkc; ++kc) {
                for (int pg = h.pg0; pg < h.pg0 + h.npg; ++pg) {
                    const TapProg& pr = hp.prog[pg];
                    for (int tap = 0; tap < pr.ntaps; ++tap) {
                        mbar_wait(&b_empty[st], eph);
                        if (elect_one()) {
                            if (ACG_DBG(p, 2)) {
                                mbar_arrive(&b_full[st]);                          // probe: no weight traffic
                            } else {
                                mbar_expect_tx(&b_full[st], b_bytes);
                                tma_load_2d(smemB + st * b_stride, &hp.map_b[hp.form ? 0 : pg],
                                            pr.w_off[tap] + kc * 64, 0, &b_full[st]);
                            }
                        }
                        __syncwarp();
                        if (++st == NBS) { st = 0; eph ^= 1u; }
                    }
                }
            }
        }
    } else if (warp == 4) {
        // ================================ MMA issuer ================================
        // The tensor pipe queues only a few MMAs: whatever the issuing warp does between two taps is time the pipe
        // idles (round-2 probe: ~165 SASS instructions = ~700 clk per tap against 384-512 clk of MMA work per tap).
        // So everything here is warp-UNIFORM by construction -- loop counters, descriptor words and the TMEM base
        // (__reduce_or_sync lands in a uniform register; a plain shared-memory load does not) -- which lets the
        // compiler keep the operands of UTCHMMA in uniform registers instead of moving them there for every MMA, the
        // ring positions are running counters (no division), and the tap offsets come precomputed from the host.
        const uint32_t tmem_u = __reduce_or_sync(0xffffffffu, tmem_base);
        const uint32_t idesc = make_idesc(N, 0, 0);
        const uint32_t ahi = desc_hi((uint32_t)WH * 128u), bhi = desc_hi(1024);
        const uint32_t blo0 = desc_lo(smemB, 16), bstep = b_stride >> 4;
        const uint32_t alo0 = desc_lo(smemH, 16), hstep = halo_stride >> 4;
        uint32_t acc_row8[NACC];               // first halo row of accumulator q, in 16-byte units (8 per row)
#pragma unroll
        for (int q = 0; q < NACC; ++q) {
            const int tb = q / XG, xg = q - tb * XG;
            acc_row8[q] = hp.il ? (uint32_t)(16 * q * WH) * 8u : (uint32_t)(tb * HR + xg * 8) * 8u;
        }
        const uint32_t bar_bfull = smem_u32(&b_full[0]), bar_bempty = smem_u32(&b_empty[0]);
        int st = 0, hb = 0, tcount = 0;
        uint32_t bph = 0u, hph = 0u;           // parities of the weight ring / halo ring
        for (int t = blockIdx.x; t < ntiles; t += gridDim.x, ++tcount) {
            const H2Tile h = h2_tile(hp, t);
            const int abuf = tcount % NB, ause = tcount / NB;
            if (ause >= 1) {      // the epilogue has drained this accumulator buffer
                mbar_wait(&acc_empty[abuf], (uint32_t)((ause - 1) & 1));
                tc_fence_after();
            }
            const uint32_t tacc = tmem_u + (uint32_t)(abuf * NACC * N);
            uint32_t first = 0u;               // 0 until the first MMA of the tile has been issued
            for (int kc = 0; kc < hp.nkc; ++kc) {
                const int nk = kc == hp.nkc - 1 ? hp.nk16_last : BK / 16;
                for (int pg = h.pg0; pg < h.pg0 + h.npg; ++pg) {
                    const TapProg& pr = hp.prog[pg];
                    mbar_wait(&halo_full[hb], hph);
                    const uint32_t alo_h = alo0 + (uint32_t)hb * hstep;
                    const int ntaps = pr.ntaps;
                    for (int tap = 0; tap < ntaps; ++tap) {
                        mbar_wait_addr(bar_bfull + 8u * (uint32_t)st, bph);
                        tc_fence_after();
                        const uint32_t blo = blo0 + (uint32_t)st * bstep;
                        const uint32_t alo_t = alo_h + (uint32_t)pr.a_off[tap];
                        if (elect_one()) {
                            if (!ACG_DBG(p, 4)) {                                  // probe: no MMAs
                                // K step outer, accumulator inner: consecutive MMAs write different accumulators
                                switch (nk) {      // straight-line code per K-step count (uniform branch)
                                    case 4: issue_tap<NACC, 4>(tacc, N, alo_t, acc_row8, ahi, blo, bhi, idesc, first); break;
                                    case 3: issue_tap<NACC, 3>(tacc, N, alo_t, acc_row8, ahi, blo, bhi, idesc, first); break;
                                    case 2: issue_tap<NACC, 2>(tacc, N, alo_t, acc_row8, ahi, blo, bhi, idesc, first); break;
                                    default: issue_tap<NACC, 1>(tacc, N, alo_t, acc_row8, ahi, blo, bhi, idesc, first); break;
                                }
                            }
                            tc_commit_addr(bar_bempty + 8u * (uint32_t)st);
                        }
                        __syncwarp();
                        first = 1u;
                        if (++st == NBS) { st = 0; bph ^= 1u; }
                    }
                    if (elect_one()) tc_commit(&halo_empty[hb]);
                    __syncwarp();
                    if (++hb == NH) { hb = 0; hph ^= 1u; }
                }
            }
            if (elect_one()) tc_commit(&acc_full[abuf]);
            __syncwarp();
        }
    } else if (warp < 4 || warp >= 8) {
        // ================================ epilogue: two groups of 4 warps ================================
        // warp w may read TMEM lanes 32*(w&3) .. +31; group 0 = warps 0-3 takes accumulators 0, 2, group 1 = warps 8-11
        // takes 1, 3
        const int grp = warp >> 3, wq = warp & 3, ew = grp * 4 + wq;     // ew: epilogue warp index 0..7
        const int ml = wq * 32 + lane, yy = ml >> 3, xi = ml & 7;
        int tcount = 0;
        for (int t = blockIdx.x; t < ntiles; t += gridDim.x, ++tcount) {
            const H2Tile h = h2_tile(hp, t);
            const int abuf = tcount % NB, ause = tcount / NB;
            mbar_wait(&acc_full[abuf], (uint32_t)(ause & 1));
            tc_fence_after();
            const uint32_t tacc = tmem_base + ((uint32_t)(wq * 32) << 16) + (uint32_t)(abuf * NACC * N);
            for (int q = grp; q < NACC && !ACG_DBG(p, 32); q += 2) {                  // probe bit 32: no epilogue work
                int tb = q / XG, oy, ox;
                if (hp.il) {          // patch row G = 16 q + yy of the interleaved tile: image G % TB, row G / TB
                    const int G = 16 * q + yy;
                    tb = G % hp.TB;
                    oy = G / hp.TB;
                    ox = xi;
                } else {
                    const int xg = q - tb * XG;
                    oy = h.y0 + yy;
                    ox = h.x0 + xg * 8 + xi;
                }
                if (hp.form == 0) { oy = (oy << 1) + (h.cls >> 1); ox = (ox << 1) + (h.cls & 1); }
                const size_t pix = (size_t)((h.b0 + tb) * hp.out_H + oy) * hp.out_W + ox;
                const size_t row_off = pix * p.ldo;
                if (stage_rb) {
                    const uint32_t stile = smemE + (uint32_t)ew * 32u * stage_rb;
                    const uint32_t rowtab = smemE + 8u * 32u * stage_rb + (uint32_t)ew * 256u;
                    if (stage_rb == 128u)
                        epilogue_acc_staged<128>(p, tacc + q * N, N, stile, rowtab, lane, row_off, pix, &sm_stats[ew][0][0],
                                                 &sm_stats[ew][1][0]);
                    else
                        epilogue_acc_staged<64>(p, tacc + q * N, N, stile, rowtab, lane, row_off, pix, &sm_stats[ew][0][0],
                                                &sm_stats[ew][1][0]);
                    continue;
                }
                // fused batch-norm backward reduction of the layer that consumes this gradient: its pre-activation row
                const __nv_bfloat16* zrow = p.rz ? p.rz + pix * p.rz_ld : nullptr;
                for (int cb0 = 0; cb0 < N; cb0 += 16 * kEpiBatch) {
                    uint32_t v[kEpiBatch][16];
#pragma unroll
                    for (int i = 0; i < kEpiBatch; ++i)
                        if (cb0 + 16 * i < N) tmem_ld16_nowait(tacc + q * N + cb0 + 16 * i, v[i]);
                    tmem_ld_wait();
                    if (ACG_DBG(p, 16)) continue;                                     // probe bit 16: loads only
#pragma unroll
                    for (int i = 0; i < kEpiBatch; ++i)
                        if (cb0 + 16 * i < N)
                            epilogue_chunk(p, v[i], cb0 + 16 * i, true, row_off, 0u, lane, &sm_stats[ew][0][cb0 + 16 * i],
                                           &sm_stats[ew][1][cb0 + 16 * i],
                                           (zrow && cb0 + 16 * i < p.n_stat) ? zrow + cb0 + 16 * i : nullptr);
                }
            }
            // this accumulator buffer may be overwritten by the MMAs of the tile after next
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&acc_empty[abuf]);
            if (p.stats) {   // per-tile flush of the eight warps' fp32 column sums into the CTA's fp64 totals, fixed order
                asm volatile("bar.sync 1, 256;" ::: "memory");
                if (tid < N && tid < p.n_stat) {
                    float a = 0.f, b = 0.f;
#pragma unroll
                    for (int w = 0; w < 8; ++w) {
                        a += sm_stats[w][0][tid]; b += sm_stats[w][1][tid];
                        sm_stats[w][0][tid] = 0.f; sm_stats[w][1][tid] = 0.f;
                    }
                    cta_acc[0][tid] += (double)a;
                    cta_acc[1][tid] += (double)b;
                }
                asm volatile("bar.sync 1, 256;" ::: "memory");
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 4) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(tmem_cols) : "memory");
    }
    if (p.stats) {
        const bool has_col = tid < N && tid < p.n_stat;
        cta_stats_finish(p, tid, kH2Threads, has_col, tid, has_col ? cta_acc[0][tid] : 0.0,
                         has_col ? cta_acc[1][tid] : 0.0, &last_cta_sh);
    }
}

// ---- host side ---------------------------------------------------------------------------------------------------
namespace {

int encode_map(CUtensorMap* map, const void* base, int rank, const cuuint64_t* dims, const cuuint64_t* strides,
               const cuuint32_t* box, const char* who) {
    EncodeTiledFn enc = encode_tiled_fn();
    ACG_REQUIRE(enc, ACG_ERR_CUDA, "%s: cuTensorMapEncodeTiled is not available", who);
    cuuint32_t estr[4] = {1, 1, 1, 1};
    CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, (cuuint32_t)rank, const_cast<void*>(base), dims, strides, box,
                     estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    ACG_REQUIRE(r == CUDA_SUCCESS, ACG_ERR_CUDA, "%s: tensor map failed (%d)", who, (int)r);
    return ACG_OK;
}

// tile shape for a grid of Hs x Ws pixels and N output columns; false when the kernel does not cover the shape
bool pick_tile(int B, int Hs, int Ws, int N, int* nacc, int* TW, int* TB, int* il) {
    *il = 0;
    if (N > BN || N % 16 != 0) return false;
    if (Hs == 8 && Ws == 8) {        // interleaved mode: one accumulator = two whole 8 x 8 images (B x 64 pixels are too
        *il = 1;                      // few to give every SM more than one accumulator)
        *nacc = 1;
        *TW = 8;
        *TB = 2;
        return B % 2 == 0;
    }
    if (Hs % 16 != 0 || (Ws != 16 && Ws != 32)) return false;
    if (N > 64) { *nacc = 2; *TW = 16; *TB = 1; return true; }              // 2 x 2 x N <= 512 TMEM columns
    *nacc = 4;
    *TW = Ws;
    *TB = Ws == 32 ? 1 : 2;
    return B % *TB == 0;
}

}  // namespace

// shape gates (also used by the dispatchers in conv_tc.cu)
bool halo2_adj_ok(const acg_conv_shape* s, const acg_tc_args* t, int N) {
    if (getenv("ACG_NO_HALO")) return false;
    if (s->stride != 2 || s->KH > 6 || s->KW > 6 || s->KH < 2 || s->KW < 2) return false;
    if (s->OH != s->H / 2 || s->OW != s->W / 2 || (s->H & 1) || (s->W & 1)) return false;
    if (t->ld_in % 64 != 0) return false;
    int nacc, TW, TB, il;
    return pick_tile(s->B, s->OH, s->OW, N, &nacc, &TW, &TB, &il);
}
bool halo2_conv_ok(const acg_conv_shape* s, const acg_tc_args* t, int N) {
    if (getenv("ACG_NO_HALO") || getenv("ACG_NO_HALO_CONV")) return false;
    if (s->stride != 2 || s->KH > 6 || s->KW > 6 || s->KH < 2 || s->KW < 2) return false;
    if (s->OH != s->H / 2 || s->OW != s->W / 2 || (s->H & 1) || (s->W & 1)) return false;
    if (t->ld_in % 16 != 0 || t->ld_in < 16) return false;     // a ragged last 64-channel block issues fewer K steps
    // every plane's taps must fit the 3 x 3 window of the staged halo
    for (int ax = 0; ax < 2; ++ax) {
        const int K = ax ? s->KW : s->KH, pad = ax ? s->pad_l : s->pad_t;
        for (int par = 0; par < 2; ++par) {
            int dmin = 1 << 20, dmax = -(1 << 20);
            for (int a = 0; a < K; ++a) {
                const int r = a - pad, pi = ((r % 2) + 2) % 2, d = (r - pi) / 2;
                if (pi != par) continue;
                dmin = d < dmin ? d : dmin;
                dmax = d > dmax ? d : dmax;
            }
            if (dmax >= dmin && dmax - dmin > 2) return false;
        }
    }
    int nacc, TW, TB, il;
    return pick_tile(s->B, s->OH, s->OW, N, &nacc, &TW, &TB, &il);
}

// Pixel-pair mode of the CONV form, for the FIRST layers (8-channel frame operands: K = 8 per tap is half an MMA step and
// the small-K kernel's 16-byte gather is LSU bound).  x [B,H,W,8] is read as [B,H,W/2,16]: two horizontally adjacent
// pixels are one 16-channel "pixel".  Output column ox reads input columns 2 ox + c - pad_l = pair ox + q, half h with
// q = floor((c - pad_l) / 2), h = (c - pad_l) & 1: along x the stride-2 filter becomes a STRIDE-1 filter of nq <= 3 pair
// taps (zero weights where a half has no tap), along y nothing changes (two row-parity planes).  Every tap is again a
// pure shift of a staged dense patch and one K=16 MMA step: 15 steps per accumulator for a 5x5 filter instead of 13, no
// gather at all.  The weights come as the pair pack (acg_pack_weights which = 2): [N][KH * nq][16].
void pair_taps(const acg_conv_shape* s, int* qmin, int* nq) {
    int lo = 1 << 20, hi = -(1 << 20);
    for (int c = 0; c < s->KW; ++c) {
        const int r = c - s->pad_l, h = ((r % 2) + 2) % 2, q = (r - h) / 2;
        lo = q < lo ? q : lo;
        hi = q > hi ? q : hi;
    }
    *qmin = lo;
    *nq = hi - lo + 1;
}
bool halo2_pair_ok(const acg_conv_shape* s, const acg_tc_args* t, int N) {
    if (getenv("ACG_NO_HALO") || getenv("ACG_NO_PAIR")) return false;
    if (s->stride != 2 || s->KH > 6 || s->KW > 6 || s->KH < 2 || s->KW < 2 || t->ld_in != 8 || t->red_z) return false;
    if (s->OH != s->H / 2 || s->OW != s->W / 2 || (s->H & 1) || (s->W & 1)) return false;
    int qmin, nq;
    pair_taps(s, &qmin, &nq);
    if (nq > 3) return false;
    for (int par = 0; par < 2; ++par) {          // row taps of a parity plane must fit the 3-row window of the halo
        int dmin = 1 << 20, dmax = -(1 << 20);
        for (int a = 0; a < s->KH; ++a) {
            const int r = a - s->pad_t, pi = ((r % 2) + 2) % 2, d = (r - pi) / 2;
            if (pi != par) continue;
            dmin = d < dmin ? d : dmin;
            dmax = d > dmax ? d : dmax;
        }
        if (dmax >= dmin && dmax - dmin > 2) return false;
    }
    int nacc, TW, TB, il;
    return pick_tile(s->B, s->OH, s->OW, N, &nacc, &TW, &TB, &il) && !il;
}

// form 0: dx / deconv output [B,H,W,N] from dy [B,OH,OW,lda] (ADJ);  form 1: y [B,OH,OW,N] from x [B,H,W,lda] (CONV);
// pair: form 1 in the pixel-pair mode (lda == 8, weights = the pair pack)
int launch_halo2(int form, const acg_conv_shape* s, const acg_tc_args* t, const Params& p_in, const void* src,
                 const void* w_pack, int N, cudaStream_t stream, const char* who, bool pair) {
    Halo2Params hp;
    memset(&hp, 0, sizeof(hp));
    hp.p = p_in;
    hp.form = form;
    hp.Hs = s->OH;
    hp.Ws = s->OW;
    int nacc = 4;
    if (!pick_tile(s->B, s->OH, s->OW, N, &nacc, &hp.TW, &hp.TB, &hp.il)) {
        set_error("%s: shape not covered by the halo kernel", who);
        return ACG_ERR_UNSUPPORTED;
    }
    const int lda = pair ? 2 * t->ld_in : t->ld_in;       // pixel-pair mode: 16-channel pair pixels
    hp.npg = pair ? 2 : 4;
    hp.nbs_cap = kH2MaxBStages;
#ifdef ACG_PROBES
    if (const char* e = getenv("ACG_H2_NBS")) {
        const int v = atoi(e);
        if (v >= 1 && v <= kH2MaxBStages) hp.nbs_cap = v;
    }
#endif
    hp.nkc = (lda + 63) / 64;
    hp.nk16_last = ((lda - (hp.nkc - 1) * 64) + 15) / 16;
    hp.out_H = form == 0 ? s->H : s->OH;
    hp.out_W = form == 0 ? s->W : s->OW;
    hp.rows = hp.il ? 10 : kH2Rows;
    const int WHh = hp.TW + 2;                       // halo row pitch in pixels
    const int ystep = hp.il ? hp.TB * WHh : WHh;     // patch rows per unit of vertical shift
    // box: (C, X, Y, B), or (C, X, B, Y) in the interleaved mode
    const cuuint32_t box_a[4] = {64, (cuuint32_t)WHh, (cuuint32_t)(hp.il ? hp.TB : hp.rows),
                                 (cuuint32_t)(hp.il ? hp.rows : hp.TB)};
    int rc;
    if (form == 0) {
        const cuuint64_t row_b = (cuuint64_t)s->OW * lda * 2, img_b = (cuuint64_t)s->OH * s->OW * lda * 2;
        const cuuint64_t dims[4] = {(cuuint64_t)lda, (cuuint64_t)s->OW, (cuuint64_t)(hp.il ? s->B : s->OH),
                                    (cuuint64_t)(hp.il ? s->OH : s->B)};
        const cuuint64_t strides[3] = {(cuuint64_t)lda * 2, hp.il ? img_b : row_b, hp.il ? row_b : img_b};
        rc = encode_map(&hp.map_a[0], src, 4, dims, strides, box_a, who);
        if (rc) return rc;
        for (int c = 1; c < 4; ++c) hp.map_a[c] = hp.map_a[0];
        for (int cls = 0; cls < 4; ++cls) {
            const int ph = cls >> 1, pw = cls & 1;
            const int a0 = (ph + s->pad_t) & 1, c0 = (pw + s->pad_l) & 1;
            const int na = (s->KH - a0 + 1) >> 1, nc = (s->KW - c0 + 1) >> 1;
            const int ea = (ph + s->pad_t - a0) >> 1, ec = (pw + s->pad_l - c0) >> 1;
            TapProg& pr = hp.prog[cls];
            pr.ntaps = na * nc;
            ACG_REQUIRE(pr.ntaps <= kMaxTaps && na <= 3 && nc <= 3, ACG_ERR_UNSUPPORTED, "%s: %d x %d class taps", who, na, nc);
            pr.oy = ea - (na - 1);
            pr.ox = ec - (nc - 1);
            for (int ta = 0; ta < na; ++ta)
                for (int tcx = 0; tcx < nc; ++tcx) {
                    const int i = ta * nc + tcx;
                    pr.a_off[i] = ((na - 1 - ta) * ystep + (nc - 1 - tcx)) * 8;
                    pr.w_off[i] = i * lda;
                }
            const cuuint64_t Kc = (cuuint64_t)pr.ntaps * lda;
            const cuuint64_t wd[2] = {Kc, (cuuint64_t)N};
            const cuuint64_t ws[1] = {Kc * 2};
            const cuuint32_t wb[2] = {64, (cuuint32_t)N};
            const void* base = static_cast<const __nv_bfloat16*>(w_pack) + p_in.w_class_off[cls];
            rc = encode_map(&hp.map_b[cls], base, 2, wd, ws, wb, who);
            if (rc) return rc;
        }
    } else if (pair) {
        const int Hp = s->H / 2, Wp = s->W / 2;           // plane rows; pixel PAIRS per row
        int qmin, nq;
        pair_taps(s, &qmin, &nq);
        for (int pi = 0; pi < 2; ++pi) {
            TapProg& pr = hp.prog[pi];
            int dmin_i = 1 << 20;
            for (int a = 0; a < s->KH; ++a) {
                const int r = a - s->pad_t, q = ((r % 2) + 2) % 2;
                if (q == pi && (r - q) / 2 < dmin_i) dmin_i = (r - q) / 2;
            }
            int n = 0;
            for (int a = 0; a < s->KH; ++a) {
                const int ra = a - s->pad_t, qa = ((ra % 2) + 2) % 2;
                if (qa != pi) continue;
                for (int qi = 0; qi < nq; ++qi) {
                    ACG_REQUIRE(n < kMaxTaps, ACG_ERR_UNSUPPORTED, "%s: more than %d taps per plane", who, kMaxTaps);
                    pr.a_off[n] = (((ra - qa) / 2 - dmin_i) * ystep + qi) * 8;
                    pr.w_off[n] = (a * nq + qi) * lda;
                    ++n;
                }
            }
            pr.ntaps = n;
            pr.oy = n ? dmin_i : 0;
            pr.ox = qmin;
            // row plane pi of the pair view xp [B,H,Wp,16]: xp_p[b][i][j][k] = xp[b][2i+pi][j][k]
            const cuuint64_t dims[4] = {(cuuint64_t)lda, (cuuint64_t)Wp, (cuuint64_t)Hp, (cuuint64_t)s->B};
            const cuuint64_t strides[3] = {(cuuint64_t)lda * 2, (cuuint64_t)2 * Wp * lda * 2, (cuuint64_t)s->H * Wp * lda * 2};
            const void* base = static_cast<const __nv_bfloat16*>(src) + (size_t)pi * Wp * lda;
            rc = encode_map(&hp.map_a[pi], base, 4, dims, strides, box_a, who);
            if (rc) return rc;
        }
        hp.map_a[2] = hp.map_a[0]; hp.map_a[3] = hp.map_a[1];
        const cuuint64_t Kt = (cuuint64_t)s->KH * nq * lda;
        const cuuint64_t wd[2] = {Kt, (cuuint64_t)N};
        const cuuint64_t ws[1] = {Kt * 2};
        const cuuint32_t wb[2] = {64, (cuuint32_t)N};
        rc = encode_map(&hp.map_b[0], w_pack, 2, wd, ws, wb, who);
        if (rc) return rc;
        for (int c = 1; c < 4; ++c) hp.map_b[c] = hp.map_b[0];
    } else {
        const int Hp = s->H / 2, Wp = s->W / 2;
        for (int pl = 0; pl < 4; ++pl) {
            const int pi = pl >> 1, pj = pl & 1;
            TapProg& pr = hp.prog[pl];
            int dmin_i = 1 << 20, dmin_j = 1 << 20;
            for (int a = 0; a < s->KH; ++a) {
                const int r = a - s->pad_t, q = ((r % 2) + 2) % 2;
                if (q == pi && (r - q) / 2 < dmin_i) dmin_i = (r - q) / 2;
            }
            for (int c = 0; c < s->KW; ++c) {
                const int r = c - s->pad_l, q = ((r % 2) + 2) % 2;
                if (q == pj && (r - q) / 2 < dmin_j) dmin_j = (r - q) / 2;
            }
            int n = 0;
            for (int a = 0; a < s->KH; ++a) {
                const int ra = a - s->pad_t, qa = ((ra % 2) + 2) % 2;
                if (qa != pi) continue;
                for (int c = 0; c < s->KW; ++c) {
                    const int rc2 = c - s->pad_l, qc = ((rc2 % 2) + 2) % 2;
                    if (qc != pj) continue;
                    ACG_REQUIRE(n < kMaxTaps, ACG_ERR_UNSUPPORTED, "%s: more than %d taps per plane", who, kMaxTaps);
                    pr.a_off[n] = (((ra - qa) / 2 - dmin_i) * ystep + ((rc2 - qc) / 2 - dmin_j)) * 8;
                    pr.w_off[n] = (a * s->KW + c) * lda;
                    ++n;
                }
            }
            pr.ntaps = n;
            pr.oy = n ? dmin_i : 0;
            pr.ox = n ? dmin_j : 0;
            // plane (pi, pj) of x [B,H,W,lda]: x_p[b][i][j][k] = x[b][2i+pi][2j+pj][k]
            const cuuint64_t row_b = (cuuint64_t)2 * s->W * lda * 2, img_b = (cuuint64_t)s->H * s->W * lda * 2;
            const cuuint64_t dims[4] = {(cuuint64_t)lda, (cuuint64_t)Wp, (cuuint64_t)(hp.il ? s->B : Hp),
                                        (cuuint64_t)(hp.il ? Hp : s->B)};
            const cuuint64_t strides[3] = {(cuuint64_t)2 * lda * 2, hp.il ? img_b : row_b, hp.il ? row_b : img_b};
            const void* base = static_cast<const __nv_bfloat16*>(src) + ((size_t)pi * s->W + pj) * lda;
            rc = encode_map(&hp.map_a[pl], base, 4, dims, strides, box_a, who);
            if (rc) return rc;
        }
        const cuuint64_t Kt = (cuuint64_t)s->KH * s->KW * lda;
        const cuuint64_t wd[2] = {Kt, (cuuint64_t)N};
        const cuuint64_t ws[1] = {Kt * 2};
        const cuuint32_t wb[2] = {64, (cuuint32_t)N};
        rc = encode_map(&hp.map_b[0], w_pack, 2, wd, ws, wb, who);
        if (rc) return rc;
        for (int c = 1; c < 4; ++c) hp.map_b[c] = hp.map_b[0];
    }
    const int n_sp = hp.il ? s->B / hp.TB : (s->B / hp.TB) * (s->OH / 16) * (s->OW / hp.TW);
    const int ntiles = form == 0 ? n_sp * 4 : n_sp;
    const int ctas = ntiles < num_sms() ? ntiles : num_sms();
    rc = fill_bn(&hp.p, t, (unsigned int)ctas, who);
    if (rc) return rc;
    set_stats_fix(&hp.p, t);
    // staged epilogue (column-wise moments, full-line stores) for every bf16 output without bias / activation
    // (ACG_EPI_DIRECT switches it off); the ring keeps at least 4 stages
    hp.staged_rb = 0;
    {
        const Params& q = hp.p;
        const bool ok = !q.bias && q.out_act == ACG_ACT_NONE && q.out_dtype == ACG_BF16 &&
                        (q.ldo & 7) == 0 && q.n_store >= N && (!q.stats || q.n_stat == N) && ((uintptr_t)q.out & 15) == 0 &&
                        !q.direct_store && !pair;
        const long long halo_stride = (((long long)hp.TB * hp.rows * (hp.TW + 2) * 128) + 1023) / 1024 * 1024;
        const long long b_stride = ((long long)N * 128 + 1023) / 1024 * 1024;
        for (int rb = 128; ok && rb >= 64 && !hp.staged_rb; rb >>= 1)
            if (N % (rb / 2) == 0 && (kH2Data - 8LL * 32 * rb - 8 * 256 - 2 * halo_stride) / b_stride >= 4) hp.staged_rb = rb;
    }
    if (nacc == 1) {
        rc = set_smem((const void*)conv_halo2_kernel<1>, kH2Smem);
        if (rc) return rc;
        launch_pdl(conv_halo2_kernel<1>, ctas, kH2Threads, kH2Smem, stream, hp, ntiles);
    } else if (nacc == 2) {
        rc = set_smem((const void*)conv_halo2_kernel<2>, kH2Smem);
        if (rc) return rc;
        launch_pdl(conv_halo2_kernel<2>, ctas, kH2Threads, kH2Smem, stream, hp, ntiles);
    } else {
        rc = set_smem((const void*)conv_halo2_kernel<4>, kH2Smem);
        if (rc) return rc;
        launch_pdl(conv_halo2_kernel<4>, ctas, kH2Threads, kH2Smem, stream, hp, ntiles);
    }
    return check_launch(who);
}

}  // namespace tc
}  // namespace acg
