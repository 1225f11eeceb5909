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
#include <mutex>
#include <utility>
#include <vector>

#include "conv_tc.cuh"

namespace acg {
namespace tc {

// persistent halo-tile kernel in both gather forms (conv_halo.cu)
bool halo2_adj_ok(const acg_conv_shape* s, const acg_tc_args* t, int N);
bool halo2_conv_ok(const acg_conv_shape* s, const acg_tc_args* t, int N);
bool halo2_pair_ok(const acg_conv_shape* s, const acg_tc_args* t, int N);
void pair_taps(const acg_conv_shape* s, int* qmin, int* nq);
int launch_halo2(int form, const acg_conv_shape* s, const acg_tc_args* t, const Params& p_in, const void* src,
                 const void* w_pack, int N, cudaStream_t stream, const char* who, bool pair = false);

constexpr int HALO_ACC = 4;
constexpr int HALO_H = 18;    // staged halo rows per image: 16 output rows + (na-1) <= 2

struct alignas(64) HaloParams {
    Params p;
    CUtensorMap map_a;        // bf16 [B][OH][OW][lda], box {64 ch, OW+2, 18, TB}, 128B swizzle, zero OOB fill
    CUtensorMap map_b[4];     // per parity class: bf16 [N][Kc], box {64, N}, 128B swizzle
};

constexpr int kHaloBuf = 648 * 128;                       // 18 x 34 (one image) or 2 x 18 x 18 rows of 128 B
constexpr int kHaloBStage = BN * BK * 2;
constexpr int kHaloBStages = 3;
constexpr int kHaloSmem = 2 * kHaloBuf + kHaloBStages * kHaloBStage + 1024;


struct alignas(64) ConvParams {
    Params p;
    CUtensorMap map_b[4];     // weights: CONV [N][Ktot] (entry 0) / ADJ per parity class [N][Kc]; box {64, 128}, 128B swizzle
};

enum { CONV = 0, ADJ = 1 };

// timing experiments (ACG_DBG_SKIP bit 3): per-phase nanoseconds summed over CTAs
__device__ unsigned long long g_phase_ns[8];      // written only under ACG_PROBES
__device__ __forceinline__ unsigned long long gtime() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}

// NST = pipeline depth.  3 stages (96 KB) keep two CTAs per SM so that one CTA's epilogue overlaps the other's main loop;
// launches with fewer CTAs than SMs (the 4x4 / 8x8 layers: few tiles, K up to 6400) are latency bound per K slice and
// take the 6-stage variant (192 KB, one CTA per SM) instead.
// The 6-stage variant runs ONE CTA per SM: with 4 producer warps that is one warp per scheduler, and the gather
// (~70 dependent instructions per K slice per thread) issues at one instruction per ~8 cycles -- the phase probe showed
// 0.57 us per K slice with the L2 -> SM path at 60 % and the tensor pipe at 20 %.  It therefore uses EIGHT producer
// warps (4 rows per thread instead of 8); the MMA-issuing warp is the one after the producers.
template <int MODE, int NST>
__global__ void __launch_bounds__(NST == 3 ? kThreads : kThreads6, NST == 3 ? 2 : 1)
conv_tc_kernel(const __grid_constant__ ConvParams cp) {
    const Params& p = cp.p;
    constexpr int STAGES = NST;
    constexpr int PW = NST == 3 ? 4 : 8;            // producer warps
    constexpr int NPROD = PW * 32;                  // producer threads
    constexpr int RPT = BM / (NPROD / 8);           // A rows per producer thread (8 threads cover one 128-byte row)
    constexpr int NTHREADS = NPROD + 32;
    extern __shared__ unsigned char smem_raw[];
    __shared__ __align__(8) uint64_t full_bar[STAGES], empty_bar[STAGES], acc_bar;
    __shared__ uint32_t tmem_base_sh;
    __shared__ float sm_stats[4][2][BN];       // per epilogue warp: no atomics, fixed summation order
    __shared__ int last_cta_sh;

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t smemA = smem_base, smemB = smem_base + STAGES * kStageA;
    for (int i = tid; i < 4 * 2 * BN; i += NTHREADS) (&sm_stats[0][0][0])[i] = 0.f;
    const bool timing = ACG_DBG(p, 8) && tid == 0;
    unsigned long long t0 = 0, t1 = 0, t2 = 0, t3 = 0;
    if (timing) t0 = gtime();

    // ---- geometry of this CTA ---------------------------------------------------------------------------
    const int s = p.stride;
    const int split = p.splits > 1 ? (int)(blockIdx.z % (unsigned)p.splits) : 0;
    const int zcls = p.splits > 1 ? (int)(blockIdx.z / (unsigned)p.splits) : (int)blockIdx.z;
    int M, ntaps, nc = 1, a0 = 0, c0 = 0, ph = 0, pw = 0, Hp = 0, Wp = 0;
    const __nv_bfloat16* wmat = p.w_pack;
    if (MODE == CONV) {
        M = p.B * p.OH * p.OW;
        ntaps = p.KH * p.KW;
    } else {
        ph = zcls / s; pw = zcls % s;
        Hp = (p.H - ph + s - 1) / s; Wp = (p.W - pw + s - 1) / s;
        a0 = (ph + p.pad_t) % s; c0 = (pw + p.pad_l) % s;
        const int na = a0 < p.KH ? (p.KH - a0 + s - 1) / s : 0;
        nc = c0 < p.KW ? (p.KW - c0 + s - 1) / s : 0;
        ntaps = na * nc;
        M = p.B * Hp * Wp;
        wmat += p.w_class_off[zcls];
        if (nc == 0) nc = 1;
    }
    const int tile_m = blockIdx.x * BM;
    if (tile_m >= M) return;   // whole CTA exits together (smaller parity classes)
    const int n0 = blockIdx.y * BN;
    const int n_cta = min(BN, p.N - n0);
    const int Ktot = ntaps * p.lda;
    const int nkb_all = (Ktot + BK - 1) / BK;
    const int kb0 = split * p.kb_per_split;                              // first K block of this CTA (0 without split-K)
    const int nkb = p.splits > 1 ? max(0, min(nkb_all - kb0, p.kb_per_split)) : nkb_all;

    if (tid == 0) {
        for (int i = 0; i < STAGES; ++i) { mbar_init(&full_bar[i], NPROD + 1); mbar_init(&empty_bar[i], 1); }
        mbar_init(&acc_bar, 1);
        fence_mbar_init();
        tma_prefetch_desc(&cp.map_b[MODE == ADJ ? zcls : 0]);
    }
    if (warp == PW) {  // tensor-memory allocation is warp-collective; this warp also owns the dealloc
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_sh)),
                     "r"((uint32_t)BN) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    // PDL: dependents may be scheduled now that this CTA owns its tensor memory; nothing above touched global memory
    if (p.px.world <= 1) pdl_launch_dependents();     // a launch that exchanges with peers must not (peer.cu)
    pdl_wait();
    const uint32_t tmem_base = tmem_base_sh;
    if (timing) t1 = gtime();

    // epilogue row of this thread (warps 0-3): output offset and, for the fused backward reduction, the offset of
    // the consumer's pre-activation row, prefetched into L2 now so that the epilogue finds it there
    size_t ep_row_off = 0, ep_rz_off = 0;
    const int ep_m = tile_m + warp * 32 + lane;
    const int nz = p.rz ? max(0, min(n_cta, p.n_stat - n0)) : 0;      // staged z columns of this CTA
    if (warp < 4 && ep_m < M) {
        if (MODE == CONV) { ep_row_off = (size_t)ep_m * p.ldo; ep_rz_off = (size_t)ep_m * p.rz_ld; }
        else {
            const int b = ep_m / (Hp * Wp), r = ep_m - b * Hp * Wp;
            const int ih = (r / Wp) * s + ph, iw = (r % Wp) * s + pw;
            const size_t pix = (size_t)(b * p.H + ih) * p.W + iw;
            ep_row_off = pix * p.ldo;
            ep_rz_off = pix * p.rz_ld;
        }
        if (nz > 0) {
            prefetch_l2(p.rz + ep_rz_off + n0);
            if (nz > 64) prefetch_l2(p.rz + ep_rz_off + n0 + 64);
        }
    }

    if (warp < PW) {
        // ================================ producers ================================
        // A (activations): cp.async 16 B gathers.  Everything that does not change along K is hoisted: per row an
        // element offset of its tap-(0,0) source pixel and a bit mask of the taps that fall inside the image, so a
        // K slice costs ~6 instructions per row (mask test, 64-bit add, LDGSTS) instead of the full index math.
        // B (weights): ONE 2-D TMA copy per K slice issued by thread 0 (hardware swizzle, zero fill past N / Ktot).
        const int j = tid & 7, rslot = tid >> 3;
        long long row_off[RPT];
        uint32_t row_mask[RPT];
#pragma unroll
        for (int i = 0; i < RPT; ++i) {
            const int m = tile_m + rslot + (NPROD / 8) * i;
            row_off[i] = 0;
            row_mask[i] = 0;
            if (m < M) {
                if (MODE == CONV) {
                    const int b = m / (p.OH * p.OW), r = m - b * p.OH * p.OW;
                    const int oh = r / p.OW, ow = r - oh * p.OW;
                    const int y0 = oh * s - p.pad_t, x0 = ow * s - p.pad_l;
                    row_off[i] = ((long long)(b * p.H + y0) * p.W + x0) * p.lda;
                    // taps a in [a_lo, a_hi) x c in [c_lo, c_hi) fall inside the image
                    const int a_lo = max(0, -y0), a_hi = min(p.KH, p.H - y0);
                    const int c_lo = max(0, -x0), c_hi = min(p.KW, p.W - x0);
                    if (c_hi > c_lo) {
                        const uint32_t cm = ((1u << c_hi) - 1u) & ~((1u << c_lo) - 1u);
                        for (int a = a_lo; a < a_hi; ++a) row_mask[i] |= cm << (a * p.KW);
                    }
                } else {
                    const int b = m / (Hp * Wp), r = m - b * Hp * Wp;
                    const int ih = (r / Wp) * s + ph, iw = (r % Wp) * s + pw;
                    const int ohb = (ih + p.pad_t - a0) / s, owb = (iw + p.pad_l - c0) / s;   // tap (0,0) of the class
                    row_off[i] = ((long long)(b * p.OH + ohb) * p.OW + owb) * p.lda;
                    // class taps ta in [a_lo, a_hi) x tc in [c_lo, c_hi): 0 <= ohb - ta < OH, 0 <= owb - tc < OW
                    const int a_lo = max(0, ohb - p.OH + 1), a_hi = min(ntaps / nc, ohb + 1);
                    const int c_lo = max(0, owb - p.OW + 1), c_hi = min(nc, owb + 1);
                    if (c_hi > c_lo) {
                        const uint32_t cm = ((1u << c_hi) - 1u) & ~((1u << c_lo) - 1u);
                        for (int ta = a_lo; ta < a_hi; ++ta) row_mask[i] |= cm << (ta * nc);
                    }
                }
            }
        }
        const CUtensorMap* bmap = &cp.map_b[MODE == ADJ ? zcls : 0];
        int tap = (kb0 * BK + j * 8) / p.lda, ci = (kb0 * BK + j * 8) % p.lda;
        for (int kb = 0; kb < nkb; ++kb) {
            const int stage = kb % STAGES;
            if (kb >= STAGES) mbar_wait(&empty_bar[stage], (uint32_t)(((kb / STAGES) - 1) & 1));
            if (tid == 0) {
                mbar_expect_tx(&full_bar[stage], (uint32_t)min(p.N, BN) * 128u);
                tma_load_2d(smemB + stage * kStageB, bmap, (kb0 + kb) * BK, n0, &full_bar[stage]);
            }
            const uint32_t tbit = tap < ntaps ? (1u << tap) : 0u;
            long long koff;
            if (MODE == CONV) { const int a = tap / p.KW, c = tap - a * p.KW; koff = (long long)(a * p.W + c) * p.lda + ci; }
            else { const int ta = tap / nc, tcc = tap - ta * nc; koff = ci - (long long)(ta * p.OW + tcc) * p.lda; }
            const uint32_t dstA = smemA + stage * kStageA + (rslot * 128) + (j << 4);
#pragma unroll
            for (int i = 0; i < RPT; ++i) {
                // row r = rslot + (NPROD/8) i: r & 7 == rslot & 7, so the swizzle term is the same for all rows
                const bool ok = (row_mask[i] & tbit) != 0;
                cp_async16((dstA ^ ((uint32_t)(rslot & 7) << 4)) + i * (NPROD / 8) * 128,
                           ok ? (const void*)(p.a_src + row_off[i] + koff) : (const void*)p.a_src, ok ? 16u : 0u);
            }
            cp_async_arrive_noinc(&full_bar[stage]);
            ci += BK;
            while (ci >= p.lda) { ci -= p.lda; ++tap; }
        }
    } else if (warp == PW) {
        // ================================ MMA issuer ================================
        // the whole warp walks the loop, elect.sync picks the issuing lane (see elect_one() in conv_tc.cuh: a plain
        // `lane == 0` branch costs ~95 clk of uniformisation code per MMA)
        const uint32_t idesc = make_idesc(n_cta, 0, 0);
        const uint32_t hi = desc_hi(1024);
        const uint32_t alo0 = desc_lo(smemA, 16), blo0 = desc_lo(smemB, 16);
        const uint32_t tmem_u = __reduce_or_sync(0xffffffffu, tmem_base);    // warp-uniform register (see conv_halo.cu)
        for (int kb = 0; kb < nkb; ++kb) {
            const int stage = kb % STAGES;
            mbar_wait(&full_bar[stage], (uint32_t)((kb / STAGES) & 1));
            tc_fence_after();
            const uint32_t alo = alo0 + stage * (kStageA >> 4), blo = blo0 + stage * (kStageB >> 4);
            if (elect_one()) {
#pragma unroll
                for (int k = 0; k < BK / 16; ++k)   // +32 B per K=16 step inside the swizzle atom
                    tc_mma2(tmem_u, alo + 2 * k, hi, blo + 2 * k, hi, idesc, (kb | k) != 0 ? 1u : 0u);
                tc_commit(&empty_bar[stage]);   // arrives when the MMAs above have finished reading this stage
            }
            __syncwarp();
        }
        if (elect_one()) tc_commit(&acc_bar);
        __syncwarp();
    }

    // ================================ epilogue (warps 0-3) ================================
    if (warp < 4) {
        if (nkb > 0) {
            mbar_wait(&acc_bar, 0);
            tc_fence_after();
        }
        if (timing) t2 = gtime();
    }
    // ---- split-K: park the partial tile, take a ticket; only the last CTA of the tile goes on to the epilogue ----
    bool final_cta = true;
    float* ws_tile = nullptr;
    if (p.splits > 1) {
        const unsigned int tile_id = ((unsigned)zcls * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x;
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
            const unsigned int t = atomicAdd(&p.tickets[tile_id], 1u);
            last_cta_sh = (t == (unsigned)p.splits - 1u);
            if (last_cta_sh) p.tickets[tile_id] = 0u;       // ready for the next launch
        }
        __syncthreads();
        final_cta = last_cta_sh != 0;
        if (final_cta) __threadfence();
    }
    if (warp < 4 && final_cta) {
        const int m = ep_m;
        const size_t row_off = ep_row_off;
        const uint32_t zrow = smem_base + (uint32_t)(warp * 32 + lane) * kZRowBytes;
        if (nz > 0) {       // every MMA has completed: the pipeline stages are free
            stage_z_row(zrow, p.rz + ep_rz_off + n0, nz, m < M);
            cp_async_commit();
            cp_async_wait<0>();
        }
        for (int cb = 0; cb < n_cta; cb += 16) {
            uint32_t v[16];
            if (nkb > 0 && p.splits == 1) tmem_ld16(tmem_base + ((uint32_t)(warp * 32) << 16) + cb, v);
            else {
#pragma unroll
                for (int i = 0; i < 16; ++i) v[i] = 0u;
            }
            if (p.splits > 1) {     // add ALL partial tiles in split order (L2-resident; bypass L1) -- this CTA's own one is
                                    // read back from the workspace too, so the sum does not depend on which CTA came last
                for (int sp = 0; sp < p.splits; ++sp) {
                    const float4* o = reinterpret_cast<const float4*>(
                        ws_tile + (size_t)sp * (BM * BN) + (size_t)(cb >> 4) * (BM * 16) + (warp * 32 + lane) * 16);
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        const float4 t = __ldcg(o + i);
                        v[4 * i] = __float_as_uint(__uint_as_float(v[4 * i]) + t.x);
                        v[4 * i + 1] = __float_as_uint(__uint_as_float(v[4 * i + 1]) + t.y);
                        v[4 * i + 2] = __float_as_uint(__uint_as_float(v[4 * i + 2]) + t.z);
                        v[4 * i + 3] = __float_as_uint(__uint_as_float(v[4 * i + 3]) + t.w);
                    }
                }
            }
            epilogue_chunk(p, v, n0 + cb, m < M, row_off, zrow + 2 * cb, lane, &sm_stats[warp][0][cb], &sm_stats[warp][1][cb]);
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == PW) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)BN) : "memory");
    }
    if (timing) {
        t3 = gtime();
        atomicAdd(&g_phase_ns[0], t1 - t0);
        atomicAdd(&g_phase_ns[1], t2 - t1);
        atomicAdd(&g_phase_ns[2], t3 - t2);
        atomicAdd(&g_phase_ns[3], 1ull);
        atomicAdd(&g_phase_ns[4], (unsigned long long)nkb);
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
        cta_stats_finish(p, tid, NTHREADS, has_col, n0 + tid, s0, s1, &last_cta_sh);
    }
}


// ---- CONV gather, persistent small-K variant (the first layers: K = 25 taps x 8 channels = 200) ---------------------
// With 4 K slices per tile the generic kernel never reaches a steady state, and -- ncu source view -- its ~10 k warp
// instructions per tile (row decode with divisions, per-chunk moment butterflies) run on 2 warps per scheduler at one
// instruction per ~8 cycles: instruction ISSUE LATENCY, not memory or the tensor pipe, bounds it (66 / 74 us for
// 1.3 / 5 GFLOP).  This kernel is specialised for what those layers are (32-wide output rows, N = 32 or 64):
//   * one CTA per SM walks its tiles; the whole weight matrix is resident in shared memory (<= 4 TMA slices);
//   * gather producers (warps 5-8), MMA issuer (warp 4) and epilogue (warps 0-3) are different warps; an 8-stage A
//     ring keeps two tiles in flight and two accumulators in tensor memory decouple MMAs from the epilogue;
//   * everything tile-invariant is computed once per kernel: per-slice tap offsets and tap bits, per-row column
//     masks and offsets (a tile is 4 whole output rows); a tile costs the producer ~300 instructions instead of ~720;
//   * the batch-norm moments are kept PER THREAD in registers across all tiles of the CTA (2 x N accumulators) and
//     reduced over the 32 rows of a warp once at the end, instead of two 16-shuffle butterflies per 16-column chunk
//     (epilogue ~75 instead of ~250 instructions per chunk).
constexpr int kSmallKMaxKb = 4;
// warps 0-3 epilogue, warp 4 MMA issue + TMEM, warps 5.. gather producers.  N = 32: EIGHT producer warps (with four --
// one per scheduler -- the ~300 dependent instructions a tile costs a producer thread issue at one per ~6 clk: 25.5 ->
// 23.3 us for g/conv1).  N = 64 keeps four: its epilogue threads hold 2 x 64 moment accumulators, and the 128-register
// cap of a 416-thread CTA spilled them (32 -> 45 us).
// Knock-out probe (scripts/smallk_knockout.py): with gather, MMAs and epilogue all switched off the launch still takes
// ~16 us of its 23 / 32 us -- and neither one arrival per warp instead of per thread, nor plain arrivals instead of
// tcgen05.commit, nor a deeper ring, nor dropping the 3 padding MMAs of the last K block changed that.
template <int NT> struct SmallK {
    static constexpr int kProd = NT == 32 ? 256 : 128;
    static constexpr int kRows = kProd / 8;               // tile rows covered by one pass of the producer threads
    static constexpr int kPass = BM / kRows;              // passes (rows per producer thread) per K slice
    static constexpr int kThreads = 160 + kProd;
    static constexpr int kStageBN = kStageB;              // one resident weight slice (N rows x 128 B used)
    static constexpr int kRing = 8;
    static constexpr int kSmem = kSmallKMaxKb * kStageBN + kRing * kStageA + 4 * kEpiStageBytes + 1024;
};

template <int NT>
__global__ void __launch_bounds__(SmallK<NT>::kThreads, 1)
conv_smallk_persistent_kernel(const __grid_constant__ ConvParams cp, int ntiles) {
    constexpr int kSmallKRing = SmallK<NT>::kRing, kSmallKProd = SmallK<NT>::kProd, kSmallKRows = SmallK<NT>::kRows,
                  kSmallKPass = SmallK<NT>::kPass, kSmallKThreads = SmallK<NT>::kThreads, kStageBN = SmallK<NT>::kStageBN;
    const Params& p = cp.p;
    extern __shared__ unsigned char smem_raw[];
    __shared__ __align__(8) uint64_t full_bar[kSmallKRing], empty_bar[kSmallKRing], b_full, acc_full[2], acc_empty[2];
    __shared__ uint32_t tmem_base_sh;
    __shared__ float sm_stats[4][2][BN];       // per epilogue warp
    __shared__ int last_cta_sh;

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t smemB = smem_base, smemA = smem_base + kSmallKMaxKb * kStageBN;
    const uint32_t smemE = smemA + kSmallKRing * kStageA;      // per epilogue warp: staging tile of the transposed stores
    for (int i = tid; i < 4 * 2 * BN; i += kSmallKThreads) (&sm_stats[0][0][0])[i] = 0.f;

    const int s = p.stride;
    const int M = p.B * p.OH * p.OW;
    const int ntaps = p.KH * p.KW;
    const int nkb = (ntaps * p.lda + BK - 1) / BK;       // <= kSmallKMaxKb
    constexpr uint32_t tmem_cols = 2 * NT < 32 ? 32 : 2 * NT;

    if (tid == 0) {
        // probe bit 64 (only together with bit 1): one plain arrival per producer warp instead of one per thread
        for (int i = 0; i < kSmallKRing; ++i) {
            mbar_init(&full_bar[i], ACG_DBG(p, 64) ? kSmallKProd / 32 : kSmallKProd);
            mbar_init(&empty_bar[i], 1);
        }
        mbar_init(&b_full, 1);
        for (int i = 0; i < 2; ++i) { mbar_init(&acc_full[i], 1); mbar_init(&acc_empty[i], 4); }
        fence_mbar_init();
        tma_prefetch_desc(&cp.map_b[0]);
    }
    if (warp == 4) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_sh)),
                     "r"(tmem_cols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    if (p.px.world <= 1) pdl_launch_dependents();     // a launch that exchanges with peers must not (peer.cu)
    pdl_wait();
    const uint32_t tmem_base = tmem_base_sh;

    if (warp >= 5) {
        // ================================ gather producers ================================
        const int ptid = tid - 160;
        if (ptid == 0) {     // resident weights: one TMA slice per K block, all on one barrier
            mbar_expect_tx(&b_full, (uint32_t)(nkb * NT) * 128u);
            for (int kb = 0; kb < nkb; ++kb) tma_load_2d(smemB + kb * kStageBN, &cp.map_b[0], kb * BK, 0, &b_full);
        }
        const int j = ptid & 7, rslot = ptid >> 3;
        // per K slice (kernel constants): tap bit and source offset of this thread's 16-byte column
        uint32_t tbit[kSmallKMaxKb];
        int koff[kSmallKMaxKb];
        {
            int tap = (j * 8) / p.lda, ci = (j * 8) % p.lda;
#pragma unroll
            for (int kb = 0; kb < kSmallKMaxKb; ++kb) {
                tbit[kb] = tap < ntaps ? (1u << tap) : 0u;
                const int a = tap / p.KW, c = tap - a * p.KW;
                koff[kb] = (a * p.W + c) * p.lda + ci;
                ci += BK;
                while (ci >= p.lda) { ci -= p.lda; ++tap; }
            }
        }
        // per row slot (kernel constants; OW == 32: a tile is 4 whole output rows): row delta, column mask, offset
        int doh[kSmallKPass], xoff[kSmallKPass];
        uint32_t cmask[kSmallKPass];
        uint32_t fullsel = 0;
        for (int a = 0; a < p.KH; ++a) fullsel |= 1u << (a * p.KW);
#pragma unroll
        for (int i = 0; i < kSmallKPass; ++i) {
            const int q = rslot + kSmallKRows * i;
            doh[i] = q >> 5;
            const int x0 = (q & 31) * s - p.pad_l;
            const int c_lo = max(0, -x0), c_hi = min(p.KW, p.W - x0);
            cmask[i] = c_hi > c_lo ? (((1u << c_hi) - 1u) & ~((1u << c_lo) - 1u)) : 0u;
            xoff[i] = x0 * p.lda;
        }
        const int tiles_per_img = (p.OH * p.OW) >> 7;
        int cnt = 0;
        for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
            const int b = tile / tiles_per_img, oh0 = (tile - b * tiles_per_img) << 2;
            long long row_off[kSmallKPass];
            uint32_t row_mask[kSmallKPass];
#pragma unroll
            for (int i = 0; i < kSmallKPass; ++i) {
                const int y0 = (oh0 + doh[i]) * s - p.pad_t;
                const int a_lo = max(0, -y0), a_hi = min(p.KH, p.H - y0);
                const uint32_t rowsel = fullsel & ((1u << (a_hi * p.KW)) - 1u) & ~((1u << (a_lo * p.KW)) - 1u);
                row_mask[i] = b < p.B ? cmask[i] * rowsel : 0u;       // cmask < 2^KW: the product has no carries
                row_off[i] = (long long)(b * p.H + y0) * p.W * p.lda + xoff[i];
            }
#pragma unroll
            for (int kb = 0; kb < kSmallKMaxKb; ++kb) {
                if (kb < nkb) {
                    const int stage = cnt % kSmallKRing, use = cnt / kSmallKRing;
                    if (use >= 1) mbar_wait(&empty_bar[stage], (uint32_t)((use - 1) & 1));
                    const uint32_t dstA = (smemA + stage * kStageA + (rslot * 128) + (j << 4)) ^ ((uint32_t)(rslot & 7) << 4);
#pragma unroll
                    for (int i = 0; i < kSmallKPass; ++i) {
                        if (ACG_DBG(p, 1)) break;                                  // probe: no gather traffic
                        const bool ok = (row_mask[i] & tbit[kb]) != 0;
                        cp_async16(dstA + i * (kSmallKRows * 128), ok ? (const void*)(p.a_src + row_off[i] + koff[kb]) : (const void*)p.a_src,
                                   ok ? 16u : 0u);
                    }
                    if (ACG_DBG(p, 64)) {
                        __syncwarp();
                        if (lane == 0) mbar_arrive(&full_bar[stage]);
                    } else {
                        cp_async_arrive_noinc(&full_bar[stage]);
                    }
                    ++cnt;
                }
            }
        }
    } else if (warp < 4) {
        // ================================ epilogue ================================
        float q1[NT], q2[NT];                 // per-thread batch-norm moments of this thread's row, over all tiles
#pragma unroll
        for (int i = 0; i < NT; ++i) { q1[i] = 0.f; q2[i] = 0.f; }
        const bool bf16_out = p.out_dtype == ACG_BF16;
        // bf16 rows of NT columns = 64 (NT = 32) or 128 (NT = 64) bytes go out through the warp's staging tile
        constexpr int RB = NT * 2;
        const bool staged = bf16_out && (p.ldo & 7) == 0 && !p.direct_store;
        const uint32_t stile = smemE + (uint32_t)warp * kEpiStageBytes;
        int it = 0;
        for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++it) {
            const int abuf = it & 1;
            mbar_wait(&acc_full[abuf], (uint32_t)((it >> 1) & 1));
            tc_fence_after();
            const int m = tile * BM + warp * 32 + lane;
            const bool row_ok = m < M;
            const size_t row_off = (size_t)m * p.ldo;
            const uint32_t tacc = tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)(abuf * NT);
#pragma unroll
            for (int cb = 0; cb < NT; cb += 16) {
                if (ACG_DBG(p, 32)) break;                                         // probe: no epilogue work
                uint32_t v[16];
                tmem_ld16(tacc + cb, v);
                if (ACG_DBG(p, 16)) continue;                                      // probe: TMEM loads only
                if (bf16_out) {
                    uint32_t w[8];
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        __nv_bfloat162 h = __floats2bfloat162_rn(__uint_as_float(v[2 * i]), __uint_as_float(v[2 * i + 1]));
                        w[i] = *reinterpret_cast<uint32_t*>(&h);
                    }
                    if (row_ok) {
#pragma unroll
                        for (int i = 0; i < 8; ++i) {     // moments of exactly what is stored
                            const float lo = __uint_as_float(w[i] << 16), hi = __uint_as_float(w[i] & 0xffff0000u);
                            q1[cb + 2 * i] += lo; q2[cb + 2 * i] = fmaf(lo, lo, q2[cb + 2 * i]);
                            q1[cb + 2 * i + 1] += hi; q2[cb + 2 * i + 1] = fmaf(hi, hi, q2[cb + 2 * i + 1]);
                        }
                        if (!staged) {
                            uint4* o = reinterpret_cast<uint4*>(static_cast<__nv_bfloat16*>(p.out) + row_off + cb);
                            o[0] = make_uint4(w[0], w[1], w[2], w[3]);
                            o[1] = make_uint4(w[4], w[5], w[6], w[7]);
                        }
                    }
                    if (staged) {
                        epi_stage_put<RB>(stile, lane, cb >> 3, w[0], w[1], w[2], w[3]);
                        epi_stage_put<RB>(stile, lane, (cb >> 3) + 1, w[4], w[5], w[6], w[7]);
                    }
                } else if (row_ok) {
#pragma unroll
                    for (int i = 0; i < 16; ++i) {
                        const float f = __uint_as_float(v[i]);
                        q1[cb + i] += f; q2[cb + i] = fmaf(f, f, q2[cb + i]);
                    }
                    float4* o = reinterpret_cast<float4*>(static_cast<float*>(p.out) + row_off + cb);
#pragma unroll
                    for (int i = 0; i < 4; ++i)
                        o[i] = make_float4(__uint_as_float(v[4 * i]), __uint_as_float(v[4 * i + 1]),
                                           __uint_as_float(v[4 * i + 2]), __uint_as_float(v[4 * i + 3]));
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&acc_empty[abuf]);
            // (after the arrive: the MMAs of the tile after next may start while the rows leave)
            if (staged && !ACG_DBG(p, 16 | 32))
                epi_stage_flush<RB>(stile, lane, static_cast<unsigned char*>(p.out), (unsigned long long)row_off * 2ull, row_ok);
        }
        if (p.stats) {      // one reduction over the warp's 32 rows per 16-column chunk, for the whole kernel
#pragma unroll
            for (int cb = 0; cb < NT; cb += 16) {
                float a16[16], b16[16];
#pragma unroll
                for (int i = 0; i < 16; ++i) { a16[i] = q1[cb + i]; b16[i] = q2[cb + i]; }
                const float cs = warp_colsum16(a16, lane), cs2 = warp_colsum16(b16, lane);
                if ((lane & 1) == 0) {
                    sm_stats[warp][0][cb + (lane >> 1)] = cs;
                    sm_stats[warp][1][cb + (lane >> 1)] = cs2;
                }
            }
        }
    } else if (warp == 4) {
        // ================================ MMA issuer (whole warp walks, elect.sync issues) ================================
        const uint32_t idesc = make_idesc(NT, 0, 0);
        const uint32_t hi = desc_hi(1024);
        const uint32_t alo0 = desc_lo(smemA, 16), blo0 = desc_lo(smemB, 16);
        mbar_wait(&b_full, 0);
        int cnt = 0, it = 0;
        for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++it) {
            const int abuf = it & 1;
            if (it >= 2) {
                mbar_wait(&acc_empty[abuf], (uint32_t)(((it >> 1) - 1) & 1));
                tc_fence_after();
            }
            const uint32_t tacc = __reduce_or_sync(0xffffffffu, tmem_base) + (uint32_t)(abuf * NT);
            for (int kb = 0; kb < nkb; ++kb, ++cnt) {
                const int stage = cnt % kSmallKRing;
                mbar_wait(&full_bar[stage], (uint32_t)((cnt / kSmallKRing) & 1));
                tc_fence_after();
                const uint32_t alo = alo0 + stage * (kStageA >> 4), blo = blo0 + kb * (kStageBN >> 4);
                if (elect_one()) {
                    if (!ACG_DBG(p, 4)) {                                          // probe: no MMAs
#pragma unroll
                        for (int k = 0; k < BK / 16; ++k)
                            tc_mma2(tacc, alo + 2 * k, hi, blo + 2 * k, hi, idesc, (kb | k) != 0 ? 1u : 0u);
                    }
                    if (ACG_DBG(p, 128)) mbar_arrive(&empty_bar[stage]);   // probe (with bit 4): plain arrive, no commit
                    else tc_commit(&empty_bar[stage]);
                }
                __syncwarp();
            }
            if (elect_one()) tc_commit(&acc_full[abuf]);
            __syncwarp();
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 4) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(tmem_cols) : "memory");
    }
    if (p.stats) {
        const bool has_col = tid < NT && tid < p.n_stat;
        double s0 = 0.0, s1 = 0.0;
        if (has_col) {
#pragma unroll
            for (int w = 0; w < 4; ++w) { s0 += (double)sm_stats[w][0][tid]; s1 += (double)sm_stats[w][1][tid]; }
        }
        cta_stats_finish(p, tid, kSmallKThreads, has_col, tid, s0, s1, &last_cta_sh);
    }
}

// ---- ADJ gather, halo-tile variant ---------------------------------------------------------------------------------
// For a stride-2 transposed convolution every tap of an output-parity class is a pure 2-D shift of the (small-grid)
// input: out_class[b][y][x] = sum_{ta,tc,k} in[b][y + ea - ta][x + ec - tc][k] * W[class][n][(ta,tc)][k].
// The generic kernel above re-gathers the shifted 128-row A slice from L2 for every tap and re-reads the 16 KB
// weight slice for every 128 output rows; it sits on the L2->SM bandwidth (~4.8 TB/s), not on the tensor pipe.
// Here one CTA owns 4 accumulators = 512 output pixels (16 rows x 32 columns of one image, or 16 x 16 of two):
//   * per 64-channel block the (16+na-1) x (W+nc-1) input halo patch of each image is staged ONCE (<= 83 KB, two
//     buffers) in the 128-byte-swizzled K-major layout with swizzle phase = absolute row index;
//   * each tap's A operand is then just a descriptor into that patch: start row = shift(tap) + 8*column-group, 8-row
//     groups spaced by the halo width (SBO = WH*128 B) -- tcgen05.mma applies the swizzle on absolute smem address
//     bits, so a start that is not 1024-B aligned is fine (verified on hardware by acg_debug_umma_shift);
//   * every streamed weight slice (tap x 64 channels, N x 128 B) feeds 16 MMAs (4 accumulators x K=64) instead of 4.
// L2->SM traffic per output tile drops ~5x (A: taps x 16 KB -> 1/4 of a shared 83 KB halo; B: 16 KB -> 4 KB).
__global__ void __launch_bounds__(kThreads, 1)
conv_adj_halo_kernel(const __grid_constant__ HaloParams hp) {
    const Params& p = hp.p;
    extern __shared__ unsigned char smem_raw[];
    __shared__ __align__(8) uint64_t halo_full[2], halo_empty[2], b_full[kHaloBStages], b_empty[kHaloBStages], acc_bar;
    __shared__ uint32_t tmem_base_sh;
    __shared__ float sm_stats[4][2][BN];       // per epilogue warp
    __shared__ int last_cta_sh;

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t smemH = smem_base, smemB = smem_base + 2 * kHaloBuf;
    for (int i = tid; i < 4 * 2 * BN; i += kThreads) (&sm_stats[0][0][0])[i] = 0.f;
    const bool timing = ACG_DBG(p, 8) && tid == 0;
    const bool mtiming = ACG_DBG(p, 8) && tid == 128;
    unsigned long long t0 = 0, t1 = 0, t2 = 0, t3 = 0, w_halo = 0, w_b = 0;
    if (timing) t0 = gtime();

    // ---- geometry (stride 2) -----------------------------------------------------------------------------
    const int ph = blockIdx.z >> 1, pw = blockIdx.z & 1;
    const int Hs = p.H >> 1, Ws = p.W >> 1;                 // class grid == conv-output grid (OH x OW)
    const int a0 = (ph + p.pad_t) & 1, c0 = (pw + p.pad_l) & 1;
    const int na = (p.KH - a0 + 1) >> 1, nc = (p.KW - c0 + 1) >> 1;
    const int ea = (ph + p.pad_t - a0) >> 1, ec = (pw + p.pad_l - c0) >> 1;
    const int ntaps = na * nc;
    const int XG = Ws >> 3, TB = HALO_ACC / XG;             // column groups of 8 per image row, images per CTA
    const int WH = Ws + 2, HR = HALO_H * WH;                // staged halo: 18 rows x (Ws+2) pixels per image
    const int tiles_y = Hs >> 4;
    const int b0 = (blockIdx.x / tiles_y) * TB, y0 = (blockIdx.x % tiles_y) << 4;
    const int N = p.N;                                      // <= 128, multiple of 16
    const int nkc = p.lda >> 6;                             // 64-channel blocks
    uint32_t tmem_cols = 32;
    while (tmem_cols < (uint32_t)(HALO_ACC * N)) tmem_cols <<= 1;

    if (tid == 0) {
        for (int i = 0; i < 2; ++i) { mbar_init(&halo_full[i], 1); mbar_init(&halo_empty[i], 1); }
        for (int i = 0; i < kHaloBStages; ++i) { mbar_init(&b_full[i], 1); mbar_init(&b_empty[i], 1); }
        mbar_init(&acc_bar, 1);
        fence_mbar_init();
        tma_prefetch_desc(&hp.map_a);
        tma_prefetch_desc(&hp.map_b[blockIdx.z]);
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
    if (timing) t1 = gtime();

    if (p.rz && warp < 4) {   // fused backward reduction: pull this thread's 4 pre-activation rows into L2 now
        const int ml = warp * 32 + lane, yy = ml >> 3, xi = ml & 7;
        const int nz = min(N, p.n_stat);
        for (int q = 0; q < HALO_ACC; ++q) {
            const int tb = q / XG, xg = q - tb * XG;
            const int ih = ((y0 + yy) << 1) + ph, iw = ((xg * 8 + xi) << 1) + pw;
            const __nv_bfloat16* zr = p.rz + ((size_t)((b0 + tb) * p.H + ih) * p.W + iw) * p.rz_ld;
            prefetch_l2(zr);
            if (nz > 64) prefetch_l2(zr + 64);
        }
    }

    if (tid == 0) {
        // ================================ producer: one thread, TMA only ================================
        // The whole halo patch of a 64-channel block is ONE 4-D tensor copy (hardware zero fill outside the image,
        // hardware 128B swizzle); every (tap, channel block) weight slice is ONE 2-D tensor copy.
        const uint32_t a_bytes = (uint32_t)(TB * HR) * 128u, b_bytes = (uint32_t)N * 128u;
        const int oh0 = y0 + ea - (na - 1), ow0 = ec - (nc - 1);
        auto load_halo = [&](int kc) {
            mbar_expect_tx(&halo_full[kc & 1], a_bytes);
            tma_load_4d(smemH + (kc & 1) * kHaloBuf, &hp.map_a, kc * 64, ow0, oh0, b0, &halo_full[kc & 1]);
        };
        load_halo(0);
        if (nkc > 1) load_halo(1);
        int bcount = 0;
        for (int kc = 0; kc < nkc; ++kc) {
            for (int tap = 0; tap < ntaps; ++tap, ++bcount) {
                const int st = bcount % kHaloBStages;
                if (bcount >= kHaloBStages) mbar_wait(&b_empty[st], (uint32_t)(((bcount / kHaloBStages) - 1) & 1));
                mbar_expect_tx(&b_full[st], b_bytes);
                tma_load_2d(smemB + st * kHaloBStage, &hp.map_b[blockIdx.z], tap * p.lda + kc * 64, 0, &b_full[st]);
            }
            if (kc + 2 < nkc) {   // the buffer of block kc is reused by block kc+2 once its MMAs have drained
                mbar_wait(&halo_empty[kc & 1], (uint32_t)((kc >> 1) & 1));
                load_halo(kc + 2);
            }
        }
    } else if (warp == 4 && lane == 0) {
        // ================================ MMA issuer ================================
        const uint32_t idesc = make_idesc(N, 0, 0);
        const uint32_t ahi = desc_hi((uint32_t)WH * 128u), bhi = desc_hi(1024);
        const uint32_t blo0 = desc_lo(smemB, 16);
        uint32_t acc_row8[HALO_ACC];           // first halo row of accumulator q, in 16-byte units (x8 per row)
#pragma unroll
        for (int q = 0; q < HALO_ACC; ++q) {
            const int tb = q / XG, xg = q - tb * XG;
            acc_row8[q] = (uint32_t)(tb * HR + xg * 8) * 8u;
        }
        int bcount = 0;
        for (int kc = 0; kc < nkc; ++kc) {
            unsigned long long ta0 = 0;
            if (mtiming) ta0 = gtime();
            mbar_wait(&halo_full[kc & 1], (uint32_t)((kc >> 1) & 1));
            if (mtiming) w_halo += gtime() - ta0;
            const uint32_t hbase = smemH + (kc & 1) * kHaloBuf;
            for (int tap = 0; tap < ntaps; ++tap, ++bcount) {
                const int st = bcount % kHaloBStages;
                if (mtiming) ta0 = gtime();
                mbar_wait(&b_full[st], (uint32_t)((bcount / kHaloBStages) & 1));
                if (mtiming) w_b += gtime() - ta0;
                tc_fence_after();
                const int ta = tap / nc, tcc = tap - ta * nc;
                const int shift = (na - 1 - ta) * WH + (nc - 1 - tcc);
                const uint32_t blo = blo0 + st * (kHaloBStage >> 4);
                const uint32_t alo_t = desc_lo(hbase, 16) + (uint32_t)shift * 8u;   // 128 B per halo row
#pragma unroll
                for (int q = 0; q < HALO_ACC; ++q) {
                    const uint32_t alo = alo_t + acc_row8[q];
#pragma unroll
                    for (int k = 0; k < BK / 16; ++k)
                        tc_mma2(tmem_base + q * N, alo + 2 * k, ahi, blo + 2 * k, bhi, idesc,
                                (kc | tap | k) != 0 ? 1u : 0u);
                }
                tc_commit(&b_empty[st]);
            }
            tc_commit(&halo_empty[kc & 1]);
        }
        tc_commit(&acc_bar);
        if (mtiming) { atomicAdd(&g_phase_ns[6], w_halo + w_b); }
    }

    // ================================ epilogue (warps 0-3): 4 accumulators ================================
    if (warp < 4) {
        mbar_wait(&acc_bar, 0);
        tc_fence_after();
        if (timing) t2 = gtime();
        const int ml = warp * 32 + lane, yy = ml >> 3, xi = ml & 7;
        const int nz = p.rz ? min(N, p.n_stat) : 0;
        const uint32_t zrow0 = smem_base + (uint32_t)ml * kZRowBytes;     // + q * 128 rows; the halo buffers are idle now
        if (nz > 0) {
            for (int q = 0; q < HALO_ACC; ++q) {
                const int tb = q / XG, xg = q - tb * XG;
                const int ih = ((y0 + yy) << 1) + ph, iw = ((xg * 8 + xi) << 1) + pw;
                const size_t pix = (size_t)((b0 + tb) * p.H + ih) * p.W + iw;
                stage_z_row(zrow0 + (uint32_t)q * (BM * kZRowBytes), p.rz + pix * p.rz_ld, nz, true);
            }
            cp_async_commit();
            cp_async_wait<0>();
        }
        for (int q = 0; q < HALO_ACC; ++q) {
            const int tb = q / XG, xg = q - tb * XG;
            const int ih = ((y0 + yy) << 1) + ph, iw = ((xg * 8 + xi) << 1) + pw;
            const size_t pix = (size_t)((b0 + tb) * p.H + ih) * p.W + iw;
            const size_t row_off = pix * p.ldo;
            const uint32_t zrow = zrow0 + (uint32_t)q * (BM * kZRowBytes);
            for (int cb = 0; cb < N; cb += 16) {
                uint32_t v[16];
                tmem_ld16(tmem_base + ((uint32_t)(warp * 32) << 16) + q * N + cb, v);
                epilogue_chunk(p, v, cb, true, row_off, zrow + 2 * cb, lane, &sm_stats[warp][0][cb], &sm_stats[warp][1][cb]);
            }
        }
    }
    unsigned long long tb4 = 0;
    if (ACG_DBG(p, 8) && lane == 0) tb4 = gtime();
    tc_fence_before();
    __syncthreads();
    if (ACG_DBG(p, 8) && lane == 0 && warp == 1) atomicAdd(&g_phase_ns[7], gtime() - tb4);   // warp 1's barrier wait
    if (warp == 4) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(tmem_cols) : "memory");
    }
    if (timing) {
        t3 = gtime();
        atomicAdd(&g_phase_ns[5], tb4 - t2);   // thread 0: epilogue loop only (overwrites the halo-wait slot semantics)
        atomicAdd(&g_phase_ns[0], t1 - t0);
        atomicAdd(&g_phase_ns[1], t2 - t1);
        atomicAdd(&g_phase_ns[2], t3 - t2);
        atomicAdd(&g_phase_ns[3], 1ull);
        atomicAdd(&g_phase_ns[4], (unsigned long long)(nkc * ntaps));
    }
    if (p.stats) {
        const bool has_col = tid < N && tid < p.n_stat;
        double s0 = 0.0, s1 = 0.0;
        if (has_col) {
#pragma unroll
            for (int w = 0; w < 4; ++w) { s0 += (double)sm_stats[w][0][tid]; s1 += (double)sm_stats[w][1][tid]; }
        }
        cta_stats_finish(p, tid, kThreads, has_col, tid, s0, s1, &last_cta_sh);
    }
}


// ---- all packs of a parameter store in ONE launch ---------------------------------------------------------------
// After every optimizer step ~22 weight tensors x 2 packs have to be refreshed; one launch per pack costs more in
// launch latency than in work.  The job table and a TILE table live in device memory (built once by the host,
// acg_pack_plan): a tile is 32 pack rows x 32 pack columns of one tap.  The CONV pack is a per-tap transpose of the
// HWIO matrix ([ci][n] -> [n][ci]), done through shared memory so that both the fp32 reads (128 B per warp row) and
// the bf16 writes (64 B per warp row) are coalesced; the ADJ pack keeps the channel order and only re-groups taps.
__global__ void __launch_bounds__(256)
pack_tiles_kernel(const acg_pack_job* __restrict__ jobs, const int4* __restrict__ tiles, int ntiles) {
    pdl_prologue();
    __shared__ float sm[32][33];
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;       // 32 x 8
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const int4 T = tiles[tile];
        const acg_pack_job j = jobs[T.x];
        const int tap = T.y & 255, t = (T.y >> 8) & 255, nt = (T.y >> 16) & 255;
        const int n0 = T.z & 0xffff, c0 = (int)((unsigned)T.z >> 16);
        const float* __restrict__ w = static_cast<const float*>(j.w);
        __nv_bfloat16* __restrict__ out = static_cast<__nv_bfloat16*>(j.pack) + T.w;
        const int ld = j.ld_k;
        if (j.which == 0) {                  // out[(n*taps + tap)*ld + ci] = w[(tap*Cin + ci)*Cout + n]
            const int taps = j.KH * j.KW;
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const int ci = c0 + ty + 8 * k, n = n0 + tx;
                sm[ty + 8 * k][tx] = (ci < j.Cin && n < j.Cout) ? w[((size_t)tap * j.Cin + ci) * j.Cout + n] : 0.f;
            }
            __syncthreads();
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const int n = n0 + ty + 8 * k, ci = c0 + tx;
                if (n < j.N && ci < ld) out[((size_t)n * taps + tap) * ld + ci] = __float2bfloat16_rn(sm[tx][ty + 8 * k]);
            }
            __syncthreads();
        } else if (j.which == 2) {           // pixel-pair CONV pack: out[(n*nt + t)*16 + c0 + ci] = w[(tap*Cin + ci)*Cout + n]
            if (tx < 8) {                    // c0 = 8 * half; tap 255: this half of the pair has no filter tap (zeros)
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const int n = n0 + ty + 8 * k;
                    if (n < j.N) {
                        const float v = (tap != 255 && tx < j.Cin && n < j.Cout) ? w[((size_t)tap * j.Cin + tx) * j.Cout + n] : 0.f;
                        out[((size_t)n * nt + t) * ld + c0 + tx] = __float2bfloat16_rn(v);
                    }
                }
            }
        } else {                             // out[class offset + (n*nt + t)*ld + co] = w[(tap*Cin + n)*Cout + co]
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const int n = n0 + ty + 8 * k, co = c0 + tx;
                if (n < j.N && co < ld) {
                    const float v = (co < j.Cout && n < j.Cin) ? w[((size_t)tap * j.Cin + n) * j.Cout + co] : 0.f;
                    out[((size_t)n * nt + t) * ld + co] = __float2bfloat16_rn(v);
                }
            }
        }
    }
}

int ru(int v, int m) { return (v + m - 1) / m * m; }

void class_taps(const acg_conv_shape* s, int cls, int* na, int* nc) {
    const int ph = cls / s->stride, pw = cls % s->stride;
    const int a0 = (ph + s->pad_t) % s->stride, c0 = (pw + s->pad_l) % s->stride;
    *na = a0 < s->KH ? (s->KH - a0 + s->stride - 1) / s->stride : 0;
    *nc = c0 < s->KW ? (s->KW - c0 + s->stride - 1) / s->stride : 0;
}

EncodeTiledFn encode_tiled_fn() {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void* ptr = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &q) != cudaSuccess ||
            q != cudaDriverEntryPointSuccess)
            return nullptr;
        fn = reinterpret_cast<EncodeTiledFn>(ptr);
    }
    return fn;
}

// bf16 [rows][K] row-major weight pack -> 2-D map with a {64 x 128} box (128B swizzle, zero fill out of bounds)
int encode_weight_map(CUtensorMap* map, const void* base, long long K, int rows, const char* who, int n_used) {
    EncodeTiledFn enc = encode_tiled_fn();
    ACG_REQUIRE(enc, ACG_ERR_CUDA, "%s: cuTensorMapEncodeTiled is not available", who);
    if (n_used <= 0 || n_used > rows) n_used = rows;                 // rows of the matrix the launch actually reads
    cuuint64_t dims[2] = {(cuuint64_t)K, (cuuint64_t)rows};
    cuuint64_t strides[1] = {(cuuint64_t)K * 2};
    cuuint32_t box[2] = {64, (cuuint32_t)(n_used < BN ? n_used : BN)};   // kernels expect min(N, 128) x 128 B per slice
    cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    ACG_REQUIRE(r == CUDA_SUCCESS, ACG_ERR_CUDA, "%s: weight tensor map failed (%d)", who, (int)r);
    return ACG_OK;
}

int encode_halo_maps(HaloParams* hp, const acg_conv_shape* s, const acg_tc_args* t, int N, int TB, const void* dy,
                     const void* w_pack) {
    EncodeTiledFn enc = encode_tiled_fn();
    ACG_REQUIRE(enc, ACG_ERR_CUDA, "acg_conv_dgrad_tc: cuTensorMapEncodeTiled is not available");
    const cuuint64_t lda = (cuuint64_t)t->ld_in;
    {   // activations: [B][OH][OW][lda] bf16, innermost first
        cuuint64_t dims[4] = {lda, (cuuint64_t)s->OW, (cuuint64_t)s->OH, (cuuint64_t)s->B};
        cuuint64_t strides[3] = {lda * 2, (cuuint64_t)s->OW * lda * 2, (cuuint64_t)s->OH * s->OW * lda * 2};
        cuuint32_t box[4] = {64, (cuuint32_t)(s->OW + 2), (cuuint32_t)HALO_H, (cuuint32_t)TB};
        cuuint32_t estr[4] = {1, 1, 1, 1};
        CUresult r = enc(&hp->map_a, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(dy), dims, strides, box, estr,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        ACG_REQUIRE(r == CUDA_SUCCESS, ACG_ERR_CUDA, "acg_conv_dgrad_tc: activation tensor map failed (%d)", (int)r);
    }
    for (int cls = 0; cls < 4; ++cls) {   // weights of each parity class: [N][Kc] bf16
        int na, nc;
        class_taps(s, cls, &na, &nc);
        const cuuint64_t Kc = (cuuint64_t)na * nc * lda;
        cuuint64_t dims[2] = {Kc, (cuuint64_t)N};
        cuuint64_t strides[1] = {Kc * 2};
        cuuint32_t box[2] = {64, (cuuint32_t)N};
        cuuint32_t estr[2] = {1, 1};
        void* base = const_cast<__nv_bfloat16*>(static_cast<const __nv_bfloat16*>(w_pack) + hp->p.w_class_off[cls]);
        CUresult r = enc(&hp->map_b[cls], CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, base, dims, strides, box, estr,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        ACG_REQUIRE(r == CUDA_SUCCESS, ACG_ERR_CUDA, "acg_conv_dgrad_tc: weight tensor map %d failed (%d)", cls, (int)r);
    }
    return ACG_OK;
}

// the halo variant covers stride-2 layers whose class grid is 16k rows x 16 or 32 columns, N <= 128, 64-channel K blocks
bool halo_ok(const acg_conv_shape* s, const acg_tc_args* t, int N) {
    if (getenv("ACG_NO_HALO")) return false;
    if (s->stride != 2 || s->KH > 6 || s->KW > 6 || s->KH < 2 || s->KW < 2) return false;
    if (s->H % 32 != 0 || (s->W != 32 && s->W != 64)) return false;
    if (s->OH != s->H / 2 || s->OW != s->W / 2) return false;
    if (N > BN || t->ld_in % 64 != 0) return false;
    const int TB = HALO_ACC / (s->W / 16);
    if (s->B % TB != 0) return false;
    return true;
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
    int fast;        // 1: a 64-pixel K slice is whole rows of one image (OW | 64 | OH*OW) or whole images (OH*OW | 64)
    int a_tma;       // 1: the dy operand (a plain [pixels][ldy] matrix) arrives by TMA: two 64 x 64 boxes per K slice
    int b_tma;       // 1: the x operand arrives by TMA too (regular geometry): per 64-channel atom ONE box
                     //    {64 channels, bw, bh, bn pixels} of the tap's input-parity plane, zero fill outside the image
                     //    AND past the last channel (ldx = 48, 144, 272: the last atom of a tap is partly padding)
    int ldn;         // channels per tap in the N index of the GEMM: ldx, or ldx rounded up to 64 with b_tma
};
struct alignas(64) WgradTmaParams {
    WgradParams p;
    CUtensorMap map_a;        // dy as bf16 [Kd][ldy], box {64 channels, 64 pixels}, 128B swizzle, zero OOB fill
    CUtensorMap map_x[4];     // x parity planes (stride 2; entry 0 alone for stride 1) as (C, X, Y, B)
};

// Producers: EIGHT warps (the cp.async gather is issue-latency bound: a variant with half the producer warps per SM ran 2x
// slower), 4 pixels of a K slice per thread; warps 0-3 are also the epilogue, warp 8 issues the MMAs and owns the TMEM.
constexpr int kWgProd = 256, kWgThreads = kWgProd + 32, kWgPx = BK / (kWgProd / 16);

__global__ void __launch_bounds__(kWgThreads, 2)
conv_wgrad_tc_kernel(const __grid_constant__ WgradTmaParams wp) {
    const WgradParams& p = wp.p;
    extern __shared__ unsigned char smem_raw[];
    __shared__ __align__(8) uint64_t full_bar[STAGES], empty_bar[STAGES], acc_bar;
    __shared__ uint32_t tmem_base_sh;

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t smemA = smem_base, smemB = smem_base + STAGES * kStageA;

    const int Ntot = p.KH * p.KW * p.ldn;
    const int n0 = blockIdx.x * BN;
    const int n_valid = min(BN, Ntot - n0);          // multiple of 8 (ldx % 8 == 0)
    const int n_cta = (n_valid + 15) & ~15;          // MMA N: multiple of 16, columns past n_valid are zero filled
    const int co0 = blockIdx.y * BM;
    const int Kd = p.B * p.OH * p.OW;
    const int k_begin = blockIdx.z * p.k_chunk;
    const int k_end = min(Kd, k_begin + p.k_chunk);
    const int nkb = k_end > k_begin ? (k_end - k_begin + BK - 1) / BK : 0;
    if (nkb == 0) return;

    if (tid == 0) {
        for (int i = 0; i < STAGES; ++i) {
            mbar_init(&full_bar[i], p.b_tma ? 1 : kWgProd + (p.a_tma ? 1 : 0));
            mbar_init(&empty_bar[i], 1);
        }
        if (p.a_tma) tma_prefetch_desc(&wp.map_a);
        mbar_init(&acc_bar, 1);
        fence_mbar_init();
    }
    if (warp == kWgProd / 32) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_sh)),
                     "r"((uint32_t)BN) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    // PDL: dependents may be scheduled now that this CTA owns its tensor memory; nothing above touched global memory
    pdl_launch_dependents();
    pdl_wait();
    const uint32_t tmem_base = tmem_base_sh;

    if (p.b_tma && warp < kWgProd / 32) {
        // ---- both operands by TMA: warp 0 is the whole producer (warps 1-7 go straight to the epilogue / exit) ----
        // A K slice is 64 consecutive output pixels = whole rows of one image or whole images; for filter tap (ta, tc) the
        // matching input pixels are a box of the parity plane ((ta - pad_t) & 1, (tc - pad_l) & 1) of x, shifted by
        // floor((ta - pad_t) / 2) rows and floor((tc - pad_l) / 2) columns -- out-of-image pixels are zero filled by the
        // copy engine.  Each 64-column atom of the N tile is one tap (ldx % 64 == 0).
        if (warp == 0) {
            const int S = p.OH * p.OW;
            const int natoms = n_valid > 64 ? 2 : 1;
            int pl[2], dj[2], di[2], cc[2];
#pragma unroll
            for (int a = 0; a < 2; ++a) {
                const int nn = n0 + 64 * a;
                const int tap = nn / p.ldn;
                cc[a] = nn - tap * p.ldn;
                const int ta = tap / p.KW, tcx = tap - ta * p.KW;
                const int ry = ta - p.pad_t, rx = tcx - p.pad_l;
                if (p.stride == 2) {
                    const int qy = ry & 1, qx = rx & 1;
                    pl[a] = qy * 2 + qx;
                    di[a] = (ry - qy) >> 1;       // floor(ry / 2): ry - qy is even
                    dj[a] = (rx - qx) >> 1;
                } else {
                    pl[a] = 0; di[a] = ry; dj[a] = rx;
                }
            }
            const uint32_t tx = (uint32_t)kStageA + (uint32_t)natoms * (BK * 128);
            int bP = k_begin / S;
            int ohP = (k_begin - bP * S) / p.OW;
            const int rows_per_slice = S >= BK ? BK / p.OW : 0, imgs_per_slice = S >= BK ? 0 : BK / S;
            for (int kb = 0; kb < nkb; ++kb) {
                const int stage = kb % STAGES;
                if (kb >= STAGES) mbar_wait(&empty_bar[stage], (uint32_t)(((kb / STAGES) - 1) & 1));
                if (elect_one()) {
                    mbar_expect_tx(&full_bar[stage], tx);
                    tma_load_2d(smemA + stage * kStageA, &wp.map_a, co0, k_begin + kb * BK, &full_bar[stage]);
                    tma_load_2d(smemA + stage * kStageA + BK * 128, &wp.map_a, co0 + 64, k_begin + kb * BK, &full_bar[stage]);
                    tma_load_4d(smemB + stage * kStageB, &wp.map_x[pl[0]], cc[0], dj[0], ohP + di[0], bP, &full_bar[stage]);
                    if (natoms == 2)
                        tma_load_4d(smemB + stage * kStageB + BK * 128, &wp.map_x[pl[1]], cc[1], dj[1], ohP + di[1], bP,
                                    &full_bar[stage]);
                }
                __syncwarp();
                ohP += rows_per_slice;
                if (ohP >= p.OH) { ohP = 0; ++bP; }
                bP += imgs_per_slice;
            }
        }
    } else if (warp < kWgProd / 32) {
        // producers: thread = (16-byte channel chunk c16 of the 128-wide tile, pixel slot 0..15)
        const int c16 = tid & 15, pslot = tid >> 4;
        const int atom = c16 >> 3, jc = c16 & 7;
        // A: dy channels co0 + c16*8 .. +8
        const int a_ch = co0 + c16 * 8;
        const bool a_ch_ok = a_ch + 8 <= p.ldy;
        // B: n = n0 + c16*8 -> (tap, ci) fixed for the whole kernel
        const int nn = n0 + c16 * 8;
        const bool b_ch_ok = c16 * 8 < n_valid;
        const int tap = nn / p.ldx, ci = nn - tap * p.ldx;
        const int ta = tap / p.KW, tcc = tap - ta * p.KW;
        const int S = p.OH * p.OW;
        const int soff0 = pslot * 128 + ((jc ^ (pslot & 7)) << 4);  // k & 7 == pslot & 7 for every k = pslot + 16 i
        if (p.fast) {
            // Regular geometry (a 64-pixel K slice is whole rows of one image, or whole images): everything about the
            // slice-relative position q = pslot + 8 i of this thread's 8 pixels is a constant of the kernel -- source
            // offset relative to the slice's first pixel, the input row delta, whether the input column is inside the
            // image.  A K slice then costs ~12 instructions per 16-byte copy pair instead of ~59 (the running decode
            // with its wrap-around loops and divergent branches made the producers, not the tensor pipe, the limiter:
            // ncu source view, 61 % of the samples in this loop).
            int offB[kWgPx], dih[kWgPx];
            uint32_t iw_ok = 0;
#pragma unroll
            for (int i = 0; i < kWgPx; ++i) {
                const int q = pslot + (kWgProd / 16) * i;
                int db = 0, rem = q;
                if (S < BK) { db = q / S; rem = q - db * S; }
                const int doh = rem / p.OW, dow = rem - doh * p.OW;
                const int iw = dow * p.stride + tcc - p.pad_l;
                if ((unsigned)iw < (unsigned)p.W) iw_ok |= 1u << i;
                dih[i] = doh * p.stride;
                offB[i] = ((db * p.H + doh * p.stride) * p.W + dow * p.stride) * p.ldx;
            }
            // slice base: first pixel P = k_begin (multiple of 64) -> image bP, row ohP (column 0)
            int bP = k_begin / S;
            int ohP = (k_begin - bP * S) / p.OW;                  // 0 when S < 64
            const int rows_per_slice = S >= BK ? BK / p.OW : 0;   // rows a slice advances inside an image
            const int imgs_per_slice = S >= BK ? 0 : BK / S;
            const __nv_bfloat16* dyP = p.dy + (long long)(k_begin + pslot) * p.ldy + a_ch;
            const long long tap_off = ((long long)(ta - p.pad_t) * p.W + (tcc - p.pad_l)) * p.ldx + ci;
            int kleft = k_end - k_begin;                          // valid pixels from the slice base on
            for (int kb = 0; kb < nkb; ++kb) {
                const int stage = kb % STAGES;
                if (kb >= STAGES) mbar_wait(&empty_bar[stage], (uint32_t)(((kb / STAGES) - 1) & 1));
                const uint32_t dstA = smemA + stage * kStageA + atom * (BK * 128) + soff0;
                const uint32_t dstB = smemB + stage * kStageB + atom * (BK * 128) + soff0;
                const __nv_bfloat16* xP = p.x + ((long long)(bP * p.H + ohP * p.stride) * p.W) * p.ldx + tap_off;
                const int ihP = ohP * p.stride + ta - p.pad_t;
                if (p.a_tma) {
                    if (tid == 0) {     // dy slice: two 64-channel atoms x 64 pixels, hardware swizzle / zero fill
                        mbar_expect_tx(&full_bar[stage], (uint32_t)kStageA);
                        tma_load_2d(smemA + stage * kStageA, &wp.map_a, co0, k_begin + kb * BK, &full_bar[stage]);
                        tma_load_2d(smemA + stage * kStageA + BK * 128, &wp.map_a, co0 + 64, k_begin + kb * BK, &full_bar[stage]);
                    }
                }
#pragma unroll
                for (int i = 0; i < kWgPx; ++i) {
                    const bool pv = pslot + (kWgProd / 16) * i < kleft;
                    if (!p.a_tma) {
                        const bool aok = pv && a_ch_ok;
                        cp_async16(dstA + i * (kWgProd / 16) * 128,
                                   aok ? (const void*)(dyP + (long long)((kWgProd / 16) * i) * p.ldy) : (const void*)p.dy,
                                   aok ? 16u : 0u);
                    }
                    const bool bok = pv && b_ch_ok && ((iw_ok >> i) & 1u) && (unsigned)(ihP + dih[i]) < (unsigned)p.H;
                    cp_async16(dstB + i * (kWgProd / 16) * 128, bok ? (const void*)(xP + offB[i]) : (const void*)p.x,
                               bok ? 16u : 0u);
                }
                cp_async_arrive_noinc(&full_bar[stage]);
                dyP += (long long)BK * p.ldy;
                kleft -= BK;
                ohP += rows_per_slice;
                if (ohP >= p.OH) { ohP = 0; ++bP; }
                bP += imgs_per_slice;
            }
        } else {
        // running decode of this thread's pixel (it advances by 8 per step, 64 per K slice): no divisions in the loop
        int pix = k_begin + pslot;
        int pb = pix / (p.OH * p.OW), pr = pix - pb * p.OH * p.OW;
        int poh = pr / p.OW, pow_ = pr - poh * p.OW;
        for (int kb = 0; kb < nkb; ++kb) {
            const int stage = kb % STAGES;
            if (kb >= STAGES) mbar_wait(&empty_bar[stage], (uint32_t)(((kb / STAGES) - 1) & 1));
            const uint32_t dstA = smemA + stage * kStageA + atom * (BK * 128) + soff0;
            const uint32_t dstB = smemB + stage * kStageB + atom * (BK * 128) + soff0;
            if (p.a_tma && tid == 0) {
                mbar_expect_tx(&full_bar[stage], (uint32_t)kStageA);
                tma_load_2d(smemA + stage * kStageA, &wp.map_a, co0, k_begin + kb * BK, &full_bar[stage]);
                tma_load_2d(smemA + stage * kStageA + BK * 128, &wp.map_a, co0 + 64, k_begin + kb * BK, &full_bar[stage]);
            }
#pragma unroll
            for (int i = 0; i < kWgPx; ++i) {
                const bool pv = pix < k_end;
                if (!p.a_tma) {
                    const bool aok = pv && a_ch_ok;
                    cp_async16(dstA + i * (kWgProd / 16) * 128,
                               aok ? (const void*)(p.dy + (long long)pix * p.ldy + a_ch) : (const void*)p.dy, aok ? 16u : 0u);
                }
                const int ih = poh * p.stride + ta - p.pad_t, iw = pow_ * p.stride + tcc - p.pad_l;
                const bool bok = pv && b_ch_ok && (unsigned)ih < (unsigned)p.H && (unsigned)iw < (unsigned)p.W;
                const long long boff = ((long long)(pb * p.H + ih) * p.W + iw) * p.ldx + ci;
                cp_async16(dstB + i * (kWgProd / 16) * 128, bok ? (const void*)(p.x + boff) : (const void*)p.x, bok ? 16u : 0u);
                pix += kWgProd / 16;
                pow_ += kWgProd / 16;
                while (pow_ >= p.OW) { pow_ -= p.OW; ++poh; }
                while (poh >= p.OH) { poh -= p.OH; ++pb; }
            }
            cp_async_arrive_noinc(&full_bar[stage]);
        }
        }
    } else if (warp == kWgProd / 32) {
        // MMA issuer: the whole warp walks the loop, elect.sync picks the issuing lane
        const uint32_t idesc = make_idesc(n_cta, 1, 1);
        const uint32_t hi = desc_hi(1024);
        const uint32_t alo0 = desc_lo(smemA, BK * 128), blo0 = desc_lo(smemB, BK * 128);
        const uint32_t tmem_u = __reduce_or_sync(0xffffffffu, tmem_base);    // warp-uniform register
        for (int kb = 0; kb < nkb; ++kb) {
            const int stage = kb % STAGES;
            mbar_wait(&full_bar[stage], (uint32_t)((kb / STAGES) & 1));
            tc_fence_after();
            const uint32_t alo = alo0 + stage * (kStageA >> 4), blo = blo0 + stage * (kStageB >> 4);
            if (elect_one()) {
#pragma unroll
                for (int k = 0; k < BK / 16; ++k)   // 16 pixels = two 8-row groups (2048 B) further down each atom
                    tc_mma2(tmem_u, alo + 128 * k, hi, blo + 128 * k, hi, idesc, (kb | k) != 0 ? 1u : 0u);
                tc_commit(&empty_bar[stage]);
            }
            __syncwarp();
        }
        if (elect_one()) tc_commit(&acc_bar);
        __syncwarp();
    }

    if (warp < 4) {
        mbar_wait(&acc_bar, 0);
        tc_fence_after();
        const int co = co0 + warp * 32 + lane;
        for (int cb = 0; cb < n_cta; cb += 16) {
            uint32_t v[16];
            tmem_ld16(tmem_base + ((uint32_t)(warp * 32) << 16) + cb, v);
            if (co < p.Cout) {
#pragma unroll
                for (int h = 0; h < 2; ++h) {       // 8-column groups never straddle a tap (ldx % 8 == 0)
                    const int nn = n0 + cb + 8 * h;
                    if (nn >= Ntot) break;
                    const int tap = nn / p.ldn, ci0 = nn - tap * p.ldn;
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        const int ci = ci0 + i;
                        if (ci < p.Cin)
                            atomicAdd(p.dw + ((size_t)tap * p.Cin + ci) * p.Cout + co, __uint_as_float(v[8 * h + i]));
                    }
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == kWgProd / 32) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)BN) : "memory");
    }
}

// ---- weight packing ------------------------------------------------------------------------------------------
// HWIO fp32 w[a][c][ci][co] -> CONV pack bf16 Wf[n = co (N rows, zero padded)][tap][cis (zero padded)]
__global__ void __launch_bounds__(256)
pack_conv_kernel(const float* __restrict__ w, int taps, int Cin, int Cout, int Cis, int N, __nv_bfloat16* __restrict__ out) {
    pdl_prologue();
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
// HWIO fp32 -> pixel-pair CONV pack (conv_halo.cu, pair mode): Wp[n][a * nq + qi][half * 8 + ci] = w[a][c][ci][n] with
// c = 2 (qmin + qi) + half + pad_l; zeros where that c is outside the filter or ci >= Cin
__global__ void __launch_bounds__(256)
pack_pair_kernel(const float* __restrict__ w, int KH, int KW, int Cin, int Cout, int N, int pad_l, int qmin, int nq,
                 __nv_bfloat16* __restrict__ out) {
    pdl_prologue();
    const long long total = (long long)N * KH * nq * 16;
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= total) return;
    const int k = (int)(idx % 16), half = k >> 3, ci = k & 7;
    const long long r = idx / 16;
    const int t = (int)(r % (KH * nq)), n = (int)(r / (KH * nq));
    const int a = t / nq, qi = t - a * nq;
    const int c = 2 * (qmin + qi) + half + pad_l;
    float v = 0.f;
    if (c >= 0 && c < KW && ci < Cin && n < Cout) v = w[((size_t)(a * KW + c) * Cin + ci) * Cout + n];
    out[idx] = __float2bfloat16_rn(v);
}
// HWIO fp32 -> ADJ pack: per parity class Wb[class][n = ci (N rows)][class tap][cos (zero padded)]
__global__ void __launch_bounds__(256)
pack_adj_kernel(const float* __restrict__ w, int KH, int KW, int Cin, int Cout, int Cos, int N, int stride, int pad_t,
                int pad_l, __nv_bfloat16* __restrict__ out) {
    pdl_prologue();
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

int check(const acg_conv_shape* s, const acg_tc_args* t, const char* who) {
    ACG_REQUIRE(s && t, ACG_ERR_INVALID, "%s: null shape/args", who);
    ACG_REQUIRE(s->stride == 1 || s->stride == 2, ACG_ERR_UNSUPPORTED, "%s: stride %d", who, s->stride);
    ACG_REQUIRE(s->KH * s->KW <= 32, ACG_ERR_UNSUPPORTED, "%s: more than 32 filter taps", who);
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

struct SplitPlan {
    int splits, kb_per;
    long long ws_bytes;
    int tickets;
};
// Split-K only pays for launches that leave most SMs idle while each CTA walks a long K loop (the 4x4 / 2x2 layers).
// tiles = CTAs that would reach the epilogue, tile_slots = tile ids the grid can produce (tickets / workspace index).
SplitPlan plan_split(long long tiles, long long tile_slots, int max_nkb) {
    SplitPlan r{1, max_nkb, 0, 0};
    if (getenv("ACG_NO_SPLITK") || tiles <= 0 || tiles * 2 > num_sms() || max_nkb < 8) return r;
    long long splits = num_sms() / tiles;
    if (splits > max_nkb / 4) splits = max_nkb / 4;          // at least 4 K blocks per CTA
    if (splits < 2) return r;
    const int kb_per = (int)((max_nkb + splits - 1) / splits);
    splits = (max_nkb + kb_per - 1) / kb_per;
    if (splits < 2) return r;
    r.splits = (int)splits;
    r.kb_per = kb_per;
    r.ws_bytes = tile_slots * splits * (long long)(BM * BN) * 4;
    r.tickets = (int)tile_slots;
    return r;
}
int max_class_taps(const acg_conv_shape* s) {
    int mx = 0;
    for (int cls = 0; cls < s->stride * s->stride; ++cls) {
        int na, nc;
        class_taps(s, cls, &na, &nc);
        if (na * nc > mx) mx = na * nc;
    }
    return mx;
}
long long adj_active_tiles(const acg_conv_shape* s, int N) {
    long long active = 0;
    for (int cls = 0; cls < s->stride * s->stride; ++cls) {
        const int ph = cls / s->stride, pw = cls % s->stride;
        const long long Mc = (long long)s->B * ((s->H - ph + s->stride - 1) / s->stride) *
                             ((s->W - pw + s->stride - 1) / s->stride);
        active += ((Mc + BM - 1) / BM) * ((N + BN - 1) / BN);
    }
    return active;
}
SplitPlan plan_fprop(const acg_conv_shape* s, int ld_in) {
    const int N = ru(s->Cout, 16);
    const long long M = (long long)s->B * s->OH * s->OW;
    const long long tiles = ((M + BM - 1) / BM) * ((N + BN - 1) / BN);
    return plan_split(tiles, tiles, (s->KH * s->KW * ld_in + BK - 1) / BK);
}
SplitPlan plan_dgrad(const acg_conv_shape* s, int ld_in) {
    const int N = ru(s->Cin, 16);
    const int Hp = (s->H + s->stride - 1) / s->stride, Wp = (s->W + s->stride - 1) / s->stride;
    const long long gx = ((long long)s->B * Hp * Wp + BM - 1) / BM, gy = (N + BN - 1) / BN;
    return plan_split(adj_active_tiles(s, N), gx * gy * s->stride * s->stride, (max_class_taps(s) * ld_in + BK - 1) / BK);
}
// applies a plan to the launch when the caller passed a large enough workspace; returns the grid.z multiplier
int apply_split(Params* p, const acg_tc_args* t, const SplitPlan& pl) {
    p->splits = 1;
    p->kb_per_split = 0;
    p->ws = nullptr;
    p->tickets = nullptr;
    if (pl.splits > 1 && t->splitk_ws && t->splitk_tickets && t->splitk_ws_bytes >= pl.ws_bytes &&
        t->splitk_n_tickets >= pl.tickets) {
        p->splits = pl.splits;
        p->kb_per_split = pl.kb_per;
        p->ws = static_cast<float*>(t->splitk_ws);
        p->tickets = t->splitk_tickets;
    }
    return p->splits;
}

int fill_bn(Params* p, const acg_tc_args* t, unsigned int total_ctas, const char* who) {
    p->dbg_skip = 0;
    p->direct_store = getenv("ACG_EPI_DIRECT") ? 1 : 0;
#ifdef ACG_PROBES
    {
        const char* e = getenv("ACG_DBG_SKIP");
        p->dbg_skip = e ? atoi(e) : 0;
    }
#endif
    p->stats = t->stats;
    p->counter = nullptr;
    p->total_ctas = total_ctas;
    p->n_stat = p->n_bias;
    p->rz = nullptr;
    if (t->red_z) {
        ACG_REQUIRE(t->stats && !t->bn_counter, ACG_ERR_INVALID,
                    "%s: the fused backward reduction accumulates into `stats` and excludes the forward finalize", who);
        ACG_REQUIRE(t->red_mean && t->red_rstd && t->red_shift, ACG_ERR_INVALID,
                    "%s: the fused backward reduction needs mean / rstd / shift of the consumer layer", who);
        ACG_REQUIRE(t->red_C > 0 && t->red_C % 16 == 0 && t->red_C <= p->n_bias && t->red_ldz % 8 == 0 &&
                        t->red_ldz >= t->red_C && t->out_dtype == ACG_BF16,
                    ACG_ERR_UNSUPPORTED, "%s: fused backward reduction: C=%d (multiple of 16, <= output channels), ldz=%d",
                    who, t->red_C, t->red_ldz);
        ACG_REQUIRE(((uintptr_t)t->red_z & 15) == 0 && ((uintptr_t)t->red_mean & 15) == 0 &&
                        ((uintptr_t)t->red_rstd & 15) == 0 && ((uintptr_t)t->red_shift & 15) == 0,
                    ACG_ERR_UNSUPPORTED, "%s: fused backward reduction needs 16-byte aligned buffers", who);
        p->rz = static_cast<const __nv_bfloat16*>(t->red_z);
        p->rz_ld = t->red_ldz;
        p->r_act = t->red_act;
        p->r_mean = t->red_mean; p->r_rstd = t->red_rstd; p->r_shift = t->red_shift;
        p->n_stat = t->red_C;
    }
    p->stats_fix = nullptr;
    p->bn_rows = 0;
    p->px.world = 0;
    if (t->stats && t->bn_counter) {
        // bn_rows == 0: the last CTA only completes the totals (deterministic workspace sum), the caller finalises
        ACG_REQUIRE(t->bn_rows == 0 || (t->bn_mean && t->bn_rstd && t->bn_scale && t->bn_shift && t->bn_rows > 0),
                    ACG_ERR_INVALID,
                    "%s: in-kernel batch-norm finalize needs mean/rstd/scale/shift buffers and the row count", who);
        p->counter = t->bn_counter;
        if (t->peer && t->peer->world > 1) {
            int prc = fill_peer_exchange(&p->px, t->peer, 2 * p->n_stat, who);
            if (prc) return prc;
        }
        p->beta = t->bn_beta;
        p->bn_mean = t->bn_mean; p->bn_rstd = t->bn_rstd; p->bn_scale = t->bn_scale; p->bn_shift = t->bn_shift;
        p->bn_rows = t->bn_rows;
        p->bn_eps = t->bn_eps;
    }
    return ACG_OK;
}

// cudaFuncAttributeMaxDynamicSharedMemorySize is a per-DEVICE attribute of a kernel: remember (kernel, device) pairs, so a
// process that drives several GPUs sets it on each of them.
// reproducible moments: the caller's limb accumulators.  With a ticket the last CTA converts them into `stats` and
// leaves them zeroed; WITHOUT one the launch only adds its limbs (fire and forget: no fence, no ticket, no last-CTA
// pass) and whoever consumes the moments completes them (acg_bn_finalize_act_fwd) and zeroes the accumulators.
void set_stats_fix(Params* p, const acg_tc_args* t) {
    p->stats_fix = nullptr;
    if (p->stats && !p->rz && t->stats_fix && t->stats_fix_len >= 6LL * p->n_stat && ((uintptr_t)t->stats_fix & 7) == 0)
        p->stats_fix = t->stats_fix;
}

int set_smem(const void* kern, int bytes) {
    static std::mutex mu;
    static std::vector<std::pair<const void*, int>> done;
    int dev = 0;
    cudaGetDevice(&dev);
    std::lock_guard<std::mutex> lk(mu);
    for (const auto& e : done)
        if (e.first == kern && e.second == dev) return ACG_OK;
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes) != cudaSuccess) {
        cudaError_t e = cudaGetLastError();
        set_error("conv_tc: cannot set %d B dynamic smem: %s", bytes, cudaGetErrorString(e));
        return ACG_ERR_CUDA;
    }
    done.emplace_back(kern, dev);
    return ACG_OK;
}

}  // namespace tc
}  // namespace acg

extern "C" {

#ifdef ACG_PROBES
/* timing experiments: returns {setup ns, main loop ns, epilogue ns, CTAs, K blocks} summed since the last call */
int acg_debug_phase_times(unsigned long long* out5 /* 8 entries */) {
    using namespace acg::tc;
    unsigned long long zero[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    if (cudaMemcpyFromSymbol(out5, g_phase_ns, sizeof(zero)) != cudaSuccess) return ACG_ERR_CUDA;
    if (cudaMemcpyToSymbol(g_phase_ns, zero, sizeof(zero)) != cudaSuccess) return ACG_ERR_CUDA;
    return ACG_OK;
}
#endif

long long acg_pack_size(const acg_conv_shape* s, int which, int ld_k) {
    using namespace acg::tc;
    if (!s || ld_k <= 0) return -1;
    if (which == 0) return (long long)ru(s->Cout, 16) * s->KH * s->KW * ld_k;      // CONV pack
    if (which == 2) {                                                              // pixel-pair CONV pack (ld_k == 16)
        int qmin, nq;
        pair_taps(s, &qmin, &nq);
        return ld_k == 16 ? (long long)ru(s->Cout, 16) * s->KH * nq * 16 : -1;
    }
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
    if (which == 2) {
        ACG_REQUIRE(ld_k == 16 && s->Cin <= 8, ACG_ERR_INVALID, "acg_pack_weights: the pixel-pair pack needs ld_k == 16, Cin <= 8");
        int qmin, nq;
        pair_taps(s, &qmin, &nq);
        const int N = ru(s->Cout, 16);
        const long long total = (long long)N * s->KH * nq * 16;
        launch_pdl(pack_pair_kernel, (int)((total + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream), w, s->KH, s->KW,
                   s->Cin, s->Cout, N, s->pad_l, qmin, nq, static_cast<__nv_bfloat16*>(pack));
        return check_launch("acg_pack_weights");
    }
    ACG_REQUIRE(ld_k % 8 == 0 && ld_k >= (which == 0 ? s->Cin : s->Cout), ACG_ERR_INVALID,
                "acg_pack_weights: ld_k=%d", ld_k);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (which == 0) {
        const int N = ru(s->Cout, 16);
        const long long total = (long long)N * s->KH * s->KW * ld_k;
        long long blocks = (total + 255) / 256;
        if (blocks > num_sms() * 8) blocks = num_sms() * 8;
        launch_pdl(pack_conv_kernel, (int)blocks, 256, 0, st, w, s->KH * s->KW, s->Cin, s->Cout, ld_k, N,
                                                     static_cast<__nv_bfloat16*>(pack));
    } else {
        const int N = ru(s->Cin, 16);
        dim3 grid(num_sms(), s->stride * s->stride);
        launch_pdl(pack_adj_kernel, grid, 256, 0, st, w, s->KH, s->KW, s->Cin, s->Cout, ld_k, N, s->stride, s->pad_t, s->pad_l,
                                             static_cast<__nv_bfloat16*>(pack));
    }
    return check_launch("acg_pack_weights");
}

int acg_pack_weights_batched(const acg_pack_job* jobs_dev, int njobs, const void* tiles_dev, int ntiles, void* stream) {
    using namespace acg;
    using namespace acg::tc;
    ACG_REQUIRE(jobs_dev && tiles_dev && njobs > 0 && ntiles > 0, ACG_ERR_INVALID,
                "acg_pack_weights_batched: bad argument");
    int blocks = ntiles < num_sms() * 8 ? ntiles : num_sms() * 8;
    launch_pdl(pack_tiles_kernel, blocks, 256, 0, static_cast<cudaStream_t>(stream), jobs_dev, static_cast<const int4*>(tiles_dev), ntiles);
    return check_launch("acg_pack_weights_batched");
}

long long acg_pack_plan(const acg_pack_job* host_jobs, int njobs, int* host_tiles, long long capacity) {
    using namespace acg;
    using namespace acg::tc;
    if (!host_jobs || njobs <= 0) return -1;
    long long n = 0;
    auto emit = [&](int job, int a, int n0, int c0, long long off) {
        if (host_tiles && n < capacity) {
            int* t = host_tiles + 4 * n;
            t[0] = job; t[1] = a; t[2] = n0 | (c0 << 16); t[3] = (int)off;
        }
        ++n;
    };
    for (int ji = 0; ji < njobs; ++ji) {
        const acg_pack_job& j = host_jobs[ji];
        if (j.KH * j.KW > 255 || j.N > 0xffff || j.ld_k > 0xffff || j.N <= 0 || j.ld_k <= 0) return -1;
        if (j.which == 0) {
            for (int tap = 0; tap < j.KH * j.KW; ++tap)
                for (int n0 = 0; n0 < j.N; n0 += 32)
                    for (int c0 = 0; c0 < j.ld_k; c0 += 32) emit(ji, tap, n0, c0, 0);
        } else if (j.which == 2) {
            if (j.ld_k != 16 || j.Cin > 8 || j.stride != 2) return -1;
            acg_conv_shape sh{};
            sh.KW = j.KW; sh.pad_l = j.pad_l;
            int qmin, nq;
            pair_taps(&sh, &qmin, &nq);
            for (int a = 0; a < j.KH; ++a)
                for (int qi = 0; qi < nq; ++qi)
                    for (int half = 0; half < 2; ++half) {
                        const int c = 2 * (qmin + qi) + half + j.pad_l;
                        const int tap = (c >= 0 && c < j.KW) ? a * j.KW + c : 255;
                        for (int n0 = 0; n0 < j.N; n0 += 32)
                            emit(ji, tap | ((a * nq + qi) << 8) | ((j.KH * nq) << 16), n0, 8 * half, 0);
                    }
        } else {
            long long off = 0;
            for (int cls = 0; cls < j.stride * j.stride; ++cls) {
                const int a0 = (cls / j.stride + j.pad_t) % j.stride, c0s = (cls % j.stride + j.pad_l) % j.stride;
                const int na = a0 < j.KH ? (j.KH - a0 + j.stride - 1) / j.stride : 0;
                const int nc = c0s < j.KW ? (j.KW - c0s + j.stride - 1) / j.stride : 0;
                const int nt = na * nc;
                for (int t = 0; t < nt; ++t) {
                    const int tap = (a0 + j.stride * (t / nc)) * j.KW + c0s + j.stride * (t % nc);
                    for (int n0 = 0; n0 < j.N; n0 += 32)
                        for (int c0 = 0; c0 < j.ld_k; c0 += 32) emit(ji, tap | (t << 8) | (nt << 16), n0, c0, off);
                }
                off += (long long)j.N * nt * j.ld_k;
                if (off > 0x7fffffffll) return -1;
            }
        }
    }
    return n;
}

int acg_conv_splitk_plan(const acg_conv_shape* s, int which, int ld_in, int* splits, long long* ws_bytes, int* n_tickets) {
    using namespace acg;
    using namespace acg::tc;
    ACG_REQUIRE(s && splits && ws_bytes && n_tickets && ld_in > 0 && (which == 0 || which == 1), ACG_ERR_INVALID,
                "acg_conv_splitk_plan: bad argument");
    const SplitPlan pl = which == 0 ? plan_fprop(s, ld_in) : plan_dgrad(s, ld_in);
    *splits = pl.splits;
    *ws_bytes = pl.splits > 1 ? pl.ws_bytes : 0;
    *n_tickets = pl.splits > 1 ? pl.tickets : 0;
    acg_tc_args t{};
    t.ld_in = ld_in;
    if (px_ok(s, &t, which)) {      // the pixel-major kernel may take the launch: room for whichever plan is larger
        int sp, tk;
        long long wb;
        px_split_plan(s, which, ld_in, ru(which == 0 ? s->Cout : s->Cin, 16), &sp, &wb, &tk);
        if (sp > *splits) *splits = sp;
        if (wb > *ws_bytes) *ws_bytes = wb;
        if (tk > *n_tickets) *n_tickets = tk;
    }
    return ACG_OK;
}

int acg_conv_kernel_kind(const acg_conv_shape* s, int which, int ld_in, int n_limit) {
    using namespace acg::tc;
    if (!s || ld_in <= 0 || (which != 0 && which != 1)) return -1;
    acg_tc_args t{};
    t.ld_in = ld_in;
    int N = ru(which == 0 ? s->Cout : s->Cin, 16);
    if (n_limit > 0 && ru(n_limit, 16) < N) N = ru(n_limit, 16);
    if (px_ok(s, &t, which)) return 2;
    if (which == 0) return (N == ru(s->Cout, 16) && halo2_conv_ok(s, &t, N)) ? 1 : 0;
    return halo2_adj_ok(s, &t, N) ? 1 : 0;
}

int acg_conv_pair_ok(const acg_conv_shape* s, int ld_in) {
    using namespace acg::tc;
    if (!s) return 0;
    acg_tc_args t{};
    t.ld_in = ld_in;
    return halo2_pair_ok(s, &t, ru(s->Cout, 16)) ? 1 : 0;
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
    int N = ru(s->Cout, 16);
    if (t->n_limit > 0 && ru(t->n_limit, 16) < N) N = ru(t->n_limit, 16);   // only the first n_limit output channels
    ACG_REQUIRE(t->ld_out >= s->Cout, ACG_ERR_INVALID, "acg_conv_fprop_tc: ld_out=%d < Cout=%d", t->ld_out, s->Cout);
    rc = set_smem((const void*)conv_tc_kernel<CONV, 3>, kSmemBytes);
    if (!rc) rc = set_smem((const void*)conv_tc_kernel<CONV, 6>, kSmemBytes6);
    if (rc) return rc;
    Params p{};
    p.a_src = static_cast<const __nv_bfloat16*>(x_bf16); p.w_pack = static_cast<const __nv_bfloat16*>(w_pack);
    p.out = y; p.bias = t->bias;
    p.B = s->B; p.H = s->H; p.W = s->W; p.OH = s->OH; p.OW = s->OW; p.KH = s->KH; p.KW = s->KW;
    p.stride = s->stride; p.pad_t = s->pad_t; p.pad_l = s->pad_l;
    p.lda = t->ld_in; p.ldo = t->ld_out; p.N = N; p.n_bias = s->Cout < N ? s->Cout : N; p.n_store = N < t->ld_out ? N : t->ld_out; p.out_dtype = t->out_dtype; p.out_act = t->out_act;
    const long long M = (long long)s->B * s->OH * s->OW;
    dim3 grid((unsigned)((M + BM - 1) / BM), (N + BN - 1) / BN, 1);
    rc = fill_bn(&p, t, grid.x * grid.y, "acg_conv_fprop_tc");
    if (rc) return rc;
    if (t->pair_x) {
        // first layers, pixel-pair mode of the halo kernel (w_pack = the pair pack, which = 2)
        ACG_REQUIRE(N == ru(s->Cout, 16) && halo2_pair_ok(s, t, N), ACG_ERR_UNSUPPORTED,
                    "acg_conv_fprop_tc: the pixel-pair mode does not cover this shape (acg_conv_pair_ok)");
        return launch_halo2(1, s, t, p, x_bf16, w_pack, N, static_cast<cudaStream_t>(stream),
                            "acg_conv_fprop_tc(halo, pixel pairs)", true);
    }
    {   // first layers (tiny K, many tiles of 4 whole 32-pixel output rows, N = 32 / 64, plain bias-free epilogue)
        const int nkb = (s->KH * s->KW * t->ld_in + BK - 1) / BK;
        const long long tiles = (M + BM - 1) / BM;
        const bool shape_ok = s->OW == 32 && (s->OH * s->OW) % BM == 0 && (N == 32 || N == 64) && s->Cout == N &&
                              s->KH * s->KW < 32 && (long long)s->H * s->W * t->ld_in < (1ll << 30);
        const bool epi_ok = !t->bias && t->out_act == ACG_ACT_NONE && !t->red_z && t->ld_out % 8 == 0 &&
                            ((uintptr_t)y & 15) == 0;
        if (nkb <= kSmallKMaxKb && shape_ok && epi_ok && tiles >= 2ll * num_sms() && !getenv("ACG_NO_SMALLK")) {
            rc = set_smem((const void*)conv_smallk_persistent_kernel<32>, SmallK<32>::kSmem);
            if (!rc) rc = set_smem((const void*)conv_smallk_persistent_kernel<64>, SmallK<64>::kSmem);
            if (rc) return rc;
            p.splits = 1;
            p.total_ctas = (unsigned int)num_sms();
            set_stats_fix(&p, t);
            ConvParams scp;
            scp.p = p;
            rc = encode_weight_map(&scp.map_b[0], w_pack, (long long)s->KH * s->KW * t->ld_in, N, "acg_conv_fprop_tc");
            if (rc) return rc;
            if (N == 32)
                launch_pdl(conv_smallk_persistent_kernel<32>, num_sms(), SmallK<32>::kThreads, SmallK<32>::kSmem,
                           static_cast<cudaStream_t>(stream), scp, (int)tiles);
            else
                launch_pdl(conv_smallk_persistent_kernel<64>, num_sms(), SmallK<64>::kThreads, SmallK<64>::kSmem,
                           static_cast<cudaStream_t>(stream), scp, (int)tiles);
            return check_launch("acg_conv_fprop_tc(small K, persistent)");
        }
    }
    if (px_ok(s, t, 0))
        // small feature maps: one output pixel x 128 images per tile, operands by TMA only (conv_px.cu)
        return launch_px(0, s, t, p, x_bf16, w_pack, N, ru(s->Cout, 16), static_cast<cudaStream_t>(stream),
                         "acg_conv_fprop_tc(pixel-major)");
    if (N == ru(s->Cout, 16) && halo2_conv_ok(s, t, N))
        // stride-2 layers with a 16- or 32-wide output: parity planes staged by TMA, every tap a shifted descriptor
        return launch_halo2(1, s, t, p, x_bf16, w_pack, N, static_cast<cudaStream_t>(stream), "acg_conv_fprop_tc(halo)");
    grid.z = (unsigned)apply_split(&p, t, plan_fprop(s, t->ld_in));
    set_stats_fix(&p, t);
    ConvParams cp;
    cp.p = p;
    rc = encode_weight_map(&cp.map_b[0], w_pack, (long long)s->KH * s->KW * t->ld_in, N, "acg_conv_fprop_tc");
    if (rc) return rc;
    if ((long long)grid.x * grid.y <= num_sms() && !getenv("ACG_CONV_3STAGE"))
        launch_pdl(conv_tc_kernel<CONV, 6>, grid, kThreads6, kSmemBytes6, static_cast<cudaStream_t>(stream), cp);
    else
        launch_pdl(conv_tc_kernel<CONV, 3>, grid, kThreads, kSmemBytes, static_cast<cudaStream_t>(stream), cp);
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
    const int Npack = ru(s->Cin, 16);        // rows per parity class of the weight pack
    int N = Npack;
    if (t->n_limit > 0 && ru(t->n_limit, 16) < N) N = ru(t->n_limit, 16);   // only the first n_limit input channels
    ACG_REQUIRE(t->ld_out >= s->Cin, ACG_ERR_INVALID, "acg_conv_dgrad_tc: ld_out=%d < Cin=%d", t->ld_out, s->Cin);
    rc = set_smem((const void*)conv_tc_kernel<ADJ, 3>, kSmemBytes);
    if (!rc) rc = set_smem((const void*)conv_tc_kernel<ADJ, 6>, kSmemBytes6);
    if (rc) return rc;
    Params p{};
    p.a_src = static_cast<const __nv_bfloat16*>(dy_bf16); p.w_pack = static_cast<const __nv_bfloat16*>(w_pack);
    p.out = dx; p.bias = t->bias;
    p.B = s->B; p.H = s->H; p.W = s->W; p.OH = s->OH; p.OW = s->OW; p.KH = s->KH; p.KW = s->KW;
    p.stride = s->stride; p.pad_t = s->pad_t; p.pad_l = s->pad_l;
    p.lda = t->ld_in; p.ldo = t->ld_out; p.N = N; p.n_bias = s->Cin < N ? s->Cin : N; p.n_store = N < t->ld_out ? N : t->ld_out; p.out_dtype = t->out_dtype; p.out_act = t->out_act;
    long long off = 0;
    const int ncls = s->stride * s->stride;
    unsigned int active = 0;
    for (int cls = 0; cls < ncls; ++cls) {
        int na, nc;
        class_taps(s, cls, &na, &nc);
        p.w_class_off[cls] = off;
        off += (long long)Npack * na * nc * t->ld_in;
        const int ph = cls / s->stride, pw = cls % s->stride;
        const long long Mc = (long long)s->B * ((s->H - ph + s->stride - 1) / s->stride) *
                             ((s->W - pw + s->stride - 1) / s->stride);
        active += (unsigned int)((Mc + BM - 1) / BM) * (unsigned int)((N + BN - 1) / BN);
    }
    const int Hp = (s->H + s->stride - 1) / s->stride, Wp = (s->W + s->stride - 1) / s->stride;
    const long long M = (long long)s->B * Hp * Wp;
    dim3 grid((unsigned)((M + BM - 1) / BM), (N + BN - 1) / BN, ncls);
    if (px_ok(s, t, 1))
        return launch_px(1, s, t, p, dy_bf16, w_pack, N, Npack, static_cast<cudaStream_t>(stream),
                         "acg_conv_dgrad_tc(pixel-major)");
    if (halo2_adj_ok(s, t, N))        // N < Npack (n_limit): the first N rows of every class' weight matrix
        // one CTA per SM walks the tile list: copies, MMAs and epilogue of consecutive tiles overlap (conv_halo.cu)
        return launch_halo2(0, s, t, p, dy_bf16, w_pack, N, static_cast<cudaStream_t>(stream), "acg_conv_dgrad_tc(halo)");
    if (N == Npack && halo_ok(s, t, N)) {
        // fused backward reduction requested: the one-tile-per-CTA predecessor stages the consumer's pre-activations
        rc = set_smem((const void*)conv_adj_halo_kernel, kHaloSmem);
        if (rc) return rc;
        const int TB = HALO_ACC / (s->W / 16);
        dim3 hgrid((unsigned)((s->B / TB) * (s->H / 32)), 1, 4);
        rc = fill_bn(&p, t, hgrid.x * 4u, "acg_conv_dgrad_tc");
        if (rc) return rc;
        set_stats_fix(&p, t);
        HaloParams hp;
        hp.p = p;
        rc = encode_halo_maps(&hp, s, t, N, TB, dy_bf16, w_pack);
        if (rc) return rc;
        launch_pdl(conv_adj_halo_kernel, hgrid, kThreads, kHaloSmem, static_cast<cudaStream_t>(stream), hp);
        return check_launch("acg_conv_dgrad_tc(halo)");
    }
    rc = fill_bn(&p, t, active, "acg_conv_dgrad_tc");
    if (rc) return rc;
    grid.z = (unsigned)(ncls * apply_split(&p, t, plan_dgrad(s, t->ld_in)));
    set_stats_fix(&p, t);
    ConvParams cp;
    cp.p = p;
    for (int cls = 0; cls < ncls; ++cls) {
        int na, nc;
        class_taps(s, cls, &na, &nc);
        if (na * nc == 0) { cp.map_b[cls] = cp.map_b[0]; continue; }
        rc = encode_weight_map(&cp.map_b[cls], static_cast<const __nv_bfloat16*>(w_pack) + p.w_class_off[cls],
                               (long long)na * nc * t->ld_in, Npack, "acg_conv_dgrad_tc", N);
        if (rc) return rc;
    }
    if ((long long)active <= num_sms() && !getenv("ACG_CONV_3STAGE"))
        launch_pdl(conv_tc_kernel<ADJ, 6>, grid, kThreads6, kSmemBytes6, static_cast<cudaStream_t>(stream), cp);
    else
        launch_pdl(conv_tc_kernel<ADJ, 3>, grid, kThreads, kSmemBytes, static_cast<cudaStream_t>(stream), cp);
    return check_launch("acg_conv_dgrad_tc");
}

int acg_conv_wgrad_tc(const acg_conv_shape* s, const void* x_bf16, const void* dy_bf16, float* dw, const acg_tc_args* t,
                      void* stream) {
    using namespace acg;
    using namespace acg::tc;
    ACG_REQUIRE(s && t && x_bf16 && dy_bf16 && dw, ACG_ERR_INVALID, "acg_conv_wgrad_tc: null pointer");
    ACG_REQUIRE(s->stride == 1 || s->stride == 2, ACG_ERR_UNSUPPORTED, "acg_conv_wgrad_tc: stride %d", s->stride);
    ACG_REQUIRE(t->ld_in % 8 == 0 && t->ld_in >= s->Cin, ACG_ERR_UNSUPPORTED,
                "acg_conv_wgrad_tc: ld_x=%d must be a multiple of 8 and >= Cin", t->ld_in);
    ACG_REQUIRE(t->ld_out % 8 == 0 && t->ld_out >= s->Cout, ACG_ERR_UNSUPPORTED,
                "acg_conv_wgrad_tc: ld_dy=%d must be a multiple of 8 and >= Cout", t->ld_out);
    ACG_REQUIRE((long long)s->B * s->H * s->W * (long long)t->ld_in < (1ll << 40) &&
                    (long long)s->B * s->OH * s->OW < (1ll << 31),
                ACG_ERR_UNSUPPORTED, "acg_conv_wgrad_tc: tensor too large");
    { int rc = set_smem((const void*)conv_wgrad_tc_kernel, kSmemBytes); if (rc) return rc; }
    WgradParams p{};
    p.x = static_cast<const __nv_bfloat16*>(x_bf16); p.dy = static_cast<const __nv_bfloat16*>(dy_bf16); p.dw = dw;
    p.B = s->B; p.H = s->H; p.W = s->W; p.OH = s->OH; p.OW = s->OW; p.KH = s->KH; p.KW = s->KW;
    p.stride = s->stride; p.pad_t = s->pad_t; p.pad_l = s->pad_l;
    p.Cin = s->Cin; p.Cout = s->Cout; p.ldx = t->ld_in; p.ldy = t->ld_out;
    // x operand by TMA (see the kernel): regular geometry, 16-byte aligned tensors, channel stride a multiple of 64.
    // (The kernel also takes strides that leave the last 64-channel atom of a tap partly empty -- p.ldn -- but measured
    // on B200 the padded MMA work costs more than the gather it replaces: g/tconv4 at ld 48: 115 vs 104 us, g/conv2 at
    // ld 32: 28.8 vs 25.4 us; ACG_WGRAD_TMA_PAD=1 enables it for strides >= 32.)
    bool x_tma = false;
    {
        const int S = s->OH * s->OW;
        const bool rows_ok = S % BK == 0 && BK % s->OW == 0, imgs_ok = S < BK && BK % S == 0;
        const long long span = (long long)(imgs_ok ? BK / S : 1) * s->H * s->W * t->ld_in;
        x_tma = (rows_ok || imgs_ok) && span < (1ll << 30) && !getenv("ACG_WGRAD_SLOW") && !getenv("ACG_WGRAD_NO_TMA") &&
                !getenv("ACG_WGRAD_NO_TMA_X") && (t->ld_in % 64 == 0 || (getenv("ACG_WGRAD_TMA_PAD") && t->ld_in >= 32)) &&
                ((uintptr_t)x_bf16 & 15) == 0 &&
                ((uintptr_t)dy_bf16 & 15) == 0 && encode_tiled_fn() != nullptr;
    }
    p.ldn = x_tma ? ru(t->ld_in, 64) : t->ld_in;
    const int Ntot = s->KH * s->KW * p.ldn;
    const int gx = (Ntot + BN - 1) / BN, gy = (s->Cout + BM - 1) / BM;
    const long long Kd = (long long)s->B * s->OH * s->OW;
    // split the pixel reduction so that the CTAs fill whole waves (two CTAs are co-resident per SM; a count just above
    // a multiple of 2 x SMs leaves a nearly empty last wave: 300 CTAs ran 30 % slower than 290), >= 4 K blocks per split
    const char* wv = getenv("ACG_WGRAD_WAVES");
    const int waves = wv ? atoi(wv) : 1;
    long long splits = ((long long)num_sms() * 2 * waves) / ((long long)gx * gy);
    const long long max_splits = (Kd + 4 * BK - 1) / (4 * BK);
    if (splits > max_splits) splits = max_splits;
    if (splits < 1) splits = 1;
    long long chunk = (Kd + splits - 1) / splits;
    chunk = (chunk + BK - 1) / BK * BK;
    splits = (Kd + chunk - 1) / chunk;
    p.k_chunk = (int)chunk;
    {
        const int S = s->OH * s->OW;
        const bool rows_ok = S % BK == 0 && BK % s->OW == 0, imgs_ok = S < BK && BK % S == 0;
        // slice-relative offsets are 32-bit: one slice spans at most 64 images or 64 rows of the input
        const long long span = (long long)(imgs_ok ? BK / S : 1) * s->H * s->W * t->ld_in;
        p.fast = (rows_ok || imgs_ok) && span < (1ll << 30) && !getenv("ACG_WGRAD_SLOW") ? 1 : 0;
    }
    dim3 grid(gx, gy, (unsigned)splits);
    WgradTmaParams wp;
    wp.p = p;
    wp.p.a_tma = 0;
    if (!getenv("ACG_WGRAD_NO_TMA") && ((uintptr_t)dy_bf16 & 15) == 0) {
        EncodeTiledFn enc = encode_tiled_fn();
        if (enc) {
            cuuint64_t dims[2] = {(cuuint64_t)t->ld_out, (cuuint64_t)Kd};
            cuuint64_t strides[1] = {(cuuint64_t)t->ld_out * 2};
            cuuint32_t box[2] = {64, (cuuint32_t)BK};
            cuuint32_t estr[2] = {1, 1};
            if (enc(&wp.map_a, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(dy_bf16), dims, strides, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS)
                wp.p.a_tma = 1;
        }
    }
    if (!wp.p.a_tma) memset(&wp.map_a, 0, sizeof(wp.map_a));
    memset(wp.map_x, 0, sizeof(wp.map_x));
    wp.p.b_tma = 0;
    if (wp.p.a_tma && x_tma) {
        // x operand by TMA: one map per input-parity plane, box = the input pixels of one 64-pixel K slice
        EncodeTiledFn enc = encode_tiled_fn();
        const int S = s->OH * s->OW;
        const cuuint32_t bw = (cuuint32_t)s->OW, bh = (cuuint32_t)(S >= BK ? BK / s->OW : s->OH),
                         bn = (cuuint32_t)(S >= BK ? 1 : BK / S);
        const int st = s->stride, nplanes = st * st;
        const cuuint64_t ld = (cuuint64_t)t->ld_in;
        bool ok = enc != nullptr;
        for (int pl = 0; ok && pl < nplanes; ++pl) {
            const int qy = pl / st, qx = pl % st;
            if (qy >= s->H || qx >= s->W) { wp.map_x[pl] = wp.map_x[0]; continue; }
            cuuint64_t dims[4] = {ld, (cuuint64_t)((s->W - qx + st - 1) / st), (cuuint64_t)((s->H - qy + st - 1) / st),
                                  (cuuint64_t)s->B};
            cuuint64_t strides[3] = {(cuuint64_t)st * ld * 2, (cuuint64_t)st * s->W * ld * 2, (cuuint64_t)s->H * s->W * ld * 2};
            cuuint32_t box[4] = {64, bw, bh, bn};
            cuuint32_t estr[4] = {1, 1, 1, 1};
            const void* base = static_cast<const __nv_bfloat16*>(x_bf16) + ((size_t)qy * s->W + qx) * t->ld_in;
            ok = enc(&wp.map_x[pl], CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(base), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
        }
        ACG_REQUIRE(ok, ACG_ERR_CUDA, "acg_conv_wgrad_tc: tensor map of the x operand failed");
        wp.p.b_tma = 1;
    } else {
        ACG_REQUIRE(!x_tma, ACG_ERR_CUDA, "acg_conv_wgrad_tc: tensor map of the dy operand failed");
    }
    launch_pdl(conv_wgrad_tc_kernel, grid, kWgThreads, kSmemBytes, static_cast<cudaStream_t>(stream), wp);
    return check_launch("acg_conv_wgrad_tc");
}

}  // extern "C"
