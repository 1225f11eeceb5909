// Shared pieces of the tcgen05 / TMEM implicit-GEMM convolution kernels (conv_tc.cu, conv_halo.cu): tile constants,
// the launch parameter block, PTX wrappers (cp.async, tcgen05.mma / commit / ld, TMA tensor copies, elect.sync),
// shared-memory matrix descriptors and the common epilogue (bias / tanh / bf16 rounding / batch-norm moments / store).
#pragma once
#include <cuda.h>
#include <stdlib.h>

#include "common.cuh"
#include "peer.cuh"

namespace acg {
namespace tc {

constexpr int BM = 128, BK = 64, BN = 128, STAGES = 3;
constexpr int kProducers = 128, kThreads = 160, kThreads6 = 288;
constexpr int kStageA = BM * BK * 2, kStageB = BN * BK * 2;
constexpr int kSmemBytes = STAGES * (kStageA + kStageB) + 1024;  // + alignment slack
constexpr int kSmemBytes6 = 6 * (kStageA + kStageB) + 1024;

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
    // fused batch-norm moments of THIS layer's output (optional)
    double* stats;              // [2][n_bias] fp64 (sum | sum of squares), accumulated with atomics
    unsigned int* counter;      // when non-NULL the last CTA to finish also finalises mean/rstd/scale/shift
    unsigned int total_ctas;    // CTAs that reach the epilogue (smaller parity classes exit early)
    const float* beta;
    float* bn_mean; float* bn_rstd; float* bn_scale; float* bn_shift;
    long long bn_rows;
    float bn_eps;
    // fused batch-norm BACKWARD reduction (optional; then `stats` is the consumer layer's `red` buffer): this launch
    // computes dA = d loss / d activation of a layer with pre-activation rz [rows][rz_ld] (bf16), and the epilogue adds
    // sum_r dzh and sum_r dzh*xhat (dzh = dA*act'(rz*rstd + shift), xhat = (rz - mean)*rstd) of its tile to stats[2][n_stat]
    const __nv_bfloat16* rz;
    int rz_ld, r_act;
    const float* r_mean; const float* r_rstd; const float* r_shift;
    int n_stat;                 // columns of stats (== n_bias for forward moments, the consumer's C for the reduction)
    // split-K (generic kernel, launches with far fewer tiles than SMs: 4x4 / 2x2 feature maps, K up to 6400): grid.z
    // carries `splits` K ranges of kb_per_split K blocks per tile; every CTA parks its fp32 accumulator tile in `ws`
    // ([tile][split][16-column chunk][128 rows][16]) and takes a ticket; the LAST CTA of a tile adds the other
    // partial tiles to its own accumulator and runs the normal epilogue (bias, moments, store).  No CTA ever waits.
    int splits, kb_per_split;
    float* ws;
    unsigned int* tickets;
    // reproducible moments (optional): integer limb accumulators [3][2][n_stat], see fix_add / cta_stats_finish
    unsigned long long* stats_fix;
    // data parallel (optional): the last CTA sums the totals over the ranks before finalising (peer.cuh); px.world == 0: off
    PeerExchange px;
    int dbg_skip;               // probe library only (-DACG_PROBES, env ACG_DBG_SKIP): see ACG_DBG below
    int direct_store;           // A/B switch ACG_EPI_DIRECT: every thread stores its own row (no warp-local transposition)
};
// Profiling probes (per-phase timing, pipelines with one stage switched off) exist only in libacg_b200_probe.so, which
// scripts/ load explicitly; in the product library ACG_DBG() is a compile-time false and the code below it vanishes.
//   1: no halo TMA   2: no weight TMA   4: no MMAs   8: per-phase %globaltimer stamps   16: epilogue loads TMEM only
//   32: no epilogue work
#ifdef ACG_PROBES
#define ACG_DBG(p, bit) (((p).dbg_skip & (bit)) != 0)
#else
#define ACG_DBG(p, bit) false
#endif

// Column sums of a 32-lane x 16-column register tile: after the butterfly lane L holds the total of column L>>1.
__device__ __forceinline__ float warp_colsum16(const float (&v)[16], int lane) {
    float w[8], x[4], y[2], z;
    const bool h16 = lane & 16, h8 = lane & 8, h4 = lane & 4, h2 = lane & 2;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const float send = h16 ? v[i] : v[i + 8];
        const float recv = __shfl_xor_sync(0xffffffffu, send, 16);
        w[i] = (h16 ? v[i + 8] : v[i]) + recv;
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const float send = h8 ? w[i] : w[i + 4];
        const float recv = __shfl_xor_sync(0xffffffffu, send, 8);
        x[i] = (h8 ? w[i + 4] : w[i]) + recv;
    }
#pragma unroll
    for (int i = 0; i < 2; ++i) {
        const float send = h4 ? x[i] : x[i + 2];
        const float recv = __shfl_xor_sync(0xffffffffu, send, 4);
        y[i] = (h4 ? x[i + 2] : x[i]) + recv;
    }
    {
        const float send = h2 ? y[0] : y[1];
        const float recv = __shfl_xor_sync(0xffffffffu, send, 2);
        z = (h2 ? y[1] : y[0]) + recv;
    }
    z += __shfl_xor_sync(0xffffffffu, z, 1);
    return z;
}

// ---- warp-local transposed stores ----------------------------------------------------------------------------------
// In every epilogue one THREAD owns one output row (tcgen05.ld 32x32b), so a warp-wide 16-byte store touches 32 different
// 128-byte lines and the L1 tag stage -- shared with the cp.async gather of the same SM -- takes ~2 clk per line: the
// 33 MB of d/conv1's output cost ~15 us of it.  Here every lane first parks the 16-byte chunks of ITS row in a per-warp
// shared-memory tile (32 rows x RB bytes, RB = 64 or 128; chunk index XOR-ed with row bits so that both the row-wise
// writes and the chunk-wise reads are bank-conflict free), then the warp stores the tile so that RB/16 consecutive lanes
// write the RB contiguous bytes of one row: 4 (RB = 128) or 8 (RB = 64) lines per instruction instead of 32.
constexpr int kEpiStageBytes = 32 * 128;           // per epilogue warp
template <int RB> __device__ __forceinline__ uint32_t epi_stage_addr(uint32_t tile, int row, int chunk) {
    static_assert(RB == 64 || RB == 128, "staging rows are 64 or 128 bytes");
    if (RB == 128) return tile + (uint32_t)row * 128u + (uint32_t)((chunk ^ (row & 7)) << 4);
    return tile + (uint32_t)row * 64u + (uint32_t)((chunk ^ ((row >> 1) & 3)) << 4);
}
template <int RB> __device__ __forceinline__ void epi_stage_put(uint32_t tile, int lane, int chunk, uint32_t a, uint32_t b,
                                                                uint32_t c, uint32_t d) {
    asm volatile("st.shared.v4.u32 [%0], {%1,%2,%3,%4};" ::"r"(epi_stage_addr<RB>(tile, lane, chunk)), "r"(a), "r"(b),
                 "r"(c), "r"(d) : "memory");
}
// out_bytes: the tensor's base + the byte offset of the tile's first column; row_off_bytes: byte offset of THIS lane's row
// (any value when !row_ok).  All 32 lanes must call.
template <int RB> __device__ __forceinline__ void epi_stage_flush(uint32_t tile, int lane, unsigned char* out_bytes,
                                                                  unsigned long long row_off_bytes, bool row_ok) {
    constexpr int LPR = RB / 16, RPI = 32 / LPR;        // lanes per row, rows per store instruction
    const unsigned long long mine = row_ok ? row_off_bytes : ~0ull;
    __syncwarp();
    const int sub = lane / LPR, c = lane % LPR;
#pragma unroll
    for (int i = 0; i < 32 / RPI; ++i) {
        const int r = i * RPI + sub;
        const unsigned long long ro = __shfl_sync(0xffffffffu, mine, r);
        uint32_t v0, v1, v2, v3;
        asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];"
                     : "=r"(v0), "=r"(v1), "=r"(v2), "=r"(v3) : "r"(epi_stage_addr<RB>(tile, r, c)));
        if (ro != ~0ull) *reinterpret_cast<uint4*>(out_bytes + ro + c * 16) = make_uint4(v0, v1, v2, v3);
    }
    __syncwarp();       // the tile may be refilled
}

// Staged epilogue of the batch-norm layers: chunk by chunk the fp32 accumulator values of a lane's row are rounded to
// bf16 and parked in the warp's staging tile (zeros for rows outside the tensor) ...
template <int RB> __device__ __forceinline__ void epi_stage_put_chunk(uint32_t tile, int lane, int chunk, const uint32_t (&v)[16],
                                                                      bool row_ok) {
    uint32_t w[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        __nv_bfloat162 h = __floats2bfloat162_rn(__uint_as_float(v[2 * k]), __uint_as_float(v[2 * k + 1]));
        w[k] = row_ok ? *reinterpret_cast<uint32_t*>(&h) : 0u;
    }
    epi_stage_put<RB>(tile, lane, 2 * chunk, w[0], w[1], w[2], w[3]);
    epi_stage_put<RB>(tile, lane, 2 * chunk + 1, w[4], w[5], w[6], w[7]);
}
// ... then the column moments of the group are read DOWN the tile -- lane L sums the 32 rows of its bf16 column pair
// (RB = 128: 64 columns per group) or column (RB = 64: 32 columns), ~7 instructions per row, instead of two 16-shuffle
// butterflies per 16-column chunk (~160 instructions per chunk: with moments the epilogue, not the MMAs, bounded
// g/tconv3's forward, 55 -> 66 us) -- and the tile leaves with full-line stores.  The sums are over exactly the values
// stored, rows in a fixed order.  sm_sum / sm_sq: THIS warp's slots of the group's first column; do_stats is warp-uniform.
template <int RB> __device__ __forceinline__ void epi_stage_moments_flush(uint32_t tile, int lane, unsigned char* out_group,
                                                                          unsigned long long row_off_bytes, bool row_ok,
                                                                          bool do_stats, float* sm_sum, float* sm_sq) {
    __syncwarp();
    if (do_stats) {
        if (RB == 128) {        // lane owns the bf16 pair (columns 2 L, 2 L + 1) = 32-bit word L of every row
            float s0a = 0.f, s0b = 0.f, s1a = 0.f, s1b = 0.f;
#pragma unroll 8
            for (int r = 0; r < 32; ++r) {
                uint32_t wv;
                asm volatile("ld.shared.u32 %0, [%1];" : "=r"(wv)
                             : "r"(epi_stage_addr<RB>(tile, r, lane >> 2) + (uint32_t)((lane & 3) << 2)));
                const float lo = __uint_as_float(wv << 16), hi = __uint_as_float(wv & 0xffff0000u);
                s0a += lo; s0b += hi;
                s1a = fmaf(lo, lo, s1a); s1b = fmaf(hi, hi, s1b);
            }
            sm_sum[2 * lane] += s0a; sm_sum[2 * lane + 1] += s0b;
            sm_sq[2 * lane] += s1a; sm_sq[2 * lane + 1] += s1b;
        } else {                // lane owns column L = 16-bit word L of every row
            float s0 = 0.f, s1 = 0.f;
#pragma unroll 8
            for (int r = 0; r < 32; ++r) {
                uint16_t hv;
                asm volatile("ld.shared.u16 %0, [%1];" : "=h"(hv)
                             : "r"(epi_stage_addr<RB>(tile, r, lane >> 3) + (uint32_t)((lane & 7) << 1)));
                const float x = __uint_as_float((uint32_t)hv << 16);
                s0 += x;
                s1 = fmaf(x, x, s1);
            }
            sm_sum[lane] += s0;
            sm_sq[lane] += s1;
        }
    }
    epi_stage_flush<RB>(tile, lane, out_group, row_off_bytes, row_ok);
}

// The same for a DATA-GRADIENT launch whose output is the activation gradient dA of a batch-norm layer (acg_tc_args.red_*):
// the backward reduction terms  sum_r dzh  and  sum_r dzh * xhat  (dzh = dA * act'(z * rstd + shift), xhat = (z - mean) * rstd,
// dA as stored) are read down the staged tile, the consumer's pre-activation z straight from global memory -- the 32
// lanes of a row read 128 (RB = 128) or 64 contiguous bytes.  rowtab: shared-memory table of this warp's 32 rows' BYTE
// offsets into z (written by epi_stage_rowtab; rows outside the tensor pass 0: their staged dA is zero, they add nothing).
// Replaces a whole acg_bn_act_bwd_reduce pass over dA and z.
__device__ __forceinline__ void epi_stage_rowtab(uint32_t rowtab, int lane, unsigned long long z_row_bytes) {
    __syncwarp();
    asm volatile("st.shared.u64 [%0], %1;" ::"r"(rowtab + (uint32_t)lane * 8u), "l"(z_row_bytes) : "memory");
    __syncwarp();
}
template <int ACT> __device__ __forceinline__ float act_bwd_ct(float u, int act_rt) {
    if (ACT == ACG_ACT_RELU) return u > 0.f ? 1.f : 0.f;
    if (ACT == ACG_ACT_LRELU) return 0.6f + 0.4f * (u > 0.f ? 1.f : (u < 0.f ? -1.f : 0.f));
    return act_bwd(u, act_rt);
}
// column sums of one group; ACT is a compile-time activation (or -1: run-time switch)
template <int RB, int ACT> __device__ __forceinline__ void epi_stage_redux_cols(const Params& p, const unsigned char* tb,
                                                                              const unsigned long long* rt,
                                                                              const unsigned char* zb, int lane, int col0,
                                                                              float* sm_sum, float* sm_sq) {
    if (RB == 128) {
        const int c = col0 + 2 * lane;
        const float mu0 = p.r_mean[c], mu1 = p.r_mean[c + 1], rs0 = p.r_rstd[c], rs1 = p.r_rstd[c + 1];
        const float sh0 = p.r_shift[c], sh1 = p.r_shift[c + 1];
        float s0a = 0.f, s0b = 0.f, s1a = 0.f, s1b = 0.f;
#pragma unroll 16
        for (int r = 0; r < 32; ++r) {
            const uint32_t zw = __ldg(reinterpret_cast<const uint32_t*>(zb + rt[r]) + lane);
            const uint32_t wv = *reinterpret_cast<const uint32_t*>(tb + (epi_stage_addr<RB>(0u, r, lane >> 2) + ((lane & 3) << 2)));
            const float dl = __uint_as_float(wv << 16), dh = __uint_as_float(wv & 0xffff0000u);
            const float zl = __uint_as_float(zw << 16), zh = __uint_as_float(zw & 0xffff0000u);
            const float ql = dl * act_bwd_ct<ACT>(fmaf(zl, rs0, sh0), p.r_act), qh = dh * act_bwd_ct<ACT>(fmaf(zh, rs1, sh1), p.r_act);
            s0a += ql; s0b += qh;
            s1a = fmaf(ql, (zl - mu0) * rs0, s1a); s1b = fmaf(qh, (zh - mu1) * rs1, s1b);
        }
        sm_sum[2 * lane] += s0a; sm_sum[2 * lane + 1] += s0b;
        sm_sq[2 * lane] += s1a; sm_sq[2 * lane + 1] += s1b;
    } else {
        const int c = col0 + lane;
        const float mu = p.r_mean[c], rs = p.r_rstd[c], sh = p.r_shift[c];
        float s0 = 0.f, s1 = 0.f;
#pragma unroll 16
        for (int r = 0; r < 32; ++r) {
            const uint16_t zv = __ldg(reinterpret_cast<const uint16_t*>(zb + rt[r]) + lane);
            const uint16_t hv = *reinterpret_cast<const uint16_t*>(tb + (epi_stage_addr<RB>(0u, r, lane >> 3) + ((lane & 7) << 1)));
            const float d = __uint_as_float((uint32_t)hv << 16), z = __uint_as_float((uint32_t)zv << 16);
            const float q = d * act_bwd_ct<ACT>(fmaf(z, rs, sh), p.r_act);
            s0 += q;
            s1 = fmaf(q, (z - mu) * rs, s1);
        }
        sm_sum[lane] += s0;
        sm_sq[lane] += s1;
    }
}
template <int RB> __device__ __forceinline__ void epi_stage_redux_flush(const Params& p, uint32_t tile, uint32_t rowtab, int lane,
                                                                        int col0, unsigned char* out_group,
                                                                        unsigned long long row_off_bytes, bool row_ok,
                                                                        float* sm_sum, float* sm_sq) {
    __syncwarp();
    // Plain C++ loads through generic pointers (not volatile asm), an UNCONDITIONAL global load per row and the activation
    // as a compile-time parameter (one switch per group, none per element), so that the row loop is ONE basic block and
    // the compiler keeps the z loads of 16 rows in flight: earlier versions paid one full memory latency per ROW
    // (g/tconv4's data gradient 77 -> 257 us) -- first behind a predicated load, then behind the branches of act_bwd().
    const unsigned char* zb = reinterpret_cast<const unsigned char*>(p.rz) + (size_t)col0 * 2;
    const unsigned long long* rt = reinterpret_cast<const unsigned long long*>(__cvta_shared_to_generic((size_t)rowtab));
    const unsigned char* tb = reinterpret_cast<const unsigned char*>(__cvta_shared_to_generic((size_t)tile));
    if (p.r_act == ACG_ACT_RELU) epi_stage_redux_cols<RB, ACG_ACT_RELU>(p, tb, rt, zb, lane, col0, sm_sum, sm_sq);
    else if (p.r_act == ACG_ACT_LRELU) epi_stage_redux_cols<RB, ACG_ACT_LRELU>(p, tb, rt, zb, lane, col0, sm_sum, sm_sq);
    else epi_stage_redux_cols<RB, -1>(p, tb, rt, zb, lane, col0, sm_sum, sm_sq);
    epi_stage_flush<RB>(tile, lane, out_group, row_off_bytes, row_ok);
}

__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src, uint32_t src_bytes) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait() {
    asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
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
// Same instruction with the two shared-memory descriptors passed as (lo, hi) 32-bit halves.  Only the low word
// (start address >> 4 | LBO << 16) changes from MMA to MMA; the issuing thread adds a constant to it instead of
// rebuilding a 64-bit descriptor with shifts and masks (the single issuing thread is latency bound: ~40 dependent
// instructions per MMA made descriptor arithmetic, not the tensor pipe, the limiter of the first version).
__device__ __forceinline__ void tc_mma2(uint32_t tmem_d, uint32_t alo, uint32_t ahi, uint32_t blo, uint32_t bhi,
                                        uint32_t idesc, uint32_t acc) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\t"
        "mov.b64 da, {%1, %2};\n\t"
        "mov.b64 db, {%3, %4};\n\t"
        "setp.ne.b32 p, %6, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %5, p;\n\t}"
        ::"r"(tmem_d), "r"(alo), "r"(ahi), "r"(blo), "r"(bhi), "r"(idesc), "r"(acc)
        : "memory");
}
// descriptor halves for 128-byte swizzle: lo = start>>4 | (LBO>>4)<<16 ; hi = SBO>>4 | version 1 | SWIZZLE_128B
__device__ __forceinline__ uint32_t desc_lo(uint32_t smem_addr, uint32_t lbo_bytes) {
    return ((smem_addr >> 4) & 0x3FFFu) | (((lbo_bytes >> 4) & 0x3FFFu) << 16);
}
__device__ __forceinline__ uint32_t desc_hi(uint32_t sbo_bytes) {
    return ((sbo_bytes >> 4) & 0x3FFFu) | (1u << 14) | (2u << 29);
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

__device__ __forceinline__ float tanh_fast(float x) {
    float y;
    asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

// One 32-row x 16-column accumulator chunk: (+bias) (tanh) -> moments -> store.  All mode tests are kernel-uniform
// and sit OUTSIDE the unrolled element loops so that they compile to branches, not to predicated instruction bloat
// (a first version that tested bias / tanh per element spent ~800 issue slots per chunk on predicated-off code).
__device__ __forceinline__ void epilogue_chunk(const Params& p, const uint32_t (&v)[16], int ncol, bool row_ok,
                                               size_t row_off, uint32_t z_smem, int lane, float* sm_sum, float* sm_sq,
                                               const __nv_bfloat16* z_gmem = nullptr) {
    float f[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) f[i] = __uint_as_float(v[i]);
    if (p.bias) {
        if (ncol + 16 <= p.n_bias) {
            const float4* b4 = reinterpret_cast<const float4*>(p.bias + ncol);
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const float4 b = b4[i];
                f[4 * i] += b.x; f[4 * i + 1] += b.y; f[4 * i + 2] += b.z; f[4 * i + 3] += b.w;
            }
        } else {
#pragma unroll
            for (int i = 0; i < 16; ++i)
                if (ncol + i < p.n_bias) f[i] += p.bias[ncol + i];
        }
    }
    if (p.out_act == ACG_ACT_TANH) {
#pragma unroll
        for (int i = 0; i < 16; ++i) f[i] = tanh_fast(f[i]);
    }
    const bool bf16_out = p.out_dtype == ACG_BF16;
    uint32_t w[8];
    if (bf16_out) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            __nv_bfloat162 h = __floats2bfloat162_rn(f[2 * i], f[2 * i + 1]);
            w[i] = *reinterpret_cast<uint32_t*>(&h);
        }
    }
    if (p.stats && ncol < p.n_stat) {      // n_stat % 16 == 0 in the reduction mode (host check)
        // batch-norm moments of exactly what is stored (bf16-rounded when the output is bf16)
        float q[16], q2[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) {
            float t = f[i];
            if (bf16_out) t = __uint_as_float((i & 1) ? (w[i >> 1] & 0xffff0000u) : (w[i >> 1] << 16));
            t = row_ok ? t : 0.f;
            q[i] = t;
        }
        if (p.rz) {
            // backward reduction terms of the layer that consumes this gradient (same arithmetic as
            // vec_col_reduce_kernel<1>): q = dA*act'(u), q2 = q*xhat
            // z_smem: this row's 16 pre-activations, staged in shared memory by stage_z_row (zero filled for rows
            // outside the tensor, whose q is zero anyway)
            float zf[16];
            {
                uint32_t zw[8];
                if (z_gmem) {     // halo kernel: dedicated epilogue warps read the row's 32 bytes straight from global
                    const uint4 z0 = __ldg(reinterpret_cast<const uint4*>(z_gmem));
                    const uint4 z1 = __ldg(reinterpret_cast<const uint4*>(z_gmem) + 1);
                    zw[0] = z0.x; zw[1] = z0.y; zw[2] = z0.z; zw[3] = z0.w;
                    zw[4] = z1.x; zw[5] = z1.y; zw[6] = z1.z; zw[7] = z1.w;
                } else {
                asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];"
                             : "=r"(zw[0]), "=r"(zw[1]), "=r"(zw[2]), "=r"(zw[3]) : "r"(z_smem));
                asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];"
                             : "=r"(zw[4]), "=r"(zw[5]), "=r"(zw[6]), "=r"(zw[7]) : "r"(z_smem + 16u));
                }
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    zf[2 * i] = __uint_as_float(zw[i] << 16);
                    zf[2 * i + 1] = __uint_as_float(zw[i] & 0xffff0000u);
                }
            }
            const float4* mu4 = reinterpret_cast<const float4*>(p.r_mean + ncol);
            const float4* rs4 = reinterpret_cast<const float4*>(p.r_rstd + ncol);
            const float4* sh4 = reinterpret_cast<const float4*>(p.r_shift + ncol);
            if (p.r_act == ACG_ACT_RELU) {
#pragma unroll
                for (int g = 0; g < 4; ++g) {
                    const float4 mu = mu4[g], rs = rs4[g], sh = sh4[g];
                    const float m_[4] = {mu.x, mu.y, mu.z, mu.w}, r_[4] = {rs.x, rs.y, rs.z, rs.w},
                                s_[4] = {sh.x, sh.y, sh.z, sh.w};
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        const int i = 4 * g + k;
                        const float u = zf[i] * r_[k] + s_[k];
                        const float d = u > 0.f ? q[i] : 0.f;
                        q[i] = d;
                        q2[i] = d * ((zf[i] - m_[k]) * r_[k]);
                    }
                }
            } else {
#pragma unroll
                for (int g = 0; g < 4; ++g) {
                    const float4 mu = mu4[g], rs = rs4[g], sh = sh4[g];
                    const float m_[4] = {mu.x, mu.y, mu.z, mu.w}, r_[4] = {rs.x, rs.y, rs.z, rs.w},
                                s_[4] = {sh.x, sh.y, sh.z, sh.w};
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        const int i = 4 * g + k;
                        const float u = zf[i] * r_[k] + s_[k];
                        const float d = q[i] * act_bwd(u, p.r_act);
                        q[i] = d;
                        q2[i] = d * ((zf[i] - m_[k]) * r_[k]);
                    }
                }
            }
        } else {
#pragma unroll
            for (int i = 0; i < 16; ++i) q2[i] = q[i] * q[i];
        }
        // sm_sum / sm_sq point into THIS WARP's slot (single writer per column: no atomics, fixed summation order)
        const float cs = warp_colsum16(q, lane), cs2 = warp_colsum16(q2, lane);
        if ((lane & 1) == 0) {
            sm_sum[lane >> 1] += cs;
            sm_sq[lane >> 1] += cs2;
        }
    }
    if (!row_ok) return;
    const bool full = ncol + 16 <= p.n_store;
    if (bf16_out) {
        __nv_bfloat16* o = static_cast<__nv_bfloat16*>(p.out) + row_off + ncol;
        if (full && (p.ldo & 7) == 0) {
            reinterpret_cast<uint4*>(o)[0] = make_uint4(w[0], w[1], w[2], w[3]);
            reinterpret_cast<uint4*>(o)[1] = make_uint4(w[4], w[5], w[6], w[7]);
        } else {
            for (int i = 0; i < 16; ++i)
                if (ncol + i < p.n_store) o[i] = __float2bfloat16_rn(f[i]);
        }
    } else {
        float* o = static_cast<float*>(p.out) + row_off + ncol;
        if (full && (p.ldo & 3) == 0) {
#pragma unroll
            for (int i = 0; i < 4; ++i)
                reinterpret_cast<float4*>(o)[i] = make_float4(f[4 * i], f[4 * i + 1], f[4 * i + 2], f[4 * i + 3]);
        } else {
            for (int i = 0; i < 16; ++i)
                if (ncol + i < p.n_store) o[i] = f[i];
        }
    }
}

// Fused backward reduction: the consumer's pre-activation rows of a tile are staged in the (by then idle) pipeline
// shared memory with cp.async right after the accumulator barrier -- one exposed L2 latency per tile instead of one
// global-load latency per 16-column chunk; the rows were pulled into L2 by a prefetch at kernel start.
constexpr int kZRowBytes = BN * 2 + 16;     // +16 B: 16-byte row accesses of a quarter warp hit distinct banks
__device__ __forceinline__ void prefetch_l2(const void* p) {
    asm volatile("prefetch.global.L2 [%0];" ::"l"(p));
}
__device__ __forceinline__ void stage_z_row(uint32_t dst, const __nv_bfloat16* src, int ncols, bool ok) {
    for (int c = 0; c < ncols; c += 8) cp_async16(dst + 2 * c, ok ? (const void*)(src + c) : (const void*)src, ok ? 16u : 0u);
}

__device__ __forceinline__ void tma_load_4d(uint32_t dst, const CUtensorMap* map, int c0, int c1, int c2, int c3,
                                            uint64_t* bar) {
    asm volatile(
        "cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5}], [%6];"
        ::"r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(smem_u32(bar))
        : "memory");
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, int c0, int c1, uint64_t* bar) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
        ::"r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(c0), "r"(c1), "r"(smem_u32(bar))
        : "memory");
}
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* map) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(map)) : "memory");
}

// Order-independent accumulation of fp64 values: v is split into three integer limbs,
//     v ~= H * 2^24 + M * 2^-8 + L * 2^-40     (H signed, M and L in [0, 2^32); truncation < 2^-40),
// each added to its own 64-bit accumulator with an integer atomic.  Integer addition is associative, so the totals do not
// depend on the order in which the CTAs arrive -- unlike fp64 atomics -- and the launch stays fire-and-forget.  The
// mapping v -> (H, M, L) is a pure function of v (its roundings need not be exact, only repeatable).  Range |v| < 2^80;
// beyond it, and for inf / nan, the value goes to the fp64 total directly so that a diverged run still shows up as such.
__device__ __forceinline__ void fix_add(const Params& p, int c, double v) {
    if (!(fabs(v) < 0x1p80)) {
        atomicAdd(&p.stats[c], v);
        return;
    }
    const int ncols = 2 * p.n_stat;
    const double h = floor(v * 0x1p-24);
    const double r = (v - h * 0x1p24) * 0x1p8;
    const double m = floor(r);
    const double l = floor((r - m) * 0x1p32);
    atomicAdd(p.stats_fix + c, (unsigned long long)(long long)h);
    atomicAdd(p.stats_fix + ncols + c, (unsigned long long)m);
    atomicAdd(p.stats_fix + 2 * ncols + c, (unsigned long long)l);
}

// CTA-level end of the fused batch-norm moments; called by ALL threads of the CTA.  s0 / s1: the fp64 totals of column
// `col` over this CTA's rows (added in a fixed order inside the CTA), held by the calling thread when has_col.
//   p.stats_fix == NULL: fp64 atomics into p.stats (the order, hence the last bits, vary from run to run);
//   p.stats_fix: integer limb accumulators (fix_add) -> the totals are bitwise reproducible; the last CTA of the launch
//     (ticket p.counter) converts them to fp64 in p.stats and leaves the accumulators zeroed for the next launch.
// The last CTA also finalises mean / rstd / scale / shift (slim.batch_norm, eps 1e-3) when p.bn_rows > 0.
__device__ __forceinline__ void cta_stats_finish(const Params& p, int tid, int nthreads, bool has_col, int col, double s0,
                                                 double s1, int* last_sh) {
    const int ncols = 2 * p.n_stat;
    if (has_col) {
        if (p.stats_fix) {
            fix_add(p, col, s0);
            fix_add(p, p.n_stat + col, s1);
        } else {
            atomicAdd(&p.stats[col], s0);
            atomicAdd(&p.stats[p.n_stat + col], s1);
        }
    }
    if (!p.counter) return;
    __threadfence();
    __syncthreads();
    if (tid == 0) *last_sh = (atomicAdd(p.counter, 1u) == p.total_ctas - 1u);
    __syncthreads();
    if (!*last_sh) return;
    __threadfence();
    if (p.stats_fix) {
        for (int c = tid; c < ncols; c += nthreads) {
            unsigned long long* a = p.stats_fix + c;
            const unsigned long long H = __ldcg(a), M = __ldcg(a + ncols), L = __ldcg(a + 2 * ncols);
            a[0] = 0ull; a[ncols] = 0ull; a[2 * ncols] = 0ull;
            // p.stats[c] is zero on entry unless a non-finite value was added to it
            p.stats[c] = __ldcg(&p.stats[c]) + limbs_to_double(H, M, L);
        }
        __threadfence();
        __syncthreads();
    }
    if (p.px.world > 1) peer_exchange(p.px, p.stats, ncols, tid, nthreads, [] {});      // SyncBN: totals over all ranks
    if (p.bn_rows > 0) {
        const double inv = 1.0 / (double)p.bn_rows;
        for (int c = tid; c < p.n_bias; c += nthreads) {
            float mu, rs, sh;
            bn_finalize_channel(__ldcg(&p.stats[c]), __ldcg(&p.stats[p.n_bias + c]), inv, p.bn_eps,
                                p.beta ? p.beta[c] : 0.f, &mu, &rs, &sh);
            p.bn_mean[c] = mu;
            p.bn_rstd[c] = rs;
            p.bn_scale[c] = rs;
            p.bn_shift[c] = sh;
        }
    }
    if (tid == 0) *p.counter = 0u;   // ready for the next launch
}

// One lane of a CONVERGED warp (elect.sync).  tcgen05.mma / TMA instructions take uniform-register operands; under a
// plain `lane == 0` branch the compiler wraps EVERY such instruction in a uniformisation loop (ELECT / PLOP3 / BRA.U.ANY,
// ~10 extra instructions at ~8 clk each: measured 95-105 clk per MMA regardless of N, round 2 knock-out probe), while
// code under an elect.sync predicate in warp-uniform control flow compiles to bare UTCHMMA + 2-3 uniform adds.
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
    return pred != 0;
}
// mbarrier wait / tcgen05.commit on a shared-memory ADDRESS (ring slots addressed by a running counter)
__device__ __forceinline__ void mbar_wait_addr(uint32_t bar, uint32_t parity) {
    uint32_t spins = 0;
    for (;;) {
        uint32_t ok;
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.b32 %0, 1, 0, p;\n\t}"
            : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
        if (ok) return;
        if (++spins > 200000000u) {
            printf("acg: mbarrier wait timed out (block %d thread %d)\n", blockIdx.x, threadIdx.x);
            __trap();
        }
    }
}
__device__ __forceinline__ void tc_commit_addr(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tmem_ld16_nowait(uint32_t taddr, uint32_t (&r)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// host helpers defined in conv_tc.cu
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn encode_tiled_fn();
int ru(int v, int m);
void class_taps(const acg_conv_shape* s, int cls, int* na, int* nc);
int fill_bn(Params* p, const acg_tc_args* t, unsigned int total_ctas, const char* who);
void set_stats_fix(Params* p, const acg_tc_args* t);
int encode_weight_map(CUtensorMap* map, const void* base, long long K, int rows, const char* who, int n_used = 0);
// pixel-major kernel for small feature maps (conv_px.cu)
bool px_ok(const acg_conv_shape* s, const acg_tc_args* t, int form);
void px_split_plan(const acg_conv_shape* s, int form, int ld_in, int N, int* splits, long long* ws_bytes, int* tickets);
int launch_px(int form, const acg_conv_shape* s, const acg_tc_args* t, const Params& p_in, const void* src,
              const void* w_pack, int N, int Npack, cudaStream_t stream, const char* who);
int set_smem(const void* kern, int bytes);

}  // namespace tc
}  // namespace acg
