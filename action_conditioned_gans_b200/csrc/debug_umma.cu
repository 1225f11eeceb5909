// Probe kernel (tests only): checks how tcgen05.mma addresses a 128-byte-swizzled K-major operand whose start is NOT
// 1024-byte aligned and whose 8-row groups are spaced by an arbitrary pitch -- the "shifted window" access that lets
// every filter tap of a transposed convolution read the same shared-memory halo patch.
#ifdef ACG_PROBES
#include "common.cuh"

namespace acg {
namespace {

__global__ void __launch_bounds__(128)
umma_shift_probe(const __nv_bfloat16* __restrict__ a_rows, int n_rows, const __nv_bfloat16* __restrict__ b_rows, int N,
                 int shift, int pitch, int base_offset_mode, float* __restrict__ out) {
    pdl_prologue();
    extern __shared__ unsigned char smem_raw[];
    __shared__ __align__(8) uint64_t bar;
    __shared__ uint32_t tmem_base_sh;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t smemA = base, smemB = base + 32768;
    unsigned char* gen = smem_raw + (base - smem_u32(smem_raw));
    // A: n_rows rows of 64 bf16, chunk j of absolute row R stored at R*128 + ((j ^ (R & 7)) << 4)
    for (int idx = tid; idx < n_rows * 8; idx += 128) {
        const int R = idx >> 3, j = idx & 7;
        const uint4 v = reinterpret_cast<const uint4*>(a_rows + (size_t)R * 64)[j];
        *reinterpret_cast<uint4*>(gen + R * 128 + ((j ^ (R & 7)) << 4)) = v;
    }
    for (int idx = tid; idx < N * 8; idx += 128) {
        const int R = idx >> 3, j = idx & 7;
        const uint4 v = reinterpret_cast<const uint4*>(b_rows + (size_t)R * 64)[j];
        *reinterpret_cast<uint4*>(gen + 32768 + R * 128 + ((j ^ (R & 7)) << 4)) = v;
    }
    if (tid == 0) { mbar_init(&bar, 1); fence_mbar_init(); }
    fence_proxy_async();
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_sh)),
                     "r"(128u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = tmem_base_sh;
    if (tid == 0) {
        const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | (8u << 24);
        for (int k = 0; k < 4; ++k) {
            const uint32_t a_addr = smemA + shift * 128 + k * 32, b_addr = smemB + k * 32;
            uint64_t ad = 0, bd = 0;
            ad |= (uint64_t)((a_addr >> 4) & 0x3FFF) | ((uint64_t)1 << 16) | ((uint64_t)((pitch * 128) >> 4) << 32) |
                  ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
            if (base_offset_mode == 1) ad |= (uint64_t)((a_addr >> 7) & 7) << 49;
            bd |= (uint64_t)((b_addr >> 4) & 0x3FFF) | ((uint64_t)1 << 16) | ((uint64_t)(1024 >> 4) << 32) |
                  ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
            asm volatile(
                "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem),
                "l"(ad), "l"(bd), "r"(idesc), "r"(k ? 1u : 0u)
                : "memory");
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar))
                     : "memory");
    }
    mbar_wait(&bar, 0);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    for (int cb = 0; cb < N; cb += 16) {
        uint32_t r[16];
        asm volatile(
            "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
            : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
              "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
            : "r"(tmem + ((uint32_t)(warp * 32) << 16) + cb));
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        for (int i = 0; i < 16; ++i) out[(size_t)(warp * 32 + lane) * N + cb + i] = __uint_as_float(r[i]);
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(128u) : "memory");
    }
}

}  // namespace
}  // namespace acg

extern "C" int acg_debug_umma_shift(const void* a_rows, int n_rows, const void* b_rows, int N, int shift, int pitch,
                                    int base_offset_mode, float* out, void* stream) {
    using namespace acg;
    ACG_REQUIRE(a_rows && b_rows && out, ACG_ERR_INVALID, "acg_debug_umma_shift: null pointer");
    ACG_REQUIRE(n_rows > 0 && n_rows <= 256 && N >= 16 && N <= 128 && N % 16 == 0 && shift >= 0 && pitch >= 8 &&
                    shift + 15 * pitch + 8 <= n_rows,
                ACG_ERR_INVALID, "acg_debug_umma_shift: bad geometry");
    static bool ready = false;
    if (!ready) {
        if (cudaFuncSetAttribute(umma_shift_probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 66560) != cudaSuccess) {
            cudaGetLastError();
            set_error("acg_debug_umma_shift: smem attribute");
            return ACG_ERR_CUDA;
        }
        ready = true;
    }
    launch_pdl(umma_shift_probe, 1, 128, 66560, static_cast<cudaStream_t>(stream), static_cast<const __nv_bfloat16*>(a_rows), n_rows, static_cast<const __nv_bfloat16*>(b_rows), N, shift, pitch,
        base_offset_mode, out);
    return check_launch("acg_debug_umma_shift");
}
#endif  // ACG_PROBES
