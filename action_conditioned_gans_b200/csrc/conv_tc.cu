// tcgen05 / TMEM implicit-GEMM convolution kernels (bf16 operands, fp32 accumulation in tensor memory).
// Placeholder translation unit: the weight packer is live, the tensor-core entry points report
// ACG_ERR_UNSUPPORTED until the kernels land (acg_conv_tc_supported() == 0 routes every layer to the
// fp32 SIMT kernels meanwhile).
#include "common.cuh"

namespace acg {
namespace {

// w[tap][Cin][Cout] fp32 -> fwd pack [tap][Cout][Cin] bf16, bwd pack [tap][Cin][Cout] bf16
__global__ void __launch_bounds__(256)
pack_weights_kernel(const float* __restrict__ w, int taps, int Cin, int Cout, __nv_bfloat16* __restrict__ pf,
                    __nv_bfloat16* __restrict__ pb) {
    const long long total = (long long)taps * Cin * Cout;
    for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
         idx += (long long)gridDim.x * blockDim.x) {
        const int co = (int)(idx % Cout);
        const long long r = idx / Cout;
        const int ci = (int)(r % Cin);
        const int tap = (int)(r / Cin);
        const __nv_bfloat16 v = __float2bfloat16_rn(w[idx]);
        if (pb) pb[idx] = v;
        if (pf) pf[((size_t)tap * Cout + co) * Cin + ci] = v;
    }
}

}  // namespace
}  // namespace acg

extern "C" {

int acg_pack_weights(const float* w, int taps, int Cin, int Cout, void* pack_fprop, void* pack_dgrad,
                     void* stream) {
    using namespace acg;
    ACG_REQUIRE(w, ACG_ERR_INVALID, "acg_pack_weights: null pointer");
    ACG_REQUIRE(taps > 0 && Cin > 0 && Cout > 0, ACG_ERR_INVALID, "acg_pack_weights: bad size");
    const long long total = (long long)taps * Cin * Cout;
    long long blocks = (total + 255) / 256;
    if (blocks > num_sms() * 8) blocks = num_sms() * 8;
    pack_weights_kernel<<<(int)blocks, 256, 0, static_cast<cudaStream_t>(stream)>>>(
        w, taps, Cin, Cout, static_cast<__nv_bfloat16*>(pack_fprop), static_cast<__nv_bfloat16*>(pack_dgrad));
    return check_launch("acg_pack_weights");
}

int acg_conv_tc_supported(const acg_conv_shape*, int) { return 0; }

int acg_conv_fprop_tc(const acg_conv_shape*, const void*, const void*, void*, const acg_tc_fusion*, void*) {
    acg::set_error("acg_conv_fprop_tc: not built yet");
    return ACG_ERR_UNSUPPORTED;
}
int acg_conv_dgrad_tc(const acg_conv_shape*, const void*, const void*, void*, const acg_tc_fusion*, void*) {
    acg::set_error("acg_conv_dgrad_tc: not built yet");
    return ACG_ERR_UNSUPPORTED;
}
int acg_conv_wgrad_tc(const acg_conv_shape*, const void*, const void*, float*, const acg_tc_fusion*, void*) {
    acg::set_error("acg_conv_wgrad_tc: not built yet");
    return ACG_ERR_UNSUPPORTED;
}

}  // extern "C"
