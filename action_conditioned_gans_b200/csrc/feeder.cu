// Device-side feeder and rollout glue.
//
// acg_gather_frames replaces the host-side frame-pair sampling of the reference's training loop -- the one-hot masks
// of util.py:10-16 applied at train.py:231-237,249-263 (`img[start_mask]`, `img[end_mask]`, `actions[start_mask]`,
// `state[end_mask]`) -- and the [-1,1] scaling of ops.py:195: sequences stay resident in HBM (uint8 or fp32), a step
// only ships B sample indices and B frame indices.  HBM-bound: 2 x 12 KB read (uint8) and 2 x 48 KB written per sample.
//
// acg_rollout_actions builds the action++state vector of one recursive rollout step (train.py:163, :290) from the
// action track and the state the generator predicted in the previous step, so that the whole rollout can be one
// captured graph with the state fed back on the device.
#include "common.cuh"

namespace acg {
namespace {

constexpr int kThreads = 256;

// one thread = 16 consecutive elements of one (sample, which in {img, next}) frame
template <bool U8>
__global__ void __launch_bounds__(kThreads)
gather_frames_kernel(const void* __restrict__ frames, const float* __restrict__ actions, const int* __restrict__ sample,
                     const int* __restrict__ t0, int N, int T, int pair_stride, int frame_elems, int A, int S, int B,
                     float* __restrict__ img, float* __restrict__ next, float* __restrict__ act,
                     float* __restrict__ next_state) {
    pdl_prologue();
    const int chunks = frame_elems >> 4;
    const long long total = (long long)B * 2 * chunks;
    for (long long idx = (long long)blockIdx.x * kThreads + threadIdx.x; idx < total;
         idx += (long long)gridDim.x * kThreads) {
        const int ch = (int)(idx % chunks);
        const int bw = (int)(idx / chunks);
        const int b = bw >> 1, which = bw & 1;
        int n = sample[b], t = t0[b] + which * pair_stride;
        n = min(max(n, 0), N - 1);                      // indices are validated on the host; never read out of bounds
        t = min(max(t, 0), T - 1);
        const size_t src = ((size_t)n * T + t) * frame_elems + (size_t)ch * 16;
        float4* dst = reinterpret_cast<float4*>((which ? next : img) + (size_t)b * frame_elems + (size_t)ch * 16);
        if (U8) {
            const uint4 v = *reinterpret_cast<const uint4*>(static_cast<const unsigned char*>(frames) + src);
            const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                float4 o;
                o.x = (float)(w[i] & 0xffu) * (1.f / 127.5f) - 1.f;
                o.y = (float)((w[i] >> 8) & 0xffu) * (1.f / 127.5f) - 1.f;
                o.z = (float)((w[i] >> 16) & 0xffu) * (1.f / 127.5f) - 1.f;
                o.w = (float)(w[i] >> 24) * (1.f / 127.5f) - 1.f;
                dst[i] = o;
            }
        } else {
            const float4* s4 = reinterpret_cast<const float4*>(static_cast<const float*>(frames) + src);
#pragma unroll
            for (int i = 0; i < 4; ++i) dst[i] = s4[i];
        }
    }
    // action++state of the first frame, state of the second (train.py:199,231-237)
    if (!actions) return;
    for (int i = blockIdx.x * kThreads + threadIdx.x; i < B * A; i += gridDim.x * kThreads) {
        const int b = i / A, a = i - b * A;
        const int n = min(max(sample[b], 0), N - 1), t = min(max(t0[b], 0), T - 1), t1 = min(t + pair_stride, T - 1);
        act[i] = actions[((size_t)n * T + t) * A + a];
        if (a >= A - S && next_state) next_state[b * S + (a - (A - S))] = actions[((size_t)n * T + t1) * A + a];
    }
}

__global__ void rollout_actions_kernel(const float* __restrict__ acts, int T, int j, const float* __restrict__ state,
                                       float* __restrict__ out, int B, int A, int S) {
    pdl_prologue();
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= B * A) return;
    const int b = i / A, a = i - b * A;
    if (a < A - S) out[i] = acts[((size_t)b * T + j) * A + a];
    else out[i] = state ? state[b * S + (a - (A - S))] : acts[(size_t)b * T * A + a];
}

}  // namespace
}  // namespace acg

extern "C" {

int acg_gather_frames(const void* frames, int frames_dtype, const float* actions, const int* sample, const int* t0,
                      int N, int T, int pair_stride, int frame_elems, int A, int S, int B, float* img, float* next,
                      float* act, float* next_state, void* stream) {
    using namespace acg;
    ACG_REQUIRE(frames && sample && t0 && img && next && (act || !actions), ACG_ERR_INVALID,
                "acg_gather_frames: null pointer");
    ACG_REQUIRE(pair_stride >= 1, ACG_ERR_INVALID, "acg_gather_frames: pair_stride=%d", pair_stride);
    ACG_REQUIRE(frames_dtype == ACG_F32 || frames_dtype == ACG_U8, ACG_ERR_UNSUPPORTED,
                "acg_gather_frames: frames must be uint8 or fp32");
    ACG_REQUIRE(N > 0 && T >= 2 && B > 0 && A > 0 && S >= 0 && S <= A && frame_elems > 0 && frame_elems % 16 == 0,
                ACG_ERR_INVALID, "acg_gather_frames: N=%d T=%d B=%d A=%d S=%d frame_elems=%d", N, T, B, A, S, frame_elems);
    ACG_REQUIRE(((uintptr_t)frames & 15) == 0 && ((uintptr_t)img & 15) == 0 && ((uintptr_t)next & 15) == 0,
                ACG_ERR_UNSUPPORTED, "acg_gather_frames: buffers must be 16-byte aligned");
    const long long total = (long long)B * 2 * (frame_elems >> 4);
    long long blocks = (total + kThreads - 1) / kThreads;
    const long long cap = (long long)num_sms() * 8;
    if (blocks > cap) blocks = cap;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (frames_dtype == ACG_U8)
        launch_pdl(gather_frames_kernel<true>, (int)blocks, kThreads, 0, st, frames, actions, sample, t0, N, T,
                   pair_stride, frame_elems, A, S, B, img, next, act, next_state);
    else
        launch_pdl(gather_frames_kernel<false>, (int)blocks, kThreads, 0, st, frames, actions, sample, t0, N, T,
                   pair_stride, frame_elems, A, S, B, img, next, act, next_state);
    return check_launch("acg_gather_frames");
}

int acg_rollout_actions(const float* acts, int T, int j, const float* state, float* out, int B, int A, int S,
                        void* stream) {
    using namespace acg;
    ACG_REQUIRE(acts && out, ACG_ERR_INVALID, "acg_rollout_actions: null pointer");
    ACG_REQUIRE(B > 0 && A > 0 && S >= 0 && S <= A && T > 0 && j >= 0 && j < T, ACG_ERR_INVALID,
                "acg_rollout_actions: B=%d A=%d S=%d T=%d j=%d", B, A, S, T, j);
    launch_pdl(rollout_actions_kernel, (B * A + 127) / 128, 128, 0, static_cast<cudaStream_t>(stream), acts, T, j,
               state, out, B, A, S);
    return check_launch("acg_rollout_actions");
}

}  // extern "C"
