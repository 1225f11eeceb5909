/*
 * acg_b200.h -- C-ABI of the B200-native hot path of action_conditioned_GANs.
 *
 * The reference (TensorFlow 1.0 graph code) has NO native layer to mirror: every device op is a
 * library call inside TF.  These entry points are therefore the boundary a
 * maintainer of the reference would bind instead of the TF ops at the cited call sites; the
 * binding stub is shown in INTEGRATION.md.  One entry point per kernel family.
 *
 * Conventions
 *  - plain pointers and sizes only; every pointer is a DEVICE pointer unless named host_*.
 *  - tensors are NHWC, weights HWIO ([kh,kw,Cin,Cout]); transposed-conv weights are
 *    [kh,kw,Cout,Cin] exactly as slim.conv2d_transpose stores them.
 *  - the caller owns all buffers; nothing here allocates device memory; workspaces are passed in.
 *  - every call only ENQUEUES work on `stream` (a cudaStream_t cast to void*), never synchronises,
 *    and is CUDA-graph capturable.
 *  - return value: 0 on success, negative on failure (ACG_ERR_*); acg_last_error() gives the text.
 *    There is no CPU fallback: without a CUDA device every compute call fails with ACG_ERR_CUDA.
 */
#ifndef ACG_B200_H_
#define ACG_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ACG_OK 0
#define ACG_ERR_INVALID (-1)     /* bad argument (null pointer, non-positive size, ...) */
#define ACG_ERR_UNSUPPORTED (-2) /* shape / dtype outside what the kernels were built for */
#define ACG_ERR_CUDA (-3)        /* a CUDA runtime call or launch failed */

/* element types of activation / logit buffers */
#define ACG_F32 0
#define ACG_BF16 1
#define ACG_U8 2   /* acg_gather_frames only: frames stored as uint8, decoded to x/127.5 - 1 (ops.py:195) */

/* activations (ops.py:22-26 lrelu; tf.nn.relu / tf.tanh at models.py:10,20,31,80) */
#define ACG_ACT_NONE 0
#define ACG_ACT_RELU 1
#define ACG_ACT_LRELU 2 /* leak 0.2 */
#define ACG_ACT_TANH 3

/* adversarial loss kinds (ops.py:28-50) */
#define ACG_LOSS_BCE 0
#define ACG_LOSS_WASS 1

int acg_version(void);
const char* acg_last_error(void);
/* number of kernel launches enqueued through this library since load (for bench.py's gpu_launches) */
long long acg_launch_count(void);

/* ------------------------------------------------------------------------------------------
 * DNA transform: replaces tf.nn.softmax + tf.extract_image_patches + stack/mul/reduce_sum at
 * models.py:60-72 (forward) and TF autodiff of it (backward w.r.t. the logits).
 *   logits [B,H,W,K*K] (f32 or bf16), img [B,H,W,C] f32, out [B,H,W,C] f32.
 *   y[b,i,j,c] = sum_p softmax(logits[b,i,j,:])_p * img[b, i+p/K-pb, j+p%K-pb, c], pb=(K-1)/2,
 *   zero outside the frame (TF SAME: K=5 pads 2/2, K=6 pads 2/3).
 * K in {5,6}, C == 3, W <= 64 and (W*C) % 4 == 0, H % 4 == 0.
 * ------------------------------------------------------------------------------------------ */
int acg_dna_fwd(const void* logits, int logits_dtype, const float* img, float* out,
                int B, int H, int W, int C, int K, void* stream);
/* dlogits = s_p * (g_p - sum_q s_q g_q), g_p = sum_c dy_c * img_{p,c}; written either dense in the logits dtype
 * (ld_dlogits == K*K) or, for fp32 logits, as bf16 rows of ld_dlogits == ru16(K*K) channels with zero pad channels
 * (the operand layout of the tensor-core convolutions, i.e. directly the dz of g/tconv4).
 * img is a fed placeholder in the reference (train.py:31-34) so no image gradient is produced. */
int acg_dna_bwd(const void* logits, int logits_dtype, const float* img, const float* dy,
                void* dlogits, int dlogits_dtype, int ld_dlogits, int B, int H, int W, int C, int K, void* stream);

/* ------------------------------------------------------------------------------------------
 * Convolution family: replaces slim.conv2d (models.py:12-15,34-37,42-51,82-88),
 * slim.conv2d_transpose (models.py:17-21,39-40,53-59) and their TF-autodiff gradients.
 * The shape always describes the FORWARD CONVOLUTION x[B,H,W,Cin] -> y[B,OH,OW,Cout]:
 *     y[b,oh,ow,co] = sum_{a,c,ci} x[b, oh*stride+a-pad_t, ow*stride+c-pad_l, ci] * w[a,c,ci,co]
 * (cross-correlation, zero padding).  conv2d_transpose IS acg_conv_dgrad of that convolution
 * (exact adjoint, which is how TF defines it), its data gradient is acg_conv_fprop and its weight
 * gradient acg_conv_wgrad with the roles of x / y swapped by the caller.
 * ------------------------------------------------------------------------------------------ */
typedef struct acg_conv_shape {
    int B, H, W, Cin;     /* conv input */
    int OH, OW, Cout;     /* conv output */
    int KH, KW, stride;   /* filter, stride (1 or 2) */
    int pad_t, pad_l;     /* zero padding before (TF SAME puts the odd element after) */
} acg_conv_shape;

/* fp32 SIMT implicit-GEMM kernels (thin layers + high-precision device reference) */
int acg_conv_fprop_f32(const acg_conv_shape* s, const float* x, const float* w, float* y, void* stream);
int acg_conv_dgrad_f32(const acg_conv_shape* s, const float* dy, const float* w, float* dx, void* stream);
/* dw += sum over the batch (caller zeroes dw; split over the batch with fp32 atomics) */
int acg_conv_wgrad_f32(const acg_conv_shape* s, const float* x, const float* dy, float* dw, void* stream);

/* tcgen05 / TMEM implicit-GEMM kernels (bf16 operands, fp32 accumulation in tensor memory).
 *   Operands are bf16 NHWC with an explicit channel stride (ld_in, a multiple of 8; the channels between the real
 *   count and ld_in must be zero) so that concat buffers and channel-padded thin layers (3 -> 16) are consumed in
 *   place.  Weights are the bf16 GEMM-ready packs made by acg_pack_weights with the SAME ld (the packed K dimension
 *   is taps x ld).  The output is written for ru16(N) channels (pad channels come out as exact zeros) with row
 *   stride ld_out, as bf16 or f32; `bias` (real channel count entries, may be NULL) and tanh (models.py:20) are
 *   applied in the epilogue. */
/* Data-parallel exchange slot of a launch whose last CTA sums its result over the ranks itself (see "Data parallelism"
 * below for the mailbox protocol and acg_peer_allreduce_f64 for the meaning of the fields). */
typedef struct acg_peer_exchange {
    void* const* mailboxes;         /* HOST array of `world` DEVICE pointers, [rank] = own segment */
    unsigned long long* epoch;      /* device uint64 of this slot on this rank */
    long long slot_off;
    int rank, world, cap;
    float timeout_s;                /* <= 0: 30 s */
} acg_peer_exchange;

typedef struct acg_tc_args {
    int ld_in;          /* channel stride of the gathered operand */
    int ld_out;         /* channel stride of the output rows, >= ru16(output channels) */
    const float* bias;  /* [output channels] or NULL */
    int out_dtype;      /* ACG_F32 | ACG_BF16 */
    int out_act;        /* ACG_ACT_NONE | ACG_ACT_TANH */
    /* optional: batch-norm moments of THIS layer fused into the epilogue (slim.batch_norm of models.py:11,32,81).
     * stats [2][C] fp64 (sum | sum of squares over all rows, of the values as stored) is accumulated with atomics
     * (caller zeroes).  If bn_counter (a zeroed uint32 in device memory) is also given, the last CTA to finish
     * writes mean / rstd / scale = rstd / shift = beta - mean*rstd (bn_rows rows, epsilon bn_eps) and resets the
     * counter, so no separate acg_bn_stats / acg_bn_finalize launch is needed on a single GPU. */
    double* stats;
    unsigned int* bn_counter;
    const float* bn_beta;   /* [C] or NULL */
    float* bn_mean; float* bn_rstd; float* bn_scale; float* bn_shift;
    long long bn_rows;
    float bn_eps;
    /* optional (data-gradient launches): fused batch-norm BACKWARD reduction.  This launch writes dA, the gradient
     * w.r.t. the activation of a layer with pre-activation red_z [rows, red_C] (bf16, row stride red_ldz, same rows
     * as the output) that was normalised with red_mean / red_rstd / red_shift and activation red_act; the epilogue
     * adds sum_r dzh and sum_r dzh*xhat (dzh = dA*act'(z*rstd + shift), xhat = (z - mean)*rstd, dA as stored) of
     * its tile to stats[2][red_C] -- exactly what acg_bn_act_bwd_reduce computes, without re-reading dA.
     * red_C % 16 == 0, red_C <= output channels, bf16 output, bn_counter == NULL. */
    const void* red_z;
    int red_ldz, red_C, red_act;
    const float* red_mean; const float* red_rstd; const float* red_shift;
    /* optional: split-K workspace for launches with far fewer output tiles than SMs (acg_conv_splitk_plan says how
     * many bytes / tickets a shape wants; tickets are uint32, zeroed once by the caller, and are left zeroed by every
     * launch).  Without a workspace the launch runs unsplit.  One workspace must not be shared by launches that can
     * run concurrently. */
    void* splitk_ws;
    long long splitk_ws_bytes;
    unsigned int* splitk_tickets;
    int splitk_n_tickets;
    /* optional: compute only the first n_limit output channels (0 = all).  The data gradient w.r.t. a concat buffer
     * whose tail is the tiled action map (models.py:16,38,84) is only needed for the feature channels. */
    int n_limit;
    /* optional: run-to-run reproducible moments.  stats_fix WITHOUT a ticket: the launch only adds its limbs and
     * acg_bn_finalize_act_fwd completes them (caller zeroes the accumulators before the launch).
     * With a ticket (bn_counter) AND stats_fix -- 6*C uint64 integer
     * accumulators (stats_fix_len elements >= 6*C), zeroed once by the caller and left zeroed by every launch -- the
     * per-CTA column totals are added as fixed-point limbs with integer atomics (associative: the CTA arrival order
     * does not matter) and the last CTA converts them into `stats` (which must be zero on entry).  Without it: fp64
     * atomics whose order varies.  bn_rows == 0 with a ticket: totals only, no finalize (the data-parallel path
     * finalises after its exchange). */
    unsigned long long* stats_fix;
    long long stats_fix_len;
    /* optional (needs bn_counter): SyncBN without an extra launch -- the last CTA exchanges the [2C] totals with the
     * peers through this slot (every rank must launch the same layer) and then finalises over bn_rows GLOBAL rows.
     * Such a launch does not trigger its programmatic dependents early (see csrc/peer.cu on why). */
    const acg_peer_exchange* peer;
    /* optional (acg_conv_fprop_tc, first layers): pixel-pair mode.  x [B,H,W,8] is read as [B,H,W/2,16] (two neighbouring
     * pixels = one 16-channel pixel), which turns the stride-2 filter into stride 1 along x: every tap is a pure shift
     * of a staged tile, no gather.  w_pack must then be the PAIR pack (acg_pack_weights which = 2, ld_k = 16);
     * acg_conv_pair_ok says whether a shape qualifies. */
    int pair_x;
} acg_tc_args;

/* y = conv(x): x [B,H,W,ld_in] -> y [B,OH,OW,ld_out]; w_pack = acg_pack_weights(which=0, ld_k=ld_in) */
int acg_conv_fprop_tc(const acg_conv_shape* s, const void* x_bf16, const void* w_pack, void* y,
                      const acg_tc_args* t, void* stream);
/* dx = conv^T(dy): dy [B,OH,OW,ld_in] -> dx [B,H,W,ld_out]; w_pack = acg_pack_weights(which=1, ld_k=ld_in) */
int acg_conv_dgrad_tc(const acg_conv_shape* s, const void* dy_bf16, const void* w_pack, void* dx,
                      const acg_tc_args* t, void* stream);
/* dw[a,c,ci,co] += sum x*dy over the batch: x [B,H,W,ld_in], dy [B,OH,OW,ld_out] (both bf16); dw fp32 HWIO */
int acg_conv_wgrad_tc(const acg_conv_shape* s, const void* x_bf16, const void* dy_bf16, float* dw,
                      const acg_tc_args* t, void* stream);
/* fp32 HWIO weights -> bf16 pack.  which=2: pixel-pair fprop pack [ru16(Cout)][KH * pair taps][16] (acg_tc_args.pair_x);
 * which=0: fprop pack [ru16(Cout)][tap][ld_k>=Cin]; which=1: dgrad pack, one
 * [ru16(Cin)][class taps][ld_k>=Cout] matrix per output-parity class.  acg_pack_size gives the element count. */
long long acg_pack_size(const acg_conv_shape* s, int which, int ld_k);
int acg_pack_weights(const acg_conv_shape* s, const float* w, int which, int ld_k, void* pack, void* stream);
/* All packs of a parameter store in one launch.  jobs_dev is a DEVICE copy of the job array (each job is what one
 * acg_pack_weights call would do; `first` is the running element count and is informational), tiles_dev a DEVICE
 * copy of the tile table acg_pack_plan wrote for the same jobs: int[4] per tile, one tile = 32 pack rows x 32 pack
 * columns of one filter tap (the CONV pack is a per-tap transpose, done through shared memory). */
typedef struct acg_pack_job {
    const void* w;       /* fp32 HWIO weights */
    void* pack;          /* bf16 output */
    long long first;
    int which, ld_k;     /* as in acg_pack_weights */
    int KH, KW, Cin, Cout, stride, pad_t, pad_l;
    int N;               /* rows of the pack: ru16(Cout) for which=0, ru16(Cin) for which=1 */
} acg_pack_job;
int acg_pack_weights_batched(const acg_pack_job* jobs_dev, int njobs, const void* tiles_dev, int ntiles,
                             void* stream);
/* HOST: number of tiles of the jobs (host_jobs is a HOST array); when host_tiles != NULL also writes up to
 * `capacity` tiles (4 ints each).  -1 on invalid jobs. */
long long acg_pack_plan(const acg_pack_job* host_jobs, int njobs, int* host_tiles, long long capacity);
/* which = 0: acg_conv_fprop_tc, 1: acg_conv_dgrad_tc (ld_in as in acg_tc_args).  splits == 1: the launch never splits. */
int acg_conv_splitk_plan(const acg_conv_shape* s, int which, int ld_in, int* splits, long long* ws_bytes,
                         int* n_tickets);
/* 1 when the tcgen05 kernels accept the shape, 0 otherwise (which: 0 fprop, 1 dgrad, 2 wgrad) */
/* which kernel a launch of this shape takes: 2 = the pixel-major kernel for small feature maps, 1 = the persistent
 * halo-tile kernel (dedicated epilogue warps), 0 = the generic kernel, -1 = bad argument.  Host only. */
int acg_conv_kernel_kind(const acg_conv_shape* s, int which, int ld_in, int n_limit);
int acg_conv_tc_supported(const acg_conv_shape* s, int which);
/* 1 when acg_conv_fprop_tc accepts acg_tc_args.pair_x for this shape (stride 2, ld_in == 8, even extents, <= 3 pair taps) */
int acg_conv_pair_ok(const acg_conv_shape* s, int ld_in);

/* ------------------------------------------------------------------------------------------
 * Batch-norm (slim.batch_norm defaults: batch statistics always, biased variance, eps 1e-3, beta
 * only; models.py:11,32,81), activations and the action-tile concat (train.py:48-50,
 * models.py:16,38,84).  `rows` = B*H*W; tensors are [rows, C] with row stride ld (in elements).
 * ------------------------------------------------------------------------------------------ */
/* stats[0:C] += sum_r z, stats[C:2C] += sum_r z^2   (fp64 accumulators, caller zeroes).
 * `groups` splits rows into equal contiguous groups with independent statistics
 * (stats is [groups][2][C]) -- the two discriminator applications of train.py:63-70. */
int acg_bn_stats(const void* z, int dtype, long long rows, int C, int ld, int groups, double* stats,
                 void* stream);
/* mean/rstd/scale/shift [groups][C] from stats; rows_per_group rows each.  With beta==NULL it is 0.
 * scale = rstd, shift = beta - mean*rstd so that bn(z) = z*scale + shift. */
int acg_bn_finalize(const double* stats, const float* beta, long long rows_per_group, int C, int groups,
                    float eps, float* mean, float* rstd, float* scale, float* shift, void* stream);
/* a[r, 0:C] = act(z[r,0:C]*scale + shift) written with row stride ld_out at channel offset 0;
 * scale==NULL means identity scale; shift==NULL means 0 (bias layers pass shift = biases). */
int acg_bn_act_fwd(const void* z, int z_dtype, long long rows, int C, int ld_in, int groups,
                   const float* scale, const float* shift, int act, void* out, int out_dtype,
                   int ld_out, void* stream);
/* The same pass writing into a CONCAT buffer: besides out[r][0:C] = act(z*scale+shift) it fills out[r][act_off : act_off +
 * n_act] with actions[r / hw][0:n_act] -- tf.concat([features, tf.tile(action)], 3) of models.py:16,38,84 without a
 * launch of its own (acg_tile_actions remains for callers that build the buffer separately). */
int acg_bn_act_fwd_cat(const void* z, int z_dtype, long long rows, int C, int ld_in, const float* scale,
                       const float* shift, int act, void* out, int out_dtype, int ld_out, const float* actions,
                       int n_act, int hw, int act_off, void* stream);
/* slim.batch_norm + activation straight from the RAW moments a convolution launch left behind (models.py:10-11,31-32,
 * 80-81): `stats` [2][C] fp64 and, optionally, the integer limb accumulators `stats_fix` [3][2][C] of a launch that got
 * acg_tc_args.stats / stats_fix but NO ticket (bn_counter == NULL: such a launch only adds its limbs -- no fence, no
 * ticket, no last-CTA pass at its end).  Every block completes the totals of its own channels, finalises them over
 * norm_rows rows (beta may be NULL), exports mean / rstd / scale / shift [C] for the backward pass and applies
 * out = act(z*scale + shift); actions != NULL additionally writes the action concat like acg_bn_act_fwd_cat.
 * bf16 z / out, C % 8 == 0, act in {none, relu, lrelu} (acg_bn_finalize_act_fwd_ok).  The caller zeroes stats AND
 * stats_fix before the convolution launch (one memset covers every layer of a network). */
int acg_bn_finalize_act_fwd_ok(int C, int ld_in, int ld_out, int act);
int acg_bn_finalize_act_fwd(const void* z, long long rows, int C, int ld_in, const double* stats,
                            const unsigned long long* stats_fix, const float* beta, long long norm_rows, float eps,
                            float* mean, float* rstd, float* scale, float* shift, int act, void* out, int ld_out,
                            const float* actions, int n_act, int hw, int act_off, void* stream);
/* backward, pass 1: dzh = (dA + dA2) * act'(z*scale+shift)   (dA2 may be NULL; it is the second consumer's
 * gradient where the graph forks -- g/tconv2 feeds both g/tconv3 and g/sconv3, models.py:40-53); red[0:C] += sum dzh, red[C:2C] += sum dzh*xhat
 * (xhat = (z-mean)*rstd; fp64, caller zeroes; [groups][2][C]). */
int acg_bn_act_bwd_reduce(const void* dA, const void* dA2, int d_dtype, int ld_d, const void* z, int z_dtype, int ld_z,
                          long long rows, int C, int groups, const float* mean, const float* rstd,
                          const float* shift, int act, double* red, void* stream);
/* Data parallel (SyncBN): pass 1 with the sum over the ranks inside the launch -- the last block pushes red[0:2C] into the
 * peers' mailboxes, waits for theirs and adds them in rank order (counter: zeroed uint32 in device memory, left zeroed).
 * Shapes the one-launch form does not cover run as acg_bn_act_bwd_reduce + acg_peer_allreduce_f64: same result. */
int acg_bn_act_bwd_reduce_sync(const void* dA, const void* dA2, int d_dtype, int ld_d, const void* z, int z_dtype,
                               int ld_z, long long rows, int C, const float* mean, const float* rstd,
                               const float* shift, int act, double* red, unsigned int* counter,
                               const acg_peer_exchange* peer, void* stream);
/* backward, pass 2: dz = rstd*(dzh - red0/R - xhat*red1/R) when has_bn, else dz = dzh, with
 * R = norm_rows (0 -> rows/groups; data-parallel SyncBN passes the GLOBAL row count after all-reducing red).
 * dbeta[c] += dbeta_scale * sum over groups of red0 (also the bias gradient of non-BN layers).
 * dz rows have stride ld_dz.  z may be NULL for a layer without batch-norm and activation (dz = dA). */
int acg_bn_act_bwd_apply(const void* dA, const void* dA2, int d_dtype, int ld_d, const void* z, int z_dtype, int ld_z,
                         long long rows, int C, int groups, const float* mean, const float* rstd,
                         const float* shift, int act, int has_bn, const double* red, void* dz,
                         int dz_dtype, int ld_dz, float* dbeta, long long norm_rows, float dbeta_scale,
                         void* stream);
/* dbias[c] += scale * red[c], c < C: bias gradient from the column sums acg_bn_act_bwd_reduce leaves in red[0:C] */
int acg_bias_grad(const double* red, int C, float scale, float* dbias, void* stream);
/* dst[r, off_dst:off_dst+n] = src[r, off_src:off_src+n]  (channel-slice copy with dtype conversion) */
int acg_copy_channels(const void* src, int src_dtype, int ld_src, int off_src, void* dst, int dst_dtype,
                      int ld_dst, int off_dst, long long rows, int n, void* stream);
/* out[r, 0:ld_out] (bf16) = [a[r,0:C] | b[r,0:C] | 0...]: the conv-operand form of a frame (g/conv1 input, b == NULL)
 * or of concat([img, frame], 3) (d/conv1 input, train.py:64,68) in one pass.  C == 3, ld_out == 8 or 16. */
int acg_pack_frames(const float* a, const float* b, int C, void* out_bf16, int ld_out, long long rows,
                    void* stream);
/* dst[(b*hw + p), off:off+A] = actions[b, 0:A]  (tf.tile of the [B,1,1,A] action map) */
int acg_tile_actions(const float* actions, int B, int hw, int A, void* dst, int dst_dtype, int ld_dst,
                     int off, void* stream);

/* ------------------------------------------------------------------------------------------
 * Losses (ops.py:19-50,100-120; train.py:72-85)
 * ------------------------------------------------------------------------------------------ */
/* frame losses in one pass over g_out / next_frame [B,H,W,3] f32:
 *   sums[0] = sum|g-n| (tf.norm ord=1, train.py:73), sums[1] = sum (g-n)^2 (build_psnr),
 *   sums[2] = gdl(next, g) (ops.py:100-120, alpha=1).   sums is fp64[3], caller zeroes.
 *   dg (may be NULL) = w_l1*sign(g-n) + w_gdl*dGDL/dg + (dadv ? dadv[..., adv_off:adv_off+3] : 0)
 *   where dadv has row stride ld_adv (the discriminator's input gradient, [B,H,W,6]). */
int acg_frame_losses(const float* g, const float* n, int B, int H, int W, double* sums, float* dg,
                     float w_l1, float w_gdl, const float* dadv, int ld_adv, int adv_off, void* stream);
/* discriminator-logit losses (tf.losses.sigmoid_cross_entropy / reduce_mean; ops.py:28-50).
 *   loss_out[0] = mean over n of  bce: max(x,0) - x*label + log1p(exp(-|x|));  wass: sign*x
 *   dlogits (may be NULL) = grad_scale * d loss / d x.   label is 1, 0.9 or 0; sign is +1/-1. */
int acg_dlogit_loss(const float* x, int n, int kind, float label_or_sign, float grad_scale,
                    float* loss_out, float* dlogits, void* stream);
/* state loss ||s - t||_F / B (train.py:77) over [B,5]; dstate (may be NULL) = grad_scale * d/ds.
 * Batch-sharded form (the norm spans the GLOBAL batch): call once with sumsq_out != NULL (writes only the local
 * fp64 sum of squares), sum that scalar over the ranks, call again with sumsq_in pointing at the global sum. */
int acg_state_loss(const float* s, const float* t, int n, float inv_batch, float grad_scale,
                   float* loss_out, float* dstate, double* sumsq_out, const double* sumsq_in, void* stream);

/* ------------------------------------------------------------------------------------------
 * Optimizers: tf.train.AdamOptimizer / RMSPropOptimizer as TF 1.0 implements them
 * (train.py:91-102) with the weight clip of train.py:89 fused (update, THEN clip).
 * Flat fp32 buffers of n elements; clip_lo > clip_hi disables the clip.
 * ------------------------------------------------------------------------------------------ */
/* lr_t = lr*sqrt(1-b2^t)/(1-b1^t) is computed by the caller: host scalar, or -- when lr_t_dev != NULL -- read from
 * device memory at kernel time so that a captured CUDA graph can be replayed with a new rate every step */
int acg_adam_step(float* p, const float* g, float* m, float* v, long long n, float lr_t, float b1,
                  float b2, float eps, float clip_lo, float clip_hi, float grad_scale, const float* lr_t_dev,
                  void* stream);
/* ms starts at ONE; p -= lr*g/sqrt(ms+eps) */
int acg_rmsprop_step(float* p, const float* g, float* ms, long long n, float lr, float decay, float eps,
                     float clip_lo, float clip_hi, float grad_scale, const float* lr_dev, void* stream);

/* ------------------------------------------------------------------------------------------
 * Data parallelism over NVLink peer memory.  The reference is single-GPU and normalises
 * over the whole batch (slim.batch_norm, models.py:11,32,81); with the batch sharded over one process per GPU the
 * [2C] fp64 moment / reduction vectors of every batch-norm layer are summed over the ranks by ONE single-CTA kernel
 * that pushes its vector into every peer's mailbox as self-validating 16-byte cells (value + epoch tags), polls the
 * cells of its own mailbox, sums in rank order (identical bits on every rank) and optionally finalises mean / rstd / scale / shift.  The flat gradient buckets stay on NCCL.
 *
 * Mailboxes are the one exception to "the caller owns all buffers": a segment that peers can map must come from
 * cudaMalloc directly, so the library allocates it.  Handles are cudaIpcMemHandle_t (64 bytes, HOST memory); the host
 * side all-gathers them (torch.distributed) and opens the peers' segments once.
 * ------------------------------------------------------------------------------------------ */
#define ACG_MAX_PEERS 8
#define ACG_PEER_HANDLE_BYTES 64
int acg_peer_alloc(long long bytes, void** out_ptr);                 /* zero-filled device segment */
int acg_peer_free(void* ptr);
int acg_peer_export(void* ptr, void* host_handle64);                  /* ptr from acg_peer_alloc */
int acg_peer_open(const void* host_handle64, void** out_ptr);         /* another process's segment */
int acg_peer_close(void* ptr);
/* bytes one exchange slot for vectors of up to `cap` doubles takes in a mailbox (256-byte multiple); -1 if invalid */
long long acg_peer_slot_bytes(int cap, int world);
/* vec[0:n] (fp64, this rank's partial sums) <- sum over ranks, through the slot at byte offset slot_off (same offset
 * on every rank) of the mailboxes host_mailboxes[0..world) (HOST array of DEVICE pointers, [rank] = own segment).
 * epoch: device uint64 owned by this slot on this rank, starts at 0, advanced by the kernel (graph-replay safe).
 * Every rank must enqueue the same sequence of exchanges per slot.  A peer that does not show up within timeout_s
 * (<= 0: 30 s) traps the kernel instead of hanging the GPU.  bn_C > 0 (then n == 2*bn_C): also write
 * mean / rstd / scale = rstd / shift = beta - mean*rstd [bn_C] from the summed moments over bn_rows GLOBAL rows. */
int acg_peer_allreduce_f64(double* vec, int n, int cap, long long slot_off, int rank, int world,
                           void* const* host_mailboxes, unsigned long long* epoch, float timeout_s, int bn_C,
                           const float* beta, long long bn_rows, float eps, float* mean, float* rstd, float* scale,
                           float* shift, void* stream);

/* ------------------------------------------------------------------------------------------
 * Device-side feeder and rollout glue.
 * acg_gather_frames replaces the host gather of the training loop: util.py:10-16 (one-hot frame masks) applied at
 * train.py:231-237,249-263 and the [-1,1] scaling of ops.py:195.  frames [N,T,frame_elems] uint8 (decoded as
 * x/127.5-1) or fp32, actions [N,T,A] fp32 stay resident in device memory; sample[B], t0[B] (int32, device) pick
 * sequence and frame index of every batch element:
 *   img[b] = frames[sample[b], t0[b]], next[b] = frames[sample[b], t0[b]+p], act[b] = actions[sample[b], t0[b]],
 *   next_state[b] = actions[sample[b], t0[b]+p, A-S:A]   (p = pair_stride, 1 for consecutive frames; next_state may
 *   be NULL; actions NULL = frames only).  The same kernel decodes a host-staged uint8 batch laid out [2,B,...]
 *   (N=1, T=2B, sample=0, t0=b, p=B).
 * acg_rollout_actions: out[b] = [acts[b, j, 0:A-S] | state[b]] (train.py:163,290); state NULL = acts[b, 0, A-S:A].
 * ------------------------------------------------------------------------------------------ */
int acg_gather_frames(const void* frames, int frames_dtype, const float* actions, const int* sample, const int* t0,
                      int N, int T, int pair_stride, int frame_elems, int A, int S, int B, float* img, float* next,
                      float* act, float* next_state, void* stream);
int acg_rollout_actions(const float* acts, int T, int j, const float* state, float* out, int B, int A, int S,
                        void* stream);

#ifdef __cplusplus
}
#endif
#endif /* ACG_B200_H_ */
