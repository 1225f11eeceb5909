"""PyTorch custom ops over the C-ABI (`torch.library`, namespace `acg::`), with autograd wiring.

This is the boundary SURVEY.md section 8(b) names: "Python host code calls hand-written sm_100a CUDA kernels as PyTorch
custom ops through a thin C-ABI layer".  Every op below is CUDA-only (`device_types="cuda"`: a CPU tensor raises
NotImplementedError -- there is no CPU fallback), takes NHWC float32 tensors like the reference's TF graph does, and
its backward is again one of our kernels:

    acg::dna(logits, img, ksize)                        models.py:60-72          bwd: acg_dna_bwd (no image gradient:
                                                                                 the frame is a placeholder, train.py:31-34)
    acg::conv2d(x, w, stride, same, bf16)               slim.conv2d              bwd: acg_conv_dgrad_* / acg_conv_wgrad_*
    acg::conv2d_transpose(x, w, stride, bf16)           slim.conv2d_transpose    bwd: acg_conv_fprop_* / acg_conv_wgrad_*
    acg::bn_act(z, beta, act)                           slim.batch_norm (batch statistics, beta only, eps 1e-3) + relu /
                                                        lrelu / none             bwd: acg_bn_act_bwd_reduce / _apply
    acg::bias_act(z, bias, act)                         bias add + activation (layers without a normalizer)
    acg::frame_losses(g, n) -> [sum|g-n|, sum (g-n)^2, gdl]   ops.py:19-20,100-120, train.py:73   bwd: same kernel
    acg::dlogit_loss(x, kind, label_or_sign)            ops.py:28-50             bwd: same kernel
    acg::adam_step / acg::rmsprop_step                  train.py:91-102 (+ clip of train.py:89); in-place, no autograd
    acg::generator_transform / acg::generator / acg::discriminator
                                                        whole networks of models.py:8-88 on the hand-scheduled engine
                                                        (tcgen05 kernels, bf16 operands); bwd = engine backward

`models.build_*` and `ops.*` are thin wrappers over these, so `build_generator_transform(...)[0].sum().backward()`
works and is tested against the oracle's autograd (tests/test_torch_ops_gpu.py).
"""
import torch
from torch.library import custom_op

from . import engine as E
from . import kernels as K

BN_EPS = E.BN_EPS


def _c(t):
    return t.contiguous()


# ---- DNA ----------------------------------------------------------------------------------------------------
@custom_op("acg::dna", mutates_args=(), device_types="cuda")
def dna(logits: torch.Tensor, img: torch.Tensor, ksize: int) -> torch.Tensor:
    logits, img = _c(logits.float()), _c(img.float())
    out = torch.empty_like(img)
    K.dna_fwd(logits, img, out, ksize)
    return out


@dna.register_fake
def _(logits, img, ksize):
    return torch.empty_like(img, dtype=torch.float32)


def _dna_setup(ctx, inputs, output):
    logits, img, ksize = inputs
    ctx.save_for_backward(logits, img)
    ctx.ksize = ksize


def _dna_bwd(ctx, dy):
    logits, img = ctx.saved_tensors
    return dna_grad(_c(logits.float()), _c(img.float()), _c(dy.float()), ctx.ksize), None, None


@custom_op("acg::dna_grad", mutates_args=(), device_types="cuda")
def dna_grad(logits: torch.Tensor, img: torch.Tensor, dy: torch.Tensor, ksize: int) -> torch.Tensor:
    dl = torch.empty_like(logits)
    K.dna_bwd(logits, img, dy, dl, ksize)
    return dl


@dna_grad.register_fake
def _(logits, img, dy, ksize):
    return torch.empty_like(logits)


dna.register_autograd(_dna_bwd, setup_context=_dna_setup)


# ---- convolutions --------------------------------------------------------------------------------------------
def _ru(v, m):
    return (v + m - 1) // m * m


def _conv_shape(x_shape, w_shape, stride, same):
    B, H, W, Cin = x_shape
    k = w_shape[0]
    return K.conv_shape(B, H, W, Cin, w_shape[3], k, stride, "SAME" if same else "VALID")


def _pad_bf16(t, ld):
    """[B,H,W,C] fp32 -> bf16 [B,H,W,ld] with zero pad channels (the tensor-core operand layout)"""
    B, H, W, Cc = t.shape
    if ld == Cc:
        return _c(t.to(torch.bfloat16))
    out = torch.zeros(B, H, W, ld, dtype=torch.bfloat16, device=t.device)
    out[..., :Cc] = t
    return out


def _conv_fwd_impl(shape, x, w, bf16):
    y = torch.empty(shape.B, shape.OH, shape.OW, shape.Cout, device=x.device)
    if not bf16:
        K.conv_fprop_f32(shape, x, w, y)
        return y
    ld = _ru(shape.Cin, 16)
    pack = torch.empty(K.pack_size(shape, 0, ld), dtype=torch.bfloat16, device=x.device)
    K.pack_weights(shape, w, 0, ld, pack)
    yp = torch.empty(shape.B, shape.OH, shape.OW, _ru(shape.Cout, 16), device=x.device)
    K.conv_fprop_tc(shape, _pad_bf16(x, ld), pack, yp, ld, yp.shape[3])
    return _c(yp[..., :shape.Cout])


def _conv_dgrad_impl(shape, dy, w, bf16):
    dx = torch.empty(shape.B, shape.H, shape.W, shape.Cin, device=dy.device)
    if not bf16:
        K.conv_dgrad_f32(shape, dy, w, dx)
        return dx
    ld = _ru(shape.Cout, 16)
    pack = torch.empty(K.pack_size(shape, 1, ld), dtype=torch.bfloat16, device=dy.device)
    K.pack_weights(shape, w, 1, ld, pack)
    dxp = torch.empty(shape.B, shape.H, shape.W, _ru(shape.Cin, 16), device=dy.device)
    K.conv_dgrad_tc(shape, _pad_bf16(dy, ld), pack, dxp, ld, dxp.shape[3])
    return _c(dxp[..., :shape.Cin])


def _conv_wgrad_impl(shape, x, dy, bf16):
    dw = torch.zeros(shape.KH, shape.KW, shape.Cin, shape.Cout, device=x.device)
    if not bf16:
        K.conv_wgrad_f32(shape, x, dy, dw)
        return dw
    ldx, ldy = _ru(shape.Cin, 16), _ru(shape.Cout, 16)
    K.conv_wgrad_tc(shape, _pad_bf16(x, ldx), _pad_bf16(dy, ldy), dw, ldx, ldy)
    return dw


@custom_op("acg::conv2d", mutates_args=(), device_types="cuda")
def conv2d(x: torch.Tensor, w: torch.Tensor, stride: int, same: bool, bf16: bool) -> torch.Tensor:
    """x [B,H,W,Cin], w HWIO [k,k,Cin,Cout]; TF SAME / VALID padding; cross-correlation"""
    x, w = _c(x.float()), _c(w.float())
    return _conv_fwd_impl(_conv_shape(x.shape, w.shape, stride, same), x, w, bf16)


@conv2d.register_fake
def _(x, w, stride, same, bf16):
    s = _conv_shape(x.shape, w.shape, stride, same)
    return x.new_empty(s.B, s.OH, s.OW, s.Cout, dtype=torch.float32)


@custom_op("acg::conv2d_dgrad", mutates_args=(), device_types="cuda")
def conv2d_dgrad(dy: torch.Tensor, w: torch.Tensor, H: int, W: int, stride: int, same: bool, bf16: bool) -> torch.Tensor:
    """adjoint of acg::conv2d w.r.t. x ( == slim.conv2d_transpose with w read as [k,k,Cout_t,Cin_t])"""
    dy, w = _c(dy.float()), _c(w.float())
    shape = _conv_shape((dy.shape[0], H, W, w.shape[2]), w.shape, stride, same)
    return _conv_dgrad_impl(shape, dy, w, bf16)


@conv2d_dgrad.register_fake
def _(dy, w, H, W, stride, same, bf16):
    return dy.new_empty(dy.shape[0], H, W, w.shape[2], dtype=torch.float32)


@custom_op("acg::conv2d_wgrad", mutates_args=(), device_types="cuda")
def conv2d_wgrad(x: torch.Tensor, dy: torch.Tensor, k: int, stride: int, same: bool, bf16: bool) -> torch.Tensor:
    x, dy = _c(x.float()), _c(dy.float())
    shape = _conv_shape(x.shape, (k, k, x.shape[3], dy.shape[3]), stride, same)
    return _conv_wgrad_impl(shape, x, dy, bf16)


@conv2d_wgrad.register_fake
def _(x, dy, k, stride, same, bf16):
    return x.new_empty(k, k, x.shape[3], dy.shape[3], dtype=torch.float32)


def _conv2d_setup(ctx, inputs, output):
    x, w, stride, same, bf16 = inputs
    ctx.save_for_backward(x, w)
    ctx.args = (stride, same, bf16)


def _conv2d_bwd(ctx, dy):
    x, w = ctx.saved_tensors
    stride, same, bf16 = ctx.args
    dx = conv2d_dgrad(dy, w, x.shape[1], x.shape[2], stride, same, bf16) if ctx.needs_input_grad[0] else None
    dw = conv2d_wgrad(x, dy, w.shape[0], stride, same, bf16) if ctx.needs_input_grad[1] else None
    return dx, dw, None, None, None


conv2d.register_autograd(_conv2d_bwd, setup_context=_conv2d_setup)


def _dgrad_setup(ctx, inputs, output):
    dy, w, H, W, stride, same, bf16 = inputs
    ctx.save_for_backward(dy, w)
    ctx.args = (stride, same, bf16)


def _dgrad_bwd(ctx, g):
    """y = conv_dgrad(dy, w) is linear in both: d/d(dy) = conv2d(g, w), d/dw = wgrad(x = g, dy = dy)"""
    dy, w = ctx.saved_tensors
    stride, same, bf16 = ctx.args
    d_dy = conv2d(g, w, stride, same, bf16) if ctx.needs_input_grad[0] else None
    d_w = conv2d_wgrad(g, dy, w.shape[0], stride, same, bf16) if ctx.needs_input_grad[1] else None
    return d_dy, d_w, None, None, None, None, None


conv2d_dgrad.register_autograd(_dgrad_bwd, setup_context=_dgrad_setup)


def conv2d_transpose(x, w, stride=2, bf16=False):
    """slim.conv2d_transpose, SAME: x [B,h,w,Cin_t], w [k,k,Cout_t,Cin_t] -> [B,h*stride,w*stride,Cout_t].
    It IS the data gradient of the SAME conv [B,h*s,w*s,Cout_t] -> [B,h,w,Cin_t] with HWIO weights w."""
    return conv2d_dgrad(x, w, x.shape[1] * stride, x.shape[2] * stride, stride, True, bf16)


# ---- batch-norm / bias + activation -----------------------------------------------------------------------------
@custom_op("acg::bn_act", mutates_args=(), device_types="cuda")
def bn_act(z: torch.Tensor, beta: torch.Tensor, act: str) -> torch.Tensor:
    z, beta = _c(z.float()), _c(beta.float())
    Cc = z.shape[-1]
    rows = z.numel() // Cc
    stats = torch.zeros(2 * Cc, dtype=torch.float64, device=z.device)
    coef = torch.empty(4, Cc, device=z.device)
    K.bn_stats(z, rows, Cc, Cc, 1, stats)
    K.bn_finalize(stats, beta, rows, Cc, 1, coef[0], coef[1], coef[2], coef[3], BN_EPS)
    out = torch.empty_like(z)
    K.bn_act_fwd(z, rows, Cc, Cc, 1, coef[2], coef[3], act, out, Cc)
    return out


@bn_act.register_fake
def _(z, beta, act):
    return torch.empty_like(z, dtype=torch.float32)


@custom_op("acg::bn_act_grad", mutates_args=(), device_types="cuda")
def bn_act_grad(dA: torch.Tensor, z: torch.Tensor, beta: torch.Tensor, act: str, has_bn: bool) -> list[torch.Tensor]:
    """-> [dz, dbeta (or dbias)]; recomputes the batch statistics (cheap next to saving them through autograd)"""
    dA, z, beta = _c(dA.float()), _c(z.float()), _c(beta.float())
    Cc = z.shape[-1]
    rows = z.numel() // Cc
    red = torch.zeros(2 * Cc, dtype=torch.float64, device=z.device)
    dz, dpar = torch.empty_like(z), torch.zeros(Cc, device=z.device)
    if has_bn:
        stats = torch.zeros(2 * Cc, dtype=torch.float64, device=z.device)
        coef = torch.empty(4, Cc, device=z.device)
        K.bn_stats(z, rows, Cc, Cc, 1, stats)
        K.bn_finalize(stats, beta, rows, Cc, 1, coef[0], coef[1], coef[2], coef[3], BN_EPS)
        mean, rstd, shift = coef[0], coef[1], coef[3]
    else:
        mean, rstd, shift = None, None, beta
    K.bn_act_bwd_reduce(dA, None, Cc, z, Cc, rows, Cc, 1, mean, rstd, shift, act, red)
    K.bn_act_bwd_apply(dA, None, Cc, z, Cc, rows, Cc, 1, mean, rstd, shift, act, has_bn, red, dz, dpar)
    return [dz, dpar]


@bn_act_grad.register_fake
def _(dA, z, beta, act, has_bn):
    return [torch.empty_like(z, dtype=torch.float32), torch.empty_like(beta, dtype=torch.float32)]


def _bn_setup(ctx, inputs, output):
    z, beta, act = inputs
    ctx.save_for_backward(z, beta)
    ctx.act = act


def _bn_bwd(ctx, dA):
    z, beta = ctx.saved_tensors
    dz, dbeta = bn_act_grad(dA, z, beta, ctx.act, True)
    return dz, dbeta, None


bn_act.register_autograd(_bn_bwd, setup_context=_bn_setup)


@custom_op("acg::bias_act", mutates_args=(), device_types="cuda")
def bias_act(z: torch.Tensor, bias: torch.Tensor, act: str) -> torch.Tensor:
    z, bias = _c(z.float()), _c(bias.float())
    Cc = z.shape[-1]
    out = torch.empty_like(z)
    K.bn_act_fwd(z, z.numel() // Cc, Cc, Cc, 1, None, bias, act, out, Cc)
    return out


@bias_act.register_fake
def _(z, bias, act):
    return torch.empty_like(z, dtype=torch.float32)


def _bias_bwd(ctx, dA):
    z, bias = ctx.saved_tensors
    dz, dbias = bn_act_grad(dA, z, bias, ctx.act, False)
    return dz, dbias, None


bias_act.register_autograd(_bias_bwd, setup_context=_bn_setup)


# ---- losses ---------------------------------------------------------------------------------------------------
@custom_op("acg::frame_losses", mutates_args=(), device_types="cuda")
def frame_losses(g: torch.Tensor, n: torch.Tensor) -> torch.Tensor:
    """-> float64 [3]: sum|g-n| (tf.norm ord=1, train.py:73), sum (g-n)^2 (build_psnr), gdl(n, g) (ops.py:100-120)"""
    g, n = _c(g.float()), _c(n.float())
    sums = torch.zeros(3, dtype=torch.float64, device=g.device)
    K.frame_losses(g, n, sums)
    return sums


@frame_losses.register_fake
def _(g, n):
    return g.new_empty(3, dtype=torch.float64)


@custom_op("acg::frame_losses_grad", mutates_args=(), device_types="cuda")
def frame_losses_grad(g: torch.Tensor, n: torch.Tensor, w_l1: float, w_gdl: float) -> torch.Tensor:
    g, n = _c(g.float()), _c(n.float())
    sums = torch.zeros(3, dtype=torch.float64, device=g.device)
    dg = torch.empty_like(g)
    K.frame_losses(g, n, sums, dg, w_l1, w_gdl)
    return dg


@frame_losses_grad.register_fake
def _(g, n, w_l1, w_gdl):
    return torch.empty_like(g, dtype=torch.float32)


def _fl_setup(ctx, inputs, output):
    ctx.save_for_backward(*inputs)


def _fl_bwd(ctx, dsums):
    """d/dg of w0*L1 + w1*MSE_sum + w2*GDL; the fused kernel has the L1 and GDL parts, the squared error is 2(g-n)"""
    g, n = ctx.saved_tensors
    w = dsums.double().cpu().tolist()
    dg = frame_losses_grad(g, n, float(w[0]), float(w[2]))
    if w[1] != 0.0:
        dg = dg + (2.0 * w[1]) * (g - n)
    return dg, None


frame_losses.register_autograd(_fl_bwd, setup_context=_fl_setup)


@custom_op("acg::dlogit_loss", mutates_args=(), device_types="cuda")
def dlogit_loss(x: torch.Tensor, kind: str, label_or_sign: float) -> torch.Tensor:
    x = _c(x.float())
    out = torch.zeros(1, device=x.device)
    K.dlogit_loss(x, x.numel(), kind, label_or_sign, 1.0, out, None)
    return out[0].clone()


@dlogit_loss.register_fake
def _(x, kind, label_or_sign):
    return x.new_empty((), dtype=torch.float32)


@custom_op("acg::dlogit_loss_grad", mutates_args=(), device_types="cuda")
def dlogit_loss_grad(x: torch.Tensor, kind: str, label_or_sign: float) -> torch.Tensor:
    x = _c(x.float())
    out = torch.zeros(1, device=x.device)
    dx = torch.empty(x.numel(), device=x.device)
    K.dlogit_loss(x, x.numel(), kind, label_or_sign, 1.0, out, dx)
    return dx.view(x.shape)


@dlogit_loss_grad.register_fake
def _(x, kind, label_or_sign):
    return torch.empty_like(x, dtype=torch.float32)


def _dl_setup(ctx, inputs, output):
    x, kind, los = inputs
    ctx.save_for_backward(x)
    ctx.args = (kind, los)


def _dl_bwd(ctx, g):
    (x,) = ctx.saved_tensors
    return g * dlogit_loss_grad(x, *ctx.args), None, None


dlogit_loss.register_autograd(_dl_bwd, setup_context=_dl_setup)


# ---- optimizers (in place; train.py:89-102) ------------------------------------------------------------------------
@custom_op("acg::adam_step", mutates_args=("p", "m", "v"), device_types="cuda")
def adam_step(p: torch.Tensor, g: torch.Tensor, m: torch.Tensor, v: torch.Tensor, lr_t: float, clip_lo: float,
              clip_hi: float) -> None:
    """tf.train.AdamOptimizer update with lr_t = lr*sqrt(1-b2^t)/(1-b1^t) (epsilon outside the corrected root);
    clip_lo > clip_hi disables the weight clip of train.py:89"""
    K.adam_step(p, g, m, v, lr_t, clip=(clip_lo, clip_hi))


@custom_op("acg::rmsprop_step", mutates_args=("p", "ms"), device_types="cuda")
def rmsprop_step(p: torch.Tensor, g: torch.Tensor, ms: torch.Tensor, lr: float, clip_lo: float, clip_hi: float) -> None:
    K.rmsprop_step(p, g, ms, lr, clip=(clip_lo, clip_hi))


# ---- whole networks on the hand-scheduled engine ------------------------------------------------------------------
# The op takes the scope's FLAT parameter buffer (engine.ParamStore.flat: every variable of the scope in TF order), so
# that autograd sees one leaf per scope; the registry maps the buffer to its store and to the cached NetRun.
_STORES = {}      # data_ptr of flat -> ParamStore
_RUNS = {}


def register_store(store):
    _STORES[store.flat.data_ptr()] = store
    return store.flat


def _store_of(flat):
    st = _STORES.get(flat.data_ptr())
    if st is None:
        raise RuntimeError("acg network ops need the flat parameter buffer of a registered engine.ParamStore")
    return st


def _run(kind, store, B, ksize):
    key = (kind, store.flat.data_ptr(), B, ksize)
    if key not in _RUNS:
        dev = store.flat.device
        if kind == "d":
            _RUNS[key] = E.DiscriminatorRun(store, B, dev)
        else:
            _RUNS[key] = E.GeneratorRun(store, B, dev, kind == "g_dna", ksize)
        store.refresh_packs()
    return _RUNS[key]


@custom_op("acg::generator_transform", mutates_args=(), device_types="cuda")
def generator_transform(images: torch.Tensor, actions: torch.Tensor, flat: torch.Tensor, ksize: int) -> list[torch.Tensor]:
    store = _store_of(flat)
    run = _run("g_dna", store, images.shape[0], ksize)
    store.refresh_packs()
    out, state = run.forward(_c(images.float()), _c(actions.float()))
    return [out.clone(), state.clone()]


@generator_transform.register_fake
def _(images, actions, flat, ksize):
    return [torch.empty_like(images, dtype=torch.float32), images.new_empty(images.shape[0], E.STATE_DIM)]


@custom_op("acg::generator_transform_grad", mutates_args=(), device_types="cuda")
def generator_transform_grad(d_out: torch.Tensor, d_state: torch.Tensor, flat: torch.Tensor, B: int, ksize: int) -> torch.Tensor:
    """gradient w.r.t. the flat parameter buffer of the LAST forward of this (store, batch, ksize)"""
    store = _store_of(flat)
    run = _run("g_dna", store, B, ksize)
    run.dg_out.copy_(d_out)
    run.dstate.copy_(d_state)
    store.grad.zero_()
    run.backward(with_state=True)
    return store.grad.clone()


@generator_transform_grad.register_fake
def _(d_out, d_state, flat, B, ksize):
    return torch.empty_like(flat)


def _gt_setup(ctx, inputs, output):
    images, actions, flat, ksize = inputs
    ctx.save_for_backward(flat)
    ctx.args = (images.shape[0], ksize)


def _gt_bwd(ctx, grads):
    (flat,) = ctx.saved_tensors
    d_out, d_state = grads
    B, ksize = ctx.args
    d_out = torch.zeros(B, E.IMG, E.IMG, 3, device=flat.device) if d_out is None else _c(d_out.float())
    d_state = torch.zeros(B, E.STATE_DIM, device=flat.device) if d_state is None else _c(d_state.float())
    return None, None, generator_transform_grad(d_out, d_state, flat, B, ksize), None


generator_transform.register_autograd(_gt_bwd, setup_context=_gt_setup)


@custom_op("acg::generator", mutates_args=(), device_types="cuda")
def generator(images: torch.Tensor, actions: torch.Tensor, flat: torch.Tensor) -> torch.Tensor:
    store = _store_of(flat)
    run = _run("g_direct", store, images.shape[0], 5)
    store.refresh_packs()
    out, _ = run.forward(_c(images.float()), _c(actions.float()))
    return out.clone()


@generator.register_fake
def _(images, actions, flat):
    return torch.empty_like(images, dtype=torch.float32)


@custom_op("acg::generator_grad", mutates_args=(), device_types="cuda")
def generator_grad(d_out: torch.Tensor, flat: torch.Tensor, B: int) -> torch.Tensor:
    store = _store_of(flat)
    run = _run("g_direct", store, B, 5)
    run.dg_out.copy_(d_out)
    store.grad.zero_()
    run.backward(with_state=False)
    return store.grad.clone()


@generator_grad.register_fake
def _(d_out, flat, B):
    return torch.empty_like(flat)


def _g_setup(ctx, inputs, output):
    images, actions, flat = inputs
    ctx.save_for_backward(flat)
    ctx.B = images.shape[0]


def _g_bwd(ctx, d_out):
    (flat,) = ctx.saved_tensors
    return None, None, generator_grad(_c(d_out.float()), flat, ctx.B)


generator.register_autograd(_g_bwd, setup_context=_g_setup)


@custom_op("acg::discriminator", mutates_args=(), device_types="cuda")
def discriminator(inputs: torch.Tensor, actions: torch.Tensor, flat: torch.Tensor) -> torch.Tensor:
    store = _store_of(flat)
    run = _run("d", store, inputs.shape[0], 0)
    store.refresh_packs()
    inputs = inputs.float()
    return run.forward(_c(inputs[..., :3]), _c(inputs[..., 3:]), _c(actions.float())).clone()


@discriminator.register_fake
def _(inputs, actions, flat):
    return inputs.new_empty(inputs.shape[0], 2, 2, 1, dtype=torch.float32)


@custom_op("acg::discriminator_grad", mutates_args=(), device_types="cuda")
def discriminator_grad(d_logits: torch.Tensor, flat: torch.Tensor, B: int) -> list[torch.Tensor]:
    """-> [d flat parameters, d inputs [B,64,64,6]] of the LAST forward of this (store, batch)"""
    store = _store_of(flat)
    run = _run("d", store, B, 0)
    run.dlogits.copy_(d_logits.reshape(-1))
    store.grad.zero_()
    dx = run.backward(need_dw=True, need_dinput=True)
    return [store.grad.clone(), _c(dx[..., :6].float())]


@discriminator_grad.register_fake
def _(d_logits, flat, B):
    return [torch.empty_like(flat), flat.new_empty(B, E.IMG, E.IMG, 6)]


def _d_setup(ctx, inputs, output):
    x, actions, flat = inputs
    ctx.save_for_backward(flat)
    ctx.B = x.shape[0]


def _d_bwd(ctx, d_logits):
    (flat,) = ctx.saved_tensors
    dflat, dx = discriminator_grad(_c(d_logits.float()), flat, ctx.B)
    return dx, None, dflat


discriminator.register_autograd(_d_bwd, setup_context=_d_setup)
