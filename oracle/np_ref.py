"""CPU ORACLE (test infrastructure, NOT product code) -- NumPy restatement.

PARITY UNPINNED: the reference (/root/reference, TensorFlow 1.0 graph code) ships no
tests or golden vectors and TensorFlow 1.0 cannot be installed here, so nothing in this
file could be checked against the reference's own outputs.  It is pinned only against
(a) an independently written PyTorch-CPU restatement (oracle/torch_ref.py) and (b) analytic
identities (tests/test_oracle.py).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this module.  The product package (action_conditioned_gans_b200/) never does.

Every function cites the reference file:line it restates.  Semantics that live inside
TF-1.0 / tf.contrib.slim (not under /root/reference) are written out explicitly:

* conv SAME padding: out=ceil(in/s); pad_total=max((out-1)*s+k-in,0); before=pad_total//2,
  the odd element goes AFTER.  Cross-correlation, NHWC activations, HWIO weights.
* conv2d_transpose SAME stride 2: exact adjoint of that conv, weights [kh,kw,Cout,Cin].
* slim layer order: conv -> (bias only if no normalizer) -> batch_norm -> activation.
* slim.batch_norm defaults: batch statistics always (is_training=True), biased variance,
  epsilon=1e-3, beta only (scale=False), moving averages never updated.
* extract_image_patches depth order (kh, kw, c) with c fastest.
* Adam / RMSProp exactly as TF-1.0 implements them (see adam_step / rmsprop_step).

Repairs made to reference HEAD (which does not run), per SURVEY.md section 0:
  R1 train.py:94 `=` -> `==`; R2 models.py:10,31,80 `argscope` -> `arg_scope`;
  R3 train.py:50 discriminator action map tiled to 16x16 (matches d/conv2 output);
  R4 test.py parses its args; R5 non-DNA rollout carries the fed state forward;
  R6 --adv/--dna accept True|False and default True.
Determinisation: the D step is optimizer update THEN clip (train.py:140,143 leave the
order unspecified).
"""
import numpy as np

BN_EPS = 1e-3          # slim.batch_norm default epsilon
L2_WEIGHT = 0.05       # train.py:22
IMG = 64               # train.py:17-18


# --------------------------------------------------------------------------------------
# padding helpers (TF SAME semantics)
# --------------------------------------------------------------------------------------
def same_pad(n_in, k, s):
    out = -(-n_in // s)
    total = max((out - 1) * s + k - n_in, 0)
    before = total // 2
    return out, before, total - before


# --------------------------------------------------------------------------------------
# slim.conv2d (models.py:12-15,34-37,42-51,82-88)
# --------------------------------------------------------------------------------------
def conv2d(x, w, stride, padding="SAME"):
    """x [B,H,W,Cin], w [kh,kw,Cin,Cout] -> [B,OH,OW,Cout]; cross-correlation."""
    B, H, W, Cin = x.shape
    kh, kw, _, Cout = w.shape
    if padding == "SAME":
        OH, pt, pb = same_pad(H, kh, stride)
        OW, pl, pr = same_pad(W, kw, stride)
    else:
        OH = (H - kh) // stride + 1
        OW = (W - kw) // stride + 1
        pt = pb = pl = pr = 0
    xp = np.zeros((B, H + pt + pb, W + pl + pr, Cin), x.dtype)
    xp[:, pt:pt + H, pl:pl + W] = x
    y = np.zeros((B, OH, OW, Cout), x.dtype)
    for a in range(kh):
        for b in range(kw):
            patch = xp[:, a:a + (OH - 1) * stride + 1:stride, b:b + (OW - 1) * stride + 1:stride]
            y += patch @ w[a, b]
    return y


# --------------------------------------------------------------------------------------
# slim.conv2d_transpose, stride 2, SAME (models.py:17-21,39-40,53-59)
# --------------------------------------------------------------------------------------
def conv2d_transpose(x, w, stride=2):
    """x [B,H,W,Cin], w [kh,kw,Cout,Cin] -> [B,sH,sW,Cout].

    Adjoint of the SAME conv that maps [sH,sW] -> [H,W]:
      y[n, s*i+a-pt, s*j+b-pl, co] += x[n,i,j,ci] * w[a,b,co,ci], clipped to the output.
    """
    B, H, W, Cin = x.shape
    kh, kw, Cout, _ = w.shape
    OH, OW = H * stride, W * stride
    _, pt, _ = same_pad(OH, kh, stride)
    _, pl, _ = same_pad(OW, kw, stride)
    full = np.zeros((B, (H - 1) * stride + kh, (W - 1) * stride + kw, Cout), x.dtype)
    for a in range(kh):
        for b in range(kw):
            full[:, a:a + (H - 1) * stride + 1:stride, b:b + (W - 1) * stride + 1:stride] += \
                x @ w[a, b].T
    need_h, need_w = pt + OH, pl + OW
    if full.shape[1] < need_h or full.shape[2] < need_w:
        grown = np.zeros((B, max(full.shape[1], need_h), max(full.shape[2], need_w), Cout), x.dtype)
        grown[:, :full.shape[1], :full.shape[2]] = full
        full = grown
    return full[:, pt:pt + OH, pl:pl + OW]


# --------------------------------------------------------------------------------------
# slim.batch_norm (models.py:11,32,81), activations (ops.py:22-26)
# --------------------------------------------------------------------------------------
def batch_norm(x, beta):
    mu = x.mean(axis=(0, 1, 2))
    var = x.var(axis=(0, 1, 2))           # biased
    return (x - mu) / np.sqrt(var + BN_EPS) + beta


def relu(x):
    return np.maximum(x, 0)


def lrelu(x, leak=0.2):
    """ops.py:22-26"""
    f1 = 0.5 * (1 + leak)
    f2 = 0.5 * (1 - leak)
    return f1 * x + f2 * np.abs(x)


# --------------------------------------------------------------------------------------
# DNA transform (models.py:60-72)
# --------------------------------------------------------------------------------------
def softmax_last(z):
    z = z - z.max(axis=-1, keepdims=True)
    e = np.exp(z)
    return e / e.sum(axis=-1, keepdims=True)


def extract_patches(img, k):
    """tf.extract_image_patches(ksizes=k, strides=1, SAME) reshaped [B,H,W,k*k,C]
    (models.py:62-68): patch index p = a*k+b, channel minor."""
    B, H, W, C = img.shape
    _, pt, pb = same_pad(H, k, 1)
    _, pl, pr = same_pad(W, k, 1)
    xp = np.zeros((B, H + pt + pb, W + pl + pr, C), img.dtype)
    xp[:, pt:pt + H, pl:pl + W] = img
    out = np.zeros((B, H, W, k * k, C), img.dtype)
    for a in range(k):
        for b in range(k):
            out[:, :, :, a * k + b, :] = xp[:, a:a + H, b:b + W, :]
    return out


def dna_forward(logits, img, k):
    """models.py:60-72.  logits [B,H,W,k*k], img [B,H,W,C] -> [B,H,W,C]."""
    s = softmax_last(logits)
    P = extract_patches(img, k)
    return (s[..., None] * P).sum(axis=3)


def dna_backward(logits, img, dy, k):
    """Autodiff of dna_forward w.r.t. logits (img is a placeholder, train.py:31-34):
    g_p = sum_c dy_c * x_{p,c};  dz_p = s_p * (g_p - sum_q s_q g_q)."""
    s = softmax_last(logits)
    P = extract_patches(img, k)
    g = (P * dy[:, :, :, None, :]).sum(axis=4)
    return s * (g - (s * g).sum(axis=-1, keepdims=True))


def dna_forward_loops(logits, img, k):
    """Same as dna_forward with explicit per-pixel loops (tiny sizes only)."""
    B, H, W, C = img.shape
    pb = (k - 1) // 2
    out = np.zeros((B, H, W, C), logits.dtype)
    for n in range(B):
        for i in range(H):
            for j in range(W):
                z = logits[n, i, j]
                e = np.exp(z - z.max())
                s = e / e.sum()
                for a in range(k):
                    for b in range(k):
                        ii, jj = i + a - pb, j + b - pb
                        if 0 <= ii < H and 0 <= jj < W:
                            out[n, i, j] += s[a * k + b] * img[n, ii, jj]
    return out


# --------------------------------------------------------------------------------------
# parameters: TF variable names, HWIO layouts, xavier-uniform init (slim defaults)
# --------------------------------------------------------------------------------------
def g_dna_spec(ksize):
    """(name, kind, k, cin, cout, bn, bias) for build_generator_transform (models.py:24-74)."""
    return [
        ("g/conv1", "conv", 5, 3, 32, True, False),
        ("g/conv2", "conv", 5, 32, 64, True, False),
        ("g/conv3", "conv", 5, 64, 128, True, False),
        ("g/conv4", "conv", 5, 128, 256, True, False),
        ("g/tconv1", "deconv", 5, 266, 128, True, False),
        ("g/tconv2", "deconv", 5, 128, 128, True, False),
        ("g/sconv3", "conv", 3, 128, 32, True, False),
        ("g/sconv4", "conv", 3, 32, 16, True, False),
        ("g/sconv5", "conv", 4, 16, 5, False, True),
        ("g/tconv3", "deconv", 5, 128, 128, True, False),
        ("g/tconv4", "deconv", 5, 128, ksize * ksize, False, True),
    ]


def g_direct_spec():
    """build_generator (models.py:8-22)."""
    return [
        ("g/conv1", "conv", 5, 3, 64, True, False),
        ("g/conv2", "conv", 5, 64, 128, True, False),
        ("g/conv3", "conv", 5, 128, 256, True, False),
        ("g/conv4", "conv", 5, 256, 512, True, False),
        ("g/tconv1", "deconv", 5, 522, 256, True, False),
        ("g/tconv2", "deconv", 5, 256, 128, True, False),
        ("g/tconv3", "deconv", 5, 128, 64, True, False),
        ("g/tconv4", "deconv", 5, 64, 3, False, True),
    ]


def d_spec():
    """build_discriminator (models.py:76-88); conv6 keeps the arg_scope's batch_norm."""
    return [
        ("d/conv1", "conv", 5, 6, 64, True, False),
        ("d/conv2", "conv", 5, 64, 128, True, False),
        ("d/conv3", "conv", 5, 138, 128, True, False),
        ("d/conv4", "conv", 5, 128, 256, True, False),
        ("d/conv5", "conv", 5, 256, 512, True, False),
        ("d/conv6", "conv", 2, 512, 1, True, False),
    ]


def init_params(spec, rng, dtype=np.float32):
    """xavier_initializer() uniform, zero biases / betas (slim defaults)."""
    p = {}
    for name, kind, k, cin, cout, bn, bias in spec:
        shape = (k, k, cin, cout) if kind == "conv" else (k, k, cout, cin)
        fan_in, fan_out = k * k * shape[2], k * k * shape[3]
        lim = np.sqrt(6.0 / (fan_in + fan_out))
        p[name + "/weights"] = rng.uniform(-lim, lim, size=shape).astype(dtype)
        if bn:
            p[name + "/BatchNorm/beta"] = np.zeros(cout, dtype)
        if bias:
            p[name + "/biases"] = np.zeros(cout, dtype)
    return p


# --------------------------------------------------------------------------------------
# networks
# --------------------------------------------------------------------------------------
def _layer(p, name, x, kind, stride=2, padding="SAME", act=relu):
    w = p[name + "/weights"]
    y = conv2d(x, w, stride, padding) if kind == "conv" else conv2d_transpose(x, w, stride)
    if name + "/biases" in p:
        y = y + p[name + "/biases"]
    if name + "/BatchNorm/beta" in p:
        y = batch_norm(y, p[name + "/BatchNorm/beta"])
    return act(y) if act is not None else y


def tile_actions(actions, size):
    """train.py:48-50 (R3: the discriminator map is 16x16)."""
    B = actions.shape[0]
    return np.broadcast_to(actions.reshape(B, 1, 1, -1), (B, size, size, actions.shape[1])).copy()


def generator_transform(p, images, actions, ksize):
    """build_generator_transform (models.py:24-74) -> (frame, state[B,5], logits)."""
    out = _layer(p, "g/conv1", images, "conv")
    out = _layer(p, "g/conv2", out, "conv")
    out = _layer(p, "g/conv3", out, "conv")
    out = _layer(p, "g/conv4", out, "conv")
    out = np.concatenate([out, tile_actions(actions, 4)], axis=3)
    out = _layer(p, "g/tconv1", out, "deconv")
    out = _layer(p, "g/tconv2", out, "deconv")
    st = _layer(p, "g/sconv3", out, "conv")
    st = _layer(p, "g/sconv4", st, "conv")
    st = _layer(p, "g/sconv5", st, "conv", stride=1, padding="VALID", act=None)
    out = _layer(p, "g/tconv3", out, "deconv")
    logits = _layer(p, "g/tconv4", out, "deconv", act=None)
    frame = dna_forward(logits, images, ksize)
    return frame, st.reshape(st.shape[0], -1), logits


def generator_direct(p, images, actions):
    """build_generator (models.py:8-22)."""
    out = _layer(p, "g/conv1", images, "conv")
    out = _layer(p, "g/conv2", out, "conv")
    out = _layer(p, "g/conv3", out, "conv")
    out = _layer(p, "g/conv4", out, "conv")
    out = np.concatenate([out, tile_actions(actions, 4)], axis=3)
    out = _layer(p, "g/tconv1", out, "deconv")
    out = _layer(p, "g/tconv2", out, "deconv")
    out = _layer(p, "g/tconv3", out, "deconv")
    return _layer(p, "g/tconv4", out, "deconv", act=np.tanh)


def discriminator(p, inputs, actions):
    """build_discriminator (models.py:76-88); inputs [B,64,64,6] -> logits [B,2,2,1]."""
    out = _layer(p, "d/conv1", inputs, "conv", act=lrelu)
    out = _layer(p, "d/conv2", out, "conv", act=lrelu)
    out = np.concatenate([out, tile_actions(actions, 16)], axis=3)
    out = _layer(p, "d/conv3", out, "conv", act=lrelu)
    out = _layer(p, "d/conv4", out, "conv", act=lrelu)
    out = _layer(p, "d/conv5", out, "conv", act=lrelu)
    return _layer(p, "d/conv6", out, "conv", stride=1, act=None)


# --------------------------------------------------------------------------------------
# losses (ops.py:19-50, 100-120; train.py:72-85)
# --------------------------------------------------------------------------------------
def sigmoid_cross_entropy(labels, logits):
    """tf.losses.sigmoid_cross_entropy: mean of max(x,0) - x*z + log1p(exp(-|x|))."""
    x = logits
    return np.mean(np.maximum(x, 0) - x * labels + np.log1p(np.exp(-np.abs(x))))


def psnr(true, pred):
    """ops.py:19-20"""
    return 10.0 * np.log(1.0 / np.mean((true - pred) ** 2)) / np.log(10.0)


def g_adv_loss(d_out_gen, arg_loss):
    """ops.py:28-35"""
    if arg_loss == "bce":
        return sigmoid_cross_entropy(np.ones_like(d_out_gen), d_out_gen)
    elif arg_loss == "wass":
        return np.mean(d_out_gen)
    raise ValueError("unexpected loss argument")


def d_loss(d_out_direct, d_out_gen, arg_loss):
    """ops.py:37-50 -> (total, direct, gen)"""
    if arg_loss == "bce":
        direct = sigmoid_cross_entropy(0.9 * np.ones_like(d_out_direct), d_out_direct)
        gen = sigmoid_cross_entropy(np.zeros_like(d_out_gen), d_out_gen)
    elif arg_loss == "wass":
        direct = np.mean(d_out_direct)
        gen = -np.mean(d_out_gen)
    else:
        raise ValueError("unexpected loss argument")
    return direct + gen, direct, gen


def gdl(a, b, alpha=1):
    """ops.py:100-120: 1x2 / 2x1 identity-channel difference filters, SAME (pad after)."""
    def dxdy(x):
        dx = -x.copy()
        dx[:, :, :-1] += x[:, :, 1:]
        dy = x.copy()
        dy[:, :-1] -= x[:, 1:]
        return np.abs(dx), np.abs(dy)
    adx, ady = dxdy(a)
    bdx, bdy = dxdy(b)
    return np.sum(np.abs(bdx - adx) ** alpha + np.abs(bdy - ady) ** alpha)


def generator_losses(g_out, next_frame, d_out_gen, batch, arg_adv, arg_loss,
                     state_out=None, next_state=None):
    """train.py:72-83 -> dict(g_l2_loss, g_adv_loss, g_loss, g_psnr)."""
    l2 = np.sum(np.abs(g_out - next_frame)) / batch
    if state_out is not None:
        l2 = l2 * L2_WEIGHT + np.sqrt(np.sum((state_out - next_state) ** 2)) / batch
    res = {"g_l2_loss": l2, "g_psnr": psnr(next_frame, g_out)}
    if arg_adv:
        res["g_adv_loss"] = g_adv_loss(d_out_gen, arg_loss)
        res["g_loss"] = l2 + res["g_adv_loss"] + gdl(next_frame, g_out)
    else:
        res["g_loss"] = l2
    return res


# --------------------------------------------------------------------------------------
# optimizers (train.py:91-102), TF-1.0 formulas; clip (train.py:89)
# --------------------------------------------------------------------------------------
def adam_step(p, g, m, v, t, lr=1e-3, b1=0.9, b2=0.999, eps=1e-8):
    """tf.train.AdamOptimizer: epsilon is NOT bias corrected. t is 1-based."""
    lr_t = lr * np.sqrt(1 - b2 ** t) / (1 - b1 ** t)
    m = b1 * m + (1 - b1) * g
    v = b2 * v + (1 - b2) * g * g
    return p - lr_t * m / (np.sqrt(v) + eps), m, v


def rmsprop_step(p, g, ms, lr=5e-5, decay=0.9, eps=1e-10):
    """tf.train.RMSPropOptimizer (momentum 0, uncentered): ms starts at ONE, eps inside sqrt."""
    ms = decay * ms + (1 - decay) * g * g
    return p - lr * g / np.sqrt(ms + eps), ms


def clip(p, lo=-0.01, hi=0.01):
    return np.clip(p, lo, hi)
