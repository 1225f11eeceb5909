"""CPU ORACLE (test infrastructure, NOT product code) -- PyTorch-CPU functional restatement.

PARITY UNPINNED (see oracle/np_ref.py header: no reference tests / golden vectors exist and
TF 1.0 cannot run here).  This file is written independently of np_ref.py (torch conv /
conv_transpose / unfold primitives instead of explicit loops) so that the two can pin each
other; autograd supplies every backward pass.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this module.

Restates: models.py:8-88, ops.py:19-50,100-120, train.py:28-176 (class Trainer) with the
repairs R1-R6 and the update-then-clip determinisation listed in np_ref.py.
"""
import math

import numpy as np
import torch
import torch.nn.functional as F

BN_EPS = 1e-3
L2_WEIGHT = 0.05      # train.py:22
ADAM_LR = 1e-3        # train.py:20
RMSPROP_LR = 5e-5     # train.py:93


def _same(n, k, s):
    out = -(-n // s)
    tot = max((out - 1) * s + k - n, 0)
    return tot // 2, tot - tot // 2


def conv2d(x, w, stride, padding="SAME"):
    """NHWC in/out; w HWIO.  slim.conv2d (models.py:12-15 etc.)."""
    xn = x.permute(0, 3, 1, 2)
    if padding == "SAME":
        pt, pb = _same(x.shape[1], w.shape[0], stride)
        pl, pr = _same(x.shape[2], w.shape[1], stride)
        xn = F.pad(xn, (pl, pr, pt, pb))
    y = F.conv2d(xn.contiguous(), w.permute(3, 2, 0, 1).contiguous(), stride=stride)
    return y.permute(0, 2, 3, 1)


def conv2d_transpose(x, w, stride=2):
    """NHWC; w [kh,kw,Cout,Cin].  slim.conv2d_transpose SAME (models.py:17-21 etc.):
    adjoint of the SAME conv => torch conv_transpose2d(padding=pad_before), cropped."""
    k = w.shape[0]
    H, W = x.shape[1], x.shape[2]
    pt, _ = _same(H * stride, k, stride)
    y = F.conv_transpose2d(x.permute(0, 3, 1, 2).contiguous(), w.permute(3, 2, 0, 1).contiguous(), stride=stride,
                           padding=pt)
    need = stride * H
    if y.shape[2] < need:
        y = F.pad(y, (0, need - y.shape[3], 0, need - y.shape[2]))
    return y[:, :, :need, :stride * W].permute(0, 2, 3, 1)


def batch_norm(x, beta):
    mu = x.mean(dim=(0, 1, 2))
    var = x.var(dim=(0, 1, 2), unbiased=False)
    return (x - mu) * torch.rsqrt(var + BN_EPS) + beta


def lrelu(x, leak=0.2):
    """ops.py:22-26"""
    return 0.5 * (1 + leak) * x + 0.5 * (1 - leak) * x.abs()


# Test hook.  relu / lrelu are piecewise linear; an element whose pre-activation is within fp32 rounding of 0 can
# take the other branch on the device, and with batch-norm over a small batch ONE such element moves a whole
# channel's gradient by ~1/rows.  To compare gradients on the SAME linear piece, a test may set
# GATES = {"<layer>#<k-th application>": bool array (pre-activation > 0 as the device saw it)}; the oracle then
# evaluates act(u) = u*gate (relu) or u*(0.2+0.8*gate) (lrelu) for those layers.  Unset (None) in normal use.
GATES = None
_gate_calls = {}


def reset_gate_calls():
    _gate_calls.clear()


def _apply_act(name, y, act):
    if act is None:
        return y
    if GATES is None or act not in (torch.relu, lrelu):
        return act(y)
    idx = _gate_calls.get(name, 0)
    _gate_calls[name] = idx + 1
    key = "%s#%d" % (name, idx)
    if key not in GATES:
        return act(y)
    g = torch.as_tensor(np.asarray(GATES[key]), dtype=y.dtype).reshape(y.shape)
    return y * g if act is torch.relu else y * (0.2 + 0.8 * g)


def _layer(p, name, x, kind, stride=2, padding="SAME", act=torch.relu):
    w = p[name + "/weights"]
    y = conv2d(x, w, stride, padding) if kind == "conv" else conv2d_transpose(x, w, stride)
    if name + "/biases" in p:
        y = y + p[name + "/biases"]
    if name + "/BatchNorm/beta" in p:
        y = batch_norm(y, p[name + "/BatchNorm/beta"])
    return _apply_act(name, y, act)


def tile_actions(actions, size):
    B = actions.shape[0]
    return actions.reshape(B, 1, 1, -1).expand(B, size, size, actions.shape[1])


def dna_transform(logits, img, k):
    """models.py:60-72 via F.unfold (which is channel-major, so re-ordered to (p, c))."""
    B, H, W, C = img.shape
    s = torch.softmax(logits, dim=-1)
    pt, pb = _same(H, k, 1)
    xn = F.pad(img.permute(0, 3, 1, 2), (pt, pb, pt, pb))
    cols = F.unfold(xn, kernel_size=k)                       # [B, C*k*k, H*W], c major
    cols = cols.reshape(B, C, k * k, H, W).permute(0, 3, 4, 2, 1)   # [B,H,W,k*k,C]
    return (s.unsqueeze(-1) * cols).sum(dim=3)


def generator_transform(p, images, actions, ksize):
    out = _layer(p, "g/conv1", images, "conv")
    out = _layer(p, "g/conv2", out, "conv")
    out = _layer(p, "g/conv3", out, "conv")
    out = _layer(p, "g/conv4", out, "conv")
    out = torch.cat([out, tile_actions(actions, 4)], dim=3)
    out = _layer(p, "g/tconv1", out, "deconv")
    out = _layer(p, "g/tconv2", out, "deconv")
    st = _layer(p, "g/sconv3", out, "conv")
    st = _layer(p, "g/sconv4", st, "conv")
    st = _layer(p, "g/sconv5", st, "conv", stride=1, padding="VALID", act=None)
    out = _layer(p, "g/tconv3", out, "deconv")
    logits = _layer(p, "g/tconv4", out, "deconv", act=None)
    return dna_transform(logits, images, ksize), st.reshape(st.shape[0], -1), logits


def generator_direct(p, images, actions):
    out = _layer(p, "g/conv1", images, "conv")
    out = _layer(p, "g/conv2", out, "conv")
    out = _layer(p, "g/conv3", out, "conv")
    out = _layer(p, "g/conv4", out, "conv")
    out = torch.cat([out, tile_actions(actions, 4)], dim=3)
    out = _layer(p, "g/tconv1", out, "deconv")
    out = _layer(p, "g/tconv2", out, "deconv")
    out = _layer(p, "g/tconv3", out, "deconv")
    return _layer(p, "g/tconv4", out, "deconv", act=torch.tanh)


def discriminator(p, inputs, actions):
    out = _layer(p, "d/conv1", inputs, "conv", act=lrelu)
    out = _layer(p, "d/conv2", out, "conv", act=lrelu)
    out = torch.cat([out, tile_actions(actions, 16)], dim=3)
    out = _layer(p, "d/conv3", out, "conv", act=lrelu)
    out = _layer(p, "d/conv4", out, "conv", act=lrelu)
    out = _layer(p, "d/conv5", out, "conv", act=lrelu)
    return _layer(p, "d/conv6", out, "conv", stride=1, act=None)


def sigmoid_ce(label, x):
    return (x.clamp(min=0) - x * label + torch.log1p(torch.exp(-x.abs()))).mean()


def build_psnr(true, pred):
    return 10.0 * torch.log(1.0 / ((true - pred) ** 2).mean()) / math.log(10.0)


def build_g_adv_loss(d_out_gen, arg_loss):
    if arg_loss == "bce":
        return sigmoid_ce(1.0, d_out_gen)
    elif arg_loss == "wass":
        return d_out_gen.mean()
    raise ValueError("unexpected loss argument")


def build_d_loss(d_out_direct, d_out_gen, arg_loss):
    if arg_loss == "bce":
        direct = sigmoid_ce(0.9, d_out_direct)
        gen = sigmoid_ce(0.0, d_out_gen)
    elif arg_loss == "wass":
        direct = d_out_direct.mean()
        gen = -d_out_gen.mean()
    else:
        raise ValueError("unexpected loss argument")
    return direct + gen, direct, gen


def build_gdl(a, b, alpha=1):
    """ops.py:100-120 through real convs with the +/- identity filters."""
    C = a.shape[3]
    eye = torch.eye(C, dtype=a.dtype)
    fx = torch.stack([-eye, eye]).unsqueeze(0)          # [1,2,C,C] HWIO
    fy = torch.stack([eye.unsqueeze(0), -eye.unsqueeze(0)])   # [2,1,C,C]
    adx, ady = conv2d(a, fx, 1).abs(), conv2d(a, fy, 1).abs()
    bdx, bdy = conv2d(b, fx, 1).abs(), conv2d(b, fy, 1).abs()
    return ((bdx - adx).abs() ** alpha + (bdy - ady).abs() ** alpha).sum()


class TFAdam:
    """tf.train.AdamOptimizer (TF 1.0)."""
    def __init__(self, params, lr=ADAM_LR, b1=0.9, b2=0.999, eps=1e-8):
        self.lr, self.b1, self.b2, self.eps, self.t = lr, b1, b2, eps, 0
        self.m = {k: torch.zeros_like(v) for k, v in params.items()}
        self.v = {k: torch.zeros_like(v) for k, v in params.items()}

    def step(self, params, grads):
        self.t += 1
        lr_t = self.lr * math.sqrt(1 - self.b2 ** self.t) / (1 - self.b1 ** self.t)
        for k in params:
            g = grads[k]
            self.m[k] = self.b1 * self.m[k] + (1 - self.b1) * g
            self.v[k] = self.b2 * self.v[k] + (1 - self.b2) * g * g
            params[k] = params[k] - lr_t * self.m[k] / (self.v[k].sqrt() + self.eps)


class TFRMSProp:
    """tf.train.RMSPropOptimizer (TF 1.0): ms initialised to ones, eps inside the sqrt."""
    def __init__(self, params, lr=RMSPROP_LR, decay=0.9, eps=1e-10):
        self.lr, self.decay, self.eps = lr, decay, eps
        self.ms = {k: torch.ones_like(v) for k, v in params.items()}

    def step(self, params, grads):
        for k in params:
            g = grads[k]
            self.ms[k] = self.decay * self.ms[k] + (1 - self.decay) * g * g
            params[k] = params[k] - self.lr * g / torch.sqrt(self.ms[k] + self.eps)


class Trainer:
    """train.py:27-176.  `params` maps TF variable names -> numpy arrays (np_ref.init_params)."""

    def __init__(self, params, arg_adv, arg_loss, arg_opt, arg_transform, ksize=6,
                 dtype=torch.float64, batch_size=None):
        self.dtype = dtype
        self.p = {k: torch.tensor(np.asarray(v), dtype=dtype) for k, v in params.items()}
        self.arg_adv, self.arg_loss, self.arg_transform, self.ksize = arg_adv, arg_loss, arg_transform, ksize
        if arg_loss not in ("bce", "wass"):
            raise ValueError("unexpected loss argument")
        if arg_opt == "rmsprop":
            mk = TFRMSProp
        elif arg_opt == "adam":
            mk = TFAdam
        else:
            raise ValueError("unexpected opt argument")
        gp = {k: v for k, v in self.p.items() if k.startswith("g/")}
        dp = {k: v for k, v in self.p.items() if k.startswith("d/")}
        self.g_opt, self.g_pretrain_opt, self.d_opt = mk(gp), mk(gp), mk(dp)
        self.last = {}

    # -- graph (train.py:48-85) --------------------------------------------------------
    def _t(self, a):
        return torch.tensor(np.asarray(a), dtype=self.dtype)

    def _forward(self, p, img, nxt, act, state, need_real=True):
        reset_gate_calls()
        B = img.shape[0]
        if self.arg_transform:
            g_out, g_state, logits = generator_transform(p, img, act, self.ksize)
        else:
            g_out, g_state, logits = generator_direct(p, img, act), None, None
        d_gen = discriminator(p, torch.cat([img, g_out], dim=3), act)
        out = {"g_out": g_out, "g_state": g_state, "d_gen": d_gen, "logits": logits}
        out["g_psnr"] = build_psnr(nxt, g_out)
        l2 = (g_out - nxt).abs().sum() / B
        if self.arg_transform:
            l2 = l2 * L2_WEIGHT + torch.sqrt(((g_state - state) ** 2).sum()) / B
        out["g_l2_loss"] = l2
        if self.arg_adv:
            out["g_adv_loss"] = build_g_adv_loss(d_gen, self.arg_loss)
            out["g_loss"] = l2 + out["g_adv_loss"] + build_gdl(nxt, g_out)
        else:
            out["g_loss"] = l2
        if need_real:
            d_real = discriminator(p, torch.cat([img, nxt], dim=3), act)
            out["d_real"] = d_real
            tot, direct, gen = build_d_loss(d_real, d_gen, self.arg_loss)
            out["discriminator_loss"], out["discriminator_direct_loss"], out["discriminator_gen_loss"] = tot, direct, gen
        return out

    def _step(self, img, nxt, act, state, loss_key, scope, opt, need_real):
        p = {k: v.clone().requires_grad_(k.startswith(scope)) for k, v in self.p.items()}
        out = self._forward(p, self._t(img), self._t(nxt), self._t(act), self._t(state), need_real)
        names = [k for k in p if k.startswith(scope)]
        grads = torch.autograd.grad(out[loss_key], [p[k] for k in names], allow_unused=True)
        gd = {k: (g if g is not None else torch.zeros_like(p[k])) for k, g in zip(names, grads)}
        sub = {k: self.p[k] for k in names}
        opt.step(sub, gd)
        self.p.update(sub)
        self.last = {k: (v.detach() if torch.is_tensor(v) else v) for k, v in out.items() if v is not None}
        self.last_grads = gd
        return out

    # -- train.py:114-144 --------------------------------------------------------------
    def pretrain_g(self, img, nxt, act, state):
        out = self._step(img, nxt, act, state, "g_l2_loss", "g/", self.g_pretrain_opt, need_real=False)
        return float(out["g_loss"].detach())

    def train_g(self, img, nxt, act, state):
        out = self._step(img, nxt, act, state, "g_loss", "g/", self.g_opt, need_real=False)
        return out["g_out"].detach().numpy()

    def train_d(self, img, nxt, act, summarize=False):
        state = np.zeros((np.asarray(img).shape[0], 5))
        out = self._step(img, nxt, act, state, "discriminator_loss", "d/", self.d_opt, need_real=True)
        for k in self.p:                       # train.py:89 (update, then clip)
            if k.startswith("d/"):
                self.p[k] = self.p[k].clamp(-0.01, 0.01)
        if summarize:
            return self.summaries()
        return None

    def summaries(self):
        keys = ["discriminator_direct_loss", "discriminator_gen_loss", "discriminator_loss",
                "g_loss", "g_l2_loss", "g_adv_loss", "g_psnr"]
        return {k: float(self.last[k]) for k in keys if k in self.last}

    # -- train.py:146-176 --------------------------------------------------------------
    def test(self, img, nxt, act):
        B = np.asarray(img).shape[0]
        with torch.no_grad():
            out = self._forward(self.p, self._t(img), self._t(nxt), self._t(act),
                                torch.zeros(B, 5, dtype=self.dtype), need_real=True)
        self.last = {k: v for k, v in out.items() if v is not None}
        st = out["g_state"].numpy() if out["g_state"] is not None else None
        return out["g_out"].numpy(), st, self.summaries()

    def test_sequence(self, input_images, test_next_frame, test_actions):
        predicted = []
        current_frame = input_images[:, 0]
        current_state = test_actions[:, 0, 5:]
        for j in range(6):
            acs = np.concatenate((test_actions[:, j * 2, :5], current_state), axis=1)
            out, st, _ = self.test(current_frame, test_next_frame[:, j * 2], acs)
            predicted.append(out)
            current_frame = out
            if st is not None:                 # R5: non-DNA keeps the fed state
                current_state = st
        predicted = np.transpose(np.array(predicted), (1, 0, 2, 3, 4))
        return predicted, current_frame[1:7]

    def numpy_params(self):
        return {k: v.detach().numpy().copy() for k, v in self.p.items()}
