"""Step engine: layer tables, flat parameter stores and the hand-scheduled forward/backward of the
generator (DNA and direct-pixel) and the discriminator on top of the acg_b200 kernels.

This is the host-side replacement of the TF graph that `Trainer.__init__` builds in the reference
(train.py:28-112) plus TF autodiff: every tensor op is one of our CUDA kernels (kernels.py); torch tensors are
device storage only.  Layer tables restate models.py:8-88 with the slim/TF-1.0 defaults written out
(SURVEY.md section 8(c)): conv -> (bias only if no normalizer) -> batch_norm(beta only, batch statistics,
eps 1e-3) -> activation; SAME padding puts the odd element after; conv2d_transpose is the exact adjoint.
"""
import math
from dataclasses import dataclass

import numpy as np
import torch

from . import kernels as K

IMG = 64            # train.py:17-18
ACTION_DIM = 10     # train.py:39-42 (5-D action ++ 5-D state)
STATE_DIM = 5       # train.py:43-46
BN_EPS = 1e-3       # slim.batch_norm default


@dataclass(frozen=True)
class LayerSpec:
    name: str
    kind: str       # 'conv' | 'deconv'
    k: int
    stride: int
    padding: str    # 'SAME' | 'VALID'
    cin: int
    cout: int
    bn: bool
    bias: bool
    act: str        # 'relu' | 'lrelu' | 'tanh' | 'none'


def g_dna_spec(ksize):
    """build_generator_transform, models.py:24-74."""
    c = lambda n, k, ci, co: LayerSpec(n, "conv", k, 2, "SAME", ci, co, True, False, "relu")
    t = lambda n, ci, co: LayerSpec(n, "deconv", 5, 2, "SAME", ci, co, True, False, "relu")
    return [
        c("g/conv1", 5, 3, 32), c("g/conv2", 5, 32, 64), c("g/conv3", 5, 64, 128), c("g/conv4", 5, 128, 256),
        t("g/tconv1", 256 + ACTION_DIM, 128), t("g/tconv2", 128, 128),
        c("g/sconv3", 3, 128, 32), c("g/sconv4", 3, 32, 16),
        LayerSpec("g/sconv5", "conv", 4, 1, "VALID", 16, STATE_DIM, False, True, "none"),
        t("g/tconv3", 128, 128),
        LayerSpec("g/tconv4", "deconv", 5, 2, "SAME", 128, ksize * ksize, False, True, "none"),
    ]


def g_direct_spec():
    """build_generator, models.py:8-22."""
    c = lambda n, ci, co: LayerSpec(n, "conv", 5, 2, "SAME", ci, co, True, False, "relu")
    t = lambda n, ci, co: LayerSpec(n, "deconv", 5, 2, "SAME", ci, co, True, False, "relu")
    return [
        c("g/conv1", 3, 64), c("g/conv2", 64, 128), c("g/conv3", 128, 256), c("g/conv4", 256, 512),
        t("g/tconv1", 512 + ACTION_DIM, 256), t("g/tconv2", 256, 128), t("g/tconv3", 128, 64),
        LayerSpec("g/tconv4", "deconv", 5, 2, "SAME", 64, 3, False, True, "tanh"),
    ]


def d_spec():
    """build_discriminator, models.py:76-88 (conv6 keeps the arg_scope's batch_norm and has no bias)."""
    c = lambda n, ci, co: LayerSpec(n, "conv", 5, 2, "SAME", ci, co, True, False, "lrelu")
    return [
        c("d/conv1", 6, 64), c("d/conv2", 64, 128), c("d/conv3", 128 + ACTION_DIM, 128),
        c("d/conv4", 128, 256), c("d/conv5", 256, 512),
        LayerSpec("d/conv6", "conv", 2, 1, "SAME", 512, 1, True, False, "none"),
    ]


def weight_shape(L):
    """TF variable shape: conv HWIO [k,k,cin,cout]; conv2d_transpose [k,k,cout,cin]."""
    return (L.k, L.k, L.cin, L.cout) if L.kind == "conv" else (L.k, L.k, L.cout, L.cin)


def variable_list(spec):
    """[(tf_variable_name, shape)] in creation order (weights, then beta or biases)."""
    out = []
    for L in spec:
        out.append((L.name + "/weights", weight_shape(L)))
        if L.bn:
            out.append((L.name + "/BatchNorm/beta", (L.cout,)))
        if L.bias:
            out.append((L.name + "/biases", (L.cout,)))
    return out


def xavier_init(spec, rng):
    """slim defaults: xavier_initializer() uniform for weights, zeros for beta / biases.  NumPy fp32 dict."""
    p = {}
    for name, shape in variable_list(spec):
        if name.endswith("/weights"):
            fan_in, fan_out = shape[0] * shape[1] * shape[2], shape[0] * shape[1] * shape[3]
            lim = math.sqrt(6.0 / (fan_in + fan_out))
            p[name] = rng.uniform(-lim, lim, size=shape).astype(np.float32)
        else:
            p[name] = np.zeros(shape, np.float32)
    return p


class ParamStore:
    """All variables of one scope ('g' or 'd') in ONE flat fp32 buffer (+ a same-layout gradient buffer), so
    that the optimizer, the weight clip and the gradient all-reduce are each a single launch / message."""

    def __init__(self, spec, device, init=None):
        self.spec = spec
        self.device = device
        self.offsets = {}
        off = 0
        for name, shape in variable_list(spec):
            n = int(np.prod(shape))
            self.offsets[name] = (off, n, shape)
            off += (n + 3) // 4 * 4          # keep every variable 16-byte aligned
        self.numel = off
        self.flat = torch.zeros(off, dtype=torch.float32, device=device)
        self.grad = torch.zeros(off, dtype=torch.float32, device=device)
        self.views = {n: self.flat[o:o + k].view(shape) for n, (o, k, shape) in self.offsets.items()}
        self.gviews = {n: self.grad[o:o + k].view(shape) for n, (o, k, shape) in self.offsets.items()}
        if init is not None:
            self.load(init)

    def load(self, arrays):
        for name, (o, k, shape) in self.offsets.items():
            a = np.asarray(arrays[name], dtype=np.float32).reshape(shape)
            self.views[name].copy_(torch.from_numpy(np.ascontiguousarray(a)))

    def numpy(self):
        return {n: v.detach().cpu().numpy().copy() for n, v in self.views.items()}

    def grads_numpy(self):
        return {n: v.detach().cpu().numpy().copy() for n, v in self.gviews.items()}


class _LayerState:
    pass


class NetRun:
    """Activation / gradient buffers of ONE application of a network at a fixed local batch size.
    The discriminator is applied twice per step (generated and real pair, train.py:63-70) with shared
    ParamStore and two NetRuns, which is exactly TF's reuse=True."""

    def __init__(self, store, batch, in_hw, device, dp=None):
        self.store, self.B, self.device, self.dp = store, batch, device, dp
        self.layers = {}
        n_stat = 0
        for L in store.spec:
            n_stat += 4 * L.cout
        self.f64 = torch.zeros(n_stat, dtype=torch.float64, device=device)   # [stats | red] per layer
        soff = 0
        for L in store.spec:
            st = _LayerState()
            st.spec = L
            st.soff = soff
            st.stats = self.f64[soff:soff + 2 * L.cout]
            st.red = self.f64[soff + 2 * L.cout:soff + 4 * L.cout]
            soff += 4 * L.cout
            st.mean = torch.zeros(L.cout, device=device)
            st.rstd = torch.ones(L.cout, device=device)
            st.scale = torch.ones(L.cout, device=device)
            st.shift = torch.zeros(L.cout, device=device)
            self.layers[L.name] = st

    def plan(self, name, h, w):
        """Fix the geometry of layer `name` for an input of h x w pixels; allocate z / dz.  Returns output hw."""
        st = self.layers[name]
        L = st.spec
        B = self.B
        if L.kind == "conv":
            st.shape = K.conv_shape(B, h, w, L.cin, L.cout, L.k, L.stride, L.padding)
            oh, ow = st.shape.OH, st.shape.OW
        else:  # conv2d_transpose: the adjoint of the SAME conv [2h,2w,cout] -> [h,w,cin]
            oh, ow = h * L.stride, w * L.stride
            st.shape = K.conv_shape(B, oh, ow, L.cout, L.cin, L.k, L.stride, L.padding)
            assert st.shape.OH == h and st.shape.OW == w
        st.in_hw, st.out_hw = (h, w), (oh, ow)
        st.rows = B * oh * ow
        st.in_rows = B * h * w
        st.z = torch.empty(B, oh, ow, L.cout, device=self.device)
        st.dz = torch.empty(B, oh, ow, L.cout, device=self.device)
        st.dx = None
        return oh, ow

    def zero_reductions(self):
        self.f64.zero_()

    # -- one layer forward: conv -> (bias | batch-norm) -> activation ------------------------------
    def layer_fwd(self, name, x, out, ld_out):
        st = self.layers[name]
        L = st.spec
        w = self.store.views[name + "/weights"]
        if L.kind == "conv":
            K.conv_fprop_f32(st.shape, x, w, st.z)
        else:
            K.conv_dgrad_f32(st.shape, x, w, st.z)
        st.x = x
        if L.bn:
            K.bn_stats(st.z, st.rows, L.cout, L.cout, 1, st.stats)
            world = 1
            if self.dp is not None:
                world = self.dp.world
                self.dp.allreduce_sum(st.stats)        # SyncBN: statistics over the GLOBAL batch
            K.bn_finalize(st.stats, self.store.views[name + "/BatchNorm/beta"], st.rows * world, L.cout, 1,
                          st.mean, st.rstd, st.scale, st.shift, BN_EPS)
            K.bn_act_fwd(st.z, st.rows, L.cout, L.cout, 1, st.scale, st.shift, L.act, out, ld_out)
        else:
            bias = self.store.views[name + "/biases"] if L.bias else None
            K.bn_act_fwd(st.z, st.rows, L.cout, L.cout, 1, None, bias, L.act, out, ld_out)

    # -- one layer backward ----------------------------------------------------------------------------
    def layer_bwd(self, name, dA, ld_d, dA2=None, need_dx=True, need_dw=True):
        st = self.layers[name]
        L = st.spec
        if L.bn:
            mean, rstd, shift = st.mean, st.rstd, st.shift
        else:
            mean, rstd, shift = None, None, (self.store.views[name + "/biases"] if L.bias else None)
        K.bn_act_bwd_reduce(dA, dA2, ld_d, st.z, L.cout, st.rows, L.cout, 1, mean, rstd, shift, L.act, st.red)
        world = 1
        if self.dp is not None and L.bn:
            world = self.dp.world
            self.dp.allreduce_sum(st.red)
        dpar = None
        if need_dw:
            dpar = self.store.gviews[name + ("/BatchNorm/beta" if L.bn else "/biases")]
        K.bn_act_bwd_apply(dA, dA2, ld_d, st.z, L.cout, st.rows, L.cout, 1, mean, rstd, shift, L.act, L.bn,
                           st.red, st.dz, dpar, norm_rows=st.rows * world, dbeta_scale=1.0 / world)
        w = self.store.views[name + "/weights"]
        if need_dw:
            dw = self.store.gviews[name + "/weights"]
            if L.kind == "conv":
                K.conv_wgrad_f32(st.shape, st.x, st.dz, dw)
            else:
                K.conv_wgrad_f32(st.shape, st.dz, st.x, dw)
        if need_dx:
            if st.dx is None:
                st.dx = torch.empty(self.B, st.in_hw[0], st.in_hw[1], L.cin, device=self.device)
            if L.kind == "conv":
                K.conv_dgrad_f32(st.shape, st.dz, w, st.dx)
            else:
                K.conv_fprop_f32(st.shape, st.dz, w, st.dx)
        return st.dx


class GeneratorRun(NetRun):
    """One application of the generator (DNA: models.py:24-74, direct: models.py:8-22)."""

    def __init__(self, store, batch, device, dna, ksize, dp=None):
        super().__init__(store, batch, IMG, device, dp)
        self.dna, self.ksize = dna, ksize
        B = batch
        h = w = IMG
        self.enc = ["g/conv1", "g/conv2", "g/conv3", "g/conv4"]
        self.act_bufs = {}
        for n in self.enc:
            h, w = self.plan(n, h, w)
        c4 = self.layers["g/conv4"].spec.cout
        self.cat = torch.empty(B, h, w, c4 + ACTION_DIM, device=device)      # models.py:16,38
        self.cat_c = c4
        h1, w1 = self.plan("g/tconv1", h, w)
        h2, w2 = self.plan("g/tconv2", h1, w1)
        if dna:
            hs, ws = self.plan("g/sconv3", h2, w2)
            hs, ws = self.plan("g/sconv4", hs, ws)
            hs, ws = self.plan("g/sconv5", hs, ws)
            assert (hs, ws) == (1, 1)
        h3, w3 = self.plan("g/tconv3", h2, w2)
        h4, w4 = self.plan("g/tconv4", h3, w3)
        assert (h4, w4) == (IMG, IMG)
        for n, st in self.layers.items():
            if n != "g/conv4":
                st.a = torch.empty(B, st.out_hw[0], st.out_hw[1], st.spec.cout, device=device)
        self.g_out = torch.empty(B, IMG, IMG, 3, device=device)
        self.dg_out = torch.empty(B, IMG, IMG, 3, device=device)
        if dna:
            self.dlogits = torch.empty(B, IMG, IMG, ksize * ksize, device=device)
            self.state = self.layers["g/sconv5"].a.view(B, STATE_DIM)
            self.dstate = torch.empty(B, STATE_DIM, device=device)
        else:
            self.state = None

    def forward(self, img, actions):
        self.zero_reductions()
        self.img = img
        Ls = self.layers
        x = img
        for n in self.enc[:-1]:
            self.layer_fwd(n, x, Ls[n].a, Ls[n].spec.cout)
            x = Ls[n].a
        ld = self.cat.shape[3]
        self.layer_fwd("g/conv4", x, self.cat, ld)
        hw = self.cat.shape[1] * self.cat.shape[2]
        K.tile_actions(actions, self.B, hw, self.cat, ld, self.cat_c)            # train.py:48-49
        self.layer_fwd("g/tconv1", self.cat, Ls["g/tconv1"].a, Ls["g/tconv1"].spec.cout)
        self.layer_fwd("g/tconv2", Ls["g/tconv1"].a, Ls["g/tconv2"].a, Ls["g/tconv2"].spec.cout)
        t2 = Ls["g/tconv2"].a
        if self.dna:
            self.layer_fwd("g/sconv3", t2, Ls["g/sconv3"].a, 32)
            self.layer_fwd("g/sconv4", Ls["g/sconv3"].a, Ls["g/sconv4"].a, 16)
            self.layer_fwd("g/sconv5", Ls["g/sconv4"].a, Ls["g/sconv5"].a, STATE_DIM)
        self.layer_fwd("g/tconv3", t2, Ls["g/tconv3"].a, Ls["g/tconv3"].spec.cout)
        if self.dna:
            kk = self.ksize * self.ksize
            self.layer_fwd("g/tconv4", Ls["g/tconv3"].a, Ls["g/tconv4"].a, kk)   # logits (+bias)
            K.dna_fwd(Ls["g/tconv4"].a, img, self.g_out, self.ksize)             # models.py:60-72
        else:
            self.layer_fwd("g/tconv4", Ls["g/tconv3"].a, self.g_out, 3)          # tanh image
        return self.g_out, self.state

    def backward(self, with_state):
        """dg_out (and dstate when with_state) must be filled; accumulates into store.grad."""
        Ls = self.layers
        if self.dna:
            K.dna_bwd(Ls["g/tconv4"].a, self.img, self.dg_out, self.dlogits, self.ksize)
            d = self.layer_bwd("g/tconv4", self.dlogits, self.ksize * self.ksize)
        else:
            d = self.layer_bwd("g/tconv4", self.dg_out, 3)
        d3 = self.layer_bwd("g/tconv3", d, Ls["g/tconv3"].spec.cout)
        d2b = None
        if self.dna and with_state:
            ds = self.layer_bwd("g/sconv5", self.dstate, STATE_DIM)
            ds = self.layer_bwd("g/sconv4", ds, 16)
            d2b = self.layer_bwd("g/sconv3", ds, 32)
        d = self.layer_bwd("g/tconv2", d3, Ls["g/tconv2"].spec.cout, dA2=d2b)
        d = self.layer_bwd("g/tconv1", d, Ls["g/tconv1"].spec.cout)            # [B,4,4,c4+10]
        d = self.layer_bwd("g/conv4", d, self.cat.shape[3])                     # first c4 channels
        d = self.layer_bwd("g/conv3", d, Ls["g/conv3"].spec.cout)
        d = self.layer_bwd("g/conv2", d, Ls["g/conv2"].spec.cout)
        self.layer_bwd("g/conv1", d, Ls["g/conv1"].spec.cout, need_dx=False)


class DiscriminatorRun(NetRun):
    """One application of build_discriminator (models.py:76-88) to concat([img, frame], 3)."""

    def __init__(self, store, batch, device, dp=None):
        super().__init__(store, batch, IMG, device, dp)
        B = batch
        self.d_in = torch.empty(B, IMG, IMG, 6, device=device)                   # train.py:64,68
        h = w = IMG
        h, w = self.plan("d/conv1", h, w)
        h, w = self.plan("d/conv2", h, w)
        self.cat = torch.empty(B, h, w, 128 + ACTION_DIM, device=device)         # models.py:84 (R3: 16x16)
        for n in ["d/conv3", "d/conv4", "d/conv5", "d/conv6"]:
            h, w = self.plan(n, h, w)
        for n, st in self.layers.items():
            if n != "d/conv2":
                st.a = torch.empty(B, st.out_hw[0], st.out_hw[1], st.spec.cout, device=device)
        self.logits = self.layers["d/conv6"].a                                   # [B,2,2,1]
        self.n_logits = self.logits.numel()
        self.dlogits = torch.empty(self.n_logits, device=device)

    def forward(self, img, frame, actions):
        self.zero_reductions()
        rows = self.B * IMG * IMG
        K.copy_channels(img, 3, 0, self.d_in, 6, 0, rows, 3)
        K.copy_channels(frame, 3, 0, self.d_in, 6, 3, rows, 3)
        Ls = self.layers
        self.layer_fwd("d/conv1", self.d_in, Ls["d/conv1"].a, 64)
        ld = self.cat.shape[3]
        self.layer_fwd("d/conv2", Ls["d/conv1"].a, self.cat, ld)
        K.tile_actions(actions, self.B, self.cat.shape[1] * self.cat.shape[2], self.cat, ld, 128)
        x = self.cat
        for n in ["d/conv3", "d/conv4", "d/conv5", "d/conv6"]:
            self.layer_fwd(n, x, Ls[n].a, Ls[n].spec.cout)
            x = Ls[n].a
        return self.logits

    def backward(self, need_dw, need_dinput):
        """dlogits must be filled.  Returns d(d_in) [B,64,64,6] when need_dinput."""
        d = self.layer_bwd("d/conv6", self.dlogits, 1, need_dw=need_dw)
        d = self.layer_bwd("d/conv5", d, 512, need_dw=need_dw)
        d = self.layer_bwd("d/conv4", d, 256, need_dw=need_dw)
        d = self.layer_bwd("d/conv3", d, 128, need_dw=need_dw)                   # [B,16,16,138]
        d = self.layer_bwd("d/conv2", d, self.cat.shape[3], need_dw=need_dw)
        return self.layer_bwd("d/conv1", d, 64, need_dw=need_dw, need_dx=need_dinput)


class TFOptimizer:
    """tf.train.AdamOptimizer / RMSPropOptimizer slots over a ParamStore (train.py:91-102)."""

    def __init__(self, store, kind):
        if kind not in ("adam", "rmsprop"):
            raise ValueError("unexpected opt argument")
        self.kind, self.store, self.t = kind, store, 0
        if kind == "adam":
            self.lr = 1e-3                                                       # train.py:20
            self.m = torch.zeros_like(store.flat)
            self.v = torch.zeros_like(store.flat)
        else:
            self.lr = 5e-5                                                       # train.py:93
            self.ms = torch.ones_like(store.flat)                                # TF initialises ms to ONE

    def step(self, clip=K.NO_CLIP, grad_scale=1.0):
        self.t += 1
        s = self.store
        if self.kind == "adam":
            lr_t = self.lr * math.sqrt(1.0 - 0.999 ** self.t) / (1.0 - 0.9 ** self.t)
            K.adam_step(s.flat, s.grad, self.m, self.v, lr_t, clip=clip, grad_scale=grad_scale)
        else:
            K.rmsprop_step(s.flat, s.grad, self.ms, self.lr, clip=clip, grad_scale=grad_scale)
