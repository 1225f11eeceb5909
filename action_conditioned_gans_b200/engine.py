"""Step engine: layer tables, flat parameter stores and the hand-scheduled forward/backward of the
generator (DNA and direct-pixel) and the discriminator on top of the acg_b200 kernels.

This is the host-side replacement of the TF graph that `Trainer.__init__` builds in the reference
(train.py:28-112) plus TF autodiff: every tensor op is one of our CUDA kernels (kernels.py); torch tensors are
device storage only.  Layer tables restate models.py:8-88 with the slim/TF-1.0 defaults written out
(SURVEY.md section 8(c)): conv -> (bias only if no normalizer) -> batch_norm(beta only, batch statistics,
eps 1e-3) -> activation; SAME padding puts the odd element after; conv2d_transpose is the exact adjoint.
"""
import math
import os
from dataclasses import dataclass

import numpy as np
import torch

from . import kernels as K

IMG = 64            # train.py:17-18
ACTION_DIM = 10     # train.py:39-42 (5-D action ++ 5-D state)
STATE_DIM = 5       # train.py:43-46
BN_EPS = 1e-3       # slim.batch_norm default
# batch-norm backward sums in the epilogue of the producing data-gradient kernel (bf16 path) instead of the separate
# acg_bn_act_bwd_reduce pass.  In the generic kernel the epilogue warps are also the gather producers and the fusion
# costs more than the removed pass saves (B200, B=256: 3.81 ms with every layer fused vs 3.59 ms).  The halo-tile kernel
# has dedicated epilogue warps, but there each thread reads its pixel's 32 bytes of z per 16-column chunk straight from
# global memory -- 32 different cache lines per warp instruction, the same L1 tag-stage cost that bounds its stores -- and
# the data-gradient kernels sit on the critical path: 3.74 ms with the halo producers fused vs 3.54 ms without.  So the
# default is "0" (separate acg_bn_act_bwd_reduce pass, which streams at 4.4-5.8 TB/s); "halo" fuses where the producer is
# the halo kernel, "1" everywhere (both parity-tested, tests/test_conv_tc_gpu.py).
FUSE_BWD_REDUCE = os.environ.get("ACG_FUSE_BWD_REDUCE", "0")
# Fused batch-norm moments through integer limb accumulators (order-independent: bitwise reproducible forward pass).
# "0": fp64 atomics instead (order varies from run to run; kept to measure what reproducibility costs).
DETERMINISTIC = os.environ.get("ACG_DETERMINISTIC", "1") != "0"
# single GPU: batch-norm finalize inside the activation pass instead of the conv launch's last CTA (A/B switch)
RAW_MOMENTS = os.environ.get("ACG_RAW_MOMENTS", "1") != "0"
# Data parallel, peer-memory exchange: the conv kernel's last CTA pushes its batch-norm totals to the peers, waits for
# theirs and finalises over the global batch -- no exchange launch per SyncBN layer.  "0": separate exchange kernel.
DP_FUSED = os.environ.get("ACG_DP_FUSED", "1") != "0"
# First layers (g/conv1, d/conv1: 8-channel frame operands): "1" sends their forward through the halo kernel's pixel-pair
# mode instead of the small-K gather kernel.  Built, parity-tested (tests/test_conv_halo_gpu.py) and measured: no gather,
# but no faster either (B=256: g/conv1 30.9 vs 29.9 us, d/conv1 42.5 vs 36.7 us) -- with 16-channel pair pixels in
# 128-byte-swizzled rows the halo kernel streams an 8 KB weight slice and waits on one mbarrier round trip per K=16 step;
# the knock-out probe (scripts/pair_bench.py probe) shows 20-28 us with MMAs and epilogue switched off.  Default off.
PAIR_FIRST = os.environ.get("ACG_PAIR_FIRST", "0") != "0"


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
        self._pack_table = None
        self.packs = {}      # layer name -> (shape, fwd_which, fwd_ld, fwd_pack, bwd_which, bwd_ld, bwd_pack)
        self.pair_packs = {}  # first layers: layer name -> (shape, pixel-pair forward pack)
        if init is not None:
            self.load(init)

    def refresh_packs(self):
        """bf16 GEMM-ready copies of the fp32 master weights (after init / load and after every optimizer step):
        every pack of the store in ONE launch through a device-side job table."""
        if not self.packs:
            return
        npk = len(self.packs) + len(self.pair_packs)
        if self._pack_table is None or self._pack_table[4] != npk:
            entries = []
            for name, (shape, fw, fld, pf, bw, bld, pb) in self.packs.items():
                w = self.views[name + "/weights"]
                entries.append((shape, w, fw, fld, pf))
                entries.append((shape, w, bw, bld, pb))
            for name, (shape, pp) in self.pair_packs.items():     # first layers: pixel-pair forward pack
                entries.append((shape, self.views[name + "/weights"], 2, 16, pp))
            self._pack_table = K.make_pack_jobs(entries, self.device) + (npk,)
        K.pack_weights_batched(*self._pack_table[:4])

    def load(self, arrays):
        for name, (o, k, shape) in self.offsets.items():
            a = np.asarray(arrays[name], dtype=np.float32).reshape(shape)
            self.views[name].copy_(torch.from_numpy(np.ascontiguousarray(a)))
        self.refresh_packs()

    def numpy(self):
        torch.cuda.synchronize(self.device)      # optimizer steps may still be running on a side stream
        return {n: v.detach().cpu().numpy().copy() for n, v in self.views.items()}

    def grads_numpy(self):
        torch.cuda.synchronize(self.device)
        return {n: v.detach().cpu().numpy().copy() for n, v in self.gviews.items()}


class _LayerState:
    pass


class Branch:
    """A side stream for an independent chain of kernels of a step (`with branch: ...` forks from the current stream,
    `branch.join()` makes the current stream wait for it).  Inside a captured CUDA graph this becomes a parallel
    branch of the graph: most layers of this model launch fewer CTAs than the 148 SMs x 2 hold or are latency bound
    (4x4 / 8x8 feature maps), so independent chains -- D(real) beside G and D(generated), every weight gradient beside
    the data-gradient chain, the state head beside the frame head -- fill each other's idle SMs."""

    enabled = True      # class-wide switch: False runs every chain on the caller's stream (per-kernel timing passes)

    def __init__(self, device):
        self.stream = torch.cuda.Stream(device=device)
        self._ctx = None
        self.dirty = False

    def __enter__(self):
        if not Branch.enabled:
            return self
        self.stream.wait_stream(torch.cuda.current_stream())
        self._ctx = torch.cuda.stream(self.stream)
        self._ctx.__enter__()
        self.dirty = True
        return self

    def __exit__(self, *exc):
        ctx, self._ctx = self._ctx, None
        return ctx.__exit__(*exc) if ctx is not None else False

    def join(self):
        if self.dirty:
            torch.cuda.current_stream().wait_stream(self.stream)
            self.dirty = False


class _NoBranch:
    """Same interface, no fork (used when a chain holds an NCCL collective: one communicator, one stream order)."""
    dirty = False

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        return False

    def join(self):
        pass


def ru16(v):
    return (v + 15) // 16 * 16


FRAME_LD = 8        # channel stride of the bf16 frame operands (3 / 6 real channels): K of the first layers = 25 x 8


class NetRun:
    """Activation / gradient buffers of ONE application of a network at a fixed local batch size.
    The discriminator is applied twice per step (generated and real pair, train.py:63-70) with a shared
    ParamStore and two NetRuns, which is exactly TF's reuse=True.

    precision 'bf16': every convolution runs on the tcgen05 kernels; activations, raw conv outputs and their
      gradients are bf16 NHWC with the channel count rounded up to 16 (pad channels are zero), weights are bf16
      GEMM-ready packs refreshed from the fp32 master copy after every optimizer step.
    precision 'fp32': fp32 SIMT kernels end to end (the tight-tolerance parity mode)."""

    def __init__(self, store, batch, device, dp=None, precision="bf16", branches=True):
        self.store, self.B, self.device, self.dp = store, batch, device, dp
        self.bf16 = precision == "bf16"
        self.adt = torch.bfloat16 if self.bf16 else torch.float32
        self.layers = {}
        # per layer [stats 2C | red 2C | integer limb accumulators of the fused moments 6C]: ONE buffer, zeroed by one
        # launch at the start of every forward pass (zero_reductions)
        n_stat = sum(10 * ru16(L.cout) for L in store.spec)
        self.f64 = torch.zeros(n_stat, dtype=torch.float64, device=device)
        # "last CTA" tickets of the conv kernels, one per layer (layers of one network may run concurrently)
        self.counters = torch.zeros(2 * len(store.spec), dtype=torch.int32, device=device)
        # side streams: weight gradients (no collective inside -> also with data parallelism) and a second chain
        self.branches = bool(branches)
        self.wgrad_branch = Branch(device) if self.branches else _NoBranch()
        self.side_branch = Branch(device) if self.branches and not self._chain_has_nccl() else _NoBranch()
        soff = 0
        for li, L in enumerate(store.spec):
            st = _LayerState()
            st.spec = L
            st.counter = self.counters[li:li + 1]
            st.counter_b = self.counters[len(store.spec) + li:len(store.spec) + li + 1]     # backward reduction pass
            if dp is not None and getattr(dp, "peer_sync", False) and L.bn:
                # exchange slots of this layer's forward moments / backward reduction terms (same order on every rank)
                st.slot_f = dp.mailbox.new_slot(2 * L.cout)
                st.slot_b = dp.mailbox.new_slot(2 * L.cout)
            cp = ru16(L.cout)
            st.stats = self.f64[soff:soff + 2 * cp]
            st.red = self.f64[soff + 2 * cp:soff + 4 * cp]
            st.fix_view = self.f64[soff + 4 * cp:soff + 10 * cp].view(torch.int64)
            soff += 10 * cp
            st.mean = torch.zeros(L.cout, device=device)
            st.rstd = torch.ones(L.cout, device=device)
            st.scale = torch.ones(L.cout, device=device)
            st.shift = torch.zeros(L.cout, device=device)
            self.layers[L.name] = st

    def _chain_has_nccl(self):
        """True when the batch-norm statistics of this network are all-reduced through NCCL (one communicator: its
        collectives must stay on one stream, in one order); the peer-memory exchange has no such restriction."""
        return self.dp is not None and not getattr(self.dp, "peer_sync", False)

    def join(self):
        """Make the current stream wait for every side chain of this network."""
        self.wgrad_branch.join()
        self.side_branch.join()

    def ld(self, c):
        """channel stride of a buffer holding c channels"""
        return ru16(c) if self.bf16 else c

    def act_buffer(self, h, w, c, dtype=None):
        """zero-initialised so that pad channels stay zero forever (kernels only write the real channels)"""
        return torch.zeros(self.B, h, w, self.ld(c), dtype=dtype or self.adt, device=self.device)

    def plan(self, name, h, w, ld_in, fused_out=False, dx=True, dx_dtype=None):
        """Fix the geometry of layer `name` for an input of h x w pixels with channel stride ld_in.
        fused_out: (bf16 only) a layer without batch-norm and activation whose bias is added in the conv epilogue,
        writing the caller's output buffer directly; no raw z is kept."""
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
        st.ld_in = ld_in
        st.ldz = self.ld(L.cout)
        st.fused = bool(fused_out and self.bf16 and not L.bn and L.act == "none")
        st.z = None if st.fused else torch.empty(B, oh, ow, st.ldz, dtype=self.adt, device=self.device)
        st.dz = torch.zeros(B, oh, ow, st.ldz, dtype=self.adt, device=self.device)
        # gradient w.r.t. the layer input (preallocated: nothing is allocated inside a step / a CUDA graph)
        st.dx = torch.zeros(B, h, w, ld_in, dtype=dx_dtype or self.adt, device=self.device) if dx else None
        if self.bf16:
            # split-K workspaces of the forward and the data-gradient launch (only the few-tile layers get one);
            # per layer and per network application, because applications run concurrently on different streams
            fwd_w, bwd_w = (0, 1) if L.kind == "conv" else (1, 0)
            st.splitk_f = K.splitk_workspace(st.shape, fwd_w, ld_in, self.device)
            st.splitk_b = K.splitk_workspace(st.shape, bwd_w, st.ldz, self.device) if dx else None
            # integer limb accumulators of the fused batch-norm moments (order-independent atomics): the forward pass
            # is bitwise reproducible (ACG_DETERMINISTIC=0: fp64 atomics, for A/B timing only)
            st.stats_fix = st.fix_view if L.bn and DETERMINISTIC else None
        if self.bf16 and name not in self.store.packs:
            # forward / backward-data packs; conv2d_transpose swaps the roles (see include/acg_b200.h)
            fwd_which, bwd_which = (0, 1) if L.kind == "conv" else (1, 0)
            pf = torch.empty(K.pack_size(st.shape, fwd_which, ld_in), dtype=torch.bfloat16, device=self.device)
            pb = torch.empty(K.pack_size(st.shape, bwd_which, st.ldz), dtype=torch.bfloat16, device=self.device)
            self.store.packs[name] = (st.shape, fwd_which, ld_in, pf, bwd_which, st.ldz, pb)
        # first layers (8-channel frame operands): forward through the halo kernel's pixel-pair mode
        st.pair = bool(self.bf16 and PAIR_FIRST and L.kind == "conv" and ld_in == 8 and K.pair_ok(st.shape, ld_in))
        if st.pair and name not in self.store.pair_packs:
            self.store.pair_packs[name] = (st.shape, torch.empty(K.pack_size(st.shape, 2, 16), dtype=torch.bfloat16,
                                                                 device=self.device))
        return oh, ow

    def zero_reductions(self):
        self.f64.zero_()

    # -- convolution dispatch ------------------------------------------------------------------------------
    def _conv_fwd(self, st, x, out, ld_out, bias=None, stats=None, bn=None, peer=None):
        L = st.spec
        if self.bf16:
            pk = self.store.packs[L.name]
            fn = K.conv_fprop_tc if L.kind == "conv" else K.conv_dgrad_tc
            pair = getattr(st, "pair", False)
            fn(st.shape, x, self.store.pair_packs[L.name][1] if pair else pk[3], out, st.ld_in, ld_out, bias=bias,
               stats=stats, bn=bn, splitk=st.splitk_f, stats_fix=st.stats_fix if stats is not None else None, peer=peer,
               pair_x=pair)
        else:
            w = self.store.views[L.name + "/weights"]
            (K.conv_fprop_f32 if L.kind == "conv" else K.conv_dgrad_f32)(st.shape, x, w, out)

    def _sync_moments(self, st, beta):
        """Sum the [2C] moments over the ranks and finalise mean / rstd / scale / shift over the global batch."""
        L, dp = st.spec, self.dp
        if dp.peer_sync:       # one launch: push to the peers' mailboxes, wait, sum in rank order, finalise
            dp.mailbox.allreduce_f64(st.stats, 2 * L.cout, st.slot_f,
                                     bn=(L.cout, beta, st.rows * dp.world, BN_EPS, st.mean, st.rstd, st.scale, st.shift))
        else:
            dp.allreduce_sum(st.stats)
            K.bn_finalize(st.stats, beta, st.rows * dp.world, L.cout, 1, st.mean, st.rstd, st.scale, st.shift, BN_EPS)

    # -- one layer forward: conv -> (bias | batch-norm) -> activation --------------------------------------
    def layer_fwd(self, name, x, out, ld_out, cat=None):
        """cat = (actions [B, A], channel offset): `out` is a concat buffer whose channels behind the layer's features
        hold the tiled action vector (models.py:16,38,84); written by the activation pass itself on the bf16 path."""
        st = self.layers[name]
        L = st.spec
        st.x = x
        if st.fused:
            self._conv_fwd(st, x, out, ld_out, bias=self.store.views[name + "/biases"] if L.bias else None)
            return
        if L.bn and self.bf16:
            # moments in the conv epilogue; on one GPU the last CTA also finalises mean / rstd / scale / shift
            beta = self.store.views[name + "/BatchNorm/beta"]
            if self.dp is None and RAW_MOMENTS and K.bn_finalize_act_fwd_ok(L.cout, st.ldz, ld_out, L.act):
                # the conv launch only ADDS its moments (no ticket, no last-CTA pass at its end); the activation pass
                # completes and finalises them per block (~2.8 us less per layer than the in-kernel finalize below)
                self._conv_fwd(st, x, st.z, st.ldz, stats=st.stats)
                K.bn_finalize_act_fwd(st.z, st.rows, L.cout, st.ldz, st.stats, st.stats_fix, beta, st.rows, BN_EPS,
                                      st.mean, st.rstd, st.scale, st.shift, L.act, out, ld_out, cat=cat,
                                      hw=st.out_hw[0] * st.out_hw[1])
                return
            if self.dp is None:
                self._conv_fwd(st, x, st.z, st.ldz, stats=st.stats,
                               bn=(st.counter, beta, st.mean, st.rstd, st.scale, st.shift, st.rows, BN_EPS))
            elif self.dp.peer_sync and DP_FUSED and st.stats_fix is not None:
                # SyncBN inside the conv launch: its last CTA exchanges the totals with the peers and finalises
                self._conv_fwd(st, x, st.z, st.ldz, stats=st.stats,
                               bn=(st.counter, beta, st.mean, st.rstd, st.scale, st.shift, st.rows * self.dp.world, BN_EPS),
                               peer=(self.dp.mailbox, st.slot_f))
            else:
                # the ticket alone (rows 0): this rank's totals in a fixed order, finalised after the exchange
                self._conv_fwd(st, x, st.z, st.ldz, stats=st.stats,
                               bn=(st.counter, None, None, None, None, None, 0, BN_EPS) if st.stats_fix is not None else None)
                self._sync_moments(st, beta)           # SyncBN: statistics over the GLOBAL batch
            if cat is not None:
                K.bn_act_fwd_cat(st.z, st.rows, L.cout, st.ldz, st.scale, st.shift, L.act, out, ld_out, cat[0],
                                 st.out_hw[0] * st.out_hw[1], cat[1])
            else:
                K.bn_act_fwd(st.z, st.rows, L.cout, st.ldz, 1, st.scale, st.shift, L.act, out, ld_out)
            return
        self._conv_fwd(st, x, st.z, st.ldz)
        if L.bn:
            K.bn_stats(st.z, st.rows, L.cout, st.ldz, 1, st.stats)
            beta = self.store.views[name + "/BatchNorm/beta"]
            if self.dp is not None:
                self._sync_moments(st, beta)           # SyncBN: statistics over the GLOBAL batch
            else:
                K.bn_finalize(st.stats, beta, st.rows, L.cout, 1, st.mean, st.rstd, st.scale, st.shift, BN_EPS)
            K.bn_act_fwd(st.z, st.rows, L.cout, st.ldz, 1, st.scale, st.shift, L.act, out, ld_out)
        else:
            bias = self.store.views[name + "/biases"] if L.bias else None
            K.bn_act_fwd(st.z, st.rows, L.cout, st.ldz, 1, None, bias, L.act, out, ld_out)

    # -- one layer backward ----------------------------------------------------------------------------
    def _fused_red(self, consumer, producer=None):
        """Arguments of the batch-norm backward reduction of layer `consumer` for the epilogue of the data-gradient
        kernel (of layer state `producer`) that produces its dA (None when the separate acg_bn_act_bwd_reduce pass has
        to run)."""
        if consumer is None or not self.bf16 or FUSE_BWD_REDUCE == "0":
            return None
        if FUSE_BWD_REDUCE == "halo":
            if producer is None:
                return None
            if not hasattr(producer, "dgrad_kind"):
                which = 1 if producer.spec.kind == "conv" else 0
                producer.dgrad_kind = K.kernel_kind(producer.shape, which, producer.ldz,
                                                    getattr(producer, "dx_channels", 0))
            if producer.dgrad_kind != 1:
                return None
        sc = self.layers[consumer]
        Lc = sc.spec
        if not Lc.bn or Lc.cout % 16 != 0 or sc.z is None:
            return None
        # the producer's dx tensor is (one of) the consumer's dA: remember that its share of the sums is taken care of
        if not hasattr(sc, "red_fused"):
            sc.red_fused = set()
        if producer is not None and producer.dx is not None:
            sc.red_fused.add(producer.dx.data_ptr())
        else:
            sc.red_fused.add(None)          # legacy "everything fused" mode
        return (sc.red, sc.z, sc.ldz, Lc.cout, Lc.act, sc.mean, sc.rstd, sc.shift)

    def layer_bwd(self, name, dA, ld_d, dA2=None, need_dx=True, need_dw=True, dx_dtype=None, consumer=None):
        """consumer: the layer whose activation gradient this layer's dx is (its batch-norm backward sums are then
        accumulated in the epilogue of the data-gradient kernel, and its own layer_bwd skips the reduction pass)."""
        st = self.layers[name]
        L = st.spec
        if L.bn:
            mean, rstd, shift = st.mean, st.rstd, st.shift
        else:
            mean, rstd, shift = None, None, (self.store.views[name + "/biases"] if L.bias else None)
        if dA is st.dz:
            # a fused (no batch-norm, no activation) layer whose producer already wrote dz in operand layout
            # (acg_dna_bwd -> g/tconv4): only the bias gradient (column sums) is left to do
            assert st.fused and dA2 is None
            if need_dw and L.bias:
                with self.wgrad_branch:     # feeds the optimizer only: off the data-gradient chain
                    K.bn_act_bwd_reduce(st.dz, None, st.ldz, None, st.ldz, st.rows, st.ldz, 1, None, None, None,
                                        "none", st.red)
                    K.bias_grad(st.red, L.cout, 1.0, self.store.gviews[name + "/biases"])
        else:
            # batch-norm backward sums: the share of a gradient tensor whose producing kernel accumulated it in its
            # epilogue is already in st.red; the others go through the reduction pass (it accumulates as well)
            fused = getattr(st, "red_fused", set())
            todo = [d for d in (dA, dA2) if d is not None and not (None in fused or d.data_ptr() in fused)]
            sync_fused = (self.dp is not None and L.bn and self.dp.peer_sync and DP_FUSED and self.bf16
                          and len(todo) == len([d for d in (dA, dA2) if d is not None]))
            if todo and sync_fused:
                # SyncBN backward: the reduction pass's last block sums the terms over the ranks itself
                K.bn_act_bwd_reduce_sync(todo[0], todo[1] if len(todo) > 1 else None, ld_d, st.z, st.ldz, st.rows, L.cout,
                                         mean, rstd, shift, L.act, st.red, st.counter_b, self.dp.mailbox, st.slot_b)
            elif todo:
                K.bn_act_bwd_reduce(todo[0], todo[1] if len(todo) > 1 else None, ld_d, st.z, st.ldz, st.rows, L.cout, 1,
                                    mean, rstd, shift, L.act, st.red)
            world = 1
            if self.dp is not None:
                world = self.dp.world
                if todo and sync_fused:
                    pass
                elif L.bn and self.dp.peer_sync:
                    self.dp.mailbox.allreduce_f64(st.red, 2 * L.cout, st.slot_b)
                elif L.bn:
                    self.dp.allreduce_sum(st.red)
            dpar = None
            if need_dw:
                dpar = self.store.gviews[name + ("/BatchNorm/beta" if L.bn else "/biases")]
            K.bn_act_bwd_apply(dA, dA2, ld_d, st.z, st.ldz, st.rows, L.cout, 1, mean, rstd, shift, L.act, L.bn,
                               st.red, st.dz, dpar, norm_rows=st.rows * world,
                               dbeta_scale=(1.0 / world if L.bn else 1.0), ld_dz=st.ldz)
        if need_dw:
            # the weight gradient only feeds the optimizer: it runs beside the data-gradient chain (join() before
            # the gradient all-reduce / optimizer step)
            dw = self.store.gviews[name + "/weights"]
            with self.wgrad_branch:
                if self.bf16:
                    if L.kind == "conv":
                        K.conv_wgrad_tc(st.shape, st.x, st.dz, dw, st.ld_in, st.ldz)
                    else:
                        K.conv_wgrad_tc(st.shape, st.dz, st.x, dw, st.ldz, st.ld_in)
                elif L.kind == "conv":
                    K.conv_wgrad_f32(st.shape, st.x, st.dz, dw)
                else:
                    K.conv_wgrad_f32(st.shape, st.dz, st.x, dw)
        if need_dx:
            if st.dx is None:
                raise RuntimeError("layer %s was planned without an input-gradient buffer" % name)
            if self.bf16:
                pk = self.store.packs[name]
                fn = K.conv_dgrad_tc if L.kind == "conv" else K.conv_fprop_tc
                # dx_channels: the input is a concat buffer whose tail (tiled actions) needs no gradient
                fn(st.shape, st.dz, pk[6], st.dx, st.ldz, st.ld_in, red=self._fused_red(consumer, st), splitk=st.splitk_b,
                   n_limit=getattr(st, "dx_channels", 0))
            else:
                w = self.store.views[name + "/weights"]
                (K.conv_dgrad_f32 if L.kind == "conv" else K.conv_fprop_f32)(st.shape, st.dz, w, st.dx)
        return st.dx


class GeneratorRun(NetRun):
    """One application of the generator (DNA: models.py:24-74, direct: models.py:8-22)."""

    def __init__(self, store, batch, device, dna, ksize, dp=None, precision="bf16", branches=True):
        super().__init__(store, batch, device, dp, precision, branches)
        self.dna, self.ksize = dna, ksize
        B = batch
        dev = device
        Ls = self.layers
        # bf16 copy of the frame, 3 -> 8 channels (zero pad)
        self.img_in = torch.zeros(B, IMG, IMG, FRAME_LD, dtype=self.adt, device=dev) if self.bf16 else None
        h = w = IMG
        ld = FRAME_LD if self.bf16 else 3
        for n in ["g/conv1", "g/conv2", "g/conv3", "g/conv4"]:
            h, w = self.plan(n, h, w, ld, dx=(n != "g/conv1"))
            ld = self.ld(Ls[n].spec.cout)
        c4 = Ls["g/conv4"].spec.cout
        self.cat_c = c4
        self.cat = self.act_buffer(h, w, c4 + ACTION_DIM)                        # models.py:16,38
        cat_ld = self.cat.shape[3]
        h1, w1 = self.plan("g/tconv1", h, w, cat_ld)
        Ls["g/tconv1"].dx_channels = c4
        h2, w2 = self.plan("g/tconv2", h1, w1, self.ld(Ls["g/tconv1"].spec.cout))
        ld2 = self.ld(Ls["g/tconv2"].spec.cout)
        if dna:
            hs, ws = self.plan("g/sconv3", h2, w2, ld2)
            hs, ws = self.plan("g/sconv4", hs, ws, self.ld(32))
            hs, ws = self.plan("g/sconv5", hs, ws, self.ld(16), fused_out=True)
            assert (hs, ws) == (1, 1)
        h3, w3 = self.plan("g/tconv3", h2, w2, ld2)
        h4, w4 = self.plan("g/tconv4", h3, w3, self.ld(Ls["g/tconv3"].spec.cout), fused_out=dna)
        assert (h4, w4) == (IMG, IMG)
        for n, st in Ls.items():
            if n in ("g/conv4", "g/tconv4", "g/sconv5"):
                continue
            st.a = self.act_buffer(st.out_hw[0], st.out_hw[1], st.spec.cout)
        self.g_out = torch.empty(B, IMG, IMG, 3, device=dev)
        self.dg_out = torch.empty(B, IMG, IMG, 3, device=dev)
        if dna:
            kk = ksize * ksize
            self.logits = torch.empty(B, IMG, IMG, kk, device=dev)               # fp32, dense: what acg_dna_* reads
            # bf16 path: acg_dna_bwd writes g/tconv4's dz (bf16, channels padded to 16) directly
            self.dlogits = Ls["g/tconv4"].dz if self.bf16 else torch.empty(B, IMG, IMG, kk, device=dev)
            self.state = torch.empty(B, STATE_DIM, device=dev)
            self.dstate = torch.empty(B, STATE_DIM, device=dev)
        else:
            self.state = None

    def forward(self, img, actions, need_state=True, out=None):
        """need_state=False skips the state head (sconv3-5): train_d only fetches the frame (train.py:132-144).
        out: write the generated frame there instead of self.g_out (recursive rollout: predicted[j])."""
        g_out = self.g_out if out is None else out
        self.zero_reductions()
        self.img = img
        Ls = self.layers
        rows = self.B * IMG * IMG
        if self.bf16:
            K.pack_frames(img, None, self.img_in, rows)
            x = self.img_in
        else:
            x = img
        for n in ["g/conv1", "g/conv2", "g/conv3"]:
            self.layer_fwd(n, x, Ls[n].a, Ls[n].a.shape[3])
            x = Ls[n].a
        ld = self.cat.shape[3]
        if self.bf16:       # concat([conv4, tile(action)], 3) (train.py:48-49): written by conv4's activation pass
            self.layer_fwd("g/conv4", x, self.cat, ld, cat=(actions, self.cat_c))
        else:
            self.layer_fwd("g/conv4", x, self.cat, ld)
            K.tile_actions(actions, self.B, self.cat.shape[1] * self.cat.shape[2], self.cat, ld, self.cat_c)
        self.layer_fwd("g/tconv1", self.cat, Ls["g/tconv1"].a, Ls["g/tconv1"].a.shape[3])
        self.layer_fwd("g/tconv2", Ls["g/tconv1"].a, Ls["g/tconv2"].a, Ls["g/tconv2"].a.shape[3])
        t2 = Ls["g/tconv2"].a
        if self.dna and need_state:
            with self.side_branch:      # the state head runs beside tconv3 / tconv4 / DNA
                self.layer_fwd("g/sconv3", t2, Ls["g/sconv3"].a, Ls["g/sconv3"].a.shape[3])
                self.layer_fwd("g/sconv4", Ls["g/sconv3"].a, Ls["g/sconv4"].a, Ls["g/sconv4"].a.shape[3])
                self.layer_fwd("g/sconv5", Ls["g/sconv4"].a, self.state, STATE_DIM)
        self.layer_fwd("g/tconv3", t2, Ls["g/tconv3"].a, Ls["g/tconv3"].a.shape[3])
        if self.dna:
            kk = self.ksize * self.ksize
            self.layer_fwd("g/tconv4", Ls["g/tconv3"].a, self.logits, kk)        # logits (+bias), fp32 dense
            K.dna_fwd(self.logits, img, g_out, self.ksize)                       # models.py:60-72
        else:
            self.layer_fwd("g/tconv4", Ls["g/tconv3"].a, g_out, 3)               # tanh image
        self.side_branch.join()
        return g_out, self.state

    def backward(self, with_state):
        """dg_out (and dstate when with_state) must be filled; accumulates into store.grad."""
        Ls = self.layers
        d2b = None
        if self.dna and with_state:
            with self.side_branch:      # state head beside the frame head
                ds = self.layer_bwd("g/sconv5", self.dstate, STATE_DIM, consumer="g/sconv4")
                ds = self.layer_bwd("g/sconv4", ds, ds.shape[3], consumer="g/sconv3")
                d2b = self.layer_bwd("g/sconv3", ds, ds.shape[3], consumer="g/tconv2")
        if self.dna:
            K.dna_bwd(self.logits, self.img, self.dg_out, self.dlogits, self.ksize)
            d = self.layer_bwd("g/tconv4", self.dlogits, self.dlogits.shape[3], consumer="g/tconv3")
        else:
            d = self.layer_bwd("g/tconv4", self.dg_out, 3, consumer="g/tconv3")
        d3 = self.layer_bwd("g/tconv3", d, d.shape[3], consumer="g/tconv2")
        self.side_branch.join()
        d = self.layer_bwd("g/tconv2", d3, d3.shape[3], dA2=d2b, consumer="g/tconv1")
        d = self.layer_bwd("g/tconv1", d, d.shape[3], consumer="g/conv4")       # [B,4,4,ld(c4+10)]
        d = self.layer_bwd("g/conv4", d, d.shape[3], consumer="g/conv3")        # first c4 channels
        d = self.layer_bwd("g/conv3", d, d.shape[3], consumer="g/conv2")
        d = self.layer_bwd("g/conv2", d, d.shape[3], consumer="g/conv1")
        self.layer_bwd("g/conv1", d, d.shape[3], need_dx=False)
        self.join()


class DiscriminatorRun(NetRun):
    """One application of build_discriminator (models.py:76-88) to concat([img, frame], 3)."""

    def __init__(self, store, batch, device, dp=None, precision="bf16", branches=True):
        super().__init__(store, batch, device, dp, precision, branches)
        B = batch
        Ls = self.layers
        if self.bf16:                                                            # train.py:64,68; 6 -> 8 channels
            self.d_in = torch.zeros(B, IMG, IMG, FRAME_LD, dtype=self.adt, device=device)
        else:
            self.d_in = self.act_buffer(IMG, IMG, 6)
        h = w = IMG
        h, w = self.plan("d/conv1", h, w, self.d_in.shape[3], dx_dtype=torch.float32)
        h, w = self.plan("d/conv2", h, w, self.ld(64))
        self.cat = self.act_buffer(h, w, 128 + ACTION_DIM)                       # models.py:84 (R3: 16x16)
        ld = self.cat.shape[3]
        for n in ["d/conv3", "d/conv4", "d/conv5", "d/conv6"]:
            h, w = self.plan(n, h, w, ld)
            ld = self.ld(Ls[n].spec.cout)
        Ls["d/conv3"].dx_channels = 128
        for n, st in Ls.items():
            if n in ("d/conv2", "d/conv6"):
                continue
            st.a = self.act_buffer(st.out_hw[0], st.out_hw[1], st.spec.cout)
        self.logits = torch.empty(B, 2, 2, 1, device=device)                     # fp32, dense
        self.n_logits = self.logits.numel()
        self.dlogits = torch.empty(self.n_logits, device=device)

    def forward(self, img, frame, actions):
        self.zero_reductions()
        rows = self.B * IMG * IMG
        ld_in = self.d_in.shape[3]
        if self.bf16:
            K.pack_frames(img, frame, self.d_in, rows)                           # concat([img, frame], 3), one pass
        else:
            K.copy_channels(img, 3, 0, self.d_in, ld_in, 0, rows, 3)
            K.copy_channels(frame, 3, 0, self.d_in, ld_in, 3, rows, 3)
        Ls = self.layers
        self.layer_fwd("d/conv1", self.d_in, Ls["d/conv1"].a, Ls["d/conv1"].a.shape[3])
        ld = self.cat.shape[3]
        if self.bf16:       # concat([conv2, tile(action)], 3) (models.py:84): written by conv2's activation pass
            self.layer_fwd("d/conv2", Ls["d/conv1"].a, self.cat, ld, cat=(actions, 128))
        else:
            self.layer_fwd("d/conv2", Ls["d/conv1"].a, self.cat, ld)
            K.tile_actions(actions, self.B, self.cat.shape[1] * self.cat.shape[2], self.cat, ld, 128)
        x = self.cat
        for n in ["d/conv3", "d/conv4", "d/conv5"]:
            self.layer_fwd(n, x, Ls[n].a, Ls[n].a.shape[3])
            x = Ls[n].a
        self.layer_fwd("d/conv6", x, self.logits, 1)
        return self.logits

    def backward(self, need_dw, need_dinput):
        """dlogits must be filled.  Returns d(d_in) [B,64,64,ld(6)] fp32 when need_dinput."""
        d = self.layer_bwd("d/conv6", self.dlogits, 1, need_dw=need_dw, consumer="d/conv5")
        d = self.layer_bwd("d/conv5", d, d.shape[3], need_dw=need_dw, consumer="d/conv4")
        d = self.layer_bwd("d/conv4", d, d.shape[3], need_dw=need_dw, consumer="d/conv3")
        d = self.layer_bwd("d/conv3", d, d.shape[3], need_dw=need_dw, consumer="d/conv2")   # [B,16,16,ld(138)]
        d = self.layer_bwd("d/conv2", d, d.shape[3], need_dw=need_dw, consumer="d/conv1")
        d = self.layer_bwd("d/conv1", d, d.shape[3], need_dw=need_dw, need_dx=need_dinput, dx_dtype=torch.float32)
        self.join()
        return d


class TFOptimizer:
    """tf.train.AdamOptimizer / RMSPropOptimizer slots over a ParamStore (train.py:91-102)."""

    def __init__(self, store, kind):
        if kind not in ("adam", "rmsprop"):
            raise ValueError("unexpected opt argument")
        self.kind, self.store, self.t = kind, store, 0
        self.lr_dev = torch.zeros(1, device=store.flat.device)
        if kind == "adam":
            self.lr = 1e-3                                                       # train.py:20
            self.m = torch.zeros_like(store.flat)
            self.v = torch.zeros_like(store.flat)
        else:
            self.lr = 5e-5                                                       # train.py:93
            self.ms = torch.ones_like(store.flat)                                # TF initialises ms to ONE

    def tick(self):
        """Host side of a step: advance t and publish this step's (bias-corrected) rate to the device scalar the
        kernel reads.  Kept apart from enqueue() so that the kernel launches can live inside a CUDA graph."""
        self.t += 1
        lr_t = self.lr
        if self.kind == "adam":
            lr_t = self.lr * math.sqrt(1.0 - 0.999 ** self.t) / (1.0 - 0.9 ** self.t)
        self.lr_dev.fill_(lr_t)

    def enqueue(self, clip=K.NO_CLIP, grad_scale=1.0):
        s = self.store
        if self.kind == "adam":
            K.adam_step(s.flat, s.grad, self.m, self.v, 0.0, clip=clip, grad_scale=grad_scale, lr_dev=self.lr_dev)
        else:
            K.rmsprop_step(s.flat, s.grad, self.ms, 0.0, clip=clip, grad_scale=grad_scale, lr_dev=self.lr_dev)
        s.refresh_packs()

    def step(self, clip=K.NO_CLIP, grad_scale=1.0):
        self.tick()
        self.enqueue(clip, grad_scale)
