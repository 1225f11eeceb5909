"""`Trainer`: the step wrappers of train.py:27-176 on top of the acg_b200 engine.

Same constructor flags and methods as the reference (`Trainer(sess, arg_adv, arg_loss, arg_opt, arg_transform)`,
`pretrain_g`, `train_g`, `train_d`, `test`, `test_sequence`); `sess` is accepted and ignored (there is no TF
session -- each method enqueues one device step).  Repairs R1-R6 of SURVEY.md section 0 are applied; the D step
is update-THEN-clip (the reference leaves the order of train.py:140,143 unspecified).
"""
import contextlib
import functools
import math

import numpy as np
import torch

from . import engine as E
from . import kernels as K

BATCH_SIZE = 64      # train.py:19
L2_WEIGHT = 0.05     # train.py:22
DNA_KSIZE = 6        # train.py:53-54 passes ksize=6

import os as _os

# run train_d's optimizer part on a side stream under the next call's generator forward (ACG_OVERLAP_D_UPDATE=0: inline)
OVERLAP_D_UPDATE = _os.environ.get("ACG_OVERLAP_D_UPDATE", "1") != "0"

# replay train_g's generator forward on a second stream under the discriminator backward of the preceding train_d
OVERLAP_G_FWD = _os.environ.get("ACG_OVERLAP_G_FWD", "1") != "0"


def _on_device(fn):
    """Every kernel is launched on torch's CURRENT stream of the CURRENT device: make the trainer's device current for
    the duration of the call, so that Trainer(device='cuda:1') works without a global torch.cuda.set_device."""
    @functools.wraps(fn)
    def wrapped(self, *a, **kw):
        with torch.cuda.device(self.device):
            return fn(self, *a, **kw)
    return wrapped


SUMMARY_KEYS = ["discriminator_direct_loss", "discriminator_gen_loss", "discriminator_loss", "g_loss",
                "g_l2_loss", "g_adv_loss", "g_psnr"]


class DataParallel:
    """Batch-sharded data parallelism (one process per GPU, torch.distributed for the plumbing).
    Collectives on the path: the flat gradient bucket of the network being updated, the per-layer batch-norm
    moment vectors (SyncBN: the reference normalises over the WHOLE batch) and the loss sums."""

    def __init__(self, group=None, device=None, peer_sync=None):
        """peer_sync: exchange the batch-norm vectors through NVLink peer memory (csrc/peer.cu) instead of NCCL.
        Default: on for CUDA runs (ACG_DP_SYNC=nccl switches it off), off for the CPU/gloo host-logic tests."""
        import os
        import torch.distributed as dist
        self.dist = dist
        self.group = group
        self.world = dist.get_world_size(group)
        self.rank = dist.get_rank(group)
        if peer_sync is None:
            peer_sync = (dist.get_backend(group) == "nccl" and torch.cuda.is_available()
                         and os.environ.get("ACG_DP_SYNC", "peer") != "nccl")
        self.peer_sync = bool(peer_sync)
        self.mailbox = None
        if self.peer_sync:
            from .peer import Mailbox
            dev = device if device is not None else torch.device("cuda", torch.cuda.current_device())
            self.mailbox = Mailbox(dist, group, dev)

    def allreduce_sum(self, t):
        self.dist.all_reduce(t, op=self.dist.ReduceOp.SUM, group=self.group)


class Trainer:
    def __init__(self, sess=None, arg_adv=True, arg_loss="bce", arg_opt="adam", arg_transform=True,
                 batch_size=BATCH_SIZE, ksize=DNA_KSIZE, device=None, params=None, seed=7, dp=None,
                 precision="bf16", use_graphs=True, branches=True):
        if arg_loss not in ("bce", "wass"):
            raise ValueError("unexpected loss argument")          # ops.py:35,47
        if arg_opt not in ("adam", "rmsprop"):
            raise ValueError("unexpected opt argument")           # train.py:98
        if not torch.cuda.is_available():
            raise RuntimeError("acg_b200 needs a CUDA device: there is no CPU fallback")
        self.device = torch.device(device if device is not None else "cuda")
        self.arg_adv, self.arg_loss, self.arg_opt, self.arg_transform = arg_adv, arg_loss, arg_opt, arg_transform
        self.B = batch_size                      # LOCAL batch (per rank)
        self.dp = dp
        self.world = dp.world if dp is not None else 1
        self.GB = self.B * self.world            # global batch: every normaliser of the losses uses this
        self.ksize = ksize
        if precision not in ("bf16", "fp32"):
            raise ValueError("precision must be 'bf16' (tcgen05 kernels) or 'fp32' (SIMT kernels)")
        self.precision = precision
        g_spec = E.g_dna_spec(ksize) if arg_transform else E.g_direct_spec()
        d_spec = E.d_spec()
        if params is None:
            rng = np.random.RandomState(seed)    # train.py:14 seeds NumPy with 7
            params = E.xavier_init(g_spec, rng)
            params.update(E.xavier_init(d_spec, rng))
        dev = self.device
        self.g_store = E.ParamStore(g_spec, dev, params)
        self.d_store = E.ParamStore(d_spec, dev, params)
        self.g_run = E.GeneratorRun(self.g_store, self.B, dev, arg_transform, ksize, dp, precision, branches)
        self.d_gen = E.DiscriminatorRun(self.d_store, self.B, dev, dp, precision, branches)
        self.d_real = E.DiscriminatorRun(self.d_store, self.B, dev, dp, precision, branches)
        # D(real) is independent of G and D(generated) until the optimizer step: it gets its own chain of the step
        nccl_bn = dp is not None and not getattr(dp, "peer_sync", False)
        self.real_branch = E.Branch(dev) if branches and not nccl_bn else E._NoBranch()
        self.g_store.refresh_packs()
        self.d_store.refresh_packs()
        self.g_opt = E.TFOptimizer(self.g_store, arg_opt)            # train.py:100
        self.g_pretrain_opt = E.TFOptimizer(self.g_store, arg_opt)   # train.py:101
        self.d_opt = E.TFOptimizer(self.d_store, arg_opt)            # train.py:102
        # device-resident scalars: fsum = [sum|g-n|, sum(g-n)^2, gdl]; sc = [adv, d_direct, d_gen, state]
        self.fsum = torch.zeros(3, dtype=torch.float64, device=dev)
        self.sc = torch.zeros(4, dtype=torch.float32, device=dev)
        self.zero_state = torch.zeros(self.B, E.STATE_DIM, device=dev)
        self.state_ss = torch.zeros(1, dtype=torch.float64, device=dev)     # sum of squares of the state residual
        self._state_slot = dp.mailbox.new_slot(1) if dp is not None and dp.peer_sync else None
        self._have = set()
        # static feed buffers + one captured CUDA graph per step kind (the step is ~300 small launches; replaying a
        # graph removes the per-launch host cost and the idle gaps between tiny kernels)
        self.use_graphs = bool(use_graphs)
        self.in_img = torch.zeros(self.B, E.IMG, E.IMG, 3, device=dev)
        self.in_next = torch.zeros(self.B, E.IMG, E.IMG, 3, device=dev)
        self.in_act = torch.zeros(self.B, E.ACTION_DIM, device=dev)
        self.in_state = torch.zeros(self.B, E.STATE_DIM, device=dev)
        self._copy_stream = None
        self._staging = None
        self._staging_u8 = None
        self._u8_pair = None
        self._arange = None
        self._idx_ring = None
        self._frames_host = None
        self._pending_fetch = None
        # the discriminator update (gradient all-reduce, optimizer, weight packs) of train_d runs on its own stream so
        # that the generator forward of the following train_g overlaps it; everything that reads D weights waits
        self._d_update_stream = None
        self._d_update_done = None
        self._g_stream = None           # train_g's generator forward under the preceding discriminator backward
        self._d_fwd_done = None
        self._rollout_bufs = {}
        self._graphs, self._calls, self._graph_launches = {}, {}, {}
        self.replayed_launches = 0      # kernels launched through graph replays (acg_launch_count sees eager ones)

    def _stage(self, img, nxt, act, st):
        """Copy the feeds into the static buffers the kernels (and the captured graphs) read.

        Host feeds go through a COPY STREAM into one of two device staging sets, so the host->device transfer of this
        call overlaps the device work of the previous call (train_d's graph is still running when train_g's feeds
        arrive); the compute stream then only does a device->device copy into the static buffers.  uint8 frames
        (what a decoded dataset holds) are shipped as uint8 -- a quarter of the bytes -- and scaled to [-1,1]
        (ops.py:195) by acg_gather_frames on the device."""
        if self._pending_fetch is not None:
            torch.cuda.current_stream().wait_event(self._pending_fetch)
            self._pending_fetch = None
        u8 = img.dtype == torch.uint8
        if u8 != (nxt.dtype == torch.uint8):
            raise ValueError("input_images and next_frame must have the same dtype (both uint8 or both float)")
        feeds = [(self.in_img, img), (self.in_next, nxt), (self.in_act, act)]
        if st is not None:
            feeds.append((self.in_state, st))
        else:
            self.in_state.zero_()
        if all(t.is_cuda for _, t in feeds):
            if u8:
                self._decode_u8(img.reshape(self.B, -1), nxt.reshape(self.B, -1))
                feeds = feeds[2:]
            for dst, t in feeds:
                dst.copy_(t.reshape(dst.shape), non_blocking=True)
            return
        if self._staging is None:
            if self._copy_stream is None:
                self._copy_stream = torch.cuda.Stream(device=self.device)
            self._staging = [[torch.empty_like(b) for b in (self.in_img, self.in_next, self.in_act, self.in_state)]
                             for _ in range(2)]
            self._staging_free = [torch.cuda.Event(), torch.cuda.Event()]
            self._staging_idx = 0
        if u8 and self._staging_u8 is None:
            fe = E.IMG * E.IMG * 3
            self._staging_u8 = [torch.empty(2, self.B, fe, dtype=torch.uint8, device=self.device) for _ in range(2)]
        k = self._staging_idx
        self._staging_idx ^= 1
        main, cs = torch.cuda.current_stream(), self._copy_stream
        cs.wait_event(self._staging_free[k])          # the compute stream is done with this set (two calls ago)
        ready = torch.cuda.Event()
        with torch.cuda.stream(cs):
            for i, (dst, t) in enumerate(feeds):
                if u8 and i < 2:
                    self._staging_u8[k][i].copy_(t.reshape(self.B, -1), non_blocking=True)
                else:
                    self._staging[k][i].copy_(t.reshape(dst.shape), non_blocking=True)
            ready.record(cs)
        main.wait_event(ready)
        for i, (dst, t) in enumerate(feeds):
            if u8 and i < 2:
                continue
            dst.copy_(self._staging[k][i], non_blocking=True)
        if u8:
            self._decode_u8(self._staging_u8[k])
        self._staging_free[k].record(main)

    def _decode_u8(self, a, b=None):
        """uint8 frames -> the fp32 static feeds.  a: [2,B,F] (img | next) or a, b: [B,F] each."""
        if b is not None:
            if self._u8_pair is None:
                self._u8_pair = torch.empty(2, self.B, a.shape[1], dtype=torch.uint8, device=self.device)
            self._u8_pair[0].copy_(a, non_blocking=True)
            self._u8_pair[1].copy_(b, non_blocking=True)
            a = self._u8_pair
        if self._arange is None:
            self._arange = torch.arange(self.B, dtype=torch.int32, device=self.device)
            self._zeros_i = torch.zeros(self.B, dtype=torch.int32, device=self.device)
        K.gather_frames(a, None, self._zeros_i, self._arange, self.in_img, self.in_next, None, None,
                        pair_stride=self.B, geometry=(1, 2 * self.B))

    def _stage_indexed(self, feeder, sample, t0, with_state):
        """Device-side feeder: the sequences are resident in HBM; a step ships 2 x B int32 indices and
        one gather kernel writes all four static feeds (frame pair, action++state, next state)."""
        if self._pending_fetch is not None:
            torch.cuda.current_stream().wait_event(self._pending_fetch)
            self._pending_fetch = None
        if self._idx_ring is None:
            self._idx_ring = [(torch.empty(2, self.B, dtype=torch.int32).pin_memory(), torch.cuda.Event())
                              for _ in range(8)]
            self._idx_dev = torch.empty(2, self.B, dtype=torch.int32, device=self.device)
            self._idx_k = 0
        host, ev = self._idx_ring[self._idx_k]
        self._idx_k = (self._idx_k + 1) % len(self._idx_ring)
        ev.synchronize()                               # the copy that last used this pinned slot has completed
        feeder.check(sample, t0, self.B)
        host[0].copy_(torch.as_tensor(sample, dtype=torch.int32))
        host[1].copy_(torch.as_tensor(t0, dtype=torch.int32))
        self._idx_dev.copy_(host, non_blocking=True)
        ev.record(torch.cuda.current_stream())
        K.gather_frames(feeder.frames, feeder.actions, self._idx_dev[0], self._idx_dev[1], self.in_img, self.in_next,
                        self.in_act, self.in_state)
        if not with_state:
            self.in_state.zero_()                      # train_d feeds zeros (train.py:137)

    def synchronize(self):
        """Block the host until every enqueued step (including a discriminator update on its side stream) is done."""
        self._wait_d_update()
        torch.cuda.synchronize(self.device)

    def _wait_d_update(self):
        """Make the current stream wait for a discriminator update that is still running on its side stream."""
        if self._d_update_done is not None:
            torch.cuda.current_stream().wait_event(self._d_update_done)
            self._d_update_done = None

    def _run(self, key, fn):
        """Eager on the first call (one-time kernel attribute setup), captured on the second, replayed after."""
        if not self.use_graphs:
            fn()
            return
        n = self._calls.get(key, 0)
        self._calls[key] = n + 1
        if n == 0:
            fn()
            return
        if key not in self._graphs:
            torch.cuda.synchronize()
            from . import _lib
            n0 = _lib.launch_count()
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                fn()
            self._graphs[key] = g
            self._graph_launches[key] = _lib.launch_count() - n0
        self._graphs[key].replay()
        self.replayed_launches += self._graph_launches[key]

    # ---- feeds -----------------------------------------------------------------------------------
    @staticmethod
    def _as_tensor(a):
        """numpy / torch (host or device) -> float32 torch tensor, no device copy yet"""
        if isinstance(a, torch.Tensor):
            return a if a.dtype in (torch.float32, torch.uint8) else a.float()
        a = np.asarray(a)
        if a.dtype != np.uint8:
            a = a.astype(np.float32, copy=False)
        return torch.from_numpy(np.ascontiguousarray(a))

    def _dev(self, a, shape):
        return self._as_tensor(a).to(self.device, non_blocking=True).reshape(shape).contiguous()

    def _feed(self, img, nxt, act, state=None):
        """Feeds stay where they are (host or device); _stage() copies them straight into the static buffers."""
        return (self._as_tensor(img), self._as_tensor(nxt), self._as_tensor(act),
                self._as_tensor(state) if state is not None else None)

    # ---- shared pieces -----------------------------------------------------------------------------
    def _g_losses(self, nxt, state_gt, want_grad, with_adv_grad):
        """Frame / state / adversarial generator losses (train.py:72-83) and, when want_grad, dL/dg_out and
        dL/dstate.  with_adv_grad: back-propagate g_adv_loss through D(gen) first (train_g)."""
        g = self.g_run
        self._have = {"g"}
        dadv = None
        w_l1 = (L2_WEIGHT if self.arg_transform else 1.0) / self.GB
        if with_adv_grad:
            sign_or_label = 1.0                                     # bce: labels = ones; wass: +mean
            K.dlogit_loss(self.d_gen.logits, self.d_gen.n_logits, self.arg_loss, sign_or_label,
                          1.0 / self.world, self.sc[0:1], self.d_gen.dlogits)
            dadv = self.d_gen.backward(need_dw=False, need_dinput=True)
        elif self.arg_adv:
            K.dlogit_loss(self.d_gen.logits, self.d_gen.n_logits, self.arg_loss, 1.0, 1.0, self.sc[0:1], None)
        self.fsum.zero_()
        K.frame_losses(g.g_out, nxt, self.fsum, g.dg_out if want_grad else None, w_l1,
                       1.0 if with_adv_grad else 0.0, dadv, dadv.shape[3] if dadv is not None else 0, 3)
        if self.arg_transform:
            n_st = self.B * E.STATE_DIM
            if self.dp is None:
                K.state_loss(g.state, state_gt, n_st, 1.0 / self.GB, 1.0, self.sc[3:4], g.dstate if want_grad else None)
            else:
                # train.py:77 is a Frobenius norm over the GLOBAL batch: local sum of squares -> sum over the ranks ->
                # loss and gradient from the global norm (kernels + the exchange only; no torch op on the path)
                K.state_loss(g.state, state_gt, n_st, 1.0 / self.GB, 1.0, None, None, sumsq_out=self.state_ss)
                if self.dp.peer_sync:
                    self.dp.mailbox.allreduce_f64(self.state_ss, 1, self._state_slot)
                else:
                    self.dp.allreduce_sum(self.state_ss)
                K.state_loss(g.state, state_gt, n_st, 1.0 / self.GB, 1.0, self.sc[3:4], g.dstate if want_grad else None,
                             sumsq_in=self.state_ss)

    def _scalars(self):
        """Host values of the loss scalars of the last step (synchronises; logging only)."""
        self._wait_d_update()        # eager collectives below must not interleave with the side-stream gradient all-reduce
        fs = self.fsum.clone()
        sc = self.sc.clone().double()
        if self.dp is not None:
            self.dp.allreduce_sum(fs)
            loc = sc.clone()
            loc[:3] /= self.world        # means over the local logits -> global mean
            loc[3] = 0.0                 # the state loss is already the global norm on every rank (_g_losses)
            self.dp.allreduce_sum(loc)
            loc[3] = sc[3]
            sc = loc
        fs = fs.cpu().numpy()
        sc = sc.cpu().numpy()
        n_el = self.GB * E.IMG * E.IMG * 3
        out = {}
        l2 = fs[0] / self.GB
        if self.arg_transform:
            l2 = l2 * L2_WEIGHT + sc[3]
        out["g_l2_loss"] = float(l2)
        out["g_psnr"] = float(10.0 * math.log10(1.0 / max(fs[1] / n_el, 1e-300)))
        if self.arg_adv:
            out["g_adv_loss"] = float(sc[0])
            out["g_loss"] = float(l2 + sc[0] + fs[2])              # train.py:81 (gdl is a SUM)
        else:
            out["g_loss"] = float(l2)
        if "d" in self._have:
            out["discriminator_direct_loss"] = float(sc[1])
            out["discriminator_gen_loss"] = float(sc[2])
            out["discriminator_loss"] = float(sc[1] + sc[2])
        return out

    def summaries(self):
        s = self._scalars()
        return {k: s[k] for k in SUMMARY_KEYS if k in s}

    def _sync_grads(self, store):
        if self.dp is not None:
            self.dp.allreduce_sum(store.grad)

    # ---- train.py:114-121 -------------------------------------------------------------------------------
    def pretrain_g(self, input_images, next_frame, actions, state):
        img, nxt, act, st = self._feed(input_images, next_frame, actions, state)
        self.enqueue_pretrain_g(img, nxt, act, st)
        return self._scalars()["g_loss"]

    def pretrain_g_indexed(self, feeder, sample, t0):
        """pretrain_g on frame pairs gathered on the device from a resident dataset (feeder.DeviceFeeder)."""
        self.enqueue_pretrain_g(stager=lambda: self._stage_indexed(feeder, sample, t0, True))
        return self._scalars()["g_loss"]

    @_on_device
    def enqueue_pretrain_g(self, img=None, nxt=None, act=None, st=None, stager=None):
        self._wait_d_update()
        self._join_g_stream()
        (stager or (lambda: self._stage(img, nxt, act, st)))()
        self.g_pretrain_opt.tick()
        self._run("pretrain_g", self._body_pretrain_g)

    def _body_pretrain_g(self):
        img, nxt, act, st = self.in_img, self.in_next, self.in_act, self.in_state
        self.g_store.grad.zero_()
        g_out, _ = self.g_run.forward(img, act)
        if self.arg_adv:
            self.d_gen.forward(img, g_out, act)      # self.g_loss is fetched -> D(gen) forward runs too
        self._g_losses(nxt, st, want_grad=True, with_adv_grad=False)
        self.g_run.backward(with_state=self.arg_transform)
        self._sync_grads(self.g_store)
        self.g_pretrain_opt.enqueue()

    # ---- train.py:123-130 -------------------------------------------------------------------------------
    def train_g(self, input_images, next_frame, actions, state):
        """Returns the generated frames of this step (train.py:124,130) as a NumPy array.  The array is a view of a
        pinned host buffer from a ring of 4: it stays valid until the fourth following train_g call."""
        img, nxt, act, st = self._feed(input_images, next_frame, actions, state)
        done = self.enqueue_train_g(img, nxt, act, st, fetch=True)
        done.synchronize()          # only the device->host copy of the frames; the backward pass keeps running
        return self._frames_host[self._frames_idx].numpy()

    def train_g_indexed(self, feeder, sample, t0):
        done = self.enqueue_train_g(fetch=True, stager=lambda: self._stage_indexed(feeder, sample, t0, True))
        done.synchronize()
        return self._frames_host[self._frames_idx].numpy()

    @_on_device
    def enqueue_train_g(self, img=None, nxt=None, act=None, st=None, fetch=False, stager=None):
        """The step is two captured graphs: (a) generator forward, (b) everything else.  (a) reads nothing the
        discriminator backward of a preceding train_d writes, so it is replayed on a second stream as soon as that
        train_d's FORWARD graph is done and overlaps the discriminator backward (most kernels of either chain leave
        SMs idle).  With fetch=True the frames leave for the host on the copy stream as soon as (a) is done,
        overlapping (b)."""
        main = torch.cuda.current_stream()
        gs = self._g_stream_for_overlap()
        if gs is not None and self._d_fwd_done is not None:
            gs.wait_event(self._d_fwd_done)            # feeds + generator buffers are free; D backward still runs
            self._d_fwd_done = None
            ctx = torch.cuda.stream(gs)
        else:
            gs, ctx = None, contextlib.nullcontext()
        with ctx:
            (stager or (lambda: self._stage(img, nxt, act, st)))()
            self.g_opt.tick()
            self._run("train_g_a", self._body_train_g_a)
            fwd_done = torch.cuda.Event()
            fwd_done.record(torch.cuda.current_stream())
        done = None
        if fetch:
            if self._frames_host is None:
                self._frames_host = [torch.empty(self.g_run.g_out.shape, dtype=torch.float32).pin_memory()
                                     for _ in range(4)]
                self._frames_idx = 0
                if self._copy_stream is None:
                    self._copy_stream = torch.cuda.Stream(device=self.device)
            self._frames_idx = (self._frames_idx + 1) % 4
            cs = self._copy_stream
            cs.wait_event(fwd_done)
            done = torch.cuda.Event()
            with torch.cuda.stream(cs):
                self._frames_host[self._frames_idx].copy_(self.g_run.g_out, non_blocking=True)
                done.record(cs)
            self._pending_fetch = done     # the NEXT step's forward overwrites g_out: it waits for this copy
        if gs is not None:
            main.wait_event(fwd_done)
        self._wait_d_update()              # D(generated) reads the discriminator weights / packs
        self._run("train_g_b", self._body_train_g_b)
        return done

    def _g_stream_for_overlap(self):
        nccl_bn = self.dp is not None and not self.dp.peer_sync    # NCCL batch-norm: one communicator, one stream
        if not self.use_graphs or not OVERLAP_G_FWD or nccl_bn:
            return None
        if self._g_stream is None:
            self._g_stream = torch.cuda.Stream(device=self.device)
        return self._g_stream

    def _join_g_stream(self):
        """Steps other than train_g that follow a train_d: nothing is pending on the generator stream, but the marker of
        the train_d forward must not leak into a later train_g."""
        self._d_fwd_done = None

    def _body_train_g_a(self):
        self.g_store.grad.zero_()
        self.g_run.forward(self.in_img, self.in_act)

    def _body_train_g_b(self):
        img, nxt, act, st = self.in_img, self.in_next, self.in_act, self.in_state
        g_out = self.g_run.g_out
        if self.arg_adv:
            self.d_gen.forward(img, g_out, act)
        self._g_losses(nxt, st, want_grad=True, with_adv_grad=self.arg_adv)
        self.g_run.backward(with_state=self.arg_transform)
        self._sync_grads(self.g_store)
        self.g_opt.enqueue()

    # ---- train.py:132-144 -------------------------------------------------------------------------------
    def train_d(self, input_images, next_frame, actions, summarize=False):
        img, nxt, act, st = self._feed(input_images, next_frame, actions, None)
        self.enqueue_train_d(img, nxt, act, need_state=summarize)
        return self._d_summaries(summarize)

    def train_d_indexed(self, feeder, sample, t0, summarize=False):
        self.enqueue_train_d(need_state=summarize, stager=lambda: self._stage_indexed(feeder, sample, t0, False))
        return self._d_summaries(summarize)

    def _d_summaries(self, summarize):
        if not summarize:
            return None
        # merged_summaries also holds the generator scalars (train.py:112,140)
        self._join_g_stream()
        self._g_losses(self.in_next, self.in_state, want_grad=False, with_adv_grad=False)
        self._have = {"g", "d"}
        return self.summaries()

    @_on_device
    def enqueue_train_d(self, img=None, nxt=None, act=None, need_state=False, stager=None):
        """need_state: also run the generator's state head (only the summaries of train.py:140 read it).
        Three graphs: (fwd) D(real), G, D(generated) forward + the logit losses, (bwd) the two discriminator backward
        passes, (update) gradient all-reduce + optimizer + clip + weight packs.  The update is replayed on a side
        stream: the next train_g's generator forward does not read D weights and overlaps it (with data parallelism
        that hides the 19 MB gradient all-reduce); that generator forward itself may start right after (fwd)."""
        self._wait_d_update()
        (stager or (lambda: self._stage(img, nxt, act, None)))()
        self.d_opt.tick()
        self._run("train_d_fwd_state" if need_state else "train_d_fwd", lambda: self._body_train_d_fwd(need_state))
        self._d_fwd_done = torch.cuda.Event()
        self._d_fwd_done.record(torch.cuda.current_stream())
        self._run("train_d_bwd", self._body_train_d_bwd)
        nccl_bn = self.dp is not None and not self.dp.peer_sync    # batch-norm all-reduces share the communicator
        if not self.use_graphs or not OVERLAP_D_UPDATE or nccl_bn:
            self._run("train_d_update", self._body_train_d_update)
        else:
            if self._d_update_stream is None:
                self._d_update_stream = torch.cuda.Stream(device=self.device)
            main, side = torch.cuda.current_stream(), self._d_update_stream
            side.wait_stream(main)
            with torch.cuda.stream(side):
                self._run("train_d_update", self._body_train_d_update)
                self._d_update_done = torch.cuda.Event()
                self._d_update_done.record(side)
        self._have = {"d"}

    def _body_train_d_fwd(self, need_state=False):
        img, nxt, act = self.in_img, self.in_next, self.in_act
        self.d_store.grad.zero_()
        with self.real_branch:
            self.d_real.forward(img, nxt, act)
        g_out, _ = self.g_run.forward(img, act, need_state=need_state)
        self.d_gen.forward(img, g_out, act)
        self.real_branch.join()
        gs = 1.0 / self.world
        if self.arg_loss == "bce":                                   # ops.py:38-42
            K.dlogit_loss(self.d_real.logits, self.d_real.n_logits, "bce", 0.9, gs, self.sc[1:2], self.d_real.dlogits)
            K.dlogit_loss(self.d_gen.logits, self.d_gen.n_logits, "bce", 0.0, gs, self.sc[2:3], self.d_gen.dlogits)
        else:                                                        # ops.py:43-45
            K.dlogit_loss(self.d_real.logits, self.d_real.n_logits, "wass", 1.0, gs, self.sc[1:2], self.d_real.dlogits)
            K.dlogit_loss(self.d_gen.logits, self.d_gen.n_logits, "wass", -1.0, gs, self.sc[2:3], self.d_gen.dlogits)

    def _body_train_d_bwd(self):
        with self.real_branch:
            self.d_real.backward(need_dw=True, need_dinput=False)
        self.d_gen.backward(need_dw=True, need_dinput=False)
        self.real_branch.join()

    def _body_train_d_update(self):
        self._sync_grads(self.d_store)
        self.d_opt.enqueue(clip=(-0.01, 0.01))                       # train.py:89, update then clip

    # ---- train.py:146-155 -------------------------------------------------------------------------------
    def test(self, input_images, next_frame, actions):
        img, nxt, act, st = self._feed(input_images, next_frame, actions, None)
        g_out, g_state = self.enqueue_test(img, nxt, act, st)
        state = g_state.cpu().numpy() if g_state is not None else None
        return g_out.cpu().numpy(), state, self.summaries()

    @_on_device
    def enqueue_test(self, img, nxt, act, st):
        self._wait_d_update()
        self._join_g_stream()
        self._stage(img, nxt, act, st)
        self._run("test", self._body_test)
        self._have = {"g", "d"}
        return self.g_run.g_out, self.g_run.state

    def _body_test(self):
        img, nxt, act, st = self.in_img, self.in_next, self.in_act, self.in_state
        with self.real_branch:
            self.d_real.forward(img, nxt, act)
        g_out, g_state = self.g_run.forward(img, act)
        self.d_gen.forward(img, g_out, act)
        self.real_branch.join()
        self._g_losses(nxt, st, want_grad=False, with_adv_grad=False)
        if self.arg_loss == "bce":
            K.dlogit_loss(self.d_real.logits, self.d_real.n_logits, "bce", 0.9, 1.0, self.sc[1:2], None)
            K.dlogit_loss(self.d_gen.logits, self.d_gen.n_logits, "bce", 0.0, 1.0, self.sc[2:3], None)
        else:
            K.dlogit_loss(self.d_real.logits, self.d_real.n_logits, "wass", 1.0, 1.0, self.sc[1:2], None)
            K.dlogit_loss(self.d_gen.logits, self.d_gen.n_logits, "wass", -1.0, 1.0, self.sc[2:3], None)

    # ---- train.py:157-176, :286-299 ----------------------------------------------------------------------
    def test_sequence(self, input_images, test_next_frame, test_actions):
        """train.py:157-176: 6 recursive steps with action index 2j; returns (predicted [B,6,64,64,3],
        current_frame[1:7]).  The reference runs Trainer.test per step and throws its summaries away; here the rollout
        is ONE captured graph of generator-only forwards (rollout())."""
        pred = self.rollout(np.asarray(input_images)[:, 0], test_actions, steps=6, action_stride=2)
        out = pred.permute(1, 0, 2, 3, 4).contiguous().cpu().numpy()
        return out, pred[5][1:7].cpu().numpy()

    @_on_device
    def rollout(self, first_frame, actions, steps, action_stride=1):
        """Recursive prediction: frame_{j+1} = G(frame_j, [actions[:, j*stride, :5] | state_j]) with the
        generated frame AND the predicted state fed back on the device; state_0 = actions[:, 0, 5:].  The direct-pixel
        generator predicts no state and keeps the fed one (repair R5).  One CUDA graph per (steps, stride, T): per step
        one tiny kernel builds the action vector, the generator writes straight into predicted[j].
        Returns the device tensor predicted [steps, B, 64, 64, 3]."""
        B = self.B
        acts = self._dev(actions, (B, -1, E.ACTION_DIM))
        T = acts.shape[1]
        if (steps - 1) * action_stride >= T:
            raise ValueError("rollout: %d steps with action stride %d need more than %d action frames"
                             % (steps, action_stride, T))
        self._wait_d_update()
        self._join_g_stream()
        if self._pending_fetch is not None:
            torch.cuda.current_stream().wait_event(self._pending_fetch)
            self._pending_fetch = None
        key = ("rollout", steps, action_stride, T)
        if key not in self._rollout_bufs:
            self._rollout_bufs[key] = (torch.empty(B, T, E.ACTION_DIM, device=self.device),
                                       torch.empty(steps, B, E.IMG, E.IMG, 3, device=self.device))
        ro_acts, predicted = self._rollout_bufs[key]
        ro_acts.copy_(acts, non_blocking=True)
        f0 = self._as_tensor(first_frame)
        if f0.dtype == torch.uint8:
            f0 = f0.float() / 127.5 - 1.0
        self.in_img.copy_(f0.reshape(self.in_img.shape), non_blocking=True)

        def body():
            frame = self.in_img
            for j in range(steps):
                state = self.g_run.state if (j > 0 and self.arg_transform) else None
                K.rollout_actions(ro_acts, j * action_stride, state, self.in_act, E.STATE_DIM)
                self.g_run.forward(frame, self.in_act, out=predicted[j])
                frame = predicted[j]

        self._run(key, body)
        return predicted

    # ---- checkpoints (replaces tf.train.Saver, train.py:215,274 / test.py:29-30) -------------------------------
    def state_dict(self):
        """All variables keyed by their TF names (HWIO layouts) + optimizer slots + step counters."""
        self._wait_d_update()
        out = {}
        out.update(self.g_store.numpy())
        out.update(self.d_store.numpy())
        for oname, opt in (("g_opt", self.g_opt), ("g_pretrain_opt", self.g_pretrain_opt), ("d_opt", self.d_opt)):
            out[oname + "/t"] = np.array(opt.t, dtype=np.int64)
            for slot in ("m", "v", "ms"):
                if hasattr(opt, slot):
                    out[oname + "/" + slot] = getattr(opt, slot).cpu().numpy()
        return out

    def load_state_dict(self, sd):
        self._wait_d_update()
        self.g_store.load(sd)
        self.d_store.load(sd)
        for oname, opt in (("g_opt", self.g_opt), ("g_pretrain_opt", self.g_pretrain_opt), ("d_opt", self.d_opt)):
            if oname + "/t" in sd:
                opt.t = int(sd[oname + "/t"])
            for slot in ("m", "v", "ms"):
                if hasattr(opt, slot) and oname + "/" + slot in sd:
                    getattr(opt, slot).copy_(torch.from_numpy(np.asarray(sd[oname + "/" + slot])))

    def save(self, path):
        np.savez(path, **self.state_dict())

    def restore(self, path):
        with np.load(path) as f:
            self.load_state_dict({k: f[k] for k in f.files})
