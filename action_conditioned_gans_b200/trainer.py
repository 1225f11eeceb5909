"""`Trainer`: the step wrappers of train.py:27-176 on top of the acg_b200 engine.

Same constructor flags and methods as the reference (`Trainer(sess, arg_adv, arg_loss, arg_opt, arg_transform)`,
`pretrain_g`, `train_g`, `train_d`, `test`, `test_sequence`); `sess` is accepted and ignored (there is no TF
session -- each method enqueues one device step).  Repairs R1-R6 of SURVEY.md section 0 are applied; the D step
is update-THEN-clip (the reference leaves the order of train.py:140,143 unspecified).
"""
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
        self._frames_host = None
        self._pending_fetch = None
        # the discriminator update (gradient all-reduce, optimizer, weight packs) of train_d runs on its own stream so
        # that the generator forward of the following train_g overlaps it; everything that reads D weights waits
        self._d_update_stream = None
        self._d_update_done = None
        self._graphs, self._calls, self._graph_launches = {}, {}, {}
        self.replayed_launches = 0      # kernels launched through graph replays (acg_launch_count sees eager ones)

    def _stage(self, img, nxt, act, st):
        """Copy the feeds into the static buffers the kernels (and the captured graphs) read.

        Host feeds go through a COPY STREAM into one of two device staging sets, so the host->device transfer of this
        call overlaps the device work of the previous call (train_d's graph is still running when train_g's feeds
        arrive); the compute stream then only does a device->device copy into the static buffers."""
        if self._pending_fetch is not None:
            torch.cuda.current_stream().wait_event(self._pending_fetch)
            self._pending_fetch = None
        feeds = [(self.in_img, img), (self.in_next, nxt), (self.in_act, act)]
        if st is not None:
            feeds.append((self.in_state, st))
        else:
            self.in_state.zero_()
        if all(t.is_cuda for _, t in feeds):
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
        k = self._staging_idx
        self._staging_idx ^= 1
        main, cs = torch.cuda.current_stream(), self._copy_stream
        cs.wait_event(self._staging_free[k])          # the compute stream is done with this set (two calls ago)
        ready = torch.cuda.Event()
        with torch.cuda.stream(cs):
            for i, (dst, t) in enumerate(feeds):
                self._staging[k][i].copy_(t.reshape(dst.shape), non_blocking=True)
            ready.record(cs)
        main.wait_event(ready)
        for i, (dst, t) in enumerate(feeds):
            dst.copy_(self._staging[k][i], non_blocking=True)
        self._staging_free[k].record(main)

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
            return a if a.dtype == torch.float32 else a.float()
        return torch.from_numpy(np.ascontiguousarray(np.asarray(a, dtype=np.float32)))

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

    def enqueue_pretrain_g(self, img, nxt, act, st):
        self._wait_d_update()
        self._stage(img, nxt, act, st)
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

    def enqueue_train_g(self, img, nxt, act, st, fetch=False):
        """The step is two captured graphs: (a) generator forward, (b) everything else.  With fetch=True the frames
        leave for the host on the copy stream as soon as (a) is done, overlapping (b)."""
        self._stage(img, nxt, act, st)
        self.g_opt.tick()
        self._run("train_g_a", self._body_train_g_a)
        done = None
        if fetch:
            if self._frames_host is None:
                self._frames_host = [torch.empty(self.g_run.g_out.shape, dtype=torch.float32).pin_memory()
                                     for _ in range(4)]
                self._frames_idx = 0
                if self._copy_stream is None:
                    self._copy_stream = torch.cuda.Stream(device=self.device)
            self._frames_idx = (self._frames_idx + 1) % 4
            main, cs = torch.cuda.current_stream(), self._copy_stream
            fwd_done = torch.cuda.Event()
            fwd_done.record(main)
            cs.wait_event(fwd_done)
            done = torch.cuda.Event()
            with torch.cuda.stream(cs):
                self._frames_host[self._frames_idx].copy_(self.g_run.g_out, non_blocking=True)
                done.record(cs)
            self._pending_fetch = done     # the NEXT step's forward overwrites g_out: it waits for this copy
        self._wait_d_update()              # D(generated) reads the discriminator weights / packs
        self._run("train_g_b", self._body_train_g_b)
        return done

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
        if summarize:
            # merged_summaries also holds the generator scalars (train.py:112,140)
            self._g_losses(self.in_next, self.in_state, want_grad=False, with_adv_grad=False)
            self._have = {"g", "d"}
            return self.summaries()
        return None

    def enqueue_train_d(self, img, nxt, act, need_state=False):
        """need_state: also run the generator's state head (only the summaries of train.py:140 read it).
        Two graphs: (main) forward + backward of D on both pairs, (update) gradient all-reduce + optimizer + clip +
        weight packs.  The update is replayed on a side stream: the next train_g's generator forward does not read D
        weights and overlaps it (with data parallelism that hides the 19 MB gradient all-reduce)."""
        self._wait_d_update()
        self._stage(img, nxt, act, None)
        self.d_opt.tick()
        if need_state:
            self._run("train_d_state", lambda: self._body_train_d(True))
        else:
            self._run("train_d", lambda: self._body_train_d(False))
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

    def _body_train_d(self, need_state=False):
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

    def enqueue_test(self, img, nxt, act, st):
        self._wait_d_update()
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

    # ---- train.py:157-176 -------------------------------------------------------------------------------
    def test_sequence(self, input_images, test_next_frame, test_actions):
        """6-step recursive rollout; frames and the predicted state stay on the device between steps
        (the reference round-trips host<->device every step)."""
        B = self.B
        seq = self._dev(input_images, (B, -1, E.IMG, E.IMG, 3))
        nxt = self._dev(test_next_frame, (B, -1, E.IMG, E.IMG, 3))
        acts = self._dev(test_actions, (B, -1, E.ACTION_DIM))
        predicted = torch.empty(6, B, E.IMG, E.IMG, 3, device=self.device)
        current_frame = seq[:, 0].contiguous()
        current_state = acts[:, 0, 5:].contiguous()
        for j in range(6):
            acs = torch.cat((acts[:, j * 2, :5], current_state), dim=1).contiguous()   # train.py:163
            g_out, g_state = self.enqueue_test(current_frame, nxt[:, j * 2].contiguous(), acs, self.zero_state)
            predicted[j].copy_(g_out)
            current_frame = predicted[j]
            if g_state is not None:                    # R5: the direct generator keeps the fed state
                current_state = g_state.clone()
        pred = predicted.permute(1, 0, 2, 3, 4).contiguous().cpu().numpy()
        return pred, current_frame[1:7].cpu().numpy()

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
