"""Where does the captured step spend its time?  Replays the train_d + train_g graphs with one kernel family (or all
kernels of one layer) turned into a no-op and prints the change of the iteration time.  The numbers are garbage
numerically (that is the point: timing of these kernels does not depend on the data); with parallel branches in the
graph the delta is the family's contribution to the CRITICAL PATH, which a per-kernel sum cannot show.

    python scripts/ablate.py [batch]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from action_conditioned_gans_b200 import engine as E
from action_conditioned_gans_b200 import kernels as Kn
from action_conditioned_gans_b200.trainer import Trainer

B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
dev = torch.device("cuda:0")
rng = np.random.RandomState(0)
img = torch.from_numpy(rng.uniform(-1, 1, (B, 64, 64, 3)).astype(np.float32)).to(dev)
nxt = (img + 0.1 * torch.randn_like(img)).clamp(-1, 1)
act = torch.randn(B, 10, device=dev)
state = torch.randn(B, 5, device=dev)

orig_call = Kn.call
orig_fwd, orig_bwd = E.NetRun.layer_fwd, E.NetRun.layer_bwd


def measure(skip_calls=(), skip_layers=(), branches=True, iters=10):
    skip_calls, skip_layers = set(skip_calls), set(skip_layers)

    def call(name, *a):
        if name in skip_calls:
            return
        orig_call(name, *a)

    def layer_fwd(self, name, x, out, ld_out, cat=None):
        if name in skip_layers:
            self.layers[name].x = x
            return
        return orig_fwd(self, name, x, out, ld_out, cat=cat)

    def layer_bwd(self, name, dA, ld_d, **kw):
        if name in skip_layers:
            return self.layers[name].dx
        return orig_bwd(self, name, dA, ld_d, **kw)

    Kn.call, E.NetRun.layer_fwd, E.NetRun.layer_bwd = call, layer_fwd, layer_bwd
    try:
        trn = Trainer(None, True, "bce", "adam", True, batch_size=B, ksize=6, device=dev, seed=7, branches=branches)
        for _ in range(4):
            trn.enqueue_train_d(img, nxt, act)
            trn.enqueue_train_g(img, nxt, act, state)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(iters):
            trn.enqueue_train_d(img, nxt, act)
            trn.enqueue_train_g(img, nxt, act, state)
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / iters
    finally:
        Kn.call, E.NetRun.layer_fwd, E.NetRun.layer_bwd = orig_call, orig_fwd, orig_bwd
        del trn
        torch.cuda.empty_cache()


base = measure()
print("baseline (branches)      %.3f ms" % base)
print("baseline (one stream)    %.3f ms" % measure(branches=False))
fams = ["acg_conv_fprop_tc", "acg_conv_dgrad_tc", "acg_conv_wgrad_tc", "acg_bn_finalize_act_fwd", "acg_bn_act_fwd", "acg_bn_act_bwd_reduce",
        "acg_bn_act_bwd_apply", "acg_pack_weights_batched", "acg_copy_channels", "acg_tile_actions", "acg_dna_fwd",
        "acg_dna_bwd", "acg_frame_losses", "acg_adam_step"]
for f in fams:
    t = measure(skip_calls=[f])
    print("without %-28s %.3f ms  (-%.3f)" % (f, t, base - t))
t = measure(skip_calls=[f for f in fams if "conv" in f])
print("without all conv kernels             %.3f ms  (-%.3f)" % (t, base - t))
t = measure(skip_calls=[f for f in fams if "conv" not in f])
print("conv kernels only                    %.3f ms" % t)
layers = [L.name for L in E.g_dna_spec(6)] + [L.name for L in E.d_spec()]
for n in layers:
    t = measure(skip_layers=[n])
    print("without layer %-10s             %.3f ms  (-%.3f)" % (n, t, base - t))
