"""ms per training iteration (train_d + train_g graphs, feeds resident) for the current environment switches.
    python scripts/step_time.py [batch] [iters]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from action_conditioned_gans_b200.trainer import Trainer

B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 20
dev = torch.device("cuda:0")
img = torch.rand(B, 64, 64, 3, device=dev) * 2 - 1
nxt = (img + 0.1 * torch.randn_like(img)).clamp(-1, 1)
act = torch.randn(B, 10, device=dev)
state = torch.randn(B, 5, device=dev)
trn = Trainer(None, True, "bce", "adam", True, batch_size=B, ksize=6, device=dev, seed=7)
for _ in range(4):
    trn.enqueue_train_d(img, nxt, act)
    trn.enqueue_train_g(img, nxt, act, state)
torch.cuda.synchronize()
best = 1e9
for rep in range(3):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        trn.enqueue_train_d(img, nxt, act)
        trn.enqueue_train_g(img, nxt, act, state)
    e1.record()
    torch.cuda.synchronize()
    best = min(best, e0.elapsed_time(e1) / iters)
print("%s  %.3f ms / iteration  (%.0f frames/s)" % (" ".join("%s=%s" % kv for kv in sorted(os.environ.items())
                                                          if kv[0].startswith("ACG_")), best, B / best * 1e3))
