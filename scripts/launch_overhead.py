import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from action_conditioned_gans_b200.trainer import Trainer
from action_conditioned_gans_b200 import _lib
dev = torch.device("cuda:0")
for B in (32, 64, 128, 256):
    trn = Trainer(None, True, "bce", "adam", True, batch_size=B, ksize=6, device=dev)
    img, nxt, act, state = [t.to(dev) for t in bench.synth_batch(B, 1, False)]
    for _ in range(3):
        trn.enqueue_train_d(img, nxt, act); trn.enqueue_train_g(img, nxt, act, state)
    torch.cuda.synchronize()
    n = 5
    l0 = _lib.launch_count()
    t0 = time.perf_counter()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        trn.enqueue_train_d(img, nxt, act); trn.enqueue_train_g(img, nxt, act, state)
    t_cpu = time.perf_counter() - t0
    e1.record(); torch.cuda.synchronize()
    print("B=%d: cpu enqueue %.2f ms/iter, gpu %.2f ms/iter, launches/iter %d" % (B, 1e3 * t_cpu / n, e0.elapsed_time(e1) / n, (_lib.launch_count() - l0) // n))
    del trn
