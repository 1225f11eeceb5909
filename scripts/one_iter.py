"""One eager training iteration (train_d + train_g, B=256, one stream) inside a cudaProfiler range, for

    ncu --profile-from-start off --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum \
        --clock-control none --csv --log-file gpurun_out/r2_iter_launches.csv python scripts/one_iter.py

Every launch of the iteration with its device time and DRAM traffic (cold-cache, serialised: compare SHARES); the
summary (scripts/summarize_iter.py) becomes profiles/traffic.json, which bench.py reports as `roofline.traffic`.
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from action_conditioned_gans_b200 import engine as E  # noqa: E402
from action_conditioned_gans_b200.trainer import Trainer  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
dev = torch.device("cuda:0")
g = torch.Generator(device=dev).manual_seed(0)
img = torch.rand(B, 64, 64, 3, device=dev, generator=g) * 2 - 1
nxt = (img + 0.1 * torch.randn(B, 64, 64, 3, device=dev, generator=g)).clamp(-1, 1)
act = torch.randn(B, 10, device=dev, generator=g)
state = torch.randn(B, 5, device=dev, generator=g)
trn = Trainer(None, True, "bce", "adam", True, batch_size=B, ksize=6, device=dev, seed=7, use_graphs=False)
E.Branch.enabled = False
for _ in range(2):
    trn.enqueue_train_d(img, nxt, act)
    trn.enqueue_train_g(img, nxt, act, state)
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStart()
trn.enqueue_train_d(img, nxt, act)
trn.enqueue_train_g(img, nxt, act, state)
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStop()
print("one_iter: done")
