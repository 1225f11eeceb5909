"""Data-parallel iteration time under torchrun, with timing-only switches to attribute the overhead:
  ACG_DP_SKIP=grad   no gradient-bucket all-reduce      ACG_DP_SKIP=bn   no batch-norm exchange (local statistics)
    python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 scripts/dp_step_time.py"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist

local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
sys.stdout.flush()
saved = os.dup(1)
os.dup2(2, 1)
dist.init_process_group("nccl", device_id=dev)
dist.barrier()
torch.cuda.synchronize()
os.dup2(saved, 1)
from action_conditioned_gans_b200 import engine as E
from action_conditioned_gans_b200 import kernels as K
from action_conditioned_gans_b200.trainer import DataParallel, Trainer

# timing-only switches live HERE (monkeypatches), not in the product: results are wrong with them
SKIP = set(filter(None, os.environ.get("ACG_DP_SKIP", "").split(",")))
if "grad" in SKIP:
    Trainer._sync_grads = lambda self, store: None
if "bn" in SKIP:
    def _local_moments(self, st, beta):
        L = st.spec
        K.bn_finalize(st.stats, beta, st.rows, L.cout, 1, st.mean, st.rstd, st.scale, st.shift, E.BN_EPS)
    E.NetRun._sync_moments = _local_moments

B = 256
dp = DataParallel(device=dev)
trn = Trainer(None, True, "bce", "adam", True, batch_size=B, ksize=6, device=dev, seed=7, dp=dp)
g = torch.Generator(device=dev).manual_seed(dist.get_rank())
img = torch.rand(B, 64, 64, 3, device=dev, generator=g) * 2 - 1
nxt = (img + 0.1 * torch.randn(B, 64, 64, 3, device=dev, generator=g)).clamp(-1, 1)
act = torch.randn(B, 10, device=dev, generator=g)
state = torch.randn(B, 5, device=dev, generator=g)
for _ in range(6):
    trn.enqueue_train_d(img, nxt, act)
    trn.enqueue_train_g(img, nxt, act, state)
dist.barrier()
torch.cuda.synchronize()
best = 1e9
for rep in range(3):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    dist.barrier()
    torch.cuda.synchronize()
    e0.record()
    for _ in range(20):
        trn.enqueue_train_d(img, nxt, act)
        trn.enqueue_train_g(img, nxt, act, state)
    e1.record()
    torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) / 20], device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    best = min(best, float(t.item()))
if dist.get_rank() == 0:
    print("world %d  ACG_DP_SKIP=%s ACG_DP_SYNC=%s  %.3f ms / iteration  (%.0f frames/s)" % (
        dist.get_world_size(), os.environ.get("ACG_DP_SKIP", ""), os.environ.get("ACG_DP_SYNC", "peer"), best,
        B * dist.get_world_size() / best * 1e3), flush=True)
torch.cuda.synchronize()
os._exit(0)
