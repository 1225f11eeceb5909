"""Data-parallel host logic on CPU (gloo, world_size 2): the collectives the step needs -- SyncBN moment
all-reduce, gradient-bucket all-reduce, loss-sum reduction -- give the single-process (global batch) answer."""
import os

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import np_ref


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from action_conditioned_gans_b200.trainer import DataParallel
    dp = DataParallel()
    rng = np.random.RandomState(0)
    z = rng.randn(8, 4, 4, 6)                     # global batch of 8, sharded 4 + 4
    beta = rng.randn(6)
    shard = z[rank * 4:(rank + 1) * 4].reshape(-1, 6)
    stats = torch.tensor(np.concatenate([shard.sum(0), (shard ** 2).sum(0)]))
    dp.allreduce_sum(stats)                       # what NetRun.layer_fwd does with st.stats
    rows = 8 * 16
    mean = stats[:6].numpy() / rows
    var = stats[6:].numpy() / rows - mean ** 2
    y_shard = (z[rank * 4:(rank + 1) * 4] - mean) / np.sqrt(var + 1e-3) + beta
    ref = np_ref.batch_norm(z, beta)[rank * 4:(rank + 1) * 4]
    ok_bn = np.abs(y_shard - ref).max() < 1e-12
    # gradient bucket: sum of per-rank partial gradients of a loss normalised by the GLOBAL batch
    g0 = rng.randn(10)                            # same draw on both ranks
    g = torch.tensor(g0 + rank)
    dp.allreduce_sum(g)
    ok_grad = np.allclose(g.numpy(), 2 * g0 + 1)
    # the state loss is a Frobenius norm over the global batch: combine squared local norms (Trainer._scalars)
    s = rng.randn(8, 5)
    loc = torch.tensor([float((s[rank * 4:(rank + 1) * 4] ** 2).sum())], dtype=torch.float64)
    dp.allreduce_sum(loc)
    ok_norm = abs(float(loc.sqrt()) - np.linalg.norm(s)) < 1e-12
    q.put((rank, dp.world, bool(ok_bn), bool(ok_grad), bool(ok_norm)))
    dist.destroy_process_group()


def test_dp_collectives_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, world, ok_bn, ok_grad, ok_norm in res:
        assert world == 2 and ok_bn and ok_grad and ok_norm, (rank, ok_bn, ok_grad, ok_norm)
