"""Elementwise (batch-norm / activation) kernels at the largest layer's size (g/tconv3 at B=256: 262144 rows x 128)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from action_conditioned_gans_b200 import kernels as K
dev = torch.device("cuda:0")
rows, C = 256 * 32 * 32, 128
z = torch.randn(rows, C, device=dev).to(torch.bfloat16)
dA = torch.randn(rows, C, device=dev).to(torch.bfloat16)
a = torch.empty_like(z); dz = torch.empty_like(z)
f64 = torch.zeros(4 * C, dtype=torch.float64, device=dev)
mean, rstd, scale, shift = (torch.randn(C, device=dev).abs() + 0.5 for _ in range(4))
dbeta = torch.zeros(C, device=dev)
def t(fn, n=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); e1.synchronize()
    return e0.elapsed_time(e1) / n * 1e3
mb = rows * C * 2 / 1e6
for name, fn, nbytes in (
    ("bn_stats", lambda: K.bn_stats(z, rows, C, C, 1, f64[:2 * C]), mb),
    ("bn_act_fwd", lambda: K.bn_act_fwd(z, rows, C, C, 1, scale, shift, "relu", a, C), 2 * mb),
    ("bwd_reduce", lambda: K.bn_act_bwd_reduce(dA, None, C, z, C, rows, C, 1, mean, rstd, shift, "relu", f64[2 * C:]), 2 * mb),
    ("bwd_apply", lambda: K.bn_act_bwd_apply(dA, None, C, z, C, rows, C, 1, mean, rstd, shift, "relu", True, f64[2 * C:], dz, dbeta, ld_dz=C), 3 * mb)):
    us = t(fn)
    print("%-12s %7.1f us  %6.0f MB  %6.0f GB/s" % (name, us, nbytes, nbytes / us * 1e3))
