"""Batch-norm / activation kernels at every layer shape of the B=256 step, timed as CUDA-graph replays of 10 launches
(no host launch cost; small layers are L2-warm, as in the step where their inputs were just produced)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from action_conditioned_gans_b200 import kernels as K
dev = torch.device("cuda:0")
B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
shapes = [("g/conv1", B * 1024, 32, "relu"), ("d/conv1", B * 1024, 64, "lrelu"), ("g/conv2", B * 256, 64, "relu"),
          ("d/conv2,g/tconv2", B * 256, 128, "lrelu"), ("conv3,tconv1", B * 64, 128, "relu"), ("conv4", B * 16, 256, "relu"),
          ("d/conv5", B * 4, 512, "lrelu"), ("g/tconv3", B * 1024, 128, "relu"), ("g/sconv3", B * 64, 32, "relu"),
          ("g/sconv4", B * 16, 16, "relu")]
def graph_time(fn, n=10, reps=5):
    fn(); torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(n): fn()
    g.replay(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): g.replay()
    e1.record(); e1.synchronize()
    return e0.elapsed_time(e1) / (n * reps) * 1e3
print("%-18s %9s %5s | %-22s | %-22s | %-22s" % ("layer", "rows", "C", "fwd us (GB/s)", "bwd reduce us (GB/s)", "bwd apply us (GB/s)"))
tot = [0, 0, 0]
for name, rows, C, act in shapes:
    z = torch.randn(rows, C, device=dev).to(torch.bfloat16)
    dA = torch.randn(rows, C, device=dev).to(torch.bfloat16)
    a = torch.empty_like(z); dz = torch.empty_like(z)
    f64 = torch.zeros(4 * C, dtype=torch.float64, device=dev)
    mean, rstd, scale, shift = (torch.randn(C, device=dev).abs() + 0.5 for _ in range(4))
    dbeta = torch.zeros(C, device=dev)
    mb = rows * C * 2 / 1e6
    t0 = graph_time(lambda: K.bn_act_fwd(z, rows, C, C, 1, scale, shift, act, a, C))
    t1 = graph_time(lambda: K.bn_act_bwd_reduce(dA, None, C, z, C, rows, C, 1, mean, rstd, shift, act, f64[2 * C:]))
    t2 = graph_time(lambda: K.bn_act_bwd_apply(dA, None, C, z, C, rows, C, 1, mean, rstd, shift, act, True, f64[2 * C:], dz, dbeta, ld_dz=C))
    for i, t in enumerate((t0, t1, t2)): tot[i] += t
    print("%-18s %9d %5d | %7.1f (%6.0f)       | %7.1f (%6.0f)       | %7.1f (%6.0f)" %
          (name, rows, C, t0, 2 * mb / t0 * 1e3, t1, 2 * mb / t1 * 1e3, t2, 3 * mb / t2 * 1e3))
print("sum", ["%.1f" % t for t in tot])
