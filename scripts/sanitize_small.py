"""Every kernel family once at small shapes, for `compute-sanitizer --tool memcheck|racecheck`:

    compute-sanitizer --tool memcheck python scripts/sanitize_small.py

Covers the ~30 hand-rolled mbarrier / TMA / tcgen05 pipelines: DNA forward / backward (K = 5, 6), the generic tcgen05
conv kernel in both gather forms (3- and 6-stage variants, split-K), the persistent small-K kernel, the halo kernel in
both forms and both accumulator counts, the weight-gradient kernel, weight packs, batch-norm / activation kernels,
losses, optimizers, the feeder gather and one whole training iteration (eager and graph-replayed) at batch 4.
Prints `sanitize_small: done` at the end; the sanitizer's own summary follows.
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402

from action_conditioned_gans_b200 import kernels as K  # noqa: E402
from action_conditioned_gans_b200.feeder import DeviceFeeder  # noqa: E402
from action_conditioned_gans_b200.trainer import Trainer  # noqa: E402

dev = torch.device("cuda:0")


def ru(v, m):
    return (v + m - 1) // m * m


def conv_family(B, H, W, Cin, Cout, k, tag):
    shape = K.conv_shape(B, H, W, Cin, Cout, k, 2, "SAME")
    ldi, ldo = (8 if Cin <= 8 else ru(Cin, 16)), ru(Cout, 16)
    if Cout >= 64:
        ldo = ru(Cout, 64)
    x = torch.zeros(B, H, W, ldi, dtype=torch.bfloat16, device=dev)
    x[..., :Cin] = torch.randn(B, H, W, Cin, device=dev)
    dy = torch.zeros(B, shape.OH, shape.OW, ldo, dtype=torch.bfloat16, device=dev)
    dy[..., :Cout] = torch.randn(B, shape.OH, shape.OW, Cout, device=dev)
    w = torch.randn(k, k, Cin, Cout, device=dev) / (k * Cin ** 0.5)
    pf = torch.empty(K.pack_size(shape, 0, ldi), dtype=torch.bfloat16, device=dev)
    pb = torch.empty(K.pack_size(shape, 1, ldo), dtype=torch.bfloat16, device=dev)
    K.pack_weights(shape, w, 0, ldi, pf)
    K.pack_weights(shape, w, 1, ldo, pb)
    y = torch.empty(B, shape.OH, shape.OW, ldo, dtype=torch.bfloat16, device=dev)
    dx = torch.empty(B, H, W, ldi, dtype=torch.bfloat16, device=dev)
    dw = torch.zeros(k, k, Cin, Cout, device=dev)
    st = torch.zeros(2 * Cout, dtype=torch.float64, device=dev)
    K.conv_fprop_tc(shape, x, pf, y, ldi, ldo, stats=st, splitk=K.splitk_workspace(shape, 0, ldi, dev))
    K.conv_dgrad_tc(shape, dy, pb, dx, ldo, ldi, splitk=K.splitk_workspace(shape, 1, ldo, dev))
    K.conv_wgrad_tc(shape, x, dy, dw, ldi, ldo)
    torch.cuda.synchronize()
    print("  conv family ok:", tag, flush=True)


def main():
    for Kk in (5, 6):
        lg = torch.randn(2, 64, 64, Kk * Kk, device=dev)
        im = torch.rand(2, 64, 64, 3, device=dev) * 2 - 1
        dy = torch.randn(2, 64, 64, 3, device=dev)
        o, dl = torch.empty_like(im), torch.empty_like(lg)
        K.dna_fwd(lg, im, o, Kk)
        K.dna_bwd(lg, im, dy, dl, Kk)
    torch.cuda.synchronize()
    print("  dna ok", flush=True)
    conv_family(2, 16, 16, 64, 128, 5, "generic 6-stage (8x8 grid)")
    conv_family(5, 8, 8, 128, 256, 5, "generic + split-K (4x4 grid)")
    conv_family(2, 32, 32, 64, 128, 5, "halo: CONV NACC=2 / ADJ N=64, two images per tile")
    conv_family(2, 64, 64, 36, 128, 5, "halo: CONV 48-channel rows / ADJ N=48, 32-wide tiles")
    conv_family(2, 32, 32, 128, 128, 5, "halo: ADJ N=128 NACC=2")
    conv_family(40, 64, 64, 6, 64, 5, "small-K persistent forward (320 tiles) / halo ADJ N=16")
    conv_family(130, 8, 8, 128, 256, 5, "pixel-major kernel: 4x4 / 8x8 grids, ragged batch tile, two N tiles")
    conv_family(64, 4, 4, 256, 512, 5, "pixel-major kernel + split-K (2x2 grid)")
    torch.cuda.synchronize()
    # feeder + one whole iteration, eager then captured + replayed (all elementwise / loss / optimizer kernels)
    rng = np.random.RandomState(0)
    frames = rng.randint(0, 256, size=(6, 7, 64, 64, 3)).astype(np.uint8)
    acts = rng.randn(6, 7, 10).astype(np.float32)
    fd = DeviceFeeder(frames, acts, dev)
    for flags in ((True, "bce", "adam", True), (True, "wass", "rmsprop", False)):
        trn = Trainer(None, *flags, batch_size=4, device=dev)
        for it in range(3):
            s, t = fd.sample(4, rng)
            trn.train_d_indexed(fd, s, t, summarize=(it == 2))
            trn.train_g_indexed(fd, s, t)
        trn.rollout(frames[:4, 0], acts[:4], steps=6, action_stride=1)
        trn.synchronize()
        print("  trainer ok:", flags, flush=True)
        del trn
    print("sanitize_small: done", flush=True)


if __name__ == "__main__":
    main()
