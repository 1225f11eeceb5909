"""Prints how closely the CUDA path's 200-step loss curves follow the oracle's (tests/golden/loss_curve_v1.npz)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
import test_loss_curve_gpu as T
ref = dict(np.load(os.path.join(ROOT, "tests", "golden", "loss_curve_v1.npz")))
for prec in ("bf16", "fp32"):
    ours = T.run_curve(torch.device("cuda:0"), prec)
    print(prec)
    for k, v in T.curve_stats(ours, ref).items():
        print("  %-30s first5 %.4f  mean %.4f  window %.4f   | ours end %.4f ref end %.4f" % (
            k, v[0], v[1], v[2], ours[k][-20:].mean(), ref[k][-20:].mean()))
