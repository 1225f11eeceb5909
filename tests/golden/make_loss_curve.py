"""Generates tests/golden/loss_curve_v1.npz: per-step losses of the CPU oracle (fp64) over 200 iterations of the
adversarial DNA schedule (train.py:241-263: one train_d with summaries, one train_g per iteration; --loss bce
--opt adam --dna, ksize 6, batch 8) on a reproducible synthetic stream.  Like golden_v1.npz these are ORACLE outputs
(the reference ships no tests and TensorFlow 1.0 cannot run here), generated once in this container:

    python tests/golden/make_loss_curve.py          (~2 minutes on 8 host threads)

tests/test_loss_curve_gpu.py replays the same stream through the CUDA path (north star: "loss curves over 200 steps
tracking the reference")."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import np_ref, torch_ref  # noqa: E402

STEPS, BATCH, KSIZE = 200, 8, 6
KEYS = ["discriminator_direct_loss", "discriminator_gen_loss", "discriminator_loss", "g_loss", "g_l2_loss",
        "g_adv_loss", "g_psnr"]


def params(seed=7):
    rng = np.random.RandomState(seed)
    p = np_ref.init_params(np_ref.g_dna_spec(KSIZE), rng)
    p.update(np_ref.init_params(np_ref.d_spec(), rng))
    return p


def feeds(it):
    """Push-shaped stream with structure to learn: the next frame is the frame shifted by an action-dependent
    offset plus noise; the state target is a linear function of the action."""
    rng = np.random.RandomState(5000 + it)
    base = rng.uniform(-1, 1, (BATCH, 16, 16, 3))
    img = np.repeat(np.repeat(base, 4, axis=1), 4, axis=2)                      # 64x64, blocky
    act = rng.randn(BATCH, 10)
    shift = np.clip(np.round(act[:, 0] * 2), -3, 3).astype(int)
    nxt = np.stack([np.roll(img[b], shift[b], axis=1) for b in range(BATCH)])
    nxt = np.clip(nxt + 0.05 * rng.randn(*nxt.shape), -1, 1)
    state = 0.5 * act[:, :5] + 0.1 * rng.randn(BATCH, 5)
    return [a.astype(np.float32) for a in (img, nxt, act, state)]


def main():
    ora = torch_ref.Trainer(params(), True, "bce", "adam", True, ksize=KSIZE)
    d_curve = {k: [] for k in KEYS}
    g_curve = {k: [] for k in ("g_loss", "g_l2_loss", "g_adv_loss", "g_psnr")}
    for it in range(STEPS):
        img, nxt, act, state = feeds(it)
        s = ora.train_d(img, nxt, act, summarize=True)
        for k in KEYS:
            d_curve[k].append(s[k])
        ora.train_g(img, nxt, act, state)
        sg = ora.summaries()
        for k in g_curve:
            g_curve[k].append(sg[k])
        if it % 20 == 0:
            print(it, {k: round(v, 4) for k, v in s.items()}, flush=True)
    out = {"d/" + k: np.array(v) for k, v in d_curve.items()}
    out.update({"g/" + k: np.array(v) for k, v in g_curve.items()})
    np.savez(os.path.join(os.path.dirname(os.path.abspath(__file__)), "loss_curve_v1.npz"), **out)


if __name__ == "__main__":
    main()
