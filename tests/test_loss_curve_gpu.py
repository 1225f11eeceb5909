"""North star: "loss curves over 200 steps tracking the reference".  The CUDA product path (bf16 tcgen05 kernels, CUDA
graphs) replays the 200-iteration stream of tests/golden/make_loss_curve.py and is compared with the fp64 oracle's
curve stored in tests/golden/loss_curve_v1.npz.  GAN training is chaotic, so after the first iterations the two runs
are compared as CURVES (windowed means, overall relative deviation), not element by element."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

HERE = os.path.dirname(os.path.abspath(__file__))


def _load_generator():
    import importlib.util
    spec = importlib.util.spec_from_file_location("make_loss_curve", os.path.join(HERE, "golden", "make_loss_curve.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def run_curve(cuda, precision="bf16"):
    from action_conditioned_gans_b200.trainer import Trainer
    gen = _load_generator()
    trn = Trainer(None, True, "bce", "adam", True, batch_size=gen.BATCH, ksize=gen.KSIZE, params=gen.params(),
                  precision=precision, device=cuda)
    d_curve = {k: [] for k in gen.KEYS}
    g_curve = {k: [] for k in ("g_loss", "g_l2_loss", "g_adv_loss", "g_psnr")}
    for it in range(gen.STEPS):
        img, nxt, act, state = gen.feeds(it)
        s = trn.train_d(img, nxt, act, summarize=True)
        for k in gen.KEYS:
            d_curve[k].append(s[k])
        trn.train_g(img, nxt, act, state)
        sg = trn.summaries()
        for k in g_curve:
            g_curve[k].append(sg[k])
    out = {"d/" + k: np.array(v) for k, v in d_curve.items()}
    out.update({"g/" + k: np.array(v) for k, v in g_curve.items()})
    return out


def curve_stats(ours, ref):
    """max relative error over the first 5 iterations, mean relative deviation over all, and of 20-step window means"""
    st = {}
    for k in ref:
        a, b = ours[k], ref[k]
        den = np.maximum(1.0, np.abs(b))
        win = lambda x: x.reshape(-1, 20).mean(1)
        st[k] = (float((np.abs(a - b) / den)[:5].max()), float((np.abs(a - b) / den).mean()),
                 float((np.abs(win(a) - win(b)) / np.maximum(1.0, np.abs(win(b)))).max()))
    return st


def test_loss_curves_track_the_oracle_over_200_steps(cuda):
    ref = dict(np.load(os.path.join(HERE, "golden", "loss_curve_v1.npz")))
    ours = run_curve(cuda)
    for k, v in ours.items():
        assert np.isfinite(v).all(), k
    st = curve_stats(ours, ref)
    for k, (first5, mean_dev, win_dev) in st.items():
        tol5, tolm, tolw = LIMITS["psnr" if k.endswith("psnr") else "loss"]
        assert first5 <= tol5, (k, "first 5 iterations", first5)
        assert mean_dev <= tolm, (k, "mean deviation over 200 iterations", mean_dev)
        assert win_dev <= tolw, (k, "20-iteration window means", win_dev)


# (first-5 pointwise, mean over 200, worst 20-step window), relative to max(1, |oracle|).  Measured on B200
# (scripts/loss_curve_probe.py): bf16 path first-5 <= 0.16e-2, mean <= 1.06e-2 (discriminator loss; the frame losses
# stay within 0.2e-2), window means <= 0.9e-2; the fp32 engine itself sits at mean 0.97e-2 on the discriminator loss,
# i.e. that part is the chaos of the adversarial game, not precision.  Limits = ~3x the measured values.
LIMITS = {"loss": (1e-2, 3e-2, 3e-2), "psnr": (2e-2, 3e-2, 3e-2)}
