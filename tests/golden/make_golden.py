"""Generates tests/golden/golden_v1.npz from the CPU oracle (fp64).

The reference ships NO tests or golden vectors and TensorFlow 1.0 cannot be installed here,
so these known-answer vectors are ORACLE outputs ("parity unpinned" by the reference itself): they pin the oracle
against regressions and give the CUDA path a fixture that does not depend on the oracle code at test time.
Run from the repo root:  python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import np_ref, torch_ref  # noqa: E402


def main():
    rng = np.random.RandomState(20260101)
    out = {}
    # (i) DNA forward / backward, K=5 and K=6, border pixels included (H=W=8 is all border for K=6)
    for K in (5, 6):
        lg = rng.randn(2, 8, 8, K * K) * 2
        img = rng.uniform(-1, 1, (2, 8, 8, 3))
        dy = rng.randn(2, 8, 8, 3)
        out["dna%d_logits" % K], out["dna%d_img" % K], out["dna%d_dy" % K] = lg, img, dy
        out["dna%d_out" % K] = np_ref.dna_forward(lg, img, K)
        out["dna%d_dlogits" % K] = np_ref.dna_backward(lg, img, dy, K)
    # (ii) conv / conv_transpose incl. the asymmetric SAME padding
    x = rng.randn(2, 8, 8, 6)
    w = rng.randn(5, 5, 6, 8) * 0.1
    out["conv_x"], out["conv_w"] = x, w
    out["conv_y"] = np_ref.conv2d(x, w, 2, "SAME")
    xt = rng.randn(2, 4, 4, 8)
    wt = rng.randn(5, 5, 6, 8) * 0.1          # [kh,kw,Cout,Cin]
    out["deconv_x"], out["deconv_w"] = xt, wt
    out["deconv_y"] = np_ref.conv2d_transpose(xt, wt)
    xs = rng.randn(2, 8, 8, 4)
    ws = rng.randn(3, 3, 4, 5) * 0.2
    out["conv3_x"], out["conv3_w"], out["conv3_y"] = xs, ws, np_ref.conv2d(xs, ws, 2, "SAME")
    xv = rng.randn(3, 4, 4, 4)
    wv = rng.randn(4, 4, 4, 5) * 0.2
    out["convv_x"], out["convv_w"], out["convv_y"] = xv, wv, np_ref.conv2d(xv, wv, 1, "VALID")
    # (iii) batch norm, lrelu
    z = rng.randn(2, 4, 4, 5) * 2 + 1
    beta = rng.randn(5)
    out["bn_z"], out["bn_beta"], out["bn_y"] = z, beta, np_ref.batch_norm(z, beta)
    out["lrelu_y"] = np_ref.lrelu(z)
    # (iv) losses
    g = rng.uniform(-1, 1, (2, 8, 8, 3))
    n = np.clip(g + 0.2 * rng.randn(2, 8, 8, 3), -1, 1)
    logit_r, logit_g = rng.randn(2, 2, 2, 1) * 2, rng.randn(2, 2, 2, 1) * 2
    out["loss_g"], out["loss_n"], out["loss_lr"], out["loss_lg"] = g, n, logit_r, logit_g
    out["gdl"] = np.array(np_ref.gdl(n, g))
    out["psnr"] = np.array(np_ref.psnr(n, g))
    for kind in ("bce", "wass"):
        out["g_adv_" + kind] = np.array(np_ref.g_adv_loss(logit_g, kind))
        out["d_loss_" + kind] = np.array(np_ref.d_loss(logit_r, logit_g, kind))
    # (v) optimizers: TF-style slots, steps t=1,2
    p0, gr = rng.randn(37) * 0.05, rng.randn(37) * 0.1
    out["opt_p0"], out["opt_g"] = p0, gr
    p, m, v = p0.copy(), np.zeros(37), np.zeros(37)
    for t in (1, 2):
        p, m, v = np_ref.adam_step(p, gr, m, v, t)
        out["adam_p%d" % t] = p.copy()
    p, ms = p0.copy(), np.ones(37)
    for t in (1, 2):
        p, ms = np_ref.rmsprop_step(p, gr, ms)
        out["rmsprop_p%d" % t] = p.copy()
    # (vi) full steps at B=2 (pretrain_g, train_d, train_g), every flag combination the CLI offers
    B = 2
    img = rng.uniform(-1, 1, (B, 64, 64, 3)).astype(np.float32)
    nxt = np.clip(img + 0.1 * rng.randn(B, 64, 64, 3), -1, 1).astype(np.float32)
    act = rng.randn(B, 10).astype(np.float32)
    state = rng.randn(B, 5).astype(np.float32)
    out["step_img"], out["step_next"], out["step_act"], out["step_state"] = img, nxt, act, state
    for dna in (True, False):
        for loss, opt in (("bce", "adam"), ("wass", "rmsprop")):
            prng = np.random.RandomState(7)
            params = np_ref.init_params(np_ref.g_dna_spec(6) if dna else np_ref.g_direct_spec(), prng)
            params.update(np_ref.init_params(np_ref.d_spec(), prng))
            tr = torch_ref.Trainer(params, True, loss, opt, dna, ksize=6)
            tag = "step_%s_%s_%s" % ("dna" if dna else "direct", loss, opt)
            out[tag + "_pretrain_g_loss"] = np.array(tr.pretrain_g(img, nxt, act, state))
            s = tr.train_d(img, nxt, act, summarize=True)
            frames = tr.train_g(img, nxt, act, state)
            s2 = tr.summaries()
            out[tag + "_d_loss"] = np.array(s["discriminator_loss"])
            out[tag + "_g_loss"] = np.array(s2["g_loss"])
            out[tag + "_g_l2_loss"] = np.array(s2["g_l2_loss"])
            out[tag + "_g_psnr"] = np.array(s2["g_psnr"])
            out[tag + "_frame_mean"] = np.array(frames.mean())
            out[tag + "_frame_abs_mean"] = np.array(np.abs(frames).mean())
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden_v1.npz")
    np.savez_compressed(path, **{k: np.asarray(v, dtype=np.float64) for k, v in out.items()})
    print("wrote", path, os.path.getsize(path), "bytes,", len(out), "arrays")


if __name__ == "__main__":
    main()
