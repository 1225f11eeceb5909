"""CPU oracle: the NumPy restatement vs the independently written torch restatement, analytic identities, and the
committed golden vectors (tests/golden/golden_v1.npz, made by tests/golden/make_golden.py).  No GPU needed."""
import os

import numpy as np
import pytest
import torch

from oracle import np_ref, torch_ref

GOLD = np.load(os.path.join(os.path.dirname(__file__), "golden", "golden_v1.npz"))


def t64(a):
    return torch.tensor(np.asarray(a), dtype=torch.float64)


@pytest.mark.parametrize("k,s,pad,shape", [(5, 2, "SAME", (2, 8, 8, 3)), (3, 2, "SAME", (2, 16, 16, 4)),
                                           (2, 1, "SAME", (3, 2, 2, 6)), (4, 1, "VALID", (2, 4, 4, 5)),
                                           (5, 2, "SAME", (1, 7, 9, 2))])
def test_conv_np_vs_torch(k, s, pad, shape):
    rng = np.random.RandomState(k + s)
    x, w = rng.randn(*shape), rng.randn(k, k, shape[3], 7)
    assert np.abs(np_ref.conv2d(x, w, s, pad) - torch_ref.conv2d(t64(x), t64(w), s, pad).numpy()).max() < 1e-12


def test_same_padding_table():
    # k5 s2 even-in (1,2); k3 s2 (0,1); k2 s1 (0,1); k6 s1 (2,3); k5 s1 (2,2)
    assert np_ref.same_pad(64, 5, 2) == (32, 1, 2)
    assert np_ref.same_pad(16, 3, 2) == (8, 0, 1)
    assert np_ref.same_pad(2, 2, 1) == (2, 0, 1)
    assert np_ref.same_pad(64, 6, 1) == (64, 2, 3)
    assert np_ref.same_pad(64, 5, 1) == (64, 2, 2)


@pytest.mark.parametrize("h", [2, 4, 5])
def test_deconv_np_vs_torch_and_adjoint(h):
    rng = np.random.RandomState(h)
    x, w = rng.randn(2, h, h, 6), rng.randn(5, 5, 4, 6)
    y = np_ref.conv2d_transpose(x, w)
    assert y.shape == (2, 2 * h, 2 * h, 4)
    assert np.abs(y - torch_ref.conv2d_transpose(t64(x), t64(w)).numpy()).max() < 1e-12
    u = rng.randn(2, 2 * h, 2 * h, 4)
    # exact adjoint of the SAME conv with weights read as HWIO [k,k,I=4,O=6]
    assert abs((np_ref.conv2d(u, w, 2) * x).sum() - (u * y).sum()) < 1e-9


@pytest.mark.parametrize("K", [5, 6])
def test_dna_np_vs_torch_vs_loops(K):
    rng = np.random.RandomState(K)
    lg, img = rng.randn(2, 6, 7, K * K), rng.uniform(-1, 1, (2, 6, 7, 3))
    y = np_ref.dna_forward(lg, img, K)
    assert np.abs(y - np_ref.dna_forward_loops(lg, img, K)).max() < 1e-12
    lt = t64(lg).requires_grad_(True)
    yt = torch_ref.dna_transform(lt, t64(img), K)
    assert np.abs(y - yt.detach().numpy()).max() < 1e-12
    dy = rng.randn(*y.shape)
    (gl,) = torch.autograd.grad(yt, [lt], t64(dy))
    bw = np_ref.dna_backward(lg, img, dy, K)
    assert np.abs(bw - gl.numpy()).max() < 1e-12
    assert np.abs(bw.sum(-1)).max() < 1e-12            # rows of the softmax Jacobian sum to zero


@pytest.mark.parametrize("K", [5, 6])
def test_dna_analytic(K):
    rng = np.random.RandomState(1)
    img = rng.uniform(-1, 1, (1, 8, 8, 3))
    pb = (K - 1) // 2
    pad = np.zeros((1, 8 + K - 1, 8 + K - 1, 3))
    pad[:, pb:pb + 8, pb:pb + 8] = img
    box = sum(pad[:, a:a + 8, b:b + 8] for a in range(K) for b in range(K)) / (K * K)
    assert np.abs(np_ref.dna_forward(np.zeros((1, 8, 8, K * K)), img, K) - box).max() < 1e-12
    lg = np.full((1, 8, 8, K * K), -1e3)
    lg[..., 1 * K + 3] = 0
    assert np.abs(np_ref.dna_forward(lg, img, K) - pad[:, 1:9, 3:11]).max() < 1e-12
    out = np_ref.dna_forward(rng.randn(1, 8, 8, K * K) * 4, img, K)
    assert out.max() <= max(img.max(), 0) and out.min() >= min(img.min(), 0)


def test_networks_np_vs_torch():
    rng = np.random.RandomState(0)
    img, act = rng.uniform(-1, 1, (2, 64, 64, 3)), rng.randn(2, 10)
    for ks in (5, 6):
        p = np_ref.init_params(np_ref.g_dna_spec(ks), rng, np.float64)
        f, s, lg = np_ref.generator_transform(p, img, act, ks)
        pt = {k: t64(v) for k, v in p.items()}
        ft, st_, lgt = torch_ref.generator_transform(pt, t64(img), t64(act), ks)
        assert f.shape == (2, 64, 64, 3) and s.shape == (2, 5) and lg.shape == (2, 64, 64, ks * ks)
        assert np.abs(f - ft.numpy()).max() < 1e-10 and np.abs(s - st_.numpy()).max() < 1e-10
    p = np_ref.init_params(np_ref.g_direct_spec(), rng, np.float64)
    f = np_ref.generator_direct(p, img, act)
    assert np.abs(f - torch_ref.generator_direct({k: t64(v) for k, v in p.items()}, t64(img), t64(act)).numpy()).max() < 1e-10
    p = np_ref.init_params(np_ref.d_spec(), rng, np.float64)
    d = np_ref.discriminator(p, np.concatenate([img, f], 3), act)
    dt = torch_ref.discriminator({k: t64(v) for k, v in p.items()}, t64(np.concatenate([img, f], 3)), t64(act))
    assert d.shape == (2, 2, 2, 1) and np.abs(d - dt.numpy()).max() < 1e-9


def test_parameter_counts_and_names():
    # SURVEY.md 8(a) a15
    def count(spec):
        return sum(v.size for v in np_ref.init_params(spec, np.random.RandomState(0)).values())
    assert count(np_ref.g_dna_spec(5)) == 2871694
    assert count(np_ref.g_dna_spec(6)) == 2906905
    assert count(np_ref.g_direct_spec()) == 8676611
    assert count(np_ref.d_spec()) == 4755137
    names = set(np_ref.init_params(np_ref.g_dna_spec(6), np.random.RandomState(0)))
    assert {"g/conv1/weights", "g/conv1/BatchNorm/beta", "g/tconv4/biases", "g/sconv5/biases"} <= names
    assert "g/tconv4/BatchNorm/beta" not in names
    dn = set(np_ref.init_params(np_ref.d_spec(), np.random.RandomState(0)))
    assert "d/conv6/BatchNorm/beta" in dn and "d/conv6/biases" not in dn


def test_losses_and_identities():
    rng = np.random.RandomState(3)
    x = rng.randn(100)
    assert np.allclose(np_ref.lrelu(x), np.maximum(x, 0.2 * x))
    assert abs(np_ref.sigmoid_cross_entropy(np.ones(4), np.zeros(4)) - np.log(2)) < 1e-12
    a, b = rng.uniform(-1, 1, (2, 8, 8, 3)), rng.uniform(-1, 1, (2, 8, 8, 3))
    assert abs(np_ref.gdl(a, b) - float(torch_ref.build_gdl(t64(a), t64(b)))) < 1e-9
    assert abs(np_ref.psnr(a, b) - float(torch_ref.build_psnr(t64(a), t64(b)))) < 1e-9
    lr, lg = rng.randn(2, 2, 2, 1), rng.randn(2, 2, 2, 1)
    for kind in ("bce", "wass"):
        assert abs(np_ref.d_loss(lr, lg, kind)[0] - float(torch_ref.build_d_loss(t64(lr), t64(lg), kind)[0])) < 1e-12
        assert abs(np_ref.g_adv_loss(lg, kind) - float(torch_ref.build_g_adv_loss(t64(lg), kind))) < 1e-12
    with pytest.raises(ValueError, match="unexpected loss argument"):
        np_ref.g_adv_loss(lg, "hinge")
    with pytest.raises(ValueError, match="unexpected loss argument"):
        torch_ref.build_d_loss(t64(lr), t64(lg), "hinge")


def test_optimizers_tf_semantics():
    p, g = np.array([1.0, -2.0]), np.array([0.5, -0.25])
    # Adam t=1: m=0.1g, v=0.001g^2, lr_t = lr*sqrt(0.001)/0.1 -> p - lr*g/(|g| + eps*sqrt(1000)...) ~ p - lr*sign(g)
    p1, m, v = np_ref.adam_step(p, g, np.zeros(2), np.zeros(2), 1)
    assert np.allclose(p1, p - 1e-3 * np.sign(g), atol=1e-8)
    # RMSProp: ms starts at ONE, eps inside the sqrt
    p1, ms = np_ref.rmsprop_step(p, g, np.ones(2))
    assert np.allclose(ms, 0.9 + 0.1 * g * g) and np.allclose(p1, p - 5e-5 * g / np.sqrt(ms + 1e-10))
    tp = {"w": t64(p)}
    opt = torch_ref.TFAdam(tp)
    opt.step(tp, {"w": t64(g)})
    assert np.allclose(tp["w"].numpy(), np_ref.adam_step(p, g, np.zeros(2), np.zeros(2), 1)[0])


def test_golden_vectors():
    G = GOLD
    for K in (5, 6):
        lg, img, dy = G["dna%d_logits" % K], G["dna%d_img" % K], G["dna%d_dy" % K]
        assert np.abs(np_ref.dna_forward(lg, img, K) - G["dna%d_out" % K]).max() < 1e-13
        assert np.abs(np_ref.dna_backward(lg, img, dy, K) - G["dna%d_dlogits" % K]).max() < 1e-13
        assert np.abs(torch_ref.dna_transform(t64(lg), t64(img), K).numpy() - G["dna%d_out" % K]).max() < 1e-12
    assert np.abs(torch_ref.conv2d(t64(G["conv_x"]), t64(G["conv_w"]), 2).numpy() - G["conv_y"]).max() < 1e-12
    assert np.abs(torch_ref.conv2d_transpose(t64(G["deconv_x"]), t64(G["deconv_w"])).numpy() - G["deconv_y"]).max() < 1e-12
    assert np.abs(np_ref.conv2d(G["conv3_x"], G["conv3_w"], 2) - G["conv3_y"]).max() < 1e-13
    assert np.abs(np_ref.conv2d(G["convv_x"], G["convv_w"], 1, "VALID") - G["convv_y"]).max() < 1e-13
    assert np.abs(torch_ref.batch_norm(t64(G["bn_z"]), t64(G["bn_beta"])).numpy() - G["bn_y"]).max() < 1e-12
    assert abs(float(torch_ref.build_gdl(t64(G["loss_n"]), t64(G["loss_g"]))) - float(G["gdl"])) < 1e-9
    p, m, v = G["opt_p0"].copy(), np.zeros(37), np.zeros(37)
    for t in (1, 2):
        p, m, v = np_ref.adam_step(p, G["opt_g"], m, v, t)
        assert np.abs(p - G["adam_p%d" % t]).max() < 1e-15


def test_golden_full_step_dna_bce_adam():
    """One pretrain_g + train_d + train_g at B=2 reproduces the committed losses (oracle regression pin)."""
    G = GOLD
    prng = np.random.RandomState(7)
    params = np_ref.init_params(np_ref.g_dna_spec(6), prng)
    params.update(np_ref.init_params(np_ref.d_spec(), prng))
    tr = torch_ref.Trainer(params, True, "bce", "adam", True, ksize=6)
    img, nxt = G["step_img"].astype(np.float32), G["step_next"].astype(np.float32)
    act, state = G["step_act"].astype(np.float32), G["step_state"].astype(np.float32)
    tag = "step_dna_bce_adam"
    assert abs(tr.pretrain_g(img, nxt, act, state) - float(G[tag + "_pretrain_g_loss"])) < 1e-6 * abs(float(G[tag + "_pretrain_g_loss"]))
    s = tr.train_d(img, nxt, act, summarize=True)
    assert abs(s["discriminator_loss"] - float(G[tag + "_d_loss"])) < 1e-8
    tr.train_g(img, nxt, act, state)
    assert abs(tr.summaries()["g_loss"] - float(G[tag + "_g_loss"])) < 1e-6 * abs(float(G[tag + "_g_loss"]))
