"""CUDA path vs the committed golden vectors (tests/golden/golden_v1.npz) -- no oracle code involved at test time --
and the reference-named API surface (ops.py / models.py / Trainer) end to end."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
GOLD = np.load(os.path.join(os.path.dirname(__file__), "golden", "golden_v1.npz"))


def _t(a, cuda):
    return torch.from_numpy(np.ascontiguousarray(np.asarray(a, dtype=np.float32))).to(cuda)


@pytest.mark.parametrize("K", [5, 6])
def test_dna_golden(cuda, K):
    from action_conditioned_gans_b200 import kernels as Kn
    lg, img, dy = (_t(GOLD["dna%d_%s" % (K, n)], cuda) for n in ("logits", "img", "dy"))
    out, dl = torch.empty_like(img), torch.empty_like(lg)
    Kn.dna_fwd(lg, img, out, K)
    Kn.dna_bwd(lg, img, dy, dl, K)
    assert np.abs(out.cpu().numpy() - GOLD["dna%d_out" % K]).max() <= 1e-5
    assert np.abs(dl.cpu().numpy() - GOLD["dna%d_dlogits" % K]).max() <= 1e-5 * np.abs(GOLD["dna%d_dlogits" % K]).max() + 1e-6


def test_conv_golden_fp32(cuda):
    from action_conditioned_gans_b200 import kernels as Kn
    for tag, k, s, pad in (("conv", 5, 2, "SAME"), ("conv3", 3, 2, "SAME"), ("convv", 4, 1, "VALID")):
        x, w, y_ref = GOLD[tag + "_x"], GOLD[tag + "_w"], GOLD[tag + "_y"]
        shape = Kn.conv_shape(x.shape[0], x.shape[1], x.shape[2], x.shape[3], w.shape[3], k, s, pad)
        y = torch.empty(y_ref.shape, device=cuda)
        Kn.conv_fprop_f32(shape, _t(x, cuda), _t(w, cuda), y)
        assert np.abs(y.cpu().numpy() - y_ref).max() < 2e-5
    x, w, y_ref = GOLD["deconv_x"], GOLD["deconv_w"], GOLD["deconv_y"]
    shape = Kn.conv_shape(2, 8, 8, 6, 8, 5, 2, "SAME")
    y = torch.empty(y_ref.shape, device=cuda)
    Kn.conv_dgrad_f32(shape, _t(x, cuda), _t(w, cuda), y)
    assert np.abs(y.cpu().numpy() - y_ref).max() < 2e-5


def test_ops_surface_golden(cuda):
    """ops.py names: lrelu, build_psnr, build_gdl, build_g_adv_loss, build_d_loss."""
    from action_conditioned_gans_b200 import ops
    g, n = _t(GOLD["loss_g"], cuda), _t(GOLD["loss_n"], cuda)
    lr, lg = _t(GOLD["loss_lr"], cuda), _t(GOLD["loss_lg"], cuda)
    assert abs(float(ops.build_gdl(n, g)) - float(GOLD["gdl"])) < 1e-4 * float(GOLD["gdl"])
    assert abs(float(ops.build_psnr(n, g)) - float(GOLD["psnr"])) < 1e-4
    for kind in ("bce", "wass"):
        assert abs(float(ops.build_g_adv_loss(lg, kind)) - float(GOLD["g_adv_" + kind])) < 1e-5
        assert abs(float(ops.build_d_loss(lr, lg, kind)) - float(GOLD["d_loss_" + kind][0])) < 1e-5
    with pytest.raises(ValueError, match="unexpected loss argument"):
        ops.build_g_adv_loss(lg, "hinge")
    z = _t(GOLD["bn_z"], cuda)
    assert np.abs(ops.lrelu(z).cpu().numpy() - GOLD["lrelu_y"]).max() < 1e-6


@pytest.mark.parametrize("dna,loss,opt", [(True, "bce", "adam"), (True, "wass", "rmsprop"), (False, "bce", "adam"),
                                          (False, "wass", "rmsprop")])
@pytest.mark.parametrize("precision,tol", [("fp32", 5e-4), ("bf16", 1e-2)])
def test_full_step_golden(cuda, dna, loss, opt, precision, tol):
    """pretrain_g -> train_d -> train_g at B=2 from seed-7 weights: losses against the committed fp64 values."""
    from action_conditioned_gans_b200.trainer import Trainer
    from oracle import np_ref            # only for the seeded initial weights (same draw as make_golden.py)
    prng = np.random.RandomState(7)
    params = np_ref.init_params(np_ref.g_dna_spec(6) if dna else np_ref.g_direct_spec(), prng)
    params.update(np_ref.init_params(np_ref.d_spec(), prng))
    trn = Trainer(None, True, loss, opt, dna, batch_size=2, ksize=6, params=params, precision=precision)
    img, nxt, act, state = (GOLD["step_" + k].astype(np.float32) for k in ("img", "next", "act", "state"))
    tag = "step_%s_%s_%s" % ("dna" if dna else "direct", loss, opt)
    gl = trn.pretrain_g(img, nxt, act, state)
    ref = float(GOLD[tag + "_pretrain_g_loss"])
    assert abs(gl - ref) <= tol * abs(ref)
    s = trn.train_d(img, nxt, act, summarize=True)
    # B=2 means batch-norm over 8 logits: the D loss is the most rounding-sensitive scalar of the step
    assert abs(s["discriminator_loss"] - float(GOLD[tag + "_d_loss"])) <= 5 * tol * max(1.0, abs(float(GOLD[tag + "_d_loss"])))
    frames = trn.train_g(img, nxt, act, state)
    s2 = trn.summaries()
    assert abs(s2["g_l2_loss"] - float(GOLD[tag + "_g_l2_loss"])) <= 5 * tol * abs(float(GOLD[tag + "_g_l2_loss"]))
    assert abs(s2["g_loss"] - float(GOLD[tag + "_g_loss"])) <= 5 * tol * abs(float(GOLD[tag + "_g_loss"]))
    assert frames.shape == (2, 64, 64, 3) and np.isfinite(frames).all()


def test_models_surface(cuda):
    """models.py names with TF-like variable reuse."""
    from action_conditioned_gans_b200 import models
    models.reset_default_graph()
    B = 3
    img = torch.rand(B, 64, 64, 3, device=cuda) * 2 - 1
    act = torch.randn(B, 10, device=cuda)
    tiled = act.view(B, 1, 1, 10).expand(B, 4, 4, 10)
    frame, state = models.build_generator_transform(img, tiled, batch_size=B, ksize=6)
    assert frame.shape == (B, 64, 64, 3) and state.shape == (B, 5)
    assert float(frame.max()) <= 1 + 1e-5 and float(frame.min()) >= -1 - 1e-5      # DNA is a convex combination
    with pytest.raises(ValueError, match="already exists"):
        models.build_generator_transform(img, tiled, batch_size=B, ksize=6)
    frame2, _ = models.build_generator_transform(img, act, batch_size=B, ksize=6, reuse=True)
    # the fused batch-norm moments are added in a fixed order (tests/test_determinism_gpu.py): a rerun is bitwise equal
    assert torch.equal(frame, frame2)
    d1 = models.build_discriminator(torch.cat([img, frame], 3), act)
    d2 = models.build_discriminator(torch.cat([img, img], 3), act, reuse=True)
    assert d1.shape == d2.shape == (B, 2, 2, 1)
    with pytest.raises(ValueError, match="does not exist"):
        models.reset_default_graph()
        models.build_discriminator(torch.cat([img, img], 3), act, reuse=True)
    models.reset_default_graph()
    out = models.build_generator(img, act)
    assert out.shape == (B, 64, 64, 3) and float(out.abs().max()) <= 1.0


def test_checkpoint_roundtrip_and_cli(cuda, tmp_path):
    """Trainer.save / restore keyed by TF variable names, and the train.py / test.py entry points on synthetic data."""
    from action_conditioned_gans_b200 import train as T
    from action_conditioned_gans_b200 import test as TT
    out = tmp_path / "run"
    T.main(["synthetic", str(out), "--dna", "True", "--adv", "True", "--loss", "bce", "--opt", "adam",
            "--iters", "2", "--pretrain_iters", "0", "--batch_size", "4"])
    ck = T.latest_checkpoint(str(out / "models"))
    assert ck is not None and (out / "logs" / "train.jsonl").exists()
    with np.load(ck) as f:
        assert "g/conv1/weights" in f.files and "d/conv6/BatchNorm/beta" in f.files and "d_opt/m" in f.files
        assert f["g/tconv4/weights"].shape == (5, 5, 36, 128)
    with pytest.raises(FileExistsError):
        T.main(["synthetic", str(out)])                         # train.py:324 os.makedirs fails if OUT exists
    seq = np.random.uniform(-1, 1, (100, 19, 64, 64, 3)).astype(np.float32)
    act = np.random.randn(100, 19, 10).astype(np.float32)
    np.save(tmp_path / "f.npy", seq)
    np.save(tmp_path / "a.npy", act)
    TT.main([str(out / "models"), str(tmp_path / "f.npy"), str(tmp_path / "a.npy"), str(tmp_path / "gifs"), "--dna"])
    assert (tmp_path / "gifs" / "sample0" / "vid0" / "generated.gif").exists()
