"""conv / conv_transpose / batch-norm / loss / optimizer kernels vs the CPU oracle (fp64), through the C-ABI."""
import numpy as np
import pytest
import torch

from oracle import np_ref, torch_ref

pytestmark = pytest.mark.gpu

# (B, H, W, Cin, Cout, k, stride, padding): every distinct geometry of models.py, shrunk in B / channels
CONV_CASES = [
    (2, 64, 64, 3, 32, 5, 2, "SAME"),     # g/conv1
    (2, 16, 16, 20, 24, 5, 2, "SAME"),    # mid encoder layer, ragged channel counts
    (3, 8, 8, 138, 70, 5, 2, "SAME"),     # d/conv3 (concat channels), Cout not a tile multiple
    (2, 16, 16, 128, 32, 3, 2, "SAME"),   # g/sconv3 (pad 0/1)
    (4, 4, 4, 16, 5, 4, 1, "VALID"),      # g/sconv5
    (4, 2, 2, 512, 1, 2, 1, "SAME"),      # d/conv6 (pad 0/1, Cout = 1)
    (1, 7, 9, 5, 6, 5, 2, "SAME"),        # odd spatial extents
    (5, 8, 8, 128, 256, 5, 2, "SAME"),    # d/conv4 at B=5: partial second M tile in every dgrad parity class
    (3, 64, 64, 3, 64, 5, 2, "SAME"),     # adjoint of the direct generator's tconv4
    (3, 32, 32, 64, 128, 5, 2, "SAME"),   # adjoint of the direct generator's tconv3
]


def _t(a, cuda):
    return torch.from_numpy(np.ascontiguousarray(a.astype(np.float32))).to(cuda)


@pytest.mark.parametrize("case", CONV_CASES)
def test_conv_fprop_dgrad_wgrad(cuda, case):
    from action_conditioned_gans_b200 import kernels as Kn
    B, H, W, Cin, Cout, k, s, padding = case
    rng = np.random.RandomState(sum(case[:7]))
    x = rng.randn(B, H, W, Cin)
    w = rng.randn(k, k, Cin, Cout) / np.sqrt(k * k * Cin)
    shape = Kn.conv_shape(B, H, W, Cin, Cout, k, s, padding)
    y_ref = np_ref.conv2d(x, w, s, padding)
    assert y_ref.shape == (B, shape.OH, shape.OW, Cout)
    dy = rng.randn(*y_ref.shape)
    # oracle gradients by autograd of the independent torch restatement
    xt = torch.tensor(x, requires_grad=True)
    wt = torch.tensor(w, requires_grad=True)
    yt = torch_ref.conv2d(xt, wt, s, padding)
    assert np.abs(yt.detach().numpy() - y_ref).max() < 1e-10
    gx, gw = torch.autograd.grad(yt, [xt, wt], torch.tensor(dy))
    y = torch.full(y_ref.shape, float("nan"), device=cuda)
    dx = torch.full(x.shape, float("nan"), device=cuda)
    dw = torch.zeros(w.shape, device=cuda)
    Kn.conv_fprop_f32(shape, _t(x, cuda), _t(w, cuda), y)
    Kn.conv_dgrad_f32(shape, _t(dy, cuda), _t(w, cuda), dx)
    Kn.conv_wgrad_f32(shape, _t(x, cuda), _t(dy, cuda), dw)
    torch.cuda.synchronize()
    for got, ref in ((y, y_ref), (dx, gx.numpy()), (dw, gw.numpy())):
        assert np.abs(got.cpu().numpy() - ref).max() <= 2e-5 * max(1.0, np.abs(ref).max())


@pytest.mark.parametrize("h,cin,cout", [(4, 266, 128), (8, 16, 24), (32, 128, 25), (32, 64, 3)])
def test_conv2d_transpose_is_dgrad(cuda, h, cin, cout):
    """slim.conv2d_transpose (SAME, stride 2, weights [k,k,Cout,Cin]) == acg_conv_dgrad of the adjoint conv."""
    from action_conditioned_gans_b200 import kernels as Kn
    B, k = 2, 5
    rng = np.random.RandomState(h + cin)
    x = rng.randn(B, h, h, cin)
    w = rng.randn(k, k, cout, cin) / np.sqrt(k * k * cin)
    y_ref = np_ref.conv2d_transpose(x, w)
    y_ref2 = torch_ref.conv2d_transpose(torch.tensor(x), torch.tensor(w)).numpy()
    assert np.abs(y_ref - y_ref2).max() < 1e-10
    shape = Kn.conv_shape(B, 2 * h, 2 * h, cout, cin, k, 2, "SAME")
    y = torch.full(y_ref.shape, float("nan"), device=cuda)
    Kn.conv_dgrad_f32(shape, _t(x, cuda), _t(w, cuda), y)
    assert np.abs(y.cpu().numpy() - y_ref).max() <= 2e-5 * max(1.0, np.abs(y_ref).max())
    # adjoint identity <conv(u), g> == <u, deconv(g)>
    u = rng.randn(B, 2 * h, 2 * h, cout)
    cu = torch.empty(B, h, h, cin, device=cuda)
    Kn.conv_fprop_f32(shape, _t(u, cuda), _t(w, cuda), cu)
    lhs = float((cu.double().cpu().numpy() * x).sum())
    rhs = float((u * y.double().cpu().numpy()).sum())
    assert abs(lhs - rhs) <= 1e-4 * max(1.0, abs(lhs))


@pytest.mark.parametrize("act", ["relu", "lrelu", "none", "tanh"])
@pytest.mark.parametrize("has_bn", [True, False])
@pytest.mark.parametrize("rows,Cc", [(256, 37), (3072, 64), (320, 128)])
def test_bn_act_forward_backward(cuda, act, has_bn, rows, Cc):
    from action_conditioned_gans_b200 import kernels as Kn
    rng = np.random.RandomState(5)
    z = rng.randn(rows, Cc) * 2 + 0.5
    beta = rng.randn(Cc) * 0.3
    dA = rng.randn(rows, Cc)
    dA2 = rng.randn(rows, Cc)
    zt = torch.tensor(z, requires_grad=True)
    bt = torch.tensor(beta, requires_grad=True)
    fn = {"relu": torch.relu, "lrelu": torch_ref.lrelu, "none": (lambda v: v), "tanh": torch.tanh}[act]
    if has_bn:
        a_ref = fn(torch_ref.batch_norm(zt.view(1, 1, rows, Cc), bt).view(rows, Cc))
    else:
        a_ref = fn(zt + bt)
    gz, gb = torch.autograd.grad(a_ref, [zt, bt], torch.tensor(dA + dA2))

    zc = _t(z, cuda)
    f64 = torch.zeros(4 * Cc, dtype=torch.float64, device=cuda)
    stats, red = f64[:2 * Cc], f64[2 * Cc:]
    mean, rstd, scale, shift = (torch.empty(Cc, device=cuda) for _ in range(4))
    ld_out = Cc + 3
    out = torch.zeros(rows, ld_out, device=cuda)
    bc = _t(beta, cuda)
    if has_bn:
        Kn.bn_stats(zc, rows, Cc, Cc, 1, stats)
        Kn.bn_finalize(stats, bc, rows, Cc, 1, mean, rstd, scale, shift)
        Kn.bn_act_fwd(zc, rows, Cc, Cc, 1, scale, shift, act, out, ld_out)
        m, r, sh = mean, rstd, shift
    else:
        Kn.bn_act_fwd(zc, rows, Cc, Cc, 1, None, bc, act, out, ld_out)
        m, r, sh = None, None, bc
    assert np.abs(out[:, :Cc].cpu().numpy() - a_ref.detach().numpy()).max() < 2e-5
    assert float(out[:, Cc:].abs().max()) == 0.0          # the wider concat buffer is left alone
    dz = torch.empty(rows, Cc, device=cuda)
    dbeta = torch.zeros(Cc, device=cuda)
    Kn.bn_act_bwd_reduce(_t(dA, cuda), _t(dA2, cuda), Cc, zc, Cc, rows, Cc, 1, m, r, sh, act, red)
    Kn.bn_act_bwd_apply(_t(dA, cuda), _t(dA2, cuda), Cc, zc, Cc, rows, Cc, 1, m, r, sh, act, has_bn, red, dz, dbeta)
    assert np.abs(dz.cpu().numpy() - gz.numpy()).max() < 5e-5
    assert np.abs(dbeta.cpu().numpy() - gb.numpy()).max() < 5e-4


def test_bn_groups_and_bf16(cuda):
    """Two row groups with independent statistics (the two discriminator applications) and bf16 storage."""
    from action_conditioned_gans_b200 import kernels as Kn
    rows, Cc = 512, 64
    rng = np.random.RandomState(9)
    z = rng.randn(rows, Cc)
    z[rows // 2:] = z[rows // 2:] * 3 + 1
    zc = _t(z, cuda)
    stats = torch.zeros(2 * 2 * Cc, dtype=torch.float64, device=cuda)
    mean, rstd, scale, shift = (torch.empty(2 * Cc, device=cuda) for _ in range(4))
    Kn.bn_stats(zc, rows, Cc, Cc, 2, stats)
    Kn.bn_finalize(stats, None, rows // 2, Cc, 2, mean, rstd, scale, shift)
    out = torch.empty(rows, Cc, device=cuda, dtype=torch.bfloat16)
    Kn.bn_act_fwd(zc, rows, Cc, Cc, 2, scale, shift, "none", out, Cc)
    ref = np.concatenate([np_ref.batch_norm(z[:rows // 2].reshape(1, 1, -1, Cc), 0),
                          np_ref.batch_norm(z[rows // 2:].reshape(1, 1, -1, Cc), 0)], axis=2).reshape(rows, Cc)
    assert np.abs(out.float().cpu().numpy() - ref).max() < 2e-2


def test_concat_helpers(cuda):
    from action_conditioned_gans_b200 import kernels as Kn
    B, hw, Cc = 3, 16, 12
    a = torch.randn(B * hw, Cc, device=cuda)
    acts = torch.randn(B, 10, device=cuda)
    cat = torch.zeros(B * hw, Cc + 10, device=cuda)
    Kn.copy_channels(a, Cc, 0, cat, Cc + 10, 0, B * hw, Cc)
    Kn.tile_actions(acts, B, hw, cat, Cc + 10, Cc)
    ref = np.concatenate([a.cpu().numpy().reshape(B, 4, 4, Cc), np_ref.tile_actions(acts.cpu().numpy(), 4)], 3)
    assert np.array_equal(cat.cpu().numpy().reshape(B, 4, 4, Cc + 10), ref)


@pytest.mark.parametrize("B", [1, 3])
def test_frame_losses(cuda, B):
    from action_conditioned_gans_b200 import kernels as Kn
    rng = np.random.RandomState(B)
    g = rng.uniform(-1, 1, (B, 64, 64, 3))
    n = np.clip(g + 0.3 * rng.randn(B, 64, 64, 3), -1, 1)
    dadv = rng.randn(B, 64, 64, 6)
    gt = torch.tensor(g, requires_grad=True)
    nt = torch.tensor(n)
    w_l1, w_gdl = 0.05 / B, 1.0
    l1 = (gt - nt).abs().sum()
    gdl = torch_ref.build_gdl(nt, gt)
    assert abs(float(gdl) - np_ref.gdl(n, g)) < 1e-8 * float(gdl)
    (grad,) = torch.autograd.grad(w_l1 * l1 + w_gdl * gdl, [gt])
    grad = grad.numpy() + dadv[..., 3:6]
    sums = torch.zeros(3, dtype=torch.float64, device=cuda)
    dg = torch.empty(B, 64, 64, 3, device=cuda)
    Kn.frame_losses(_t(g, cuda), _t(n, cuda), sums, dg, w_l1, w_gdl, _t(dadv, cuda), 6, 3)
    s = sums.cpu().numpy()
    g32, n32 = g.astype(np.float32).astype(np.float64), n.astype(np.float32).astype(np.float64)
    assert abs(s[0] - np.abs(g32 - n32).sum()) <= 1e-6 * s[0]
    assert abs(s[1] - ((g32 - n32) ** 2).sum()) <= 1e-6 * s[1]
    assert abs(s[2] - np_ref.gdl(n32, g32)) <= 1e-5 * s[2]
    # the GDL gradient is a sum of signs: compare away from the kinks of |.|, where fp32 rounding may flip one
    diff = np.abs(dg.cpu().numpy() - grad)
    assert (diff > 1e-5).mean() < 1e-3


@pytest.mark.parametrize("kind,label", [("bce", 1.0), ("bce", 0.9), ("bce", 0.0), ("wass", 1.0), ("wass", -1.0)])
def test_dlogit_loss(cuda, kind, label):
    from action_conditioned_gans_b200 import kernels as Kn
    rng = np.random.RandomState(4)
    x = rng.randn(64 * 4) * 3
    xt = torch.tensor(x, requires_grad=True)
    loss = torch_ref.sigmoid_ce(label, xt) if kind == "bce" else label * xt.mean()
    (gx,) = torch.autograd.grad(loss, [xt])
    lo = torch.zeros(1, device=cuda)
    dl = torch.empty(x.size, device=cuda)
    Kn.dlogit_loss(_t(x, cuda), x.size, kind, label, 1.0, lo, dl)
    assert abs(float(lo) - float(loss)) < 1e-6 * max(1.0, abs(float(loss)))
    assert np.abs(dl.cpu().numpy() - gx.numpy()).max() < 1e-8
    if kind == "bce" and label == 1.0:          # BCE(0, target 1) = ln 2
        Kn.dlogit_loss(torch.zeros(8, device=cuda), 8, "bce", 1.0, 1.0, lo, None)
        assert abs(float(lo) - np.log(2)) < 1e-6
    with pytest.raises(ValueError):
        Kn.dlogit_loss(_t(x, cuda), x.size, "hinge", 1.0, 1.0, lo, dl)


def test_state_loss(cuda):
    from action_conditioned_gans_b200 import kernels as Kn
    rng = np.random.RandomState(2)
    s, t = rng.randn(16, 5), rng.randn(16, 5)
    st = torch.tensor(s, requires_grad=True)
    loss = torch.sqrt(((st - torch.tensor(t)) ** 2).sum()) / 16
    (gs,) = torch.autograd.grad(loss, [st])
    lo = torch.zeros(1, device=cuda)
    ds = torch.empty(16, 5, device=cuda)
    Kn.state_loss(_t(s, cuda), _t(t, cuda), 80, 1.0 / 16, 1.0, lo, ds)
    assert abs(float(lo) - float(loss)) < 1e-6
    assert np.abs(ds.cpu().numpy() - gs.numpy()).max() < 1e-7


@pytest.mark.parametrize("n", [1, 7, 4096, 100003])
def test_optimizers(cuda, n):
    from action_conditioned_gans_b200 import kernels as Kn
    rng = np.random.RandomState(n)
    n_pad = (n + 3) // 4 * 4
    p0, g = rng.randn(n_pad) * 0.05, rng.randn(n_pad) * 0.1
    # Adam, steps t = 1 and 2 (bias correction), TF-1.0 formula (eps outside the corrected sqrt)
    p, m, v = p0.copy(), np.zeros(n_pad), np.zeros(n_pad)
    pc, mc, vc, gc = _t(p0, cuda), torch.zeros(n_pad, device=cuda), torch.zeros(n_pad, device=cuda), _t(g, cuda)
    for t in (1, 2):
        p, m, v = np_ref.adam_step(p, g.astype(np.float32).astype(np.float64), m, v, t)
        lr_t = 1e-3 * np.sqrt(1 - 0.999 ** t) / (1 - 0.9 ** t)
        Kn.adam_step(pc, gc, mc, vc, float(lr_t))
        assert np.abs(pc.cpu().numpy() - p).max() < 2e-6
    # RMSProp from ms = ones, then clipped like d_vars (train.py:89)
    p, ms = p0.copy(), np.ones(n_pad)
    pc, msc = _t(p0, cuda), torch.ones(n_pad, device=cuda)
    for _ in range(2):
        p, ms = np_ref.rmsprop_step(p, g.astype(np.float32).astype(np.float64), ms)
        p = np_ref.clip(p)
        Kn.rmsprop_step(pc, gc, msc, 5e-5, clip=(-0.01, 0.01))
        assert np.abs(pc.cpu().numpy() - p).max() < 1e-6
        assert float(pc.abs().max()) <= 0.01


def test_pack_frames(cuda):
    """acg_pack_frames == concat([img, frame], 3) rounded to bf16, zero pad channels (train.py:64,68 / g/conv1 input)."""
    from action_conditioned_gans_b200 import kernels as Kn
    rng = np.random.RandomState(4)
    a = torch.from_numpy(rng.uniform(-1, 1, (3, 64, 64, 3)).astype(np.float32)).to(cuda)
    b = torch.from_numpy(rng.uniform(-1, 1, (3, 64, 64, 3)).astype(np.float32)).to(cuda)
    rows = 3 * 64 * 64
    for second in (b, None):
        for ld in (8, 16):
            out = torch.full((3, 64, 64, ld), 5.0, dtype=torch.bfloat16, device=cuda)
            Kn.pack_frames(a, second, out, rows)
            ref = torch.zeros(3, 64, 64, ld, device=cuda)
            ref[..., 0:3] = a
            if second is not None:
                ref[..., 3:6] = second
            assert torch.equal(out, ref.to(torch.bfloat16))


@pytest.mark.parametrize("case", [(256, 4, 4, 256, 272, "relu"), (6, 16, 16, 128, 144, "lrelu"), (3, 5, 7, 24, 40, "none")])
def test_activation_pass_writes_the_action_concat(cuda, case):
    """acg_bn_act_fwd_cat == acg_bn_act_fwd + acg_tile_actions (tf.concat([features, tf.tile(action)], 3) of
    models.py:16,38,84) on the bf16 fast path and, for a shape outside it (C % 8 != 0 is impossible with padded buffers,
    so: fp32 output), through the two-launch form."""
    from action_conditioned_gans_b200 import kernels as Kn
    B, H, W, Cc, ld, act = case
    g = torch.Generator(device=cuda).manual_seed(B + Cc)
    rows = B * H * W
    z = torch.randn(rows, Cc, device=cuda, generator=g).to(torch.bfloat16)
    scale = torch.rand(Cc, device=cuda, generator=g) + 0.5
    shift = torch.randn(Cc, device=cuda, generator=g)
    actions = torch.randn(B, 10, device=cuda, generator=g)
    for dt in (torch.bfloat16, torch.float32):
        zz = z if dt == torch.bfloat16 else z.float()
        want = torch.full((rows, ld), 7.0, dtype=dt, device=cuda)
        Kn.bn_act_fwd(zz, rows, Cc, Cc, 1, scale, shift, act, want, ld)
        Kn.tile_actions(actions, B, H * W, want, ld, Cc)
        got = torch.full((rows, ld), 7.0, dtype=dt, device=cuda)
        Kn.bn_act_fwd_cat(zz, rows, Cc, Cc, scale, shift, act, got, ld, actions, H * W, Cc)
        torch.cuda.synchronize()
        assert torch.equal(got, want)
        assert bool((got[:, Cc + 10:] == 7.0).all())                  # beyond the concat: untouched
        ref = actions.to(dt).repeat_interleave(H * W, dim=0)
        assert torch.equal(got[:, Cc:Cc + 10], ref)
