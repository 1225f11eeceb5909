"""The `acg::` torch.library custom ops (torch_ops.py): forward AND autograd of every op against the CPU oracle's
autograd, and the differentiable models.py / ops.py surface built on them (models.py:8,24,76; ops.py:19-50,100)."""
import numpy as np
import pytest
import torch

from oracle import np_ref, torch_ref
from tests._gates import device_gates

pytestmark = pytest.mark.gpu


@pytest.fixture(autouse=True)
def _clear_gates():
    yield
    torch_ref.GATES = None


def _t64(a, grad=False):
    return torch.tensor(np.asarray(a), dtype=torch.float64, requires_grad=grad)


@pytest.mark.parametrize("K", [5, 6])
def test_dna_op_forward_and_backward(cuda, K):
    from action_conditioned_gans_b200 import torch_ops as T
    rng = np.random.RandomState(K)
    lg, img, r = rng.randn(3, 64, 64, K * K), rng.uniform(-1, 1, (3, 64, 64, 3)), rng.randn(3, 64, 64, 3)
    x = torch.tensor(lg, dtype=torch.float32, device=cuda, requires_grad=True)
    y = T.dna(x, torch.tensor(img, dtype=torch.float32, device=cuda), K)
    (y * torch.tensor(r, dtype=torch.float32, device=cuda)).sum().backward()
    xr = _t64(lg, True)
    yr = torch_ref.dna_transform(xr, _t64(img), K)
    (yr * _t64(r)).sum().backward()
    assert np.abs(y.detach().cpu().numpy() - yr.detach().numpy()).max() <= 1e-5
    assert np.abs(x.grad.cpu().numpy() - xr.grad.numpy()).max() <= 1e-5 * max(1.0, float(xr.grad.abs().max()))


@pytest.mark.parametrize("bf16,tol", [(False, 2e-5), (True, 1e-2)])
def test_conv_ops_forward_and_backward(cuda, bf16, tol):
    from action_conditioned_gans_b200 import torch_ops as T
    rng = np.random.RandomState(0)
    x, w = rng.randn(2, 16, 16, 24), rng.randn(5, 5, 24, 40) / 25
    wt = rng.randn(5, 5, 12, 24) / 25            # conv2d_transpose: [k,k,Cout,Cin]
    r, rt = rng.randn(2, 8, 8, 40), rng.randn(2, 32, 32, 12)
    dev = lambda a, g=False: torch.tensor(a, dtype=torch.float32, device=cuda, requires_grad=g)
    xd, wd, wtd = dev(x, True), dev(w, True), dev(wt, True)
    y = T.conv2d(xd, wd, 2, True, bf16)
    yt = T.conv2d_transpose(xd, wtd, 2, bf16)
    ((y * dev(r)).sum() + (yt * dev(rt)).sum()).backward()
    xr, wr, wtr = _t64(x, True), _t64(w, True), _t64(wt, True)
    yr = torch_ref.conv2d(xr, wr, 2)
    ytr = torch_ref.conv2d_transpose(xr, wtr, 2)
    ((yr * _t64(r)).sum() + (ytr * _t64(rt)).sum()).backward()
    for got, ref in ((y, yr), (yt, ytr), (xd.grad, xr.grad), (wd.grad, wr.grad), (wtd.grad, wtr.grad)):
        g, rf = got.detach().cpu().numpy(), ref.detach().numpy()
        assert np.abs(g - rf).max() <= tol * max(1.0, np.abs(rf).max())


@pytest.mark.parametrize("act", ["relu", "lrelu", "none"])
def test_bn_act_op(cuda, act):
    from action_conditioned_gans_b200 import torch_ops as T
    rng = np.random.RandomState(1)
    z, beta, r = rng.randn(3, 8, 8, 20) * 2 + 0.5, rng.randn(20) * 0.3, rng.randn(3, 8, 8, 20)
    zd = torch.tensor(z, dtype=torch.float32, device=cuda, requires_grad=True)
    bd = torch.tensor(beta, dtype=torch.float32, device=cuda, requires_grad=True)
    a = T.bn_act(zd, bd, act)
    (a * torch.tensor(r, dtype=torch.float32, device=cuda)).sum().backward()
    zr, br = _t64(z, True), _t64(beta, True)
    u = torch_ref.batch_norm(zr, br)
    ar = {"relu": torch.relu, "lrelu": torch_ref.lrelu, "none": lambda t: t}[act](u)
    (ar * _t64(r)).sum().backward()
    assert np.abs(a.detach().cpu().numpy() - ar.detach().numpy()).max() <= 2e-5
    assert np.abs(zd.grad.cpu().numpy() - zr.grad.numpy()).max() <= 2e-4 * max(1.0, float(zr.grad.abs().max()))
    assert np.abs(bd.grad.cpu().numpy() - br.grad.numpy()).max() <= 2e-4 * max(1.0, float(br.grad.abs().max()))


def test_ops_surface_is_differentiable(cuda):
    """ops.py:19-50,100-120: values and gradients of build_gdl / build_psnr / build_g_adv_loss / build_d_loss / lrelu."""
    from action_conditioned_gans_b200 import ops
    rng = np.random.RandomState(2)
    g, n = rng.uniform(-1, 1, (2, 64, 64, 3)), rng.uniform(-1, 1, (2, 64, 64, 3))
    xg, xr_ = rng.randn(2, 2, 2, 1), rng.randn(2, 2, 2, 1)
    dev = lambda a, gr=False: torch.tensor(a, dtype=torch.float32, device=cuda, requires_grad=gr)
    gd, xgd, xrd = dev(g, True), dev(xg, True), dev(xr_, True)
    loss = ops.build_gdl(gd, dev(n)) + 0.05 * ops.build_psnr(dev(n), gd) + ops.build_g_adv_loss(xgd, "bce") \
        + ops.build_d_loss(xrd, xgd, "bce") + ops.build_d_loss(xrd, xgd, "wass") + ops.lrelu(xgd).sum()
    loss.backward()
    g6, xg6, xr6 = _t64(g, True), _t64(xg, True), _t64(xr_, True)
    ref = torch_ref.build_gdl(_t64(n), g6) + 0.05 * torch_ref.build_psnr(_t64(n), g6) \
        + torch_ref.build_g_adv_loss(xg6, "bce") + torch_ref.build_d_loss(xr6, xg6, "bce")[0] \
        + torch_ref.build_d_loss(xr6, xg6, "wass")[0] + torch_ref.lrelu(xg6).sum()
    ref.backward()
    assert abs(float(loss) - float(ref)) <= 1e-5 * abs(float(ref))
    for got, rf in ((gd.grad, g6.grad), (xgd.grad, xg6.grad), (xrd.grad, xr6.grad)):
        assert np.abs(got.cpu().numpy() - rf.numpy()).max() <= 1e-4 * max(1.0, float(rf.abs().max()))
    with pytest.raises(ValueError, match="unexpected loss argument"):
        ops.build_g_adv_loss(xgd, "hinge")


def test_models_surface_backward_matches_oracle_autograd(cuda):
    """build_generator_transform(...) -> loss -> .backward(): the gradient lands in models.trainable('g').grad and agrees
    with torch_ref's autograd (bf16 engine: relative Frobenius norm per variable, the oracle evaluated on the linear
    pieces the device took -- this is a GRADIENT check, see torch_ref.GATES); same for the discriminator including the
    gradient w.r.t. its input frames."""
    from action_conditioned_gans_b200 import models
    from action_conditioned_gans_b200 import torch_ops as T
    models.reset_default_graph()
    B, ksize = 4, 6
    rng = np.random.RandomState(5)
    img = rng.uniform(-1, 1, (B, 64, 64, 3)).astype(np.float32)
    act = rng.randn(B, 10).astype(np.float32)
    r, rs = rng.randn(B, 64, 64, 3).astype(np.float32), rng.randn(B, 5).astype(np.float32)
    dev = lambda a: torch.from_numpy(a).to(cuda)
    frame, state = models.build_generator_transform(dev(img), dev(act), batch_size=B, ksize=ksize)
    ((frame * dev(r)).sum() + (state * dev(rs)).sum()).backward()
    store = models.VARIABLES["g"]
    flat_grad = models.trainable("g").grad
    assert flat_grad is not None and flat_grad.shape == store.flat.shape
    run = [v for k, v in T._RUNS.items() if k[0] == "g_dna" and k[1] == store.flat.data_ptr()][0]
    torch_ref.GATES = device_gates(run)
    torch_ref.reset_gate_calls()
    p = {k: _t64(v.cpu().numpy(), True) for k, v in store.views.items()}
    f_ref, s_ref, _ = torch_ref.generator_transform(p, _t64(img), _t64(act), ksize)
    ((f_ref * _t64(r)).sum() + (s_ref * _t64(rs)).sum()).backward()
    assert np.abs(frame.detach().cpu().numpy() - f_ref.detach().numpy()).max() <= 2e-2
    for name, (off, n, shape) in store.offsets.items():
        got = flat_grad[off:off + n].view(shape).cpu().numpy()
        ref = p[name].grad.numpy()
        rel = np.linalg.norm(got - ref) / max(np.linalg.norm(ref), 1e-9)
        assert rel <= 5e-2, (name, rel)
    # discriminator: parameters AND input gradient
    x_in = torch.cat([dev(img), frame.detach()], 3).requires_grad_(True)
    rd = rng.randn(B, 2, 2, 1).astype(np.float32)
    logits = models.build_discriminator(x_in, dev(act))
    (logits * dev(rd)).sum().backward()
    dstore = models.VARIABLES["d"]
    drun = [v for k, v in T._RUNS.items() if k[0] == "d" and k[1] == dstore.flat.data_ptr()][0]
    torch_ref.GATES = device_gates(drun)
    torch_ref.reset_gate_calls()
    pd = {k: _t64(v.cpu().numpy(), True) for k, v in dstore.views.items()}
    xr = _t64(x_in.detach().cpu().numpy(), True)
    lr = torch_ref.discriminator(pd, xr, _t64(act))
    (lr * _t64(rd)).sum().backward()
    rel = np.linalg.norm(x_in.grad.cpu().numpy() - xr.grad.numpy()) / np.linalg.norm(xr.grad.numpy())
    assert rel <= 5e-2, rel
    dgrad = models.trainable("d").grad
    for name, (off, n, shape) in dstore.offsets.items():
        got = dgrad[off:off + n].view(shape).cpu().numpy()
        ref = pd[name].grad.numpy()
        rel = np.linalg.norm(got - ref) / max(np.linalg.norm(ref), 1e-9)
        assert rel <= 5e-2, (name, rel)
    models.reset_default_graph()
