"""Generator / discriminator forward + hand-scheduled backward (engine.py) vs autograd of the CPU oracle."""
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


def _params(spec, seed):
    rng = np.random.RandomState(seed)
    p = np_ref.init_params(spec, rng)
    for k in p:
        if not k.endswith("weights"):
            p[k] = (rng.randn(*p[k].shape) * 0.1).astype(np.float32)
    return p


def _report(got, ref, tol, l2=False):
    """fp32 path: max-norm per variable.  bf16 path: relative Frobenius norm per variable (bf16 operands put ~0.4 %
    rounding noise on every activation; accumulation is fp32)."""
    bad = []
    for k, r in ref.items():
        if l2:
            d, s = np.linalg.norm(got[k] - r), max(np.linalg.norm(r), 1e-9)
        else:
            d, s = np.abs(got[k] - r).max(), max(np.abs(r).max(), 1e-6)
        if d > tol * s:
            bad.append("%s: diff %.3g at scale %.3g" % (k, d, s))
    assert not bad, "\n".join(bad)


PREC = [("fp32", 2e-4, 1e-4), ("bf16", 5e-2, 5e-2)]     # (precision, gradient tolerance, forward tolerance)


@pytest.mark.parametrize("prec,gtol,ftol", PREC)
@pytest.mark.parametrize("B", [2, 5])
def test_discriminator_fwd_bwd(cuda, B, prec, gtol, ftol):
    from action_conditioned_gans_b200 import engine as E
    p = _params(np_ref.d_spec(), 3)
    rng = np.random.RandomState(B)
    img = rng.uniform(-1, 1, (B, 64, 64, 3)).astype(np.float32)
    frame = rng.uniform(-1, 1, (B, 64, 64, 3)).astype(np.float32)
    act = rng.randn(B, 10).astype(np.float32)
    r = rng.randn(B, 2, 2, 1).astype(np.float32)
    # engine
    store = E.ParamStore(E.d_spec(), cuda, p)
    run = E.DiscriminatorRun(store, B, cuda, precision=prec)
    store.refresh_packs()
    t = lambda a: torch.from_numpy(a).to(cuda)
    out = run.forward(t(img), t(frame), t(act))
    # oracle, evaluated on the linear pieces the device took (see torch_ref.GATES)
    torch_ref.GATES = device_gates(run)
    torch_ref.reset_gate_calls()
    pt = {k: torch.tensor(v, dtype=torch.float64, requires_grad=True) for k, v in p.items()}
    xin = torch.tensor(np.concatenate([img, frame], 3), dtype=torch.float64, requires_grad=True)
    logits = torch_ref.discriminator(pt, xin, torch.tensor(act, dtype=torch.float64))
    names = list(pt)
    grads = torch.autograd.grad((logits * torch.tensor(r, dtype=torch.float64)).sum(), [pt[k] for k in names] + [xin])
    gref = {k: g.numpy() for k, g in zip(names, grads[:-1])}
    assert np.abs(out.cpu().numpy() - logits.detach().numpy()).max() < ftol * max(1.0, float(logits.abs().max()))
    run.dlogits.copy_(t(r).reshape(-1))
    store.grad.zero_()
    dx = run.backward(need_dw=True, need_dinput=True)
    _report(store.grads_numpy(), gref, gtol, l2=prec == "bf16")
    _report({"d_in": dx.cpu().numpy()[..., :6]}, {"d_in": grads[-1].numpy()}, gtol, l2=prec == "bf16")


@pytest.mark.parametrize("prec,gtol,ftol", PREC)
@pytest.mark.parametrize("dna,ksize", [(True, 6), (True, 5), (False, 5)])
def test_generator_fwd_bwd(cuda, dna, ksize, prec, gtol, ftol):
    from action_conditioned_gans_b200 import engine as E
    B = 3
    spec = np_ref.g_dna_spec(ksize) if dna else np_ref.g_direct_spec()
    p = _params(spec, 4)
    rng = np.random.RandomState(1)
    img = rng.uniform(-1, 1, (B, 64, 64, 3)).astype(np.float32)
    act = rng.randn(B, 10).astype(np.float32)
    r = rng.randn(B, 64, 64, 3).astype(np.float32)
    rs = rng.randn(B, 5).astype(np.float32)
    store = E.ParamStore(E.g_dna_spec(ksize) if dna else E.g_direct_spec(), cuda, p)
    run = E.GeneratorRun(store, B, cuda, dna, ksize, precision=prec)
    store.refresh_packs()
    t = lambda a: torch.from_numpy(a).to(cuda)
    g_out, g_state = run.forward(t(img), t(act))
    torch_ref.GATES = device_gates(run)
    torch_ref.reset_gate_calls()
    pt = {k: torch.tensor(v, dtype=torch.float64, requires_grad=True) for k, v in p.items()}
    ti, ta = torch.tensor(img, dtype=torch.float64), torch.tensor(act, dtype=torch.float64)
    if dna:
        frame, state, logits = torch_ref.generator_transform(pt, ti, ta, ksize)
        loss = (frame * torch.tensor(r, dtype=torch.float64)).sum() + (state * torch.tensor(rs, dtype=torch.float64)).sum()
    else:
        frame = torch_ref.generator_direct(pt, ti, ta)
        loss = (frame * torch.tensor(r, dtype=torch.float64)).sum()
    names = list(pt)
    grads = torch.autograd.grad(loss, [pt[k] for k in names])
    gref = {k: g.numpy() for k, g in zip(names, grads)}
    assert np.abs(g_out.cpu().numpy() - frame.detach().numpy()).max() < ftol
    if dna:
        assert np.abs(g_state.cpu().numpy() - state.detach().numpy()).max() < ftol * max(1.0, float(state.abs().max()))
        run.dstate.copy_(t(rs))
    run.dg_out.copy_(t(r))
    store.grad.zero_()
    run.backward(with_state=dna)
    _report(store.grads_numpy(), gref, gtol, l2=prec == "bf16")
