"""DNA transform (models.py:60-72): CUDA kernels through the C-ABI vs the CPU oracle (fp64)."""
import numpy as np
import pytest
import torch

from oracle import np_ref

pytestmark = pytest.mark.gpu


def _inputs(B, H, W, K, seed, scale=3.0):
    rng = np.random.RandomState(seed)
    logits = (rng.randn(B, H, W, K * K) * scale).astype(np.float32)
    img = rng.uniform(-1, 1, size=(B, H, W, 3)).astype(np.float32)
    dy = rng.randn(B, H, W, 3).astype(np.float32)
    return logits, img, dy


@pytest.mark.parametrize("rows", ["2", "4"])     # both band heights of the kernel (ACG_DNA_ROWS; default: by batch size)
@pytest.mark.parametrize("K", [5, 6])
@pytest.mark.parametrize("B,H,W", [(1, 4, 4), (2, 8, 16), (3, 64, 64), (5, 12, 64)])
def test_dna_fwd_bwd_vs_oracle(cuda, monkeypatch, rows, K, B, H, W):
    from action_conditioned_gans_b200 import kernels as Kn
    monkeypatch.setenv("ACG_DNA_ROWS", rows)
    logits, img, dy = _inputs(B, H, W, K, seed=100 + K + B)
    ref = np_ref.dna_forward(logits.astype(np.float64), img.astype(np.float64), K)
    ref_b = np_ref.dna_backward(logits.astype(np.float64), img.astype(np.float64), dy.astype(np.float64), K)
    tl, ti, td = (torch.from_numpy(a).to(cuda) for a in (logits, img, dy))
    out = torch.full((B, H, W, 3), float("nan"), device=cuda)
    dl = torch.full((B, H, W, K * K), float("nan"), device=cuda)
    Kn.dna_fwd(tl, ti, out, K)
    Kn.dna_bwd(tl, ti, td, dl, K)
    torch.cuda.synchronize()
    got, got_b = out.cpu().numpy(), dl.cpu().numpy()
    # north star: DNA outputs within 1e-5 relative error in fp32
    assert np.abs(got - ref).max() <= 1e-5 * max(1.0, np.abs(ref).max())
    assert np.abs(got_b - ref_b).max() <= 1e-5 * max(1.0, np.abs(ref_b).max())
    # softmax-Jacobian identity: the logit gradient of every pixel sums to zero
    assert np.abs(got_b.sum(-1)).max() < 1e-5


@pytest.mark.parametrize("K", [5, 6])
def test_dna_analytic_cases(cuda, K):
    from action_conditioned_gans_b200 import kernels as Kn
    B, H, W = 2, 16, 16
    rng = np.random.RandomState(3)
    img = rng.uniform(-1, 1, size=(B, H, W, 3)).astype(np.float32)
    ti = torch.from_numpy(img).to(cuda)
    out = torch.empty(B, H, W, 3, device=cuda)
    pb = (K - 1) // 2
    # uniform logits -> K x K box filter of the zero padded frame
    Kn.dna_fwd(torch.zeros(B, H, W, K * K, device=cuda), ti, out, K)
    pad = np.zeros((B, H + K - 1, W + K - 1, 3), np.float64)
    pad[:, pb:pb + H, pb:pb + W] = img
    box = sum(pad[:, a:a + H, b:b + W] for a in range(K) for b in range(K)) / (K * K)
    assert np.abs(out.cpu().numpy() - box).max() < 1e-6
    # one-hot logits -> the frame shifted by the selected tap (zero outside)
    for (a, b) in [(0, 0), (pb, pb), (K - 1, K - 1), (1, K - 2)]:
        lg = torch.full((B, H, W, K * K), -1e4, device=cuda)
        lg[..., a * K + b] = 0.0
        Kn.dna_fwd(lg, ti, out, K)
        assert np.abs(out.cpu().numpy() - pad[:, a:a + H, b:b + W]).max() < 1e-6
    # convex combination: output stays inside [min, max] of the zero padded input
    lg = torch.randn(B, H, W, K * K, device=cuda) * 5
    Kn.dna_fwd(lg, ti, out, K)
    o = out.cpu().numpy()
    assert o.max() <= max(img.max(), 0) + 1e-6 and o.min() >= min(img.min(), 0) - 1e-6


@pytest.mark.parametrize("K", [5, 6])
def test_dna_bf16_logits(cuda, K):
    """bf16 logits / dlogits storage (what the tensor-core tconv4 epilogue can emit)."""
    from action_conditioned_gans_b200 import kernels as Kn
    B, H, W = 2, 16, 32
    logits, img, dy = _inputs(B, H, W, K, seed=11)
    tl = torch.from_numpy(logits).to(cuda).to(torch.bfloat16)
    lq = tl.float().cpu().numpy().astype(np.float64)
    ref = np_ref.dna_forward(lq, img.astype(np.float64), K)
    ref_b = np_ref.dna_backward(lq, img.astype(np.float64), dy.astype(np.float64), K)
    ti, td = torch.from_numpy(img).to(cuda), torch.from_numpy(dy).to(cuda)
    out = torch.empty(B, H, W, 3, device=cuda)
    dl = torch.empty(B, H, W, K * K, device=cuda, dtype=torch.bfloat16)
    Kn.dna_fwd(tl, ti, out, K)
    Kn.dna_bwd(tl, ti, td, dl, K)
    assert np.abs(out.cpu().numpy() - ref).max() <= 1e-5
    assert np.abs(dl.float().cpu().numpy() - ref_b).max() <= 1e-2 * np.abs(ref_b).max()   # bf16 storage


def test_dna_full_size_properties(cuda):
    """BASELINE config 2 sizes (B=64 and B=256, 64x64x3): size-independent properties."""
    from action_conditioned_gans_b200 import kernels as Kn
    for K, B in [(5, 64), (6, 64), (5, 256)]:
        g = torch.Generator(device=cuda).manual_seed(5)
        lg = torch.randn(B, 64, 64, K * K, device=cuda, generator=g) * 2
        img = torch.rand(B, 64, 64, 3, device=cuda, generator=g) * 2 - 1
        dy = torch.randn(B, 64, 64, 3, device=cuda, generator=g)
        out = torch.empty(B, 64, 64, 3, device=cuda)
        dl = torch.empty_like(lg)
        Kn.dna_fwd(lg, img, out, K)
        Kn.dna_bwd(lg, img, dy, dl, K)
        assert torch.isfinite(out).all() and torch.isfinite(dl).all()
        assert out.max() <= 1 + 1e-5 and out.min() >= -1 - 1e-5          # convex combination of [-1,1] / 0
        assert dl.sum(-1).abs().max() < 1e-4                              # rows of the softmax Jacobian
        # shift invariance of softmax: adding a per-pixel constant to the logits changes nothing
        out2 = torch.empty_like(out)
        Kn.dna_fwd(lg + 3.0, img, out2, K)
        assert (out - out2).abs().max() < 1e-5
        # linearity in the frame
        out3 = torch.empty_like(out)
        Kn.dna_fwd(lg, img * 0.5, out3, K)
        assert (out3 - 0.5 * out).abs().max() < 1e-6
        # spot-check 2 samples against the oracle
        idx = [0, B - 1]
        ref = np_ref.dna_forward(lg[idx].double().cpu().numpy(), img[idx].double().cpu().numpy(), K)
        assert np.abs(out[idx].cpu().numpy() - ref).max() <= 1e-5


def test_dna_rejects_bad_arguments(cuda):
    from action_conditioned_gans_b200 import kernels as Kn
    lg = torch.zeros(1, 8, 8, 16, device=cuda)
    img = torch.zeros(1, 8, 8, 3, device=cuda)
    out = torch.zeros(1, 8, 8, 3, device=cuda)
    with pytest.raises(RuntimeError):
        Kn.dna_fwd(lg, img, out, 4)                      # K must be 5 or 6
    with pytest.raises(RuntimeError):
        Kn.dna_fwd(lg.cpu(), img, out, 5)                # no CPU path


@pytest.mark.parametrize("rows", ["2", "4"])
@pytest.mark.parametrize("K", [5, 6])
def test_dna_bwd_padded_bf16_output(cuda, monkeypatch, rows, K):
    """dlogits written as bf16 rows of ru16(K*K) channels with zero pad channels (the dz operand of g/tconv4)."""
    from action_conditioned_gans_b200 import kernels as Kn
    monkeypatch.setenv("ACG_DNA_ROWS", rows)
    B, H, W = 3, 16, 64
    logits, img, dy = _inputs(B, H, W, K, seed=21)
    ref_b = np_ref.dna_backward(logits.astype(np.float64), img.astype(np.float64), dy.astype(np.float64), K)
    tl, ti, td = (torch.from_numpy(a).to(cuda) for a in (logits, img, dy))
    ld = (K * K + 15) // 16 * 16
    dl = torch.full((B, H, W, ld), float("nan"), device=cuda, dtype=torch.bfloat16)
    Kn.dna_bwd(tl, ti, td, dl, K)
    got = dl.float().cpu().numpy()
    assert np.abs(got[..., :K * K] - ref_b).max() <= 8e-3 * np.abs(ref_b).max()      # bf16 storage
    assert np.abs(got[..., K * K:]).max() == 0.0
