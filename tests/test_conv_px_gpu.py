"""Pixel-major tcgen05 kernel for small feature maps (csrc/conv_px.cu): one GEMM tile = one output pixel x 128 images,
both operands by TMA, out-of-image taps skipped.  Checked against the fp32 SIMT convolution on bf16-rounded operands,
against the generic tcgen05 kernel (ACG_NO_PX=1), with split-K, ragged batches, channel strides that are not a multiple
of 64, fused batch-norm moments and the in-kernel finalize.  Layers: models.py:36-41 (g/conv4, g/tconv1), :82-87
(d/conv4, d/conv5, d/conv6), state head :42-50."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def ru(v, m):
    return (v + m - 1) // m * m


# (B, H, W, Cin, Cout, k, stride, padding)
CASES = [
    (256, 8, 8, 128, 256, 5, 2, "SAME"),    # g/conv4, d/conv4 (fwd: 4x4 grid; dgrad: 8x8 grid)
    (256, 4, 4, 256, 512, 5, 2, "SAME"),    # d/conv5 (fwd: 2x2 grid, split-K; dgrad: 4x4 grid)
    (200, 8, 8, 266, 128, 5, 2, "SAME"),    # adjoint of g/tconv1 with the action concat (266 -> ld 272), ragged batch
    (64, 8, 8, 32, 16, 3, 2, "SAME"),       # g/sconv4: half a batch tile, N = 16, 3x3 taps
    (130, 4, 4, 16, 5, 4, 1, "VALID"),      # g/sconv5: 4x4 VALID -> 1x1
    (128, 2, 2, 512, 1, 2, 1, "SAME"),      # d/conv6: k=2 stride 1
    (96, 16, 16, 138, 128, 5, 2, "SAME"),   # d/conv3: ld 144 (last K slice of a tap holds 16 channels); dgrad not pixel-major
    (64, 6, 10, 40, 48, 5, 2, "SAME"),      # odd grid (3x5), ragged channels
]


def _operands(case, cuda, seed):
    from action_conditioned_gans_b200 import kernels as Kn
    B, H, W, Cin, Cout, k, s, padding = case
    g = torch.Generator(device=cuda).manual_seed(seed)
    shape = Kn.conv_shape(B, H, W, Cin, Cout, k, s, padding)
    w = (torch.randn(k, k, Cin, Cout, device=cuda, generator=g) / (k * Cin ** 0.5)).to(torch.bfloat16).float()
    return shape, w, g


@pytest.mark.parametrize("case", CASES)
def test_pixel_major_forward(cuda, case, monkeypatch):
    from action_conditioned_gans_b200 import kernels as Kn
    B, H, W, Cin, Cout, k, s, padding = case
    shape, w, g = _operands(case, cuda, 3)
    ld_in, ld_out = ru(Cin, 16), ru(Cout, 16)
    assert Kn.kernel_kind(shape, 0, ld_in) == 2
    x = torch.zeros(B, H, W, ld_in, dtype=torch.bfloat16, device=cuda)
    x[..., :Cin] = torch.randn(B, H, W, Cin, device=cuda, generator=g).to(torch.bfloat16)
    ref = torch.empty(B, shape.OH, shape.OW, Cout, device=cuda)
    Kn.conv_fprop_f32(shape, x[..., :Cin].float().contiguous(), w, ref)
    pack = torch.empty(Kn.pack_size(shape, 0, ld_in), dtype=torch.bfloat16, device=cuda)
    Kn.pack_weights(shape, w, 0, ld_in, pack)
    rows = B * shape.OH * shape.OW
    beta = torch.randn(Cout, device=cuda, generator=g)
    wsp = Kn.splitk_workspace(shape, 0, ld_in, cuda)
    res = {}
    for mode in ("px", "px_nosplit", "generic"):
        if mode == "generic":
            monkeypatch.setenv("ACG_NO_PX", "1")
        out = torch.full((rows, ld_out), 7.0, dtype=torch.bfloat16, device=cuda)
        stats = torch.zeros(2 * Cout, dtype=torch.float64, device=cuda)
        fix = Kn.stats_accumulators(Cout, cuda)
        counter = torch.zeros(1, dtype=torch.int32, device=cuda)
        mean, rstd, scale, shift = (torch.zeros(Cout, device=cuda) for _ in range(4))
        for rep in range(2):     # twice: tickets, counter and accumulators must be left ready
            stats.zero_()
            Kn.conv_fprop_tc(shape, x, pack, out, ld_in, ld_out, stats=stats,
                             bn=(counter, beta, mean, rstd, scale, shift, rows, 1e-3),
                             splitk=None if mode == "px_nosplit" else wsp, stats_fix=fix)
        torch.cuda.synchronize()
        assert int(counter.item()) == 0 and int(fix.abs().sum()) == 0
        if wsp is not None:
            assert int(wsp[1].abs().sum()) == 0
        res[mode] = (out.float(), stats.clone(), torch.stack([mean, rstd, shift]))
    monkeypatch.delenv("ACG_NO_PX")
    sc = max(1.0, float(ref.abs().max()))
    for mode, (out, stats, fin) in res.items():
        got = out.view(B, shape.OH, shape.OW, ld_out)
        assert float((got[..., :Cout] - ref).abs().max()) <= 1.2e-2 * sc, mode
        assert float(got[..., Cout:].abs().max()) == 0.0 if ld_out > Cout else True
        o = out[:, :Cout].double()
        tot = torch.cat([o.sum(0), (o * o).sum(0)])
        assert float((stats - tot).abs().max()) <= 2e-6 * max(1.0, float(tot.abs().max())), mode
        mu = tot[:Cout] / rows
        assert float((fin[0].double() - mu).abs().max()) <= 1e-5 * max(1.0, float(mu.abs().max())), mode
    # the three paths agree up to fp32 summation order (one bf16 ulp where a rounding flips)
    a, b = res["px"][0], res["generic"][0]
    assert float((a - b).abs().max()) <= 1e-2 * sc and float((a != b).float().mean()) < 0.02
    # unsplit pixel-major == generic bit for bit when taps are whole numbers of K=16 steps: same products, same order
    if ld_in % 64 == 0 and Kn.splitk_workspace(shape, 0, ld_in, cuda) is None:
        assert torch.equal(res["px_nosplit"][0], res["generic"][0])
    # fp32 output with bias and tanh (the fused-output path of g/sconv5 / direct generator heads)
    bias = torch.randn(Cout, device=cuda, generator=g)
    out32 = torch.zeros(rows, ld_out, device=cuda)
    Kn.conv_fprop_tc(shape, x, pack, out32, ld_in, ld_out, bias=bias, out_act="tanh", splitk=wsp)
    want = torch.tanh(ref + bias)
    assert float((out32.view(B, shape.OH, shape.OW, ld_out)[..., :Cout] - want).abs().max()) <= 4e-3


@pytest.mark.parametrize("case", CASES[:6] + CASES[7:])
def test_pixel_major_adjoint(cuda, case, monkeypatch):
    """Data gradient of conv2d == forward of conv2d_transpose (ADJ gather: one weight matrix per output-parity class)."""
    from action_conditioned_gans_b200 import kernels as Kn
    B, H, W, Cin, Cout, k, s, padding = case
    shape, w, g = _operands(case, cuda, 5)
    ld_in, ld_out = ru(Cout, 16), ru(Cin, 16)
    assert Kn.kernel_kind(shape, 1, ld_in) == 2
    dy = torch.zeros(B, shape.OH, shape.OW, ld_in, dtype=torch.bfloat16, device=cuda)
    dy[..., :Cout] = torch.randn(B, shape.OH, shape.OW, Cout, device=cuda, generator=g).to(torch.bfloat16)
    ref = torch.empty(B, H, W, Cin, device=cuda)
    Kn.conv_dgrad_f32(shape, dy[..., :Cout].float().contiguous(), w, ref)
    pack = torch.empty(Kn.pack_size(shape, 1, ld_in), dtype=torch.bfloat16, device=cuda)
    Kn.pack_weights(shape, w, 1, ld_in, pack)
    wsp = Kn.splitk_workspace(shape, 1, ld_in, cuda)
    rows = B * H * W
    sc = max(1.0, float(ref.abs().max()))
    outs = {}
    for mode in ("px", "generic"):
        if mode == "generic":
            monkeypatch.setenv("ACG_NO_PX", "1")
        dx = torch.full((rows, ld_out), float("nan"), dtype=torch.bfloat16, device=cuda)
        stats = torch.zeros(2 * Cin, dtype=torch.float64, device=cuda)
        counter = torch.zeros(1, dtype=torch.int32, device=cuda)
        fix = Kn.stats_accumulators(Cin, cuda)
        for rep in range(2):
            stats.zero_()
            Kn.conv_dgrad_tc(shape, dy, pack, dx, ld_in, ld_out, stats=stats,
                             bn=(counter, None, None, None, None, None, 0, 1e-3), splitk=wsp, stats_fix=fix)
        torch.cuda.synchronize()
        got = dx.float().view(B, H, W, ld_out)
        assert bool(torch.isfinite(got).all()), mode
        assert float((got[..., :Cin] - ref).abs().max()) <= 1.2e-2 * sc, mode
        if ld_out > Cin:
            assert float(got[..., Cin:].abs().max()) == 0.0
        o = dx[:, :Cin].double()
        tot = torch.cat([o.sum(0), (o * o).sum(0)])
        assert float((stats - tot).abs().max()) <= 2e-6 * max(1.0, float(tot.abs().max())), mode
        outs[mode] = got
    monkeypatch.delenv("ACG_NO_PX")
    assert float((outs["px"] - outs["generic"]).abs().max()) <= 1e-2 * sc
    # fp32 gradient restricted to the first channels (n_limit: gradient w.r.t. the feature part of a concat buffer)
    nlim = min(Cin, 16) if Cin > 16 else 0
    if nlim:
        dx32 = torch.full((rows, ld_out), float("nan"), device=cuda)
        Kn.conv_dgrad_tc(shape, dy, pack, dx32, ld_in, ld_out, splitk=wsp, n_limit=nlim)
        got = dx32.view(B, H, W, ld_out)
        assert float((got[..., :nlim] - ref[..., :nlim]).abs().max()) <= 2e-3 * sc
        assert bool(torch.isnan(got[..., nlim:]).all())


def test_small_batches_keep_the_generic_kernel(cuda):
    from action_conditioned_gans_b200 import kernels as Kn
    assert Kn.kernel_kind(Kn.conv_shape(16, 8, 8, 128, 256, 5, 2, "SAME"), 0, 128) != 2      # configs[0]: batch 16
    assert Kn.kernel_kind(Kn.conv_shape(256, 32, 32, 64, 128, 5, 2, "SAME"), 0, 64) != 2     # 16x16 output grid
