"""tcgen05 implicit-GEMM convolution kernels vs the CPU oracle on bf16-rounded operands (fp32 accumulation, so
the only difference left is summation order): conv2d forward, conv2d_transpose forward (== dgrad), both data
gradients, channel padding / concat strides, ragged tiles."""
import numpy as np
import pytest
import torch

from oracle import np_ref, torch_ref

pytestmark = pytest.mark.gpu


def ru(v, m):
    return (v + m - 1) // m * m


def _bf16_round(a):
    return torch.from_numpy(a.astype(np.float32)).to(torch.bfloat16).float().numpy().astype(np.float64)


def _pad_channels(a, ld, cuda):
    """[B,H,W,C] float64 -> bf16 CUDA tensor [B,H,W,ld] with zero pad channels"""
    B, H, W, C = a.shape
    t = torch.zeros(B, H, W, ld, dtype=torch.bfloat16, device=cuda)
    t[..., :C] = torch.from_numpy(a.astype(np.float32)).to(cuda).to(torch.bfloat16)
    return t


# (B, H, W, Cin, Cout, k, stride, padding)
CASES = [
    (2, 8, 8, 64, 64, 1, 1, "SAME"),       # plain GEMM through the conv path: M=128, K=64, N=64
    (2, 16, 16, 64, 128, 5, 2, "SAME"),    # g/conv3-like
    (3, 64, 64, 3, 32, 5, 2, "SAME"),      # g/conv1: Cin 3 padded to 16
    (2, 16, 16, 138, 128, 5, 2, "SAME"),   # d/conv3: concat channels 138 -> ld 144
    (5, 8, 8, 128, 256, 5, 2, "SAME"),     # d/conv4: two N tiles, partial M tile
    (2, 64, 64, 36, 128, 5, 2, "SAME"),    # adjoint of g/tconv4 (ksize 6)
    (3, 8, 8, 272, 128, 5, 2, "SAME"),     # adjoint of g/tconv1 with the action concat (266 -> 272)
    (2, 16, 16, 128, 32, 3, 2, "SAME"),    # g/sconv3
    (1, 7, 9, 24, 40, 5, 2, "SAME"),       # odd extents, ragged N
    (4, 32, 32, 32, 64, 5, 2, "SAME"),     # g/conv2: its dgrad takes the halo-tile kernel (16x16 class grid, 2 images per CTA)
    (3, 64, 64, 6, 64, 5, 2, "SAME"),      # d/conv1: halo dgrad with N=16, one 64-channel block
    (2, 32, 32, 128, 128, 3, 2, "SAME"),   # 3x3 stride 2 (pad 0/1): 2x2 / 2x1 / 1x2 / 1x1 taps per class in the halo kernel
    (2, 64, 64, 128, 192, 5, 2, "SAME"),   # three 64-channel blocks: the halo double buffer is recycled
]


@pytest.mark.parametrize("case", CASES)
@pytest.mark.parametrize("out_dtype", [torch.float32, torch.bfloat16])
def test_conv_fprop_tc(cuda, case, out_dtype):
    from action_conditioned_gans_b200 import kernels as Kn
    B, H, W, Cin, Cout, k, s, padding = case
    rng = np.random.RandomState(sum(case[:7]))
    x = _bf16_round(rng.randn(B, H, W, Cin))
    w = rng.randn(k, k, Cin, Cout) / np.sqrt(k * k * Cin)
    bias = rng.randn(Cout)
    shape = Kn.conv_shape(B, H, W, Cin, Cout, k, s, padding)
    ld_in, ld_out = ru(Cin, 16), ru(Cout, 16) + 16
    y_ref = np_ref.conv2d(x, _bf16_round(w), s, padding) + bias
    wt = torch.from_numpy(w.astype(np.float32)).to(cuda)
    pack = torch.empty(Kn.pack_size(shape, 0, ld_in), dtype=torch.bfloat16, device=cuda)
    Kn.pack_weights(shape, wt, 0, ld_in, pack)
    y = torch.full((B, shape.OH, shape.OW, ld_out), 7.0, dtype=out_dtype, device=cuda)
    Kn.conv_fprop_tc(shape, _pad_channels(x, ld_in, cuda), pack, y, ld_in, ld_out,
                     bias=torch.from_numpy(bias.astype(np.float32)).to(cuda))
    torch.cuda.synchronize()
    got = y.float().cpu().numpy()
    tol = 2e-3 if out_dtype == torch.float32 else 1.2e-2
    assert np.abs(got[..., :Cout] - y_ref).max() <= tol * max(1.0, np.abs(y_ref).max())
    assert np.abs(got[..., Cout:ru(Cout, 16)]).max(initial=0.0) == 0.0      # pad channels are exact zeros
    assert (got[..., ru(Cout, 16):] == 7.0).all()                            # beyond ru16(Cout): untouched


@pytest.mark.parametrize("case", CASES)
def test_conv_dgrad_tc(cuda, case):
    """dgrad == conv2d_transpose forward: checked against the independent torch restatement's autograd."""
    from action_conditioned_gans_b200 import kernels as Kn
    B, H, W, Cin, Cout, k, s, padding = case
    rng = np.random.RandomState(sum(case[:7]) + 1)
    shape = Kn.conv_shape(B, H, W, Cin, Cout, k, s, padding)
    dy = _bf16_round(rng.randn(B, shape.OH, shape.OW, Cout))
    w = rng.randn(k, k, Cin, Cout) / np.sqrt(k * k * Cout)
    xt = torch.zeros(B, H, W, Cin, dtype=torch.float64, requires_grad=True)
    yt = torch_ref.conv2d(xt, torch.tensor(_bf16_round(w)), s, padding)
    (dx_ref,) = torch.autograd.grad(yt, [xt], torch.tensor(dy))
    dx_ref = dx_ref.numpy()
    ld_in, ld_out = ru(Cout, 8), ru(Cin, 16)
    wt = torch.from_numpy(w.astype(np.float32)).to(cuda)
    pack = torch.empty(Kn.pack_size(shape, 1, ld_in), dtype=torch.bfloat16, device=cuda)
    Kn.pack_weights(shape, wt, 1, ld_in, pack)
    dx = torch.full((B, H, W, ld_out), float("nan"), dtype=torch.float32, device=cuda)
    Kn.conv_dgrad_tc(shape, _pad_channels(dy, ld_in, cuda), pack, dx, ld_in, ld_out)
    torch.cuda.synchronize()
    got = dx.cpu().numpy()
    assert np.isfinite(got).all()
    assert np.abs(got[..., :Cin] - dx_ref).max() <= 2e-3 * max(1.0, np.abs(dx_ref).max())
    assert np.abs(got[..., Cin:]).max(initial=0.0) == 0.0


def test_conv_tc_matches_simt_at_full_size(cuda):
    """Largest layer of the DNA generator (tconv3 as a dgrad, B=64): tensor-core path vs the fp32 SIMT kernel."""
    from action_conditioned_gans_b200 import kernels as Kn
    B = 64
    shape = Kn.conv_shape(B, 32, 32, 128, 128, 5, 2, "SAME")
    g = torch.Generator(device=cuda).manual_seed(1)
    dy = torch.randn(B, 16, 16, 128, device=cuda, generator=g).to(torch.bfloat16)
    w = (torch.randn(5, 5, 128, 128, device=cuda, generator=g) / 40).to(torch.bfloat16).float()
    ref = torch.empty(B, 32, 32, 128, device=cuda)
    Kn.conv_dgrad_f32(shape, dy.float(), w, ref)
    pack = torch.empty(Kn.pack_size(shape, 1, 128), dtype=torch.bfloat16, device=cuda)
    Kn.pack_weights(shape, w, 1, 128, pack)
    out = torch.empty(B, 32, 32, 128, device=cuda)
    Kn.conv_dgrad_tc(shape, dy, pack, out, 128, 128)
    assert (out - ref).abs().max() <= 2e-3 * ref.abs().max()


@pytest.mark.parametrize("case", CASES + [(16, 32, 32, 32, 64, 5, 2, "SAME"), (4, 32, 32, 128, 48, 5, 2, "SAME")])
def test_conv_wgrad_tc(cuda, case):
    """dW[a,c,ci,co] = sum x*dy (MN-major operands, split over pixels with fp32 atomics) vs oracle autograd."""
    from action_conditioned_gans_b200 import kernels as Kn
    B, H, W, Cin, Cout, k, s, padding = case
    rng = np.random.RandomState(sum(case[:7]) + 2)
    shape = Kn.conv_shape(B, H, W, Cin, Cout, k, s, padding)
    x = _bf16_round(rng.randn(B, H, W, Cin))
    dy = _bf16_round(rng.randn(B, shape.OH, shape.OW, Cout))
    wt = torch.zeros(k, k, Cin, Cout, dtype=torch.float64, requires_grad=True)
    yt = torch_ref.conv2d(torch.tensor(x), wt, s, padding)
    (dw_ref,) = torch.autograd.grad(yt, [wt], torch.tensor(dy))
    dw_ref = dw_ref.numpy()
    ld_x, ld_dy = ru(Cin, 16), ru(Cout, 8)
    dw = torch.zeros(k, k, Cin, Cout, device=cuda)
    Kn.conv_wgrad_tc(shape, _pad_channels(x, ld_x, cuda), _pad_channels(dy, ld_dy, cuda), dw, ld_x, ld_dy)
    torch.cuda.synchronize()
    got = dw.cpu().numpy()
    assert np.abs(got - dw_ref).max() <= 2e-3 * max(1.0, np.abs(dw_ref).max())
    # accumulate semantics: a second call doubles the result
    Kn.conv_wgrad_tc(shape, _pad_channels(x, ld_x, cuda), _pad_channels(dy, ld_dy, cuda), dw, ld_x, ld_dy)
    assert np.abs(dw.cpu().numpy() - 2 * dw_ref).max() <= 4e-3 * max(1.0, np.abs(dw_ref).max())


def test_batched_tile_pack_equals_single_packs(cuda):
    """acg_pack_weights_batched (one launch, tile table from acg_pack_plan, shared-memory transpose) writes exactly
    what one acg_pack_weights call per pack writes -- every layer geometry of models.py, ragged channel counts."""
    from action_conditioned_gans_b200 import kernels as Kn
    rng = np.random.RandomState(11)
    geos = [(64, 64, 3, 32, 5, 2, "SAME", 16, 32), (16, 16, 138, 70, 5, 2, "SAME", 144, 80),
            (16, 16, 128, 32, 3, 2, "SAME", 128, 32), (4, 4, 16, 5, 4, 1, "VALID", 16, 16),
            (2, 2, 512, 1, 2, 1, "SAME", 512, 16), (64, 64, 36, 128, 5, 2, "SAME", 48, 128),
            (8, 8, 266, 128, 5, 2, "SAME", 272, 128)]
    entries, singles = [], []
    for (H, W, Cin, Cout, k, s, pad, ld_ci, ld_co) in geos:
        shape = Kn.conv_shape(2, H, W, Cin, Cout, k, s, pad)
        w = torch.from_numpy(rng.randn(k, k, Cin, Cout).astype(np.float32)).to(cuda)
        for which, ld in ((0, ld_ci), (1, ld_co)):
            n = Kn.pack_size(shape, which, ld)
            batched = torch.full((n,), 7.0, dtype=torch.bfloat16, device=cuda)
            single = torch.full((n,), -7.0, dtype=torch.bfloat16, device=cuda)
            Kn.pack_weights(shape, w, which, ld, single)
            entries.append((shape, w, which, ld, batched))
            singles.append(single)
    table, njobs, tiles, ntiles = Kn.make_pack_jobs(entries, cuda)
    assert njobs == len(entries) and ntiles > njobs
    Kn.pack_weights_batched(table, njobs, tiles, ntiles)
    torch.cuda.synchronize()
    for (shape, w, which, ld, batched), single in zip(entries, singles):
        assert torch.equal(batched.view(torch.int16), single.view(torch.int16)), (shape.Cin, shape.Cout, which)


# (which data-gradient kernel, B, H, W, Cin, Cout, C of the consumer's batch-norm, activation)
RED_CASES = [
    ("dgrad", 3, 16, 16, 32, 64, 32, "relu"),        # generic ADJ kernel (small grid -> 6-stage variant)
    ("dgrad", 5, 8, 8, 138, 128, 128, "lrelu"),      # d/conv3 -> d/conv2: concat gradient, only 128 of 138 channels
    ("dgrad", 4, 32, 32, 128, 128, 128, "relu"),     # TMA halo-tile kernel (H=W=32, Cout=128)
    ("dgrad", 2, 64, 64, 64, 128, 64, "lrelu"),      # halo-tile kernel, 64-wide rows
    ("fprop", 3, 16, 16, 48, 128, 128, "relu"),      # data gradient of a conv2d_transpose = forward conv form
    ("fprop", 2, 8, 8, 272, 128, 128, "relu"),
    ("fprop", 4, 32, 32, 64, 128, 128, "relu"),      # CONV-form halo kernel (g/tconv3's data gradient shape), NACC 2
    ("fprop", 2, 64, 64, 48, 128, 128, "lrelu"),     # CONV-form halo kernel, 48-channel rows (g/tconv4's data gradient)
    ("dgrad", 6, 32, 32, 32, 64, 32, "relu"),        # ADJ halo kernel, N = 32, two images per tile (g/conv2 -> g/conv1)
]


@pytest.mark.parametrize("case", RED_CASES)
def test_fused_bwd_reduction_equals_separate_pass(cuda, case):
    """The batch-norm backward sums accumulated in the epilogue of the data-gradient kernels (acg_tc_args.red_*) equal
    what acg_bn_act_bwd_reduce computes from the stored gradient (fp32 partial sums in a different order)."""
    from action_conditioned_gans_b200 import kernels as Kn
    which, B, H, W, Cin, Cout, Cc, act = case
    g = torch.Generator(device=cuda).manual_seed(3)
    shape = Kn.conv_shape(B, H, W, Cin, Cout, 5, 2, "SAME")
    w = torch.randn(5, 5, Cin, Cout, device=cuda, generator=g) / 30
    if which == "dgrad":        # out = gradient w.r.t. the conv input [B,H,W,Cin]
        ld_in, ld_out, rows, n_out = ru(Cout, 16), ru(Cin, 16), B * H * W, Cin
        src = torch.randn(B, shape.OH, shape.OW, ld_in, device=cuda, generator=g).to(torch.bfloat16)
        pack = torch.empty(Kn.pack_size(shape, 1, ld_in), dtype=torch.bfloat16, device=cuda)
        Kn.pack_weights(shape, w, 1, ld_in, pack)
        fn = Kn.conv_dgrad_tc
    else:                       # out = forward conv output [B,OH,OW,Cout]
        ld_in, ld_out, rows, n_out = ru(Cin, 16), ru(Cout, 16), B * shape.OH * shape.OW, Cout
        src = torch.randn(B, H, W, ld_in, device=cuda, generator=g).to(torch.bfloat16)
        src[..., Cin:] = 0
        pack = torch.empty(Kn.pack_size(shape, 0, ld_in), dtype=torch.bfloat16, device=cuda)
        Kn.pack_weights(shape, w, 0, ld_in, pack)
        fn = Kn.conv_fprop_tc
    assert Cc <= n_out
    ldz = Cc
    z = torch.randn(rows, ldz, device=cuda, generator=g).to(torch.bfloat16)
    mean = torch.randn(Cc, device=cuda, generator=g) * 0.3
    rstd = torch.rand(Cc, device=cuda, generator=g) + 0.5
    shift = torch.randn(Cc, device=cuda, generator=g) * 0.3
    red_fused = torch.zeros(2 * Cc, dtype=torch.float64, device=cuda)
    red_sep = torch.zeros(2 * Cc, dtype=torch.float64, device=cuda)
    out = torch.zeros(rows, ld_out, dtype=torch.bfloat16, device=cuda)
    out_plain = torch.zeros(rows, ld_out, dtype=torch.bfloat16, device=cuda)
    fn(shape, src, pack, out, ld_in, ld_out, red=(red_fused, z, ldz, Cc, act, mean, rstd, shift))
    fn(shape, src, pack, out_plain, ld_in, ld_out)
    assert torch.equal(out, out_plain)                        # the output itself is untouched by the fusion
    Kn.bn_act_bwd_reduce(out, None, ld_out, z, ldz, rows, Cc, 1, mean, rstd, shift, act, red_sep)
    torch.cuda.synchronize()
    a, b = red_fused.cpu().numpy(), red_sep.cpu().numpy()
    scale = np.abs(b).max()
    assert scale > 0
    assert np.abs(a - b).max() <= 2e-5 * scale + 1e-3


@pytest.mark.parametrize("Cin,Cout", [(3, 32), (6, 64)])
def test_first_layers_with_8_channel_frames(cuda, Cin, Cout):
    """g/conv1 / d/conv1 as the engine runs them: frame operands padded to 8 channels only (K = 25 x 8 = 200: K blocks
    straddle taps and the last one is partial), forward, weight gradient (N tile of 72 -> 80 columns) and data
    gradient into an 8-channel fp32 buffer."""
    from action_conditioned_gans_b200 import kernels as Kn
    B, H, W, k, s = 3, 64, 64, 5, 2
    rng = np.random.RandomState(Cin)
    shape = Kn.conv_shape(B, H, W, Cin, Cout, k, s, "SAME")
    x = _bf16_round(rng.randn(B, H, W, Cin))
    w = rng.randn(k, k, Cin, Cout) / np.sqrt(k * k * Cin)
    wt = torch.from_numpy(w.astype(np.float32)).to(cuda)
    xt = torch.tensor(x, requires_grad=True)
    wr = torch.tensor(_bf16_round(w), requires_grad=True)
    yt = torch_ref.conv2d(xt, wr, s, "SAME")
    dy = _bf16_round(rng.randn(*yt.shape))
    dx_ref, dw_ref = torch.autograd.grad(yt, [xt, wr], torch.tensor(dy))
    ld = 8
    x8 = _pad_channels(x, ld, cuda)
    # forward
    pack = torch.empty(Kn.pack_size(shape, 0, ld), dtype=torch.bfloat16, device=cuda)
    Kn.pack_weights(shape, wt, 0, ld, pack)
    y = torch.zeros(B, shape.OH, shape.OW, Cout, device=cuda)
    Kn.conv_fprop_tc(shape, x8, pack, y, ld, Cout)
    assert np.abs(y.cpu().numpy() - yt.detach().numpy()).max() <= 2e-3 * max(1.0, float(yt.abs().max()))
    # weight gradient
    dyp = _pad_channels(dy, Cout, cuda)
    dw = torch.zeros(k, k, Cin, Cout, device=cuda)
    Kn.conv_wgrad_tc(shape, x8, dyp, dw, ld, Cout)
    assert np.abs(dw.cpu().numpy() - dw_ref.numpy()).max() <= 2e-3 * max(1.0, float(dw_ref.abs().max()))
    # data gradient, 8-channel fp32 output (what frame_losses reads for the adversarial term)
    packb = torch.empty(Kn.pack_size(shape, 1, Cout), dtype=torch.bfloat16, device=cuda)
    Kn.pack_weights(shape, wt, 1, Cout, packb)
    dx = torch.full((B, H, W, ld), float("nan"), device=cuda)
    Kn.conv_dgrad_tc(shape, dyp, packb, dx, Cout, ld)
    got = dx.cpu().numpy()
    assert np.isfinite(got).all()
    assert np.abs(got[..., :Cin] - dx_ref.numpy()).max() <= 2e-3 * max(1.0, float(dx_ref.abs().max()))
    assert np.abs(got[..., Cin:]).max() == 0.0


@pytest.mark.parametrize("case", [(4, 4, 4, 256, 512, 5, 2), (64, 4, 4, 256, 512, 5, 2), (16, 8, 8, 128, 256, 5, 2),
                                  (8, 2, 2, 512, 1, 2, 1)])
def test_split_k_equals_unsplit(cuda, case):
    """Few-tile layers (d/conv4, d/conv5, d/conv6 ...) split their K loop over grid.z; the last CTA of a tile adds the
    parked partial tiles and runs the normal epilogue -- output, fused batch-norm moments and the in-kernel finalize
    must equal the unsplit launch (fp32 sums in a different order)."""
    from action_conditioned_gans_b200 import kernels as Kn
    B, H, W, Cin, Cout, k, s = case
    g = torch.Generator(device=cuda).manual_seed(9)
    shape = Kn.conv_shape(B, H, W, Cin, Cout, k, s, "SAME")
    w = torch.randn(k, k, Cin, Cout, device=cuda, generator=g) / (k * Cin ** 0.5)
    for which in (0, 1):
        if which == 0:
            ld_in, n_out, rows = ru(Cin, 16), Cout, B * shape.OH * shape.OW
            src = torch.randn(B, H, W, ld_in, device=cuda, generator=g).to(torch.bfloat16)
            fn = Kn.conv_fprop_tc
        else:
            ld_in, n_out, rows = ru(Cout, 16), Cin, B * H * W
            src = torch.randn(B, shape.OH, shape.OW, ld_in, device=cuda, generator=g).to(torch.bfloat16)
            if Cout < ld_in:
                src[..., Cout:] = 0
            fn = Kn.conv_dgrad_tc
        ld_out = ru(n_out, 16)
        pack = torch.empty(Kn.pack_size(shape, which, ld_in), dtype=torch.bfloat16, device=cuda)
        Kn.pack_weights(shape, w, which, ld_in, pack)
        wsp = Kn.splitk_workspace(shape, which, ld_in, cuda)
        if which == 0:
            assert wsp is not None, "the forward launch of this shape is expected to split"
        if wsp is None:
            continue
        outs, stats_all, fin = [], [], []
        for splitk in (None, wsp):
            out = torch.zeros(rows, ld_out, dtype=torch.bfloat16, device=cuda)
            stats = torch.zeros(2 * n_out, dtype=torch.float64, device=cuda)
            counter = torch.zeros(1, dtype=torch.int32, device=cuda)
            beta = torch.zeros(n_out, device=cuda)
            mean, rstd, scale, shift = (torch.zeros(n_out, device=cuda) for _ in range(4))
            for _ in range(2):          # twice: tickets and the finalize counter must be left ready for the next launch
                stats.zero_()
                fn(shape, src, pack, out, ld_in, ld_out, stats=stats,
                   bn=(counter, beta, mean, rstd, scale, shift, rows, 1e-3), splitk=splitk)
            torch.cuda.synchronize()
            outs.append(out.float().cpu().numpy())
            stats_all.append(stats.cpu().numpy())
            fin.append(torch.stack([mean, rstd, shift]).cpu().numpy())
            if splitk is not None:
                assert int(splitk[1].abs().sum()) == 0 and int(counter.item()) == 0
        sc = max(1.0, np.abs(outs[0]).max())
        assert np.abs(outs[0] - outs[1]).max() <= 1e-2 * sc                # one bf16 ulp of the largest value
        assert (outs[0] != outs[1]).mean() < 0.02                            # and only where a rounding flips
        assert np.abs(stats_all[0] - stats_all[1]).max() <= 2e-2 * max(1.0, np.abs(stats_all[0]).max())
        assert np.abs(fin[0] - fin[1]).max() <= 1e-2 * max(1.0, np.abs(fin[0]).max())


@pytest.mark.parametrize("Cin,Cout", [(3, 32), (6, 64)])
def test_small_k_persistent_kernel_equals_generic(cuda, Cin, Cout, monkeypatch):
    """The first layers at a batch that gives >= 2 tiles per SM take the persistent small-K kernel (resident weights,
    register-resident moments); output must be bit-identical to the generic kernel (same MMA order), the fused
    batch-norm moments / finalize equal to rounding, and both must match the fp32 SIMT convolution."""
    from action_conditioned_gans_b200 import kernels as Kn
    B, ld = 40, 8
    g = torch.Generator(device=cuda).manual_seed(Cin)
    shape = Kn.conv_shape(B, 64, 64, Cin, Cout, 5, 2, "SAME")
    x = torch.zeros(B, 64, 64, ld, dtype=torch.bfloat16, device=cuda)
    x[..., :Cin] = (torch.rand(B, 64, 64, Cin, device=cuda, generator=g) * 2 - 1).to(torch.bfloat16)
    w = (torch.randn(5, 5, Cin, Cout, device=cuda, generator=g) / (25 * Cin) ** 0.5).to(torch.bfloat16).float()
    pack = torch.empty(Kn.pack_size(shape, 0, ld), dtype=torch.bfloat16, device=cuda)
    Kn.pack_weights(shape, w, 0, ld, pack)
    rows = B * 32 * 32
    beta = torch.randn(Cout, device=cuda, generator=g)
    res = []
    for generic in (False, True):
        if generic:
            monkeypatch.setenv("ACG_NO_SMALLK", "1")
        else:
            monkeypatch.delenv("ACG_NO_SMALLK", raising=False)
        out = torch.zeros(rows, Cout, dtype=torch.bfloat16, device=cuda)
        stats = torch.zeros(2 * Cout, dtype=torch.float64, device=cuda)
        counter = torch.zeros(1, dtype=torch.int32, device=cuda)
        mean, rstd, scale, shift = (torch.zeros(Cout, device=cuda) for _ in range(4))
        for _ in range(2):
            stats.zero_()
            Kn.conv_fprop_tc(shape, x, pack, out, ld, Cout, stats=stats,
                             bn=(counter, beta, mean, rstd, scale, shift, rows, 1e-3))
        torch.cuda.synchronize()
        assert int(counter.item()) == 0
        res.append((out.clone(), stats.cpu().numpy(), torch.stack([mean, rstd, shift]).cpu().numpy()))
    assert torch.equal(res[0][0], res[1][0])
    assert np.abs(res[0][1] - res[1][1]).max() <= 1e-5 * np.abs(res[1][1]).max()
    assert np.abs(res[0][2] - res[1][2]).max() <= 1e-5 * max(1.0, np.abs(res[1][2]).max())
    ref = torch.empty(B, 32, 32, Cout, device=cuda)
    Kn.conv_fprop_f32(shape, x[..., :Cin].float().contiguous(), w, ref)
    assert (res[0][0].float().view_as(ref) - ref).abs().max() <= 1.2e-2 * max(1.0, float(ref.abs().max()))
    # fp32 output variant (no bf16 rounding before the moments)
    monkeypatch.delenv("ACG_NO_SMALLK", raising=False)
    out32 = torch.zeros(rows, Cout, device=cuda)
    Kn.conv_fprop_tc(shape, x, pack, out32, ld, Cout)
    assert (out32.view_as(ref) - ref).abs().max() <= 2e-3 * max(1.0, float(ref.abs().max()))
