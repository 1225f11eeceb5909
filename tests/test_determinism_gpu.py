"""Run-to-run reproducibility of the forward pass.  The batch-norm moments fused into the conv
epilogues are added in a fixed order inside a CTA and across CTAs as fixed-point integer limbs (integer atomics are
associative), and split-K tiles are added in split order -- so a rerun on the same inputs is BITWISE identical, for every
kernel family, and so is the whole generator / discriminator forward.  (The backward pass still accumulates weight gradients with fp32 atomics; see
DESIGN.md section 7.)"""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def ru(v, m):
    return (v + m - 1) // m * m


# (B, H, W, Cin, Cout, k, which, expected kernel kind or None)
CASES = [
    (8, 32, 32, 64, 128, 5, 0, 1),       # halo CONV form, two accumulators
    (6, 64, 64, 32, 64, 5, 0, 1),        # halo CONV form, 32-wide tiles
    (16, 16, 16, 128, 256, 5, 0, None),  # 8x8 output grid
    (8, 32, 32, 64, 128, 5, 1, 1),       # halo ADJ form (g/tconv2-like: 16x16x128 -> 32x32x64)
    (4, 64, 64, 32, 64, 5, 1, 1),        # halo ADJ form, 64-wide rows
    (40, 64, 64, 6, 64, 5, 0, 0),        # persistent small-K kernel
    (3, 64, 64, 6, 64, 5, 0, 0),         # generic kernel, CONV gather (too few tiles for the small-K kernel)
    (64, 4, 4, 256, 512, 5, 0, 2),       # pixel-major kernel + split-K
    (256, 8, 8, 128, 256, 5, 1, 2),      # pixel-major kernel, ADJ gather
    (32, 4, 4, 256, 512, 5, 0, 0),       # generic kernel + split-K
    (5, 8, 8, 128, 256, 5, 1, None),     # ADJ gather at an odd batch
]


@pytest.mark.parametrize("case", CASES)
def test_fused_moments_are_bitwise_reproducible(cuda, case):
    from action_conditioned_gans_b200 import kernels as Kn
    B, H, W, Cin, Cout, k, which, kind = case
    g = torch.Generator(device=cuda).manual_seed(11)
    shape = Kn.conv_shape(B, H, W, Cin, Cout, k, 2, "SAME")
    w = torch.randn(k, k, Cin, Cout, device=cuda, generator=g) / (k * Cin ** 0.5)
    if which == 0:
        ld_in, n_out, rows = (8 if Cin <= 8 else ru(Cin, 16)), Cout, B * shape.OH * shape.OW
        src = torch.zeros(B, H, W, ld_in, dtype=torch.bfloat16, device=cuda)
        src[..., :Cin] = torch.randn(B, H, W, Cin, device=cuda, generator=g).to(torch.bfloat16)
        fn = Kn.conv_fprop_tc
    else:
        ld_in, n_out, rows = ru(Cout, 64), Cin, B * H * W
        src = torch.zeros(B, shape.OH, shape.OW, ld_in, dtype=torch.bfloat16, device=cuda)
        src[..., :Cout] = torch.randn(B, shape.OH, shape.OW, Cout, device=cuda, generator=g).to(torch.bfloat16)
        fn = Kn.conv_dgrad_tc
    if kind is not None:
        assert Kn.kernel_kind(shape, which, ld_in) == kind
    ld_out = ru(n_out, 16)
    pack = torch.empty(Kn.pack_size(shape, which, ld_in), dtype=torch.bfloat16, device=cuda)
    Kn.pack_weights(shape, w, which, ld_in, pack)
    wsp = Kn.splitk_workspace(shape, which, ld_in, cuda)
    ws = Kn.stats_accumulators(n_out, cuda)
    beta = torch.randn(n_out, device=cuda, generator=g)
    counter = torch.zeros(1, dtype=torch.int32, device=cuda)
    runs = []
    for rep in range(6):
        out = torch.zeros(rows, ld_out, dtype=torch.bfloat16, device=cuda)
        stats = torch.zeros(2 * n_out, dtype=torch.float64, device=cuda)
        mean, rstd, scale, shift = (torch.zeros(n_out, device=cuda) for _ in range(4))
        fn(shape, src, pack, out, ld_in, ld_out, stats=stats, bn=(counter, beta, mean, rstd, scale, shift, rows, 1e-3),
           splitk=wsp, stats_fix=ws)
        torch.cuda.synchronize()
        assert int(counter.item()) == 0 and int(ws.abs().sum()) == 0          # left ready for the next launch
        runs.append((out, stats, torch.stack([mean, rstd, scale, shift])))
    for out, stats, fin in runs[1:]:
        assert torch.equal(out, runs[0][0]) and torch.equal(stats, runs[0][1]) and torch.equal(fin, runs[0][2])
    # the totals are those of the values as stored (fp32 partial sums per warp and tile, fp64 above)
    out, stats, fin = runs[0]
    o = out[:, :n_out].double()
    ref = torch.cat([o.sum(0), (o * o).sum(0)])
    assert float((stats - ref).abs().max()) <= 2e-6 * max(1.0, float(ref.abs().max()))
    mu = ref[:n_out] / rows
    var = (ref[n_out:] / rows - mu * mu).clamp_min(0)
    assert float((fin[0].double() - mu).abs().max()) <= 1e-5
    assert float((fin[1].double() - (var + 1e-3).rsqrt()).abs().max()) <= 1e-4 * float((var + 1e-3).rsqrt().max())
    # totals only (ticket with rows == 0: what the data-parallel path asks for), then the atomics path for comparison
    stats2 = torch.zeros(2 * n_out, dtype=torch.float64, device=cuda)
    fn(shape, src, pack, out, ld_in, ld_out, stats=stats2, bn=(counter, None, None, None, None, None, 0, 1e-3),
       splitk=wsp, stats_fix=ws)
    assert torch.equal(stats2, stats)
    stats3 = torch.zeros(2 * n_out, dtype=torch.float64, device=cuda)
    fn(shape, src, pack, out, ld_in, ld_out, stats=stats3, splitk=wsp)
    assert float((stats3 - stats).abs().max()) <= 1e-9 * max(1.0, float(stats.abs().max()))
    # accumulators that are too short are ignored (fp64 atomics), never overrun
    stats4 = torch.zeros(2 * n_out, dtype=torch.float64, device=cuda)
    short = torch.zeros(6 * n_out - 1, dtype=torch.int64, device=cuda)
    fn(shape, src, pack, out, ld_in, ld_out, stats=stats4, bn=(counter, None, None, None, None, None, 0, 1e-3),
       splitk=wsp, stats_fix=short)
    assert float((stats4 - stats).abs().max()) <= 1e-9 * max(1.0, float(stats.abs().max()))
    assert int(counter.item()) == 0 and int(short.abs().sum()) == 0
    # the same moments WITHOUT a ticket: the launch only adds its limbs (no fence / ticket / last-CTA pass) and the
    # activation pass completes, finalises and applies them -- same bits as the in-kernel finalize + plain activation pass
    for act in ("relu", "lrelu", "none"):
        stats6 = torch.zeros(2 * n_out, dtype=torch.float64, device=cuda)
        ws.zero_()
        fin6 = [torch.zeros(n_out, device=cuda) for _ in range(4)]
        fn(shape, src, pack, out, ld_in, ld_out, stats=stats6, splitk=wsp, stats_fix=ws)
        assert int(ws.abs().sum()) != 0 and float(stats6.abs().sum()) == 0.0     # limbs only, nobody converted them
        a = torch.zeros(rows, ld_out, dtype=torch.bfloat16, device=cuda)
        assert Kn.bn_finalize_act_fwd_ok(n_out, ld_out, ld_out, act)
        Kn.bn_finalize_act_fwd(out, rows, n_out, ld_out, stats6, ws, beta, rows, 1e-3, *fin6, act, a, ld_out)
        a_ref = torch.zeros(rows, ld_out, dtype=torch.bfloat16, device=cuda)
        Kn.bn_act_fwd(out, rows, n_out, ld_out, 1, fin[2], fin[3], act, a_ref, ld_out)
        assert torch.equal(torch.stack(fin6), fin) and torch.equal(a.view(torch.int16), a_ref.view(torch.int16))
    ws.zero_()
    # a diverged layer still reads as diverged: inf in the input -> non-finite totals
    bad = src.clone()
    bad.view(-1)[0] = float("inf")
    stats5 = torch.zeros(2 * n_out, dtype=torch.float64, device=cuda)
    fn(shape, bad, pack, out, ld_in, ld_out, stats=stats5, bn=(counter, None, None, None, None, None, 0, 1e-3),
       splitk=wsp, stats_fix=ws)
    assert not bool(torch.isfinite(stats5).all()) and int(ws.abs().sum()) == 0


@pytest.mark.parametrize("B", [3, 16])
@pytest.mark.parametrize("dna", [True, False])
def test_network_forward_is_bitwise_reproducible(cuda, B, dna):
    """Generator and discriminator forward (bf16 product mode, stream branches on): eight reruns, identical bits."""
    from action_conditioned_gans_b200 import engine as E
    rng = np.random.RandomState(7)
    gspec = E.g_dna_spec(6) if dna else E.g_direct_spec()
    gstore = E.ParamStore(gspec, cuda, E.xavier_init(gspec, rng))
    dstore = E.ParamStore(E.d_spec(), cuda, E.xavier_init(E.d_spec(), rng))
    grun = E.GeneratorRun(gstore, B, cuda, dna, 6)
    drun = E.DiscriminatorRun(dstore, B, cuda)
    gstore.refresh_packs()
    dstore.refresh_packs()
    img = torch.rand(B, 64, 64, 3, device=cuda) * 2 - 1
    act = torch.randn(B, 10, device=cuda)
    first = None
    for rep in range(8):
        frame, state = grun.forward(img, act)
        logits = drun.forward(img, frame, act)
        torch.cuda.synchronize()
        got = (frame.clone(), None if state is None else state.clone(), logits.clone())
        if first is None:
            first = got
            continue
        for a, b in zip(got, first):
            assert (a is None and b is None) or torch.equal(a, b)
