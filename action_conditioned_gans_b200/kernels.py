"""Thin Python shims over the C-ABI (one function per entry point of include/acg_b200.h).

Every function takes CUDA torch tensors (used only as device storage), enqueues the kernel on torch's current
stream and returns immediately.  Nothing here computes on the CPU or through torch ops.
"""
import ctypes as C

import torch

from . import _lib
from ._lib import ACT_IDS, ConvShape, TcArgs, call, dtype_id, ptr, stream


def same_pad(n_in, k, s):
    """TF SAME: out=ceil(n/s), pad_total=max((out-1)s+k-n,0), before=total//2 (odd element after)."""
    out = -(-n_in // s)
    total = max((out - 1) * s + k - n_in, 0)
    return out, total // 2


def conv_shape(B, H, W, Cin, Cout, k, stride, padding="SAME"):
    """acg_conv_shape of a forward convolution with TF padding semantics."""
    if padding == "SAME":
        OH, pt = same_pad(H, k, stride)
        OW, pl = same_pad(W, k, stride)
    elif padding == "VALID":
        OH, OW, pt, pl = (H - k) // stride + 1, (W - k) // stride + 1, 0, 0
    else:
        raise ValueError("padding must be SAME or VALID")
    return ConvShape(B, H, W, Cin, OH, OW, Cout, k, k, stride, pt, pl)


# ---- DNA ------------------------------------------------------------------------------------
def dna_fwd(logits, img, out, K):
    B, H, W, Cc = img.shape
    call("acg_dna_fwd", ptr(logits), dtype_id(logits), ptr(img), ptr(out), B, H, W, Cc, K, stream())


def dna_bwd(logits, img, dy, dlogits, K):
    """dlogits: dense [B,H,W,K*K] in the logits dtype, or bf16 [B,H,W,ru16(K*K)] (zero pad channels) for fp32 logits"""
    B, H, W, Cc = img.shape
    call("acg_dna_bwd", ptr(logits), dtype_id(logits), ptr(img), ptr(dy), ptr(dlogits), dtype_id(dlogits),
         dlogits.shape[3], B, H, W, Cc, K, stream())


def bias_grad(red, Cc, scale, dbias):
    call("acg_bias_grad", ptr(red), Cc, scale, ptr(dbias), stream())


# ---- convolutions -----------------------------------------------------------------------------
def conv_fprop_f32(shape, x, w, y):
    call("acg_conv_fprop_f32", C.byref(shape), ptr(x), ptr(w), ptr(y), stream())


def conv_dgrad_f32(shape, dy, w, dx):
    call("acg_conv_dgrad_f32", C.byref(shape), ptr(dy), ptr(w), ptr(dx), stream())


def conv_wgrad_f32(shape, x, dy, dw):
    call("acg_conv_wgrad_f32", C.byref(shape), ptr(x), ptr(dy), ptr(dw), stream())


def tc_supported(shape, which):
    return bool(_lib.load().acg_conv_tc_supported(C.byref(shape), which))


def kernel_kind(shape, which, ld_in, n_limit=0):
    """which kernel a conv_fprop_tc (which=0) / conv_dgrad_tc (which=1) launch of this shape takes: 2 pixel-major (small
    feature maps), 1 halo-tile, 0 generic"""
    return int(_lib.load().acg_conv_kernel_kind(C.byref(shape), which, ld_in, int(n_limit)))


def pair_ok(shape, ld_in):
    """True when conv_fprop_tc(..., pair_x=True) covers the shape (first layers: 8-channel frames read as pixel pairs)"""
    return bool(_lib.load().acg_conv_pair_ok(C.byref(shape), ld_in))


def splitk_workspace(shape, which, ld_in, device):
    """(workspace, tickets) tensors a conv_fprop_tc (which=0) / conv_dgrad_tc (which=1) launch of this shape wants for
    split-K, or None when it never splits (acg_conv_splitk_plan)."""
    splits, nbytes, ntick = C.c_int(), C.c_longlong(), C.c_int()
    call("acg_conv_splitk_plan", C.byref(shape), which, ld_in, C.byref(splits), C.byref(nbytes), C.byref(ntick))
    if splits.value <= 1:
        return None
    return (torch.empty(nbytes.value // 4, dtype=torch.float32, device=device),
            torch.zeros(ntick.value, dtype=torch.int32, device=device))


def _tc_args(ld_in, ld_out, bias, out, out_act, stats=None, bn=None, red=None, splitk=None, n_limit=0, stats_fix=None, peer=None, pair_x=False):
    """bn = (counter, beta, mean, rstd, scale, shift, rows, eps): finalise the moments in-kernel (single GPU); rows == 0
      with only the counter set: the last CTA completes the totals and the caller finalises (data parallel)
    red = (red_buffer, z, ldz, C, act, mean, rstd, shift): fused batch-norm backward reduction of the consumer layer
    stats_fix: int64 tensor of >= 6*C zeros: reproducible moments (integer limb accumulators).  With bn's ticket the last
      CTA converts them into `stats` and leaves them zeroed; without bn the launch only adds its limbs and
      bn_finalize_act_fwd completes them (the caller zeroes the accumulators before the next launch)"""
    t = TcArgs(ld_in, ld_out, ptr(bias), dtype_id(out), ACT_IDS[out_act], ptr(stats))
    t.n_limit = int(n_limit)
    if stats_fix is not None:
        t.stats_fix, t.stats_fix_len = ptr(stats_fix), stats_fix.numel()
    t.pair_x = int(bool(pair_x))
    if peer is not None:          # (Mailbox, slot): the launch's last CTA sums the moments over the ranks itself
        t._peer = peer[0].exchange_struct(peer[1])     # kept alive by the args object for the duration of the call
        t.peer = C.addressof(t._peer)
    if splitk is not None:
        ws, tickets = splitk
        t.splitk_ws, t.splitk_ws_bytes = ptr(ws), ws.numel() * 4
        t.splitk_tickets, t.splitk_n_tickets = ptr(tickets), tickets.numel()
    if red is not None:
        buf, z, ldz, Cc, act, mean, rstd, shift = red
        t.stats = ptr(buf)
        t.red_z, t.red_ldz, t.red_C, t.red_act = ptr(z), ldz, Cc, ACT_IDS[act]
        t.red_mean, t.red_rstd, t.red_shift = ptr(mean), ptr(rstd), ptr(shift)
    if bn is not None:
        counter, beta, mean, rstd, scale, shift, rows, eps = bn
        t.bn_counter, t.bn_beta = ptr(counter), ptr(beta)
        t.bn_mean, t.bn_rstd, t.bn_scale, t.bn_shift = ptr(mean), ptr(rstd), ptr(scale), ptr(shift)
        t.bn_rows, t.bn_eps = rows, eps
    return t


def conv_fprop_tc(shape, x, w_pack, y, ld_in, ld_out, bias=None, out_act=None, stats=None, bn=None, red=None,
                  splitk=None, n_limit=0, stats_fix=None, peer=None, pair_x=False):
    t = _tc_args(ld_in, ld_out, bias, y, out_act, stats, bn, red, splitk, n_limit, stats_fix, peer, pair_x)
    call("acg_conv_fprop_tc", C.byref(shape), ptr(x), ptr(w_pack), ptr(y), C.byref(t), stream())


def conv_dgrad_tc(shape, dy, w_pack, dx, ld_in, ld_out, bias=None, out_act=None, stats=None, bn=None, red=None,
                  splitk=None, n_limit=0, stats_fix=None, peer=None, pair_x=False):
    t = _tc_args(ld_in, ld_out, bias, dx, out_act, stats, bn, red, splitk, n_limit, stats_fix, peer, pair_x)
    call("acg_conv_dgrad_tc", C.byref(shape), ptr(dy), ptr(w_pack), ptr(dx), C.byref(t), stream())


def conv_wgrad_tc(shape, x, dy, dw, ld_x, ld_dy):
    t = TcArgs(ld_x, ld_dy, None, _lib.F32, 0)
    call("acg_conv_wgrad_tc", C.byref(shape), ptr(x), ptr(dy), ptr(dw), C.byref(t), stream())


def stats_accumulators(channels, device):
    """zeroed integer limb accumulators for reproducible fused batch-norm moments of a layer with `channels` outputs"""
    return torch.zeros(6 * channels, dtype=torch.int64, device=device)


def pack_size(shape, which, ld_k):
    n = int(_lib.load().acg_pack_size(C.byref(shape), which, ld_k))
    if n < 0:
        raise RuntimeError("acg_pack_size: invalid arguments")
    return n


def make_pack_jobs(entries, device):
    """entries: [(shape, w, which, ld_k, pack)] -> (device job table, njobs, device tile table, ntiles)"""
    jobs = (_lib.PackJob * len(entries))()
    first = 0
    for i, (shape, w, which, ld_k, pack) in enumerate(entries):
        n_rows = (shape.Cout + 15) // 16 * 16 if which in (0, 2) else (shape.Cin + 15) // 16 * 16
        jobs[i] = _lib.PackJob(ptr(w), ptr(pack), first, which, ld_k, shape.KH, shape.KW, shape.Cin, shape.Cout,
                               shape.stride, shape.pad_t, shape.pad_l, n_rows)
        first += pack_size(shape, which, ld_k)
    lib = _lib.load()
    ntiles = int(lib.acg_pack_plan(jobs, len(entries), None, 0))
    if ntiles <= 0:
        raise RuntimeError("acg_pack_plan: invalid pack jobs")
    tiles = (C.c_int * (4 * ntiles))()
    if int(lib.acg_pack_plan(jobs, len(entries), tiles, ntiles)) != ntiles:
        raise RuntimeError("acg_pack_plan: tile count changed")
    table = torch.frombuffer(bytearray(bytes(jobs)), dtype=torch.uint8).to(device)
    tile_table = torch.frombuffer(bytearray(bytes(tiles)), dtype=torch.int32).to(device)
    return table, len(entries), tile_table, ntiles


def pack_weights_batched(table, njobs, tile_table, ntiles):
    call("acg_pack_weights_batched", ptr(table), njobs, ptr(tile_table), ntiles, stream())


def pack_weights(shape, w, which, ld_k, pack):
    call("acg_pack_weights", C.byref(shape), ptr(w), which, ld_k, ptr(pack), stream())


# ---- batch-norm / activation / concat -------------------------------------------------------------
def bn_stats(z, rows, Cc, ld, groups, stats):
    call("acg_bn_stats", ptr(z), dtype_id(z), rows, Cc, ld, groups, ptr(stats), stream())


def bn_finalize(stats, beta, rows_per_group, Cc, groups, mean, rstd, scale, shift, eps=1e-3):
    call("acg_bn_finalize", ptr(stats), ptr(beta), rows_per_group, Cc, groups, eps, ptr(mean), ptr(rstd),
         ptr(scale), ptr(shift), stream())


def bn_act_fwd(z, rows, Cc, ld_in, groups, scale, shift, act, out, ld_out):
    call("acg_bn_act_fwd", ptr(z), dtype_id(z), rows, Cc, ld_in, groups, ptr(scale), ptr(shift), ACT_IDS[act],
         ptr(out), dtype_id(out), ld_out, stream())


def bn_act_fwd_cat(z, rows, Cc, ld_in, scale, shift, act, out, ld_out, actions, hw, act_off):
    """bn_act_fwd into a concat buffer + the tiled action vector behind the features, one launch"""
    call("acg_bn_act_fwd_cat", ptr(z), dtype_id(z), rows, Cc, ld_in, ptr(scale), ptr(shift), ACT_IDS[act], ptr(out),
         dtype_id(out), ld_out, ptr(actions), actions.shape[1], hw, act_off, stream())


def bn_finalize_act_fwd_ok(Cc, ld_in, ld_out, act):
    return bool(_lib.load().acg_bn_finalize_act_fwd_ok(Cc, ld_in, ld_out, ACT_IDS[act]))


def bn_finalize_act_fwd(z, rows, Cc, ld_in, stats, stats_fix, beta, norm_rows, eps, mean, rstd, scale, shift, act, out,
                        ld_out, cat=None, hw=1):
    """batch-norm finalize + activation from the RAW moments of a conv launch without a ticket; cat = (actions, offset)"""
    acts, off = cat if cat is not None else (None, 0)
    call("acg_bn_finalize_act_fwd", ptr(z), rows, Cc, ld_in, ptr(stats), ptr(stats_fix), ptr(beta), norm_rows, eps,
         ptr(mean), ptr(rstd), ptr(scale), ptr(shift), ACT_IDS[act], ptr(out), ld_out, ptr(acts),
         acts.shape[1] if acts is not None else 0, hw, off, stream())


def bn_act_bwd_reduce(dA, dA2, ld_d, z, ld_z, rows, Cc, groups, mean, rstd, shift, act, red):
    call("acg_bn_act_bwd_reduce", ptr(dA), ptr(dA2), dtype_id(dA), ld_d, ptr(z), dtype_id(z) if z is not None else 0,
         ld_z, rows, Cc,
         groups, ptr(mean), ptr(rstd), ptr(shift), ACT_IDS[act], ptr(red), stream())


def bn_act_bwd_reduce_sync(dA, dA2, ld_d, z, ld_z, rows, Cc, mean, rstd, shift, act, red, counter, mailbox, slot):
    """bn_act_bwd_reduce + the sum of red over the data-parallel ranks, in one launch where the fast path applies"""
    px = mailbox.exchange_struct(slot)
    call("acg_bn_act_bwd_reduce_sync", ptr(dA), ptr(dA2), dtype_id(dA), ld_d, ptr(z),
         dtype_id(z) if z is not None else 0, ld_z, rows, Cc, ptr(mean), ptr(rstd), ptr(shift), ACT_IDS[act], ptr(red),
         ptr(counter), C.addressof(px), stream())


def bn_act_bwd_apply(dA, dA2, ld_d, z, ld_z, rows, Cc, groups, mean, rstd, shift, act, has_bn, red, dz, dbeta,
                     norm_rows=0, dbeta_scale=1.0, ld_dz=None):
    call("acg_bn_act_bwd_apply", ptr(dA), ptr(dA2), dtype_id(dA), ld_d, ptr(z), dtype_id(z) if z is not None else 0,
         ld_z, rows, Cc, groups, ptr(mean), ptr(rstd), ptr(shift), ACT_IDS[act], int(has_bn), ptr(red), ptr(dz),
         dtype_id(dz), Cc if ld_dz is None else ld_dz, ptr(dbeta), norm_rows, dbeta_scale, stream())


def copy_channels(src, ld_src, off_src, dst, ld_dst, off_dst, rows, n):
    call("acg_copy_channels", ptr(src), dtype_id(src), ld_src, off_src, ptr(dst), dtype_id(dst), ld_dst, off_dst,
         rows, n, stream())


def pack_frames(a, b, out, rows):
    """out[r] (bf16, 16 channels) = [a[r] | b[r] | 0]; b may be None"""
    call("acg_pack_frames", ptr(a), ptr(b), 3, ptr(out), out.shape[-1], rows, stream())


def tile_actions(actions, B, hw, dst, ld_dst, off):
    A = actions.shape[1]
    call("acg_tile_actions", ptr(actions), B, hw, A, ptr(dst), dtype_id(dst), ld_dst, off, stream())


# ---- device-side feeder / rollout glue ---------------------------------------------------------------
def gather_frames(frames, actions, sample, t0, img, nxt, act, next_state, pair_stride=1, geometry=None):
    """frames [N,T,H,W,3] uint8 or fp32, actions [N,T,A] or None; sample / t0 int32 [B] (all on the device).
    geometry = (N, T) overrides the leading dimensions (host-staged [2,B,...] batches: N=1, T=2B, pair_stride=B)."""
    N, T = geometry if geometry is not None else (frames.shape[0], frames.shape[1])
    fe = img[0].numel()
    B = sample.numel()
    if frames.dtype == torch.uint8:
        fdt = _lib.U8
    elif frames.dtype == torch.float32:
        fdt = _lib.F32
    else:
        raise RuntimeError("frames must be uint8 or float32")
    if sample.dtype != torch.int32 or t0.dtype != torch.int32:
        raise RuntimeError("sample / t0 must be int32")
    A = actions.shape[-1] if actions is not None else 0
    S = next_state.shape[1] if next_state is not None else 0
    call("acg_gather_frames", ptr(frames), fdt, ptr(actions), ptr(sample), ptr(t0), N, T, pair_stride, fe, max(A, 1),
         S if actions is not None else 0, B, ptr(img), ptr(nxt), ptr(act), ptr(next_state), stream())


def rollout_actions(acts, j, state, out, S=5):
    B, T, A = acts.shape
    call("acg_rollout_actions", ptr(acts), T, j, ptr(state), ptr(out), B, A, S, stream())


# ---- losses -------------------------------------------------------------------------------------
def frame_losses(g, n, sums, dg=None, w_l1=0.0, w_gdl=0.0, dadv=None, ld_adv=0, adv_off=0):
    B, H, W, _ = g.shape
    call("acg_frame_losses", ptr(g), ptr(n), B, H, W, ptr(sums), ptr(dg), w_l1, w_gdl, ptr(dadv), ld_adv, adv_off,
         stream())


def dlogit_loss(x, n, kind, label_or_sign, grad_scale, loss_out, dlogits=None):
    if kind not in ("bce", "wass"):
        raise ValueError("unexpected loss argument")
    call("acg_dlogit_loss", ptr(x), n, _lib.LOSS_BCE if kind == "bce" else _lib.LOSS_WASS, label_or_sign,
         grad_scale, ptr(loss_out), ptr(dlogits), stream())


def state_loss(s, t, n, inv_batch, grad_scale, loss_out, dstate=None, sumsq_out=None, sumsq_in=None):
    """sumsq_out / sumsq_in (fp64 scalars): the two phases of the batch-sharded form, see include/acg_b200.h"""
    call("acg_state_loss", ptr(s), ptr(t), n, inv_batch, grad_scale, ptr(loss_out), ptr(dstate), ptr(sumsq_out),
         ptr(sumsq_in), stream())


# ---- optimizers -----------------------------------------------------------------------------------
NO_CLIP = (1.0, -1.0)  # lo > hi disables the clip


def adam_step(p, g, m, v, lr_t, b1=0.9, b2=0.999, eps=1e-8, clip=NO_CLIP, grad_scale=1.0, lr_dev=None):
    call("acg_adam_step", ptr(p), ptr(g), ptr(m), ptr(v), p.numel(), lr_t, b1, b2, eps, clip[0], clip[1],
         grad_scale, ptr(lr_dev), stream())


def rmsprop_step(p, g, ms, lr, decay=0.9, eps=1e-10, clip=NO_CLIP, grad_scale=1.0, lr_dev=None):
    call("acg_rmsprop_step", ptr(p), ptr(g), ptr(ms), p.numel(), lr, decay, eps, clip[0], clip[1], grad_scale,
         ptr(lr_dev), stream())
