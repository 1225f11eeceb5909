"""`ops.py` surface of the reference (ops.py:19-50,100-120) on the acg_b200 kernels.

Every function takes CUDA float32 torch tensors (NHWC) and returns a 0-d CUDA tensor (the losses) or a tensor
(lrelu).  They are differentiable `acg::` custom ops (torch_ops.py) whose backward is the same fused kernel; the
training step (trainer.py) calls the fused value+gradient kernels directly.  Unknown `arg_loss` raises
ValueError('unexpected loss argument') exactly like the reference.
"""
import math

import torch

from . import torch_ops as T


def _check(t):
    if not (isinstance(t, torch.Tensor) and t.is_cuda):
        raise RuntimeError("acg_b200 ops need CUDA tensors: there is no CPU fallback")
    return t.contiguous().float()


def lrelu(x, leak=0.2, name="lrelu"):
    """ops.py:22-26: f1*x + f2*abs(x) with f1 = (1+leak)/2, f2 = (1-leak)/2.  Only leak=0.2 (the only value the
    reference uses) is built into the kernel."""
    if abs(leak - 0.2) > 1e-12:
        raise RuntimeError("lrelu: only leak=0.2 is compiled in")
    x = _check(x)
    return T.bias_act(x, torch.zeros(x.shape[-1], device=x.device), "lrelu")


def build_psnr(true, pred):
    """ops.py:19-20: 10*log10(1 / mean((true-pred)^2))."""
    s = T.frame_losses(_check(pred), _check(true))
    return (10.0 * torch.log(true.numel() / s[1]) / math.log(10.0)).float()


def build_gdl(g_out, next_frames, alpha=1):
    """ops.py:100-120 with the reference's call order build_gdl(next_frame_ph, g_next_frame) in mind: the value is
    symmetric in its two arguments; the gradient flows to the first one.  Only alpha=1 (the default, the only value
    used) is compiled in."""
    if alpha != 1:
        raise RuntimeError("build_gdl: only alpha=1 is compiled in")
    return T.frame_losses(_check(g_out), _check(next_frames))[2].float()


def _logit_loss(x, kind, label_or_sign):
    return T.dlogit_loss(_check(x), kind, float(label_or_sign))


def build_g_adv_loss(d_out_gen, arg_loss):
    """ops.py:28-35"""
    if arg_loss == "bce":
        return _logit_loss(d_out_gen, "bce", 1.0)
    elif arg_loss == "wass":
        return _logit_loss(d_out_gen, "wass", 1.0)
    raise ValueError("unexpected loss argument")


def build_d_loss(d_out_direct, d_out_gen, arg_loss):
    """ops.py:37-50 (one-sided label smoothing 0.9 on the real pair; wass: mean(D(real)) - mean(D(gen)))."""
    if arg_loss == "bce":
        return _logit_loss(d_out_direct, "bce", 0.9) + _logit_loss(d_out_gen, "bce", 0.0)
    elif arg_loss == "wass":
        return _logit_loss(d_out_direct, "wass", 1.0) + _logit_loss(d_out_gen, "wass", -1.0)
    raise ValueError("unexpected loss argument")
