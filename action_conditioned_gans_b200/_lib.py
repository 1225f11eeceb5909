"""ctypes binding of libacg_b200.so (include/acg_b200.h).

This is the only place the shared library is opened.  There is NO fallback of any kind: if the library is
missing or a call fails, a RuntimeError is raised (errors are Python exceptions).
"""
import ctypes as C
import os

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libacg_b200.so")

F32, BF16, U8 = 0, 1, 2
ACT_NONE, ACT_RELU, ACT_LRELU, ACT_TANH = 0, 1, 2, 3
LOSS_BCE, LOSS_WASS = 0, 1
ACT_IDS = {None: ACT_NONE, "none": ACT_NONE, "relu": ACT_RELU, "lrelu": ACT_LRELU, "tanh": ACT_TANH}


class ConvShape(C.Structure):
    """struct acg_conv_shape: always the FORWARD convolution x[B,H,W,Cin] -> y[B,OH,OW,Cout]."""
    _fields_ = [(n, C.c_int) for n in
                ("B", "H", "W", "Cin", "OH", "OW", "Cout", "KH", "KW", "stride", "pad_t", "pad_l")]


class PeerExchange(C.Structure):
    """struct acg_peer_exchange"""
    _fields_ = [("mailboxes", C.c_void_p), ("epoch", C.c_void_p), ("slot_off", C.c_longlong), ("rank", C.c_int),
                ("world", C.c_int), ("cap", C.c_int), ("timeout_s", C.c_float)]


class TcArgs(C.Structure):
    """struct acg_tc_args"""
    _fields_ = [("ld_in", C.c_int), ("ld_out", C.c_int), ("bias", C.c_void_p), ("out_dtype", C.c_int),
                ("out_act", C.c_int), ("stats", C.c_void_p), ("bn_counter", C.c_void_p), ("bn_beta", C.c_void_p),
                ("bn_mean", C.c_void_p), ("bn_rstd", C.c_void_p), ("bn_scale", C.c_void_p), ("bn_shift", C.c_void_p),
                ("bn_rows", C.c_longlong), ("bn_eps", C.c_float),
                ("red_z", C.c_void_p), ("red_ldz", C.c_int), ("red_C", C.c_int), ("red_act", C.c_int),
                ("red_mean", C.c_void_p), ("red_rstd", C.c_void_p), ("red_shift", C.c_void_p),
                ("splitk_ws", C.c_void_p), ("splitk_ws_bytes", C.c_longlong), ("splitk_tickets", C.c_void_p),
                ("splitk_n_tickets", C.c_int), ("n_limit", C.c_int), ("stats_fix", C.c_void_p),
                ("stats_fix_len", C.c_longlong), ("peer", C.c_void_p), ("pair_x", C.c_int)]


class PackJob(C.Structure):
    """struct acg_pack_job"""
    _fields_ = [("w", C.c_void_p), ("pack", C.c_void_p), ("first", C.c_longlong), ("which", C.c_int),
                ("ld_k", C.c_int), ("KH", C.c_int), ("KW", C.c_int), ("Cin", C.c_int), ("Cout", C.c_int),
                ("stride", C.c_int), ("pad_t", C.c_int), ("pad_l", C.c_int), ("N", C.c_int)]


_P, _I, _L, _F = C.c_void_p, C.c_int, C.c_longlong, C.c_float
_SP = C.POINTER(ConvShape)
_FP = C.POINTER(TcArgs)

# name -> argtypes; must list every function include/acg_b200.h declares (tests/test_abi.py checks this)
SIGNATURES = {
    "acg_dna_fwd": [_P, _I, _P, _P, _I, _I, _I, _I, _I, _P],
    "acg_dna_bwd": [_P, _I, _P, _P, _P, _I, _I, _I, _I, _I, _I, _I, _P],
    "acg_bias_grad": [_P, _I, _F, _P, _P],
    "acg_conv_fprop_f32": [_SP, _P, _P, _P, _P],
    "acg_conv_dgrad_f32": [_SP, _P, _P, _P, _P],
    "acg_conv_wgrad_f32": [_SP, _P, _P, _P, _P],
    "acg_conv_fprop_tc": [_SP, _P, _P, _P, _FP, _P],
    "acg_conv_dgrad_tc": [_SP, _P, _P, _P, _FP, _P],
    "acg_conv_wgrad_tc": [_SP, _P, _P, _P, _FP, _P],
    "acg_pack_weights": [_SP, _P, _I, _I, _P, _P],
    "acg_pack_weights_batched": [_P, _I, _P, _I, _P],
    "acg_conv_tc_supported": [_SP, _I],
    "acg_conv_pair_ok": [_SP, _I],
    "acg_conv_kernel_kind": [_SP, _I, _I, _I],
    "acg_conv_splitk_plan": [_SP, _I, _I, C.POINTER(C.c_int), C.POINTER(C.c_longlong), C.POINTER(C.c_int)],
    "acg_bn_stats": [_P, _I, _L, _I, _I, _I, _P, _P],
    "acg_bn_finalize": [_P, _P, _L, _I, _I, _F, _P, _P, _P, _P, _P],
    "acg_bn_act_fwd": [_P, _I, _L, _I, _I, _I, _P, _P, _I, _P, _I, _I, _P],
    "acg_bn_act_fwd_cat": [_P, _I, _L, _I, _I, _P, _P, _I, _P, _I, _I, _P, _I, _I, _I, _P],
    "acg_bn_finalize_act_fwd_ok": [_I, _I, _I, _I],
    "acg_bn_finalize_act_fwd": [_P, _L, _I, _I, _P, _P, _P, _L, _F, _P, _P, _P, _P, _I, _P, _I, _P, _I, _I, _I, _P],
    "acg_bn_act_bwd_reduce": [_P, _P, _I, _I, _P, _I, _I, _L, _I, _I, _P, _P, _P, _I, _P, _P],
    "acg_bn_act_bwd_reduce_sync": [_P, _P, _I, _I, _P, _I, _I, _L, _I, _P, _P, _P, _I, _P, _P, _P, _P],
    "acg_bn_act_bwd_apply": [_P, _P, _I, _I, _P, _I, _I, _L, _I, _I, _P, _P, _P, _I, _I, _P, _P, _I, _I, _P, _L, _F,
                             _P],
    "acg_copy_channels": [_P, _I, _I, _I, _P, _I, _I, _I, _L, _I, _P],
    "acg_pack_frames": [_P, _P, _I, _P, _I, _L, _P],
    "acg_tile_actions": [_P, _I, _I, _I, _P, _I, _I, _I, _P],
    "acg_frame_losses": [_P, _P, _I, _I, _I, _P, _P, _F, _F, _P, _I, _I, _P],
    "acg_dlogit_loss": [_P, _I, _I, _F, _F, _P, _P, _P],
    "acg_state_loss": [_P, _P, _I, _F, _F, _P, _P, _P, _P, _P],
    "acg_adam_step": [_P, _P, _P, _P, _L, _F, _F, _F, _F, _F, _F, _F, _P, _P],
    "acg_rmsprop_step": [_P, _P, _P, _L, _F, _F, _F, _F, _F, _F, _P, _P],
    "acg_gather_frames": [_P, _I, _P, _P, _P, _I, _I, _I, _I, _I, _I, _I, _P, _P, _P, _P, _P],
    "acg_rollout_actions": [_P, _I, _I, _P, _P, _I, _I, _I, _P],
    "acg_peer_alloc": [_L, C.POINTER(C.c_void_p)],
    "acg_peer_free": [_P],
    "acg_peer_export": [_P, _P],
    "acg_peer_open": [_P, C.POINTER(C.c_void_p)],
    "acg_peer_close": [_P],
    "acg_peer_allreduce_f64": [_P, _I, _I, _L, _I, _I, C.POINTER(C.c_void_p), _P, _F, _I, _P, _L, _F, _P, _P, _P, _P,
                               _P],
}
# calls that return a plain value instead of a status
PLAIN = {"acg_version": ([], C.c_int), "acg_last_error": ([], C.c_char_p),
         "acg_launch_count": ([], C.c_longlong), "acg_pack_size": ([_SP, _I, _I], C.c_longlong),
         "acg_peer_slot_bytes": ([_I, _I], C.c_longlong),
         "acg_pack_plan": ([_P, _I, _P, _L], C.c_longlong)}

# include/acg_b200_probe.h: only in libacg_b200_probe.so (profiling scripts, hardware-behaviour probes)
PROBE_SIGNATURES = {
    "acg_debug_phase_times": [_P],
    "acg_debug_umma_shift": [_P, _I, _P, _I, _I, _I, _I, _P, _P],
}

_lib = None
_probe = False


def use_probe_library():
    """Profiling scripts only: load libacg_b200_probe.so (same kernels compiled with -DACG_PROBES, which adds the
    ACG_DBG_SKIP pipeline switches and the acg_debug_* entry points).  Must be called before the first load()."""
    global LIB_PATH, _probe
    if _lib is not None and not _probe:
        raise RuntimeError("use_probe_library() must be called before the library is first used")
    LIB_PATH = os.path.join(HERE, "libacg_b200_probe.so")
    _probe = True


def load():
    """Open the shared library (once).  Raises RuntimeError if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            "libacg_b200.so is missing (%s): run `python -c 'import __graft_entry__ as g; g.build()'` "
            "-- there is no CPU or PyTorch fallback for the acg_b200 kernels" % LIB_PATH)
    lib = C.CDLL(LIB_PATH)
    for name, argtypes in SIGNATURES.items():
        fn = getattr(lib, name)
        fn.argtypes = argtypes
        fn.restype = C.c_int
    for name, (argtypes, restype) in PLAIN.items():
        fn = getattr(lib, name)
        fn.argtypes = argtypes
        fn.restype = restype
    if _probe:
        for name, argtypes in PROBE_SIGNATURES.items():
            fn = getattr(lib, name)
            fn.argtypes = argtypes
            fn.restype = C.c_int
    _lib = lib
    return lib


def last_error():
    return load().acg_last_error().decode("utf-8", "replace")


def launch_count():
    return int(load().acg_launch_count())


def call(name, *args):
    """Invoke a status-returning entry point; negative status -> RuntimeError with the library's message."""
    rc = getattr(load(), name)(*args)
    if rc != 0:
        msg = last_error()
        if "unexpected loss argument" in msg or "unexpected opt argument" in msg:
            raise ValueError(msg)
        raise RuntimeError("%s failed (%d): %s" % (name, rc, msg))


def ptr(t):
    """Device pointer of a torch tensor (None -> NULL).  The tensor must be CUDA and contiguous."""
    if t is None:
        return None
    if not t.is_cuda:
        raise RuntimeError("acg_b200 kernels need CUDA tensors (got %s); there is no CPU path" % t.device)
    if not t.is_contiguous():
        raise RuntimeError("acg_b200 kernels need contiguous tensors")
    return t.data_ptr()


def stream():
    return torch.cuda.current_stream().cuda_stream


def dtype_id(t):
    if t.dtype == torch.float32:
        return F32
    if t.dtype == torch.bfloat16:
        return BF16
    raise RuntimeError("unsupported dtype %s" % t.dtype)
