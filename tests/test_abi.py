"""The C-ABI shared library: builds with nvcc (no GPU needed), loads, exports every symbol include/acg_b200.h
declares, the ctypes table covers every one of them, and argument validation works without touching the device."""
import ctypes as C
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "acg_b200.h")


def declared_functions():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(acg_[a-z0-9_]+)\s*\(", src)))


def test_header_lists_the_expected_families():
    names = declared_functions()
    for n in ("acg_dna_fwd", "acg_dna_bwd", "acg_conv_fprop_tc", "acg_conv_dgrad_tc", "acg_conv_wgrad_tc",
              "acg_conv_fprop_f32", "acg_bn_stats", "acg_frame_losses", "acg_dlogit_loss", "acg_adam_step",
              "acg_rmsprop_step", "acg_pack_weights"):
        assert n in names


def test_library_exports_every_declared_symbol(built_lib):
    from action_conditioned_gans_b200 import _lib
    out = subprocess.run(["nm", "-D", "--defined-only", _lib.LIB_PATH], capture_output=True, text=True, check=True).stdout
    exported = set(re.findall(r" T (acg_[a-z0-9_]+)", out))
    declared = set(declared_functions())
    assert declared <= exported, "declared but not exported: %s" % sorted(declared - exported)
    assert exported <= declared, "exported but not declared in the header: %s" % sorted(exported - declared)
    bound = set(_lib.SIGNATURES) | set(_lib.PLAIN)
    assert bound == declared, "ctypes table mismatch: %s" % sorted(bound ^ declared)
    assert built_lib.acg_version() >= 100


def test_sass_is_blackwell_native(built_lib):
    """tcgen05.mma -> UTCHMMA, tcgen05.ld -> LDTM, bulk copies (TMA engine) -> UBLKCP, cp.async -> LDGSTS."""
    from action_conditioned_gans_b200 import _lib
    sass = subprocess.run(["cuobjdump", "-sass", _lib.LIB_PATH], capture_output=True, text=True).stdout
    if not sass:
        pytest.skip("cuobjdump unavailable")
    for mnemonic in ("UTCHMMA", "LDTM", "UBLKCP", "LDGSTS"):
        assert mnemonic in sass, mnemonic
    assert "HMMA.16816" not in sass          # no legacy mma.sync path


def test_argument_validation_without_a_device(built_lib):
    """Invalid arguments are rejected before any CUDA call (status < 0 and a message); nothing computes on the CPU."""
    from action_conditioned_gans_b200 import _lib
    lib = built_lib
    assert lib.acg_dna_fwd(None, 0, None, None, 1, 8, 8, 3, 5, None) == -1
    assert b"null" in lib.acg_last_error()
    buf = C.create_string_buffer(4096)
    p = C.cast(C.addressof(buf) + (-C.addressof(buf)) % 16, C.c_void_p)
    assert lib.acg_dna_fwd(p, 0, p, p, 1, 8, 8, 3, 4, None) == -2          # K must be 5 or 6
    assert lib.acg_dna_fwd(p, 0, p, p, 1, 8, 8, 4, 5, None) == -2          # C must be 3
    assert lib.acg_dlogit_loss(p, 4, 7, 1.0, 1.0, p, None, None) == -1     # unknown loss kind
    assert b"unexpected loss argument" in lib.acg_last_error()
    shape = _lib.ConvShape(1, 8, 8, 4, 4, 4, 4, 5, 5, 3, 1, 1)             # stride 3
    assert lib.acg_conv_fprop_f32(C.byref(shape), p, p, p, None) == -2
    assert lib.acg_adam_step(None, p, p, p, 4, 1e-3, 0.9, 0.999, 1e-8, 1.0, -1.0, 1.0, None, None) == -1
    with pytest.raises(RuntimeError, match="acg_dna_fwd failed"):
        _lib.call("acg_dna_fwd", None, 0, None, None, 1, 8, 8, 3, 5, None)


def test_bn_finalize_act_fwd_host_checks(built_lib):
    """acg_bn_finalize_act_fwd (batch-norm finalize + activation from the raw moments of a conv launch): the shape gate
    is a host function and the argument checks reject before any CUDA call."""
    from action_conditioned_gans_b200 import _lib
    from action_conditioned_gans_b200 import kernels as K
    lib = built_lib
    relu, lrelu, none_, tanh = (_lib.ACT_IDS[a] for a in ("relu", "lrelu", "none", "tanh"))
    assert lib.acg_bn_finalize_act_fwd_ok(128, 128, 128, relu) == 1
    assert lib.acg_bn_finalize_act_fwd_ok(256, 256, 272, lrelu) == 1         # into a concat buffer
    assert lib.acg_bn_finalize_act_fwd_ok(16, 16, 16, none_) == 1
    assert lib.acg_bn_finalize_act_fwd_ok(1, 16, 16, relu) == 0              # d/conv6: C not a multiple of 8
    assert lib.acg_bn_finalize_act_fwd_ok(128, 128, 128, tanh) == 0
    assert lib.acg_bn_finalize_act_fwd_ok(128, 120, 128, relu) == 0          # row stride shorter than C
    assert K.bn_finalize_act_fwd_ok(64, 64, 64, "lrelu") and not K.bn_finalize_act_fwd_ok(12, 16, 16, "relu")
    buf = C.create_string_buffer(8192)
    p = C.cast(C.addressof(buf) + (-C.addressof(buf)) % 16, C.c_void_p)
    args = [p, 64, 16, 16, p, None, None, 64, 1e-3, p, p, p, p, relu, p, 16, None, 0, 1, 0, None]
    bad = list(args); bad[4] = None                                          # no moments
    assert lib.acg_bn_finalize_act_fwd(*bad) == -1 and b"null" in lib.acg_last_error()
    bad = list(args); bad[2] = 12                                            # C = 12
    assert lib.acg_bn_finalize_act_fwd(*bad) == -2
    bad = list(args); bad[7] = 0                                             # no rows to normalise over
    assert lib.acg_bn_finalize_act_fwd(*bad) == -1
    bad = list(args); bad[16], bad[17], bad[18], bad[19] = p, 10, 7, 16      # concat: 64 rows are not whole images of 7
    assert lib.acg_bn_finalize_act_fwd(*bad) == -1 and b"concat" in lib.acg_last_error()


def test_pack_size_is_host_only(built_lib):
    from action_conditioned_gans_b200 import kernels as K
    s = K.conv_shape(4, 16, 16, 64, 128, 5, 2, "SAME")
    assert K.pack_size(s, 0, 64) == 128 * 25 * 64
    # dgrad pack: 4 parity classes with 3x3 + 3x2 + 2x3 + 2x2 = 25 taps in total
    assert K.pack_size(s, 1, 128) == 64 * 25 * 128
    assert (s.OH, s.OW, s.pad_t, s.pad_l) == (8, 8, 1, 1)


def test_host_side_planners_need_no_device(built_lib):
    """acg_pack_plan and acg_conv_splitk_plan are host functions (tile table of the batched weight pack, split-K
    decision); they must work on the CPU build box and give the documented answers for the model's layers."""
    import ctypes as C
    from action_conditioned_gans_b200 import _lib
    from action_conditioned_gans_b200 import kernels as K
    lib = built_lib
    # d/conv5 at B=256: [B,4,4,256] -> [B,2,2,512]: 8 x 4 = 32 output tiles, 100 K blocks -> 4 splits of 25
    s = K.conv_shape(256, 4, 4, 256, 512, 5, 2)
    splits, nbytes, ntick = C.c_int(), C.c_longlong(), C.c_int()
    assert lib.acg_conv_splitk_plan(C.byref(s), 0, 256, C.byref(splits), C.byref(nbytes), C.byref(ntick)) == 0
    assert splits.value == 4 and ntick.value == 32 and nbytes.value == 32 * 4 * 128 * 128 * 4
    # d/conv2 at B=256 has 512 tiles: never splits
    s2 = K.conv_shape(256, 32, 32, 64, 128, 5, 2)
    assert lib.acg_conv_splitk_plan(C.byref(s2), 0, 64, C.byref(splits), C.byref(nbytes), C.byref(ntick)) == 0
    assert splits.value == 1 and nbytes.value == 0 and ntick.value == 0
    assert lib.acg_conv_splitk_plan(None, 0, 64, C.byref(splits), C.byref(nbytes), C.byref(ntick)) == -1
    # tile table of a CONV pack + an ADJ pack of one 5x5 layer (Cin 64 -> Cout 128, ld 64 / 128)
    jobs = (_lib.PackJob * 2)()
    jobs[0] = _lib.PackJob(None, None, 0, 0, 64, 5, 5, 64, 128, 2, 1, 1, 128)       # [128][25][64]
    jobs[1] = _lib.PackJob(None, None, 0, 1, 128, 5, 5, 64, 128, 2, 1, 1, 64)       # 4 classes of [64][taps][128]
    n = lib.acg_pack_plan(jobs, 2, None, 0)
    assert n == 25 * 4 * 2 + 25 * 2 * 4          # taps x row tiles x column tiles, per job
    tiles = (C.c_int * (4 * n))()
    assert lib.acg_pack_plan(jobs, 2, tiles, n) == n
    t = [tuple(tiles[4 * i:4 * i + 4]) for i in range(n)]
    assert all(job in (0, 1) for job, _, _, _ in t) and sum(1 for x in t if x[0] == 0) == 200
    # class offsets of the ADJ pack: 4 / 6 / 6 / 9 taps (TF SAME padding of a 5x5 stride-2 filter), 64 rows x 128 columns
    offs = sorted(set(x[3] for x in t if x[0] == 1))
    assert offs == [0, 4 * 64 * 128, 10 * 64 * 128, 16 * 64 * 128]
    assert lib.acg_pack_plan(None, 2, None, 0) == -1
