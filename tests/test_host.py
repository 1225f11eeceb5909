"""Host-side logic that needs no GPU: layer tables, parameter layout, CLI flags (repair R6), schedule helpers."""
import numpy as np
import pytest
import torch


def test_layer_tables_match_the_oracle():
    from action_conditioned_gans_b200 import engine as E
    from oracle import np_ref
    for spec, ospec in ((E.g_dna_spec(5), np_ref.g_dna_spec(5)), (E.g_dna_spec(6), np_ref.g_dna_spec(6)),
                        (E.g_direct_spec(), np_ref.g_direct_spec()), (E.d_spec(), np_ref.d_spec())):
        assert [(L.name, L.kind, L.k, L.cin, L.cout, L.bn, L.bias) for L in spec] == [tuple(o) for o in ospec]
        ours = {n: s for n, s in E.variable_list(spec)}
        theirs = {k: v.shape for k, v in np_ref.init_params(ospec, np.random.RandomState(0)).items()}
        assert ours == theirs


def test_xavier_init_is_the_oracles():
    from action_conditioned_gans_b200 import engine as E
    from oracle import np_ref
    a = E.xavier_init(E.d_spec(), np.random.RandomState(7))
    b = np_ref.init_params(np_ref.d_spec(), np.random.RandomState(7))
    assert all(np.array_equal(a[k], b[k]) for k in b)
    w = a["d/conv1/weights"]
    assert abs(np.abs(w).max() - np.sqrt(6.0 / (25 * 6 + 25 * 64))) < 1e-3


def test_param_store_layout_on_cpu():
    from action_conditioned_gans_b200 import engine as E
    st = E.ParamStore(E.d_spec(), torch.device("cpu"))
    assert st.numel >= 4755137 and st.numel % 4 == 0
    for name, (off, n, shape) in st.offsets.items():
        assert off % 4 == 0 and st.views[name].shape == torch.Size(shape)
    assert st.views["d/conv6/weights"].data_ptr() == st.flat.data_ptr() + 4 * st.offsets["d/conv6/weights"][0]


def test_cli_flags_repair_r6():
    from action_conditioned_gans_b200.train import build_parser
    p = build_parser()
    a = p.parse_args(["IN", "OUT"])
    assert (a.adv, a.dna, a.loss, a.opt) == (True, True, "bce", "adam")          # README defaults
    a = p.parse_args(["IN", "OUT", "--dna", "True", "--adv", "True", "--loss", "bce", "--opt", "adam"])
    assert (a.adv, a.dna) == (True, True)                                       # BASELINE.json configs[0] spelling
    a = p.parse_args(["IN", "OUT", "--dna", "False", "--adv", "--loss", "wass", "--opt", "rmsprop"])
    assert (a.adv, a.dna, a.loss, a.opt) == (True, False, "wass", "rmsprop")
    from action_conditioned_gans_b200.test import build_parser as tp
    t = tp().parse_args(["M", "F.npy", "A.npy", "O", "--dna"])
    assert t.dna is True and t.model_path == "M"


def test_bad_flag_values_raise_like_the_reference():
    from action_conditioned_gans_b200.trainer import Trainer
    with pytest.raises(ValueError, match="unexpected loss argument"):
        Trainer(None, True, "hinge", "adam", True)
    with pytest.raises(ValueError, match="unexpected opt argument"):
        Trainer(None, True, "bce", "sgd", True)


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU failure mode")
def test_no_cpu_fallback():
    from action_conditioned_gans_b200 import kernels as K
    from action_conditioned_gans_b200.trainer import Trainer
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        Trainer(None, True, "bce", "adam", True, batch_size=2)
    with pytest.raises(RuntimeError, match="no CPU path"):
        K.dna_fwd(torch.zeros(1, 8, 8, 25), torch.zeros(1, 8, 8, 3), torch.zeros(1, 8, 8, 3), 5)


def test_frame_pair_masks_and_schedule():
    from action_conditioned_gans_b200.util import build_all_mask
    m = build_all_mask(7)
    assert m.shape == (6, 7) and m.sum() == 6 and not m[:, 6].any()           # t in [0,5], paired with t+1
    end = np.roll(m, 1, axis=1)
    assert (np.argmax(end, 1) == np.argmax(m, 1) + 1).all()


def test_synthetic_data_shapes():
    from action_conditioned_gans_b200.feeder import DeviceFeeder
    from action_conditioned_gans_b200.train import synthetic_push
    img, act = synthetic_push(4, 0)
    assert img.shape == (4, 7, 64, 64, 3) and act.shape == (4, 7, 10)
    assert img.dtype == np.uint8 and act.dtype == np.float32          # frames as a decoded dataset holds them
    back = img.astype(np.float32) / 127.5 - 1.0
    assert np.array_equal(DeviceFeeder.quantize(back), img)            # decode / quantize round trip


def test_feeder_host_logic():
    """Frame-pair sampling of util.py:10-16 / train.py:226-263 as index draws: t in [0, T-2], paired with t+1; train_g
    re-draws the frame index on the same sequences; out-of-range indices are refused before any kernel runs."""
    from action_conditioned_gans_b200.feeder import DeviceFeeder
    from action_conditioned_gans_b200.train import synthetic_push
    img, act = synthetic_push(6, 1)
    fd = DeviceFeeder(img, act, "cpu")                                 # host logic only: no kernel is called
    rng = np.random.RandomState(0)
    sample, t0 = fd.sample(64, rng)
    assert sample.dtype == np.int32 and t0.dtype == np.int32 and sample.max() < 6 and t0.max() <= 5 and t0.min() >= 0
    s2, t2 = fd.redraw(sample, rng)
    assert np.array_equal(s2, sample) and not np.array_equal(t2, t0)
    a, b, ac, st = fd.host_pair(sample, t0)
    assert a.shape == (64, 64, 64, 3) and np.array_equal(st, act[sample, t0 + 1, 5:]) and np.array_equal(ac, act[sample, t0])
    with pytest.raises(IndexError):
        fd.check(sample, np.full(64, 6, np.int32), 64)
    with pytest.raises(ValueError):
        DeviceFeeder(img[:, :1], act[:, :1], "cpu")


def test_checkpoint_pruning_is_atomic_and_keeps_five(tmp_path):
    from action_conditioned_gans_b200 import train as T

    class FakeTrainer:
        def save(self, path):
            np.savez(path, x=np.zeros(3))

    for step in range(0, 900, 100):
        T.save_checkpoint(FakeTrainer(), str(tmp_path), step)
    left = sorted(p.name for p in tmp_path.iterdir())
    assert left == ["model%d.npz" % s for s in (400, 500, 600, 700, 800)]          # tf.train.Saver max_to_keep=5
    assert T.latest_checkpoint(str(tmp_path)).endswith("model800.npz")
    assert not any(n.endswith(".tmp.npz") for n in left)


def test_save_samples_layout(tmp_path):
    from action_conditioned_gans_b200.util import save_samples
    a = np.random.uniform(-1, 1, (2, 1, 64, 64, 3))
    save_samples(str(tmp_path), a, a, a, 3)
    assert (tmp_path / "sample3" / "vid1" / "generated0.png").exists()
    assert (tmp_path / "sample3" / "vid0" / "ground_truth0.png").exists()
    save_samples(str(tmp_path), np.repeat(a, 3, 1), np.repeat(a, 3, 1), np.array([0]), 4, gif=True)
    assert (tmp_path / "sample4" / "vid0" / "generated.gif").exists()


def test_peer_slot_plan_is_symmetric(built_lib):
    """Every rank requests its exchange slots in the same order -> same byte offsets in every mailbox (host logic of
    action_conditioned_gans_b200/peer.py; acg_peer_slot_bytes needs no device)."""
    from action_conditioned_gans_b200 import peer
    caps = [64, 128, 256, 512, 1024, 2, 64]
    plans = [peer.SlotPlan(8) for _ in range(2)]
    offs = [[p.take(c) for c in caps] for p in plans]
    assert offs[0] == offs[1]
    prev_end = 0
    for (off, idx), c in zip(offs[0], caps):
        assert off % 256 == 0 and off >= prev_end
        prev_end = off + peer.slot_bytes(c, 8)
        assert peer.slot_bytes(c, 8) >= 128 + 2 * 8 * c * 16          # 16-byte cells: value halves + epoch tags
    assert built_lib.acg_peer_slot_bytes(16, 9) == -1          # more than ACG_MAX_PEERS ranks
    small = peer.SlotPlan(2, segment_bytes=8192)
    small.take(64)
    with pytest.raises(RuntimeError, match="exhausted"):
        small.take(1024)


def test_custom_ops_are_registered_and_cuda_only():
    """The acg:: torch.library ops exist with schemas and fake (shape) kernels, and refuse CPU tensors: there is no CPU
    fallback behind the PyTorch boundary."""
    from action_conditioned_gans_b200 import torch_ops  # noqa: F401
    for name in ("dna", "conv2d", "conv2d_dgrad", "conv2d_wgrad", "bn_act", "bias_act", "frame_losses", "dlogit_loss",
                 "adam_step", "rmsprop_step", "generator_transform", "generator", "discriminator"):
        assert hasattr(torch.ops.acg, name), name
    assert "Tensor logits, Tensor img, SymInt ksize" in str(torch.ops.acg.dna.default._schema) or \
        "Tensor logits, Tensor img, int ksize" in str(torch.ops.acg.dna.default._schema)
    with pytest.raises(NotImplementedError):
        torch.ops.acg.dna(torch.randn(1, 8, 8, 25), torch.randn(1, 8, 8, 3), 5)
    # shape inference without a device (fake kernels)
    meta = torch.ops.acg.conv2d(torch.empty(2, 64, 64, 3, device="meta"), torch.empty(5, 5, 3, 32, device="meta"), 2, True, False)
    assert tuple(meta.shape) == (2, 32, 32, 32)
