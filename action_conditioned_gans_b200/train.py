"""train.py of the reference (train.py:179-333): the host training loop and the CLI, on the acg_b200 Trainer.

    python train.py IN OUT [--adv [True|False]] [--loss bce|wass] [--opt adam|rmsprop] [--dna [True|False]]

Repair R6: --adv / --dna accept `--adv`, `--adv True`, `--adv False` and default to True (README.md:30-49).
IN is a directory of .npz shards {images [N,T,64,64,3] (uint8, or float in [-1,1]), actions [N,T,10]}, or the word
`synthetic` (the Push TFRecords of ops.py:141-223 are not available offline; only their output shapes matter here).
90 % of the shards train, the last 10 % are held out for the in-loop rollout, as `build_tfrecord_input(..., .9, ...)`
does (train.py:189-196).

The sequences are uploaded once and stay resident in device memory (feeder.DeviceFeeder): every step ships 2 x B
indices and a gather kernel does the frame-pair sampling of util.py:10-16 / train.py:231-237 on the device.  Scalars go
to `OUT/logs/train.jsonl` and, as the reference's `tf.summary.FileWriter` does (train.py:211-214,275), to TensorBoard
event files in `OUT/logs`; checkpoints keep the newest 5 like `tf.train.Saver()` and are written atomically.
"""
import argparse
import glob
import json
import os
import time

import numpy as np

from .feeder import DeviceFeeder
from .trainer import BATCH_SIZE, SUMMARY_KEYS, Trainer
from .util import save_samples

HISTORY_LENGTH = 1      # train.py:16
PRETRAIN_ITER = 20      # train.py:24
TRAIN_ITER = 60000      # train.py:25
NUM_FRAMES = 7          # ops.py: frames 6,8,...,18 (synthetic data only; real shards carry their own T)
MAX_TO_KEEP = 5         # tf.train.Saver() default


def str2bool(v):
    if isinstance(v, bool):
        return v
    if v.lower() in ("true", "1", "yes", "y", "t"):
        return True
    if v.lower() in ("false", "0", "no", "n", "f"):
        return False
    raise argparse.ArgumentTypeError("expected True or False, got %r" % v)


def build_parser():
    p = argparse.ArgumentParser()
    p.add_argument("input_path", type=str)
    p.add_argument("output_path", type=str)
    p.add_argument("--adv", type=str2bool, nargs="?", const=True, default=True)
    p.add_argument("--loss", type=str, default="bce")
    p.add_argument("--opt", type=str, default="adam")
    p.add_argument("--dna", type=str2bool, nargs="?", const=True, default=True)
    # constants of train.py:16-25 promoted to optional flags with the same defaults
    p.add_argument("--iters", type=int, default=TRAIN_ITER)
    p.add_argument("--pretrain_iters", type=int, default=PRETRAIN_ITER)
    p.add_argument("--batch_size", type=int, default=BATCH_SIZE)
    p.add_argument("--seed", type=int, default=7)
    p.add_argument("--precision", type=str, default="bf16", choices=["bf16", "fp32"])
    return p


def synthetic_push(n_seq, seed, T=NUM_FRAMES):
    """Push-shaped synthetic sequences as a decoded dataset holds them: uint8 frames [N,T,64,64,3], action++state
    [N,T,10]."""
    rng = np.random.RandomState(seed)
    base = rng.uniform(-1, 1, (n_seq, 1, 64, 64, 3)).astype(np.float32)
    drift = np.cumsum(0.05 * rng.randn(n_seq, T, 64, 64, 3).astype(np.float32), axis=1)
    return DeviceFeeder.quantize(np.clip(base + drift, -1, 1)), rng.randn(n_seq, T, 10).astype(np.float32)


def load_shards(files):
    imgs, acts = [], []
    for f in files:
        with np.load(f) as z:
            im, ac = z["images"], z["actions"]
        imgs.append(im if im.dtype == np.uint8 else DeviceFeeder.quantize(im))
        acts.append(ac.astype(np.float32))
    return np.concatenate(imgs), np.concatenate(acts)


def open_data(input_path, batch, seed, device):
    """-> (train feeder, held-out feeder).  ops.py:189-196: the first 90 % of the (sorted) files train, the rest test."""
    if input_path == "synthetic" or not os.path.isdir(input_path):
        print("input_path %r is not a directory of .npz shards: using synthetic Push-shaped data" % input_path)
        tr, te = synthetic_push(max(4 * batch, 64), seed), synthetic_push(max(batch, 16), seed + 1)
    else:
        files = sorted(glob.glob(os.path.join(input_path, "*.npz")))
        if not files:
            raise FileNotFoundError("no .npz shards under %s" % input_path)
        split = int(np.floor(0.9 * len(files)))
        split = min(max(split, 1), len(files) - 1) if len(files) > 1 else 1
        tr = load_shards(files[:split])
        te = load_shards(files[split:]) if len(files) > 1 else tr
    return DeviceFeeder(tr[0], tr[1], device), DeviceFeeder(te[0], te[1], device)


def _ckpt_step(f):
    return int(os.path.basename(f)[5:-4])


def latest_checkpoint(model_dir):
    files = glob.glob(os.path.join(model_dir, "model*.npz"))
    if not files:
        return None
    return max(files, key=_ckpt_step)


def save_checkpoint(trainer, model_dir, step, max_to_keep=MAX_TO_KEEP):
    """tf.train.Saver().save (train.py:215,274): the file appears atomically (temp file + os.replace, so a crash
    mid-write never leaves a truncated model*.npz for latest_checkpoint to pick up) and only the newest `max_to_keep`
    checkpoints stay on disk (each is ~115 MB: weights + three sets of optimizer slots)."""
    final = os.path.join(model_dir, "model{:d}.npz".format(step))
    tmp = final + ".tmp.npz"
    trainer.save(tmp)
    os.replace(tmp, final)
    files = sorted(glob.glob(os.path.join(model_dir, "model*.npz")), key=_ckpt_step)
    files = [f for f in files if not f.endswith(".tmp.npz")]
    for f in files[:-max_to_keep]:
        os.remove(f)
    return final


class SummaryLog:
    """The 7 scalars of train.py:104-112 to JSONL and to TensorBoard event files (tf.summary.FileWriter, train.py:211)."""

    def __init__(self, log_dir):
        os.makedirs(log_dir, exist_ok=True)
        self.jsonl = open(os.path.join(log_dir, "train.jsonl"), "a")
        try:
            from torch.utils.tensorboard import SummaryWriter
            self.tb = SummaryWriter(log_dir)
        except Exception as e:                      # tensorboard missing: the JSONL still has everything
            print("TensorBoard writer unavailable (%s): scalars go to train.jsonl only" % e)
            self.tb = None

    def add(self, scalars, step, **extra):
        self.jsonl.write(json.dumps(dict(scalars or {}, iteration=step, **extra)) + "\n")
        self.jsonl.flush()
        if self.tb is not None:
            for k in SUMMARY_KEYS:
                if scalars and k in scalars:
                    self.tb.add_scalar(k, scalars[k], step)
            self.tb.flush()

    def close(self):
        self.jsonl.close()
        if self.tb is not None:
            self.tb.close()


def train(input_path, output_path, test_output_path, log_dir, model_dir, arg_adv, arg_loss, arg_opt, arg_transform,
          iters=TRAIN_ITER, pretrain_iters=PRETRAIN_ITER, batch_size=BATCH_SIZE, seed=7, precision="bf16"):
    """train.py:179-309: 20 pre-train G iterations, then per iteration D_per_G x train_d (5 for wass, else 1) on fresh
    batches and one train_g on the last D batch with a re-drawn frame index (train.py:217-263)."""
    np.random.seed(seed)                                         # train.py:14
    trainer = Trainer(None, arg_adv, arg_loss, arg_opt, arg_transform, batch_size=batch_size, seed=seed,
                      precision=precision)
    data, test_data = open_data(input_path, batch_size, seed, trainer.device)
    log = SummaryLog(log_dir)
    D_per_G = 5 if arg_loss == "wass" else 1                     # train.py:217-220
    B = batch_size

    t0 = time.time()
    for i in range(iters):
        if i < pretrain_iters:
            sample, t = data.sample(B)                           # get_batch + boolean_mask[randint] (train.py:226-230)
            trainer.pretrain_g_indexed(data, sample, t)
            print("pre-train iter: " + str(i))
            continue
        summ = None
        for j in range(D_per_G):
            sample, t = data.sample(B)
            make_summ = (i % 100 == 0) and (j == D_per_G - 1)
            summ = trainer.train_d_indexed(data, sample, t, summarize=make_summ)
        sample, t = data.redraw(sample)                          # same sequences, new frame index (train.py:258-263)
        gen_next_frames = trainer.train_g_indexed(data, sample, t)
        if i % 100 == 0:
            print("Iteration {:d}".format(i))
            a, b, _, _ = data.host_pair(sample[:32], t[:32])
            save_samples(output_path, np.expand_dims(a, 1), np.expand_dims(gen_next_frames[:32], 1),
                         np.expand_dims(b, 1), i)
            save_checkpoint(trainer, model_dir, i)
            log.add(summ, i, wall_s=time.time() - t0)
        if i % 500 == 0:
            idx = np.random.randint(0, test_data.N, size=B)
            timg = test_data.frames[torch_index(idx, test_data)]
            tact = test_data.actions[torch_index(idx, test_data)]
            predicted, truth = rollout(trainer, timg, tact)
            save_samples(test_output_path, truth[:16], predicted[:16], truth[:16], i)
    log.close()
    return trainer


def torch_index(idx, feeder):
    import torch
    return torch.as_tensor(idx, dtype=torch.long, device=feeder.device)


def rollout(trainer, test_input, test_actions):
    """The in-loop evaluation of train.py:286-299: T-1 recursive steps with action index j, as ONE captured graph with
    frames and state fed back on the device (Trainer.rollout).  test_input: device frames [B,T,64,64,3] (uint8 or float),
    test_actions [B,T,10].  Returns (predicted [B,T-1,64,64,3], ground truth [B,T,64,64,3]) as float NumPy arrays."""
    import torch
    T = int(test_input.shape[1])
    first = test_input[:, 0]
    pred = trainer.rollout(first.contiguous(), test_actions.contiguous(), steps=T - 1, action_stride=1)
    truth = test_input
    if truth.dtype == torch.uint8:
        truth = truth.float() / 127.5 - 1.0
    return pred.permute(1, 0, 2, 3, 4).contiguous().cpu().numpy(), truth.cpu().numpy()


def main(argv=None):
    args = build_parser().parse_args(argv)
    output_path = os.path.join(args.output_path, "train_output")
    test_output_path = os.path.join(args.output_path, "test_output")
    model_dir = os.path.join(args.output_path, "models")
    log_dir = os.path.join(args.output_path, "logs")
    os.makedirs(args.output_path)                                # train.py:324 (fails if OUT exists)
    os.makedirs(model_dir)
    train(args.input_path, output_path, test_output_path, log_dir, model_dir, args.adv, args.loss, args.opt, args.dna,
          iters=args.iters, pretrain_iters=args.pretrain_iters, batch_size=args.batch_size, seed=args.seed,
          precision=args.precision)


if __name__ == "__main__":
    main()
