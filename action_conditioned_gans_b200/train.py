"""train.py of the reference (train.py:179-333): the host training loop and the CLI, on the acg_b200 Trainer.

    python train.py IN OUT [--adv [True|False]] [--loss bce|wass] [--opt adam|rmsprop] [--dna [True|False]]

Repair R6: --adv / --dna accept `--adv`, `--adv True`, `--adv False` and default to True (README.md:30-49).
IN is a directory of .npz shards {images [N,T,64,64,3] in [-1,1], actions [N,T,10]}, or the word `synthetic`
(the Push TFRecords of ops.py:141-223 are not available offline; only their output shapes matter here).
"""
import argparse
import glob
import json
import os
import time

import numpy as np

from .trainer import BATCH_SIZE, Trainer
from .util import build_all_mask, save_samples

HISTORY_LENGTH = 1      # train.py:16
PRETRAIN_ITER = 20      # train.py:24
TRAIN_ITER = 60000      # train.py:25
NUM_FRAMES = 7          # ops.py: frames 6,8,...,18


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


class SyntheticPush:
    """Push-shaped synthetic sequences: img [B,7,64,64,3] in [-1,1], action++state [B,7,10]."""

    def __init__(self, batch, seed):
        self.B, self.rng = batch, np.random.RandomState(seed)

    def get_batch(self):
        B = self.B
        base = self.rng.uniform(-1, 1, (B, 1, 64, 64, 3)).astype(np.float32)
        drift = np.cumsum(0.05 * self.rng.randn(B, NUM_FRAMES, 64, 64, 3).astype(np.float32), axis=1)
        img = np.clip(base + drift, -1, 1)
        act = self.rng.randn(B, NUM_FRAMES, 10).astype(np.float32)
        return img, act


class NpzShards:
    def __init__(self, path, batch, seed):
        self.files = sorted(glob.glob(os.path.join(path, "*.npz")))
        if not self.files:
            raise FileNotFoundError("no .npz shards under %s" % path)
        self.B, self.rng = batch, np.random.RandomState(seed)

    def get_batch(self):
        with np.load(self.files[self.rng.randint(len(self.files))]) as f:
            img, act = f["images"], f["actions"]
        idx = self.rng.randint(0, img.shape[0], size=self.B)
        return img[idx].astype(np.float32), act[idx].astype(np.float32)


def open_data(input_path, batch, seed):
    if input_path == "synthetic" or not os.path.isdir(input_path):
        print("input_path %r is not a directory of .npz shards: using synthetic Push-shaped data" % input_path)
        return SyntheticPush(batch, seed)
    return NpzShards(input_path, batch, seed)


def latest_checkpoint(model_dir):
    files = glob.glob(os.path.join(model_dir, "model*.npz"))
    if not files:
        return None
    return max(files, key=lambda f: int(os.path.basename(f)[5:-4]))


def train(input_path, output_path, test_output_path, log_dir, model_dir, arg_adv, arg_loss, arg_opt, arg_transform,
          iters=TRAIN_ITER, pretrain_iters=PRETRAIN_ITER, batch_size=BATCH_SIZE, seed=7, precision="bf16"):
    """train.py:179-309: 20 pre-train G iterations, then per iteration D_per_G x train_d (5 for wass, else 1) on fresh
    batches and one train_g on the last D batch with a re-drawn frame index (train.py:217-263)."""
    np.random.seed(seed)                                         # train.py:14
    data = open_data(input_path, batch_size, seed)
    test_data = open_data(input_path, batch_size, seed + 1)
    boolean_mask = build_all_mask(NUM_FRAMES)                    # train.py:202
    trainer = Trainer(None, arg_adv, arg_loss, arg_opt, arg_transform, batch_size=batch_size, seed=seed,
                      precision=precision)
    os.makedirs(log_dir, exist_ok=True)
    log = open(os.path.join(log_dir, "train.jsonl"), "a")
    D_per_G = 5 if arg_loss == "wass" else 1                     # train.py:217-220
    B = batch_size

    def draw(img, act):
        start_mask = boolean_mask[np.random.randint(0, len(boolean_mask), size=B)]
        end_mask = np.roll(start_mask, 1, axis=1)
        state = act[:, :, 5:]                                    # next_state_train, train.py:199
        return img[start_mask], img[end_mask], act[start_mask], state[end_mask]

    t0 = time.time()
    for i in range(iters):
        if i < pretrain_iters:
            img, act = data.get_batch()
            a, b, c, d = draw(img, act)
            trainer.pretrain_g(a, b, c, d)
            print("pre-train iter: " + str(i))
            continue
        summ = None
        for j in range(D_per_G):
            img, act = data.get_batch()
            a, b, c, d = draw(img, act)
            make_summ = (i % 100 == 0) and (j == D_per_G - 1)
            summ = trainer.train_d(a, b, c, summarize=make_summ)
        a, b, c, d = draw(img, act)
        gen_next_frames = trainer.train_g(a, b, c, d)
        if i % 100 == 0:
            print("Iteration {:d}".format(i))
            save_samples(output_path, np.expand_dims(a[:32], 1), np.expand_dims(gen_next_frames[:32], 1),
                         np.expand_dims(b[:32], 1), i)
            trainer.save(os.path.join(model_dir, "model{:d}.npz".format(i)))
            rec = dict(summ or {}, iteration=i, wall_s=time.time() - t0)
            log.write(json.dumps(rec) + "\n")
            log.flush()
        if i % 500 == 0:
            timg, tact = test_data.get_batch()
            predicted, _ = rollout(trainer, timg, tact)
            save_samples(test_output_path, timg[:16], predicted[:16], timg[:16], i)
    log.close()
    return trainer


def rollout(trainer, test_input, test_actions):
    """The in-loop evaluation of train.py:286-299: T-1 recursive steps with action index j."""
    predicted = []
    current_frame = test_input[:, 0]
    current_state = test_actions[:, 0, 5:]
    for j in range(test_input.shape[1] - 1):
        acs = np.concatenate((test_actions[:, j, :5], current_state), axis=1)
        out, st, _ = trainer.test(current_frame, test_input[:, j + 1], acs)
        predicted.append(out)
        current_frame = out
        if st is not None:
            current_state = st
    return np.transpose(np.array(predicted), (1, 0, 2, 3, 4)), None


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
