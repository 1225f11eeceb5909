"""test.py of the reference (test.py:15-47): restore the latest checkpoint, roll out 6 steps, write GIFs.

    python test.py MODEL_DIR FRAMES.npy ACTIONS.npy OUT [--dna [True|False]]

Repair R4: the arguments are actually parsed.  Samples [32:96] and frames [:, 6:] are used as in test.py:35-40.
"""
import argparse

import numpy as np

from .train import latest_checkpoint, str2bool
from .trainer import Trainer
from .util import save_samples


def build_parser():
    p = argparse.ArgumentParser()
    p.add_argument("model_path", type=str)
    p.add_argument("input_frame_path", type=str)
    p.add_argument("input_action_path", type=str)
    p.add_argument("output_path", type=str)
    p.add_argument("--dna", type=str2bool, nargs="?", const=True, default=True)
    p.add_argument("--precision", type=str, default="bf16", choices=["bf16", "fp32"])
    return p


def main(argv=None):
    args = build_parser().parse_args(argv)
    seq = np.load(args.input_frame_path)
    act = np.load(args.input_action_path)
    seq_batch, act_batch = seq[32:96], act[32:96]
    trainer = Trainer(None, True, "bce", "adam", args.dna, batch_size=seq_batch.shape[0], precision=args.precision)
    ckpt = latest_checkpoint(args.model_path)
    if ckpt is None:
        raise FileNotFoundError("no model*.npz checkpoint under %s" % args.model_path)
    trainer.restore(ckpt)
    test_g_out, gt = trainer.test_sequence(seq_batch[:, 6:], seq_batch[:, 6:], act_batch[:, 6:])
    for i in range(test_g_out.shape[1]):
        save_samples(args.output_path, seq_batch[:, [6 + (k + 1) * 2 for k in range(6)]], test_g_out, np.array([0]),
                     i, gif=True)


if __name__ == "__main__":
    main()
