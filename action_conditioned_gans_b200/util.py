"""util.py of the reference: frame-pair masks (util.py:10-16) and the sample dump (util.py:18-57; PIL instead of
matplotlib / imageio, which this image does not have)."""
import os

import numpy as np


def build_all_mask(num_frame):
    """util.py:10-16: boolean [num_frame-1, num_frame]; row i selects frame i (the last frame has no successor)."""
    return np.eye(num_frame, dtype=bool)[:num_frame - 1]


def _to_uint8(a):
    return ((255.0 / 2) * (np.asarray(a, dtype=np.float32) + 1.0)).astype(np.uint8)


def save_samples(output_path, input_sample, generated_sample, ground_truth, sample_number, gif=False):
    """util.py:18-57: same directory layout and file names (sample{n}/vid{i}/frame{j}.png, generated{j}.png,
    ground_truth{j}.png, or *.gif)."""
    from PIL import Image
    input_sample, generated_sample = _to_uint8(input_sample), _to_uint8(generated_sample)
    ground_truth = _to_uint8(ground_truth)
    save_folder = os.path.join(output_path, "sample{:d}".format(sample_number))
    os.makedirs(save_folder, exist_ok=True)

    def dump(vid, stem, gif_name):
        if gif:
            frames = [Image.fromarray(f) for f in vid]
            frames[0].save(os.path.join(vid_folder, gif_name), save_all=True, append_images=frames[1:], duration=250,
                           loop=0)
        else:
            for j in range(int(vid.shape[0])):
                Image.fromarray(vid[j]).save(os.path.join(vid_folder, "{}{:d}.png".format(stem, j)))

    for i in range(input_sample.shape[0]):
        vid_folder = os.path.join(save_folder, "vid{:d}".format(i))
        os.makedirs(vid_folder, exist_ok=True)
        dump(input_sample[i], "frame", "ground_truth.gif")
        dump(generated_sample[i], "generated", "generated.gif")
        if not gif:
            dump(ground_truth[i], "ground_truth", "")
