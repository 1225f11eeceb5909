"""Device-side feeder.

The reference keeps the sequences on the host and, every step, gathers one frame pair per sample with one-hot masks
(`build_all_mask`, util.py:10-16, applied at train.py:231-237 and :249-263) before feeding four fp32 arrays through
placeholders.  Here the sequences live in device memory (uint8 as a decoded dataset holds them, or fp32); a step ships
2 x B int32 indices and ONE gather kernel (csrc/feeder.cu, acg_gather_frames) writes the trainer's static feed buffers:
frame t, frame t+1, action++state of frame t and the state of frame t+1, with the [-1,1] scaling of ops.py:195 applied
on the device.  64x64 Push-sized data fits easily: 60 k sequences x 7 frames x 12 KB = 5 GB of the 180 GB.
"""
import numpy as np
import torch


class DeviceFeeder:
    def __init__(self, frames, actions, device):
        """frames [N,T,64,64,3] uint8 (0..255) or float32 in [-1,1]; actions [N,T,10] float32 (action ++ state)."""
        frames = frames if isinstance(frames, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(frames))
        actions = actions if isinstance(actions, torch.Tensor) else torch.from_numpy(
            np.ascontiguousarray(np.asarray(actions, dtype=np.float32)))
        if frames.dtype not in (torch.uint8, torch.float32):
            raise ValueError("frames must be uint8 or float32")
        if frames.dim() != 5 or actions.dim() != 3 or frames.shape[:2] != actions.shape[:2]:
            raise ValueError("frames [N,T,H,W,3] and actions [N,T,A] must agree in N and T")
        if frames.shape[1] < 2:
            raise ValueError("sequences need at least two frames")
        self.device = torch.device(device)
        self.frames = frames.to(self.device).contiguous()
        self.actions = actions.to(self.device, dtype=torch.float32).contiguous()
        self.N, self.T = int(frames.shape[0]), int(frames.shape[1])

    @staticmethod
    def quantize(frames_f32):
        """[-1,1] float frames -> uint8 (the inverse of the device-side decode x/127.5 - 1, to within 1/255)."""
        a = np.asarray(frames_f32, dtype=np.float32)
        return np.clip(np.rint((a + 1.0) * 127.5), 0, 255).astype(np.uint8)

    def sample(self, batch, rng=np.random):
        """One batch of (sequence index, frame index): what `get_batch` + `boolean_mask[np.random.randint(0, T-1)]` draw
        in the reference loop (train.py:226-237).  Returns two int32 arrays of length `batch`."""
        return (rng.randint(0, self.N, size=batch).astype(np.int32),
                rng.randint(0, self.T - 1, size=batch).astype(np.int32))

    def redraw(self, sample, rng=np.random):
        """train_g runs on the LAST discriminator batch with a re-drawn frame index (train.py:258-263)."""
        return sample, rng.randint(0, self.T - 1, size=len(sample)).astype(np.int32)

    def check(self, sample, t0, batch):
        sample, t0 = np.asarray(sample), np.asarray(t0)
        if sample.shape != (batch,) or t0.shape != (batch,):
            raise ValueError("sample / t0 must have shape (%d,)" % batch)
        if sample.min() < 0 or sample.max() >= self.N or t0.min() < 0 or t0.max() >= self.T - 1:
            raise IndexError("feeder indices out of range (N=%d, T=%d)" % (self.N, self.T))

    def host_pair(self, sample, t0):
        """The same gather on the host (tests, and the reference-signature calls): img, next, act, next_state as
        float32 NumPy arrays."""
        fr = self.frames.cpu().numpy()
        ac = self.actions.cpu().numpy()
        img, nxt = fr[sample, t0], fr[sample, t0 + 1]
        if fr.dtype == np.uint8:
            img = img.astype(np.float32) * np.float32(1.0 / 127.5) - np.float32(1.0)
            nxt = nxt.astype(np.float32) * np.float32(1.0 / 127.5) - np.float32(1.0)
        return img, nxt, ac[sample, t0], ac[sample, t0 + 1, 5:]
