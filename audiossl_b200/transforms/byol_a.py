"""BYOL-A style augmentations with the API of audiossl/transforms/byol_a.py:7-141 (Mixup with a FIFO memory
bank, RandomResizeCrop with bicubic align_corners resize, MixGaussianNoise).  They sit between the fused mel kernel
and the encoder in the training recipe, so they run where the log-mel lives - on the GPU: the arithmetic is the
CUDA kernels of csrc/augment.cu (one launch per call), the random draws are made on the host from the SAME
generators in the SAME order as the reference (``np.random`` / ``random``), so that with equal seeds the per-sample
transforms reproduce the reference's outputs (tests/golden/transform.npz).  The batch-at-once versions are in
transforms/batched.py."""
import random

import numpy as np
import torch
from torch import nn

from .. import ops


def _device_only(x, who):
    if not x.is_cuda:
        raise RuntimeError("%s runs on the GPU only (no CPU fallback): it follows the fused mel kernel" % who)
    return x.contiguous().float()


class RandomResizeCrop(nn.Module):
    def __init__(self, virtual_crop_scale=(1.0, 1.5), freq_scale=(0.6, 1.5), time_scale=(0.6, 1.5)):
        super().__init__()
        assert time_scale[1] >= 1.0 and freq_scale[1] >= 1.0
        self.virtual_crop_scale, self.freq_scale, self.time_scale = virtual_crop_scale, freq_scale, time_scale
        self.interpolation = 'bicubic'

    @staticmethod
    def get_params(virtual_crop_size, in_size, time_scale, freq_scale):
        """(i, j, h, w) of the crop on the canvas; draw order of the reference: np.random height, np.random width,
        then ``random`` for the two offsets (byol_a.py:24-31)."""
        canvas_h, canvas_w = virtual_crop_size
        src_h, src_w = in_size
        h = int(np.clip(int(np.random.uniform(*freq_scale) * src_h), 1, canvas_h))
        w = int(np.clip(int(np.random.uniform(*time_scale) * src_w), 1, canvas_w))
        i = random.randint(0, canvas_h - h) if canvas_h > h else 0
        j = random.randint(0, canvas_w - w) if canvas_w > w else 0
        return i, j, h, w

    def forward(self, lms):
        """lms [C, H, W] (one clip, as the DataLoader transform sees it) -> same shape."""
        lms = _device_only(lms, "RandomResizeCrop")
        c, h, w = lms.shape
        canvas_h, canvas_w = int(h * self.virtual_crop_scale[0]), int(w * self.virtual_crop_scale[1])
        rect = self.get_params((canvas_h, canvas_w), (h, w), self.time_scale, self.freq_scale)
        r = torch.tensor([rect] * c, dtype=torch.int32, device=lms.device)
        return ops.resize_crop_fwd(lms, r, torch.empty_like(lms), canvas_h, canvas_w)

    def __repr__(self):
        return (type(self).__name__ + f'(virtual_crop_size={self.virtual_crop_scale}, '
                f'time_scale={tuple(round(s, 4) for s in self.time_scale)}, '
                f'freq_scale={tuple(round(s, 4) for s in self.freq_scale)})')


def log_mixup_exp(xa, xb, alpha):
    """log(alpha * e^xa + (1 - alpha) * e^xb + eps) for [C, H, T] clips; when the lengths differ the shorter one is
    aligned at a random frame (np.random.randint, as byol_a.py:61-82) - all three branches are one kernel launch."""
    xa, xb = _device_only(xa, "log_mixup_exp"), _device_only(xb, "log_mixup_exp")
    la, lb = xa.shape[2], xb.shape[2]
    start = 0
    if la < lb:
        start = np.random.randint(0, lb - la)
    elif la > lb:
        start = np.random.randint(0, la - lb)
    dev = xa.device
    i32 = lambda v: torch.tensor([v], dtype=torch.int32, device=dev)
    return ops.mixup_fwd(xa[None], xb[None], i32(0), torch.tensor([1.0 - alpha], dtype=torch.float32, device=dev),
                         torch.empty_like(xa[None]), zlen=i32(lb), start=i32(start))[0]


class Mixup(nn.Module):
    def __init__(self, ratio=0.4, n_memory=2000, log_mixup_exp=True):
        super().__init__()
        if not log_mixup_exp:
            raise NotImplementedError("plain (non-log) mixup is not used by any ATST recipe")
        self.ratio, self.n, self.log_mixup_exp = ratio, n_memory, log_mixup_exp
        self.memory_bank = []

    def forward(self, x):
        alpha = self.ratio * np.random.random()
        mixed = x
        if self.memory_bank:
            z = self.memory_bank[np.random.randint(len(self.memory_bank))]
            mixed = log_mixup_exp(x, z, 1. - alpha)
        self.memory_bank = (self.memory_bank + [x])[-self.n:]
        return mixed.to(torch.float)

    def __repr__(self):
        return type(self).__name__ + f'(ratio={self.ratio},n={self.n},log_mixup_exp={self.log_mixup_exp})'


class MixGaussianNoise():
    def __init__(self, ratio=0.3):
        self.ratio = ratio

    def forward(self, lms):
        x = lms.exp()
        lambd = self.ratio * np.random.rand()
        z = torch.normal(0, lambd, x.shape).exp()
        return ((1 - lambd) * x + z + torch.finfo(x.dtype).eps).log()

    def __repr__(self):
        return type(self).__name__ + f'(ratio={self.ratio})'
