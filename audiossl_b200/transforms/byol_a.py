"""BYOL-A style augmentations with the API of audiossl/transforms/byol_a.py:7-141 (Mixup with a FIFO memory
bank, RandomResizeCrop with bicubic align_corners resize, MixGaussianNoise).  These sit between the fused
mel kernel and the encoder in the training recipe; they are tensor-level ops that run on whatever device the
log-mel lives on (cuda in this framework).  SURVEY.md section 8f f1: batching them on the device is a later row."""
import random

import numpy as np
import torch
from torch import nn
from torch.nn import functional as F


class RandomResizeCrop(nn.Module):
    def __init__(self, virtual_crop_scale=(1.0, 1.5), freq_scale=(0.6, 1.5), time_scale=(0.6, 1.5)):
        super().__init__()
        assert time_scale[1] >= 1.0 and freq_scale[1] >= 1.0
        self.virtual_crop_scale, self.freq_scale, self.time_scale = virtual_crop_scale, freq_scale, time_scale
        self.interpolation = 'bicubic'

    @staticmethod
    def get_params(virtual_crop_size, in_size, time_scale, freq_scale):
        canvas_h, canvas_w = virtual_crop_size
        src_h, src_w = in_size
        h = np.clip(int(np.random.uniform(*freq_scale) * src_h), 1, canvas_h)
        w = np.clip(int(np.random.uniform(*time_scale) * src_w), 1, canvas_w)
        i = random.randint(0, canvas_h - h) if canvas_h > h else 0
        j = random.randint(0, canvas_w - w) if canvas_w > w else 0
        return i, j, h, w

    def forward(self, lms):
        c, h, w = lms.shape
        canvas_h, canvas_w = int(h * self.virtual_crop_scale[0]), int(w * self.virtual_crop_scale[1])
        canvas = torch.zeros((c, canvas_h, canvas_w), dtype=torch.float, device=lms.device)
        top, left = (canvas_h - h) // 2, (canvas_w - w) // 2
        canvas[:, top:top + h, left:left + w] = lms
        i, j, ch, cw = self.get_params((canvas_h, canvas_w), (h, w), self.time_scale, self.freq_scale)
        crop = canvas[:, i:i + ch, j:j + cw]
        out = F.interpolate(crop.unsqueeze(0), size=(h, w), mode=self.interpolation, align_corners=True)
        return out.squeeze(0).to(torch.float)

    def __repr__(self):
        return (type(self).__name__ + f'(virtual_crop_size={self.virtual_crop_scale}, '
                f'time_scale={tuple(round(s, 4) for s in self.time_scale)}, '
                f'freq_scale={tuple(round(s, 4) for s in self.freq_scale)})')


def log_mixup_exp(xa, xb, alpha):
    """log(alpha * e^xa + (1 - alpha) * e^xb + eps); a shorter operand is mixed into a random window."""
    ea, eb = xa.exp(), xb.exp()
    la, lb = ea.shape[2], eb.shape[2]
    eps = torch.finfo(ea.dtype).eps
    if la < lb:
        s = np.random.randint(0, lb - la)
        return torch.log(alpha * ea + (1. - alpha) * eb[:, :, s:s + la] + eps)
    if la > lb:
        s = np.random.randint(0, la - lb)
        ea[:, :, s:s + lb] = alpha * ea[:, :, s:s + lb] + (1. - alpha) * eb
        return torch.log(ea + eps)
    return torch.log(alpha * ea + (1. - alpha) * eb + eps)


class Mixup(nn.Module):
    def __init__(self, ratio=0.4, n_memory=2000, log_mixup_exp=True):
        super().__init__()
        self.ratio, self.n, self.log_mixup_exp = ratio, n_memory, log_mixup_exp
        self.memory_bank = []

    def forward(self, x):
        alpha = self.ratio * np.random.random()
        mixed = x
        if self.memory_bank:
            z = self.memory_bank[np.random.randint(len(self.memory_bank))]
            mixed = log_mixup_exp(x, z, 1. - alpha) if self.log_mixup_exp else alpha * z + (1. - alpha) * x
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
