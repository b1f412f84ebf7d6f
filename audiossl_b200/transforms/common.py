"""Waveform / feature helpers with the names and call semantics of audiossl/transforms/common.py:9-117.
They are thin tensor ops (index arithmetic, padding); device-agnostic so they also run on cuda tensors."""
import numpy as np
import torch
from torch.nn import functional as F


class CustomAudioTransform:
    def __repr__(self):
        return type(self).__name__ + "()"


class Identity(CustomAudioTransform):
    def __call__(self, signal):
        return signal


class GaussianNoise(CustomAudioTransform):
    def __init__(self, g):
        self.g = g

    def __call__(self, signal):
        return signal + self.g * torch.randn_like(signal)


def _right_pad(signal, target):
    missing = target - signal.shape[-1]
    return F.pad(signal, (0, missing)) if missing > 0 else signal


class PadToSize(CustomAudioTransform):
    def __init__(self, size: int):
        self.size = size

    def __call__(self, signal):
        return _right_pad(signal, self.size) if signal.shape[1] < self.size else signal


class ToSizeN(CustomAudioTransform):
    """pad or truncate-pad to the nearest multiple of `size` (rounding up past the half-way point)."""

    def __init__(self, size: int):
        self.size = size

    def __call__(self, signal):
        whole, rest = divmod(signal.shape[1], self.size)
        n = whole + 1 if (rest > self.size // 2 or whole == 0) else whole
        return F.pad(signal, (0, self.size * n - signal.shape[1]))


class CentralCrop(CustomAudioTransform):
    def __init__(self, size: int, pad: bool = True):
        self.size, self.pad = size, pad

    def __call__(self, signal):
        n = signal.shape[-1]
        if n < self.size:
            return _right_pad(signal, self.size) if self.pad else signal
        start = (n - self.size) // 2
        return signal[..., start:start + self.size]


class RandomCrop(CustomAudioTransform):
    def __init__(self, size: int, pad: bool = True):
        self.size, self.pad = size, pad

    def __call__(self, signal):
        n = signal.shape[-1]
        if signal.shape[1] < self.size:
            return _right_pad(signal, self.size) if self.pad else signal
        start = np.random.randint(0, n - self.size + 1)
        return signal[:, start:start + self.size]


class Normalize(CustomAudioTransform):
    def __init__(self, std_mean=None, reduce_dim=None):
        self.std_mean, self.reduce_dim = std_mean, reduce_dim

    def __call__(self, input):
        if self.std_mean is not None:
            std, mean = self.std_mean
        elif self.reduce_dim is not None:
            std, mean = torch.std_mean(input, dim=self.reduce_dim, keepdim=True)
        else:
            std, mean = torch.std_mean(input)
        return (input - mean) / (std + 1e-6)


class MinMax(CustomAudioTransform):
    def __init__(self, min, max):
        self.min, self.max = min, max

    def __call__(self, input):
        lo, hi = (torch.min(input), torch.max(input)) if self.min is None else (self.min, self.max)
        return (input - lo) / (hi - lo) * 2. - 1.


class div(CustomAudioTransform):
    def __init__(self, value=100):
        self.value = value

    def __call__(self, input):
        input /= 100  # the reference divides by the literal 100 whatever `value` is (common.py:112-117)
        return input
