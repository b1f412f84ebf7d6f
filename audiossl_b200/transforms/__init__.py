from .common import (CentralCrop, GaussianNoise, Identity, MinMax, Normalize, PadToSize, RandomCrop, ToSizeN, div)
from .byol_a import MixGaussianNoise, Mixup, RandomResizeCrop, log_mixup_exp
from .mel import LogMelSpectrogram
from .batched import BatchedMixup, BatchedRandomResizeCrop

__all__ = ["CentralCrop", "GaussianNoise", "Identity", "MinMax", "Normalize", "PadToSize", "RandomCrop", "ToSizeN",
           "div", "MixGaussianNoise", "Mixup", "RandomResizeCrop", "log_mixup_exp", "LogMelSpectrogram", "BatchedMixup", "BatchedRandomResizeCrop"]
