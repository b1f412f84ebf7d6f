"""ATSTTrainTransform - audiossl/methods/atst/transform.py:12-74 with the same constructor and return value
(``([crop1, crop2], [len1, len2])``); the mel_feature stage is the fused CUDA kernel, so the waveform must
live on the GPU (the reference computes it on CPU DataLoader workers)."""
import random

from torch.nn import functional as F

from ...transforms.byol_a import Mixup, RandomResizeCrop
from ...transforms.common import RandomCrop
from ...transforms.mel import LogMelSpectrogram

random.seed(1234)


class _Compose:
    def __init__(self, transforms):
        self.transforms = transforms

    def __call__(self, x):
        for t in self.transforms:
            x = t(x)
        return x


class ATSTTrainTransform:
    def __init__(self, sr=16000, mask_ratio=0.75, different_positive=True, anchor_len=(6., 6.),
                 positive_len=(6., 6.), virtual_crop=1.5):
        self.different_positive = different_positive
        self.anchor_len, self.positive_len = anchor_len, positive_len
        self.max_positive_len = max(self.positive_len + self.anchor_len)
        self.mel_feature = LogMelSpectrogram(sr, win_length=1024)
        self.positivecrop = _Compose([RandomCrop(16000 * 6), self.mel_feature])
        self.positive_transform1 = _Compose([Mixup(), RandomResizeCrop((1, virtual_crop))])
        self.positive_transform2 = _Compose([Mixup(), RandomResizeCrop((1, virtual_crop))])

    def _view(self, input, seconds, aug):
        n = int(seconds * 16000)
        self.positivecrop.transforms[0].size = n
        return self.positivecrop(input), n

    def __call__(self, input):
        anchor_len = random.uniform(*self.anchor_len)
        crop1, n1 = self._view(input, anchor_len, None)
        if self.different_positive:
            positive_len = random.uniform(*self.positive_len)
            crop2, n2 = self._view(input, positive_len, None)
        else:
            crop2, n2 = crop1, n1
        max_frames = int((self.max_positive_len * 16000) // 160)
        crops = [F.pad(self.positive_transform1(crop1), (0, max_frames - n1 // 160)),
                 F.pad(self.positive_transform2(crop2), (0, max_frames - n2 // 160))]
        lengths = [n1 // 160 + 1, n2 // 160 + 1]
        return crops, lengths
