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


class BatchedATSTTrainTransform:
    """The same recipe for a whole batch resident on the GPU: ``wav [B,1,n]`` (cuda) ->
    ``([crop1, crop2], [len1, len2])`` with crops ``[B,1,64,T_max]`` and lengths int64 ``[B]`` - the batch contract of
    ``ATSTLightningModule.training_step`` - in three kernel launches per view (fused mel with the random window in its
    addressing, log-mixup-exp against a device memory bank, bicubic resize-crop) instead of B Python calls on
    DataLoader workers.

    Differences from per-sample calls of ``ATSTTrainTransform`` (distributional, not arithmetic): one view length is
    drawn per batch and view (the recipe's default range is the single value 6 s), the Mixup memory bank is per
    process and view instead of per DataLoader worker, and the draws come from ``rng`` (a ``numpy.random.RandomState``;
    default: the global numpy generator)."""

    def __init__(self, sr=16000, mask_ratio=0.75, different_positive=True, anchor_len=(6., 6.),
                 positive_len=(6., 6.), virtual_crop=1.5, rng=None, augment=True):
        import numpy as np
        from ...transforms.batched import BatchedMixup, BatchedRandomResizeCrop
        self.np = np
        self.rng = rng or np.random
        self.different_positive = different_positive
        self.anchor_len, self.positive_len = anchor_len, positive_len
        self.max_positive_len = max(self.positive_len + self.anchor_len)
        self.max_frames = int((self.max_positive_len * 16000) // 160)
        self.mel_feature = LogMelSpectrogram(sr, win_length=1024)
        self.augment = augment
        self.mixup = [BatchedMixup(max_frames=self.max_frames + 1, rng=self.rng) for _ in range(2)]
        self.rrc = [BatchedRandomResizeCrop((1, virtual_crop), rng=self.rng) for _ in range(2)]

    def _crop_mel(self, wav, seconds):
        """random window of int(seconds * 16000) samples per clip (RandomCrop: zero-pad when the clip is shorter),
        then the fused mel."""
        import torch
        B, _, n = wav.shape
        size = int(seconds * 16000)
        if n < size:
            wav = F.pad(wav, (0, size - n))
            n = size
        start = torch.as_tensor(self.rng.randint(0, n - size + 1, B), dtype=torch.int64, device=wav.device)
        return self.mel_feature(wav, clip_start=start, clip_len=size), size

    def _view(self, mel, size, v):
        import torch
        if self.augment:
            mel = self.rrc[v](self.mixup[v](mel))
        crop = F.pad(mel, (0, self.max_frames - size // 160))
        return crop, torch.full((mel.shape[0],), size // 160 + 1, dtype=torch.int64, device=mel.device)

    def __call__(self, wav):
        if not wav.is_cuda or wav.dim() != 3:
            raise RuntimeError("BatchedATSTTrainTransform takes a [B,1,n] waveform batch on the GPU")
        anchor_len = self.rng.uniform(*self.anchor_len)
        mel1, n1 = self._crop_mel(wav, anchor_len)
        if self.different_positive:
            mel2, n2 = self._crop_mel(wav, self.rng.uniform(*self.positive_len))
        else:
            mel2, n2 = mel1, n1
        c1, l1 = self._view(mel1, n1, 0)
        c2, l2 = self._view(mel2, n2, 1)
        return [c1, c2], [l1, l2]
