"""FrameATSTTrainTransform - audiossl/methods/atstframe/transform.py:14-101: one crop, two views (teacher view
un-augmented when aug_tea=False), one block mask shared by both views.  Mel runs on the fused CUDA kernel."""
from torch.nn import functional as F

from ...models.atst.audio_transformer import get_num_patches
from ...transforms.byol_a import Mixup, RandomResizeCrop
from ...transforms.common import Identity, RandomCrop
from ...transforms.mel import LogMelSpectrogram
from ..atst.transform import _Compose
from . import random_mask


class FrameATSTTrainTransform:
    def __init__(self, sr=16000, win_length=1024, aug_tea=True, aug_stu=True, mix_up=True, freq_wrap=True,
                 mask_ratio=0.75, mask_nooverlap=False, min_mask_len=2, mask_len=5, mask_type="random",
                 anchor_len=6., patch_h=64, patch_w=4, n_mels=64, **kwargs):
        self.anchor_len = self.max_positive_len = anchor_len
        self.mask_ratio, self.mask_type = mask_ratio, mask_type
        self.patch_h, self.patch_w, self.mask_len, self.n_mels = patch_h, patch_w, mask_len, n_mels
        self.mask_nooverlap, self.min_mask_len = mask_nooverlap, min_mask_len
        self.mel_feature = LogMelSpectrogram(sr, n_mels=n_mels, win_length=win_length)
        self.positivecrop = _Compose([RandomCrop(16000 * 6), self.mel_feature])

        def aug():
            return _Compose([Mixup() if mix_up else Identity(),
                             RandomResizeCrop((1, 1.0), time_scale=(1.0, 1.0)) if freq_wrap else Identity()])
        self.positive_transform1 = aug() if aug_tea else Identity()
        self.positive_transform2 = aug() if aug_stu else Identity()

    def __call__(self, input):
        n = int(self.anchor_len * 16000)
        self.positivecrop.transforms[0].size = n
        crop = self.positivecrop(input)
        num_patches = get_num_patches(self.n_mels, n // 160 + 1, self.patch_h, self.patch_w)
        if self.mask_type == "random":
            mask = random_mask.get_mask_one(num_patches, num_patches, self.mask_ratio)
        else:
            mask = random_mask.get_mask(1, num_patches, self.mask_ratio, no_overlap=self.mask_nooverlap,
                                        min_length=self.mask_len,
                                        type="static" if self.mask_type == "block" else "uniform",
                                        other=self.min_mask_len).squeeze(0)
        pad = int((self.max_positive_len * 16000) // 160 - n // 160)
        crops = [F.pad(self.positive_transform1(crop), (0, pad)), F.pad(self.positive_transform2(crop), (0, pad))]
        return crops, [n // 160 + 1] * 2, [mask] * 2


class BatchedFrameATSTTrainTransform:
    """The ATST-Frame recipe for a whole batch on the GPU: ``wav [B,1,n]`` -> ``(crops, lengths, masks)`` as
    ``FrameATSTLightningModule.training_step`` takes them (two views of ONE crop per clip, frequency-only
    resize-crop, one block mask per clip shared by both views)."""

    def __init__(self, sr=16000, win_length=1024, aug_tea=True, aug_stu=True, mix_up=True, freq_wrap=True,
                 mask_ratio=0.75, mask_nooverlap=False, min_mask_len=2, mask_len=5, mask_type="random",
                 anchor_len=6., patch_h=64, patch_w=4, n_mels=64, rng=None, **kwargs):
        import numpy as np
        from ...transforms.batched import BatchedMixup, BatchedRandomResizeCrop
        self.rng = rng or np.random
        self.anchor_len = anchor_len
        self.mask_ratio, self.mask_type, self.mask_len, self.mask_nooverlap = mask_ratio, mask_type, mask_len, mask_nooverlap
        self.patch_h, self.patch_w, self.n_mels = patch_h, patch_w, n_mels
        self.mel_feature = LogMelSpectrogram(sr, n_mels=n_mels, win_length=win_length)
        frames = int(anchor_len * 16000) // 160 + 1
        self.aug = [aug_tea, aug_stu]
        self.mixup = [BatchedMixup(max_frames=frames, rng=self.rng) if mix_up else None for _ in range(2)]
        self.rrc = [BatchedRandomResizeCrop((1, 1.0), time_scale=(1.0, 1.0), rng=self.rng) if freq_wrap else None
                    for _ in range(2)]

    def __call__(self, wav):
        import torch
        if not wav.is_cuda or wav.dim() != 3:
            raise RuntimeError("BatchedFrameATSTTrainTransform takes a [B,1,n] waveform batch on the GPU")
        B, _, n = wav.shape
        size = int(self.anchor_len * 16000)
        if n < size:
            wav = F.pad(wav, (0, size - n))
            n = size
        start = torch.as_tensor(self.rng.randint(0, n - size + 1, B), dtype=torch.int64, device=wav.device)
        mel = self.mel_feature(wav, clip_start=start, clip_len=size)
        P = get_num_patches(self.n_mels, size // 160 + 1, self.patch_h, self.patch_w)
        if self.mask_type == "random":
            mask = random_mask.get_mask_batch(B, P, self.mask_ratio)
        else:
            mask = random_mask.get_mask(B, P, self.mask_ratio, no_overlap=self.mask_nooverlap,
                                        min_length=self.mask_len,
                                        type="static" if self.mask_type == "block" else "uniform")
        mask = mask.to(wav.device)
        crops = []
        for v in range(2):
            x = mel
            if self.aug[v]:
                if self.mixup[v] is not None:
                    x = self.mixup[v](x)
                if self.rrc[v] is not None:
                    x = self.rrc[v](x)
            crops.append(x)
        lengths = torch.full((B,), size // 160 + 1, dtype=torch.int64, device=wav.device)
        return crops, [lengths, lengths], [mask, mask]
