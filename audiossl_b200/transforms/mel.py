"""Fused log-mel front-end as a transform callable.

Drop-in for the reference's ``mel_feature`` Compose (audiossl/methods/atst/transform.py:14-29):
MelSpectrogram(16000, f_min=60, f_max=7800, hop 160, win 1024|640, n_fft 1024, 64 mels) ->
AmplitudeToDB("power", top_db=80) -> MinMax(-79.6482, 50.6842), executed by one CUDA kernel pair on the
waveform's device.  Accepts [1,n] (one clip, as the DataLoader transform does) or batched [B,1,n] /
[B,n] (as audiossl/methods/atstframe/embedding.py:57-60 does); the top_db clamp is per clip either way.
"""
import torch

from .. import ops


class LogMelSpectrogram:
    def __init__(self, sr=16000, n_mels=64, win_length=1024, hop_length=160, n_fft=1024, f_min=60, f_max=7800,
                 top_db=80, min=-79.6482, max=50.6842):
        if (sr, n_mels, hop_length, n_fft, f_min, f_max, top_db) != (16000, 64, 160, 1024, 60, 7800, 80):
            raise NotImplementedError("the fused kernel implements the ATST front-end constants only")
        if (min, max) != (-79.6482, 50.6842):
            raise NotImplementedError("MinMax constants are baked into the kernel")
        self.win_length = win_length

    def __call__(self, wav, clip_start=None, clip_len=None):
        """clip_start (int64 [B], cuda) / clip_len: transform the window [start, start + clip_len) of every clip
        (RandomCrop fused into the kernel's addressing; no cropped copy is made)."""
        if not wav.is_cuda:
            raise RuntimeError("LogMelSpectrogram runs on the GPU only (no CPU fallback); move the waveform to cuda")
        return ops.mel_forward(wav.float().contiguous(), win_length=self.win_length, clip_start=clip_start,
                               clip_len=clip_len)

    def __repr__(self):
        return "LogMelSpectrogram(win_length=%d)" % self.win_length
