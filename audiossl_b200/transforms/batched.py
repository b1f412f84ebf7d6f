"""Device-batched versions of the recipe's log-mel augmentations.

The reference applies ``Mixup`` and ``RandomResizeCrop`` per sample on CPU DataLoader workers
(audiossl/methods/atst/transform.py:35-46).  Once the mel front-end runs on the GPU these are the remaining
per-sample CPU stage, so they get batched CUDA kernels here.  Parity is distributional for the random draws
(different RNG streams; the memory bank is per process instead of per DataLoader worker) and exact for the
arithmetic given the draws (tests feed the same parameters to torch's bicubic resize / the log-mixup formula)."""
import numpy as np
import torch

from .. import ops


class BatchedMixup:
    """log-mixup-exp against a device FIFO of past inputs (byol_a.py:85-115): each clip is mixed with a random
    bank entry using alpha = ratio * U(0,1); the un-mixed batch is then pushed into the bank.  The bank is one
    [n_memory, Hm, max_frames] device tensor with a per-entry frame count, so batches of different widths mix the way
    ``log_mixup_exp`` does (random alignment of the shorter clip, byol_a.py:61-82)."""

    def __init__(self, ratio=0.4, n_memory=2000, max_frames=None, rng=None):
        self.ratio, self.n, self.max_frames = ratio, n_memory, max_frames
        self.bank, self.size, self.head = None, 0, 0
        self.lens = np.zeros(n_memory, np.int32)
        self.rng = rng or np.random

    def _ensure_bank(self, x):
        Hm, T = x.shape[-2], x.shape[-1]
        width = max(T, self.max_frames or 0)
        if self.bank is None or self.bank.shape[1] != Hm or self.bank.shape[2] < T or self.bank.device != x.device:
            old = self.bank
            self.bank = torch.zeros((self.n, Hm, width), device=x.device, dtype=torch.float32)
            if old is not None and old.shape[1] == Hm and old.device == x.device:
                self.bank[:, :, :old.shape[2]] = old  # a wider batch arrived: keep the entries
            else:
                self.size = self.head = 0

    def __call__(self, x, alpha=None, idx=None, start=None):
        """x [B,1,Hm,T] cuda.  alpha / idx / start override the random draws (tests)."""
        B, T = x.shape[0], x.shape[-1]
        x = x.contiguous()
        self._ensure_bank(x)
        if alpha is None:
            alpha = self.ratio * self.rng.random(B)
        if idx is None:
            idx = self.rng.randint(0, self.size, B) if self.size > 0 else np.full(B, -1)
        idx = np.asarray(idx, np.int32)
        zlen = np.where(idx >= 0, self.lens[np.maximum(idx, 0)], T).astype(np.int32)
        if start is None:
            span = np.abs(zlen - T)
            start = np.where(span > 0, (self.rng.random(B) * span).astype(np.int32), 0)
        dev = x.device
        out = ops.mixup_fwd(x, self.bank, torch.as_tensor(idx, device=dev),
                            torch.as_tensor(np.asarray(alpha, np.float32), device=dev), torch.empty_like(x),
                            zlen=torch.as_tensor(zlen, device=dev),
                            start=torch.as_tensor(np.asarray(start, np.int32), device=dev))
        # FIFO push of the un-mixed inputs
        pos = (self.head + np.arange(B)) % self.n
        self.bank[torch.as_tensor(pos, device=dev), :, :T] = x.reshape(B, x.shape[-2], T)
        self.lens[pos] = T
        self.head = (self.head + B) % self.n
        self.size = min(self.n, self.size + B)
        return out


class BatchedRandomResizeCrop:
    """byol_a.py:7-49 for a batch: per-clip crop of the zero "virtual canvas", bicubic align_corners resize back."""

    def __init__(self, virtual_crop_scale=(1.0, 1.5), freq_scale=(0.6, 1.5), time_scale=(0.6, 1.5), rng=None):
        assert time_scale[1] >= 1.0 and freq_scale[1] >= 1.0
        self.virtual_crop_scale, self.freq_scale, self.time_scale = virtual_crop_scale, freq_scale, time_scale
        self.rng = rng or np.random

    def get_params(self, B, canvas, size):
        canvas_h, canvas_w = canvas
        src_h, src_w = size
        rect = np.zeros((B, 4), np.int32)
        for b in range(B):
            h = int(np.clip(int(self.rng.uniform(*self.freq_scale) * src_h), 1, canvas_h))
            w = int(np.clip(int(self.rng.uniform(*self.time_scale) * src_w), 1, canvas_w))
            i = self.rng.randint(0, canvas_h - h + 1) if canvas_h > h else 0
            j = self.rng.randint(0, canvas_w - w + 1) if canvas_w > w else 0
            rect[b] = (i, j, h, w)
        return rect

    def __call__(self, lms, rect=None):
        """lms [B,1,64,T] cuda -> same shape.  rect [B,4] = (i, j, h, w) overrides the random draws (tests)."""
        B, _, Hm, T = lms.shape
        ch, cw = int(Hm * self.virtual_crop_scale[0]), int(T * self.virtual_crop_scale[1])
        if rect is None:
            rect = self.get_params(B, (ch, cw), (Hm, T))
        r = torch.as_tensor(np.asarray(rect, np.int32), device=lms.device).contiguous()
        return ops.resize_crop_fwd(lms.contiguous(), r, torch.empty_like(lms), ch, cw)
