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
    bank entry using alpha = ratio * U(0,1); the un-mixed batch is then pushed into the bank."""

    def __init__(self, ratio=0.4, n_memory=2000, rng=None):
        self.ratio, self.n = ratio, n_memory
        self.bank, self.size, self.head = None, 0, 0
        self.rng = rng or np.random

    def __call__(self, x, alpha=None, idx=None):
        """x [B,1,64,T] cuda.  alpha / idx override the random draws (tests)."""
        B = x.shape[0]
        x = x.contiguous()
        if self.bank is None or self.bank.shape[1:] != x.shape[1:]:
            self.bank = torch.empty((self.n,) + tuple(x.shape[1:]), device=x.device, dtype=torch.float32)
            self.size = self.head = 0
        if alpha is None:
            alpha = self.ratio * self.rng.random(B)
        if idx is None:
            idx = self.rng.randint(0, self.size, B) if self.size > 0 else np.full(B, -1)
        a = torch.as_tensor(np.asarray(alpha, np.float32), device=x.device)
        j = torch.as_tensor(np.asarray(idx, np.int32), device=x.device)
        out = ops.mixup_fwd(x, self.bank, j, a, torch.empty_like(x))
        # FIFO push of the un-mixed inputs
        pos = (self.head + torch.arange(B, device=x.device)) % self.n
        self.bank.index_copy_(0, pos, x)
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
