"""Frame masks: audiossl/methods/atstframe/random_mask.py:5-36.

``get_mask`` wraps fairseq's ``compute_mask_indices`` in the reference; fairseq is not available here and not
vendored by the reference, so the "static" block mask is restated from the published fairseq 0.12 algorithm as the
reference calls it (bsz=1, no padding mask, overlap allowed).  PARITY UNPINNED for this function:
the model treats masks as inputs, fixtures feed them explicitly."""
import numpy as np
import torch
from torch.nn import functional as F


def get_mask(batch_size, num_patches, mask_ratio, padding_mask=None, no_overlap=True, min_length=5, type="static",
             other=0):
    if type != "static" or no_overlap or padding_mask is not None:
        raise NotImplementedError("only the static, overlapping block mask of the ATST-Frame recipe is restated")
    masks = np.zeros((batch_size, num_patches), dtype=bool)
    for b in range(batch_size):
        num_mask = max(2, int(mask_ratio * num_patches / float(min_length) + np.random.rand()))
        if num_patches - min_length <= num_mask:
            min_len = num_patches - num_mask - 1
        else:
            min_len = min_length
        starts = np.random.choice(num_patches - min_len, num_mask, replace=False)
        for s in starts:
            masks[b, s:s + min_length] = True
    return torch.from_numpy(masks)


def get_mask_one(num_patches, available_patches, mask_ratio):
    mask_index_ = (torch.randperm(available_patches) < available_patches * mask_ratio)
    return F.pad(mask_index_, (0, num_patches - available_patches), value=1)


def get_mask_variable_length(batch_size, num_patches, available_patches, mask_ratio):
    avail = available_patches.to("cpu")
    return torch.cat([get_mask_one(num_patches, avail[i], mask_ratio).unsqueeze(0) for i in range(batch_size)])


def get_mask_batch(batch_size, num_patches, mask_ratio):
    return torch.cat([(torch.randperm(num_patches) < num_patches * mask_ratio).unsqueeze(0)
                      for _ in range(batch_size)])
