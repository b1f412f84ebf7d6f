"""Heads, loss and multi-crop wrapper shells: audiossl/models/atst/byol.py:6-121.

The modules own the parameters / BN buffers under the reference's keys (projector.0.weight,
projector.1.{weight,bias,running_mean,running_var,num_batches_tracked}, projector.3.weight, predictor.*);
the arithmetic runs in audiossl_b200.models.atst.atst._Runtime on the CUDA engine.
"""
import torch
import torch.nn as nn

from ... import ops


def build_mlp(num_layers, input_dim, mlp_dim, output_dim, last_bn=True):
    if num_layers != 2 or last_bn:
        raise NotImplementedError("ATST uses build_mlp(2, in, 4096, 256, last_bn=False) only (byol.py:97-99)")
    return nn.Sequential(nn.Linear(input_dim, mlp_dim, bias=False), nn.BatchNorm1d(mlp_dim), nn.ReLU(inplace=True),
                         nn.Linear(mlp_dim, output_dim, bias=False))


class ByolLoss(nn.Module):
    """BYOL loss + the two logged std statistics (byol.py:57-78); forward-only convenience entry.
    (The fused training step in ATST computes loss and d loss / d student in one kernel.)"""

    def __init__(self, ncrops):
        super().__init__()
        self.ncrops = ncrops

    def forward(self, student, teacher):
        B = teacher.shape[0] // 2
        _, acc = ops.byol_loss(student.contiguous(), teacher.contiguous(), self.ncrops, B)
        from ...distributed import allreduce_sum_, world
        allreduce_sum_(acc[1:])
        out = ops.byol_finalize(acc, student.shape[0] * world(), teacher.shape[0] * world(), self.ncrops, B)
        return out[0], out[1], out[2]


class MultiCropWrapper(nn.Module):
    def __init__(self, encoder, embed_dim, predictor=True):
        super().__init__()
        self.encoder = encoder
        self.projector = build_mlp(2, embed_dim, 4096, 256, last_bn=False)
        if predictor:
            self.predictor = build_mlp(2, 256, 4096, 256, last_bn=False)
        else:
            self.predictor = nn.Identity()

    @staticmethod
    def group_crops(x):
        """consecutive crops of equal width share one encoder call (byol.py:107-116)."""
        widths = [inp.shape[-1] for inp in x]
        groups, start = [], 0
        for i in range(1, len(x) + 1):
            if i == len(x) or widths[i] != widths[start]:
                groups.append((start, i))
                start = i
        return groups
