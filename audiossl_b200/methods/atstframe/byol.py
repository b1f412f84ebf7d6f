"""Frame-level heads / loss shells: audiossl/methods/atstframe/byol.py:57-138."""
import torch.nn as nn

from ... import ops
from ...models.atst.byol import build_mlp


class ByolLoss(nn.Module):
    """symmetric frame loss over the two views' masked frames + logged std statistics (byol.py:57-84)."""

    def __init__(self, symmetric):
        super().__init__()
        if not symmetric:
            raise NotImplementedError("symmetric=False is not used by the ATST-Frame recipes")
        self.symmetric = symmetric
        self.ncrops = 2

    def forward(self, student, teacher):
        from ...distributed import allreduce_sum_
        import torch
        R = student.shape[0]
        _, acc = ops.byol_loss(student.contiguous(), teacher.contiguous(), 2, R // 2)
        n = torch.tensor([float(R)], device=student.device)
        allreduce_sum_(acc[1:])
        allreduce_sum_(n)
        out = ops.byol_finalize(acc, n.item(), n.item(), 2, R // 2)
        return out[0], out[1], out[2]


class MultiCropWrapper(nn.Module):
    def __init__(self, encoder, embed_dim, projector="mlp", predictor=True):
        super().__init__()
        if projector != "mlp":
            raise NotImplementedError('projector="linear"/None belong to the data2vec variant (avg_blocks>0)')
        self.encoder = encoder
        self.projector = build_mlp(2, embed_dim, 4096, 256, last_bn=False)
        self.predictor = build_mlp(2, 256, 4096, 256, last_bn=False) if predictor else nn.Identity()
