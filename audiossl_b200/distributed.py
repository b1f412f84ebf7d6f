"""torch.distributed plumbing for the data-parallel step.

One process per GPU; NCCL over NVLink/NVSwitch on the GPU box, gloo for the CPU tests of this logic.
Collectives on the data path:
  C1  one all-reduce (AVG) of the flat student gradient buffer,
  C2  SyncBatchNorm statistics: all-gather of per-rank (mean, M2, n) in forward, all-reduce of
      (sum dy, sum dy*xhat) in backward,
  C3  compute_var: one all-reduce of the packed [1+4*256] loss / std accumulator (logging only).
"""
import torch
import torch.distributed as dist


def world():
    if dist.is_available() and dist.is_initialized():
        return dist.get_world_size()
    return 1


def combine_bn_stats(means, m2s, counts):
    """Chan et al. parallel variance: per-rank (mean [G,C], M2 [G,C], n [G]) -> global (mean, M2, n)."""
    n = counts.sum()
    w = (counts / n).view(-1, 1)
    mean = (means * w).sum(0)
    m2 = m2s.sum(0) + (counts.view(-1, 1) * (means - mean) ** 2).sum(0)
    return mean, m2, n


def bn_stats_sync(mean, m2, n, equal_counts=True):
    """SyncBatchNorm forward statistics (reference: sync_batchnorm=True, methods/atst/train.py:22).
    equal_counts: every rank contributes the same number of rows (true for ATST-clip) so the global
    count is known on the host without a device read; the frame model passes False."""
    G = world()
    if G == 1:
        return mean, m2, n
    C = mean.numel()
    packed = torch.cat([mean, m2, torch.full((1,), float(n), device=mean.device)])
    gathered = [torch.empty_like(packed) for _ in range(G)]
    dist.all_gather(gathered, packed)
    g = torch.stack(gathered)
    mean_g, m2_g, n_g = combine_bn_stats(g[:, :C], g[:, C:2 * C], g[:, 2 * C])
    n_total = float(n) * G if equal_counts else float(n_g.item())
    return mean_g.contiguous(), m2_g.contiguous(), n_total


def bn_sums_sync(s1, s2):
    G = world()
    if G == 1:
        return s1, s2
    packed = torch.cat([s1, s2])
    dist.all_reduce(packed)
    C = s1.numel()
    return packed[:C].contiguous(), packed[C:].contiguous()


def allreduce_avg_(flat):
    G = world()
    if G == 1:
        return flat
    if dist.get_backend() == "nccl":
        dist.all_reduce(flat, op=dist.ReduceOp.AVG)
    else:
        dist.all_reduce(flat)
        flat.mul_(1.0 / G)
    return flat


def allreduce_sum_(t):
    if world() > 1:
        dist.all_reduce(t)
    return t
