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


def bn_stats_sync(mean, m2, n, n_total=None):
    """SyncBatchNorm forward statistics (reference: sync_batchnorm=True, methods/atst/train.py:22).
    n_total: the global row count when the ranks contribute different numbers of rows (ATST-Frame: masked frames;
    the caller all-reduces the count once per step) - the default is n * world (ATST-clip).  No host read here."""
    G = world()
    if G == 1:
        return mean, m2, n
    (mean_g, m2_g), = bn_stats_sync_many([(mean, m2, n)])
    return mean_g, m2_g, (float(n) * G if n_total is None else float(n_total))


def bn_stats_sync_many(stats):
    """several BatchNorm layers in ONE all-gather: stats = [(mean [C], M2 [C], n)] -> [(global mean, global M2)]."""
    G = world()
    C = stats[0][0].numel()
    dev = stats[0][0].device
    packed = torch.cat([t for mean, m2, n in stats for t in (mean, m2, torch.full((1,), float(n), device=dev))])
    gathered = torch.empty((G, packed.numel()), device=dev, dtype=packed.dtype)
    dist.all_gather_into_tensor(gathered, packed) if dist.get_backend() == "nccl" else \
        gathered.copy_(torch.stack(_all_gather_list(packed, G)))
    out = []
    w = 2 * C + 1
    for k in range(len(stats)):
        g = gathered[:, k * w:(k + 1) * w]
        mean_g, m2_g, _ = combine_bn_stats(g[:, :C], g[:, C:2 * C], g[:, 2 * C])
        out.append((mean_g.contiguous(), m2_g.contiguous()))
    return out


def _all_gather_list(t, G):
    parts = [torch.empty_like(t) for _ in range(G)]
    dist.all_gather(parts, t)
    return parts


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


class GradExchange:
    """The data-parallel gradient exchange, overlapped with the backward pass: ranges of the flat gradient buffer are
    all-reduced (AVG) on a side stream as soon as the kernels that accumulate into them have been enqueued
    (``submit``), the complement goes out at the end (``finish``), and only then does the compute stream wait.
    The flat buffer is laid out in module order, so a transformer block's four weight matrices are one contiguous
    range that is complete when the backward pass leaves the block."""

    def __init__(self, grad, lo, device):
        self.grad, self.lo, self.hi = grad, lo, grad.numel()
        self.stream = torch.cuda.Stream(device=device) if device.type == "cuda" else None
        self.pending, self.done = [], []

    def _issue(self, a, b):
        chunk = self.grad[a:b]
        if dist.get_backend() == "nccl":
            self.pending.append((dist.all_reduce(chunk, op=dist.ReduceOp.AVG, async_op=True), None))
        else:
            self.pending.append((dist.all_reduce(chunk, async_op=True), chunk))

    def submit(self, a, b):
        """[a, b) of the flat buffer is final once the work enqueued so far on the current stream has run."""
        if world() == 1 or b <= a:
            return
        a = max(a, self.lo)
        self.done.append((a, b))
        if self.stream is None:
            return self._issue(a, b)
        self.stream.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(self.stream):
            self._issue(a, b)

    def finish(self):
        """exchange whatever was not submitted, then make the current stream wait for all of it."""
        G = world()
        if G == 1:
            return
        gaps, pos = [], self.lo
        for a, b in sorted(self.done) + [(self.hi, self.hi)]:
            if a > pos:
                gaps.append((pos, a))
            pos = max(pos, b)
        for a, b in gaps:
            self.submit(a, b)
        ctx = torch.cuda.stream(self.stream) if self.stream is not None else _Null()
        with ctx:
            for work, chunk in self.pending:
                work.wait()
                if chunk is not None:
                    chunk.mul_(1.0 / G)
        if self.stream is not None:
            torch.cuda.current_stream().wait_stream(self.stream)
        self.pending, self.done = [], []


class _Null:
    def __enter__(self):
        return self

    def __exit__(self, *a):
        return False


def allreduce_sum_(t):
    if world() > 1:
        dist.all_reduce(t)
    return t
