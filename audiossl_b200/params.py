"""Flat fp32 parameter / gradient storage for one network (student or teacher).

All parameters of a network live in ONE contiguous device buffer so that
  * the DDP gradient exchange is a single NCCL all-reduce,
  * the EMA teacher update and the AdamW step are single multi-tensor kernels (K17, K18),
  * the TF32 compute copy of the GEMM weights is one rounding pass.
Segments: [regularised (>=2-D, non-bias) of encoder+projector | non-regularised of encoder+projector |
           regularised of predictor | non-regularised of predictor]   (utils/common.py:41-68 grouping).
The nn.Parameters of the module tree are re-pointed at views of this buffer, so state_dict(),
load_state_dict() and checkpoints keep the reference's key layout.
"""
import torch

ALIGN = 64  # elements (256 B): keeps every tensor TMA / float4 friendly


def _is_regularized(name, p):
    return not (name.endswith(".bias") or p.ndim == 1)


class FlatParams:
    def __init__(self, named_params, device, ema_prefixes=("encoder.", "projector.")):
        """named_params: list of (name, nn.Parameter) in module order."""
        groups = [[], [], [], []]
        for name, p in named_params:
            in_ema = name.startswith(ema_prefixes)
            reg = _is_regularized(name, p)
            groups[(0 if in_ema else 2) + (0 if reg else 1)].append((name, p))
        self.offsets, self.shapes, self.order = {}, {}, []
        self.seg_bounds = []
        off = 0
        for g in groups:
            start = off
            for name, p in g:
                self.offsets[name] = off
                self.shapes[name] = tuple(p.shape)
                self.order.append(name)
                off += (p.numel() + ALIGN - 1) // ALIGN * ALIGN
            self.seg_bounds.append((start, off))
        self.total = off
        self.ema_count = self.seg_bounds[1][1]  # encoder+projector prefix of the buffer
        self.data = torch.zeros(self.total, device=device, dtype=torch.float32)
        self.grad = torch.zeros(self.total, device=device, dtype=torch.float32)
        self.compute = torch.zeros(self.total, device=device, dtype=torch.float32)  # tf32-rounded copy
        self.params = dict(named_params)
        with torch.no_grad():
            for name, p in named_params:
                v = self.view(self.data, name)
                v.copy_(p.data.to(device))
                p.data = v
        self._ptrs = {name: p.data_ptr() for name, p in named_params}

    def view(self, buf, name):
        off = self.offsets[name]
        shape = self.shapes[name]
        n = 1
        for s in shape:
            n *= s
        return buf[off:off + n].view(shape)

    def p(self, name):
        return self.view(self.data, name)

    def g(self, name):
        return self.view(self.grad, name)

    def c(self, name):
        return self.view(self.compute, name)

    def is_current(self):
        """False if someone (e.g. module.cuda()/load_state_dict with assign) re-allocated a parameter."""
        return all(self.params[n].data_ptr() == ptr for n, ptr in self._ptrs.items())

    def attach_grads(self):
        for name, p in self.params.items():
            if p.requires_grad:
                p.grad = self.g(name)

    def wd_segments(self):
        """[(start, end, regularised?)] for the fused optimizer."""
        (a0, a1), (b0, b1), (c0, c1), (d0, d1) = self.seg_bounds
        return [(a0, a1, True), (b0, b1, False), (c0, c1, True), (d0, d1, False)]
