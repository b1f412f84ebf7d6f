"""Flat fp32 parameter / gradient storage for one network (student or teacher).

All parameters of a network live in ONE contiguous device buffer so that
  * the DDP gradient exchange is a single NCCL all-reduce,
  * the EMA teacher update and the AdamW step are single multi-tensor kernels (K17, K18),
  * the TF32 compute copy of the GEMM weights is one rounding pass.
Segments: [regularised (>=2-D, non-bias) of encoder+projector | non-regularised of encoder+projector |
           regularised of predictor | non-regularised of predictor]   (utils/common.py:41-68 grouping).
Tensors that never receive a gradient ("frozen": requires_grad=False, or ATST-clip's encoder.mask_embed, which the
clip forward never reads - the reason the reference recipe needs find_unused_parameters, methods/atst/train.py:19)
sit at the head of their segment and are excluded from the optimizer ranges and from the gradient exchange:
transformers' AdamW skips a parameter whose .grad is None (no update, no weight decay), so must this.
The nn.Parameters of the module tree are re-pointed at views of this buffer, so state_dict(),
load_state_dict() and checkpoints keep the reference's key layout.
"""
import torch

ALIGN = 64  # elements (256 B): keeps every tensor TMA / float4 friendly


def _is_regularized(name, p):
    return not (name.endswith(".bias") or p.ndim == 1)


class FlatParams:
    def __init__(self, named_params, device, ema_prefixes=("encoder.", "projector."), frozen=()):
        """named_params: list of (name, nn.Parameter) in module order; frozen: names that never get a gradient."""
        self.frozen = frozenset(frozen)
        groups = [[], [], [], []]
        for name, p in named_params:
            in_ema = name.startswith(ema_prefixes)
            reg = _is_regularized(name, p)
            groups[(0 if in_ema else 2) + (0 if reg else 1)].append((name, p))
        self.offsets, self.shapes, self.order = {}, {}, []
        self.seg_bounds, self.seg_trainable = [], []
        off = 0
        for g in groups:
            start = off
            g = [x for x in g if x[0] in self.frozen] + [x for x in g if x[0] not in self.frozen]
            first_trainable = None
            for name, p in g:
                if first_trainable is None and name not in self.frozen:
                    first_trainable = off
                self.offsets[name] = off
                self.shapes[name] = tuple(p.shape)
                self.order.append(name)
                off += (p.numel() + ALIGN - 1) // ALIGN * ALIGN
            self.seg_bounds.append((start, off))
            self.seg_trainable.append(off if first_trainable is None else first_trainable)
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
        for name, p in named_params:
            p._atst_flat = (self, name)  # lets a sub-module (an encoder used stand-alone) find its storage again
        # set by the backward pass, cleared by the optimizer: a second backward before the gradients were consumed
        # would overwrite them (Lightning's accumulate_grad_batches > 1), which must not happen silently
        self._ranges = {}
        self._pv, self._gv, self._cv = {}, {}, {}  # views of the three flat buffers by name (built on first use)
        self.grads_pending = False
        self.has_optimizer = False

    def view(self, buf, name):
        off = self.offsets[name]
        shape = self.shapes[name]
        n = 1
        for s in shape:
            n *= s
        return buf[off:off + n].view(shape)

    def _cached(self, cache, buf, name):
        v = cache.get(name)
        if v is None:
            v = cache[name] = self.view(buf, name)
        return v

    def p(self, name):
        return self._cached(self._pv, self.data, name)

    def g(self, name):
        return self._cached(self._gv, self.grad, name)

    def c(self, name):
        return self._cached(self._cv, self.compute, name)

    def is_current(self):
        """False if someone (e.g. module.cuda()/load_state_dict with assign) re-allocated a parameter."""
        return all(self.params[n].data_ptr() == ptr for n, ptr in self._ptrs.items())

    def attach_grads(self):
        for name, p in self.params.items():
            if p.requires_grad:
                p.grad = None if name in self.frozen else self.g(name)

    def wd_segments(self):
        """[(start, end, regularised?)] for the fused optimizer: the trainable tail of every segment."""
        return [(t, e, i % 2 == 0) for i, ((_, e), t) in enumerate(zip(self.seg_bounds, self.seg_trainable)) if e > t]

    def matrix_range(self, prefix):
        """[lo, hi) of the flat buffers covered by the regularised (>= 2-D) tensors whose names start with ``prefix``
        (one transformer block, the projector, ...): contiguous because segments keep module order."""
        cached = self._ranges.get(prefix)
        if cached is not None:
            return cached
        self._ranges[prefix] = rng = self._matrix_range(prefix)
        return rng

    def _matrix_range(self, prefix):
        a, b = self.seg_bounds[0] if prefix.startswith(("encoder.", "projector.")) else self.seg_bounds[2]
        names = [n for n in self.order if n.startswith(prefix) and a <= self.offsets[n] < b and n not in self.frozen]
        if not names:
            return 0, 0
        lo = min(self.offsets[n] for n in names)
        hi = max(self.offsets[n] + (self.params[n].numel() + ALIGN - 1) // ALIGN * ALIGN for n in names)
        assert hi - lo == sum((self.params[n].numel() + ALIGN - 1) // ALIGN * ALIGN for n in names), prefix
        return lo, hi

    def exchange_start(self):
        """first element of the flat gradient that the data-parallel exchange carries (frozen head excluded)."""
        return self.seg_trainable[0] if self.seg_trainable[0] < self.seg_bounds[0][1] else 0

    def exchanged_grad(self):
        """the slice of the flat gradient that the data-parallel all-reduce carries (frozen head excluded)."""
        return self.grad[self.seg_trainable[0]:] if self.seg_trainable[0] < self.seg_bounds[0][1] else self.grad
