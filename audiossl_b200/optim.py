"""Fused AdamW with the semantics of transformers-4.x ``AdamW`` (the optimizer of
audiossl/methods/atst/model.py:44-48): betas (0.9, 0.999), eps 1e-6 added to sqrt(v) before bias
correction, ``correct_bias=True``, decoupled weight decay applied after the Adam update.

It is a regular ``torch.optim.Optimizer`` (param_groups with ``lr`` / ``weight_decay`` that the
Lightning module's ``schedule()`` overwrites every step), but ``step()`` is a handful of launches of one
multi-tensor kernel over the flat parameter / gradient / moment buffers instead of a Python loop over
~150 tensors.
"""
import torch

from . import ops


class FusedHFAdamW(torch.optim.Optimizer):
    def __init__(self, params, flat=None, lr=1e-3, betas=(0.9, 0.999), eps=1e-6, weight_decay=0.0,
                 correct_bias=True):
        if not correct_bias:
            raise NotImplementedError("correct_bias=False is not used by the reference")
        defaults = dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay, correct_bias=correct_bias)
        super().__init__(params, defaults)
        self.flat_provider = flat  # callable returning the FlatParams of the student
        self._m = self._v = None
        self._step = 0

    @torch.no_grad()
    def step(self, closure=None):
        loss = closure() if closure is not None else None
        fp = self.flat_provider()
        if self._m is None or self._m.numel() != fp.total:
            self._m = torch.zeros_like(fp.data)
            self._v = torch.zeros_like(fp.data)
        self._step += 1
        g_reg, g_noreg = self.param_groups[0], self.param_groups[1]
        for (a, b, reg) in fp.wd_segments():
            if b <= a:
                continue
            grp = g_reg if reg else g_noreg
            ops.adamw_step(fp.data[a:b], fp.grad[a:b], self._m[a:b], self._v[a:b], self._step, grp["lr"],
                           grp["weight_decay"], grp["betas"][0], grp["betas"][1], grp["eps"])
        return loss

    def zero_grad(self, set_to_none=True):
        # gradients live in one flat buffer that the backward pass overwrites; nothing to do per tensor
        for group in self.param_groups:
            for p in group["params"]:
                if set_to_none:
                    p.grad = None

    def state_dict(self):
        sd = super().state_dict()
        sd["flat_state"] = {"step": self._step, "m": self._m, "v": self._v}
        return sd

    def load_state_dict(self, sd):
        flat = sd.pop("flat_state", None)
        super().load_state_dict(sd)
        if flat is not None:
            self._step, self._m, self._v = flat["step"], flat["m"], flat["v"]
