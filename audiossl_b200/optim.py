"""Fused AdamW with the semantics of transformers-4.x ``AdamW`` (the optimizer of
audiossl/methods/atst/model.py:44-48): betas (0.9, 0.999), eps 1e-6 added to sqrt(v) before bias
correction, ``correct_bias=True``, decoupled weight decay applied after the Adam update, and - like its
``if p.grad is None: continue`` - no update and no decay for a parameter that never receives a gradient
(ATST-clip's ``encoder.mask_embed``, frozen tensors): those sit outside ``FlatParams.wd_segments()``.

It is a regular ``torch.optim.Optimizer`` (param_groups with ``lr`` / ``weight_decay`` that the
Lightning module's ``schedule()`` overwrites every step), but ``step()`` is a handful of launches of one
multi-tensor kernel over the flat parameter / gradient / moment buffers instead of a Python loop over
~150 tensors.

Checkpoints: ``state_dict()`` carries the per-parameter layout of transformers' AdamW (``state[i] = {step,
exp_avg, exp_avg_sq}``, views of the flat moment buffers), so a reference optimizer checkpoint loads here and
vice versa.
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
        # [4] cuda tensor {step_size, lr*wd} x {regularised, non-regularised group}: when set (graph.GraphedTrainStep),
        # the kernels read the step's scalars from it instead of taking them as launch arguments
        self.device_scalars = None

    def _moments(self, fp):
        if self._m is None or self._m.numel() != fp.total or self._m.device != fp.data.device:
            self._m = torch.zeros_like(fp.data)
            self._v = torch.zeros_like(fp.data)
        return self._m, self._v

    @torch.no_grad()
    def step(self, closure=None):
        loss = closure() if closure is not None else None
        fp = self.flat_provider()
        fp.has_optimizer, fp.grads_pending = True, False
        m, v = self._moments(fp)
        self._step += 1
        g_reg, g_noreg = self.param_groups[0], self.param_groups[1]
        for (a, b, reg) in fp.wd_segments():
            grp = g_reg if reg else g_noreg
            dyn = None if self.device_scalars is None else self.device_scalars[0 if reg else 2:2 if reg else 4]
            ops.adamw_step(fp.data[a:b], fp.grad[a:b], m[a:b], v[a:b], self._step, grp["lr"],
                           grp["weight_decay"], grp["betas"][0], grp["betas"][1], grp["eps"], dyn=dyn)
        return loss

    def step_scalars(self, step, lr, wd):
        """the four floats of ``device_scalars`` for optimizer step number ``step`` (1-based)."""
        out = []
        for grp, w in ((self.param_groups[0], wd), (self.param_groups[1], 0.0)):
            b1, b2 = grp["betas"]
            out += [lr * (1.0 - b2 ** step) ** 0.5 / (1.0 - b1 ** step), lr * w]
        return out

    def zero_grad(self, set_to_none=True):
        # gradients live in one flat buffer that the backward pass rebuilds; nothing to clear per tensor
        try:
            fp = self.flat_provider()
            fp.has_optimizer, fp.grads_pending = True, False
        except RuntimeError:  # no CUDA runtime yet (module still on the host)
            pass
        for group in self.param_groups:
            for p in group["params"]:
                if set_to_none:
                    p.grad = None

    # ------------------------------------------------------------------ checkpoints (transformers-AdamW layout)
    def _indexed_names(self, fp):
        """[(index in the param_groups numbering, flat-buffer name)] of the parameters this optimizer owns."""
        by_ptr = {p.data_ptr(): n for n, p in fp.params.items()}
        out, i = [], 0
        for group in self.param_groups:
            for p in group["params"]:
                out.append((i, by_ptr.get(p.data_ptr())))
                i += 1
        return out

    def state_dict(self):
        sd = super().state_dict()
        state = {}
        if self._m is not None and self._step > 0:
            fp = self.flat_provider()
            for i, name in self._indexed_names(fp):
                if name is None or name in fp.frozen:
                    continue  # never stepped: transformers' AdamW holds no state for it either
                state[i] = {"step": self._step, "exp_avg": fp.view(self._m, name).clone(),
                            "exp_avg_sq": fp.view(self._v, name).clone()}
        sd["state"] = state
        return sd

    def load_state_dict(self, sd):
        sd = dict(sd)  # shallow copy: the caller's dict keeps its keys
        flat = sd.pop("flat_state", None)  # round-1 checkpoints of this class
        state = sd.get("state", {})
        sd["state"] = {}
        super().load_state_dict(sd)
        if flat is not None:
            self._step, self._m, self._v = flat["step"], flat["m"], flat["v"]
            return
        if not state:
            return
        fp = self.flat_provider()
        m, v = self._moments(fp)
        m.zero_()
        v.zero_()
        steps = set()
        for i, name in self._indexed_names(fp):
            st = state.get(i, state.get(str(i)))
            if st is None:
                continue
            if name is None:
                raise RuntimeError("optimizer state for parameter %d cannot be mapped onto the flat buffers" % i)
            fp.view(m, name).copy_(st["exp_avg"])
            fp.view(v, name).copy_(st["exp_avg_sq"])
            steps.add(int(st["step"]))
        if len(steps) > 1:
            raise RuntimeError("per-parameter step counts differ (%s): the fused kernel keeps one step count" % steps)
        self._step = steps.pop() if steps else 0
