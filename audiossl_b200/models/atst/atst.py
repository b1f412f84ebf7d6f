"""ATST: student / EMA-teacher pair, BYOL loss, teacher update - audiossl/models/atst/atst.py:7-35.

Same constructor, attributes (.student/.teacher with .encoder/.projector/.predictor, .loss_fn),
``forward(melspecs, lengths) -> (loss, std_cls_s, std_cls_t)`` and ``update_teacher(m)`` as the reference.
Differences, all opt-in or invisible to existing recipes:
  * arch="large" is accepted (AST_large, embed_dim 1024),
  * the whole step runs on hand-written sm_100a kernels; ``loss.backward()`` triggers the explicit backward
    pass of the engine and leaves gradients in ``param.grad`` (views of one flat buffer),
  * under torch.distributed (one process per GPU) the module performs the data-parallel exchange itself:
    one NCCL all-reduce of the flat gradient buffer + SyncBatchNorm / compute_var statistics
    (SURVEY.md section 8e), so it must NOT additionally be wrapped in DistributedDataParallel.
"""
from functools import partial

import torch
from torch import nn

from ... import ops
from ...distributed import (GradExchange, allreduce_sum_, bn_stats_sync, bn_stats_sync_many, bn_sums_sync, world)
from ...engine import HEADS_3X, EncoderEngine, HeadEngine, Workspace, droppath_scales
from ...params import FlatParams
from .audio_transformer import AST, AST_base, AST_large, AST_small
from .byol import ByolLoss, MultiCropWrapper


class _StepFn(torch.autograd.Function):
    """autograd bridge: the forward value is the fused step's loss, backward runs the engine's backward."""

    @staticmethod
    def forward(ctx, anchor, model, loss):
        ctx.model = model
        return loss.clone()

    @staticmethod
    def backward(ctx, grad_out):
        ctx.model._rt.backward(grad_out)
        return None, None, None


class _Runtime:
    # parameters the forward pass never reads: they keep grad None and are never stepped / decayed (the reference's
    # AdamW skips them).  The frame runtime uses mask_embed and overrides this.
    never_used = ("encoder.mask_embed",)

    def __init__(self, model, device):
        self.model = model
        self.device = device
        model.student.to(device)   # BN buffers must live next to the parameters
        model.teacher.to(device)
        enc = model.student.encoder
        frozen = set(self.never_used) | {n for n, p in model.student.named_parameters() if not p.requires_grad}
        self.fs = FlatParams(list(model.student.named_parameters()), device, frozen=frozen)
        self.ft = FlatParams(list(model.teacher.named_parameters()), device, frozen=frozen)
        assert self.ft.total == self.fs.ema_count and self.ft.order == self.fs.order[:len(self.ft.order)], \
            "teacher layout must be the encoder+projector prefix of the student layout"
        self.enc = self._make_encoder(enc)
        self.proj = HeadEngine("projector.", enc.embed_dim)
        self.pred = HeadEngine("predictor.", 256)
        self.ws = Workspace(device)
        self.saved = None
        self.anchor = torch.zeros((), device=device, requires_grad=True)
        self.exchange = GradExchange(self.fs.grad, self.fs.exchange_start(), device)

    def _make_encoder(self, enc):
        return EncoderEngine(enc.embed_dim, enc.depth, enc.num_heads, use_cls=enc.use_cls, max_frames=enc.spec_w)

    def current(self):
        return self.fs.is_current() and self.ft.is_current()

    @staticmethod
    def _bn_buffers(seq):
        bn = seq[1]
        return bn.running_mean, bn.running_var, bn.num_batches_tracked

    def _encode(self, fp, net, crops, lengths, dp_list, save, tag):
        groups = MultiCropWrapper.group_crops(crops)
        outs, ctxs = [], []
        for gi, (s, e) in enumerate(groups):
            mel = crops[s] if e - s == 1 else torch.cat(crops[s:e])
            ln = None
            if lengths is not None:
                ln = lengths[s] if e - s == 1 else torch.cat(lengths[s:e])
            mel = mel.contiguous().float()
            S = mel.shape[0]
            enc = net.encoder
            dp = None
            if dp_list is not None:
                dp = dp_list[gi]
            elif net.training and enc.drop_path_rate > 0:
                dp = droppath_scales(enc.depth, enc.drop_path_rate, S, mel.device)
            out, ctx = self.enc.forward(fp, self.ws, mel, ln, dp=dp, save=save, tag="%s%d" % (tag, gi))
            outs.append(out)
            ctxs.append(ctx)
        return (outs[0] if len(outs) == 1 else torch.cat(outs)), ctxs

    def step(self, crops, lengths, dp_student=None, dp_teacher=None):
        m = self.model
        fs, ft = self.fs, self.ft
        ncrops = m.loss_fn.ncrops
        B = crops[0].shape[0]
        need_grad = torch.is_grad_enabled()
        ops.round_tf32(fs.data, fs.compute)
        ops.round_tf32(ft.data, ft.compute)
        sync = bn_stats_sync if world() > 1 else None
        # teacher on the first two crops (no grad; train-mode BN and DropPath exactly like the reference, D7)
        t_cls, _ = self._encode(ft, m.teacher, crops[:2], None if lengths is None else lengths[:2], dp_teacher,
                                False, "t")
        # student on all crops
        s_cls, enc_ctxs = self._encode(fs, m.student, crops, lengths, dp_student, need_grad, "s")
        # both projectors' first GEMM + batch statistics, then ONE SyncBatchNorm exchange for the two of them
        t_head = self.proj.forward_stats(ft, self.ws, t_cls, "t")
        s_head = self.proj.forward_stats(fs, self.ws, s_cls, "s")
        if world() > 1:
            (t_head["mean"], t_head["m2"]), (s_head["mean"], s_head["m2"]) = bn_stats_sync_many(
                [(h["mean"], h["m2"], h["n"]) for h in (t_head, s_head)])
            t_head["n"], s_head["n"] = t_head["n"] * world(), s_head["n"] * world()
        t_out, _ = self.proj.forward_finish(ft, self.ws, t_head, self._bn_buffers(m.teacher.projector), False)
        z, proj_ctx = self.proj.forward_finish(fs, self.ws, s_head, self._bn_buffers(m.student.projector), True)
        s_out, pred_ctx = self.pred.forward(fs, self.ws, z, self._bn_buffers(m.student.predictor), "s", False, sync)
        dstudent, acc = ops.byol_loss(s_out, t_out, ncrops, B, dstudent=self.ws.get("dstudent", s_out.shape),
                                      acc=self.ws.get("loss_acc", (1 + 4 * 256,)))
        allreduce_sum_(acc[1:])
        G = world()
        out3 = ops.byol_finalize(acc, ncrops * B * G, 2 * B * G, ncrops, B, out=self.ws.get("loss_out", (3,)))
        self.saved = (enc_ctxs, proj_ctx, pred_ctx, dstudent) if need_grad else None
        self.last_outputs = (s_out, t_out)
        return out3

    def _claim_gradient_buffer(self):
        """one backward pass per optimizer step: the flat gradient is rebuilt (and, under DDP, averaged) by every
        backward, so gradient accumulation over micro-batches would silently keep only the last one."""
        fs = self.fs
        if fs.has_optimizer and fs.grads_pending:
            raise RuntimeError("a second backward pass before optimizer.step() / zero_grad(): gradient accumulation "
                               "(accumulate_grad_batches > 1) is not supported by the fused step - raise the per-GPU "
                               "batch size instead")
        fs.grads_pending = True
        fs.grad.zero_()

    def backward(self, grad_out):
        if self.saved is None:
            raise RuntimeError("backward called without a recorded forward (was the step run under no_grad?)")
        enc_ctxs, proj_ctx, pred_ctx, dstudent = self.saved
        self.saved = None
        fs = self.fs
        self._claim_gradient_buffer()
        d = self.ws.get("dstudent_scaled", dstudent.shape)
        torch.mul(dstudent, grad_out.to(dstudent.dtype), out=d)
        if not HEADS_3X:
            ops.round_tf32(d, d)
        sums = bn_sums_sync if world() > 1 else None
        dz = self.pred.backward(fs, self.ws, pred_ctx, d, need_dx=True, sums_sync=sums)
        if not HEADS_3X:
            ops.round_tf32(dz, dz)
        dcls = self.proj.backward(fs, self.ws, proj_ctx, dz, need_dx=True, sums_sync=sums)
        if self.enc.debug is not None:
            self.enc.debug.append(("d_heads_in", "s", -1, dcls.clone()))
        # data-parallel exchange, overlapped: the heads' matrices now, each block's matrices as the backward pass
        # leaves it (only in the last crop group - earlier groups still accumulate into the same gradients), the
        # rest (biases, norms, embeddings) at the end
        ex = self.exchange
        ex.submit(*fs.matrix_range("predictor."))
        ex.submit(*fs.matrix_range("projector."))
        row = 0
        for gi, ctx in enumerate(enc_ctxs):
            S = ctx["S"]
            cb = (lambda i: ex.submit(*fs.matrix_range("encoder.blocks.%d." % i))) if gi == len(enc_ctxs) - 1 else None
            self.enc.backward(fs, self.ws, ctx, dcls[row:row + S], on_block_done=cb)
            row += S
        ex.finish()
        fs.attach_grads()


class ATST(nn.Module):
    def __init__(self, arch="small", ncrops=2, **kwargs):
        super().__init__()
        if isinstance(arch, dict):  # explicit sizes, e.g. dict(embed_dim=128, depth=2, num_heads=2) for tests
            cfg = dict(arch)
            embed_dim = cfg["embed_dim"]
            encoder_fn = partial(AST, patch_h=64, patch_w=4, qkv_bias=False,
                                 norm_layer=partial(nn.LayerNorm, eps=1e-6), **cfg)
        elif arch == "small":
            encoder_fn, embed_dim = AST_small, 384
        elif arch == "base":
            encoder_fn, embed_dim = AST_base, 768
        elif arch == "large":  # extension: the factory exists in the reference, the switch does not
            encoder_fn, embed_dim = AST_large, 1024
        else:
            raise RuntimeError("arch {} is not implemented".format(arch))
        self.student = MultiCropWrapper(encoder_fn(**kwargs), embed_dim, predictor=True)
        self.teacher = MultiCropWrapper(encoder_fn(**kwargs), embed_dim, predictor=False)
        for p in self.teacher.parameters():
            p.requires_grad = False
        self.teacher.load_state_dict({k: v for k, v in self.student.state_dict().items() if "predictor" not in k})
        self.loss_fn = ByolLoss(ncrops)
        self._rt = None
        self.ema_device_scalar = None  # 1-element cuda tensor: update_teacher reads m from it (CUDA-graph replay)

    def _runtime(self, device):
        if device.type != "cuda":
            raise RuntimeError("audiossl_b200 has no CPU path: move the module and the batch to a B200 (cuda) device")
        if self._rt is None or self._rt.device != device or not self._rt.current():
            self._rt = _Runtime(self, device)
        return self._rt

    def forward(self, melspecs, lengths, dp_student=None, dp_teacher=None):
        """melspecs: list of ncrops tensors [B,1,64,T_i]; lengths: list of ncrops int tensors [B].
        dp_student / dp_teacher: optional injected DropPath scales (parity tests), one list per encoder call."""
        rt = self._runtime(melspecs[0].device)
        out3 = rt.step(list(melspecs), None if lengths is None else list(lengths), dp_student, dp_teacher)
        loss = out3[0]
        if torch.is_grad_enabled():
            loss = _StepFn.apply(rt.anchor, self, loss)
        return loss, out3[1], out3[2]

    def update_teacher(self, m):
        """k <- m*k + (1-m)*q over encoder + projector parameters (not BN buffers, not the predictor)."""
        p = next(self.student.parameters())
        rt = self._runtime(p.device)
        ops.ema_update(rt.ft.data, rt.fs.data[:rt.fs.ema_count], float(m), m_dev=self.ema_device_scalar)
