"""FrameATST / FrameATSTLightningModule - audiossl/methods/atstframe/model.py:24-166 on the CUDA engine.

Step (symmetric branch, model.py:68-72): teacher(x, length, mask, mask_input=False), student(..., True), frame-level
BYOL loss on the masked frames inside the valid length.  The masked-row count is data dependent: the valid
indices are compacted once per step (one host read of the count), both networks gather those rows
(atst_gather_rows) before the heads, and the backward pass scatters the head gradients back
(atst_scatter_rows).  BatchNorm statistics use the true (per-rank unequal) row counts."""
import torch
from torch import nn

from ... import ops
from ...distributed import allreduce_sum_, bn_stats_sync, bn_sums_sync, world
from ...engine import HEADS_3X, EncoderEngine, droppath_scales
from ...models.atst.atst import _Runtime, _StepFn
from ...optim import FusedHFAdamW
from ...utils.common import bool_flag, cosine_scheduler_step, get_params_groups
from ..atst.model import LightningModule
from .audio_transformer import FrameAST_base, FrameAST_large, FrameAST_small, FrameAST
from .byol import ByolLoss, MultiCropWrapper


class _FrameRuntime(_Runtime):
    never_used = ()  # the student blends mask_embed into the masked patches

    def _make_encoder(self, enc):
        return EncoderEngine(enc.embed_dim, enc.depth, enc.num_heads, use_cls=False, norm_name="norm_frame",
                             max_frames=enc.spec_w)

    def _frames(self, fp, net, mel, ln, mask, mask_input, idx, save, tag):
        enc = net.encoder
        dp = None
        if net.training and enc.drop_path_rate > 0:
            dp = droppath_scales(enc.depth, enc.drop_path_rate, mel.shape[0], mel.device)
        xn, ctx = self.enc.forward(fp, self.ws, mel, ln, dp=dp, save=save, tag=tag, mask=mask, mask_input=mask_input)
        rows = ops.gather_rows(xn, idx, self.ws.get(tag + "/rows", (idx.numel(), self.enc.D)))
        if self.enc.debug is not None:
            self.enc.debug.append(("heads_in", tag, -1, rows.clone()))
        return rows, ctx

    def step(self, crops, lengths, masks):
        m = self.model
        fs, ft = self.fs, self.ft
        need_grad = torch.is_grad_enabled()
        ops.round_tf32(fs.data, fs.compute)
        ops.round_tf32(ft.data, ft.compute)
        mel = torch.cat(list(crops)).contiguous().float()
        ln = torch.cat(list(lengths))
        mask = torch.cat(list(masks)).to(torch.bool)
        S, P = mask.shape
        plen = (ln - ln % 4) // 4
        valid = mask & (torch.arange(P, device=mask.device)[None, :] < plen[:, None])
        idx = torch.nonzero(valid.reshape(-1)).to(torch.int32).reshape(-1).contiguous()
        R = idx.numel()  # host read of the data-dependent row count (the reference's boolean indexing syncs too)
        if R < 2 or R % 2:
            raise RuntimeError("ATST-Frame expects the same mask for both views (even masked-frame count), got %d" % R)
        G = world()
        n_glob = float(R)
        if G > 1:  # ranks hold different numbers of masked frames: ONE count exchange per step serves the three
            cnt = torch.tensor([float(R)], device=mel.device)  # BatchNorm layers and the loss (no per-layer host read)
            allreduce_sum_(cnt)
            n_glob = cnt.item()
        sync = (lambda mean, m2, n: bn_stats_sync(mean, m2, n, n_total=n_glob)) if G > 1 else None
        t_rows, _ = self._frames(ft, m.teacher, mel, ln, mask, False, idx, False, "t0")
        t_out, _ = self.proj.forward(ft, self.ws, t_rows, self._bn_buffers(m.teacher.projector), "t", False, sync)
        s_rows, enc_ctx = self._frames(fs, m.student, mel, ln, mask, True, idx, need_grad, "s0")
        z, proj_ctx = self.proj.forward(fs, self.ws, s_rows, self._bn_buffers(m.student.projector), "s", True, sync)
        s_out, pred_ctx = self.pred.forward(fs, self.ws, z, self._bn_buffers(m.student.predictor), "s", False, sync)
        dstudent, acc = ops.byol_loss(s_out, t_out, 2, R // 2, dstudent=self.ws.get("dstudent", s_out.shape),
                                      acc=self.ws.get("loss_acc", (1 + 4 * 256,)))
        if G > 1:
            allreduce_sum_(acc[1:])
        out3 = ops.byol_finalize(acc, n_glob, n_glob, 2, R // 2, out=self.ws.get("loss_out", (3,)))
        self.saved = (enc_ctx, proj_ctx, pred_ctx, dstudent, idx) if need_grad else None
        self.last_outputs = (s_out, t_out)
        return out3

    def backward(self, grad_out):
        if self.saved is None:
            raise RuntimeError("backward called without a recorded forward (was the step run under no_grad?)")
        enc_ctx, proj_ctx, pred_ctx, dstudent, idx = self.saved
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
        drows = self.proj.backward(fs, self.ws, proj_ctx, dz, need_dx=True, sums_sync=sums)
        if self.enc.debug is not None:
            self.enc.debug.append(("d_heads_in", "s", -1, drows.clone()))
        dxn = self.ws.get("s0/bwd/dxn", (enc_ctx["M"], self.enc.D))
        dxn.zero_()
        ops.scatter_rows(drows, idx, dxn)
        ex = self.exchange
        ex.submit(*fs.matrix_range("predictor."))
        ex.submit(*fs.matrix_range("projector."))
        self.enc.backward(fs, self.ws, enc_ctx, dxn,
                          on_block_done=lambda i: ex.submit(*fs.matrix_range("encoder.blocks.%d." % i)))
        ex.finish()
        fs.attach_grads()


class FrameATST(nn.Module):
    def __init__(self, arch="small", symmetric=True, pos_type="cut", avg_blocks=0, patch_embed="Linear", **kwargs):
        super().__init__()
        if isinstance(arch, dict):
            from functools import partial
            cfg = dict(arch)
            embed_dim = cfg["embed_dim"]
            encoder_fn = partial(FrameAST, patch_h=64, patch_w=4, qkv_bias=False,
                                 norm_layer=partial(nn.LayerNorm, eps=1e-6), **cfg)
        elif arch == "small":
            encoder_fn, embed_dim = FrameAST_small, 384
        elif arch == "base":
            encoder_fn, embed_dim = FrameAST_base, 768
        elif arch == "large":
            encoder_fn, embed_dim = FrameAST_large, 1024
        else:
            raise RuntimeError("arch {} is not implemented".format(arch))
        if avg_blocks != 0:
            raise NotImplementedError("avg_blocks>0 (data2vec-style targets) is not used by the ATST-Frame recipes")
        self.symmetric = symmetric
        self.student = MultiCropWrapper(encoder_fn(pos_type=pos_type, patch_embed=patch_embed, **kwargs), embed_dim,
                                        predictor=True)
        self.teacher = MultiCropWrapper(encoder_fn(pos_type=pos_type, patch_embed=patch_embed, **kwargs), embed_dim,
                                        predictor=False)
        for p in self.teacher.parameters():
            p.requires_grad = False
        self._init_teacher()
        self.loss_fn = ByolLoss(symmetric=symmetric)
        self._rt = None

    def _runtime(self, device):
        if device.type != "cuda":
            raise RuntimeError("audiossl_b200 has no CPU path: move the module and the batch to a B200 (cuda) device")
        if self._rt is None or self._rt.device != device or not self._rt.current():
            self._rt = _FrameRuntime(self, device)
        return self._rt

    def forward(self, x, length, mask):
        rt = self._runtime(x[0].device)
        out3 = rt.step(x, length, mask)
        loss = out3[0]
        if torch.is_grad_enabled():
            loss = _StepFn.apply(rt.anchor, self, loss)
        return loss, out3[1], out3[2]

    def update_teacher(self, m):
        p = next(self.student.parameters())
        rt = self._runtime(p.device)
        ops.ema_update(rt.ft.data, rt.fs.data[:rt.fs.ema_count], float(m))

    def _init_teacher(self):
        self.teacher.load_state_dict({k: v for k, v in self.student.state_dict().items() if "predictor" not in k})


class FrameATSTLightningModule(LightningModule):
    def __init__(self, arch="small", learning_rate: float = 5e-4, warmup_steps=1300, max_steps=39000, ema=0.99,
                 symmetric=True, pos_type="cut", avg_blocks=0, patch_embed="Linear", **kwargs):
        super().__init__()
        model_kwargs = {k: kwargs[k] for k in ("drop_path_rate",) if k in kwargs}
        self.model = FrameATST(arch=arch, symmetric=symmetric, pos_type=pos_type, avg_blocks=avg_blocks,
                               patch_embed=patch_embed, **model_kwargs)
        self.learning_rate, self.warmup_steps, self.max_steps, self.symmetric = learning_rate, warmup_steps, max_steps, symmetric
        self.ema_scheduler = cosine_scheduler_step(ema, 1, max_steps, 0)
        self.wd_scheduler = cosine_scheduler_step(0.04, 0.4, max_steps, 0)
        self.mylr_scheduler = cosine_scheduler_step(learning_rate, 1e-6, max_steps, warmup_steps)
        self.save_hyperparameters()

    def training_step(self, batch, batch_idx):
        self.schedule()
        (melspecs, lengths, masks), _ = batch
        total_loss_frm, std_frm_stu, std_frm_tea = self.model(melspecs, lengths, masks)
        loss = total_loss_frm
        self.log("loss", loss, prog_bar=True, logger=True)
        self.log("loss_frm", total_loss_frm, prog_bar=True, logger=True)
        self.log("std_frm_tea", std_frm_tea, prog_bar=True, logger=True)
        self.log("std_frm_stu", std_frm_stu, prog_bar=True, logger=True)
        self.log("ema", self.ema_scheduler[self.global_step], prog_bar=True, logger=True)
        self.log("step", self.global_step, prog_bar=True, logger=True)
        return loss

    def schedule(self):
        for i, param_group in enumerate(self.trainer.optimizers[0].param_groups):
            param_group["lr"] = self.mylr_scheduler[self.global_step]
            if i == 0:
                param_group["weight_decay"] = self.wd_scheduler[self.global_step]
        self.log("wd", self.wd_scheduler[self.global_step], prog_bar=True, logger=True)
        self.log("lr", param_group["lr"], prog_bar=True, logger=True)

    def configure_optimizers(self):
        def flat():
            p = next(self.model.student.parameters())
            return self.model._runtime(p.device).fs
        return [FusedHFAdamW(get_params_groups(self.model.student), flat=flat, lr=self.learning_rate, weight_decay=0.)]

    def on_train_batch_end(self, outputs, batch, batch_idx: int, unused: int = 0) -> None:
        self.model.update_teacher(self.ema_scheduler[self.global_step])

    @staticmethod
    def add_model_specific_args(parent_parser):
        parser = parent_parser.add_argument_group("FrameATSTModel")
        parser.add_argument("--arch", type=str, default="small")
        parser.add_argument("--symmetric", type=bool_flag, default=True, help="whether to use symemtric loss")
        parser.add_argument("--nprompt", type=int, default=0, help="number of prompts, not used, always 0")
        parser.add_argument("--learning_rate", default=0.0005, type=float)
        parser.add_argument('--ema', default=0.99, type=float)
        parser.add_argument('--warmup_steps', default=1300, type=int)
        parser.add_argument('--max_steps', default=39010, type=int)
        parser.add_argument('--pos_type', default="cut", type=str)
        parser.add_argument('--avg_blocks', default=0, type=int)
        parser.add_argument('--patch_embed', default="Linear", type=str)
        return parent_parser
