"""AST encoder shell: audiossl/models/atst/audio_transformer.py:56-221, 367-374.

Holds parameters under the reference's keys (cls_token, pos_embed, mask_embed, patch_embed.patch_embed.*,
blocks.N.*, norm.*) with the reference's initialisation (trunc-normal 0.02, zero biases, unit LayerNorm).
Forward runs on the CUDA engine.  Options no reference recipe uses raise NotImplementedError
(SURVEY.md section 8a "options present ... that the five configs do not exercise").
"""
from functools import partial

import torch
from torch import nn

from ...modules.transformer import Block, trunc_normal_


def get_num_patches(height=64, width=1001, patch_height=16, patch_width=16):
    return (height // patch_height) * (width // patch_width)


class PatchEmbed_v2(nn.Module):
    def __init__(self, patch_height=64, patch_width=4, embed_dim=768):
        super().__init__()
        self.patch_height, self.patch_width = patch_height, patch_width
        self.patch_embed = nn.Linear(patch_height * patch_width, embed_dim)


def get_cls_avg(output_i, cur_len, use_cls):
    """CLS token and length-masked mean of the patch tokens for each collected layer
    (audiossl/models/atst/audio_transformer.py:355-366)."""
    n_tok = output_i[0].shape[1] - (1 if use_cls else 0)
    length_mask = torch.arange(n_tok, device=output_i[0].device) < cur_len.unsqueeze(1)
    if use_cls:
        cls = [x[:, 0] for x in output_i]
        avg = [torch.sum(x[:, 1:] * length_mask.unsqueeze(-1), dim=1) / (cur_len.unsqueeze(1) + 1e-6) for x in output_i]
    else:
        cls = [torch.zeros_like(x[:, 0]) for x in output_i]
        avg = [torch.sum(x * length_mask.unsqueeze(-1), dim=1) / (cur_len.unsqueeze(1) + 1e-6) for x in output_i]
    return cls, avg


class _EncoderInference:
    """Inference entry points shared by AST and FrameAST: the same CUDA engine as
    training, no saved activations, no gradient.  The encoder gets its own flat parameter buffer the first time it
    is called stand-alone (e.g. ``model.teacher.encoder`` handed to a downstream probe)."""

    _inf = None

    def _inference_runtime(self, device):
        from ...engine import EncoderEngine, Workspace
        from ...params import FlatParams
        if device.type != "cuda":
            raise RuntimeError("audiossl_b200 has no CPU path: move the encoder and its input to a B200 (cuda) device")
        rt = self._inf
        if rt is None or rt["device"] != device or not rt["fp"].is_current():
            fp, prefix = self._owner_storage(device)
            if fp is None:  # a free-standing encoder (e.g. returned by load_model): its own flat buffer
                self.to(device)
                fp, prefix = FlatParams(list(self.named_parameters()), device, ema_prefixes=("",)), ""
            eng = EncoderEngine(self.embed_dim, self.depth, self.num_heads, use_cls=self.use_cls,
                                norm_name="norm" if self.use_cls else "norm_frame", prefix=prefix,
                                max_frames=self.spec_w)
            rt = dict(device=device, fp=fp, eng=eng, ws=Workspace(device))
            self._inf = rt
        return rt

    def _owner_storage(self, device):
        """(FlatParams, name prefix) when this encoder's parameters already live in the flat buffer of a training
        runtime (``model.teacher.encoder`` during validation): inference then reads that storage in place instead of
        re-pointing the parameters at a private copy - which would invalidate the training runtime and make every
        eval / train alternation rebuild tens of GB of buffers."""
        owners = [getattr(p, "_atst_flat", None) for p in self.parameters()]
        if not owners or any(o is None for o in owners) or any(o[0] is not owners[0][0] for o in owners):
            return None, ""
        fp = owners[0][0]
        if fp.data.device != device or not fp.is_current():
            return None, ""
        mine = next(iter(dict(self.named_parameters())))
        full = next(o[1] for o, (n, _) in zip(owners, self.named_parameters()) if n == mine)
        return fp, full[:len(full) - len(mine)]

    @torch.no_grad()
    def _run(self, x, length, collect=0, mask_index=None, mask_input=False):
        from ... import ops
        rt = self._inference_runtime(x.device)
        fp = rt["fp"]
        ops.round_tf32(fp.data, fp.compute)
        out, ctx = rt["eng"].forward(fp, rt["ws"], x.contiguous().float(), length, dp=None, save=False, tag="inf",
                                     mask=mask_index, mask_input=mask_input, collect=collect, round_final=False)
        return out, ctx

    @torch.no_grad()
    def get_intermediate_layers(self, x, length, n=1, scene=True):
        """final-norm token outputs of the last n blocks.  AST: list of [B, N, D] (models/atst/audio_transformer.py:
        235-256).  FrameAST: concatenation over layers of the length-masked mean (scene=True) or of the frame
        sequences (scene=False) (methods/atstframe/audio_transformer.py:259-281)."""
        B = x.shape[0]
        _, ctx = self._run(x, length, collect=n)
        N = ctx["N"]
        layers = [y.view(B, N, self.embed_dim).clone() for y in ctx["collected"]]
        if self.use_cls:
            return layers
        if not scene:
            return torch.cat(layers, dim=-1)
        plen = (length - length % self.patch_w) // self.patch_w
        mask = (torch.arange(N, device=x.device) < plen.unsqueeze(1)).unsqueeze(-1)
        return torch.cat([torch.sum(y * mask, dim=1) / (plen.unsqueeze(-1) + 1e-6) for y in layers], dim=-1)

    @torch.no_grad()
    def get_intermediate_layers_chunks(self, x, length, n=1, chunk_len=601, avgpool=True):
        """long audio in chunk_len-frame chunks, CLS / mean-pooled tokens of the last n blocks averaged over the
        chunks that contain audio (models/atst/audio_transformer.py:257-353)."""
        total_len = x.shape[-1]
        num_chunks = total_len // chunk_len + 1
        cls, avg, marks = [], [], []
        for i in range(num_chunks):
            cur_len = torch.clip(length - i * chunk_len, 0)
            marks.append(cur_len > 0 if i == 0 else cur_len > chunk_len // 2)
            xc = x[:, :, :, i * chunk_len:min((i + 1) * chunk_len, total_len)]
            if xc.shape[-1] < self.patch_w:
                marks.pop()
                continue
            B = xc.shape[0]
            _, ctx = self._run(xc, cur_len, collect=n)
            outs = [y.view(B, ctx["N"], self.embed_dim) for y in ctx["collected"]]
            plen = (cur_len - cur_len % self.patch_w) // self.patch_w
            c_, a_ = get_cls_avg(outs, plen, self.use_cls)
            cls.append([t.clone() for t in c_])  # views of the reused workspace: copy before the next chunk runs
            avg.append(a_)
        mark = torch.stack(marks, dim=0).unsqueeze(-1).float()  # [chunks, B, 1]
        cls_out = [torch.sum(torch.stack(list(c), 0) * mark, 0) / torch.sum(mark, 0) for c in zip(*cls)]
        avg_out = [torch.sum(torch.stack(list(a), 0) * mark, 0) / torch.sum(mark, 0) for a in zip(*avg)]
        return torch.cat(cls_out + avg_out, dim=-1) if avgpool else torch.cat(cls_out, dim=-1)


class AST(_EncoderInference, nn.Module):
    def __init__(self, use_cls=True, spec_h=64, spec_w=1001, patch_w=16, patch_h=16, in_chans=1, num_classes=0,
                 embed_dim=768, depth=12, num_heads=12, mlp_ratio=4., qkv_bias=False, qk_scale=None, drop_rate=0.,
                 attn_drop_rate=0., drop_path_rate=0.1, norm_layer=nn.LayerNorm, mask_ratio=0, pos_type="cut",
                 **kwargs):
        super().__init__()
        if patch_h != 64 or patch_w != 4 or spec_h != 64:
            raise NotImplementedError("the CUDA path implements the 64x4 patches of every ATST recipe")
        if pos_type != "cut":
            raise NotImplementedError('pos_type="interpolate" is not used by any ATST recipe')
        if mlp_ratio != 4.:
            raise NotImplementedError("mlp_ratio != 4")
        if not use_cls:
            # the reference's AST(use_cls=False) adds pos_embed[:, :T] and returns the length-masked token mean
            # (audio_transformer.py:180-186, 211-221); the engine's no-CLS path is FrameAST's (positions from 1,
            # per-token output, norm_frame) - a different function, so refuse instead of computing that silently
            raise NotImplementedError("AST(use_cls=False) is not used by any ATST recipe; the no-CLS encoder on the "
                                      "CUDA engine is audiossl_b200.methods.atstframe.audio_transformer.FrameAST")
        self.num_features = self.embed_dim = embed_dim
        self.spec_w, self.spec_h, self.patch_w, self.patch_h = spec_w, spec_h, patch_w, patch_h
        self.depth, self.num_heads, self.drop_path_rate = depth, num_heads, drop_path_rate
        self.patch_embed = PatchEmbed_v2(patch_h, patch_w, embed_dim)
        self.mask_embed = nn.Parameter(torch.zeros(1, 1, embed_dim))
        self.num_patches = get_num_patches(spec_h, spec_w, patch_h, patch_w)
        self.use_cls = use_cls
        if use_cls:
            self.cls_token = nn.Parameter(torch.zeros(1, 1, embed_dim))
        self.pos_embed = nn.Parameter(torch.zeros(1, self.num_patches + 1, embed_dim))
        self.pos_type = pos_type
        dpr = [x.item() for x in torch.linspace(0, drop_path_rate, depth)]
        self.blocks = nn.ModuleList([
            Block(dim=embed_dim, num_heads=num_heads, mlp_ratio=mlp_ratio, qkv_bias=qkv_bias, qk_scale=qk_scale,
                  drop=drop_rate, attn_drop=attn_drop_rate, drop_path=dpr[i], norm_layer=norm_layer)
            for i in range(depth)])
        self.norm = norm_layer(embed_dim)
        trunc_normal_(self.pos_embed, std=.02)
        trunc_normal_(self.mask_embed, std=.02)
        if use_cls:
            trunc_normal_(self.cls_token, std=.02)
        self.apply(self._init_weights)

    def _init_weights(self, m):
        if isinstance(m, nn.Linear):
            trunc_normal_(m.weight, std=.02)
            if m.bias is not None:
                nn.init.constant_(m.bias, 0)
        elif isinstance(m, nn.LayerNorm):
            nn.init.constant_(m.bias, 0)
            nn.init.constant_(m.weight, 1.0)

    @torch.no_grad()
    def forward(self, x, mask_index=None, length=None, avg=False):
        """stand-alone (inference) forward: final-norm CLS embedding [B, D] (audio_transformer.py:188-210).
        Training goes through ATST.forward, which drives both networks and the backward pass."""
        if avg or mask_index is not None:
            raise NotImplementedError("avg=True / mask_index are not used by the ATST-clip recipes")
        out, _ = self._run(x, length)
        return out.clone()


def AST_small(patch_h=64, patch_w=4, **kwargs):
    return AST(patch_h=patch_h, patch_w=patch_w, embed_dim=384, depth=12, num_heads=6, qkv_bias=False,
               norm_layer=partial(nn.LayerNorm, eps=1e-6), **kwargs)


def AST_base(patch_h=64, patch_w=4, **kwargs):
    return AST(patch_h=patch_h, patch_w=patch_w, embed_dim=768, depth=12, num_heads=12, qkv_bias=False,
               norm_layer=partial(nn.LayerNorm, eps=1e-6), **kwargs)


def AST_large(patch_h=64, patch_w=4, **kwargs):
    return AST(patch_h=patch_h, patch_w=patch_w, embed_dim=1024, depth=24, num_heads=16, qkv_bias=False,
               norm_layer=partial(nn.LayerNorm, eps=1e-6), **kwargs)
