"""FrameAST encoder shell: audiossl/methods/atstframe/audio_transformer.py:99-207, 283-290.

No CLS token; the student's masked frames are replaced by ``mask_embed`` before the positional embedding
(positions start at 1); the output is ``norm_frame`` of every token, from which the caller keeps the masked frames
inside the valid length.  State-dict keys: mask_embed, pos_embed, patch_embed.patch_embed.*, blocks.N.*,
norm_frame.* (138 for small, as in the reference)."""
from functools import partial

import torch
from torch import nn

from ...models.atst.audio_transformer import PatchEmbed_v2, _EncoderInference, get_num_patches
from ...modules.transformer import Block, trunc_normal_


class FrameAST(_EncoderInference, nn.Module):
    def __init__(self, nprompt=0, spec_h=64, spec_w=1001, patch_w=16, patch_h=16, pos_type="cut", avg_blocks=0,
                 in_chans=1, num_classes=0, embed_dim=768, depth=12, num_heads=12, mlp_ratio=4., qkv_bias=False,
                 qk_scale=None, drop_rate=0., attn_drop_rate=0., drop_path_rate=0.1, norm_layer=nn.LayerNorm,
                 patch_embed="Linear", **kwargs):
        super().__init__()
        if patch_h != 64 or patch_w != 4 or spec_h != 64:
            raise NotImplementedError("the CUDA path implements the 64x4 patches of every ATST recipe")
        if pos_type != "cut" or avg_blocks != 0 or nprompt != 0 or patch_embed != "Linear" or mlp_ratio != 4.:
            raise NotImplementedError("pos_type=interpolate / avg_blocks>0 (data2vec) / nprompt>0 / patch_embed=CNN "
                                      "are not exercised by the ATST-Frame recipes (SURVEY.md section 8a)")
        self.num_features = self.embed_dim = embed_dim
        self.spec_w, self.spec_h, self.patch_w, self.patch_h = spec_w, spec_h, patch_w, patch_h
        self.depth, self.num_heads, self.drop_path_rate = depth, num_heads, drop_path_rate
        self.pos_type, self.avg_blocks, self.nprompt, self.use_cls = pos_type, avg_blocks, nprompt, False
        self.patch_embed = PatchEmbed_v2(patch_h, patch_w, embed_dim)
        self.mask_embed = nn.Parameter(torch.zeros(1, 1, embed_dim))
        self.num_patches = get_num_patches(spec_h, spec_w, patch_h, patch_w)
        self.pos_embed = nn.Parameter(torch.zeros(1, self.num_patches + 1, embed_dim))
        dpr = [x.item() for x in torch.linspace(0, drop_path_rate, depth)]
        self.blocks = nn.ModuleList([
            Block(dim=embed_dim, num_heads=num_heads, mlp_ratio=mlp_ratio, qkv_bias=qkv_bias, qk_scale=qk_scale,
                  drop=drop_rate, attn_drop=attn_drop_rate, drop_path=dpr[i], norm_layer=norm_layer)
            for i in range(depth)])
        self.norm_frame = norm_layer(embed_dim)
        trunc_normal_(self.pos_embed, std=.02)
        trunc_normal_(self.mask_embed, std=.02)
        self.apply(self._init_weights)

    def _init_weights(self, m):
        if isinstance(m, nn.Linear):
            trunc_normal_(m.weight, std=.02)
            if m.bias is not None:
                nn.init.constant_(m.bias, 0)
        elif isinstance(m, nn.LayerNorm):
            nn.init.constant_(m.bias, 0)
            nn.init.constant_(m.weight, 1.0)


def FrameAST_small(patch_h=64, patch_w=4, **kwargs):
    return FrameAST(patch_h=patch_h, patch_w=patch_w, embed_dim=384, depth=12, num_heads=6, qkv_bias=False,
                    norm_layer=partial(nn.LayerNorm, eps=1e-6), **kwargs)


def FrameAST_base(patch_h=64, patch_w=4, **kwargs):
    return FrameAST(patch_h=patch_h, patch_w=patch_w, embed_dim=768, depth=12, num_heads=12, qkv_bias=False,
                    norm_layer=partial(nn.LayerNorm, eps=1e-6), **kwargs)


def FrameAST_large(patch_h=64, patch_w=4, **kwargs):
    return FrameAST(patch_h=patch_h, patch_w=patch_w, embed_dim=1024, depth=24, num_heads=16, qkv_bias=False,
                    norm_layer=partial(nn.LayerNorm, eps=1e-6), **kwargs)
