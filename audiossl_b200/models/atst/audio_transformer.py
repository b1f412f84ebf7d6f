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


class AST(nn.Module):
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


def AST_small(patch_h=64, patch_w=4, **kwargs):
    return AST(patch_h=patch_h, patch_w=patch_w, embed_dim=384, depth=12, num_heads=6, qkv_bias=False,
               norm_layer=partial(nn.LayerNorm, eps=1e-6), **kwargs)


def AST_base(patch_h=64, patch_w=4, **kwargs):
    return AST(patch_h=patch_h, patch_w=patch_w, embed_dim=768, depth=12, num_heads=12, qkv_bias=False,
               norm_layer=partial(nn.LayerNorm, eps=1e-6), **kwargs)


def AST_large(patch_h=64, patch_w=4, **kwargs):
    return AST(patch_h=patch_h, patch_w=patch_w, embed_dim=1024, depth=24, num_heads=16, qkv_bias=False,
               norm_layer=partial(nn.LayerNorm, eps=1e-6), **kwargs)
