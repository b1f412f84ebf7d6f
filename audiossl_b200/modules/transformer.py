"""Parameter containers mirroring audiossl/modules/transformer.py:70-150 (Mlp, Attention, Block).

They hold the nn.Parameters under the reference's state-dict keys and reproduce its initialisation;
the arithmetic is executed by audiossl_b200.engine.EncoderEngine on hand-written kernels.
"""
import torch
import torch.nn as nn


def trunc_normal_(tensor, mean=0., std=1., a=-2., b=2.):
    return nn.init.trunc_normal_(tensor, mean=mean, std=std, a=a, b=b)


class Mlp(nn.Module):
    def __init__(self, in_features, hidden_features=None, out_features=None, **kwargs):
        super().__init__()
        out_features = out_features or in_features
        hidden_features = hidden_features or in_features
        self.fc1 = nn.Linear(in_features, hidden_features)
        self.fc2 = nn.Linear(hidden_features, out_features)
        trunc_normal_(self.fc1.weight, std=0.02)
        trunc_normal_(self.fc2.weight, std=0.02)
        nn.init.constant_(self.fc1.bias, 0)
        nn.init.constant_(self.fc2.bias, 0)


class Attention(nn.Module):
    def __init__(self, dim, num_heads=8, qkv_bias=False, qk_scale=None, attn_drop=0., proj_drop=0.):
        super().__init__()
        if qkv_bias or attn_drop or proj_drop or qk_scale is not None:
            raise NotImplementedError("qkv_bias / dropout / qk_scale are not exercised by any ATST config")
        self.num_heads = num_heads
        self.scale = (dim // num_heads) ** -0.5
        self.qkv = nn.Linear(dim, dim * 3, bias=False)
        self.proj = nn.Linear(dim, dim)


class Block(nn.Module):
    def __init__(self, dim, num_heads, mlp_ratio=4., qkv_bias=False, qk_scale=None, drop=0., attn_drop=0.,
                 drop_path=0., act_layer=nn.GELU, norm_layer=nn.LayerNorm):
        super().__init__()
        if drop:
            raise NotImplementedError("drop_rate > 0 is not exercised by any ATST config")
        self.norm1 = norm_layer(dim)
        self.attn = Attention(dim, num_heads=num_heads, qkv_bias=qkv_bias, qk_scale=qk_scale, attn_drop=attn_drop)
        self.drop_path_rate = drop_path
        self.norm2 = norm_layer(dim)
        self.mlp = Mlp(in_features=dim, hidden_features=int(dim * mlp_ratio))


def get_attention_mask(x, length):
    """additive key-padding mask of the reference (modules/transformer.py:152-159); the CUDA path never
    materialises it (lengths go straight to the attention kernel) - kept for API parity."""
    batch_size, max_len, _ = x.shape
    mask = torch.arange(max_len, device=length.device).expand(batch_size, max_len) >= length[:, None]
    mask = -10000.0 * mask[:, None, None, :]
    return mask.expand(batch_size, 1, max_len, max_len).to(x.device)
