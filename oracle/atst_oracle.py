"""CPU oracle for the ATST pre-training hot path.  TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import this module.  The product path
(``audiossl_b200``) never imports it and has no CPU fallback.

It restates, in plain numpy / fp32 torch-CPU, the algorithm of the reference
(Audio-WestlakeU/audiossl @ ec3a14d) for SURVEY.md section 8(a) rows a2, a8-a16:

* ``mel_feature``        - torchaudio ``MelSpectrogram`` + ``AmplitudeToDB`` + ``MinMax``
                           (reference call site audiossl/methods/atst/transform.py:14-29;
                           arithmetic lives in torchaudio 2.x functional.py:54-145, 356-404,
                           492-575, an un-vendored dependency: README.md:29 pins 2.1.1).
* ``OracleAST``          - audiossl/models/atst/audio_transformer.py:56-75,78-221 and
                           audiossl/modules/transformer.py:70-159 (Block/Attention/Mlp/mask).
* ``OracleMultiCrop``    - audiossl/models/atst/byol.py:6-22,82-121.
* ``byol_loss``          - audiossl/models/atst/byol.py:24-41,42-53,57-78.
* ``ema_update``         - audiossl/models/atst/atst.py:29-34.
* ``hf_adamw_step``      - transformers 4.x ``AdamW`` (removed from the installed 5.5; restated
                           from its published algorithm; PARITY UNPINNED for this one function -
                           its Adam half is cross-checked against torch.optim.Adam through the exact
                           eps re-parametrisation in tests/test_oracle_golden.py).
* ``cosine_scheduler_step`` / ``param_groups`` - audiossl/utils/common.py:29-39,41-68.

Pinning: ``tests/golden/make_golden.py`` imports the unmodified reference from
/root/reference (in the build container only) and writes ``tests/golden/*.npz``;
``tests/test_oracle_golden.py`` checks every function here against those vectors.
"""
from __future__ import annotations

import math
from functools import partial

import numpy as np
import torch
from torch import nn
from torch.nn import functional as F

# --------------------------------------------------------------------------------------
# mel front-end (numpy fp32; optional fp64 to arbitrate fp32 disagreements)
# --------------------------------------------------------------------------------------
SR = 16000
N_FFT = 1024
HOP = 160
N_MELS = 64
F_MIN = 60.0
F_MAX = 7800.0
TOP_DB = 80.0
MINMAX_MIN = -79.6482
MINMAX_MAX = 50.6842


def hz_to_mel_htk(f):
    return 2595.0 * np.log10(1.0 + f / 700.0)


def mel_filterbank(dtype=np.float32):
    """[513, 64] HTK triangular filterbank, norm=None (torchaudio melscale_fbanks)."""
    n_freqs = N_FFT // 2 + 1
    all_freqs = np.linspace(0, SR // 2, n_freqs)
    m_min, m_max = hz_to_mel_htk(F_MIN), hz_to_mel_htk(F_MAX)
    m_pts = np.linspace(m_min, m_max, N_MELS + 2)
    f_pts = 700.0 * (10.0 ** (m_pts / 2595.0) - 1.0)
    # torchaudio computes these in fp32; do the same so the weights are bit-comparable
    all_freqs = all_freqs.astype(np.float32)
    f_pts = f_pts.astype(np.float32)
    f_diff = f_pts[1:] - f_pts[:-1]
    slopes = f_pts[None, :] - all_freqs[:, None]
    down = (-1.0 * slopes[:, :-2]) / f_diff[:-1]
    up = slopes[:, 2:] / f_diff[1:]
    fb = np.maximum(0.0, np.minimum(down, up))
    return fb.astype(dtype)


def hann_window(win_length=N_FFT, dtype=np.float32):
    """periodic Hann of win_length, zero-padded (centred) to N_FFT as torch.stft does."""
    n = np.arange(win_length, dtype=np.float64)
    w = 0.5 - 0.5 * np.cos(2.0 * np.pi * n / win_length)
    left = (N_FFT - win_length) // 2
    out = np.zeros(N_FFT, dtype=np.float64)
    out[left:left + win_length] = w
    return out.astype(dtype)


def mel_power(wav, win_length=N_FFT, dtype=np.float32):
    """wav [..., n] -> mel power [..., 64, n//160+1]."""
    wav = np.asarray(wav, dtype=dtype)
    lead = wav.shape[:-1]
    x = wav.reshape(-1, wav.shape[-1])
    pad = N_FFT // 2
    xp = np.pad(x, ((0, 0), (pad, pad)), mode="reflect")
    n = x.shape[-1]
    T = n // HOP + 1
    idx = np.arange(T)[:, None] * HOP + np.arange(N_FFT)[None, :]
    win = hann_window(win_length, dtype)
    frames = xp[:, idx] * win  # [B, T, 1024]
    spec = np.fft.rfft(frames, axis=-1)
    power = (spec.real.astype(dtype) ** 2 + spec.imag.astype(dtype) ** 2).astype(dtype)
    fb = mel_filterbank(dtype)
    mel = power @ fb  # [B, T, 64]
    mel = np.swapaxes(mel, -1, -2)
    return mel.reshape(*lead, N_MELS, T)


def mel_feature(wav, win_length=N_FFT, dtype=np.float32):
    """Normalised log-mel in [-1, 1]-ish scale: [..., n] -> [..., 64, T].

    AmplitudeToDB(power, top_db=80): 10*log10(max(x, 1e-10)), clamp at (per-clip max - 80)
    where a "clip" is the trailing [C, F, T] block (torchaudio packs 3-D input as one clip,
    4-D input per leading batch index).  MinMax: (x - min)/(max - min)*2 - 1.
    """
    mel = mel_power(wav, win_length, dtype)
    db = (10.0 * np.log10(np.maximum(mel, dtype(1e-10)))).astype(dtype)
    if db.ndim >= 4:
        mx = db.max(axis=(-3, -2, -1), keepdims=True)
    else:
        mx = db.max()
    db = np.maximum(db, mx - dtype(TOP_DB))
    out = (db - dtype(MINMAX_MIN)) / dtype(MINMAX_MAX - MINMAX_MIN) * dtype(2.0) - dtype(1.0)
    return out.astype(dtype)


# --------------------------------------------------------------------------------------
# TF32 operand emulation (test infrastructure for the tcgen05 kind::tf32 kernels)
# --------------------------------------------------------------------------------------
# The CUDA path multiplies on the tensor cores with TF32 operands (fp32 containers, 10-bit mantissa, fp32
# accumulation).  Its producers round every GEMM operand with ``cvt.rna.tf32.f32`` (nearest, ties away from zero).
# A product of two TF32 numbers is exact in fp32, so an fp32 matmul of operands rounded the same way reproduces the
# tensor-core result up to accumulation order.  Inside ``with tf32_emulation():`` the oracle rounds at exactly the
# places the kernels do (DESIGN.md "Precision"):
#   * every Linear: input, weight and - in the backward pass - the incoming gradient (dgrad and wgrad operands;
#     bias gradients are column sums of the rounded gradient),
#   * attention: Q, K, V (the qkv GEMM stores rounded values), the un-normalised probabilities exp(s - max) that
#     multiply V, the stored output O, and in the backward pass dO, P, dS and the stored dQ / dK / dV.
# Everything else (LayerNorm, softmax statistics, GELU, BatchNorm, loss, residual stream) stays fp32, as on the GPU.
#   * ``gelu_half=True`` (engine.half_dgelu(), ATST_FUSE_GELU bit 4): the student's MLP keeps gelu'(u) as fp16 for the
#     backward pass instead of the fp32 pre-activation u (gemm_epilogue.cuh EPI_GELU_H / EPI_DGELU_H); the forward
#     pass is unchanged.
# With emulation off nothing below changes the reference arithmetic.
_EMULATE_TF32 = False
_EMULATE_HEADS = True
_EMULATE_GELU_H = False


class tf32_emulation:
    """``heads=False`` leaves the projector / predictor Linears in fp32: the CUDA path runs those (< 0.1 % of the
    flops) as error-compensated 3xTF32 products (hi/lo operand split), i.e. fp32 to ~1e-6."""

    def __init__(self, on=True, heads=True, gelu_half=False):
        self.on, self.heads, self.gelu_half = on, heads, gelu_half

    def __enter__(self):
        global _EMULATE_TF32, _EMULATE_HEADS, _EMULATE_GELU_H
        self.prev = (_EMULATE_TF32, _EMULATE_HEADS, _EMULATE_GELU_H)
        _EMULATE_TF32, _EMULATE_HEADS, _EMULATE_GELU_H = self.on, self.heads, bool(self.on and self.gelu_half)

    def __exit__(self, *exc):
        global _EMULATE_TF32, _EMULATE_HEADS, _EMULATE_GELU_H
        _EMULATE_TF32, _EMULATE_HEADS, _EMULATE_GELU_H = self.prev


def rna_tf32(x):
    """cvt.rna.tf32.f32 on an fp32 tensor: keep 10 mantissa bits, round to nearest, ties away from zero."""
    i = x.detach().contiguous().view(torch.int32)
    return ((i + 0x1000) & ~0x1FFF).view(torch.float32)


class _LinearTF32(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, w, b):
        xr, wr = rna_tf32(x), rna_tf32(w)
        ctx.save_for_backward(xr, wr)
        ctx.has_bias = b is not None
        y = xr @ wr.t()
        return y + b if b is not None else y

    @staticmethod
    def backward(ctx, g):
        xr, wr = ctx.saved_tensors
        gr = rna_tf32(g)
        g2, x2 = gr.reshape(-1, gr.shape[-1]), xr.reshape(-1, xr.shape[-1])
        return gr @ wr, g2.t() @ x2, (g2.sum(0) if ctx.has_bias else None)


def linear(x, w, b=None):
    if _EMULATE_TF32:
        return _LinearTF32.apply(x, w, b)
    return F.linear(x, w, b)


class _GeluHalfGrad(torch.autograd.Function):
    """exact GELU (audiossl/modules/transformer.py:77-92, nn.GELU) whose backward multiplies by
    gelu'(u) = Phi(u) + u phi(u) rounded to fp16 - what EPI_GELU_H stores and EPI_DGELU_H reads."""

    @staticmethod
    def forward(ctx, u):
        cdf = 0.5 * (1.0 + torch.erf(u * 0.7071067811865476))
        pdf = 0.3989422804014327 * torch.exp(-0.5 * u * u)
        ctx.save_for_backward((cdf + u * pdf).to(torch.float16))
        return F.gelu(u)

    @staticmethod
    def backward(ctx, g):
        (gp,) = ctx.saved_tensors
        return g * gp.to(torch.float32)


def gelu(u):
    if _EMULATE_TF32 and _EMULATE_GELU_H and u.requires_grad:
        return _GeluHalfGrad.apply(u)
    return F.gelu(u)


class OLinear(nn.Linear):
    """Linear of the projector / predictor heads."""

    def forward(self, x):
        if _EMULATE_TF32 and not _EMULATE_HEADS:
            return F.linear(x, self.weight, self.bias)
        return linear(x, self.weight, self.bias)


class _AttentionTF32(torch.autograd.Function):
    """softmax(q k^T scale + key mask) v the way attention_tc.cu / attention_bwd_tc.cu compute it."""

    @staticmethod
    def forward(ctx, q, k, v, scale, valid):
        # q, k, v [B,H,N,d]; valid [B,1,1,N] bool (keys below the length)
        q, k, v = rna_tf32(q), rna_tf32(k), rna_tf32(v)
        s = q @ k.transpose(-2, -1)
        m = s.masked_fill(~valid, -float("inf")).amax(-1, keepdim=True)
        pu = torch.where(valid, torch.exp((s - m) * scale), torch.zeros(()))
        l = pu.sum(-1, keepdim=True)
        o = rna_tf32((rna_tf32(pu) @ v) * (1.0 / l))
        ctx.save_for_backward(q, k, v, o, m * scale + torch.log(l), valid)
        ctx.scale = scale
        return o

    @staticmethod
    def backward(ctx, d_o):
        q, k, v, o, lse, valid = ctx.saved_tensors
        d_o = rna_tf32(d_o)
        p = torch.where(valid, torch.exp((q @ k.transpose(-2, -1)) * ctx.scale - lse), torch.zeros(()))
        delta = (d_o * o).sum(-1, keepdim=True)
        ds = rna_tf32(p * (d_o @ v.transpose(-2, -1) - delta))
        dq = rna_tf32((ds @ k) * ctx.scale)
        dk = rna_tf32((ds.transpose(-2, -1) @ q) * ctx.scale)
        dv = rna_tf32(rna_tf32(p).transpose(-2, -1) @ d_o)
        return dq, dk, dv, None, None


def attention_core(q, k, v, scale, length):
    """[B,H,N,d] x3 -> [B,H,N,d]; length [B] = number of valid keys (None: all) - modules/transformer.py:111-118."""
    n_tok = q.shape[-2]
    if _EMULATE_TF32:
        if length is None:
            valid = torch.ones((q.shape[0], 1, 1, n_tok), dtype=torch.bool)
        else:
            # a length <= 0 masks every key with the same -10000: the softmax is that of the unmasked row
            ln = torch.where(length <= 0, torch.full_like(length, n_tok), length)
            valid = (torch.arange(n_tok)[None, :] < ln[:, None])[:, None, None, :]
        return _AttentionTF32.apply(q, k, v, scale, valid)
    att = (q @ k.transpose(-2, -1)) * scale
    if length is not None:
        att = att + attention_mask(n_tok, length)
    return att.softmax(dim=-1) @ v


# --------------------------------------------------------------------------------------
# transformer encoder (torch CPU fp32)
# --------------------------------------------------------------------------------------
def _trunc_normal_(t, std=0.02):
    return nn.init.trunc_normal_(t, mean=0.0, std=std, a=-2.0, b=2.0)


def attention_mask(n_tok, length):
    """-10000 on keys >= length, [B,1,1,N] broadcastable (modules/transformer.py:152-159)."""
    m = torch.arange(n_tok, device=length.device)[None, :] >= length[:, None]
    return (-10000.0 * m[:, None, None, :]).to(torch.float32)


class OracleBlock(nn.Module):
    def __init__(self, dim, heads, mlp_ratio=4.0):
        super().__init__()
        self.norm1 = nn.LayerNorm(dim, eps=1e-6)
        self.attn = nn.Module()
        self.attn.qkv = nn.Linear(dim, dim * 3, bias=False)
        self.attn.proj = nn.Linear(dim, dim)
        self.norm2 = nn.LayerNorm(dim, eps=1e-6)
        self.mlp = nn.Module()
        self.mlp.fc1 = nn.Linear(dim, int(dim * mlp_ratio))
        self.mlp.fc2 = nn.Linear(int(dim * mlp_ratio), dim)
        self.heads = heads

    def forward(self, x, length, dp_scale=None):
        """dp_scale: optional ([B] attn-branch scale, [B] mlp-branch scale) = mask/keep_prob."""
        B, N, C = x.shape
        h = self.norm1(x)
        qkv = linear(h, self.attn.qkv.weight).reshape(B, N, 3, self.heads, C // self.heads).permute(2, 0, 3, 1, 4)
        y = attention_core(qkv[0], qkv[1], qkv[2], (C // self.heads) ** -0.5, length)
        y = y.transpose(1, 2).reshape(B, N, C)
        y = linear(y, self.attn.proj.weight, self.attn.proj.bias)
        if dp_scale is not None:
            y = y * dp_scale[0][:, None, None]
        x = x + y
        z = linear(gelu(linear(self.norm2(x), self.mlp.fc1.weight, self.mlp.fc1.bias)), self.mlp.fc2.weight,
                   self.mlp.fc2.bias)
        if dp_scale is not None:
            z = z * dp_scale[1][:, None, None]
        return x + z


class OracleAST(nn.Module):
    """AST with PatchEmbed_v2 (64x4 patches), CLS token, "cut" positional embedding."""

    def __init__(self, embed_dim=768, depth=12, num_heads=12, spec_h=64, spec_w=1001,
                 patch_h=64, patch_w=4, use_cls=True, norm_name="norm"):
        super().__init__()
        self.embed_dim, self.patch_h, self.patch_w, self.use_cls = embed_dim, patch_h, patch_w, use_cls
        self.patch_embed = nn.Module()
        self.patch_embed.patch_embed = nn.Linear(patch_h * patch_w, embed_dim)
        self.mask_embed = nn.Parameter(torch.zeros(1, 1, embed_dim))
        if use_cls:
            self.cls_token = nn.Parameter(torch.zeros(1, 1, embed_dim))
        n_patches = (spec_h // patch_h) * (spec_w // patch_w)
        self.pos_embed = nn.Parameter(torch.zeros(1, n_patches + 1, embed_dim))
        self.blocks = nn.ModuleList([OracleBlock(embed_dim, num_heads) for _ in range(depth)])
        self.norm_name = norm_name
        setattr(self, norm_name, nn.LayerNorm(embed_dim, eps=1e-6))
        _trunc_normal_(self.pos_embed)
        _trunc_normal_(self.mask_embed)
        if use_cls:
            _trunc_normal_(self.cls_token)
        for m in self.modules():
            if isinstance(m, nn.Linear):
                _trunc_normal_(m.weight)
                if m.bias is not None:
                    nn.init.zeros_(m.bias)

    def patchify(self, mel):
        B, c, H, W = mel.shape
        H, W = H - H % self.patch_h, W - W % self.patch_w
        x = mel[:, :, :H, :W]
        # 'b c (h p1) (w p2) -> b (w h) (p1 p2 c)'
        x = x.reshape(B, c, H // self.patch_h, self.patch_h, W // self.patch_w, self.patch_w)
        x = x.permute(0, 4, 2, 3, 5, 1).reshape(B, (W // self.patch_w) * (H // self.patch_h), -1)
        return x

    def tokens(self, mel, length, mask_index=None, mask=True):
        pe = self.patch_embed.patch_embed
        x = linear(self.patchify(mel), pe.weight, pe.bias)
        B, T, C = x.shape
        plen = None
        if length is not None:
            plen = (mel.shape[2] // self.patch_h) * ((length - length % self.patch_w) // self.patch_w)
        if mask_index is not None and mask:
            m = mask_index.unsqueeze(2).float()
            x = (1 - m) * x + m * self.mask_embed
        if self.use_cls:
            x = torch.cat([self.cls_token.expand(B, -1, -1), x], dim=1)
            x = x + self.pos_embed[:, :T + 1]
        else:  # frame model: positions start at 1 (atstframe/audio_transformer.py:161-181)
            x = x + self.pos_embed[:, 1:T + 1]
        return x, plen

    def forward(self, mel, length=None, dp_scales=None, mask_index=None, mask_input=True):
        x, plen = self.tokens(mel, length, mask_index, mask_input)
        for i, blk in enumerate(self.blocks):
            key_len = None if plen is None else (plen + 1 if self.use_cls else plen)
            x = blk(x, key_len, None if dp_scales is None else dp_scales[i])
        x = getattr(self, self.norm_name)(x)
        if self.use_cls:
            return x[:, 0]
        # frame model returns masked frames inside the valid length (audio_transformer.py:183-207)
        lm = torch.arange(x.shape[1], device=x.device)[None, :] < plen[:, None]
        return x[mask_index.bool() & lm]


def oracle_intermediate(enc, mel, length, n=1):
    """final-norm token outputs of the last n blocks (inference; models/atst/audio_transformer.py:235-256)."""
    x, plen = enc.tokens(mel, length, None, False)
    outs = []
    for i, blk in enumerate(enc.blocks):
        x = blk(x, plen + 1 if enc.use_cls else plen)
        if len(enc.blocks) - i <= n:
            outs.append(getattr(enc, enc.norm_name)(x))
    return outs, plen


def oracle_intermediate_chunks(enc, mel, length, n=1, chunk_len=601):
    """models/atst/audio_transformer.py:257-353 (avgpool=True): chunked CLS + masked-mean pooling."""
    total = mel.shape[-1]
    cls, avg, marks = [], [], []
    for i in range(total // chunk_len + 1):
        cur = torch.clip(length - i * chunk_len, 0)
        marks.append(cur > 0 if i == 0 else cur > chunk_len // 2)
        xc = mel[..., i * chunk_len:min((i + 1) * chunk_len, total)]
        outs, plen = oracle_intermediate(enc, xc, cur, n)
        lm = (torch.arange(outs[0].shape[1] - 1)[None, :] < plen[:, None]).unsqueeze(-1)
        cls.append([o[:, 0] for o in outs])
        avg.append([(o[:, 1:] * lm).sum(1) / (plen[:, None] + 1e-6) for o in outs])
    mark = torch.stack(marks).unsqueeze(-1).float()
    co = [(torch.stack(list(c)) * mark).sum(0) / mark.sum(0) for c in zip(*cls)]
    ao = [(torch.stack(list(a)) * mark).sum(0) / mark.sum(0) for a in zip(*avg)]
    return torch.cat(co + ao, dim=-1)


def oracle_resize_crop(lms, rect, virtual_crop_scale=(1.0, 1.5)):
    """RandomResizeCrop.forward with the random draws given (transforms/byol_a.py:34-49): lms [C,H,W]."""
    c, h, w = lms.shape
    ch, cw = int(h * virtual_crop_scale[0]), int(w * virtual_crop_scale[1])
    canvas = torch.zeros((c, ch, cw))
    x0, y0 = (cw - w) // 2, (ch - h) // 2
    canvas[:, y0:y0 + h, x0:x0 + w] = lms
    i, j, hh, ww = rect
    crop = canvas[:, i:i + hh, j:j + ww]
    return F.interpolate(crop.unsqueeze(0), size=(h, w), mode="bicubic", align_corners=True).squeeze(0)


def oracle_log_mixup_exp(x, z, alpha):
    """log((1-alpha) e^x + alpha e^z + eps) (transforms/byol_a.py:61-82 with equal lengths; Mixup passes 1-alpha)."""
    return torch.log((1.0 - alpha) * x.exp() + alpha * z.exp() + torch.finfo(x.dtype).eps)


def build_mlp(in_dim, hidden, out_dim):
    return nn.Sequential(OLinear(in_dim, hidden, bias=False), nn.BatchNorm1d(hidden),
                         nn.ReLU(inplace=True), OLinear(hidden, out_dim, bias=False))


class OracleMultiCrop(nn.Module):
    def __init__(self, encoder, embed_dim, predictor=True):
        super().__init__()
        self.encoder = encoder
        self.projector = build_mlp(embed_dim, 4096, 256)
        self.predictor = build_mlp(256, 4096, 256) if predictor else nn.Identity()

    def forward(self, crops, lengths, dp_scales=None):
        """crops: list of [B,1,64,T_i]; consecutive equal-width crops share one encoder call.
        dp_scales: optional list (one per encoder call) of per-block (attn, mlp) scale pairs."""
        widths = [c.shape[-1] for c in crops]
        groups, start = [], 0
        for i in range(1, len(crops) + 1):
            if i == len(crops) or widths[i] != widths[start]:
                groups.append((start, i))
                start = i
        outs = []
        for gi, (s, e) in enumerate(groups):
            outs.append(self.encoder(torch.cat(crops[s:e]), torch.cat(lengths[s:e]),
                                     None if dp_scales is None else dp_scales[gi]))
        out = torch.cat(outs)
        return self.predictor(self.projector(out))


def compute_std(y):
    """single-rank compute_var (byol.py:42-53): unbiased per-dim std of rows, +1e-6 inside sqrt."""
    y = y.reshape(-1, y.shape[-1])
    n = float(y.shape[0])
    zs, zss = y.sum(0), (y ** 2).sum(0)
    var = zss / (n - 1) - zs ** 2 / (n * (n - 1))
    return torch.sqrt(var + 1e-6)


def byol_loss(student, teacher, ncrops):
    """returns (loss, std_student, std_teacher) exactly as ByolLoss.forward (byol.py:57-78)."""
    std_s = compute_std(F.normalize(student, dim=-1)).mean()
    std_t = compute_std(F.normalize(teacher, dim=-1)).mean()
    s_chunks = student.chunk(ncrops)
    t_chunks = teacher.detach().chunk(2)
    total, n_terms = 0.0, 0
    for iq, q in enumerate(t_chunks):
        for iv, v in enumerate(s_chunks):
            if iq == iv:
                continue
            p = F.normalize(q, dim=-1)
            z = F.normalize(v, dim=-1)
            total = total + (2 - 2 * (p * z).sum(dim=1).mean())
            n_terms += 1
    return total / n_terms, std_s, std_t


class OracleATST(nn.Module):
    CFG = {"small": (384, 12, 6), "base": (768, 12, 12), "large": (1024, 24, 16)}

    def __init__(self, arch="small", ncrops=2, embed_dim=None, depth=None, num_heads=None):
        super().__init__()
        if embed_dim is None:
            embed_dim, depth, num_heads = self.CFG[arch]
        self.ncrops = ncrops
        self.student = OracleMultiCrop(OracleAST(embed_dim, depth, num_heads), embed_dim, True)
        self.teacher = OracleMultiCrop(OracleAST(embed_dim, depth, num_heads), embed_dim, False)
        for p in self.teacher.parameters():
            p.requires_grad = False
        self.teacher.load_state_dict({k: v for k, v in self.student.state_dict().items()
                                      if "predictor" not in k})

    def forward(self, crops, lengths, dp_student=None, dp_teacher=None):
        t = self.teacher(crops[:2], lengths[:2], dp_teacher)
        s = self.student(crops, lengths, dp_student)
        return byol_loss(s, t, self.ncrops)

    @torch.no_grad()
    def update_teacher(self, m):
        ema_update(self, m)


class OracleFrameATST(nn.Module):
    """ATST-Frame, symmetric branch: audiossl/methods/atstframe/model.py:24-76, byol.py:57-84,118-138,
    audio_transformer.py:161-207 (no CLS, mask_embed blend for the student, masked valid frames returned)."""

    def __init__(self, arch="small", embed_dim=None, depth=None, num_heads=None):
        super().__init__()
        if embed_dim is None:
            embed_dim, depth, num_heads = OracleATST.CFG[arch]
        mk = lambda: OracleAST(embed_dim, depth, num_heads, use_cls=False, norm_name="norm_frame")
        self.student = OracleMultiCrop(mk(), embed_dim, True)
        self.teacher = OracleMultiCrop(mk(), embed_dim, False)
        for p in self.teacher.parameters():
            p.requires_grad = False
        self.teacher.load_state_dict({k: v for k, v in self.student.state_dict().items()
                                      if "predictor" not in k})

    @staticmethod
    def _net(net, crops, lengths, masks, mask_input):
        frames = net.encoder(torch.cat(crops), torch.cat(lengths), None, torch.cat(masks), mask_input)
        return net.predictor(net.projector(frames))

    def forward(self, crops, lengths, masks):
        t = self._net(self.teacher, crops, lengths, masks, False)
        s = self._net(self.student, crops, lengths, masks, True)
        return byol_loss(s, t, 2)

    @torch.no_grad()
    def update_teacher(self, m):
        ema_update(self, m)


@torch.no_grad()
def ema_update(model, m):
    for q, k in zip(model.student.encoder.parameters(), model.teacher.encoder.parameters()):
        k.mul_(m).add_((1 - m) * q.detach())
    for q, k in zip(model.student.projector.parameters(), model.teacher.projector.parameters()):
        k.mul_(m).add_((1 - m) * q.detach())


# --------------------------------------------------------------------------------------
# schedules / optimiser
# --------------------------------------------------------------------------------------
def cosine_scheduler_step(base_value, final_value, max_steps, warmup_steps=0, start_warmup_value=0):
    warm = np.array([])
    if warmup_steps > 0:
        warm = np.linspace(start_warmup_value, base_value, warmup_steps)
    iters = np.arange(max_steps - warmup_steps)
    sched = final_value + 0.5 * (base_value - final_value) * (1 + np.cos(np.pi * iters / len(iters)))
    sched = np.concatenate((warm, sched))
    assert len(sched) == max_steps
    return sched


def param_groups(module):
    """(regularised, not_regularised) parameter name lists (utils/common.py:41-68)."""
    reg, noreg = [], []
    for name, p in module.named_parameters():
        if not p.requires_grad:
            continue
        (noreg if (name.endswith(".bias") or p.ndim == 1) else reg).append(name)
    return reg, noreg


@torch.no_grad()
def hf_adamw_step(p, g, m, v, step, lr, wd, beta1=0.9, beta2=0.999, eps=1e-6):
    """transformers-4.x AdamW: bias-corrected step size, eps added to sqrt(v) (uncorrected),
    decoupled weight decay applied AFTER the Adam update using the same lr."""
    m.mul_(beta1).add_(g, alpha=1.0 - beta1)
    v.mul_(beta2).addcmul_(g, g, value=1.0 - beta2)
    denom = v.sqrt().add_(eps)
    step_size = lr * math.sqrt(1.0 - beta2 ** step) / (1.0 - beta1 ** step)
    p.addcdiv_(m, denom, value=-step_size)
    if wd > 0.0:
        p.add_(p, alpha=-lr * wd)
