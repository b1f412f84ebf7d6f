"""Explicit forward / backward drivers for the AST encoder and the projector / predictor heads.

No torch autograd inside: every step is a launch of a hand-written sm_100a kernel through the C ABI
(audiossl_b200.ops).  Activations needed by the backward pass are kept in a shape-keyed workspace that is
reused from step to step, so a training step performs no device allocation after the first one.

Reference semantics (file:line under /root/reference):
  encoder forward  audiossl/models/atst/audio_transformer.py:153-210 (prepare_tokens, block loop, norm, CLS)
  block            audiossl/modules/transformer.py:136-150 (pre-LN, DropPath on both branches)
  frame encoder    audiossl/methods/atstframe/audio_transformer.py:161-207
  heads            audiossl/models/atst/byol.py:6-22 (Linear -> BatchNorm1d -> ReLU -> Linear, no biases)
"""
import torch

import os

from . import ops

# GELU / GELU' inside the GEMM epilogue or as separate elementwise passes; see DESIGN.md "GEMM epilogues".
#   ATST_FUSE_GELU bit 0: forward passes that keep no activations (teacher, inference) fuse GELU into fc1 and never
#                         store the pre-activation; bit 1: the student forward fuses it too (epilogue writes u and g);
#                  bit 2: the backward fuses GELU' into the fc2 dgrad epilogue; bit 3: that epilogue also takes the
#                         fc1 bias gradient (column sums) instead of a separate pass over du - measured slower
#                         (the butterflies + red.global.add cost the dgrad GEMM ~0.9 ms to save a 0.27 ms pass): off
#                  bit 4: (with bits 1 and 2) what the student keeps for the backward pass is gelu'(u) as fp16 - the same
#                         10-bit mantissa the TF32 rounding of du leaves anyway - instead of the fp32 pre-activation u:
#                         the second stream of the fc1 and fc2-dgrad epilogues is half the bytes and the dgrad epilogue
#                         has no GELU math left.  Never in the 3xTF32 validation build.
_FG = int(os.environ.get("ATST_FUSE_GELU", "23"))
FUSE_GELU_NOSAVE, FUSE_GELU, FUSE_DGELU, FUSE_COLSUM = bool(_FG & 1), bool(_FG & 2), bool(_FG & 4), bool(_FG & 8)
HALF_DGELU = bool(_FG & 16) and FUSE_GELU and FUSE_DGELU


def half_dgelu(hidden=None):
    """whether the student's MLP keeps fp16 gelu'(u) (the parity tests configure their emulation with this).
    `hidden` (the MLP width 4 D): the fp16 epilogues exist in the CTA-pair GEMM only, i.e. for widths that are a
    multiple of 256 or above 1024 (and a multiple of 16); other widths keep the fp32 pre-activation."""
    from . import _lib
    if not HALF_DGELU or _lib.is_precise():
        return False
    return hidden is None or ((hidden % 256 == 0 or hidden > 1024) and hidden % 32 == 0)
# The projector / predictor heads (< 0.1 % of the flops) run as error-compensated 3xTF32 products on unrounded fp32
# operands: their train-mode BatchNorm over a few hundred rows doubles whatever rounding error enters it, and the BYOL
# gradient behind it is the ill-conditioned part of the step.  ATST_HEADS_3XTF32=0: plain TF32.
HEADS_3X = os.environ.get("ATST_HEADS_3XTF32", "1") != "0"


class Workspace:
    """shape-keyed cache of device tensors (activations, gradients, scratch)."""

    def __init__(self, device):
        self.device = device
        self.bufs = {}

    def get(self, tag, shape, dtype=torch.float32):
        key = (tag, tuple(shape), dtype)
        t = self.bufs.get(key)
        if t is None:
            t = torch.empty(shape, device=self.device, dtype=dtype)
            self.bufs[key] = t
        return t

    def nbytes(self):
        return sum(t.numel() * t.element_size() for t in self.bufs.values())


_KEEP_CACHE = {}


def droppath_scales(depth, drop_path_rate, S, device, generator=None):
    """per-block (attn, mlp) DropPath scales floor(keep + U[0,1)) / keep for S sequences
    (modules/transformer.py:48-57; rates linspace(0, rate, depth), audio_transformer.py:107).
    Blocks with rate 0 get None (nn.Identity in the reference: no random draw)."""
    rates = torch.linspace(0, drop_path_rate, depth).tolist()
    live = [i for i, r in enumerate(rates) if r != 0.0]
    out = [None] * depth
    if not live:
        return out
    # one draw and one floor / divide for all blocks (four launches per encoder call instead of four per block)
    key = (depth, float(drop_path_rate), str(device))
    keep = _KEEP_CACHE.get(key)
    if keep is None:
        keep = _KEEP_CACHE[key] = torch.tensor([1.0 - rates[i] for i in live], device=device).view(-1, 1, 1)
    u = torch.rand((len(live), 2, S), device=device, generator=generator)
    s = torch.floor(keep + u) / keep
    for j, i in enumerate(live):
        out[i] = (s[j, 0], s[j, 1])
    return out


class EncoderEngine:
    def __init__(self, embed_dim, depth, num_heads, use_cls=True, norm_name="norm", prefix="encoder.",
                 patch_w=4, max_frames=1001):
        assert embed_dim == num_heads * 64, "head_dim must be 64 (all reference configs)"
        self.D, self.depth, self.H = embed_dim, depth, num_heads
        self.use_cls, self.norm_name, self.px = use_cls, norm_name, prefix
        self.patch_w, self.max_frames = patch_w, max_frames
        # test hook (tests/linkwise.py): a list that collects (name, tag, layer, tensor clone) at every block boundary
        # of the forward and backward passes, so each link of the chain can be checked against the CPU reference from the
        # GPU's own inputs to that link
        self.debug = None

    # ------------------------------------------------------------------ forward
    def forward(self, fp, ws, mel, lengths, dp=None, save=True, tag="s", mask=None, mask_input=True, collect=0,
                round_final=True):
        """mel [S,1,64,T] fp32 cuda contiguous; lengths [S] (valid frames) or None.
        Returns (out [S,D] tf32-rounded final-norm CLS rows, ctx) for the clip model, or
        (x_norm [S*N, D], ctx) for the frame model (row selection is done by the caller).
        collect=n (inference, get_intermediate_layers): ctx["collected"] = final-norm of every token after each of
        the last n blocks, n tensors [S*N, D] (fp32, not tf32-rounded)."""
        px, D, H = self.px, self.D, self.H
        S, _, Hm, T = mel.shape
        if T > self.max_frames:
            raise ValueError("clip of %d frames exceeds the positional-embedding capacity %d "
                             "(reference: 10 s max, chunk longer audio)" % (T, self.max_frames))
        P = T // self.patch_w
        N = P + (1 if self.use_cls else 0)
        M = S * N
        key_len = None
        if lengths is not None:
            plen = (lengths - lengths % self.patch_w) // self.patch_w
            key_len = (plen + (1 if self.use_cls else 0)).to(torch.int32).contiguous()
        t = (lambda name, shape: ws.get(tag + "/" + name, shape))
        patches = ops.patchify(mel, out=t("patches", (S * P, 256)))
        pe = ops.gemm_nt(patches, fp.c(px + "patch_embed.patch_embed.weight"),
                         bias=fp.p(px + "patch_embed.patch_embed.bias"), out=t("pe", (S * P, D)))
        m8 = None
        if mask is not None:
            m8 = mask.to(torch.uint8).contiguous()
        x = ops.tokens_fwd(pe, fp.p(px + "cls_token") if self.use_cls else None, fp.p(px + "pos_embed"), S, P, D,
                           use_cls=self.use_cls, mask_embed=fp.p(px + "mask_embed") if (m8 is not None and mask_input) else None,
                           mask=m8 if mask_input else None, out=t("x0", (M, D)))
        ctx = {"S": S, "P": P, "N": N, "M": M, "key_len": key_len, "dp": dp, "tag": tag, "layers": [],
               "patches": patches, "mask": m8 if mask_input else None}
        dbg = (lambda name, i, t_: self.debug.append((name, tag, i, t_.clone()))) if self.debug is not None else (lambda *a: None)
        for i in range(self.depth):
            b = "%sblocks.%d." % (px, i)
            dbg("x_in", i, x)
            lt = (lambda name, shape, dtype=torch.float32, i=i:
                  ws.get("%s/L%d/%s" % (tag, i if save else 0, name), shape, dtype))
            h, mean1, rstd1 = self._ln(x, fp.p(b + "norm1.weight"), fp.p(b + "norm1.bias"), M, lt("h", (M, D)),
                                       lt("mean1", (M,)), lt("rstd1", (M,)))
            qkv = ops.gemm_nt(h, fp.c(b + "attn.qkv.weight"), round_out=True, out=lt("qkv", (M, 3 * D)))
            o, lse = ops.attention_fwd(qkv, S, N, H, key_len, out=lt("o", (M, D)), lse=lt("lse", (S, H, N)))
            s_attn = s_mlp = None
            if dp is not None and dp[i] is not None:
                s_attn, s_mlp = dp[i]
            x1 = ops.gemm_nt(o, fp.c(b + "attn.proj.weight"), bias=fp.p(b + "attn.proj.bias"), epi=ops.EPI_RESID,
                             resid=x, rowscale=s_attn, rows_per_seq=N, out=lt("x1", (M, D)))
            h2, mean2, rstd2 = self._ln(x1, fp.p(b + "norm2.weight"), fp.p(b + "norm2.bias"), M, lt("h2", (M, D)),
                                        lt("mean2", (M,)), lt("rstd2", (M,)))
            u = None
            if not save and FUSE_GELU_NOSAVE:
                g = ops.gemm_nt(h2, fp.c(b + "mlp.fc1.weight"), bias=fp.p(b + "mlp.fc1.bias"), epi=ops.EPI_GELU,
                                aux=None, round_out=True, out=lt("g", (M, 4 * D)))
            elif save and FUSE_GELU and half_dgelu(4 * D):
                u = lt("gp", (M, 4 * D), torch.float16)   # gelu'(pre-activation), not the pre-activation
                g = ops.gemm_nt(h2, fp.c(b + "mlp.fc1.weight"), bias=fp.p(b + "mlp.fc1.bias"), epi=ops.EPI_GELU_H,
                                aux=u, round_out=True, out=lt("g", (M, 4 * D)))
            elif save and FUSE_GELU:
                u = lt("u", (M, 4 * D))
                g = ops.gemm_nt(h2, fp.c(b + "mlp.fc1.weight"), bias=fp.p(b + "mlp.fc1.bias"), epi=ops.EPI_GELU,
                                aux=u, round_out=True, out=lt("g", (M, 4 * D)))
            else:
                u = lt("u", (M, 4 * D))
                ops.gemm_nt(h2, fp.c(b + "mlp.fc1.weight"), bias=fp.p(b + "mlp.fc1.bias"), out=u)
                g = ops.gelu_fwd(u, lt("g", (M, 4 * D)))
            x2 = ops.gemm_nt(g, fp.c(b + "mlp.fc2.weight"), bias=fp.p(b + "mlp.fc2.bias"), epi=ops.EPI_RESID,
                             resid=x1, rowscale=s_mlp, rows_per_seq=N, out=lt("x2", (M, D)))
            if save:
                ctx["layers"].append(dict(x=x, h=h, mean1=mean1, rstd1=rstd1, qkv=qkv, o=o, lse=lse, x1=x1, h2=h2,
                                          mean2=mean2, rstd2=rstd2, u=u, g=g))
            x = x2
            if collect and self.depth - i <= collect:
                nmw = px + self.norm_name
                j = collect - (self.depth - i)
                yn, _, _ = ops.layernorm_fwd(x, fp.p(nmw + ".weight"), fp.p(nmw + ".bias"), M, D, round_out=False,
                                             out=ws.get("%s/collect%d" % (tag, j), (M, D)))
                ctx.setdefault("collected", []).append(yn)
        ctx["x_final"] = x
        nm = px + self.norm_name
        if self.use_cls:
            out = t("cls_out", (S, D))
            mean = t("meanf", (S,))
            rstd = t("rstdf", (S,))
            ops.layernorm_fwd(x, fp.p(nm + ".weight"), fp.p(nm + ".bias"), S, D, x_stride=N * D, out=out, mean=mean,
                              rstd=rstd, round_out=round_final and not HEADS_3X)
        else:
            out = t("xn", (M, D))
            mean = t("meanf", (M,))
            rstd = t("rstdf", (M,))
            ops.layernorm_fwd(x, fp.p(nm + ".weight"), fp.p(nm + ".bias"), M, D, out=out, mean=mean, rstd=rstd,
                              round_out=round_final and not HEADS_3X)
        ctx["meanf"], ctx["rstdf"] = mean, rstd
        dbg("x_in", self.depth, x)
        dbg("enc_out", self.depth, out)
        return out, ctx

    def _ln(self, x, g, b, rows, out, mean, rstd):
        return ops.layernorm_fwd(x, g, b, rows, self.D, out=out, mean=mean, rstd=rstd)

    # ------------------------------------------------------------------ backward
    def backward(self, fp, ws, ctx, d_out, on_block_done=None):
        """d_out: gradient wrt forward()'s output ([S,D] clip model / [S*N,D] frame model).
        Accumulates parameter gradients into fp.grad.  on_block_done(i) is called once every kernel that writes the
        weight gradients of block i has been enqueued (the data-parallel exchange of that range can start).

        Every LayerNorm-backward launch also emits, for the branch that consumes its result, the GEMM-ready copy
        dys = tf32(droppath_scale * dx) and that branch's bias gradient (column sums of dys), so the DropPath
        backward and the proj / fc2 bias gradients cost no extra pass; the fc1 bias gradient rides on the GELU'
        pass."""
        px, D, H = self.px, self.D, self.H
        S, P, N, M, tag = ctx["S"], ctx["P"], ctx["N"], ctx["M"], ctx["tag"]
        t = (lambda name, shape: ws.get(tag + "/bwd/" + name, shape))
        nm = px + self.norm_name
        dp = ctx["dp"]
        scales = [(None, None) if (dp is None or dp[i] is None) else dp[i] for i in range(self.depth)]
        blk = (lambda i: "%sblocks.%d." % (px, i))
        dxa, dxb, dys = t("dxa", (M, D)), t("dxb", (M, D)), t("dys", (M, D))
        last = self.depth - 1
        if self.use_cls:
            dxa.zero_()
            dys.zero_()
            ops.layernorm_bwd(d_out, ctx["x_final"], ctx["meanf"], ctx["rstdf"], fp.p(nm + ".weight"),
                              fp.g(nm + ".weight"), fp.g(nm + ".bias"), S, D, dx=dxa, x_stride=N * D, dx_stride=N * D,
                              dys=dys, dys_stride=N * D, rowscale=scales[last][1], rows_per_seq=1,
                              colsum_out=fp.g(blk(last) + "mlp.fc2.bias"))
        else:
            ops.layernorm_bwd(d_out, ctx["x_final"], ctx["meanf"], ctx["rstdf"], fp.p(nm + ".weight"),
                              fp.g(nm + ".weight"), fp.g(nm + ".bias"), M, D, dx=dxa, dys=dys,
                              rowscale=scales[last][1], rows_per_seq=N, colsum_out=fp.g(blk(last) + "mlp.fc2.bias"))
        dx, other = dxa, dxb
        dbg = (lambda name, i, t_: self.debug.append((name, tag, i, t_.clone()))) if self.debug is not None else (lambda *a: None)
        dbg("d_enc_out", self.depth, d_out)
        dbg("dx_in", self.depth, dx)
        for i in reversed(range(self.depth)):
            b = blk(i)
            L = ctx["layers"][i]
            s_attn = scales[i][0]
            # ---- MLP branch: x2 = x1 + s * (g W2^T + b2); dys = tf32(s * dx), fc2.bias gradient already accumulated
            ops.gemm_tn_acc(dys, L["g"], fp.g(b + "mlp.fc2.weight"))
            if FUSE_DGELU:
                du = ops.gemm_nn(dys, fp.c(b + "mlp.fc2.weight"),
                                 epi=ops.EPI_DGELU_H if L["u"].dtype == torch.float16 else ops.EPI_DGELU, aux=L["u"], round_out=True,
                                 out=t("du", (M, 4 * D)), colsum_out=fp.g(b + "mlp.fc1.bias") if FUSE_COLSUM else None)
                if not FUSE_COLSUM:
                    ops.colsum_acc(du, fp.g(b + "mlp.fc1.bias"))
            else:
                du = ops.gemm_nn(dys, fp.c(b + "mlp.fc2.weight"), out=t("du", (M, 4 * D)))
                ops.gelu_bwd_(du, L["u"], colsum_out=fp.g(b + "mlp.fc1.bias"))
            dbg("du", i, du)
            ops.gemm_tn_acc(du, L["h2"], fp.g(b + "mlp.fc1.weight"))
            dh2 = ops.gemm_nn(du, fp.c(b + "mlp.fc1.weight"), out=t("dh", (M, D)))
            dx1 = ops.layernorm_bwd(dh2, L["x1"], L["mean2"], L["rstd2"], fp.p(b + "norm2.weight"),
                                    fp.g(b + "norm2.weight"), fp.g(b + "norm2.bias"), M, D, dres=dx, dx=other,
                                    dys=dys, rowscale=s_attn, rows_per_seq=N, colsum_out=fp.g(b + "attn.proj.bias"))
            dbg("dx1", i, dx1)
            # ---- attention branch: x1 = x + s * (o Wp^T + bp)
            ops.gemm_tn_acc(dys, L["o"], fp.g(b + "attn.proj.weight"))
            d_o = ops.gemm_nn(dys, fp.c(b + "attn.proj.weight"), round_out=True, out=t("d_o", (M, D)))
            dqkv = ops.attention_bwd(L["qkv"], L["o"], d_o, L["lse"], S, N, H, ctx["key_len"],
                                     dqkv=t("dqkv", (M, 3 * D)), delta_ws=t("delta", (S, H, N)))
            dbg("dqkv", i, dqkv)
            ops.gemm_tn_acc(dqkv, L["h"], fp.g(b + "attn.qkv.weight"))
            dh = ops.gemm_nn(dqkv, fp.c(b + "attn.qkv.weight"), out=t("dh", (M, D)))
            if i > 0:  # the block input feeds block i-1's MLP branch
                ops.layernorm_bwd(dh, L["x"], L["mean1"], L["rstd1"], fp.p(b + "norm1.weight"),
                                  fp.g(b + "norm1.weight"), fp.g(b + "norm1.bias"), M, D, dres=dx1, dx=dx, dys=dys,
                                  rowscale=scales[i - 1][1], rows_per_seq=N,
                                  colsum_out=fp.g(blk(i - 1) + "mlp.fc2.bias"))
            else:
                ops.layernorm_bwd(dh, L["x"], L["mean1"], L["rstd1"], fp.p(b + "norm1.weight"),
                                  fp.g(b + "norm1.weight"), fp.g(b + "norm1.bias"), M, D, dres=dx1, dx=dx)
            dbg("dx_in", i, dx)
            if on_block_done is not None:
                on_block_done(i)
        dpe = t("dpe", (S * P, D))
        ops.tokens_bwd(dx, dpe, fp.g(px + "pos_embed"), fp.g(px + "cls_token") if self.use_cls else None, S, P, D,
                       use_cls=self.use_cls, mask=ctx["mask"],
                       dmask_embed=fp.g(px + "mask_embed") if ctx["mask"] is not None else None)
        ops.colsum_acc(dpe, fp.g(px + "patch_embed.patch_embed.bias"))
        ops.gemm_tn_acc(dpe, ctx["patches"], fp.g(px + "patch_embed.patch_embed.weight"))


class HeadEngine:
    """Linear(in,4096,no bias) -> BatchNorm1d(4096, train) -> ReLU -> Linear(4096,256,no bias)."""

    def __init__(self, prefix, in_dim, hidden=4096, out_dim=256):
        self.px, self.in_dim, self.hidden, self.out_dim = prefix, in_dim, hidden, out_dim

    def forward(self, fp, ws, x, bn_buffers, tag, round_out, stats_sync=None, momentum=0.1, eps=1e-5):
        """x [R, in] (tf32-rounded).  bn_buffers: (running_mean, running_var, num_batches_tracked) or None.
        stats_sync(mean, m2, n) -> (mean, m2, n_total): SyncBatchNorm statistics exchange (DDP)."""
        head = self.forward_stats(fp, ws, x, tag)
        if stats_sync is not None:
            head["mean"], head["m2"], head["n"] = stats_sync(head["mean"], head["m2"], head["n"])
        return self.forward_finish(fp, ws, head, bn_buffers, round_out, momentum, eps)

    def forward_stats(self, fp, ws, x, tag):
        """first Linear + this rank's batch statistics (the part before the SyncBatchNorm exchange)."""
        px = self.px
        R = x.shape[0]
        z1 = ops.gemm_nt(x, self._w(fp, "0.weight"), out=ws.get(tag + "/" + px + "z1", (R, self.hidden)),
                         precise=HEADS_3X)
        mean, m2 = ops.bn_stats(z1)
        return dict(x=x, z1=z1, mean=mean, m2=m2, n=float(R), R=R, tag=tag)

    def _w(self, fp, name):
        """the weight operand: the fp32 master for the 3xTF32 heads, the TF32 compute copy otherwise"""
        return fp.p(self.px + name) if HEADS_3X else fp.c(self.px + name)

    def forward_finish(self, fp, ws, head, bn_buffers, round_out, momentum=0.1, eps=1e-5):
        """running statistics, normalise + ReLU, second Linear - from the (global) statistics in ``head``."""
        px, tag, R = self.px, head["tag"], head["R"]
        t = (lambda name, shape: ws.get(tag + "/" + px + name, shape))
        rm = rv = None
        if bn_buffers is not None:
            rm, rv, nbt = bn_buffers
            nbt += 1
        rstd = ops.bn_finalize(head["mean"], head["m2"], head["n"], rm, rv, eps=eps, momentum=momentum)
        a1 = ops.bn_relu_fwd(head["z1"], head["mean"], rstd, fp.p(px + "1.weight"), fp.p(px + "1.bias"),
                             out=t("a1", (R, self.hidden)), round_out=not HEADS_3X)
        z2 = ops.gemm_nt(a1, self._w(fp, "3.weight"), round_out=round_out and not HEADS_3X,
                         out=t("z2", (R, self.out_dim)), precise=HEADS_3X)
        ctx = dict(x=head["x"], z1=head["z1"], mean=head["mean"], rstd=rstd, a1=a1, n=head["n"], R=R, tag=tag)
        return z2, ctx

    def backward(self, fp, ws, ctx, dz2, need_dx=True, sums_sync=None):
        """dz2 [R,256] tf32-rounded.  sums_sync(s1, s2) -> all-reduced (SyncBatchNorm backward)."""
        px = self.px
        R, tag = ctx["R"], ctx["tag"]
        t = (lambda name, shape: ws.get(tag + "/bwd/" + px + name, shape))
        ops.gemm_tn_acc(dz2, ctx["a1"], fp.g(px + "3.weight"), precise=HEADS_3X)
        da1 = ops.gemm_nn(dz2, self._w(fp, "3.weight"), out=t("da1", (R, self.hidden)), precise=HEADS_3X)
        gamma, beta = fp.p(px + "1.weight"), fp.p(px + "1.bias")
        s1, s2 = ops.bn_relu_bwd_stats(da1, ctx["z1"], ctx["mean"], ctx["rstd"], gamma, beta)
        fp.g(px + "1.bias").add_(s1)
        fp.g(px + "1.weight").add_(s2)
        if sums_sync is not None:
            s1, s2 = sums_sync(s1, s2)
        dz1 = ops.bn_relu_bwd_apply(da1, ctx["z1"], ctx["mean"], ctx["rstd"], gamma, beta, s1, s2, ctx["n"],
                                    out=t("dz1", (R, self.hidden)), round_out=not HEADS_3X)
        ops.gemm_tn_acc(dz1, ctx["x"], fp.g(px + "0.weight"), precise=HEADS_3X)
        if not need_dx:
            return None
        return ops.gemm_nn(dz1, self._w(fp, "0.weight"), round_out=False, out=t("dx", (R, self.in_dim)),
                           precise=HEADS_3X)
