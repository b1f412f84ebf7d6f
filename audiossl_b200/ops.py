"""Torch-tensor front for the C ABI: shape checks, output allocation, current stream.

PyTorch is used for device memory and streams only; every function below launches hand-written
sm_100a kernels from libatst_b200.so.
"""
import torch

from . import _lib
from ._lib import check, ptr

EPI_STORE, EPI_GELU, EPI_DGELU, EPI_RESID, EPI_SCALE, EPI_RELU = 0, 1, 2, 3, 4, 5
EPI_GELU_H, EPI_DGELU_H = 10, 11  # aux = fp16 gelu'(pre-activation) instead of the fp32 pre-activation

# launch accounting for bench.py: kernels launched through the C ABI, algorithmic GEMM flops, and (optionally)
# a CUDA-event pair around every GEMM launch to measure the dominant kernel in place
STATS = {"launches": 0, "gemm_flops": 0.0, "gemm_bytes": 0.0, "gemm_launches": 0, "time_gemms": False,
         "gemm_events": [], "attn_events": []}


def reset_stats():
    STATS.update(launches=0, gemm_flops=0.0, gemm_bytes=0.0, gemm_launches=0, gemm_events=[], attn_events=[])


class _AttnTimer:
    """CUDA-event pair around an attention call when bench.py measures kernels in place (algorithmic bytes)."""

    def __init__(self, nbytes, tag):
        self.ev = None
        if STATS["time_gemms"]:
            self.ev = (torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True), nbytes, tag)
            self.ev[0].record()

    def done(self):
        if self.ev is not None:
            self.ev[1].record()
            STATS["attn_events"].append(self.ev)


class _GemmTimer:
    def __init__(self, flops, tag=None, nbytes=0.0):
        """tag: (format, args) - formatted only when a per-shape breakdown is being recorded"""
        self.flops = flops
        self.tag = tag
        STATS["gemm_bytes"] += nbytes  # algorithmic operand + result bytes of this launch
        STATS["launches"] += 1
        STATS["gemm_launches"] += 1
        STATS["gemm_flops"] += flops
        self.ev = None
        if STATS["time_gemms"]:
            self.ev = (torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
            self.ev[0].record()

    def done(self):
        if self.ev is not None:
            self.ev[1].record()
            tag = self.tag[0] % self.tag[1] if self.tag else ""
            STATS["gemm_events"].append((self.ev[0], self.ev[1], self.flops, tag))


def _count(n):
    STATS["launches"] += n


def _split3(x, pattern, along_rows):
    """3xTF32 validation build: hi/lo split of a GEMM operand, the three parts laid side by side (or stacked) so
    that ONE pass of the unchanged GEMM kernel over the tripled contraction dimension accumulates
    hi*hi + lo*hi + hi*lo in fp32 (include/atst_b200.h atst_split_tf32)."""
    rows, cols = x.shape
    assert x.stride(1) == 1
    out = torch.empty((3 * rows, cols) if along_rows else (rows, 3 * cols), device=x.device, dtype=torch.float32)
    check(_lib.lib().atst_split_tf32(ptr(x), x.stride(0), rows, cols, ptr(out), pattern, 1 if along_rows else 0,
                                     _lib.stream()), "atst_split_tf32")
    _count(1)
    return out


def _f32c(t, name):
    if t.dtype != torch.float32 or not t.is_cuda or not t.is_contiguous():
        raise ValueError("%s must be a contiguous fp32 CUDA tensor" % name)
    return t


def mel_forward(wav, win_length=1024, normalize=True, out=None, clip_start=None, clip_len=None):
    """wav [..., n] fp32 cuda -> normalised log-mel [..., 64, n//160+1] (reference mel_feature).
    clip_start (int64 [B] cuda) + clip_len: the mel of the window wav[b, clip_start[b] : clip_start[b] + clip_len] of
    every clip (the train transform's RandomCrop) without materialising the crops."""
    lead = wav.shape[:-1]
    w2 = _f32c(wav.reshape(-1, wav.shape[-1]), "wav")
    n = wav.shape[-1] if clip_len is None else int(clip_len)
    if clip_start is not None and (clip_start.dtype != torch.int64 or clip_start.numel() != w2.shape[0] or n > wav.shape[-1]):
        raise ValueError("clip_start must be int64 [B] and clip_len <= the waveform length")
    B = w2.shape[0]
    T = n // 160 + 1
    if out is None:
        out = torch.empty((B, 64, T), device=wav.device, dtype=torch.float32)
    ws = torch.empty((2 * B,), device=wav.device, dtype=torch.int32)
    check(_lib.lib().atst_mel_forward(ptr(w2), B, n, w2.stride(0), ptr(clip_start), win_length, ptr(out), 64 * T,
                                      ptr(ws), 1 if normalize else 0, _lib.stream()), "atst_mel_forward")
    _count(1)
    return out.reshape(*lead, 64, T)


def _aux_bytes(aux, epi):
    if aux is None:
        return 0.0
    want = torch.float16 if epi in (EPI_GELU_H, EPI_DGELU_H) else torch.float32
    if aux.dtype != want:
        raise TypeError("epilogue %d takes a %s aux tensor, got %s" % (epi, want, aux.dtype))
    return float(aux.numel() * aux.element_size())


def gemm_nt(A, B, bias=None, epi=EPI_STORE, resid=None, aux=None, rowscale=None, rows_per_seq=1, round_out=False,
            out=None, precise=False):
    """out[M,N] = epi(A[M,K] @ B[N,K]^T + bias).  precise: error-compensated 3xTF32 product of UNROUNDED fp32 operands
    (hi / lo split, one pass of the same kernel over the tripled contraction) - the projector / predictor heads."""
    M, K = A.shape
    N = B.shape[0]
    assert B.shape[1] == K
    if out is None:
        out = torch.empty((M, N), device=A.device, dtype=torch.float32)
    if _lib.is_precise() or precise:
        A, B, K = _split3(A, 0, False), _split3(B, 1, False), 3 * K
    _t = _GemmTimer(2.0 * M * N * K, ("nt %dx%dx%d e%d", (M, N, K, epi)),
                   4.0 * (M * K + N * K + M * N * (1 + (resid is not None))) + _aux_bytes(aux, epi))
    check(_lib.lib().atst_gemm_nt(ptr(A), A.stride(0), ptr(B), B.stride(0), ptr(out), out.stride(0), M, N, K,
                                  ptr(bias), epi, ptr(resid), resid.stride(0) if resid is not None else 0,
                                  ptr(aux), aux.stride(0) if aux is not None else 0, ptr(rowscale), rows_per_seq,
                                  1 if round_out else 0, _lib.stream()), "atst_gemm_nt")
    _t.done()
    return out


def gemm_nn(A, W, epi=EPI_STORE, aux=None, rowscale=None, rows_per_seq=1, round_out=False, out=None, colsum_out=None,
            precise=False):
    """out[M,N] = epi(A[M,K] @ W[K,N])  (dgrad against a Linear weight W[out=K, in=N]).
    colsum_out [N] (optional) += column sums of out, taken in the GEMM epilogue."""
    M, K = A.shape
    N = W.shape[1]
    assert W.shape[0] == K
    if out is None:
        out = torch.empty((M, N), device=A.device, dtype=torch.float32)
    if _lib.is_precise() or precise:
        A, W, K = _split3(A, 0, False), _split3(W, 1, True), 3 * K
    _t = _GemmTimer(2.0 * M * N * K, ("nn %dx%dx%d e%d", (M, N, K, epi)), 4.0 * (M * K + N * K + M * N) + _aux_bytes(aux, epi))
    check(_lib.lib().atst_gemm_nn(ptr(A), A.stride(0), ptr(W), W.stride(0), ptr(out), out.stride(0), M, N, K, epi,
                                  ptr(aux), aux.stride(0) if aux is not None else 0, ptr(rowscale), rows_per_seq,
                                  1 if round_out else 0, ptr(colsum_out), _lib.stream()), "atst_gemm_nn")
    _t.done()
    return out


def gemm_tn_acc(A, B, out, precise=False):
    """out[M,N] += A[T,M]^T @ B[T,N]  (wgrad; accumulates)."""
    T, M = A.shape
    N = B.shape[1]
    assert B.shape[0] == T and tuple(out.shape) == (M, N)
    if _lib.is_precise() or precise:
        A, B, T = _split3(A, 0, True), _split3(B, 1, True), 3 * T
    _t = _GemmTimer(2.0 * M * N * T, ("tn %dx%dx%d", (M, N, T)), 4.0 * (T * M + T * N + M * N))
    check(_lib.lib().atst_gemm_tn(ptr(A), A.stride(0), ptr(B), B.stride(0), ptr(out), out.stride(0), M, N, T,
                                  _lib.stream()), "atst_gemm_tn")
    _t.done()
    return out


def layernorm_fwd(x, gamma, beta, rows, D, x_stride=None, out=None, out_stride=None, eps=1e-6, round_out=True,
                  mean=None, rstd=None):
    x_stride = D if x_stride is None else x_stride
    if out is None:
        out = torch.empty((rows, D), device=x.device, dtype=torch.float32)
    out_stride = D if out_stride is None else out_stride
    if mean is None:
        mean = torch.empty((rows,), device=x.device, dtype=torch.float32)
    if rstd is None:
        rstd = torch.empty((rows,), device=x.device, dtype=torch.float32)
    check(_lib.lib().atst_layernorm_forward(ptr(x), x_stride, ptr(gamma), ptr(beta), ptr(out), out_stride, ptr(mean),
                                            ptr(rstd), rows, D, eps, 1 if round_out else 0, _lib.stream()),
          "atst_layernorm_forward")
    _count(1)
    return out, mean, rstd


def layernorm_bwd(dy, x, mean, rstd, gamma, dgamma, dbeta, rows, D, dres=None, dx=None, dy_stride=None,
                  x_stride=None, dres_stride=None, dx_stride=None, dys=None, dys_stride=None, rowscale=None,
                  rows_per_seq=1, colsum_out=None):
    dy_stride = D if dy_stride is None else dy_stride
    x_stride = D if x_stride is None else x_stride
    dres_stride = D if dres_stride is None else dres_stride
    if dx is None:
        dx = torch.empty((rows, D), device=x.device, dtype=torch.float32)
    dx_stride = D if dx_stride is None else dx_stride
    check(_lib.lib().atst_layernorm_backward(ptr(dy), dy_stride, ptr(x), x_stride, ptr(mean), ptr(rstd), ptr(gamma),
                                             ptr(dres), dres_stride, ptr(dx), dx_stride, ptr(dgamma), ptr(dbeta),
                                             rows, D, ptr(dys), D if dys_stride is None else dys_stride, ptr(rowscale),
                                             rows_per_seq, ptr(colsum_out), _lib.stream()), "atst_layernorm_backward")
    _count(1)
    return dx


def attention_fwd(qkv, S, N, H, lengths=None, out=None, lse=None):
    D = H * 64
    if out is None:
        out = torch.empty((S * N, D), device=qkv.device, dtype=torch.float32)
    if lse is None:
        lse = torch.empty((S, H, N), device=qkv.device, dtype=torch.float32)
    _t = _AttnTimer(4.0 * S * N * 4 * D, "fwd")  # qkv read once, o written once
    check(_lib.lib().atst_attention_forward(ptr(qkv), ptr(out), ptr(lse), ptr(lengths), S, N, H, _lib.stream()),
          "atst_attention_forward")
    _t.done()
    _count(1)
    return out, lse


def attention_bwd(qkv, o, d_o, lse, S, N, H, lengths=None, dqkv=None, delta_ws=None):
    if dqkv is None:
        dqkv = torch.empty_like(qkv)
    if delta_ws is None:
        delta_ws = torch.empty((S, H, N), device=qkv.device, dtype=torch.float32)
    D = H * 64
    _t = _AttnTimer(4.0 * S * N * 8 * D, "bwd")  # qkv, o, dO read once, dqkv written once
    check(_lib.lib().atst_attention_backward(ptr(qkv), ptr(o), ptr(d_o), ptr(lse), ptr(delta_ws), ptr(dqkv),
                                             ptr(lengths), S, N, H, _lib.stream()), "atst_attention_backward")
    _t.done()
    _count(3)
    return dqkv


def patchify(mel, out=None):
    """mel [S,1,64,T] -> [S*(T//4), 256] (PatchEmbed_v2 rearrange, tf32-rounded)."""
    S, _, H, T = mel.shape
    assert H == 64 and mel.is_contiguous()
    P = T // 4
    if out is None:
        out = torch.empty((S * P, 256), device=mel.device, dtype=torch.float32)
    check(_lib.lib().atst_patchify(ptr(mel), 64 * T, S, T, ptr(out), _lib.stream()), "atst_patchify")
    _count(1)
    return out


def tokens_fwd(pe, cls, pos, S, P, D, use_cls=True, mask_embed=None, mask=None, out=None):
    N = P + (1 if use_cls else 0)
    if out is None:
        out = torch.empty((S * N, D), device=pe.device, dtype=torch.float32)
    check(_lib.lib().atst_tokens_forward(ptr(pe), ptr(cls), ptr(pos), ptr(mask_embed), ptr(mask), ptr(out), S, P, D,
                                         1 if use_cls else 0, _lib.stream()), "atst_tokens_forward")
    _count(1)
    return out


def tokens_bwd(dx, dpe, dpos, dcls, S, P, D, use_cls=True, mask=None, dmask_embed=None):
    check(_lib.lib().atst_tokens_backward(ptr(dx), ptr(mask), ptr(dpe), ptr(dpos), ptr(dcls), ptr(dmask_embed), S, P,
                                          D, 1 if use_cls else 0, _lib.stream()), "atst_tokens_backward")
    _count(1)


def colsum_acc(X, out, rows=None, cols=None, ld=None):
    rows = X.shape[0] if rows is None else rows
    cols = X.shape[1] if cols is None else cols
    ld = X.stride(0) if ld is None else ld
    check(_lib.lib().atst_colsum_accumulate(ptr(X), ld, rows, cols, ptr(out), _lib.stream()), "atst_colsum")
    _count(1)


def bn_stats(X):
    rows, cols = X.shape
    mean = torch.empty((cols,), device=X.device, dtype=torch.float32)
    m2 = torch.empty((cols,), device=X.device, dtype=torch.float32)
    check(_lib.lib().atst_bn_stats(ptr(X), rows, cols, ptr(mean), ptr(m2), _lib.stream()), "atst_bn_stats")
    _count(1)
    return mean, m2


def bn_finalize(mean, m2, count, running_mean=None, running_var=None, eps=1e-5, momentum=0.1):
    rstd = torch.empty_like(mean)
    check(_lib.lib().atst_bn_finalize(ptr(mean), ptr(m2), float(count), eps, momentum, ptr(rstd), ptr(running_mean),
                                      ptr(running_var), mean.numel(), _lib.stream()), "atst_bn_finalize")
    _count(1)
    return rstd


def bn_relu_fwd(X, mean, rstd, gamma, beta, out=None, round_out=True):
    rows, cols = X.shape
    if out is None:
        out = torch.empty_like(X)
    check(_lib.lib().atst_bn_relu_forward(ptr(X), ptr(mean), ptr(rstd), ptr(gamma), ptr(beta), ptr(out), rows, cols,
                                          1 if round_out else 0, _lib.stream()), "atst_bn_relu_forward")
    _count(1)
    return out


def bn_relu_bwd_stats(dY, X, mean, rstd, gamma, beta):
    rows, cols = X.shape
    s1 = torch.empty((cols,), device=X.device, dtype=torch.float32)
    s2 = torch.empty((cols,), device=X.device, dtype=torch.float32)
    check(_lib.lib().atst_bn_relu_backward_stats(ptr(dY), ptr(X), ptr(mean), ptr(rstd), ptr(gamma), ptr(beta), rows,
                                                 cols, ptr(s1), ptr(s2), _lib.stream()), "atst_bn_relu_backward_stats")
    _count(1)
    return s1, s2


def bn_relu_bwd_apply(dY, X, mean, rstd, gamma, beta, s1, s2, count, out=None, round_out=True):
    rows, cols = X.shape
    if out is None:
        out = torch.empty_like(X)
    check(_lib.lib().atst_bn_relu_backward_apply(ptr(dY), ptr(X), ptr(mean), ptr(rstd), ptr(gamma), ptr(beta),
                                                 ptr(s1), ptr(s2), float(count), ptr(out), rows, cols,
                                                 1 if round_out else 0, _lib.stream()), "atst_bn_relu_backward_apply")
    _count(1)
    return out


def byol_loss(student, teacher, ncrops, B, dstudent=None, acc=None):
    if dstudent is None:
        dstudent = torch.empty_like(student)
    if acc is None:
        acc = torch.empty((1 + 4 * 256,), device=student.device, dtype=torch.float32)
    check(_lib.lib().atst_byol_loss(ptr(student), ptr(teacher), ncrops, B, ptr(dstudent), ptr(acc), _lib.stream()),
          "atst_byol_loss")
    _count(1)
    return dstudent, acc


def byol_finalize(acc, n_student_rows, n_teacher_rows, ncrops, B, out=None):
    if out is None:
        out = torch.empty((3,), device=acc.device, dtype=torch.float32)
    check(_lib.lib().atst_byol_finalize(ptr(acc), float(n_student_rows), float(n_teacher_rows), ncrops, B, ptr(out),
                                        _lib.stream()), "atst_byol_finalize")
    _count(1)
    return out


def ema_update(k, q, m, m_dev=None):
    """m_dev (1-element cuda tensor, optional): the momentum is read from it on the device (CUDA-graph replay)."""
    assert k.numel() == q.numel()
    check(_lib.lib().atst_ema_update(ptr(k), ptr(q), float(m), ptr(m_dev), k.numel(), _lib.stream()), "atst_ema_update")
    _count(1)


def adamw_step(p, g, m, v, step, lr, wd, beta1=0.9, beta2=0.999, eps=1e-6, grad_scale=1.0, dyn=None):
    """dyn (2-element cuda tensor, optional): {lr * sqrt(1 - beta2^t) / (1 - beta1^t), lr * wd} read on the device."""
    check(_lib.lib().atst_adamw_step(ptr(p), ptr(g), ptr(m), ptr(v), p.numel(), int(step), float(lr), float(wd),
                                     beta1, beta2, eps, float(grad_scale), ptr(dyn), _lib.stream()), "atst_adamw_step")
    _count(1)


def mixup_fwd(x, bank, idx, alpha, out, zlen=None, start=None):
    """x [B,(1,)Hm,T]; bank [n,(1,)Hm,Tb] (or None when every idx is negative); zlen / start int32 [B] or None."""
    B, Hm, T = x.shape[0], x.shape[-2], x.shape[-1]
    bank_T = T if bank is None else bank.shape[-1]
    check(_lib.lib().atst_mixup_forward(ptr(x), T, ptr(bank), bank_T, ptr(idx), ptr(zlen), ptr(start), ptr(alpha),
                                        ptr(out), Hm, B, _lib.stream()), "atst_mixup_forward")
    _count(1)
    return out


def resize_crop_fwd(lms, rect, out, canvas_h, canvas_w):
    B, Hm, T = lms.shape[0], lms.shape[-2], lms.shape[-1]
    check(_lib.lib().atst_resize_crop_forward(ptr(lms), ptr(rect), ptr(out), B, Hm, T, canvas_h, canvas_w,
                                              _lib.stream()), "atst_resize_crop_forward")
    _count(1)
    return out


def gather_rows(x, idx, out):
    rows, D = out.shape
    check(_lib.lib().atst_gather_rows(ptr(x), ptr(idx), ptr(out), rows, D, _lib.stream()), "atst_gather_rows")
    _count(1)
    return out


def scatter_rows(src, idx, dst):
    rows, D = src.shape
    check(_lib.lib().atst_scatter_rows(ptr(src), ptr(idx), ptr(dst), rows, D, _lib.stream()), "atst_scatter_rows")
    _count(1)
    return dst


def gelu_fwd(u, out):
    check(_lib.lib().atst_gelu_forward(ptr(u), ptr(out), u.numel(), _lib.stream()), "atst_gelu_forward")
    _count(1)
    return out


def gelu_bwd_(d, u, colsum_out=None):
    rows, cols = d.shape
    check(_lib.lib().atst_gelu_backward(ptr(d), ptr(u), rows, cols, ptr(colsum_out), _lib.stream()), "atst_gelu_backward")
    _count(1)
    return d


def round_tf32(src, dst=None):
    if dst is None:
        dst = torch.empty_like(src)
    check(_lib.lib().atst_round_tf32(ptr(src), ptr(dst), src.numel(), _lib.stream()), "atst_round_tf32")
    _count(1)
    return dst


def axpy(y, x, a):
    check(_lib.lib().atst_axpy(ptr(y), ptr(x), float(a), y.numel(), _lib.stream()), "atst_axpy")
    _count(1)
