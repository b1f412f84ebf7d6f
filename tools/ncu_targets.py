"""The kernels the round's profiles are taken on, at BASELINE config-2 shapes (one process, a few launches each):

    ncu --set full --clock-control none --import-source on -k regex:'mel_kernel|gemm2_tf32|attn_' -o out python tools/ncu_targets.py

  mel     64 clips x 10 s                                      (mel_kernel)
  gemm    qkv plain / fc1 + GELU (two output streams) / fc2 dgrad + GELU' / proj + residual at M = 128 512 rows
  attn    forward and backward, 512 sequences x 251 tokens x 12 heads
"""
import sys

import torch

sys.path.insert(0, __file__.rsplit("/", 2)[0])
from audiossl_b200 import ops  # noqa: E402

which = sys.argv[1] if len(sys.argv) > 1 else "all"
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
torch.manual_seed(0)
dev = "cuda"
if which in ("all", "mel"):
    wav = torch.randn(64, 1, 160000, device=dev) * 0.1
    for _ in range(reps):
        ops.mel_forward(wav)
if which in ("all", "gemm"):
    M, D = 128512, 768
    h = ops.round_tf32(torch.randn(M, D, device=dev))
    Wqkv = ops.round_tf32(torch.randn(3 * D, D, device=dev) * 0.02)
    W1 = ops.round_tf32(torch.randn(4 * D, D, device=dev) * 0.02)
    b1 = torch.zeros(4 * D, device=dev)
    Wp = ops.round_tf32(torch.randn(D, D, device=dev) * 0.02)
    bp = torch.zeros(D, device=dev)
    qkv = torch.empty(M, 3 * D, device=dev)
    u, g, du = (torch.empty(M, 4 * D, device=dev) for _ in range(3))
    x1 = torch.empty(M, D, device=dev)
    for _ in range(reps):
        ops.gemm_nt(h, Wqkv, round_out=True, out=qkv)                                   # plain
        ops.gemm_nt(h, W1, bias=b1, epi=ops.EPI_GELU, aux=u, round_out=True, out=g)     # fc1 + GELU, u and g out
        ops.gemm_nn(h, W1.t().contiguous().t() if False else W1.new_empty(D, 4 * D).normal_(0, 0.02), epi=ops.EPI_DGELU,
                    aux=u, round_out=True, out=du)                                      # fc2 dgrad + GELU'
        ops.gemm_nt(h, Wp, bias=bp, epi=ops.EPI_RESID, resid=h, out=x1)                 # proj + residual
if which in ("all", "attn"):
    S, N, H = 512, 251, 12
    qkv = ops.round_tf32(torch.randn(S * N, 3 * H * 64, device=dev))
    d_o = ops.round_tf32(torch.randn(S * N, H * 64, device=dev))
    for _ in range(reps):
        o, lse = ops.attention_fwd(qkv, S, N, H)
        ops.attention_bwd(qkv, o, d_o, lse, S, N, H)
torch.cuda.synchronize()
print("done")
