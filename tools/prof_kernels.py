"""small driver for ncu captures of the hot kernels at BASELINE config-2 shapes (fewer sequences)."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from audiossl_b200 import ops  # noqa: E402

S, N, H, D = 512, 251, 12, 768
M = S * N
A = ops.round_tf32(torch.randn(M, D, device="cuda"))
W = ops.round_tf32(torch.randn(3 * D, D, device="cuda") * 0.05)
C = torch.empty(M, 3 * D, device="cuda")
for _ in range(3):
    ops.gemm_nt(A, W, out=C, round_out=True)
o, lse = ops.attention_fwd(C, S, N, H)
d_o = ops.round_tf32(torch.randn_like(o))
dqkv = torch.empty_like(C)
for _ in range(2):
    ops.attention_fwd(C, S, N, H, out=o, lse=lse)
    ops.attention_bwd(C, o, d_o, lse, S, N, H, dqkv=dqkv)
wav = torch.randn(64, 160000, device="cuda") * 0.1
for _ in range(2):
    ops.mel_forward(wav)
torch.cuda.synchronize()
