"""ncu driver: the fc1 GEMM of config 2 with the fused GELU epilogue (with / without the pre-activation store)."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from audiossl_b200 import ops  # noqa: E402

M, D = 128512, 768
A = ops.round_tf32(torch.randn(M, D, device="cuda"))
W = ops.round_tf32(torch.randn(4 * D, D, device="cuda") * 0.05)
b = torch.randn(4 * D, device="cuda")
u = torch.empty(M, 4 * D, device="cuda")
g = torch.empty(M, 4 * D, device="cuda")
ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
for _ in range(2):
    ops.gemm_nt(A, W, bias=b, out=u)
    ops.gemm_nt(A, W, bias=b, epi=ops.EPI_GELU, aux=u, round_out=True, out=g)
    ops.gemm_nt(A, W, bias=b, epi=ops.EPI_GELU, aux=None, round_out=True, out=g)
ev[0].record(); ops.gemm_nt(A, W, bias=b, out=u)
ev[1].record(); ops.gemm_nt(A, W, bias=b, epi=ops.EPI_GELU, aux=u, round_out=True, out=g)
ev[2].record(); ops.gemm_nt(A, W, bias=b, epi=ops.EPI_GELU, aux=None, round_out=True, out=g)
ev[3].record()
torch.cuda.synchronize()
print("plain %.3f ms | gelu+aux %.3f ms | gelu (no aux) %.3f ms" % (ev[0].elapsed_time(ev[1]), ev[1].elapsed_time(ev[2]), ev[2].elapsed_time(ev[3])))
