"""ncu driver: the proj GEMM of config 2 (M x 768 x 768) with the residual + DropPath epilogue."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from audiossl_b200 import ops  # noqa: E402

M, D, N = 128512, 768, 251
A = ops.round_tf32(torch.randn(M, D, device="cuda"))
W = ops.round_tf32(torch.randn(D, D, device="cuda") * 0.05)
b = torch.randn(D, device="cuda")
x = torch.randn(M, D, device="cuda")
sc = torch.rand(M // N, device="cuda") + 0.5
out = torch.empty(M, D, device="cuda")
for _ in range(3):
    ops.gemm_nt(A, W, bias=b, out=out)
    ops.gemm_nt(A, W, bias=b, epi=ops.EPI_RESID, resid=x, rowscale=sc, rows_per_seq=N, out=out)
ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
ev[0].record(); ops.gemm_nt(A, W, bias=b, out=out)
ev[1].record(); ops.gemm_nt(A, W, bias=b, epi=ops.EPI_RESID, resid=x, rowscale=sc, rows_per_seq=N, out=out)
ev[2].record()
torch.cuda.synchronize()
print("plain %.3f ms | resid %.3f ms" % (ev[0].elapsed_time(ev[1]), ev[1].elapsed_time(ev[2])))
