"""ncu targets: the pair GEMM's epilogue classes at config-2 shapes, two launches each (profile the second).

    ncu --set full --clock-control none --import-source on -k regex:gemm2_tf32 -o out python tools/ncu_epi_targets.py
"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from audiossl_b200 import ops  # noqa: E402

M, D = 128512, 768
torch.manual_seed(0)
h = ops.round_tf32(torch.randn(M, D, device="cuda"))
W1 = ops.round_tf32(torch.randn(4 * D, D, device="cuda") * 0.05)
b1 = torch.randn(4 * D, device="cuda")
dy = ops.round_tf32(torch.randn(M, D, device="cuda"))
W2 = ops.round_tf32(torch.randn(D, 4 * D, device="cuda") * 0.05)
gp = torch.empty(M, 4 * D, device="cuda", dtype=torch.float16)
g = torch.empty(M, 4 * D, device="cuda")
du = torch.empty(M, 4 * D, device="cuda")
Wp = ops.round_tf32(torch.randn(D, D, device="cuda") * 0.05)
bp = torch.randn(D, device="cuda")
x1 = torch.empty(M, D, device="cuda")
for _ in range(2):
    ops.gemm_nt(h, W1, bias=b1, out=g)                                                   # 0/1  plain fc1
    ops.gemm_nt(h, W1, bias=b1, epi=ops.EPI_GELU, aux=None, round_out=True, out=g)       # 2/3  GELU, no side stream
    ops.gemm_nt(h, W1, bias=b1, epi=ops.EPI_GELU_H, aux=gp, round_out=True, out=g)       # 4/5  GELU + fp16 gelu'
    ops.gemm_nn(dy, W2, epi=ops.EPI_DGELU_H, aux=gp, round_out=True, out=du)             # 6/7  dgrad * fp16 gelu'
    ops.gemm_nt(h, Wp, bias=bp, out=x1)                                                  # 8/9  plain proj
    ops.gemm_nt(h, Wp, bias=bp, epi=ops.EPI_RESID, resid=dy, out=x1)                     # 10/11 proj + residual
torch.cuda.synchronize()
print("done")
