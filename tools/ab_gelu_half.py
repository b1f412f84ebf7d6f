"""A/B of the student's MLP side stream at config-2 shapes: fp32 pre-activation u (EPI_GELU / EPI_DGELU) against
fp16 gelu'(u) (EPI_GELU_H / EPI_DGELU_H).  The variants are interleaved and repeated so that the power-capped clock
treats them alike; CUDA events on the launching stream, median of the repetitions.

    python tools/ab_gelu_half.py [reps]
"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from audiossl_b200 import ops  # noqa: E402

reps = int(sys.argv[1]) if len(sys.argv) > 1 else 15
M, D = 128512, 768
torch.manual_seed(0)
h = ops.round_tf32(torch.randn(M, D, device="cuda"))
W1 = ops.round_tf32(torch.randn(4 * D, D, device="cuda") * 0.05)
b1 = torch.randn(4 * D, device="cuda")
dy = ops.round_tf32(torch.randn(M, D, device="cuda"))
W2 = ops.round_tf32(torch.randn(D, 4 * D, device="cuda") * 0.05)
u = torch.empty(M, 4 * D, device="cuda")
gp = torch.empty(M, 4 * D, device="cuda", dtype=torch.float16)
g = torch.empty(M, 4 * D, device="cuda")
du = torch.empty(M, 4 * D, device="cuda")
Wp = ops.round_tf32(torch.randn(D, D, device="cuda") * 0.05)
bp = torch.randn(D, device="cuda")
x1 = torch.empty(M, D, device="cuda")
Wq = ops.round_tf32(torch.randn(3 * D, D, device="cuda") * 0.05)
qkv = torch.empty(M, 3 * D, device="cuda")

variants = {
    "fc1 plain (u only)": lambda: ops.gemm_nt(h, W1, bias=b1, out=u),
    "fc1 + GELU, no side stream (teacher)": lambda: ops.gemm_nt(h, W1, bias=b1, epi=ops.EPI_GELU, aux=None, round_out=True, out=g),
    "fc1 + GELU, u fp32": lambda: ops.gemm_nt(h, W1, bias=b1, epi=ops.EPI_GELU, aux=u, round_out=True, out=g),
    "fc1 + GELU, gelu' fp16": lambda: ops.gemm_nt(h, W1, bias=b1, epi=ops.EPI_GELU_H, aux=gp, round_out=True, out=g),
    "fc1 + GELU, no side stream, compact kernel": lambda: ops.gemm_nt(h, W1, bias=b1, epi=ops.EPI_GELU_H, aux=None, round_out=True, out=g),
    "fc2 dgrad plain": lambda: ops.gemm_nn(dy, W2, round_out=True, out=du),
    "fc2 dgrad * gelu'(u fp32)": lambda: ops.gemm_nn(dy, W2, epi=ops.EPI_DGELU, aux=u, round_out=True, out=du),
    "fc2 dgrad * gelu' fp16": lambda: ops.gemm_nn(dy, W2, epi=ops.EPI_DGELU_H, aux=gp, round_out=True, out=du),
    "proj plain (N = K = 768)": lambda: ops.gemm_nt(h, Wp, bias=bp, out=x1),
    "proj + residual": lambda: ops.gemm_nt(h, Wp, bias=bp, epi=ops.EPI_RESID, resid=dy, out=x1),
    "fc2 + residual (K = 3072)": lambda: ops.gemm_nt(g, W2, bias=bp, epi=ops.EPI_RESID, resid=dy, out=x1),
    "qkv plain (N = 2304)": lambda: ops.gemm_nt(h, Wq, round_out=True, out=qkv),
}
for f in variants.values():  # warm-up (tensor maps, instruction cache, clocks)
    for _ in range(3):
        f()
torch.cuda.synchronize()
times = {k: [] for k in variants}
for _ in range(reps):
    for k, f in variants.items():
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        f()
        e1.record()
        e1.synchronize()
        times[k].append(e0.elapsed_time(e1))
flops = {k: 2.0 * M * 4 * D * D for k in variants}
flops["proj plain (N = K = 768)"] = flops["proj + residual"] = 2.0 * M * D * D
flops["qkv plain (N = 2304)"] = 2.0 * M * 3 * D * D
for k, t in times.items():
    t = sorted(t)
    med = t[len(t) // 2]
    print("%-44s  median %.3f ms  (min %.3f)  %.0f TFLOP/s" % (k, med, t[0], flops[k] / med * 1e-9))
