"""layer-by-layer comparison of the CUDA engine against the TF32-emulating oracle (locates rounding-point
disagreements between the kernels and oracle/atst_oracle.py "TF32 operand emulation")."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from audiossl_b200 import ops  # noqa: E402
from audiossl_b200.models.atst import ATST  # noqa: E402
from oracle import atst_oracle as O  # noqa: E402
from tests import util  # noqa: E402


def rel(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return ((a - b).norm() / b.norm().clamp_min(1e-30)).item()


case = sys.argv[1] if len(sys.argv) > 1 else "tiny2"
c = util.CASES[case]
m = ATST(arch=dict(embed_dim=c["dim"], depth=c["depth"], num_heads=c["heads"]), ncrops=c["ncrops"], drop_path_rate=0.0)
util.load_det(m)
m.cuda().train()
ref = O.OracleATST(ncrops=c["ncrops"], embed_dim=c["dim"], depth=c["depth"], num_heads=c["heads"])
util.load_det(ref)
ref.train()
crops, lengths = util.make_inputs(case, c["B"], c["widths"], c["lens"])
rt = m._runtime(torch.device("cuda", 0))
fs = rt.fs
ops.round_tf32(fs.data, fs.compute)
mel = torch.cat(crops[:2]).cuda()
ln = torch.cat(lengths[:2]).cuda()
out, ctx = rt.enc.forward(fs, rt.ws, mel, ln, dp=None, save=True, tag="dbg")
torch.cuda.synchronize()
enc = ref.student.encoder
D, H = c["dim"], c["heads"]
R = O.rna_tf32
with torch.no_grad(), O.tf32_emulation():
    melc, lnc = torch.cat(crops[:2]), torch.cat(lengths[:2])
    print("patches (rounded)", rel(ctx["patches"].reshape(melc.shape[0], -1, 256), R(enc.patchify(melc))))
    x, plen = enc.tokens(melc, lnc)
    S, N, _ = x.shape
    print("x0", rel(ctx["layers"][0]["x"].reshape(S, N, D), x))
    for i, blk in enumerate(enc.blocks):
        L = ctx["layers"][i]
        h = blk.norm1(x)
        print(i, "h (rounded)", rel(L["h"].reshape(S, N, D), R(h)), " fp32-LN itself:", rel(L["h"].reshape(S, N, D), h))
        qkv = O.linear(h, blk.attn.qkv.weight)
        print(i, "qkv (rounded)", rel(L["qkv"].reshape(S, N, 3 * D), R(qkv)))
        q = qkv.reshape(S, N, 3, H, 64).permute(2, 0, 3, 1, 4)
        o = O.attention_core(q[0], q[1], q[2], 0.125, plen + 1).transpose(1, 2).reshape(S, N, D)
        print(i, "o", rel(L["o"].reshape(S, N, D), o))
        x1 = x + O.linear(o, blk.attn.proj.weight, blk.attn.proj.bias)
        print(i, "x1", rel(L["x1"].reshape(S, N, D), x1))
        h2 = blk.norm2(x1)
        print(i, "h2 (rounded)", rel(L["h2"].reshape(S, N, D), R(h2)))
        u = O.linear(h2, blk.mlp.fc1.weight, blk.mlp.fc1.bias)
        print(i, "u", rel(L["u"].reshape(S, N, 4 * D), u))
        g = torch.nn.functional.gelu(u)
        print(i, "g (rounded)", rel(L["g"].reshape(S, N, 4 * D), R(g)))
        x = x1 + O.linear(g, blk.mlp.fc2.weight, blk.mlp.fc2.bias)
        nxt = ctx["layers"][i + 1]["x"] if i + 1 < len(enc.blocks) else ctx["x_final"]
        print(i, "x2", rel(nxt.reshape(S, N, D), x))
    cls = enc.norm(x)[:, 0]
    print("cls (rounded)", rel(out, R(cls)))
    zz, pctx = rt.proj.forward(fs, rt.ws, out, None, "dbg", True, None)
    p = ref.student.projector
    z1 = p[0](cls)
    print("proj z1", rel(pctx["z1"], z1))
    a1 = p[2](p[1](z1))
    print("proj a1 (rounded)", rel(pctx["a1"], R(a1)))
    print("proj out (rounded)", rel(zz, R(p[3](a1))))
    # weights
    w = fs.c("encoder.blocks.0.attn.qkv.weight").cpu()
    print("tf32 weight copy == rna(weight):", torch.equal(w, R(enc.blocks[0].attn.qkv.weight)))
