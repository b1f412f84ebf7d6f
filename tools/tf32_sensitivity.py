import sys, torch, numpy as np
sys.path.insert(0, '/root/repo')
from oracle import atst_oracle as O
from tests import util
import torch.nn.functional as F

def rna(x):
    i = x.contiguous().view(torch.int32)
    i = (i + 0x1000) & ~0x1FFF
    return i.view(torch.float32)

class RoundSTE(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x): return rna(x)
    @staticmethod
    def backward(ctx, g): return g

orig_linear = F.linear
def tf32_linear(x, w, b=None):
    return orig_linear(RoundSTE.apply(x), RoundSTE.apply(w), b)

case = sys.argv[1] if len(sys.argv) > 1 else "tiny2b32"
c = util.CASES[case]
def run(emu):
    m = O.OracleATST(ncrops=c["ncrops"], embed_dim=c["dim"], depth=c["depth"], num_heads=c["heads"])
    util.load_det(m); m.train()
    crops, lengths = util.make_inputs(case, c["B"], c["widths"], c["lens"])
    if emu: F.linear = tf32_linear; torch.nn.functional.linear = tf32_linear
    try:
        loss, _, _ = m(crops, lengths)
        loss.backward()
    finally:
        F.linear = orig_linear; torch.nn.functional.linear = orig_linear
    return loss.item(), {n: p.grad.clone() for n, p in m.student.named_parameters() if p.grad is not None}
l0, g0 = run(False)
l1, g1 = run(True)
print("loss", l0, l1, abs(l1-l0)/abs(l0))
for n in g0:
    e = ((g1[n]-g0[n]).norm()/g0[n].norm().clamp_min(1e-30)).item()
    if any(k in n for k in ("predictor", "projector", "blocks.1.mlp", "norm.")):
        print("%-45s %.3e  |g| %.3e" % (n, e, g0[n].norm()))
