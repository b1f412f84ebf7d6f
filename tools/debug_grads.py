"""backward bring-up: per-parameter and per-layer gradient comparison against oracle autograd."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from audiossl_b200.models.atst import ATST  # noqa: E402
from oracle import atst_oracle as O  # noqa: E402
from tests import util  # noqa: E402


def rel(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return ((a - b).norm() / b.norm().clamp_min(1e-30)).item()


case = sys.argv[1] if len(sys.argv) > 1 else "tiny2"
c = util.CASES[case]
cfg = dict(embed_dim=c["dim"], depth=c["depth"], num_heads=c["heads"])
m = ATST(arch=cfg, ncrops=c["ncrops"], drop_path_rate=0.0)
util.load_det(m)
m.cuda().train()
ref = O.OracleATST(ncrops=c["ncrops"], **cfg)
util.load_det(ref)
ref.train()
crops, lengths = util.make_inputs(case, c["B"], c["widths"], c["lens"])

# oracle with activation-gradient capture at block boundaries
acts = {}
enc = ref.student.encoder
for i, blk in enumerate(enc.blocks):
    def hook(mod, inp, out, i=i):
        out.retain_grad()
        acts[("x_out", i)] = out
    blk.register_forward_hook(hook)
rl, _, _ = ref(crops, lengths)
rl.backward()

rt = m._runtime(torch.device("cuda", 0))
rt.enc.debug = []
loss, _, _ = m([x.cuda() for x in crops], [x.cuda() for x in lengths])
loss.backward()
torch.cuda.synchronize()
print("loss", loss.item(), rl.item())
S = sum(x.shape[0] for x in crops)
D = c["dim"]
dbg = {(n, i): t for n, i, t in rt.enc.debug}
depth = c["depth"]
for i in reversed(range(depth)):
    g = acts[("x_out", i)].grad  # grad wrt output of block i
    name = ("dx_out", depth) if i == depth - 1 else ("dx_in", i + 1)
    mine = dbg[name].reshape(g.shape)
    print("grad wrt block %d output: rel %.3e  (|ref| %.3e |mine| %.3e)" % (i, rel(mine, g), g.norm(), mine.norm()))
for k in ("du", "dx1", "dqkv", "dx_in"):
    for i in reversed(range(depth)):
        print(k, i, "norm %.4e" % dbg[(k, i)].norm().item())
print("---- parameter grads")
rp = dict(ref.student.named_parameters())
for name, p in m.student.named_parameters():
    if rp[name].grad is None:
        continue
    print("%-50s rel %.3e" % (name, rel(p.grad, rp[name].grad)))
