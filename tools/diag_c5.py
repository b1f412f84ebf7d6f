"""diagnostic: internals of one transformer block of the c5 (ATST-large, B=2) step, GPU vs TF32-emulating oracle."""
import sys
import torch
import torch.nn.functional as F
sys.path.insert(0, '/root/repo')
from oracle import atst_oracle as O
from tests.test_parity_tf32_gpu import _mel, _waves, oracle_like
from audiossl_b200.models.atst import ATST

arch = sys.argv[1] if len(sys.argv) > 1 else "large"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 2
torch.set_num_threads(16)
torch.manual_seed(0)
m = ATST(arch=arch, ncrops=2, drop_path_rate=0.0).cuda().train()
ref = oracle_like(m)
crops = [_mel(_waves(B, 96000, 3)), _mel(_waves(B, 96000, 4))]
lengths = [torch.tensor(([601, 333] * B)[:B]).cuda(), torch.tensor([601] * B).cuda()]
rt = m._runtime(crops[0].device)
rt.enc.debug = []
loss, _, _ = m(crops, lengths)
loss.backward()
torch.cuda.synchronize()
hooks = {(n, t, i): x.cpu() for n, t, i, x in rt.enc.debug}
ws = {k[0]: v for k, v in rt.ws.bufs.items()}
rel = lambda a, b: ((a.double().cpu() - b.double()).norm() / b.double().norm().clamp_min(1e-30)).item()
enc = ref.student.encoder
D, depth = enc.embed_dim, len(enc.blocks)
S, N = 2 * B, 151
ln = torch.cat([l.cpu() for l in lengths])
plen = (ln - ln % 4) // 4 + 1
mine = dict(m.student.named_parameters())
for i in (depth - 1, depth // 2, 0):
    blk = enc.blocks[i]
    for p in blk.parameters():
        p.grad = None
    with O.tf32_emulation():
        x = hooks[("x_in", "s0", i)].view(S, N, D).clone().requires_grad_(True)
        h = blk.norm1(x)
        qkv_flat = O.linear(h, blk.attn.qkv.weight)
        qkv_flat.retain_grad()
        qkv = qkv_flat.reshape(S, N, 3, blk.heads, D // blk.heads).permute(2, 0, 3, 1, 4)
        y = O.attention_core(qkv[0], qkv[1], qkv[2], (D // blk.heads) ** -0.5, plen)
        o = y.transpose(1, 2).reshape(S, N, D)
        o.retain_grad()
        x1 = x + O.linear(o, blk.attn.proj.weight, blk.attn.proj.bias)
        x1.retain_grad()
        h2 = blk.norm2(x1)
        u = O.linear(h2, blk.mlp.fc1.weight, blk.mlp.fc1.bias)
        u.retain_grad()
        g = F.gelu(u)
        x2 = x1 + O.linear(g, blk.mlp.fc2.weight, blk.mlp.fc2.bias)
        x2.backward(hooks[("dx_in", "s0", i + 1)].view(S, N, D))
    L = lambda name: ws["s0/L%d/%s" % (i, name)].cpu()
    print("block %d forward : h %.2e qkv %.2e o %.2e x1 %.2e h2 %.2e u %.2e g %.2e" % (
        i, rel(L("h"), O.rna_tf32(h.detach()).view(S * N, D)), rel(L("qkv"), O.rna_tf32(qkv_flat.detach()).view(S * N, 3 * D)),
        rel(L("o"), o.detach().view(S * N, D)), rel(L("x1"), x1.detach().view(S * N, D)),
        rel(L("h2"), O.rna_tf32(h2.detach()).view(S * N, D)), rel(L("u"), u.detach().view(S * N, 4 * D)),
        rel(L("g"), O.rna_tf32(g.detach()).view(S * N, 4 * D))))
    print("block %d backward: du %.2e dqkv %.2e dx1 %.2e dx %.2e" % (
        i, rel(hooks[("du", "s0", i)], O.rna_tf32(u.grad).view(S * N, 4 * D)), rel(hooks[("dqkv", "s0", i)], qkv_flat.grad.view(S * N, 3 * D)),
        rel(hooks[("dx1", "s0", i)], x1.grad.view(S * N, D)), rel(hooks[("dx_in", "s0", i)], x.grad.view(S * N, D))))
    pre = "encoder.blocks.%d." % i
    print("block %d grads   : %s" % (i, " ".join("%s %.2e" % (n, rel(mine[pre + n].grad, p.grad)) for n, p in blk.named_parameters())))
    # the magnitudes involved
    dxin = hooks[("dx_in", "s0", i + 1)]
    print("block %d norms   : |dx_in| %.3e |dx1 - dx_in| %.3e |du| %.3e |dqkv| %.3e" % (
        i, dxin.norm(), (x1.grad.view(S * N, D) - dxin).norm(), u.grad.norm(), qkv_flat.grad.norm()))
