import sys, torch
sys.path.insert(0, '/root/repo')
from audiossl_b200 import ops
torch.backends.cuda.matmul.allow_tf32 = False
torch.manual_seed(0)
def rel(a, b):
    a, b = a.double(), b.double()
    return ((a - b).norm() / b.norm()).item()
R = lambda *s: ops.round_tf32(torch.randn(*s, device="cuda"))
for T, M, N in [(604, 1024, 4096), (604, 4096, 1024), (604, 3072, 1024), (604, 1024, 1024), (2008, 768, 3072), (604, 768, 3072), (600, 1024, 4096), (608, 1024, 4096), (1208, 1024, 4096)]:
    G, A = R(T, M) * 0.1, R(T, N)
    dW = torch.zeros(M, N, device="cuda")
    ops.gemm_tn_acc(G, A, dW)
    print("tn  T=%d M=%d N=%d  rel %.2e" % (T, M, N, rel(dW, G.double().t() @ A.double())))
for M, K, N in [(604, 4096, 1024), (604, 1024, 4096), (604, 3072, 1024), (604, 1024, 1024)]:
    A, W = R(M, K), R(K, N) * 0.05
    print("nn  M=%d K=%d N=%d  rel %.2e" % (M, K, N, rel(ops.gemm_nn(A, W), A.double() @ W.double())))
    Wt = R(N, K) * 0.05
    print("nt  M=%d N=%d K=%d  rel %.2e" % (M, N, K, rel(ops.gemm_nt(A, Wt), A.double() @ Wt.double().t())))
# layernorm backward at D = 1024
for rows, D in [(604, 1024), (604, 768), (2008, 768)]:
    x = torch.randn(rows, D, device="cuda") * 2 + 0.3
    g = torch.randn(D, device="cuda") * 0.1 + 1
    b = torch.randn(D, device="cuda") * 0.1
    dy = torch.randn(rows, D, device="cuda")
    dres = torch.randn(rows, D, device="cuda")
    xr = x.double().clone().requires_grad_(True)
    gr, br = g.double().clone().requires_grad_(True), b.double().clone().requires_grad_(True)
    y = torch.nn.functional.layer_norm(xr, (D,), gr, br, 1e-6)
    y.backward(dy.double())
    out, mean, rstd = ops.layernorm_fwd(x, g, b, rows, D)
    dg, db = torch.zeros(D, device="cuda"), torch.zeros(D, device="cuda")
    dys = torch.empty(rows, D, device="cuda")
    cs = torch.zeros(D, device="cuda")
    dx = ops.layernorm_bwd(dy, x, mean, rstd, g, dg, db, rows, D, dres=dres, dys=dys, colsum_out=cs)
    print("ln  rows=%d D=%d  fwd %.2e dx %.2e dgamma %.2e dbeta %.2e colsum %.2e" % (
        rows, D, rel(out, y.detach()), rel(dx, xr.grad + dres.double()), rel(dg, gr.grad), rel(db, br.grad), rel(cs, dys.double().sum(0))))
# attention S=4 N=151 H=16 ragged
S, N, H = 4, 151, 16
D = H * 64
qkv = ops.round_tf32(torch.randn(S * N, 3 * D, device="cuda"))
lengths = torch.tensor([151, 84, 151, 151], dtype=torch.int32, device="cuda")
q = qkv.double().clone().requires_grad_(True)
t = q.reshape(S, N, 3, H, 64).permute(2, 0, 3, 1, 4)
att = (t[0] @ t[1].transpose(-2, -1)) * 0.125 + ((torch.arange(N, device="cuda")[None] >= lengths[:, None]) * -10000.0)[:, None, None, :]
o_ref = (att.softmax(-1) @ t[2]).transpose(1, 2).reshape(S * N, D)
d_o = ops.round_tf32(torch.randn(S * N, D, device="cuda"))
o_ref.backward(d_o.double())
o, lse = ops.attention_fwd(qkv, S, N, H, lengths)
dqkv = ops.attention_bwd(qkv, o, d_o, lse, S, N, H, lengths)
print("attn S=4 N=151 H=16  o %.2e dqkv %.2e" % (rel(o, o_ref), rel(dqkv, q.grad)))
c = ops.colsum_acc(torch.randn(604, 4096, device="cuda"), torch.zeros(4096, device="cuda"))
