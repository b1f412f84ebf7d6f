"""Link-by-link parity of one full training step (test infrastructure; imports the oracle).

Why not simply compare the step end to end with the TF32-emulating oracle?  Because rounding is discontinuous:
two implementations of the same rounded pipeline that differ by fp32 noise (1e-7) in a value near a TF32 rounding
boundary round it to different neighbours (5e-4 apart), and every later rounding stage multiplies the number of
such flips - measured on the GPU (profiles/r02_layers_tf32_divergence.log): 9e-6 after the first LayerNorm, 1e-4
after one block, saturating at the TF32 noise level (6e-4 relative on activations, percents on the BatchNorm-head
gradients) after four blocks.  An end-to-end comparison therefore cannot be tighter than the distance between the
TF32 and the fp32 model, however faithful the emulation.

What can be held tightly is every LINK of the chain, given the GPU's own input to that link: the engine records
its tensors at every block boundary of the forward and the backward pass (EncoderEngine.debug), and each link -
tokens, every transformer block, final norm, heads + loss, and the same links backwards, including DropPath
routing, CLS-strided norm, multi-crop accumulation and the frame model's row gather / scatter - is recomputed by
the TF32-emulating oracle from those recorded inputs and compared with what the GPU produced next.  Within one
link only a handful of rounding stages separate the two sides, so the agreement is 1e-5 .. 3e-4, and a wiring or
kernel error of a few per cent in any link, on any parameter, fails.  Every parameter's gradient belongs to
exactly one link, so all of them are covered.
"""
import copy

import torch
from torch import nn

from oracle import atst_oracle as O
from tests import conditioning


def rel(a, b):
    a, b = torch.as_tensor(a).detach().double().cpu(), torch.as_tensor(b).detach().double().cpu()
    return ((a - b).norm() / b.norm().clamp_min(1e-30)).item()


class Report:
    def __init__(self, label):
        self.label, self.rows = label, []

    def add(self, kind, name, err, tol):
        self.rows.append((kind, name, err, tol))

    def worst(self, kind):
        rows = [r for r in self.rows if r[0] == kind]
        return max(rows, key=lambda r: r[2] / r[3]) if rows else (kind, "", 0.0, 1.0)

    def check(self):
        import os
        if os.environ.get("ATST_LINK_DUMP"):
            for r in self.rows:
                print("LINK %-5s %-50s %.3e  tol %.1e %s" % (r[0], r[1], r[2], r[3], "" if r[2] < r[3] else "<<<"))
        bad = [r for r in self.rows if not (r[2] < r[3])]
        assert not bad, "%s: %d of %d links off: %s" % (
            self.label, len(bad), len(self.rows), "; ".join("%s %s %.2e (tol %.0e)" % r for r in bad[:8]))

    def summary(self):
        kinds = []
        for r in self.rows:
            if r[0] not in kinds:
                kinds.append(r[0])
        parts = ["%s %.1e (%s)" % (k, self.worst(k)[2], self.worst(k)[1]) for k in kinds]
        tail = "" if getattr(self, "kappa_heads", None) is None else " | heads-link kappa %.0f" % self.kappa_heads
        return "%s: %d links | worst per kind: %s%s" % (self.label, len(self.rows), " | ".join(parts), tail)


class _HeadsLink(nn.Module):
    """the heads + loss link as a module of its own (private copies), so that its conditioning can be measured:
    the link inputs are parameters too, they are perturbed along with the weights and yield the d_in gradient"""

    def __init__(self, ref, t_in, s_in, ncrops):
        super().__init__()
        self.tp = copy.deepcopy(ref.teacher.projector)
        self.sp = copy.deepcopy(ref.student.projector)
        self.pr = copy.deepcopy(ref.student.predictor)
        self.t_in, self.s_in, self.ncrops = nn.Parameter(t_in.detach().clone()), nn.Parameter(s_in.detach().clone()), ncrops

    @staticmethod
    def run(m):
        for p in m.parameters():
            p.grad = None
        s_out = m.pr(m.sp(m.s_in))
        loss, _, _ = O.byol_loss(s_out, m.tp(m.t_in).detach(), m.ncrops)
        loss.backward()
        grads = {"d_in": m.s_in.grad}
        grads.update({"projector." + n: p.grad for n, p in m.sp.named_parameters()})
        grads.update({"predictor." + n: p.grad for n, p in m.pr.named_parameters()})
        return s_out, grads


def _block_kappa(blk, xin, length, dp, d_out):
    """conditioning of one block link's parameter gradients: relative gradient change per relative change of the
    block's BRANCH output (y - x; the residual itself carries no arithmetic).  With few sequences per batch the
    weight gradients are sums over tokens that nearly cancel (BatchNorm in the heads makes the per-sequence loss
    gradients sum to zero), so the 1e-4 rounding-flip differences of the branch internals show up amplified."""
    def run(m):
        for p in m.parameters():
            p.grad = None
        x = xin.clone().requires_grad_(True)
        y = m(x, length, dp)
        y.backward(d_out)
        return (y - x).detach(), {n: p.grad for n, p in m.named_parameters()}
    kappa, _ = conditioning.gradient_kappa(blk, run)
    return kappa


def _cpu_scales(blocks):
    if blocks is None:
        return None
    return [None if b is None else (b[0].cpu(), b[1].cpu()) for b in blocks]


def check_step(model, ref, crops, lengths, masks=None, dp_teacher=None, dp_student=None, ncrops=2, label="",
               tol_fwd=5e-4, tol_bwd=1e-3, tol_grad=1.5e-3):
    """Runs model(...) + backward on the GPU with the engine's hooks on, then checks every link against `ref`
    (an oracle model holding the same weights) under TF32 emulation.  Returns the Report (already asserted)."""
    frame = masks is not None
    dev = crops[0].device
    rt = model._runtime(dev)
    rt.enc.debug = []
    try:
        if frame:
            loss, std_s, std_t = model(crops, lengths, masks)
        else:
            kw = {} if dp_teacher is None else dict(dp_teacher=dp_teacher, dp_student=dp_student)
            loss, std_s, std_t = model(crops, lengths, **kw)
        loss.backward()
        torch.cuda.synchronize()
        hooks = {(n, t, i): x.cpu() for n, t, i, x in rt.enc.debug}
    finally:
        rt.enc.debug = None
    rep = Report(label)
    for p in ref.parameters():
        p.grad = None
    c_cpu = [c.cpu() for c in crops]
    l_cpu = [l.cpu() for l in lengths]
    if frame:
        groups_t = groups_s = [(0, len(crops))]
        m_cpu = torch.cat([m.cpu() for m in masks]).bool()
    else:
        groups_s = model.student.group_crops(crops)
        groups_t = model.teacher.group_crops(crops[:2])
    depth = ref.student.encoder.blocks.__len__()
    D = ref.student.encoder.embed_dim
    cls_tok = 0 if frame else 1

    def norm_of(enc):
        return getattr(enc, enc.norm_name)

    def pass_info(net, groups, dps, prefix, mask_input):
        out = []
        for gi, (s, e) in enumerate(groups):
            mel, ln = torch.cat(c_cpu[s:e]), torch.cat(l_cpu[s:e])
            dp = None if dps is None else _cpu_scales(dps[gi])
            tag = "%s%d" % (prefix, gi)
            out.append((net.encoder, mel, ln, dp, tag, mask_input))
        return out

    passes_t = pass_info(ref.teacher, groups_t, dp_teacher, "t", False)
    passes_s = pass_info(ref.student, groups_s, dp_student, "s", True)

    from audiossl_b200.engine import HEADS_3X, half_dgelu  # the heads run as 3xTF32 (~fp32) products unless switched off
    with O.tf32_emulation(True, heads=not HEADS_3X, gelu_half=half_dgelu()):
        # ------------------------------------------------------------------ forward links, both networks
        enc_out = {"t": [], "s": []}
        plens = {}
        for enc, mel, ln, dp, tag, mask_input in passes_t + passes_s:
            with torch.no_grad():
                x0, plen = enc.tokens(mel, ln, m_cpu if frame else None, mask_input)
                S, N, _ = x0.shape
                plens[tag] = (plen, S, N)
                rep.add("fwd", tag + "/tokens", rel(hooks[("x_in", tag, 0)].view(S, N, D), x0), 1e-5)
                for i, blk in enumerate(enc.blocks):
                    y = blk(hooks[("x_in", tag, i)].view(S, N, D), plen + cls_tok, None if dp is None else dp[i])
                    rep.add("fwd", "%s/block%d" % (tag, i), rel(hooks[("x_in", tag, i + 1)].view(S, N, D), y), tol_fwd)
                xn = norm_of(enc)(hooks[("x_in", tag, depth)].view(S, N, D))
                got = hooks[("enc_out", tag, depth)]
                want = xn.reshape(S * N, D) if frame else xn[:, 0]
                if not HEADS_3X:  # a plain TF32 head GEMM gets its operand pre-rounded by the LayerNorm kernel
                    want = O.rna_tf32(want)
                rep.add("fwd", tag + "/final_norm", rel(got, want), 1e-4)
                if frame:
                    valid = m_cpu & (torch.arange(N)[None, :] < plen[:, None])
                    idx = torch.nonzero(valid.reshape(-1)).reshape(-1)
                    rows = hooks[("heads_in", tag, -1)]
                    assert torch.equal(rows, got[idx]), tag + ": gathered rows are not the masked valid frames"
                    enc_out[tag[0]].append(rows)
                else:
                    enc_out[tag[0]].append(got)
        # ------------------------------------------------------------------ heads + loss link (forward and backward)
        t_in = torch.cat(enc_out["t"])
        s_in = torch.cat(enc_out["s"]).clone().requires_grad_(True)
        t_out = ref.teacher.projector(t_in)
        s_out = ref.student.predictor(ref.student.projector(s_in))
        rl, rs, rt_ = O.byol_loss(s_out, t_out, 2 if frame else ncrops)
        rl.backward()
        g_s, g_t = rt.last_outputs
        # two or three TF32 GEMMs and train-mode BatchNorm over as few as six rows: a block's worth of rounding stages
        rep.add("fwd", "heads/student_out", rel(g_s, s_out), tol_fwd)
        rep.add("fwd", "heads/teacher_out", rel(g_t, t_out), tol_fwd)
        rep.add("fwd", "loss", abs(loss.item() - rl.item()) / abs(rl.item()), 1e-4)
        rep.add("fwd", "std_s", abs(std_s.item() - rs.item()) / abs(rs.item()), 1e-4)
        rep.add("fwd", "std_t", abs(std_t.item() - rt_.item()) / abs(rt_.item()), 1e-4)
        # the heads link is where the step is ill-conditioned (BYOL loss behind train-mode BatchNorm): its backward
        # tolerance follows the measured amplification of this link's own forward difference (tests/conditioning.py)
        e_heads = max(rel(g_s, s_out), rel(g_t, t_out))
        # (BatchNorm couples every tensor of the link, and one noise draw estimates kappa to a factor of ~2: the link's
        # largest kappa over two draws is used for all of its tensors, with a safety factor of 5)
        link = _HeadsLink(ref, t_in, s_in, 2 if frame else ncrops)
        k_heads = max(max(conditioning.gradient_kappa(link, _HeadsLink.run, seed=sd)[0].values()) for sd in (0, 1))
        kappa = {n: k_heads for n in ["d_in"] + ["projector." + n for n, _ in link.sp.named_parameters()]
                 + ["predictor." + n for n, _ in link.pr.named_parameters()]}
        rep.kappa_heads = k_heads
        rep.add("bwd", "heads/d_in", rel(hooks[("d_heads_in", "s", -1)], s_in.grad),
                conditioning.allowed(tol_bwd, k_heads, e_heads, safety=5.0))
        # ------------------------------------------------------------------ backward links, student encoder
        row = 0
        tol_of = {}  # parameter name -> tolerance widened by the measured conditioning of its link
        for enc, mel, ln, dp, tag, mask_input in passes_s:
            plen, S, N = plens[tag]
            d_out = hooks[("d_enc_out", tag, depth)]
            if frame:
                d_heads = hooks[("d_heads_in", "s", -1)]
                z = torch.zeros_like(d_out)
                z[idx] = d_heads
                assert torch.equal(z, d_out), "scatter of the head gradients"
            else:
                assert torch.equal(d_out, hooks[("d_heads_in", "s", -1)][row:row + S]), "CLS gradient rows of " + tag
                row += S
            xf = hooks[("x_in", tag, depth)].view(S, N, D).clone().requires_grad_(True)
            xn = norm_of(enc)(xf)
            (xn.reshape(S * N, D) if frame else xn[:, 0]).backward(d_out)
            rep.add("bwd", tag + "/final_norm", rel(hooks[("dx_in", tag, depth)].view(S, N, D), xf.grad), tol_bwd)
            for i in reversed(range(depth)):
                xin = hooks[("x_in", tag, i)].view(S, N, D).clone().requires_grad_(True)
                y = enc.blocks[i](xin, plen + cls_tok, None if dp is None else dp[i])
                d_next = hooks[("dx_in", tag, i + 1)].view(S, N, D)
                y.backward(d_next)
                rep.add("bwd", "%s/block%d" % (tag, i), rel(hooks[("dx_in", tag, i)].view(S, N, D), xin.grad), tol_bwd)
                branch = (y - xin).detach()
                e_branch = ((hooks[("x_in", tag, i + 1)].view(S, N, D) - y.detach()).norm() / branch.norm()).item()
                kap = _block_kappa(enc.blocks[i], xin.detach(), plen + cls_tok, None if dp is None else dp[i], d_next)
                for n_, k_ in kap.items():
                    key = "encoder.blocks.%d.%s" % (i, n_)
                    base = tol_grad if enc.blocks[i].get_parameter(n_).dim() > 1 else 3 * tol_grad
                    tol_of[key] = max(tol_of.get(key, 0.0), conditioning.allowed(base, k_, e_branch))
            x0, _ = enc.tokens(mel, ln, m_cpu if frame else None, mask_input)
            x0.backward(hooks[("dx_in", tag, 0)].view(S, N, D))
    # ---------------------------------------------------------------------- every parameter gradient (one link each)
    mine = dict(model.student.named_parameters())
    big = max(p.grad.norm().item() for p in ref.student.parameters() if p.grad is not None)
    n = 0
    for name, rp in ref.student.named_parameters():
        g = mine[name].grad
        if rp.grad is None:
            assert g is None, "%s has a gradient, the reference has none" % name
            continue
        assert g is not None, name
        n += 1
        tol = tol_grad
        if name in kappa:  # projector / predictor: part of the ill-conditioned heads link
            tol = conditioning.allowed(tol_grad, kappa[name], e_heads, safety=5.0)
        elif name in tol_of:  # transformer blocks: fixed tolerance unless the link's own conditioning says otherwise
            tol = tol_of[name]
        elif rp.dim() == 1 or rp.numel() == rp.shape[-1]:  # biases, norm scales, cls / mask tokens: sums over all rows
            tol = 3 * tol_grad
        rep.add("grad", name, ((g.detach().cpu().double() - rp.grad.double()).norm()
                               / max(rp.grad.norm().item(), 1e-3 * big)).item(), tol)
    assert n >= 4 * depth + 8
    # BatchNorm running statistics of the three heads (momentum update of batch statistics)
    prod_buf = dict(model.named_buffers())
    for name, b in ref.named_buffers():
        if "running" in name:
            rep.add("fwd", "buf/" + name, rel(prod_buf[name], b), tol_fwd)
    rep.check()
    return rep
