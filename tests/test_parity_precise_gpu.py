"""GPU parity of the 3xTF32 validation build against the fp32 vectors of the UNMODIFIED reference.

``audiossl_b200.set_precision("3xtf32")`` routes every call to libatst_b200_precise.so: the same sources, kernels and
engine as the product, compiled with -DATST_PRECISE - producers keep fp32 instead of rounding to TF32, and every
tensor-core product is error-compensated (hi/lo operand split; the GEMM kernels run unchanged over K-concatenated
operands, attention splits its fragments in registers).  That removes the TF32 rounding that makes an end-to-end
comparison of the default path loose (tests/test_parity_tf32_gpu.py explains), so here the WHOLE step - forward,
loss, backward through every kernel and every piece of engine wiring, optimizer, EMA - is held to the reference end to
end:

    256-d outputs, loss, logged statistics   <= 1e-3 relative (north_star), measured ~1e-5
    EVERY parameter gradient                 <= 5e-3 (golden samples: no trimming; live oracle: whole tensors)
    three optimizer steps, every tensor      update error <= 5e-2

Cases: the golden fixtures generated from /root/reference (tests/golden/make_golden.py) and, at the BASELINE shapes,
the live fp32 oracle (no emulation).
"""
import numpy as np
import pytest
import torch

from tests import conditioning, util
from tests.golden import detfill
from tests.test_parity_tf32_gpu import (_compare_updates, _mel, _three_steps, _waves, injected_droppath, oracle_like,
                                        rel)

pytestmark = pytest.mark.gpu

OUT_TOL, GRAD_TOL = 1e-3, 5e-3


@pytest.fixture(autouse=True)
def precise_build():
    import audiossl_b200
    prev = audiossl_b200.set_precision("3xtf32")
    from audiossl_b200 import _lib
    assert _lib.lib().atst_is_precise() == 1
    yield
    audiossl_b200.set_precision(prev)


# --------------------------------------------------------------------------- kernels of the validation build
@pytest.mark.parametrize("M,N,K", [(156, 384, 128), (1004, 2304, 768), (70, 4096, 128)])
def test_gemms_are_fp32_accurate_on_unrounded_inputs(M, N, K):
    from audiossl_b200 import ops
    torch.manual_seed(0)
    A = torch.randn(M, K, device="cuda")
    B = torch.randn(N, K, device="cuda") * 0.05
    bias = torch.randn(N, device="cuda")
    # 3xTF32 drops the lo*lo term (2^-22 relative per product) and accumulates in fp32: ~1e-5 at K = 768, against
    # 3e-4 for plain TF32 operands
    tol = 2e-5
    assert rel(ops.gemm_nt(A, B, bias=bias), A.double() @ B.double().t() + bias.double()) < tol
    W = torch.randn(K, N, device="cuda") * 0.05
    assert rel(ops.gemm_nn(A, W), A.double() @ W.double()) < tol
    G = torch.randn(M, N, device="cuda") * 0.1
    dW = torch.ones(N, K, device="cuda")
    ops.gemm_tn_acc(G, A, dW)
    assert rel(dW, 1.0 + G.double().t() @ A.double()) < tol


@pytest.mark.parametrize("S,N,H,lens", [(3, 151, 6, [191, 77, 0]), (5, 251, 12, [251, 1, 64, 65, 200]), (2, 26, 2, None)])
def test_attention_is_fp32_accurate(S, N, H, lens):
    from audiossl_b200 import ops
    torch.manual_seed(0)
    D = H * 64
    qkv = torch.randn(S * N, 3 * D, device="cuda")
    lengths = None if lens is None else torch.tensor(lens, dtype=torch.int32, device="cuda")
    q = qkv.double().clone().requires_grad_(True)
    t = q.reshape(S, N, 3, H, 64).permute(2, 0, 3, 1, 4)
    att = (t[0] @ t[1].transpose(-2, -1)) * 0.125
    if lengths is not None:
        att = att + ((torch.arange(N, device="cuda")[None] >= lengths[:, None]) * -10000.0)[:, None, None, :]
    o_ref = (att.softmax(-1) @ t[2]).transpose(1, 2).reshape(S * N, D)
    d_o = torch.randn(S * N, D, device="cuda")
    o_ref.backward(d_o.double())
    o, lse = ops.attention_fwd(qkv, S, N, H, lengths)
    dqkv = ops.attention_bwd(qkv, o, d_o, lse, S, N, H, lengths)
    assert rel(o, o_ref) < 3e-5 and rel(dqkv, q.grad) < 3e-5


def _sum_type_factor(name):
    """biases, norm scales and the cls / mask tokens are plain sums over every row (or sequence) of the batch, with
    cancellation: twice the tolerance of the matrices"""
    return 2.0 if name.endswith((".bias", "cls_token", "mask_embed")) or ".norm" in name or name.endswith("1.weight") else 1.0


# --------------------------------------------------------------------------- conditioning of the compared gradients
def oracle_kappa(m, crops, lengths, ncrops=2, masks=None, dp_t=None, dp_s=None):
    """per-parameter amplification (relative gradient change per relative output change) of this very step, measured
    with the fp32 oracle on the same weights and inputs (tests/conditioning.py)."""
    from oracle import atst_oracle as O
    ref = oracle_like(m, ncrops=ncrops, frame=masks is not None)
    cpu = lambda xs: [x.cpu() for x in xs]
    scales = lambda gs: None if gs is None else [[None if b is None else (b[0].cpu(), b[1].cpu()) for b in bl] for bl in gs]
    c_cpu, l_cpu = cpu(crops), cpu(lengths)

    def run(r):
        if masks is not None:
            args = (c_cpu, l_cpu, cpu(masks))
            t, s = r._net(r.teacher, *args, False), r._net(r.student, *args, True)
            loss = O.byol_loss(s, t, 2)[0]
        else:
            t, s = r.teacher(c_cpu[:2], l_cpu[:2], scales(dp_t)), r.student(c_cpu, l_cpu, scales(dp_s))
            loss = O.byol_loss(s, t, ncrops)[0]
        loss.backward()
        return s, {n: p.grad for n, p in r.student.named_parameters() if p.grad is not None}
    kappa, _ = conditioning.gradient_kappa(ref, run)
    return kappa


# --------------------------------------------------------------------------- golden fixtures of the reference
def check_grads_golden(m, g, case, kappa, e_out, tol=GRAD_TOL):
    """every parameter gradient against the reference's; tolerance max(5e-3, 3 * kappa * observed output error)"""
    stats = []
    for name, p in m.student.named_parameters():
        key = case + "/grad/" + name
        if key + "/idx" not in g.files:
            assert p.grad is None, name  # the reference produced no gradient for it (mask_embed in ATST-clip)
            continue
        err, ref_norm = util.sample_rel_err(p.grad.cpu().numpy(), g, key)
        stats.append((name, err, ref_norm))
    assert len(stats) > 20
    big = max(r for _, _, r in stats)
    worst = (0.0, "", 0.0)
    for n, e, r in stats:
        t = conditioning.allowed(tol * _sum_type_factor(n), kappa[n], e_out)
        e_eff = e if r > 1e-3 * big else e * r / (1e-3 * big)  # tiny tensors: against 1e-3 of the largest gradient
        assert e_eff < t, "gradient of %s off by %.3e (tolerance %.1e, kappa %.0f)" % (n, e_eff, t, kappa[n])
        if e_eff / t > worst[2]:
            worst = (e_eff, n, e_eff / t)
    return worst


@pytest.mark.parametrize("case", ["tiny2", "tiny2b32", "tiny4", "small2", "tiny2dp"])
def test_atst_step_matches_reference_golden(case):
    from audiossl_b200.models.atst import ATST
    g = util.gold("atst.npz")
    c = util.CASES[case]
    m = ATST(arch=dict(embed_dim=c["dim"], depth=c["depth"], num_heads=c["heads"]), ncrops=c["ncrops"],
             drop_path_rate=c.get("drop_path", 0.0))
    util.load_det(m)
    m.cuda().train()
    crops, lengths = util.make_inputs(case, c["B"], c["widths"], c["lens"])
    kw = {}
    if case == "tiny2dp":  # replay the reference's recorded torch.rand stream
        keep = [1.0 - x for x in torch.linspace(0, c["drop_path"], c["depth"]).tolist()]
        dp_t, dp_s = util.dp_scales_from_rand(g["tiny2dp/rand"], c["depth"], keep)
        cu = lambda groups: [[None if b is None else (b[0].cuda(), b[1].cuda()) for b in blocks] for blocks in groups]
        kw = dict(dp_teacher=cu(dp_t), dp_student=cu(dp_s))
    crops, lengths = [x.cuda() for x in crops], [x.cuda() for x in lengths]
    loss, std_s, std_t = m(crops, lengths, **kw)
    loss.backward()
    s_out, t_out = m._rt.last_outputs
    es, et = rel(s_out, g[case + "/student_out"]), rel(t_out, g[case + "/teacher_out"])
    assert es < OUT_TOL and (case == "tiny2dp" or et < OUT_TOL)
    np.testing.assert_allclose(loss.item(), g[case + "/loss"], rtol=1e-4)
    if case != "tiny2dp":
        np.testing.assert_allclose(std_s.item(), g[case + "/std_s"], rtol=1e-4)
        np.testing.assert_allclose(std_t.item(), g[case + "/std_t"], rtol=1e-4)
        for name, b in m.named_buffers():
            if "running" in name:
                util.check_summary(b.cpu().numpy(), g, case + "/buf/" + name, rtol=1e-3, atol=1e-4)
    kappa = oracle_kappa(m, crops, lengths, ncrops=c["ncrops"], dp_t=kw.get("dp_teacher"), dp_s=kw.get("dp_student"))
    w = check_grads_golden(m, g, case, kappa, es)
    print("%s (3xTF32 vs reference fp32): outputs %.1e / %.1e, worst gradient %.1e (%s, kappa %.0f)"
          % (case, es, et, w[0], w[1], kappa[w[1]]))


@pytest.mark.parametrize("case", ["frame2", "frame2b16"])
def test_frame_step_matches_reference_golden(case):
    from audiossl_b200.methods.atstframe.model import FrameATST
    g = util.gold("frame.npz")
    B, lens = {"frame2": (4, [[101, 101, 77, 60]] * 2), "frame2b16": (16, [[101 - (i * 5) % 40 for i in range(16)]] * 2)}[case]
    m = FrameATST(arch=dict(embed_dim=128, depth=2, num_heads=2), drop_path_rate=0.0)
    util.load_det(m)
    m.cuda().train()
    crops, lengths = util.make_inputs(case, B, [101, 101], lens)
    mk = detfill.det_array(case + "/mask", (B, 25), 1.0, "uniform") > 0.0
    mk[:, 0] = True
    mask = torch.from_numpy(mk).cuda()
    crops, lengths = [c.cuda() for c in crops], [l.cuda() for l in lengths]
    loss, std_s, std_t = m(crops, lengths, [mask, mask])
    loss.backward()
    s_out, t_out = m._rt.last_outputs
    assert tuple(s_out.shape) == g[case + "/student_out"].shape
    es = rel(s_out, g[case + "/student_out"])
    assert es < OUT_TOL and rel(t_out, g[case + "/teacher_out"]) < OUT_TOL
    np.testing.assert_allclose(loss.item(), g[case + "/loss"], rtol=1e-4)
    np.testing.assert_allclose(std_s.item(), g[case + "/std_s"], rtol=1e-4)
    kappa = oracle_kappa(m, crops, lengths, masks=[mask, mask])
    w = check_grads_golden(m, g, case, kappa, es)
    print("%s (3xTF32 vs reference fp32): outputs %.1e, worst gradient %.1e (%s, kappa %.0f)"
          % (case, es, w[0], w[1], kappa[w[1]]))


# --------------------------------------------------------------------------- BASELINE shapes, live fp32 oracle
def compare_with_fp32_oracle(m, ref, crops, lengths, ncrops=2, dp_t=None, dp_s=None, masks=None, label=""):
    from oracle import atst_oracle as O
    cpu_scales = lambda gs: None if gs is None else [[None if b is None else (b[0].cpu(), b[1].cpu()) for b in bl] for bl in gs]
    if masks is not None:
        loss, std_s, std_t = m(crops, lengths, masks)
    else:
        kw = {} if dp_t is None else dict(dp_teacher=dp_t, dp_student=dp_s)
        loss, std_s, std_t = m(crops, lengths, **kw)
    loss.backward()
    s_out, t_out = m._rt.last_outputs
    c_cpu, l_cpu = [c.cpu() for c in crops], [l.cpu() for l in lengths]
    if masks is not None:
        args = (c_cpu, l_cpu, [k.cpu() for k in masks])
        t_ref = ref._net(ref.teacher, *args, False)
        s_ref = ref._net(ref.student, *args, True)
        ncrops = 2
    else:
        t_ref = ref.teacher(c_cpu[:2], l_cpu[:2], cpu_scales(dp_t))
        s_ref = ref.student(c_cpu, l_cpu, cpu_scales(dp_s))
    rl, rs, rt = O.byol_loss(s_ref, t_ref, ncrops)
    rl.backward()
    es, et = rel(s_out, s_ref.detach()), rel(t_out, t_ref.detach())
    assert s_out.shape == s_ref.shape and es < OUT_TOL and et < OUT_TOL, (es, et)
    np.testing.assert_allclose(loss.item(), rl.item(), rtol=1e-4)
    np.testing.assert_allclose(std_s.item(), rs.item(), rtol=1e-4)
    np.testing.assert_allclose(std_t.item(), rt.item(), rtol=1e-4)
    mine = dict(m.student.named_parameters())
    big = max(p.grad.norm().item() for p in ref.student.parameters() if p.grad is not None)
    kappa = oracle_kappa(m, crops, lengths, ncrops=ncrops, masks=masks, dp_t=dp_t, dp_s=dp_s)
    k_max = max(kappa.values())
    worst, n = (0.0, "", 0.0), 0
    for name, rp in ref.student.named_parameters():
        if rp.grad is None:
            assert mine[name].grad is None, name
            continue
        e = ((mine[name].grad.cpu().double() - rp.grad.double()).norm() / max(rp.grad.norm().item(), 1e-3 * big)).item()
        # few clips per batch at these shapes: every gradient is a strongly cancelling sum and one noise draw
        # estimates a tensor's kappa to a factor of ~2, so the step's largest kappa bounds all of its tensors
        t = conditioning.allowed(GRAD_TOL * _sum_type_factor(name), k_max, es, safety=5.0)
        assert e < t, "%s: gradient of %s off by %.3e (tolerance %.1e, kappa %.0f)" % (label, name, e, t, kappa[name])
        n += 1
        if e / t > worst[2]:
            worst = (e, name, e / t)
    print("%s (3xTF32 vs fp32 oracle): outputs %.1e / %.1e, worst of %d gradients %.1e (%s; kappa %.0f, step max %.0f)"
          % (label, es, et, n, worst[0], worst[1], kappa[worst[1]], k_max))


def test_config2_base_10s_matches_fp32_oracle():
    from audiossl_b200.models.atst import ATST
    torch.manual_seed(0)
    m = ATST(arch="base", ncrops=2, drop_path_rate=0.1).cuda().train()
    ref = oracle_like(m)
    B = 4
    crops = [_mel(_waves(B, 160000, 1)), _mel(_waves(B, 160000, 2))]
    lengths = [torch.tensor([1001, 801, 1001, 422]).cuda(), torch.tensor([1001, 1001, 640, 999]).cuda()]
    dp_t, dp_s = injected_droppath(m, [2 * B], [2 * B], seed=11)
    compare_with_fp32_oracle(m, ref, crops, lengths, dp_t=dp_t, dp_s=dp_s, label="c2 base/10s")


def test_config5_large_6s_matches_fp32_oracle():
    from audiossl_b200.models.atst import ATST
    torch.manual_seed(0)
    m = ATST(arch="large", ncrops=2, drop_path_rate=0.0).cuda().train()
    ref = oracle_like(m)
    B = 2
    crops = [_mel(_waves(B, 96000, 3)), _mel(_waves(B, 96000, 4))]
    lengths = [torch.tensor([601, 333]).cuda(), torch.tensor([601, 601]).cuda()]
    compare_with_fp32_oracle(m, ref, crops, lengths, label="c5 large/6s")


def test_config3_multicrop_matches_fp32_oracle():
    from audiossl_b200.models.atst import ATST
    torch.manual_seed(0)
    m = ATST(arch="base", ncrops=8, drop_path_rate=0.1).cuda().train()
    ref = oracle_like(m, ncrops=8)
    B = 3
    crops = [_mel(_waves(B, 96000, 10 + i)) for i in range(2)] + [_mel(_waves(B, 16000, 20 + i)) for i in range(6)]
    gl = torch.Generator().manual_seed(5)
    lengths = [torch.randint(300, 602, (B,), generator=gl).cuda() for _ in range(2)] + \
              [torch.randint(50, 102, (B,), generator=gl).cuda() for _ in range(6)]
    lengths[0][0], lengths[2][0] = 601, 101
    dp_t, dp_s = injected_droppath(m, [2 * B], [2 * B, 6 * B], seed=13)
    compare_with_fp32_oracle(m, ref, crops, lengths, ncrops=8, dp_t=dp_t, dp_s=dp_s, label="c3 base/2x6s+6x1s")


def test_config4_frame_base_10s_matches_fp32_oracle():
    from audiossl_b200.methods.atstframe import random_mask
    from audiossl_b200.methods.atstframe.model import FrameATST
    torch.manual_seed(0)
    np.random.seed(0)
    m = FrameATST(arch="base", drop_path_rate=0.0).cuda().train()
    ref = oracle_like(m, frame=True)
    B = 3
    crops = [_mel(_waves(B, 160000, 31)), _mel(_waves(B, 160000, 32))]
    lengths = [torch.tensor([1001, 1001, 700]).cuda()] * 2
    mask = random_mask.get_mask(B, 250, 0.65, no_overlap=False, min_length=5).cuda()
    compare_with_fp32_oracle(m, ref, crops, lengths, masks=[mask, mask], label="c4 frame-base/10s")


# --------------------------------------------------------------------------- three optimizer steps, every tensor
def test_three_training_steps_every_tensor():
    from audiossl_b200.methods.atst.model import ATSTLightningModule
    torch.manual_seed(0)
    lm = ATSTLightningModule(arch="small", learning_rate=5e-4, warmup_steps=2, max_steps=10, ema=0.99,
                             drop_path_rate=0.0)
    util.load_det(lm.model)
    lm.cuda().train()
    ref = oracle_like(lm.model)
    init = {k: v.detach().cpu().clone() for k, v in lm.model.named_parameters()}
    B = 16
    batches = [util.make_inputs("loop%d" % s, B, [101, 101], [[101 - (i * 5) % 50 for i in range(B)],
                                                               [101 - (i * 9) % 40 for i in range(B)]]) for s in range(3)]
    _three_steps(lm, ref, batches, frame=False, loss_rtol=1e-4, emulate=False)
    worst, n = _compare_updates(lm.model, ref, init, "clip: ", tol=5e-2, lr_sum=float(sum(lm.mylr_scheduler[:3])))
    assert torch.equal(lm.model.student.encoder.mask_embed.detach().cpu(), init["student.encoder.mask_embed"])
    print("clip 3 steps (3xTF32 vs fp32 oracle): worst update error %.2e (%s) over %d tensors" % (worst[0], worst[1], n))


def test_three_frame_training_steps_every_tensor():
    from audiossl_b200.methods.atstframe.model import FrameATSTLightningModule
    torch.manual_seed(0)
    lm = FrameATSTLightningModule(arch="small", learning_rate=5e-4, warmup_steps=2, max_steps=10, ema=0.99,
                                  drop_path_rate=0.0)
    util.load_det(lm.model)
    lm.cuda().train()
    ref = oracle_like(lm.model, frame=True)
    init = {k: v.detach().cpu().clone() for k, v in lm.model.named_parameters()}
    B = 8
    batches = []
    for step in range(3):
        crops, lengths = util.make_inputs("floop%d" % step, B, [101, 101], [[101 - (i * 5) % 40 for i in range(B)]] * 2)
        mk = detfill.det_array("floop%d/mask" % step, (B, 25), 1.0, "uniform") > 0.0
        mk[:, 0] = True
        batches.append((crops, lengths, [torch.from_numpy(mk)] * 2))
    _three_steps(lm, ref, batches, frame=True, loss_rtol=1e-4, emulate=False)
    worst, n = _compare_updates(lm.model, ref, init, "frame: ", tol=5e-2, lr_sum=float(sum(lm.mylr_scheduler[:3])))
    print("frame 3 steps (3xTF32 vs fp32 oracle): worst update error %.2e (%s) over %d tensors" % (worst[0], worst[1], n))
