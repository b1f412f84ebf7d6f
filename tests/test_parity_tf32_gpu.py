"""GPU parity against the TF32-emulating oracle, at toy sizes and at the BASELINE shapes.

The CUDA path multiplies on tcgen05 with TF32 operands.  Rounding the GEMM operands to TF32 inside the *fp32* oracle
moves these models' gradients by 4-15 % (BatchNorm heads subtract nearly equal batch means; tools/tf32_sensitivity.py),
so a comparison with the fp32 golden vectors cannot tell a wiring bug from rounding noise.  Here the oracle rounds
with cvt.rna semantics at exactly the operands the kernels round (oracle/atst_oracle.py "TF32 operand emulation"),
which leaves only fp32 accumulation-order noise between the two sides:

    256-d outputs, loss, std statistics   <= 1e-3 relative (north_star), measured ~1e-5
    EVERY parameter gradient              <= 5e-3 relative l2 over the whole tensor (no sampling, no trimming)

The fp32-golden comparison stays in tests/test_parity_gpu.py as the documented "TF32 vs fp32" distance.
Live-oracle cases at BASELINE sizes cost 1-10 s of CPU each (the oracle is the live reference algorithm).
"""
import numpy as np
import pytest
import torch

from tests import util

pytestmark = pytest.mark.gpu

OUT_TOL = 1e-3
GRAD_TOL = 5e-3


def rel(a, b):
    a, b = torch.as_tensor(a).double().cpu(), torch.as_tensor(b).double().cpu()
    return ((a - b).norm() / b.norm().clamp_min(1e-30)).item()


def compare_grads(student, ref_student, tol=GRAD_TOL, label=""):
    """whole-tensor relative l2 error of every parameter gradient.  A tensor whose oracle gradient is numerically
    zero next to the others (the final LayerNorm bias, cancelled by the projector BatchNorm) is held on the
    absolute scale of the largest gradient instead.  Returns (worst error, its name, number compared)."""
    mine = dict(student.named_parameters())
    ref = dict(ref_student.named_parameters())
    assert set(mine) == set(ref)
    big = max(p.grad.norm().item() for p in ref.values() if p.grad is not None)
    worst, n = (0.0, ""), 0
    for name, rp in ref.items():
        g = mine[name].grad
        if rp.grad is None:  # e.g. ATST-clip's mask_embed: the reference never touches it
            assert g is None or not g.any(), "%s%s has a gradient, the reference has none" % (label, name)
            continue
        assert g is not None, label + name
        g, rg = g.detach().double().cpu(), rp.grad.double()
        assert g.shape == rg.shape
        n += 1
        if rg.norm().item() > 1e-3 * big:
            e = ((g - rg).norm() / rg.norm()).item()
        else:
            e = ((g - rg).norm() / big).item()
        if e > worst[0]:
            worst = (e, name)
    assert worst[0] < tol, "%sgradient of %s off by %.3e (tolerance %.1e)" % (label, worst[1], worst[0], tol)
    return worst[0], worst[1], n


def to_cpu_scales(groups):
    if groups is None:
        return None
    return [[None if b is None else (b[0].cpu(), b[1].cpu()) for b in blocks] for blocks in groups]


def run_pair(model, ref, crops, lengths, dp_teacher=None, dp_student=None, ncrops=2):
    """one step of the CUDA model and of the TF32-emulating oracle on the same inputs / weights / DropPath draws."""
    from oracle import atst_oracle as O
    kw = {}
    if dp_teacher is not None:
        kw = dict(dp_teacher=dp_teacher, dp_student=dp_student)
    loss, std_s, std_t = model(crops, lengths, **kw)
    loss.backward()
    s_out, t_out = model._rt.last_outputs
    with O.tf32_emulation():
        c_cpu, l_cpu = [c.cpu() for c in crops], [l.cpu() for l in lengths]
        t_ref = ref.teacher(c_cpu[:2], l_cpu[:2], to_cpu_scales(dp_teacher))
        s_ref = ref.student(c_cpu, l_cpu, to_cpu_scales(dp_student))
        rl, rs, rt = O.byol_loss(s_ref, t_ref, ncrops)
        rl.backward()
    assert rel(s_out, s_ref.detach()) < OUT_TOL and rel(t_out, t_ref.detach()) < OUT_TOL
    np.testing.assert_allclose(loss.item(), rl.item(), rtol=OUT_TOL)
    np.testing.assert_allclose(std_s.item(), rs.item(), rtol=OUT_TOL)
    np.testing.assert_allclose(std_t.item(), rt.item(), rtol=OUT_TOL)
    return rel(s_out, s_ref.detach()), rel(t_out, t_ref.detach())


def oracle_like(model, ncrops=2, frame=False):
    from oracle import atst_oracle as O
    enc = model.student.encoder
    cls = O.OracleFrameATST if frame else O.OracleATST
    kw = dict(embed_dim=enc.embed_dim, depth=enc.depth, num_heads=enc.num_heads)
    ref = cls(**kw) if frame else cls(ncrops=ncrops, **kw)
    ref.load_state_dict({k: v.detach().cpu() for k, v in model.state_dict().items()})
    ref.train()
    return ref


def injected_droppath(model, group_sizes_teacher, group_sizes_student, seed):
    from audiossl_b200.engine import droppath_scales
    enc = model.student.encoder
    gen = torch.Generator(device="cuda").manual_seed(seed)
    mk = lambda S: droppath_scales(enc.depth, enc.drop_path_rate, S, torch.device("cuda"), generator=gen)
    return [mk(S) for S in group_sizes_teacher], [mk(S) for S in group_sizes_student]


# --------------------------------------------------------------------------- toy sizes, deterministic fill
@pytest.mark.parametrize("case", ["tiny2", "tiny2b32", "tiny4", "tiny2dp", "small2"])
def test_step_matches_tf32_oracle(case):
    from audiossl_b200.models.atst import ATST
    c = util.CASES[case]
    m = ATST(arch=dict(embed_dim=c["dim"], depth=c["depth"], num_heads=c["heads"]), ncrops=c["ncrops"],
             drop_path_rate=c.get("drop_path", 0.0))
    util.load_det(m)
    m.cuda().train()
    ref = oracle_like(m, c["ncrops"])
    crops, lengths = util.make_inputs(case, c["B"], c["widths"], c["lens"])
    crops, lengths = [x.cuda() for x in crops], [x.cuda() for x in lengths]
    dp_t = dp_s = None
    if c.get("drop_path", 0.0) > 0:
        groups = [e - s for s, e in m.student.group_crops(crops)]
        dp_t, dp_s = injected_droppath(m, [2 * c["B"]], [g * c["B"] for g in groups], seed=7)
    es, et = run_pair(m, ref, crops, lengths, dp_t, dp_s, c["ncrops"])
    w = compare_grads(m.student, ref.student, label=case + ": ")
    print("%s: out err %.2e / %.2e, worst grad %.2e (%s) over %d tensors" % (case, es, et, w[0], w[1], w[2]))


# --------------------------------------------------------------------------- BASELINE shapes, live oracle
def _waves(B, n, seed=1234):
    return torch.randn(B, 1, n, generator=torch.Generator().manual_seed(seed)) * 0.1


def _mel(wav):
    from audiossl_b200 import ops
    return ops.mel_forward(wav.cuda())


def test_config2_base_10s_matches_tf32_oracle():
    """BASELINE config 2 shape: ATST-base, 10 s clips (251 tokens, D 768, 12 heads), reference initialisation,
    DropPath 0.1 with shared draws, ragged lengths; B = 4 clips (8 sequences per network)."""
    from audiossl_b200.models.atst import ATST
    torch.manual_seed(0)
    m = ATST(arch="base", ncrops=2, drop_path_rate=0.1).cuda().train()
    ref = oracle_like(m)
    B = 4
    crops = [_mel(_waves(B, 160000, 1)), _mel(_waves(B, 160000, 2))]
    lengths = [torch.tensor([1001, 801, 1001, 422]).cuda(), torch.tensor([1001, 1001, 640, 999]).cuda()]
    dp_t, dp_s = injected_droppath(m, [2 * B], [2 * B], seed=11)
    es, et = run_pair(m, ref, crops, lengths, dp_t, dp_s)
    w = compare_grads(m.student, ref.student, label="c2: ")
    print("c2 base/10s: out err %.2e / %.2e, worst grad %.2e (%s) over %d tensors" % (es, et, w[0], w[1], w[2]))


def test_config5_large_6s_matches_tf32_oracle():
    """BASELINE config 5 shape: ATST-large (24 layers, D 1024, 16 heads), 6 s clips (151 tokens), B = 2."""
    from audiossl_b200.models.atst import ATST
    torch.manual_seed(0)
    m = ATST(arch="large", ncrops=2, drop_path_rate=0.0).cuda().train()
    ref = oracle_like(m)
    B = 2
    crops = [_mel(_waves(B, 96000, 3)), _mel(_waves(B, 96000, 4))]
    lengths = [torch.tensor([601, 333]).cuda(), torch.tensor([601, 601]).cuda()]
    es, et = run_pair(m, ref, crops, lengths)
    w = compare_grads(m.student, ref.student, label="c5: ")
    print("c5 large/6s: out err %.2e / %.2e, worst grad %.2e (%s) over %d tensors" % (es, et, w[0], w[1], w[2]))


def test_config3_multicrop_matches_tf32_oracle():
    """BASELINE config 3 shape: ATST-base, 2 global crops of 601 frames + 6 local crops of 101 frames (ncrops = 8,
    two encoder calls per network pass), ragged lengths, DropPath 0.1; B = 3."""
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
    es, et = run_pair(m, ref, crops, lengths, dp_t, dp_s, ncrops=8)
    w = compare_grads(m.student, ref.student, label="c3: ")
    print("c3 base/2x6s+6x1s: out err %.2e / %.2e, worst grad %.2e (%s) over %d tensors" % (es, et, w[0], w[1], w[2]))


def test_config4_frame_base_10s_lightning_step_matches_tf32_oracle():
    """BASELINE config 4 shape through FrameATSTLightningModule.training_step: ATST-Frame base, 10 s clips (250 patches),
    block masks from random_mask.get_mask (ratio 0.65, span 5), the same mask for both views; B = 3."""
    from audiossl_b200.methods.atstframe import random_mask
    from audiossl_b200.methods.atstframe.model import FrameATSTLightningModule
    from oracle import atst_oracle as O
    torch.manual_seed(0)
    np.random.seed(0)
    lm = FrameATSTLightningModule(arch="base", learning_rate=1e-4, warmup_steps=2, max_steps=10, ema=0.99,
                                  drop_path_rate=0.0)
    lm.cuda().train()
    opt = lm.configure_optimizers()[0]
    lm.trainer.optimizers = [opt]
    ref = oracle_like(lm.model, frame=True)
    B = 3
    mel = _mel(_waves(B, 160000, 31))
    crops = [mel, _mel(_waves(B, 160000, 32))]
    lengths = [torch.tensor([1001, 1001, 700]).cuda()] * 2
    mask = random_mask.get_mask(B, 250, 0.65, no_overlap=False, min_length=5)
    assert mask.shape == (B, 250) and 0.3 < mask.float().mean().item() < 0.7
    masks = [mask.cuda(), mask.cuda()]
    loss = lm.training_step(((crops, lengths, masks), None), 0)
    loss.backward()
    s_out, t_out = lm.model._rt.last_outputs
    with O.tf32_emulation():
        args = ([c.cpu() for c in crops], [l.cpu() for l in lengths], [mask, mask])
        t_ref = ref._net(ref.teacher, *args, False)
        s_ref = ref._net(ref.student, *args, True)
        rl, rs, rt = O.byol_loss(s_ref, t_ref, 2)
        rl.backward()
    assert s_out.shape == s_ref.shape  # masked-row count and order are exact
    assert rel(s_out, s_ref.detach()) < OUT_TOL and rel(t_out, t_ref.detach()) < OUT_TOL
    np.testing.assert_allclose(loss.item(), rl.item(), rtol=OUT_TOL)
    np.testing.assert_allclose(lm.logged["std_frm_stu"].item(), rs.item(), rtol=OUT_TOL)
    np.testing.assert_allclose(lm.logged["std_frm_tea"].item(), rt.item(), rtol=OUT_TOL)
    valid = mask & (torch.arange(250)[None] < torch.tensor([250, 250, 175])[:, None])
    assert s_out.shape[0] == 2 * int(valid.sum())  # masked frames inside the valid length, both views
    w = compare_grads(lm.model.student, ref.student, label="c4: ")
    print("c4 frame-base/10s: loss %.6f vs %.6f, worst grad %.2e (%s) over %d tensors" % (loss.item(), rl.item(), w[0], w[1], w[2]))


# --------------------------------------------------------------------------- three optimizer steps, every tensor
def _three_steps(lm, ref, batches, frame):
    """Lightning-surface loop (schedule -> training_step -> backward -> fused HF-AdamW -> EMA hook) next to the
    oracle doing the same with its own restated transformers-AdamW under TF32 emulation.  Returns the per-step losses."""
    from oracle import atst_oracle as O
    opt = lm.configure_optimizers()[0]
    lm.trainer.optimizers = [opt]
    reg, _ = O.param_groups(ref.student)
    sp = dict(ref.student.named_parameters())
    state = {n: (torch.zeros_like(p), torch.zeros_like(p)) for n, p in sp.items()}
    for step, batch in enumerate(batches):
        lm.global_step = step
        loss = lm.training_step((tuple([t.cuda() for t in part] for part in batch), None), step)
        opt.zero_grad()
        loss.backward()
        opt.step()
        lm.on_train_batch_end(None, None, step)
        for p in ref.student.parameters():
            p.grad = None
        with O.tf32_emulation():
            rl = ref(*batch)[0]
            rl.backward()
        lr, wd = lm.mylr_scheduler[step], lm.wd_scheduler[step]
        for n, p in sp.items():
            if p.grad is None:  # transformers' AdamW: `if p.grad is None: continue`
                continue
            O.hf_adamw_step(p.data, p.grad, state[n][0], state[n][1], step + 1, lr, wd if n in reg else 0.0)
        ref.update_teacher(lm.ema_scheduler[step])
        np.testing.assert_allclose(loss.item(), rl.item(), rtol=1e-3, err_msg="step %d" % step)
    return opt


def _compare_updates(model, ref, init, label):
    """every student and teacher tensor: the accumulated update (value - initial value) against the oracle's.
    Adam normalises gradients element-wise, so an element whose gradient is near the eps = 1e-6 scale can step
    differently under fp32 noise; the bound is on the whole-tensor l2 of the update."""
    worst = (0.0, "")
    ref_sd = dict(ref.named_parameters())
    n = 0
    for name, p in model.named_parameters():
        mine = p.detach().cpu().double() - init[name].double()
        want = ref_sd[name].detach().double() - init[name].double()
        if want.norm().item() == 0.0:
            assert mine.norm().item() == 0.0, "%s%s moved, the reference leaves it untouched" % (label, name)
            continue
        e = ((mine - want).norm() / want.norm()).item()
        n += 1
        if e > worst[0]:
            worst = (e, name)
    assert worst[0] < 2e-2, "%supdate of %s off by %.3e" % (label, worst[1], worst[0])
    return worst, n


def test_three_training_steps_follow_the_oracle_every_tensor():
    from audiossl_b200.methods.atst.model import ATSTLightningModule
    torch.manual_seed(0)
    lm = ATSTLightningModule(arch="small", learning_rate=5e-4, warmup_steps=2, max_steps=10, ema=0.99,
                             drop_path_rate=0.0)
    util.load_det(lm.model)
    lm.cuda().train()
    ref = oracle_like(lm.model)
    init = {k: v.detach().cpu().clone() for k, v in lm.model.named_parameters()}
    B = 16
    batches = []
    for step in range(3):
        crops, lengths = util.make_inputs("loop%d" % step, B, [101, 101], [[101 - (i * 5) % 50 for i in range(B)],
                                                                              [101 - (i * 9) % 40 for i in range(B)]])
        batches.append((crops, lengths))
    opt = _three_steps(lm, ref, batches, frame=False)
    worst, n = _compare_updates(lm.model, ref, init, "clip: ")
    # the clip forward never reads mask_embed: no gradient, no Adam step, no weight decay (reference: grad is None)
    assert torch.equal(lm.model.student.encoder.mask_embed.detach().cpu(), init["student.encoder.mask_embed"])
    assert lm.model.student.encoder.mask_embed.grad is None
    # optimizer checkpoint in the transformers-AdamW layout: one entry per stepped parameter
    sd = opt.state_dict()
    n_trainable = sum(len(g["params"]) for g in opt.param_groups)
    assert len(sd["state"]) == n_trainable - 1 and all(int(s["step"]) == 3 for s in sd["state"].values())
    print("clip 3 steps: worst update error %.2e (%s) over %d tensors" % (worst[0], worst[1], n))


def test_three_frame_training_steps_follow_the_oracle_every_tensor():
    from audiossl_b200.methods.atstframe.model import FrameATSTLightningModule
    from tests.golden import detfill
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
        mask = torch.from_numpy(mk)
        batches.append((crops, lengths, [mask, mask]))
    _three_steps(lm, ref, batches, frame=True)
    worst, n = _compare_updates(lm.model, ref, init, "frame: ")
    assert not torch.equal(lm.model.student.encoder.mask_embed.detach().cpu(), init["student.encoder.mask_embed"])
    print("frame 3 steps: worst update error %.2e (%s) over %d tensors" % (worst[0], worst[1], n))
