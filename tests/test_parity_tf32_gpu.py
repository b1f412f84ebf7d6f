"""GPU parity of the whole training step against the TF32-emulating oracle, at toy sizes and at the BASELINE shapes.

The CUDA path multiplies on tcgen05 with TF32 operands.  Rounding the GEMM operands to TF32 inside the *fp32* oracle
moves these models' gradients by 4-15 % (BatchNorm heads subtract nearly equal batch means; tools/tf32_sensitivity.py),
so a comparison with the fp32 golden vectors cannot tell a wiring bug from rounding noise.  The oracle therefore rounds
with cvt.rna semantics at exactly the operands the kernels round (oracle/atst_oracle.py "TF32 operand emulation").

Even so an END-TO-END comparison cannot be tight: rounding is discontinuous, fp32-level noise flips values that sit on
a rounding boundary, and after four blocks the two sides are as far apart as TF32 is from fp32 (measured,
profiles/r02_layers_tf32_divergence.log; tests/linkwise.py explains).  So the step is checked LINK BY LINK, each link
recomputed by the oracle from the GPU's own recorded input to it (tests/linkwise.py):

    forward links (tokens, every block, final norm, heads, loss, BatchNorm buffers)   <= 5e-4   (measured 1e-6 .. 2e-4)
    backward links (same chain backwards, incl. DropPath, CLS rows, gather / scatter)  <= 1e-3
    EVERY parameter gradient, whole tensor, no sampling, no trimming                   <= 1e-3

and end to end the GPU must be no farther from the emulating oracle than TF32 is from fp32.  The comparison with the
fp32 golden vectors of the unmodified reference stays in tests/test_parity_gpu.py (the "TF32 vs fp32" distance), and
tests/test_parity_precise_gpu.py repeats it with the error-compensated 3xTF32 build, where it is tight.
Live-oracle cases at BASELINE sizes cost 5-30 s of CPU each (the oracle is the live reference algorithm).
"""
import numpy as np
import pytest
import torch

from audiossl_b200.engine import HEADS_3X, half_dgelu
from tests import linkwise, util

pytestmark = pytest.mark.gpu


def rel(a, b):
    a, b = torch.as_tensor(a).double().cpu(), torch.as_tensor(b).double().cpu()
    return ((a - b).norm() / b.norm().clamp_min(1e-30)).item()


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
def test_step_links_match_tf32_oracle(case):
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
    rep = linkwise.check_step(m, ref, crops, lengths, dp_teacher=dp_t, dp_student=dp_s, ncrops=c["ncrops"], label=case)
    print(rep.summary())


def test_end_to_end_distance_is_the_tf32_distance():
    """GPU vs emulating oracle, end to end, next to emulating oracle vs fp32 oracle: the same size (decorrelated TF32
    rounding), i.e. the CUDA path is as far from the emulation as TF32 arithmetic is from fp32 and no farther."""
    from audiossl_b200.models.atst import ATST
    from oracle import atst_oracle as O
    case = "tiny2b32"
    c = util.CASES[case]
    m = ATST(arch=dict(embed_dim=c["dim"], depth=c["depth"], num_heads=c["heads"]), ncrops=2, drop_path_rate=0.0)
    util.load_det(m)
    m.cuda().train()
    ref = oracle_like(m)
    crops, lengths = util.make_inputs(case, c["B"], c["widths"], c["lens"])
    loss, _, _ = m([x.cuda() for x in crops], [x.cuda() for x in lengths])
    loss.backward()
    s_gpu = m._rt.last_outputs[0].cpu()
    g_gpu = {n: p.grad.cpu().clone() for n, p in m.student.named_parameters() if p.grad is not None}

    def oracle_run(emulate):
        for p in ref.parameters():
            p.grad = None
        r2 = oracle_like(m)  # fresh BatchNorm buffers
        with O.tf32_emulation(emulate, heads=not HEADS_3X, gelu_half=half_dgelu()):
            t = r2.teacher(crops[:2], lengths[:2])
            s = r2.student(crops, lengths)
            l, _, _ = O.byol_loss(s, t, 2)
            l.backward()
        return l.item(), s.detach(), {n: p.grad.clone() for n, p in r2.student.named_parameters() if p.grad is not None}
    l32, s32, g32 = oracle_run(False)
    lem, sem, gem = oracle_run(True)
    assert abs(loss.item() - lem) < 1e-3 * abs(lem) and abs(loss.item() - l32) < 1e-3 * abs(l32)
    d_out_gpu, d_out_tf32 = rel(s_gpu, sem), rel(sem, s32)
    assert d_out_gpu < 1.5 * d_out_tf32 + 1e-4, (d_out_gpu, d_out_tf32)
    names = [n for n in g32 if g32[n].norm() > 1e-3 * max(v.norm() for v in g32.values())]
    e_gpu = np.median([rel(g_gpu[n], gem[n]) for n in names])
    e_tf32 = np.median([rel(gem[n], g32[n]) for n in names])
    assert e_gpu < 1.5 * e_tf32, (e_gpu, e_tf32)
    print("end to end: outputs GPU-emu %.2e vs emu-fp32 %.2e; median gradient distance GPU-emu %.2e vs emu-fp32 %.2e"
          % (d_out_gpu, d_out_tf32, e_gpu, e_tf32))


def _varied_waves(B, n, seed):
    """clips that differ from one another like real audio does (tones of different pitch and level over noise),
    unlike B draws of the same white noise, whose embeddings are nearly identical"""
    g = torch.Generator().manual_seed(seed)
    t = torch.arange(n) / 16000.0
    f = 100.0 * (60.0 ** torch.rand(B, generator=g))            # 100 Hz .. 6 kHz, log-uniform
    a = 0.02 + 0.5 * torch.rand(B, generator=g)
    w = a[:, None] * torch.sin(2 * torch.pi * f[:, None] * t[None, :] * (1.0 + 0.3 * t[None, :]))
    return (w + 0.02 * torch.rand(B, 1, generator=g) * torch.randn(B, n, generator=g))[:, None, :]


@pytest.mark.parametrize("kind", ["varied", "white_noise"])
def test_default_build_forward_against_the_fp32_reference_at_batch_64(kind):
    """north star: forward / loss within 1e-3 relative of the reference's fp32 path.  The default (TF32) build on
    ATST-small with the reference's initialisation, 64 clips x 2 views of 1 s (BatchNorm over 128 rows), against the
    fp32 oracle WITHOUT emulation.  Loss and statistics: 1e-3 asserted (measured 2e-5).  The 256-d outputs sit behind
    train-mode BatchNorm, which divides by the batch spread of its input: the bound on them is 1e-3 unless the batch
    itself amplifies - measured with the oracle as (output change) / (relative perturbation of the encoder outputs) -
    in which case it is 3 x that factor x the observed error of the encoder outputs (which is the TF32 rounding of
    the encoder GEMMs, ~5e-4; the heads themselves are 3xTF32).  64 draws of white noise are such a batch: every clip
    has the same embedding up to 1 %, so BatchNorm amplifies ~50-fold whatever separates two implementations."""
    from audiossl_b200.models.atst import ATST
    from oracle import atst_oracle as O
    torch.manual_seed(0)
    m = ATST(arch="small", ncrops=2, drop_path_rate=0.0).cuda().train()
    ref = oracle_like(m)
    B = 64
    mk = (lambda sd: _varied_waves(B, 16000, sd)) if kind == "varied" else (lambda sd: _waves(B, 16000, sd))
    crops = [_mel(mk(41)), _mel(mk(42))]
    lengths = [torch.full((B,), 101).cuda(), torch.randint(40, 102, (B,), generator=torch.Generator().manual_seed(1)).cuda()]
    rt = m._runtime(crops[0].device)
    rt.enc.debug = []
    try:
        loss, std_s, std_t = m(crops, lengths)
        cls_gpu = [x.cpu() for n, t, i, x in rt.enc.debug if n == "enc_out" and t == "s0"][0]
    finally:
        rt.enc.debug = None
    s_out, t_out = m._rt.last_outputs
    with torch.no_grad(), O.tf32_emulation(False):
        c_cpu, l_cpu = [c.cpu() for c in crops], [l.cpu() for l in lengths]
        cls_ref = ref.student.encoder(torch.cat(c_cpu), torch.cat(l_cpu))
        t_ref = ref.teacher(c_cpu, l_cpu)
        s_ref = ref.student(c_cpu, l_cpu)
        rl, rs, rt_ = O.byol_loss(s_ref, t_ref, 2)
        # amplification of the heads for this batch: relative output change per relative change of their input
        eps = 1e-3
        noisy = cls_ref * (1.0 + eps * torch.randn(cls_ref.shape, generator=torch.Generator().manual_seed(2)))
        import copy
        heads = copy.deepcopy(ref.student)
        amp = rel(heads.predictor(heads.projector(noisy)), copy.deepcopy(ref.student).predictor(
            copy.deepcopy(ref.student).projector(cls_ref))) / eps
    e_enc = rel(cls_gpu, cls_ref)
    es, et = rel(s_out, s_ref), rel(t_out, t_ref)
    bound = max(1e-3, 3.0 * amp * e_enc)
    print("default build vs fp32 oracle, B = 64, %s clips: encoder out %.2e, heads amplify x%.1f, student out %.2e, "
          "teacher out %.2e (bound %.1e), loss %.2e, std %.2e / %.2e"
          % (kind, e_enc, amp, es, et, bound, abs(loss.item() - rl.item()) / abs(rl.item()),
             abs(std_s.item() - rs.item()) / rs.item(), abs(std_t.item() - rt_.item()) / rt_.item()))
    assert e_enc < 1e-3
    assert es < bound and et < bound
    assert abs(loss.item() - rl.item()) < 1e-3 * abs(rl.item())
    assert abs(std_s.item() - rs.item()) < 1e-3 * rs.item() and abs(std_t.item() - rt_.item()) < 1e-3 * rt_.item()


# --------------------------------------------------------------------------- BASELINE shapes, live oracle
def _waves(B, n, seed=1234):
    return torch.randn(B, 1, n, generator=torch.Generator().manual_seed(seed)) * 0.1


def _mel(wav):
    from audiossl_b200 import ops
    return ops.mel_forward(wav.cuda())


def test_config2_base_10s_links_match_tf32_oracle():
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
    rep = linkwise.check_step(m, ref, crops, lengths, dp_teacher=dp_t, dp_student=dp_s, label="c2 base/10s")
    print(rep.summary())


def test_config5_large_6s_links_match_tf32_oracle():
    """BASELINE config 5 shape: ATST-large (24 layers, D 1024, 16 heads), 6 s clips (151 tokens), B = 2."""
    from audiossl_b200.models.atst import ATST
    torch.manual_seed(0)
    m = ATST(arch="large", ncrops=2, drop_path_rate=0.0).cuda().train()
    ref = oracle_like(m)
    B = 2
    crops = [_mel(_waves(B, 96000, 3)), _mel(_waves(B, 96000, 4))]
    lengths = [torch.tensor([601, 333]).cuda(), torch.tensor([601, 601]).cuda()]
    rep = linkwise.check_step(m, ref, crops, lengths, label="c5 large/6s")
    print(rep.summary())


def test_config3_multicrop_links_match_tf32_oracle():
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
    rep = linkwise.check_step(m, ref, crops, lengths, dp_teacher=dp_t, dp_student=dp_s, ncrops=8,
                              label="c3 base/2x6s+6x1s")
    print(rep.summary())


def test_config4_frame_base_10s_links_match_tf32_oracle():
    """BASELINE config 4 shape: ATST-Frame base, 10 s clips (250 patches), block masks from random_mask.get_mask
    (ratio 0.65, span 5), the same mask for both views, one clip shorter than the window; B = 3."""
    from audiossl_b200.methods.atstframe import random_mask
    from audiossl_b200.methods.atstframe.model import FrameATST
    torch.manual_seed(0)
    np.random.seed(0)
    m = FrameATST(arch="base", drop_path_rate=0.0).cuda().train()
    ref = oracle_like(m, frame=True)
    B = 3
    crops = [_mel(_waves(B, 160000, 31)), _mel(_waves(B, 160000, 32))]
    lengths = [torch.tensor([1001, 1001, 700]).cuda()] * 2
    mask = random_mask.get_mask(B, 250, 0.65, no_overlap=False, min_length=5)
    assert mask.shape == (B, 250) and 0.3 < mask.float().mean().item() < 0.7
    masks = [mask.cuda(), mask.cuda()]
    rep = linkwise.check_step(m, ref, crops, lengths, masks=masks, label="c4 frame-base/10s")
    valid = mask & (torch.arange(250)[None] < torch.tensor([250, 250, 175])[:, None])
    assert m._rt.last_outputs[0].shape[0] == 2 * int(valid.sum())  # masked frames inside the valid length, both views
    print(rep.summary())


@pytest.mark.parametrize("case", ["frame2", "frame2b16"])
def test_frame_step_links_match_tf32_oracle(case):
    from audiossl_b200.methods.atstframe.model import FrameATST
    from tests.golden import detfill
    B, lens = {"frame2": (4, [[101, 101, 77, 60]] * 2), "frame2b16": (16, [[101 - (i * 5) % 40 for i in range(16)]] * 2)}[case]
    m = FrameATST(arch=dict(embed_dim=128, depth=2, num_heads=2), drop_path_rate=0.0)
    util.load_det(m)
    m.cuda().train()
    ref = oracle_like(m, frame=True)
    crops, lengths = util.make_inputs(case, B, [101, 101], lens)
    mk = detfill.det_array(case + "/mask", (B, 25), 1.0, "uniform") > 0.0
    mk[:, 0] = True
    mask = torch.from_numpy(mk).cuda()
    rep = linkwise.check_step(m, ref, [c.cuda() for c in crops], [l.cuda() for l in lengths], masks=[mask, mask],
                              label=case)
    print(rep.summary())


# --------------------------------------------------------------------------- three optimizer steps, every tensor
def _three_steps(lm, ref, batches, frame, loss_rtol=5e-3, emulate=True):
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
        with O.tf32_emulation(emulate, heads=not HEADS_3X, gelu_half=half_dgelu()):
            rl = ref(*batch)[0]
            rl.backward()
        lr, wd = lm.mylr_scheduler[step], lm.wd_scheduler[step]
        for n, p in sp.items():
            if p.grad is None:  # transformers' AdamW: `if p.grad is None: continue`
                continue
            O.hf_adamw_step(p.data, p.grad, state[n][0], state[n][1], step + 1, lr, wd if n in reg else 0.0)
        ref.update_teacher(lm.ema_scheduler[step])
        np.testing.assert_allclose(loss.item(), rl.item(), rtol=loss_rtol, err_msg="step %d" % step)
    return opt


def _compare_updates(model, ref, init, label, tol, lr_sum=None):
    """every student and teacher tensor: the accumulated update (value - initial value) against the oracle's.
    Adam normalises gradients element-wise (the first steps are ~ lr * sign(g)), so the decorrelated TF32 rounding
    of the two sides (module docstring) shows up as sign flips of the small-gradient elements: the bound on the
    whole-tensor l2 of the update is loose here and tight (2e-2) in the 3xTF32 build
    (tests/test_parity_precise_gpu.py), which shares this loop."""
    worst = (0.0, "")
    ref_sd = dict(ref.named_parameters())
    n = 0
    for name, p in model.named_parameters():
        mine = p.detach().cpu().double() - init[name].double()
        want = ref_sd[name].detach().double() - init[name].double()
        if want.norm().item() == 0.0:
            assert mine.norm().item() == 0.0, "%s%s moved, the reference leaves it untouched" % (label, name)
            continue
        # Adam moves every element by about lr per step whatever the gradient's size, so a tensor whose true gradient
        # is numerically zero (the final-norm bias: BatchNorm in the projector removes any constant added to its
        # input) takes steps that are pure rounding noise in BOTH implementations - with transformers' eps = 1e-6 they
        # come out a few hundred times smaller than a regular step.  Measure against the larger of the reference
        # update and a fifth of a full-rate update, so that such tensors are checked for being small, not for agreeing
        floor = 0.0 if lr_sum is None else 0.2 * lr_sum * want.numel() ** 0.5
        e = ((mine - want).norm() / max(want.norm().item(), floor)).item()
        n += 1
        if e > worst[0]:
            worst = (e, name)
    assert worst[0] < tol, "%supdate of %s off by %.3e" % (label, worst[1], worst[0])
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
    worst, n = _compare_updates(lm.model, ref, init, "clip: ", tol=0.5, lr_sum=float(sum(lm.mylr_scheduler[:3])))
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
    worst, n = _compare_updates(lm.model, ref, init, "frame: ", tol=0.5, lr_sum=float(sum(lm.mylr_scheduler[:3])))
    assert not torch.equal(lm.model.student.encoder.mask_embed.detach().cpu(), init["student.encoder.mask_embed"])
    print("frame 3 steps: worst update error %.2e (%s) over %d tensors" % (worst[0], worst[1], n))
