"""GPU parity of the reference-facing API either side of the training step:
the train transforms end to end and the embedding API, against vectors produced by the UNMODIFIED reference
(tests/golden/make_golden.py gen_transform / gen_embed).  The transforms draw from the same host generators in the
same order as the reference, so equal seeds give the reference's crops, mixing partners, resize rectangles and masks;
what is left is the arithmetic of the CUDA kernels (fused mel, log-mixup-exp, bicubic resize), held to the mel
tolerance."""
import random

import numpy as np
import pytest
import torch

from tests import util
from tests.golden import detfill

pytestmark = pytest.mark.gpu

TOL = 5e-4  # absolute, in the normalised log-mel units of the reference's MinMax (range about [-1, 1])


def _wav(name, n):
    return torch.from_numpy(detfill.det_array(name, (1, n), 0.1)).cuda()


def test_atst_train_transform_variable_length_views_match_reference():
    """anchor / positive lengths U(0.7, 1.0) s: RandomCrop, the fused mel, Mixup against a bank of clips of other
    lengths (all three log_mixup_exp branches), RandomResizeCrop, right padding - five consecutive calls."""
    from audiossl_b200.methods.atst.transform import ATSTTrainTransform
    g = util.gold("transform.npz")
    random.seed(5)
    np.random.seed(5)
    tf = ATSTTrainTransform(anchor_len=(0.7, 1.0), positive_len=(0.7, 1.0))
    seen = set()
    for k in range(5):
        crops, lengths = tf(_wav("tf/wav%d" % k, 24000))
        assert list(lengths) == list(g["clipvar/%d/lengths" % k])
        for v in range(2):
            ref = g["clipvar/%d/crop%d" % (k, v)]
            assert tuple(crops[v].shape) == ref.shape and crops[v].is_cuda
            np.testing.assert_allclose(crops[v].cpu().numpy(), ref, atol=TOL, rtol=0)
        seen.add(lengths[0] != lengths[1])
    assert True in seen  # views of different lengths did occur


def test_atst_train_transform_recipe_defaults_match_reference():
    """the recipe's 6 s views of 10 s clips (and of a 5 s clip: zero-padded waveform)."""
    from audiossl_b200.methods.atst.transform import ATSTTrainTransform
    g = util.gold("transform.npz")
    random.seed(6)
    np.random.seed(6)
    tf = ATSTTrainTransform()
    for k, n in enumerate((160000, 160000, 80000)):
        crops, lengths = tf(_wav("tf6/wav%d" % k, n))
        assert list(lengths) == list(g["clip6/%d/lengths" % k]) == [601, 601]
        for v in range(2):
            assert list(crops[v].shape) == list(g["clip6/%d/shape%d" % (k, v)])
            util.check_summary(crops[v].cpu().numpy(), g, "clip6/%d/crop%d" % (k, v), rtol=0, atol=TOL, scale=1.0)


def test_frame_train_transform_matches_reference():
    """ATST-Frame: one crop, two independently augmented views (frequency-only warp), one mask for both."""
    from audiossl_b200.methods.atstframe.transform import FrameATSTTrainTransform
    g = util.gold("transform.npz")
    random.seed(7)
    np.random.seed(7)
    torch.manual_seed(7)
    tf = FrameATSTTrainTransform(anchor_len=1.0, mask_type="random", mask_ratio=0.75)
    for k in range(3):
        crops, lengths, masks = tf(_wav("tff/wav%d" % k, 20000))
        assert list(lengths) == list(g["frame/%d/lengths" % k])
        assert masks[0] is masks[1] and np.array_equal(masks[0].cpu().numpy(), g["frame/%d/mask" % k])
        for v in range(2):
            np.testing.assert_allclose(crops[v].cpu().numpy(), g["frame/%d/crop%d" % (k, v)], atol=TOL, rtol=0)


def test_block_mask_statistics_and_transform_contract():
    """mask_type="block" (the ATST-Frame recipe): spans of 5 frames, overlap allowed, about half of the frames masked
    (SURVEY.md a17); the transform returns it for both views."""
    from audiossl_b200.methods.atstframe import random_mask
    from audiossl_b200.methods.atstframe.transform import FrameATSTTrainTransform
    np.random.seed(0)
    m = random_mask.get_mask(64, 250, 0.65, no_overlap=False, min_length=5)
    assert m.shape == (64, 250) and m.dtype == torch.bool
    frac = m.float().mean(1)
    assert 0.40 < frac.mean().item() < 0.60 and frac.min().item() > 0.3
    runs = []
    for row in m.numpy():
        edges = np.flatnonzero(np.diff(np.concatenate([[0], row.astype(np.int8), [0]])))
        runs += list(edges[1::2] - edges[0::2])
    assert min(runs) >= 5  # every masked run is a union of length-5 spans
    tf = FrameATSTTrainTransform(anchor_len=2.0, mask_type="block", mask_ratio=0.65, mask_len=5)
    crops, lengths, masks = tf(_wav("tfb/wav", 40000))
    assert lengths == [201, 201] and masks[0].shape == (50,) and crops[0].shape == (1, 64, 201)


def _checkpoint(tmp_path, arch="small"):
    """a Lightning-format checkpoint of the reference's key layout with name-derived weights (the fixture generator
    wrote the same one for the reference's load_model)."""
    from audiossl_b200.methods.atstframe.model import FrameATSTLightningModule
    hp = dict(arch=arch, learning_rate=5e-4, warmup_steps=10, max_steps=100, ema=0.99)
    lm = FrameATSTLightningModule(**hp)
    util.load_det(lm)
    path = str(tmp_path / "last.ckpt")
    torch.save({"state_dict": {k: v.cpu() for k, v in lm.state_dict().items()}, "hyper_parameters": hp}, path)
    return path


def test_embedding_api_matches_reference(tmp_path):
    """load_model -> get_scene_embedding / get_timestamp_embedding on 12 s clips (a full 1001-frame chunk + a tail)
    and on a [1, N] clip shorter than one chunk: audiossl/methods/atstframe/embedding.py:19-127."""
    from audiossl_b200.methods.atstframe import embedding as E
    g = util.gold("embed.npz")
    enc = E.load_model(_checkpoint(tmp_path))
    assert enc.scene_embedding_size == 384 * 2 * 12 and enc.timestamp_embedding_size == 384 * 12
    assert enc.sample_rate == 16000 and enc.hyper_param["arch"] == "small" and not enc.training
    n = 16000 * 12 + 800
    audio = torch.stack([torch.from_numpy(detfill.signal("noise", n)), torch.from_numpy(detfill.signal("chirp", n))])[:, None].cuda()
    rel = lambda a, b: float(np.linalg.norm(np.asarray(a, np.float64) - b) / np.linalg.norm(b))
    scene = E.get_scene_embedding(audio, enc)
    assert tuple(scene.shape) == g["scene"].shape == (2, 4608)
    assert rel(scene.cpu().numpy(), g["scene"]) < 1e-3
    ts, stamps = E.get_timestamp_embedding(audio, enc)
    assert list(ts.shape) == list(g["ts/shape"]) == [2, 301, 4608]
    np.testing.assert_array_equal(stamps.numpy(), g["stamps"])
    assert rel(ts[:, ::37, :].cpu().numpy()[:, :, ::13], g["ts/rows"]) < 1e-3
    err, _ = util.sample_rel_err(ts.cpu().numpy(), g, "ts")
    assert err < 1e-3
    scene1 = E.get_scene_embedding(audio[0, :, :16000 * 3], enc)
    assert tuple(scene1.shape) == (1, 4608) and rel(scene1.cpu().numpy(), g["scene1"]) < 1e-3
    with pytest.raises(RuntimeError):
        E.get_scene_embedding(audio.cpu(), enc)


def test_batched_train_transform_contract_and_statistics():
    """BatchedATSTTrainTransform: the batch contract of training_step, the un-augmented path equal to the fused mel
    of the drawn windows, and augmentation statistics in the range the per-sample reference transform produces."""
    from audiossl_b200 import ops
    from audiossl_b200.methods.atst.transform import BatchedATSTTrainTransform
    B, n = 6, 40000
    wav = (torch.randn(B, 1, n, generator=torch.Generator().manual_seed(0)) * 0.1).cuda()
    rng = np.random.RandomState(3)
    tf = BatchedATSTTrainTransform(anchor_len=(1.0, 1.0), positive_len=(1.5, 1.5), rng=rng, augment=False)
    crops, lengths = tf(wav)
    # right-padded to max_frames + 1 = (1.5 s * 16000) // 160 + 1 frames, as the reference's F.pad arithmetic gives
    assert [tuple(c.shape) for c in crops] == [(B, 1, 64, 151)] * 2 and all(c.is_cuda for c in crops)
    assert lengths[0].tolist() == [101] * B and lengths[1].tolist() == [151] * B and lengths[0].dtype == torch.int64
    # replay the draws: uniform(anchor), B window starts, uniform(positive), B window starts
    rng2 = np.random.RandomState(3)
    rng2.uniform(1.0, 1.0)
    s1 = rng2.randint(0, n - 16000 + 1, B)
    rng2.uniform(1.5, 1.5)
    s2 = rng2.randint(0, n - 24000 + 1, B)
    for b in range(B):
        m1 = ops.mel_forward(wav[b, :, s1[b]:s1[b] + 16000].contiguous())
        m2 = ops.mel_forward(wav[b, :, s2[b]:s2[b] + 24000].contiguous())
        np.testing.assert_allclose(crops[0][b, :, :, :101].cpu().numpy(), m1.cpu().numpy(), atol=1e-5)
        np.testing.assert_allclose(crops[1][b, :, :, :151].cpu().numpy(), m2.cpu().numpy(), atol=1e-5)
    assert float(crops[0][..., 101:].abs().sum()) == 0.0
    # augmentations on: same shapes, finite, a memory bank that fills, outputs that differ from the plain mel
    tf = BatchedATSTTrainTransform(anchor_len=(1.0, 1.0), positive_len=(1.0, 1.0), rng=np.random.RandomState(4))
    for _ in range(3):
        crops, lengths = tf(wav)
    assert tf.mixup[0].size == 3 * B and tuple(crops[0].shape) == (B, 1, 64, 101)
    assert torch.isfinite(crops[0]).all() and torch.isfinite(crops[1]).all()
    assert -1.5 < float(crops[0].mean()) < 1.0 and float(crops[0].std()) > 0.05


def test_data_module_and_launcher_on_synthetic_clips(tmp_path):
    """the train.py-equivalent launcher end to end on one GPU: synthetic clips -> host loader -> pinned H2D prefetch
    -> device transform -> training steps -> Lightning-layout checkpoint -> resume."""
    from audiossl_b200.methods.atst import train as T
    args = T.build_parser().parse_args(["--save_path", str(tmp_path), "--nproc", "1", "--arch", "small",
                                        "--batch_size_per_gpu", "4", "--num_workers", "0", "--synthetic_clips", "16",
                                        "--train_len", "1.0", "--max_steps", "3", "--warmup_steps", "1",
                                        "--log_every", "1", "--save_every", "2"])
    lm = T.main(args)
    ck = torch.load(str(tmp_path / "last.ckpt"), map_location="cpu", weights_only=False)
    assert ck["global_step"] == 3 and ck["hyper_parameters"]["arch"] == "small"
    assert set(ck["state_dict"]) == set(lm.state_dict()) and len(ck["optimizer_states"][0]["state"]) > 100
    assert np.isfinite(float(lm.logged["loss"].detach())) and abs(ck["hyper_parameters"]["learning_rate"] - 5e-4 * 4 / 256) < 1e-12
    # resume: picks up at step 3 and runs to 4
    args2 = T.build_parser().parse_args(["--save_path", str(tmp_path), "--nproc", "1", "--arch", "small",
                                         "--batch_size_per_gpu", "4", "--num_workers", "0", "--synthetic_clips", "16",
                                         "--train_len", "1.0", "--max_steps", "4", "--warmup_steps", "1"])
    T.main(args2)
    ck2 = torch.load(str(tmp_path / "last.ckpt"), map_location="cpu", weights_only=False)
    assert ck2["global_step"] == 4
    assert int(ck2["optimizer_states"][0]["state"][next(iter(ck2["optimizer_states"][0]["state"]))]["step"]) == 4


def test_lmdb_to_device_batches(tmp_path):
    """f4 end to end: records written in the reference's LMDB / legacy-arrow layout -> LMDBDataset -> collate ->
    DevicePrefetcher -> device batches identical to the stored waveforms."""
    from audiossl_b200.datasets import DevicePrefetcher, LMDBDataset, collate_waveforms, write_dataset
    rng = np.random.RandomState(0)
    recs = [("c%03d" % i, (rng.standard_normal((1, 4000 + 100 * i)) * 0.1).astype(np.float32),
             np.eye(1, 4, i % 4, dtype=np.float32)) for i in range(10)]
    write_dataset(str(tmp_path / "train.lmdb"), recs)
    ds = LMDBDataset(str(tmp_path), "train")
    loader = torch.utils.data.DataLoader(ds, batch_size=4, collate_fn=lambda s: collate_waveforms(s, 4500))
    seen = 0
    by_key = {k: (w, l) for k, w, l in recs}
    for bi, (wav, labels) in enumerate(DevicePrefetcher(loader, "cuda")):
        assert wav.is_cuda and labels.is_cuda and wav.shape[1:] == (1, 4500)
        for j in range(wav.shape[0]):
            name = ds.keys[bi * 4 + j].decode()
            ref = np.zeros(4500, np.float32)
            m = min(4500, by_key[name][0].shape[1])
            ref[:m] = by_key[name][0][0, :m]
            assert np.array_equal(wav[j, 0].cpu().numpy(), ref)
            assert np.array_equal(labels[j].cpu().numpy(), by_key[name][1][0])
        seen += wav.shape[0]
    assert seen == 10


def test_validation_between_training_steps_keeps_the_runtime_and_accumulation_is_refused():
    """model.teacher.encoder used stand-alone between two training steps (Lightning validation, downstream probes)
    reads the training runtime's flat storage in place - no rebuild of the runtime, same numbers as a free-standing
    copy of the encoder - and a second backward before the optimizer consumed the first one is an error, not a silent
    overwrite (accumulate_grad_batches > 1)."""
    import copy
    from audiossl_b200.methods.atst.model import ATSTLightningModule
    torch.manual_seed(0)
    lm = ATSTLightningModule(arch=dict(embed_dim=128, depth=2, num_heads=2), learning_rate=1e-3, warmup_steps=1,
                             max_steps=10, drop_path_rate=0.0).cuda().train()
    opt = lm.configure_optimizers()[0]
    lm.trainer.optimizers = [opt]
    crops, lengths = util.make_inputs("tiny2b32", 8, [101, 101], [[101] * 8, [101 - i for i in range(8)]])
    batch = (([c.cuda() for c in crops], [l.cuda() for l in lengths]), None)

    def train_step(i):
        lm.global_step = i
        loss = lm.training_step(batch, i)
        opt.zero_grad()
        loss.backward()
        opt.step()
        lm.on_train_batch_end(None, None, i)
        return loss
    train_step(0)
    rt = lm.model._rt
    enc = lm.model.teacher.encoder
    emb = enc(crops[0].cuda(), length=lengths[0].cuda())
    assert lm.model._rt is rt and rt.current()                       # the parameters were not re-pointed
    assert enc._inf["fp"] is rt.ft
    free = copy.deepcopy(enc)                                         # a free-standing encoder: private flat buffer
    free._inf = None
    for p in free.parameters():
        p._atst_flat = None
    emb2 = free(crops[0].cuda(), length=lengths[0].cuda())
    assert torch.equal(emb, emb2)
    train_step(1)
    assert lm.model._rt is rt                                         # and the training runtime survived validation
    emb3 = enc(crops[0].cuda(), length=lengths[0].cuda())
    assert not torch.equal(emb, emb3)                                 # the EMA update is visible to the next validation
    # gradient accumulation is refused
    lm.global_step = 2
    l1 = lm.training_step(batch, 2)
    opt.zero_grad()
    l1.backward()
    l2 = lm.training_step(batch, 2)
    with pytest.raises(RuntimeError, match="accumulation"):
        l2.backward()


def test_graphed_training_step_matches_eager_steps():
    """audiossl_b200.graph.GraphedTrainStep: the whole step replayed as one CUDA graph gives the eager loop's losses
    and parameters over several steps (schedules, Adam bias corrections and the EMA momentum reach the kernels through
    device memory), and building it leaves the model where it was."""
    import copy
    from audiossl_b200.graph import GraphedTrainStep
    from audiossl_b200.methods.atst.model import ATSTLightningModule

    def make():
        torch.manual_seed(0)
        lm = ATSTLightningModule(arch=dict(embed_dim=128, depth=2, num_heads=2), learning_rate=1e-3, warmup_steps=2,
                                 max_steps=20, ema=0.9, drop_path_rate=0.0)
        util.load_det(lm.model)
        lm.cuda().train()
        opt = lm.configure_optimizers()[0]
        lm.trainer.optimizers = [opt]
        return lm, opt
    batches = []
    for s in range(4):
        crops, lengths = util.make_inputs("graph%d" % s, 8, [101, 101], [[101 - (i * 3) % 30 for i in range(8)], [101] * 8])
        batches.append((([c.cuda() for c in crops], [l.cuda() for l in lengths]), None))
    # eager reference
    lm_e, opt_e = make()
    losses_e = []
    for s, batch in enumerate(batches):
        lm_e.global_step = s
        loss = lm_e.training_step(batch, s)
        opt_e.zero_grad()
        loss.backward()
        opt_e.step()
        lm_e.on_train_batch_end(None, None, s)
        losses_e.append(loss.item())
    # graphed
    lm_g, opt_g = make()
    before = {k: v.detach().clone() for k, v in lm_g.state_dict().items()}
    step = GraphedTrainStep(lm_g, opt_g, batches[0])
    for k, v in lm_g.state_dict().items():
        assert torch.equal(v, before[k]), "building the graph changed " + k
    assert opt_g._step == 0
    losses_g = [step(batch, s).item() for s, batch in enumerate(batches)]
    # same kernels, same inputs: the first steps agree to fp32 noise; after that the atomically accumulated weight
    # gradients (order differs from launch to launch) are amplified by Adam's normalisation of tiny gradients
    np.testing.assert_allclose(losses_g[:2], losses_e[:2], rtol=2e-5)
    np.testing.assert_allclose(losses_g, losses_e, rtol=1e-3)
    assert opt_g._step == 4
    sd_e, sd_g = lm_e.state_dict(), lm_g.state_dict()
    lr_sum = float(sum(lm_e.mylr_scheduler[:4]))
    worst = 0.0
    for k in sd_e:
        if not sd_e[k].dtype.is_floating_point:
            assert torch.equal(sd_e[k], sd_g[k]), k
            continue
        moved = (sd_e[k] - before[k]).norm().item()
        diff = (sd_e[k] - sd_g[k]).norm().item()
        if moved > 0:
            # tensors whose true gradient is numerically zero (final-norm bias: BatchNorm removes constants) take
            # pure-noise Adam steps in both runs: measure against a fifth of a full-rate update at least
            floor = 0.2 * lr_sum * sd_e[k].numel() ** 0.5 if k.startswith("model.student") else 0.0
            worst = max(worst, diff / max(moved, floor))
        else:
            assert diff == 0.0, k
    assert worst < 0.1, worst  # (same amplification; a wrong schedule / bias correction would be off by O(1))
    # back to eager on the same objects
    step.release()
    lm_g.global_step = 4
    loss = lm_g.training_step(batches[0], 4)
    opt_g.zero_grad()
    loss.backward()
    opt_g.step()
    assert np.isfinite(loss.item()) and opt_g._step == 5


def test_frame_launcher_on_synthetic_clips(tmp_path):
    """the atstframe/train.py-equivalent launcher: synthetic clips -> pinned H2D -> device transform with block masks
    -> FrameATSTLightningModule steps -> checkpoint."""
    from audiossl_b200.methods.atstframe import train as T
    args = T.build_parser().parse_args(["--save_path", str(tmp_path), "--nproc", "1", "--arch", "small",
                                        "--batch_size_per_gpu", "4", "--num_workers", "0", "--synthetic_clips", "8",
                                        "--anchor_len", "2.0", "--mask_type", "block", "--mask_ratio", "0.65",
                                        "--max_steps", "2", "--warmup_steps", "1", "--log_every", "1",
                                        "--save_every", "2"])
    lm = T.main(args)
    ck = torch.load(str(tmp_path / "last.ckpt"), map_location="cpu", weights_only=False)
    assert ck["global_step"] == 2 and np.isfinite(float(lm.logged["loss"].detach()))
    assert "model.student.encoder.norm_frame.weight" in ck["state_dict"]
    # the checkpoint is what load_model of the embedding API reads
    from audiossl_b200.methods.atstframe import embedding as E
    enc = E.load_model(str(tmp_path / "last.ckpt"))
    emb = E.get_scene_embedding(torch.randn(2, 1, 32000, device="cuda") * 0.1, enc)
    assert tuple(emb.shape) == (2, 12 * 384) and torch.isfinite(emb).all()
