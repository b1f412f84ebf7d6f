"""Generate golden vectors from the UNMODIFIED reference (run in the build container only).

    python tests/golden/make_golden.py            # needs /root/reference on disk

The reference has no tests of its own, so parity is pinned by running its code
in-process here and committing the outputs as small fixtures under tests/golden/.  Nothing is
copied from the reference: the script imports it, feeds deterministic inputs/weights
(tests/golden/detfill.py) and stores what comes out.  Harness: 1-rank gloo process
group + ``torch.Tensor.cuda`` identity shim because ``compute_var`` hard-codes ``.cuda()`` and
``all_reduce`` (audiossl/models/atst/byol.py:42-53).
"""
import os
import sys
from functools import partial

import numpy as np
import torch
from torch import nn

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
sys.path.insert(0, "/root/reference")

from tests.golden import detfill  # noqa: E402


def harness():
    import torch.distributed as dist
    if not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29591")
        dist.init_process_group("gloo", rank=0, world_size=1)
    torch.Tensor.cuda = lambda self, *a, **k: self


def load_det(module):
    sd = module.state_dict()
    new = detfill.fill_state_dict(sd)
    module.load_state_dict({k: (torch.from_numpy(new[k]) if k in new else v) for k, v in sd.items()})


def flat(prefix, d, out):
    for k, v in d.items():
        out[prefix + "/" + k] = v


def grads_summary(module, out, prefix):
    for name, p in module.named_parameters():
        if p.grad is None:
            continue
        flat(prefix + "/grad/" + name, detfill.summarize(p.grad.numpy()), out)


# ----------------------------------------------------------------------------- mel
def gen_mel():
    from audiossl.methods.atst.transform import ATSTTrainTransform
    import torchaudio
    from audiossl.transforms.common import MinMax
    from torchvision import transforms
    out = {}
    tf1024 = ATSTTrainTransform().mel_feature
    mel640 = torchaudio.transforms.MelSpectrogram(16000, f_min=60, f_max=7800, hop_length=160,
                                                  win_length=640, n_fft=1024, n_mels=64)
    tf640 = transforms.Compose([mel640, torchaudio.transforms.AmplitudeToDB(stype="power", top_db=80),
                                MinMax(min=-79.6482, max=50.6842)])
    cases = [("noise", 16000), ("sine_silence", 16000), ("chirp", 16000), ("zeros", 16000),
             ("impulses", 16000), ("noise", 8000), ("noise", 1600), ("noise", 40000), ("chirp", 96000),
             ("noise", 160000), ("sine_silence", 160000)]
    for kind, n in cases:
        wav = torch.from_numpy(detfill.signal(kind, n))[None]
        for win, tf in ((1024, tf1024), (640, tf640)):
            if win == 640 and n > 40000:
                continue
            y = tf(wav)  # [1,64,T]
            out["%s_%d_w%d" % (kind, n, win)] = y.numpy()
    # batched 4-D input: per-clip top_db reference (embedding.py:57-60 relies on [B,1,n] -> [B,1,64,T])
    wavs = torch.stack([torch.from_numpy(detfill.signal(k, 16000)) for k in ("noise", "sine_silence", "chirp")])
    out["batch3_16000_w1024"] = tf1024(wavs[:, None]).numpy()
    np.savez_compressed(os.path.join(HERE, "mel.npz"), **out)
    print("mel.npz", len(out))


# ----------------------------------------------------------------------------- encoder / loss
def make_inputs(tag, B, widths, lens):
    crops, lengths = [], []
    for i, (w, l) in enumerate(zip(widths, lens)):
        crops.append(torch.from_numpy(detfill.det_array("%s/crop%d" % (tag, i), (B, 1, 64, w), 1.0, "uniform")))
        lengths.append(torch.tensor(l, dtype=torch.int64))
    return crops, lengths


class RefATSTLike(nn.Module):
    """ATST.forward (models/atst/atst.py:24-28) over arbitrary encoder sizes, reference parts only."""

    def __init__(self, embed_dim, depth, heads, ncrops, drop_path_rate=0.0):
        super().__init__()
        from audiossl.models.atst.byol import MultiCropWrapper, ByolLoss
        from audiossl.models.atst.audio_transformer import AST
        mk = lambda: AST(patch_h=64, patch_w=4, embed_dim=embed_dim, depth=depth, num_heads=heads,
                         qkv_bias=False, norm_layer=partial(nn.LayerNorm, eps=1e-6),
                         drop_path_rate=drop_path_rate)
        self.student = MultiCropWrapper(mk(), embed_dim, predictor=True)
        self.teacher = MultiCropWrapper(mk(), embed_dim, predictor=False)
        for p in self.teacher.parameters():
            p.requires_grad = False
        self.loss_fn = ByolLoss(ncrops)

    def forward(self, melspecs, lengths):
        t = self.teacher(melspecs[:2], lengths[:2])
        s = self.student(melspecs, lengths)
        return self.loss_fn(s, t), s, t


def run_case(model, tag, B, widths, lens, out, record_rand=False):
    load_det(model)
    model.train()
    crops, lengths = make_inputs(tag, B, widths, lens)
    rec = []
    if record_rand:
        orig = torch.rand

        def rand(*a, **k):
            r = orig(*a, **k)
            rec.append(r.reshape(-1).clone())
            return r
        torch.rand = rand
    try:
        (loss, std_s, std_t), s, t = model(crops, lengths)
    finally:
        if record_rand:
            torch.rand = orig
    loss.backward()
    out[tag + "/loss"] = np.float32(loss.item())
    out[tag + "/std_s"] = np.float32(std_s.item())
    out[tag + "/std_t"] = np.float32(std_t.item())
    out[tag + "/student_out"] = s.detach().numpy()
    out[tag + "/teacher_out"] = t.detach().numpy()
    grads_summary(model.student, out, tag)
    for name, b in model.named_buffers():
        if "running" in name:
            flat(tag + "/buf/" + name, detfill.summarize(b.numpy()), out)
    if record_rand:
        out[tag + "/rand"] = torch.stack(rec).numpy()  # [n_calls, rows]
    return model


def gen_atst():
    harness()
    out = {}
    # tiny: D=128, depth 2, 2 heads (dh=64 as in every real config); full + ragged lengths
    m = RefATSTLike(128, 2, 2, ncrops=2)
    run_case(m, "tiny2", 3, [101, 101], [[101, 77, 50], [101, 101, 9]], out)
    # same encoder, 32 clips per crop: BatchNorm / loss statistics over 64 rows (well conditioned: used for the
    # gradient tolerance of the TF32 CUDA path)
    mb = RefATSTLike(128, 2, 2, ncrops=2)
    run_case(mb, "tiny2b32", 32, [101, 101], [[101 - (i * 7) % 60 for i in range(32)], [101 - (i * 11) % 45 for i in range(32)]], out)
    # EMA update pinned on the same module (models/atst/atst.py:29-34 semantics via ATST.update_teacher)
    from audiossl.models.atst.atst import ATST
    with torch.no_grad():
        ATST.update_teacher(m, 0.99)
    for name, p in m.teacher.named_parameters():
        flat("tiny2/ema/" + name, detfill.summarize(p.detach().numpy()), out)
    # multi-crop 2 global + 2 local (two encoder calls per network), ragged local lengths
    m = RefATSTLike(128, 2, 2, ncrops=4)
    run_case(m, "tiny4", 2, [101, 101, 41, 41], [[101, 90], [101, 101], [41, 33], [41, 41]], out)
    # drop path active in both networks (train-mode teacher, D7): record the torch.rand stream
    m = RefATSTLike(128, 2, 2, ncrops=2, drop_path_rate=0.5)
    torch.manual_seed(7)
    run_case(m, "tiny2dp", 4, [101, 101], [[101, 101, 60, 101], [101, 80, 101, 101]], out, record_rand=True)
    # the real thing: ATST("small") as the Lightning module builds it, drop path off for parity
    m = ATST(arch="small", drop_path_rate=0.0)

    class Wrap(nn.Module):
        def __init__(s, a):
            super().__init__()
            s.student, s.teacher, s.loss_fn = a.student, a.teacher, a.loss_fn

        def forward(s, c, l):
            t = s.teacher(c[:2], l[:2])
            st = s.student(c, l)
            return s.loss_fn(st, t), st, t
    run_case(Wrap(m), "small2", 2, [101, 101], [[101, 64], [101, 101]], out)
    np.savez_compressed(os.path.join(HERE, "atst.npz"), **out)
    print("atst.npz", len(out))


def frame_stubs():
    """sys.modules stand-ins for packages the frame model imports but this image lacks."""
    import types
    if "pytorch_lightning" not in sys.modules:
        pl = types.ModuleType("pytorch_lightning")

        class LightningModule(nn.Module):
            pass
        pl.LightningModule = LightningModule
        sys.modules["pytorch_lightning"] = pl
    if "fairseq" not in sys.modules:
        fs, fd, fdu = (types.ModuleType(n) for n in ("fairseq", "fairseq.data", "fairseq.data.data_utils"))
        fdu.compute_mask_indices = lambda *a, **k: None  # masks are inputs of the fixtures
        sys.modules.update({"fairseq": fs, "fairseq.data": fd, "fairseq.data.data_utils": fdu})
    import transformers.optimization as to
    if not hasattr(to, "AdamW"):
        to.AdamW = torch.optim.AdamW  # never stepped here


def frame_masks(tag, B, P):
    m = detfill.det_array(tag + "/mask", (B, P), 1.0, "uniform") > 0.0
    m[:, 0] = True  # every clip has at least one masked frame inside any length
    return torch.from_numpy(m)


def gen_frame():
    """ATST-Frame (methods/atstframe): FrameAST + symmetric frame-level BYOL loss on masked frames."""
    harness()
    frame_stubs()
    from audiossl.methods.atstframe.audio_transformer import FrameAST
    from audiossl.methods.atstframe.byol import ByolLoss, MultiCropWrapper
    out = {}

    class RefFrame(nn.Module):
        def __init__(self, dim, depth, heads):
            super().__init__()
            mk = lambda: FrameAST(patch_h=64, patch_w=4, embed_dim=dim, depth=depth, num_heads=heads, qkv_bias=False,
                                  norm_layer=partial(nn.LayerNorm, eps=1e-6), drop_path_rate=0.0)
            self.student = MultiCropWrapper(mk(), dim, predictor=True)
            self.teacher = MultiCropWrapper(mk(), dim, predictor=False)
            for p in self.teacher.parameters():
                p.requires_grad = False
            self.loss_fn = ByolLoss(symmetric=True)

        def forward(self, x, length, mask):  # FrameATST.forward, symmetric branch (model.py:68-72)
            tea = self.teacher(x, length, mask, False)
            stu = self.student(x, length, mask, True)
            return self.loss_fn(stu, tea), stu, tea

    for tag, B, lens in (("frame2", 4, [[101, 101, 77, 60], [101, 101, 77, 60]]),
                         ("frame2b16", 16, [[101 - (i * 5) % 40 for i in range(16)]] * 2)):
        m = RefFrame(128, 2, 2)
        load_det(m)
        m.train()
        crops, lengths = make_inputs(tag, B, [101, 101], lens)
        mask = frame_masks(tag, B, 25)
        (loss, std_s, std_t), s, t = m(crops, lengths, [mask, mask])
        loss.backward()
        out[tag + "/loss"] = np.float32(loss.item())
        out[tag + "/std_s"] = np.float32(std_s.item())
        out[tag + "/std_t"] = np.float32(std_t.item())
        out[tag + "/student_out"] = s.detach().numpy()
        out[tag + "/teacher_out"] = t.detach().numpy()
        grads_summary(m.student, out, tag)
    np.savez_compressed(os.path.join(HERE, "frame.npz"), **out)
    print("frame.npz", len(out))


def gen_infer():
    """inference entry points: get_intermediate_layers[_chunks] in eval mode."""
    harness()
    frame_stubs()
    from audiossl.models.atst.audio_transformer import AST
    from audiossl.methods.atstframe.audio_transformer import FrameAST
    out = {}
    kw = dict(patch_h=64, patch_w=4, embed_dim=128, depth=3, num_heads=2, qkv_bias=False,
              norm_layer=partial(nn.LayerNorm, eps=1e-6))
    enc = AST(**kw)
    load_det(enc)
    enc.eval()
    x = torch.from_numpy(detfill.det_array("infer/x", (3, 1, 64, 250), 1.0, "uniform"))
    length = torch.tensor([250, 180, 40])
    with torch.no_grad():
        out["clip/cls"] = enc(x[..., :101], length=torch.tensor([101, 101, 40])).numpy()
        layers = enc.get_intermediate_layers(x[..., :101], torch.tensor([101, 101, 40]), n=2)
        out["clip/layers"] = torch.stack(layers).numpy()
        out["clip/chunks"] = enc.get_intermediate_layers_chunks(x, length, n=2, chunk_len=101, avgpool=True).numpy()
    fenc = FrameAST(**kw)
    load_det(fenc)
    fenc.eval()
    with torch.no_grad():
        out["frame/scene"] = fenc.get_intermediate_layers(x[..., :101], torch.tensor([101, 101, 40]), n=2, scene=True).numpy()
        out["frame/seq"] = fenc.get_intermediate_layers(x[..., :101], torch.tensor([101, 101, 40]), n=2, scene=False).numpy()
    np.savez_compressed(os.path.join(HERE, "infer.npz"), **out)
    print("infer.npz", len(out))


AUG_RECTS = [(0, 0, 64, 151), (0, 25, 64, 101), (0, 10, 38, 60), (13, 40, 51, 111), (0, 150, 64, 1), (63, 0, 1, 151)]


def gen_augment():
    """Mixup arithmetic and RandomResizeCrop with injected draws."""
    from audiossl.transforms.byol_a import RandomResizeCrop, log_mixup_exp
    out = {}
    lms = torch.from_numpy(detfill.det_array("aug/lms", (len(AUG_RECTS), 1, 64, 101), 1.0, "uniform"))
    rrc = RandomResizeCrop((1, 1.5))
    res = []
    for b, rect in enumerate(AUG_RECTS):
        RandomResizeCrop.get_params = staticmethod(lambda *a, rect=rect: rect)
        res.append(rrc(lms[b]))
    out["rrc"] = torch.stack(res).numpy()
    z = torch.from_numpy(detfill.det_array("aug/bank", (3, 1, 64, 101), 1.0, "uniform"))
    alphas = [0.0, 0.13, 0.4]
    out["mixup"] = torch.stack([log_mixup_exp(lms[b].clone(), z[b].clone(), 1.0 - alphas[b]) for b in range(3)]).numpy()
    np.savez_compressed(os.path.join(HERE, "augment.npz"), **out)
    print("augment.npz", len(out))


def gen_sched():
    from audiossl.utils.common import cosine_scheduler_step, get_params_groups
    out = {"ema": cosine_scheduler_step(0.99, 1, 1000, 0), "wd": cosine_scheduler_step(0.04, 0.4, 1000, 0),
           "lr": cosine_scheduler_step(5e-4, 1e-6, 1000, 100)}
    m = RefATSTLike(128, 2, 2, ncrops=2)
    reg, noreg = get_params_groups(m.student, debug=True)
    out["reg"] = np.array(reg)
    out["noreg"] = np.array(noreg)
    np.savez_compressed(os.path.join(HERE, "sched.npz"), **out)
    print("sched.npz")


def gen_transform():
    """the train transforms end to end (crop -> mel_feature -> Mixup -> RandomResizeCrop -> pad) with seeded host
    generators: audiossl/methods/atst/transform.py:50-74, audiossl/methods/atstframe/transform.py:70-101."""
    import random
    frame_stubs()
    sys.path.insert(0, "/root/reference/audiossl/methods/atstframe")  # its transform does `import random_mask`
    from audiossl.methods.atst.transform import ATSTTrainTransform
    from audiossl.methods.atstframe.transform import FrameATSTTrainTransform
    out = {}
    # variable-length views: every log_mixup_exp branch (equal, shorter, longer bank entry) is reached
    random.seed(5)
    np.random.seed(5)
    tf = ATSTTrainTransform(anchor_len=(0.7, 1.0), positive_len=(0.7, 1.0))
    for k in range(5):
        wav = torch.from_numpy(detfill.det_array("tf/wav%d" % k, (1, 24000), 0.1))
        crops, lengths = tf(wav)
        out["clipvar/%d/crop0" % k], out["clipvar/%d/crop1" % k] = crops[0].numpy(), crops[1].numpy()
        out["clipvar/%d/lengths" % k] = np.array(lengths, np.int64)
    # the recipe's defaults (6 s views of a 10 s clip, one clip shorter than the view: zero-padded waveform)
    random.seed(6)
    np.random.seed(6)
    tf = ATSTTrainTransform()
    for k, n in enumerate((160000, 160000, 80000)):
        wav = torch.from_numpy(detfill.det_array("tf6/wav%d" % k, (1, n), 0.1))
        crops, lengths = tf(wav)
        for v in range(2):
            flat("clip6/%d/crop%d" % (k, v), detfill.summarize(crops[v].numpy()), out)
            out["clip6/%d/shape%d" % (k, v)] = np.array(crops[v].shape, np.int64)
        out["clip6/%d/lengths" % k] = np.array(lengths, np.int64)
    # ATST-Frame: one crop, two independently augmented views, frequency-only warp, random (non-block) mask
    random.seed(7)
    np.random.seed(7)
    torch.manual_seed(7)
    tf = FrameATSTTrainTransform(anchor_len=1.0, mask_type="random", mask_ratio=0.75)
    for k in range(3):
        wav = torch.from_numpy(detfill.det_array("tff/wav%d" % k, (1, 20000), 0.1))
        crops, lengths, masks = tf(wav)
        out["frame/%d/crop0" % k], out["frame/%d/crop1" % k] = crops[0].numpy(), crops[1].numpy()
        out["frame/%d/lengths" % k] = np.array(lengths, np.int64)
        out["frame/%d/mask" % k] = masks[0].numpy()
        assert masks[0] is masks[1]
    np.savez_compressed(os.path.join(HERE, "transform.npz"), **out)
    print("transform.npz", len(out))


def gen_embed():
    """audiossl/methods/atstframe/embedding.py:19-127: load_model from a Lightning-format checkpoint, scene and
    timestamp embeddings of clips longer than one 1001-frame chunk.  The checkpoint itself is not committed: the
    test rebuilds it from the same name-derived weights."""
    import tempfile
    harness()
    frame_stubs()
    import pytorch_lightning as pl

    def load_from_checkpoint(cls, path, **kw):  # what Lightning does for this module: ctor(**hyper_parameters) + weights
        ck = torch.load(path, map_location="cpu", weights_only=False)
        m = cls(**ck["hyper_parameters"])
        m.load_state_dict(ck["state_dict"])
        return m
    pl.LightningModule.load_from_checkpoint = classmethod(load_from_checkpoint)
    pl.LightningModule.save_hyperparameters = lambda self, *a, **k: None
    sys.path.insert(0, "/root/reference/audiossl/methods/atstframe")
    from audiossl.methods.atstframe import embedding as E
    from audiossl.methods.atstframe.model import FrameATSTLightningModule
    hp = dict(arch="small", learning_rate=5e-4, warmup_steps=10, max_steps=100, ema=0.99)
    lm = FrameATSTLightningModule(**hp)
    load_det(lm)
    out = {}
    with tempfile.TemporaryDirectory() as d:
        path = os.path.join(d, "last.ckpt")
        torch.save({"state_dict": lm.state_dict(), "hyper_parameters": hp}, path)
        enc = E.load_model(path)
    assert enc.scene_embedding_size == 384 * 2 * 12 and enc.timestamp_embedding_size == 384 * 12
    n = 16000 * 12 + 800  # 1206 frames: one full 1001-frame chunk + a 205-frame tail
    audio = torch.stack([torch.from_numpy(detfill.signal("noise", n)), torch.from_numpy(detfill.signal("chirp", n))])[:, None]
    with torch.no_grad():
        scene = E.get_scene_embedding(audio, enc)
        ts, stamps = E.get_timestamp_embedding(audio, enc)
        scene1 = E.get_scene_embedding(audio[0, :, :16000 * 3], enc)  # [1, n] input, shorter than one chunk
    out["scene"] = scene.numpy()
    out["scene1"] = scene1.numpy()
    out["ts/shape"] = np.array(ts.shape, np.int64)
    flat("ts", detfill.summarize(ts.numpy()), out)
    out["ts/rows"] = ts[:, ::37, :].numpy()[:, :, ::13]
    out["stamps"] = stamps.numpy()
    np.savez_compressed(os.path.join(HERE, "embed.npz"), **out)
    print("embed.npz", len(out), scene.shape, ts.shape)


if __name__ == "__main__":
    torch.manual_seed(0)
    gen_mel()
    gen_atst()
    gen_frame()
    gen_infer()
    gen_augment()
    gen_sched()
    gen_transform()
    gen_embed()
