"""CPU tests: host-side logic, C-ABI symbol table, distributed plumbing on gloo (world_size 2)."""
import os
import re
import subprocess
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    from audiossl_b200 import _lib
    from audiossl_b200.build import build
    build()
    lib = _lib.load()
    hdr = open(os.path.join(ROOT, "include", "atst_b200.h")).read()
    names = set(re.findall(r"\b(atst_\w+)\s*\(", hdr))
    assert len(names) >= 25
    for n in names:
        assert hasattr(lib, n), n
    assert names - {"atst_last_error"} == set(_lib.SIGNATURES)
    assert lib.atst_version() >= 100 and lib.atst_is_precise() == 0
    # the 3xTF32 validation build exports the same boundary; bring-up entry points are in neither
    precise = _lib.load("precise")
    for n in names:
        assert hasattr(precise, n), n
    assert precise.atst_is_precise() == 1
    dbg = open(os.path.join(ROOT, "include", "atst_b200_debug.h")).read()
    for n in set(re.findall(r"\b(atst_\w+)\s*\(", dbg)):
        assert not hasattr(lib, n) and n not in names, n


def test_ctypes_signatures_match_the_header():
    """every declaration of include/atst_b200.h against the ctypes argtypes of _lib.SIGNATURES: same parameter count
    and the same pointer / integer / float kind at every position (an ABI drift would otherwise only show on the GPU)."""
    import ctypes
    from audiossl_b200 import _lib
    hdr = open(os.path.join(ROOT, "include", "atst_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", " ", hdr, flags=re.S)
    decls = dict(re.findall(r"\bint\s+(atst_\w+)\s*\(([^;]*?)\)\s*;", hdr, flags=re.S))
    assert set(decls) == set(_lib.SIGNATURES)

    def kind(param):
        param = param.strip()
        if "*" in param:
            return "ptr"
        if param.startswith("float"):
            return "float"
        return "int"  # int, unsigned, long long

    ck = {ctypes.c_void_p: "ptr", ctypes.c_char_p: "ptr", ctypes.c_float: "float", ctypes.c_int: "int",
          ctypes.c_longlong: "int", ctypes.c_uint: "int"}
    for name, params in decls.items():
        plist = [] if params.strip() in ("", "void") else [x for x in params.split(",")]
        want = [kind(x) for x in plist]
        have = [ck[a] for a in _lib.SIGNATURES[name]]
        assert want == have, (name, want, have)
        # 64-bit integers must be declared as such on both sides
        for x, a in zip(plist, _lib.SIGNATURES[name]):
            if "long long" in x and "*" not in x:
                assert a is ctypes.c_longlong, (name, x)
            if a is ctypes.c_longlong:
                assert "long long" in x, (name, x)


def test_epilogue_codes_agree_between_the_enum_the_header_and_python():
    """GemmEpilogue (csrc/gemm.h) is what the C ABI's `epi` argument means: ops.py mirrors it by value and
    include/atst_b200.h documents every code a caller may pass."""
    from audiossl_b200 import ops
    src = open(os.path.join(ROOT, "audiossl_b200", "csrc", "gemm.h")).read()
    enum = {k: int(v) for k, v in re.findall(r"\b(EPI_[A-Z_]+)\s*=\s*(\d+)", src)}
    names = ["EPI_STORE", "EPI_GELU", "EPI_DGELU", "EPI_RESID", "EPI_SCALE", "EPI_RELU", "EPI_GELU_H", "EPI_DGELU_H"]
    for name in names:
        assert enum[name] == getattr(ops, name), name
    assert len({enum[n] for n in names}) == len(names)
    hdr = open(os.path.join(ROOT, "include", "atst_b200.h")).read()
    for code in (enum["EPI_GELU_H"], enum["EPI_DGELU_H"]):
        assert re.search(r"\|\s*%d\s" % code, hdr), "epilogue %d is not documented in the public header" % code


def test_aux_dtype_is_checked_before_the_library_is_called():
    """the fp16 gelu' epilogues take a half aux tensor and the fp32 ones a float one: a mismatch would be read with the
    wrong stride on the device, so it is refused on the host (no GPU needed to see it)."""
    import torch
    from audiossl_b200 import ops
    A, W = torch.zeros(32, 16), torch.zeros(256, 16)
    with pytest.raises(TypeError):
        ops.gemm_nt(A, W, epi=ops.EPI_GELU_H, aux=torch.zeros(32, 256))
    with pytest.raises(TypeError):
        ops.gemm_nt(A, W, epi=ops.EPI_GELU, aux=torch.zeros(32, 256, dtype=torch.float16))
    with pytest.raises(TypeError):
        ops.gemm_nn(torch.zeros(32, 16), torch.zeros(16, 256), epi=ops.EPI_DGELU_H, aux=torch.zeros(32, 256))


def test_fp16_side_stream_is_only_chosen_for_widths_the_pair_kernel_serves():
    """engine.half_dgelu(width): the fp16 gelu' epilogues live in the CTA-pair GEMM (width % 256 == 0 or width > 1024,
    and a multiple of 32 for the dgrad); any other MLP width keeps the fp32 pre-activation path instead of failing."""
    from audiossl_b200 import engine
    assert engine.HALF_DGELU == engine.half_dgelu()  # default library: the switch alone
    if engine.HALF_DGELU:
        for width in (512, 768, 1536, 3072, 4096, 1056):
            assert engine.half_dgelu(width), width
        for width in (640, 1040, 1100, 96):
            assert not engine.half_dgelu(width), width
        import audiossl_b200
        with audiossl_b200.precision("3xtf32"):
            assert not engine.half_dgelu(3072)  # the validation build keeps fp32 everywhere


def test_built_library_sass_has_the_blackwell_paths_and_no_regressions():
    """cuobjdump of the in-tree library (no GPU needed): tcgen05 / TMEM / TMA opcodes in every hot kernel, no register
    spills in the per-class pair-GEMM instantiations, and no GPU-scope fence on the accumulator hand-back (the only
    MEMBAR.ALL.GPU left in a pair GEMM are the two cluster barriers at kernel start and end)."""
    import shutil
    from audiossl_b200 import _lib
    if shutil.which("cuobjdump") is None or not os.path.exists(_lib.LIB_PATH):
        pytest.skip("needs cuobjdump and the built library")
    sass = subprocess.run(["cuobjdump", "-sass", _lib.LIB_PATH], stdout=subprocess.PIPE, text=True, check=True).stdout
    kernels, cur = {}, None
    for line in sass.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = kernels.setdefault(m.group(1), {})
            continue
        m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
        if m and cur is not None:
            op = m.group(1)
            for w in ("UTCHMMA", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UBLKCP", "STL", "MEMBAR.ALL.GPU"):
                if op == w or op.startswith(w + "."):
                    cur[w] = cur.get(w, 0) + 1
    pair = {k: v for k, v in kernels.items() if "gemm2_tf32_kernel" in k}
    assert len(pair) >= 9, sorted(pair)  # 3 layouts x plain, residual, GELU, 2 fp16 classes, 2 generic
    for k, c in pair.items():
        assert c.get("UTCHMMA", 0) > 0 and c.get("LDTM", 0) > 0 and c.get("UTMALDG", 0) > 0, k
        assert c.get("MEMBAR.ALL.GPU", 0) <= 2, (k, c)
        if not k.endswith("Li0EEEv14CUtensorMap_stS1_NS_10GemmParamsE"):  # every class but the run-time generic one
            assert c.get("STL", 0) == 0, (k, c)
    for name in ("attn_fwd_tc_kernel", "attn_bwd_tc_kernel"):
        ks = [c for k, c in kernels.items() if name in k]
        assert ks and all(c.get("UTCHMMA", 0) > 0 and c.get("LDTM", 0) > 0 and c.get("STTM", 0) > 0 and
                          c.get("UTMALDG", 0) > 0 and c.get("UTMASTG", 0) > 0 for c in ks), name
    mel = [c for k, c in kernels.items() if "mel_kernel" in k]
    assert mel and mel[0].get("UBLKCP", 0) > 0


def test_no_cpu_fallback():
    from audiossl_b200.models.atst import ATST
    from audiossl_b200.transforms import LogMelSpectrogram
    with pytest.raises(RuntimeError):
        LogMelSpectrogram()(torch.zeros(1, 16000))
    m = ATST(arch=dict(embed_dim=128, depth=1, num_heads=2))
    with pytest.raises(RuntimeError):
        m([torch.zeros(2, 1, 64, 101)] * 2, [torch.full((2,), 101)] * 2)


def test_product_never_imports_oracle():
    for dp, _, files in os.walk(os.path.join(ROOT, "audiossl_b200")):
        for f in files:
            if f.endswith(".py"):
                src = open(os.path.join(dp, f)).read()
                assert "oracle" not in src.replace("# oracle", ""), os.path.join(dp, f)


def test_state_dict_keys_match_reference_layout():
    from audiossl_b200.methods.atst.model import ATSTLightningModule
    lm = ATSTLightningModule(arch="small", max_steps=10, warmup_steps=2)
    keys = list(lm.state_dict().keys())
    assert len(keys) == 299  # SURVEY.md section 5 [verified on the reference]
    assert "model.student.encoder.patch_embed.patch_embed.weight" in keys
    assert "model.teacher.projector.1.running_var" in keys
    assert "model.student.predictor.3.weight" in keys
    assert not any(k.startswith("model.teacher.predictor") for k in keys)


def test_flat_params_layout_and_groups():
    from audiossl_b200.models.atst import ATST
    from audiossl_b200.params import FlatParams
    from audiossl_b200.utils.common import get_params_groups
    m = ATST(arch=dict(embed_dim=128, depth=2, num_heads=2))
    fs = FlatParams(list(m.student.named_parameters()), torch.device("cpu"))
    ft = FlatParams(list(m.teacher.named_parameters()), torch.device("cpu"))
    assert ft.total == fs.ema_count and ft.order == fs.order[:len(ft.order)]
    reg, noreg = get_params_groups(m.student, debug=True)
    segs = fs.wd_segments()
    for name in fs.order:
        off = fs.offsets[name]
        seg = [s for s in segs if s[0] <= off < s[1]][0]
        assert seg[2] == (name in reg), name
        assert off % 64 == 0
    # parameters are views of the flat buffer: an in-place update of the buffer is visible in the module
    fs.data.add_(1.0)
    assert torch.equal(m.student.encoder.cls_token.data.view(-1), fs.p("encoder.cls_token").view(-1))
    assert fs.is_current()


def test_schedules_match_golden():
    from audiossl_b200.utils.common import cosine_scheduler_step
    from tests import util
    g = util.gold("sched.npz")
    np.testing.assert_array_equal(cosine_scheduler_step(0.99, 1, 1000, 0), g["ema"])
    np.testing.assert_array_equal(cosine_scheduler_step(5e-4, 1e-6, 1000, 100), g["lr"])


def test_crop_grouping():
    from audiossl_b200.models.atst.byol import MultiCropWrapper
    x = [torch.zeros(1, 1, 64, w) for w in (601, 601, 101, 101, 101, 601)]
    assert MultiCropWrapper.group_crops(x) == [(0, 2), (2, 5), (5, 6)]


def test_transforms_api():
    import audiossl_b200.transforms as T
    x = torch.arange(10.)[None]
    assert T.CentralCrop(4)(x).tolist() == [[3., 4., 5., 6.]]
    assert T.PadToSize(12)(x).shape == (1, 12)
    assert T.RandomCrop(20)(x).shape == (1, 20)  # short input is zero padded (common.py:69-72)
    assert T.ToSizeN(4)(x).shape == (1, 8)  # remainder 2 is not past the half-way point: truncates
    np.testing.assert_allclose(T.MinMax(0., 10.)(torch.tensor([0., 5., 10.])).numpy(), [-1., 0., 1.])
    lms = torch.randn(1, 64, 50)
    # the augmentations are fronts of the CUDA kernels (they follow the fused mel on the device): no CPU arithmetic
    with pytest.raises(RuntimeError, match="GPU only"):
        T.RandomResizeCrop()(lms)
    mx = T.Mixup()
    assert torch.equal(mx(lms), lms) and len(mx.memory_bank) == 1  # empty bank: identity, input remembered
    with pytest.raises(RuntimeError, match="GPU only"):
        mx(lms * 0.5)
    i, j, h, w = T.RandomResizeCrop.get_params((64, 75), (64, 50), (0.6, 1.5), (0.6, 1.5))
    assert 1 <= h <= 64 and 1 <= w <= 75 and 0 <= i <= 64 - h and 0 <= j <= 75 - w


WORKER = r'''
import os, sys, torch, torch.distributed as dist
sys.path.insert(0, %r)
from audiossl_b200 import distributed as D
rank = int(os.environ["RANK"]); dist.init_process_group("gloo")
torch.manual_seed(0)
full = torch.randn(12, 8) * 3 + 1
mine = full[rank * 6:(rank + 1) * 6]
mean = mine.mean(0); m2 = ((mine - mean) ** 2).sum(0)
gm, gm2, n = D.bn_stats_sync(mean, m2, 6.0)
assert n == 12.0
assert torch.allclose(gm, full.mean(0), atol=1e-6) and torch.allclose(gm2 / n, full.var(0, unbiased=False), atol=1e-5)
s1, s2 = D.bn_sums_sync(mine.sum(0), (mine ** 2).sum(0))
assert torch.allclose(s1, full.sum(0), atol=1e-5) and torch.allclose(s2, (full ** 2).sum(0), atol=1e-4)
g = torch.full((5,), float(rank + 1)); D.allreduce_avg_(g); assert torch.allclose(g, torch.full((5,), 1.5))
# two BatchNorm layers in one exchange; unequal per-rank row counts (ATST-Frame) with the global count passed in
parts = [full[:5], full[5:]]
mine = parts[rank]
mean = mine.mean(0); m2 = ((mine - mean) ** 2).sum(0)
(ga, gb), (gc, gd) = D.bn_stats_sync_many([(mean, m2, float(len(mine))), (mean * 2, m2 * 4, float(len(mine)))])
assert torch.allclose(ga, full.mean(0), atol=1e-6) and torch.allclose(gb / 12, full.var(0, unbiased=False), atol=1e-5)
assert torch.allclose(gc, 2 * full.mean(0), atol=1e-5) and torch.allclose(gd / 12, 4 * full.var(0, unbiased=False), atol=1e-4)
gm, gm2, n = D.bn_stats_sync(mean, m2, float(len(mine)), n_total=12.0)
assert n == 12.0 and torch.allclose(gm, full.mean(0), atol=1e-6)
# overlapped gradient exchange: ranges submitted out of order, the complement sent by finish(), frozen head untouched
flat = torch.arange(100.) * (rank + 1)
ex = D.GradExchange(flat, 10, torch.device("cpu"))
ex.submit(60, 80); ex.submit(20, 30); ex.submit(5, 12)
ex.finish()
want = torch.arange(100.) * 1.5
want[:10] = torch.arange(10.) * (rank + 1)
assert torch.allclose(flat, want), flat
flat2 = torch.ones(50) * (rank + 1); ex2 = D.GradExchange(flat2, 0, torch.device("cpu")); ex2.finish()
assert torch.allclose(flat2, torch.full((50,), 1.5))
print("rank", rank, "ok")
'''


def test_distributed_helpers_gloo_world2(tmp_path):
    script = tmp_path / "w.py"
    script.write_text(WORKER % ROOT)
    env = dict(os.environ, OMP_NUM_THREADS="1")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                        "--master-addr", "127.0.0.1", "--master-port", "29671", str(script)],
                       capture_output=True, text=True, timeout=300, env=env)
    assert r.returncode == 0, r.stdout + r.stderr
    assert r.stdout.count("ok") == 2


def test_frozen_tensors_are_outside_the_optimizer_and_exchange_ranges():
    """ATST-clip never reads encoder.mask_embed: the reference's AdamW skips it (grad is None), so the flat layout
    keeps it out of the fused-optimizer segments and out of the all-reduced gradient slice; the teacher keeps the
    same prefix layout for the EMA."""
    from audiossl_b200.models.atst import ATST
    from audiossl_b200.params import FlatParams
    m = ATST(arch=dict(embed_dim=128, depth=2, num_heads=2))
    m.student.projector[1].bias.requires_grad = False
    frozen = {"encoder.mask_embed", "projector.1.bias"}
    fs = FlatParams(list(m.student.named_parameters()), torch.device("cpu"), frozen=frozen)
    ft = FlatParams(list(m.teacher.named_parameters()), torch.device("cpu"), frozen=frozen)
    assert ft.total == fs.ema_count and ft.order == fs.order[:len(ft.order)]
    segs = fs.wd_segments()
    assert len(segs) == 4

    def covered(name):
        return any(a <= fs.offsets[name] < b for a, b, _ in segs)
    assert not covered("encoder.mask_embed") and not covered("projector.1.bias")
    assert all(covered(n) for n in fs.order if n not in frozen)
    assert fs.offsets["encoder.mask_embed"] == 0
    ex = fs.exchanged_grad()
    assert ex.numel() == fs.total - 128 and ex.data_ptr() == fs.grad.data_ptr() + 4 * 128
    fs.attach_grads()
    assert m.student.encoder.mask_embed.grad is None and m.student.encoder.cls_token.grad is not None
    assert m.student.projector[1].bias.grad is None


def test_optimizer_checkpoint_uses_the_transformers_adamw_layout():
    """state_dict()['state'][i] = {step, exp_avg, exp_avg_sq} per stepped parameter (what transformers' AdamW saves),
    and such a checkpoint loads back into the flat moment buffers."""
    from audiossl_b200.models.atst import ATST
    from audiossl_b200.optim import FusedHFAdamW
    from audiossl_b200.params import FlatParams
    from audiossl_b200.utils.common import get_params_groups
    m = ATST(arch=dict(embed_dim=128, depth=1, num_heads=2))
    fs = FlatParams(list(m.student.named_parameters()), torch.device("cpu"), frozen={"encoder.mask_embed"})
    opt = FusedHFAdamW(get_params_groups(m.student), flat=lambda: fs, lr=1e-3)
    params = [p for g in opt.param_groups for p in g["params"]]
    ref_state = {i: {"step": 7, "exp_avg": torch.full_like(p, 0.5 + i), "exp_avg_sq": torch.full_like(p, 2.0 + i)}
                 for i, p in enumerate(params) if p is not m.student.encoder.mask_embed}
    sd = {"state": ref_state, "param_groups": opt.state_dict()["param_groups"]}
    keep = dict(sd)
    opt.load_state_dict(sd)
    assert sd.keys() == keep.keys()  # the caller's dict is not modified
    assert opt._step == 7
    out = opt.state_dict()["state"]
    assert set(out) == set(ref_state)
    for i in ref_state:
        assert torch.equal(out[i]["exp_avg"], ref_state[i]["exp_avg"])
        assert torch.equal(out[i]["exp_avg_sq"], ref_state[i]["exp_avg_sq"])
    name = "encoder.blocks.0.mlp.fc1.weight"
    i = [k for k, p in enumerate(params) if p is m.student.encoder.blocks[0].mlp.fc1.weight][0]
    assert torch.equal(fs.view(opt._m, name), ref_state[i]["exp_avg"])
