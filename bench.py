"""ATST pre-training step throughput (BASELINE.json metric) on N B200s of one node, or the CPU reference arm.

    python bench.py --gpus 1 --steps 5 --warmup 3                 # our arm, one JSON line on stdout
    torchrun --nproc-per-node N ... bench.py --gpus N ...          # one rank per GPU (NCCL), weak scaling
    python bench.py --impl reference --steps 3 --warmup 1          # the reference's algorithm on the host cores

One "step" = one pass of the hot path over one synthetic batch: fused log-mel of both views (raw 16 kHz
waveforms already in HBM) -> EMA-teacher forward + student forward (AST-base, 251 tokens) -> BYOL loss ->
student backward -> gradient all-reduce (N > 1) -> HF-AdamW -> teacher EMA.  Workload at N = 1 is BASELINE
config 2: ATST-base, 10 s clips, 64 mels, 256 clips per GPU, DropPath 0.1 as in the recipe.
`e2e` repeats the measurement through the public Lightning-style API with HOST (pinned) waveforms: the
host->device copy of both views and the device->host read of the loss are inside the timed region.
"""
import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

CLIP_SECONDS = 10.0
SR = 16000
STEP_GFLOP_PER_CLIP = 360.5  # BASELINE.md section 4, config 2 (teacher fwd + student fwd + student bwd)

# BASELINE.json configs (per-GPU shapes).  c2 is the benchmark line; the others are run for the record
# (profiles/) with `--config`.  crops: list of (seconds, count); gflop: algorithmic step GFLOP per clip
CONFIGS = {
    "c1": dict(kind="clip", arch="small", crops=[(1.0, 2)], batch=8, gflop=9.0,
               name="ATST-small, 1 s clips, batch 8"),
    "c2": dict(kind="clip", arch="base", crops=[(10.0, 2)], batch=256, gflop=360.5,
               name="ATST-base, 10 s clips, 256 clips/GPU"),
    "c3": dict(kind="clip", arch="base", crops=[(6.0, 2), (1.0, 6)], batch=128, gflop=292.5,
               name="ATST-base, 2x6 s + 6x1 s crops, 128 clips/GPU"),
    "c4": dict(kind="frame", arch="base", crops=[(10.0, 2)], batch=64, gflop=370.0,
               name="ATST-Frame base, 10 s clips, 64 clips/GPU"),
    "c5": dict(kind="clip", arch="large", crops=[(6.0, 2)], batch=256, gflop=748.0,
               name="ATST-large, 6 s clips, 256 clips/GPU"),
}


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        d = json.load(open(path))
        return {"hbm_gbs": d["hbm_gbs"], "tflops": d.get("bf16_tflops_sustained", d["bf16_tflops"]), "src": "measured"}
    return {"hbm_gbs": 6650.0, "tflops": 1400.0, "src": "fallback"}


class ClockSampler(threading.Thread):
    """nvidia-smi style clock / throttle-reason sampling during the timed region (pynvml, 200 ms)."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.reasons, self.stop_flag = index, [], set(), False
        self.max_mhz = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:  # noqa: BLE001
            self.nv = None

    def run(self):
        if self.nv is None:
            return
        nv = self.nv
        names = {nv.nvmlClocksThrottleReasonHwSlowdown: "hw_slowdown",
                 nv.nvmlClocksThrottleReasonHwThermalSlowdown: "hw_thermal_slowdown",
                 nv.nvmlClocksThrottleReasonSwThermalSlowdown: "sw_thermal_slowdown",
                 nv.nvmlClocksThrottleReasonSwPowerCap: "sw_power_cap"}
        while not self.stop_flag:
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, name in names.items():
                    if r & bit:
                        self.reasons.add(name)
            except Exception:  # noqa: BLE001
                pass
            time.sleep(0.2)

    def summary(self):
        s = sorted(self.samples)
        return {"sm_mhz": s[len(s) // 2] if s else None, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons)}


# ------------------------------------------------------------------------------------------------ CPU arm
def cpu_reference_step_fn(batch, threads=None):
    """The reference's algorithm on the host: oracle mel_feature x2 views + teacher fwd + student fwd + loss +
    backward + EMA on ATST-base / 10 s clips (oracle/atst_oracle.py, pinned against the reference's outputs)."""
    import numpy as np
    import torch
    from oracle import atst_oracle as O
    if threads:
        torch.set_num_threads(threads)
    torch.manual_seed(0)
    model = O.OracleATST("base")
    model.train()
    g = torch.Generator().manual_seed(1234)
    wav = (torch.randn(2, batch, 1, int(CLIP_SECONDS * SR), generator=g) * 0.1).numpy()
    depth = 12
    rates = torch.linspace(0, 0.1, depth).tolist()

    def dp(S):
        out = []
        for r in rates:
            if r == 0:
                out.append(None)
                continue
            keep = 1 - r
            out.append((torch.floor(keep + torch.rand(S)) / keep, torch.floor(keep + torch.rand(S)) / keep))
        return [out]

    def step():
        crops = [torch.from_numpy(O.mel_feature(wav[v])) for v in range(2)]
        lengths = [torch.full((batch,), crops[0].shape[-1], dtype=torch.int64)] * 2
        for p in model.student.parameters():
            p.grad = None
        loss, _, _ = model(crops, lengths, dp_student=dp(2 * batch), dp_teacher=dp(2 * batch))
        loss.backward()
        model.update_teacher(0.9995)
        return float(loss.detach())

    return step


def reference_modules_step_fn(batch, threads=None, root="/root/reference"):
    """The UNMODIFIED reference modules on the host (only where /root/reference exists, i.e. in the build container;
    the GPU box has no copy and uses the oracle port): ATSTTrainTransform.mel_feature x2 views + ATST.forward +
    backward + update_teacher, through the harness of SURVEY.md section 8c (1-rank gloo group, Tensor.cuda identity
    because compute_var hard-codes .cuda() + all_reduce).  Returns None when the reference cannot be imported."""
    if not os.path.isdir(os.path.join(root, "audiossl")):
        return None
    try:
        import torch
        import torch.distributed as dist
        sys.path.insert(0, root)
        for k in [k for k in sys.modules if k == "audiossl" or k.startswith("audiossl.")]:
            del sys.modules[k]  # the repo's own `audiossl` alias package must not shadow the reference
        from audiossl.methods.atst.transform import ATSTTrainTransform
        from audiossl.models.atst.atst import ATST
        if not dist.is_initialized():
            os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
            os.environ.setdefault("MASTER_PORT", "29593")
            dist.init_process_group("gloo", rank=0, world_size=1)
        torch.Tensor.cuda = lambda self, *a, **k: self
    except Exception:  # noqa: BLE001
        return None
    finally:
        if root in sys.path:
            sys.path.remove(root)
    if threads:
        torch.set_num_threads(threads)
    torch.manual_seed(0)
    model = ATST(arch="base")  # drop_path_rate 0.1 (the AST default), train mode in both networks
    model.train()
    mel = ATSTTrainTransform().mel_feature
    g = torch.Generator().manual_seed(1234)
    wav = torch.randn(2, batch, 1, int(CLIP_SECONDS * SR), generator=g) * 0.1

    def step():
        crops = [mel(wav[v]) for v in range(2)]
        lengths = [torch.full((batch,), crops[0].shape[-1], dtype=torch.int64)] * 2
        for p in model.student.parameters():
            p.grad = None
        loss, _, _ = model(crops, lengths)
        loss.backward()
        model.update_teacher(0.9995)
        return float(loss.detach())

    return step


def cpu_step_fn(batch, threads):
    """(step function, kind): the reference's own modules when they are importable here, else the oracle port."""
    fn = reference_modules_step_fn(batch, threads)
    if fn is not None:
        return fn, "reference"
    return cpu_reference_step_fn(batch, threads), "port"


def one_thread_number(seconds=20.0):
    """the reference's own host policy, OMP/MKL_NUM_THREADS=1 (audiossl/__init__.py:1-3): one clip per step."""
    import torch
    prev = torch.get_num_threads()
    step, kind = cpu_step_fn(1, 1)
    step()
    t0 = time.perf_counter()
    k = 0
    while k < 1 or (time.perf_counter() - t0 < seconds and k < 3):
        step()
        k += 1
    dt = (time.perf_counter() - t0) / k
    torch.set_num_threads(prev)
    return {"value": 1 / dt, "unit": "clips/s", "cores": 1, "kind": kind, "sample": "%d steps of 1 clip" % k}


def pick_cpu_threads():
    """a big host (128+ hardware threads) is slower with every thread on these small per-clip GEMMs/FFTs:
    probe a forward pass at a few thread counts and keep the fastest."""
    import torch
    from oracle import atst_oracle as O
    cores = os.cpu_count() or 1
    cands = sorted({c for c in (8, 16, 32, 64, cores) if c <= cores})
    enc = O.OracleAST(768, 2, 12)
    x = torch.randn(4, 1, 64, 1001)
    ln = torch.full((4,), 1001, dtype=torch.int64)
    best, best_t = cands[0], 1e30
    for c in cands:
        torch.set_num_threads(c)
        with torch.no_grad():
            enc(x, ln)
            t0 = time.perf_counter()
            enc(x, ln)
            dt = time.perf_counter() - t0
        if dt < best_t:
            best, best_t = c, dt
    return best


def run_reference(args):
    import torch
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    batch = 4
    cores = pick_cpu_threads()
    step, kind = cpu_step_fn(batch, cores)
    for _ in range(max(args.warmup, 1)):
        step()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step()
    dt = (time.perf_counter() - t0) / args.steps
    val = batch / dt
    what = "the reference's own modules" if kind == "reference" else "oracle port of the reference algorithm"
    line = {"impl": "reference", "metric": "ATST-base clips/sec (student+teacher fwd + bwd)", "value": val,
            "unit": "clips/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic",
            "config": {"workload": "ATST-base, 10 s clips, 256 clips/GPU [c2], 16 kHz, 64 mel, 2 crops, DropPath 0.1; "
                                   "CPU sample of %d clips/step" % batch},
            "cpu_baseline": {"value": val, "unit": "clips/s", "cores": torch.get_num_threads(), "kind": kind,
                             "sample": "%d steps of %d clips (%s, torch CPU fp32)" % (args.steps, batch, what),
                             "one_thread": one_thread_number()},
            "e2e": {"value": val, "unit": "clips/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------ stock PyTorch on the GPU
def run_torch_gpu(args):
    """Comparator: the same modules in stock PyTorch-CUDA on one B200 - torchaudio's
    MelSpectrogram / AmplitudeToDB front-end, the oracle's nn.Module restatement of ATST (cuBLAS / ATen kernels,
    autograd), fused torch AdamW, EMA - in fp32 with TF32 matmuls allowed (or --no-tf32).  This is the only GPU path
    the reference had before this repo; none of this repo's kernels run here."""
    import torch
    from oracle import atst_oracle as O
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    dev = torch.device("cuda", 0)
    torch.backends.cuda.matmul.allow_tf32 = not args.no_tf32
    torch.backends.cudnn.allow_tf32 = not args.no_tf32
    torch.manual_seed(0)
    model = O.OracleATST("base").to(dev).train()
    opt = torch.optim.AdamW([p for p in model.student.parameters()], lr=2e-4, eps=1e-6, weight_decay=0.04, fused=True)
    try:
        import torchaudio
        melspec = torchaudio.transforms.MelSpectrogram(16000, f_min=60, f_max=7800, hop_length=160, win_length=1024,
                                                       n_fft=1024, n_mels=64).to(dev)
        to_db = torchaudio.transforms.AmplitudeToDB(stype="power", top_db=80)
        mel = lambda w: (to_db(melspec(w)) + 79.6482) / (50.6842 + 79.6482) * 2.0 - 1.0
        front = "torchaudio"
    except Exception:  # noqa: BLE001
        fb = torch.from_numpy(O.mel_filterbank()).to(dev)
        win = torch.hann_window(1024, device=dev)

        def mel(w):
            spec = torch.stft(w[:, 0], 1024, 160, 1024, win, center=True, pad_mode="reflect", return_complex=True).abs() ** 2
            db = 10.0 * torch.log10(torch.clamp(fb.t() @ spec, min=1e-10))
            db = torch.maximum(db, db.amax(dim=(-2, -1), keepdim=True) - 80.0)
            return ((db + 79.6482) / (50.6842 + 79.6482) * 2.0 - 1.0)[:, None]
        front = "torch.stft"
    B = args.batch if args.batch > 0 else 256
    depth, rates = 12, torch.linspace(0, 0.1, 12).tolist()

    def dp(S):
        out = []
        for r in rates:
            if r == 0:
                out.append(None)
                continue
            keep = 1 - r
            u = torch.rand(2, S, device=dev)
            out.append((torch.floor(keep + u[0]) / keep, torch.floor(keep + u[1]) / keep))
        return [out]

    while True:
        try:
            g = torch.Generator(device=dev).manual_seed(1234)
            wav = torch.randn(2, B, 1, int(CLIP_SECONDS * SR), device=dev, generator=g) * 0.1
            lengths = [torch.full((B,), 1001, dtype=torch.int64, device=dev)] * 2

            def step():
                crops = [mel(wav[v]) for v in range(2)]
                opt.zero_grad(set_to_none=True)
                loss, _, _ = model(crops, lengths, dp_student=dp(2 * B), dp_teacher=dp(2 * B))
                loss.backward()
                opt.step()
                model.update_teacher(0.9995)
                return loss
            for _ in range(max(args.warmup, 1)):
                step()
            torch.cuda.synchronize()
            break
        except torch.OutOfMemoryError:
            opt.zero_grad(set_to_none=True)
            torch.cuda.empty_cache()
            B //= 2
            if B < 8:
                raise
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        step()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / args.steps
    print(json.dumps({"impl": "torch-gpu", "metric": "ATST-base clips/sec (student+teacher fwd + bwd)",
                      "value": B / (ms / 1e3), "unit": "clips/s", "n_gpus": 1, "steps": args.steps,
                      "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
                      "dtype": "f32" if args.no_tf32 else "tf32 (torch allow_tf32, cuBLAS)", "data": "synthetic",
                      "config": {"workload": "ATST-base, 10 s clips, %d clips/GPU [c2], DropPath 0.1, %s mel + "
                                             "nn.Module ATST (autograd) + fused torch AdamW + EMA" % (B, front),
                                 "peak_mem_gb": torch.cuda.max_memory_allocated() / 1e9}}), flush=True)


# ------------------------------------------------------------------------------------------------ GPU arm
def _half_stream():
    from audiossl_b200 import engine
    return bool(engine.half_dgelu())


def run_ours(args):
    import torch
    import torch.distributed as dist
    from audiossl_b200 import ops
    from audiossl_b200.methods.atst.model import ATSTLightningModule
    from audiossl_b200.transforms import LogMelSpectrogram

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        # stdout carries the JSON line only: NCCL prints its version banner (and any NCCL_DEBUG output) to fd 1 when the
        # communicator is created, so create it with fd 1 pointed at stderr
        sys.stdout.flush()
        saved_fd = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=dev)
            warm = torch.zeros(1, device=dev)
            dist.all_reduce(warm)
            torch.cuda.synchronize()
        finally:
            sys.stdout.flush()
            os.dup2(saved_fd, 1)
            os.close(saved_fd)
    cfg = CONFIGS[args.config]
    B = args.batch if args.batch > 0 else cfg["batch"]
    arch = args.arch or cfg["arch"]
    torch.manual_seed(0)
    ncrops = sum(c for _, c in cfg["crops"])
    if cfg["kind"] == "frame":
        from audiossl_b200.methods.atstframe.model import FrameATSTLightningModule
        from audiossl_b200.methods.atstframe import random_mask
        lm = FrameATSTLightningModule(arch=arch, learning_rate=8e-5, warmup_steps=10, max_steps=100000, ema=0.9996)
    else:
        lm = ATSTLightningModule(arch=arch, learning_rate=2e-4, warmup_steps=10, max_steps=100000, ema=0.9995,
                                 ncrops=ncrops)
    lm.cuda().train()
    opt = lm.configure_optimizers()[0]
    lm.trainer.optimizers = [opt]
    mel = LogMelSpectrogram()
    g = torch.Generator(device=dev).manual_seed(1234 + rank)
    # one device / pinned-host waveform tensor per crop group: [count, B, 1, n]
    wav_dev = [torch.randn(cnt, B, 1, int(sec * SR), device=dev, generator=g) * 0.1 for sec, cnt in cfg["crops"]]
    wav_host = [w.cpu().pin_memory() for w in wav_dev]
    lengths = []
    for sec, cnt in cfg["crops"]:
        lengths += [torch.full((B,), int(sec * SR) // 160 + 1, device=dev, dtype=torch.int64)] * cnt
    n = int(cfg["crops"][0][0] * SR)
    masks = None
    if cfg["kind"] == "frame":
        import numpy as np
        np.random.seed(1234 + rank)
        P = (n // 160 + 1) // 4
        mk = random_mask.get_mask(B, P, 0.65, no_overlap=False, min_length=5).to(dev)
        masks = [mk, mk]
    h2d_bytes = sum(w.numel() * 4 for w in wav_host)

    mel_events = []
    # e2e input pipeline (audiossl_b200.datasets.DevicePrefetcher): device staging filled from pinned host memory on a
    # copy stream, so the host->device copy of step i+1 (inside the timed region) overlaps the compute of step i
    from audiossl_b200.datasets import DevicePrefetcher

    def host_batches():
        while True:
            yield tuple(wav_host)
    feed = {"it": None, "aug": None}

    def step(i, from_host, augment=False):
        if from_host:
            if feed["it"] is None:
                feed["it"] = iter(DevicePrefetcher(host_batches(), dev))
            src = next(feed["it"])
        else:
            src = wav_dev
        if ops.STATS["time_gemms"]:
            ev = (torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
            ev[0].record()
        if augment:  # the recipe's device transform: window crop + mel + Mixup + RandomResizeCrop per view
            crops, lens = feed["aug"](src[0][0])
        else:
            crops, lens = [mel(w[k]) for w in src for k in range(w.shape[0])], lengths
        if ops.STATS["time_gemms"]:
            ev[1].record()
            mel_events.append(ev)
        lm.global_step = i
        batch = ((crops, lens, masks), None) if masks is not None else ((crops, lens), None)
        loss = lm.training_step(batch, i)
        opt.zero_grad()
        loss.backward()
        opt.step()
        lm.on_train_batch_end(None, None, i)
        if from_host:
            return loss.item()  # device -> host read of the step's result
        return loss

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(from_host, steps, start_i, augment=False):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(steps):
            step(start_i + i, from_host, augment)
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return ms.item()

    for i in range(args.warmup):
        step(i, False)
    if args.graph:  # the step after the mel as one CUDA graph (audiossl_b200.graph; launch-bound small batches)
        from audiossl_b200.graph import GraphedTrainStep
        if cfg["kind"] != "clip" or world > 1:
            raise SystemExit("--graph: single-GPU ATST-clip configs only")
        crops0 = [mel(w[k]) for w in wav_dev for k in range(w.shape[0])]
        gstep = GraphedTrainStep(lm, opt, ((crops0, lengths), None))
        eager_step = step

        def step(i, from_host, augment=False):  # noqa: F811
            if from_host or augment:
                return eager_step(i, from_host, augment)
            crops = [mel(w[k]) for w in wav_dev for k in range(w.shape[0])]
            return gstep(((crops, lengths), None), i)
        for i in range(3):
            step(args.warmup + i, False)
    if args.profile:  # short run for ncu: one more step, nothing else
        torch.cuda.synchronize()
        step(args.warmup, False)
        torch.cuda.synchronize()
        return
    if args.gaps:  # in-situ kernel timeline of one step (CUPTI through torch.profiler): busy vs idle GPU time
        from torch.profiler import profile, ProfilerActivity
        torch.cuda.synchronize()
        with profile(activities=[ProfilerActivity.CUDA]) as prof:
            for k_ in range(max(args.steps, 1)):
                step(args.warmup + k_, False)
            torch.cuda.synchronize()
        ks = sorted(((e.time_range.start, e.time_range.end, e.name) for e in prof.events()
                     if e.device_type == torch.autograd.DeviceType.CUDA), key=lambda t: t[0])
        span = ks[-1][1] - ks[0][0]
        # union of the busy intervals (kernels of the compute stream and NCCL kernels of the exchange stream overlap)
        busy, cur_a, cur_b = 0.0, None, None
        for a, b, _ in ks:
            if cur_b is None or a > cur_b:
                if cur_b is not None:
                    busy += cur_b - cur_a
                cur_a, cur_b = a, b
            else:
                cur_b = max(cur_b, b)
        busy += cur_b - cur_a
        agg = {}
        for a, b, nm in ks:
            d = agg.setdefault(nm[:60], [0, 0.0])
            d[0] += 1
            d[1] += (b - a) / 1e3
        nccl = [(a, b, nm) for a, b, nm in ks if "nccl" in nm.lower()]
        mine = [(a, b, nm) for a, b, nm in ks if "nccl" not in nm.lower()]
        # NCCL time that no kernel of ours overlaps = the exposed part of the exchange
        exposed = 0.0
        j = 0
        for a, b, _ in nccl:
            covered, pos = 0.0, a
            while j < len(mine) and mine[j][1] <= a:
                j += 1
            k = j
            while k < len(mine) and mine[k][0] < b:
                lo, hi = max(mine[k][0], pos), min(mine[k][1], b)
                if hi > lo:
                    covered += hi - lo
                    pos = hi
                k += 1
            exposed += (b - a) - covered
        gaps = sorted(((mine[i + 1][0] - mine[i][1], mine[i][2][:40], mine[i + 1][2][:40]) for i in range(len(mine) - 1)),
                      reverse=True)
        nsteps = max(args.steps, 1)
        out = open(args.gaps_out % rank, "w") if args.gaps_out else sys.stdout
        print("rank %d of %d  steps %d  kernels %d  span %.2f ms/step  busy %.2f ms/step  idle %.2f ms/step" %
              (rank, world, nsteps, len(ks), span / 1e3 / nsteps, busy / 1e3 / nsteps, (span - busy) / 1e3 / nsteps),
              file=out)
        print("NCCL kernels: %d per step, %.3f ms per step, of which %.3f ms not overlapped by a kernel of this repo" %
              (len(nccl) // nsteps, sum(b - a for a, b, _ in nccl) / 1e3 / nsteps, exposed / 1e3 / nsteps), file=out)
        for nm, (n_, ms_) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:25]:
            print("  %-60s n=%4d %8.3f ms/step" % (nm, n_ // nsteps, ms_ / nsteps), file=out)
        print("largest gaps between this repo's kernels (us):", file=out)
        for g_, a_, b_ in gaps[:15]:
            print("  %8.1f  %s -> %s" % (g_, a_, b_), file=out)
        if out is not sys.stdout:
            out.close()
        if world > 1:
            dist.barrier()
            dist.destroy_process_group()
        return
    sampler = ClockSampler(local)
    sampler.start()
    ops.reset_stats()
    ms = timed(False, args.steps, args.warmup)
    launches = ops.STATS["launches"]
    if args.graph:  # the remaining legs (host inputs, per-kernel timing) run the eager step
        launches = args.steps * (len(crops0) // max(len(crops0), 1) * 2 + 1)  # per step: the mel launches + 1 graph
        gstep.release()
        step = eager_step
        for i in range(2):
            step(args.warmup + args.steps + 10 + i, False)
        ops.reset_stats()
        step(args.warmup + args.steps + 12, False)
        ops.STATS["gemm_flops"] *= args.steps
        ops.STATS["gemm_bytes"] *= 1
    gemm_flops_step = ops.STATS["gemm_flops"] / args.steps
    gemm_bytes_launch = ops.STATS["gemm_bytes"] / max(ops.STATS["gemm_launches"], 1)
    feed["it"] = None
    for i in range(2):  # untimed: the staging buffers of the host path are allocated and touched once
        step(args.warmup + args.steps + i, True)
    ms_e2e = timed(True, args.steps, args.warmup + args.steps + 2)
    ms_aug = None
    if cfg["kind"] == "clip" and len(cfg["crops"]) == 1 and not args.no_augment:
        # SURVEY.md 8d "second number with augmentations on": same e2e loop, the batch produced by the device
        # train transform (two random full-length windows of each clip, mel, Mixup memory bank, RandomResizeCrop)
        from audiossl_b200.methods.atst.transform import BatchedATSTTrainTransform
        import numpy as np
        sec = cfg["crops"][0][0]
        feed["aug"] = BatchedATSTTrainTransform(anchor_len=(sec, sec), positive_len=(sec, sec),
                                                rng=np.random.RandomState(1234 + rank))
        feed["it"] = None
        for i in range(2):
            step(i, True, True)
        ms_aug = timed(True, args.steps, args.warmup + 2 * args.steps, True)
    sampler.stop_flag = True
    sampler.join(timeout=2)
    # dominant kernel (gemm_tf32_kernel) measured in place: one extra step with an event pair around every launch
    ops.STATS["time_gemms"] = True
    ops.reset_stats()
    torch.cuda.synchronize()
    step(args.warmup + 2 * args.steps, False)
    torch.cuda.synchronize()
    ops.STATS["time_gemms"] = False
    gemm_ms = sum(a.elapsed_time(b) for a, b, _, _ in ops.STATS["gemm_events"])
    gemm_fl = sum(f for _, _, f, _ in ops.STATS["gemm_events"])
    n_gemm = len(ops.STATS["gemm_events"])
    if rank == 0 and args.breakdown:
        agg = {}
        for a, b2, f, tag in ops.STATS["gemm_events"]:
            e = agg.setdefault(tag, [0, 0.0, 0.0])
            e[0] += 1
            e[1] += a.elapsed_time(b2)
            e[2] += f
        with open(args.breakdown, "w") as fh:
            for tag, (cnt, t_ms, fl) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
                fh.write("%-28s n=%3d  %8.3f ms  %7.1f TFLOP/s\n" % (tag, cnt, t_ms, fl / t_ms / 1e9))

    mel_ms = sum(a.elapsed_time(b2) for a, b2 in mel_events)
    attn = {}
    for a_, b_, nb_, tag_ in ops.STATS["attn_events"]:
        d_ = attn.setdefault(tag_, [0, 0.0, 0.0])
        d_[0] += 1
        d_[1] += a_.elapsed_time(b_)
        d_[2] += nb_
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    pk = peaks()
    # 4 n + 4 * 64 * (n // 160 + 1) bytes per clip-view (896 256 B for 10 s)
    mel_bytes = sum(cnt * B * (4 * int(sec * SR) + 4 * 64 * (int(sec * SR) // 160 + 1)) for sec, cnt in cfg["crops"])
    step_gflop = cfg["gflop"]
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "gemm_traffic.json")
    if args.config == "c2" and os.path.exists(tpath):
        traffic = json.load(open(tpath)).get("dram_bytes_per_launch")
    mel_traffic = None
    mpath = os.path.join(ROOT, "profiles", "mel_traffic.json")
    if args.config == "c2" and os.path.exists(mpath):  # ncu DRAM bytes per 10 s clip-view x the views of one launch
        mel_traffic = json.load(open(mpath))["dram_bytes_per_clip_view_10s"] * B
    ms_step = ms / args.steps
    value = B * world / (ms_step / 1e3)
    e2e_val = B * world / (ms_e2e / args.steps / 1e3)
    achieved = gemm_fl / (gemm_ms / 1e3) / 1e12 if gemm_ms > 0 else 0.0
    line = {
        "metric": "ATST-base clips/sec (student+teacher fwd + bwd)", "value": value, "unit": "clips/s",
        "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_step,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "tf32",
        "data": "synthetic",
        "config": {"workload": "%s [%s], 16 kHz, 64 mel, %d crops, %d clips/GPU, DropPath 0.1, "
                               "mel+teacher fwd+student fwd+loss+bwd+AdamW+EMA" % (cfg["name"], args.config, ncrops, B),
                   "parallelism": "dp%d" % world, "l2": "inputs_exceed_l2 (tens of GB of activations per step)",
                   "step_gflop_per_clip_algorithmic": step_gflop,
                   "arithmetic": "TF32 tcgen05 products with fp32 accumulation (heads: 3xTF32), everything else fp32"
                                 + ("; stored activations fp32 except the student MLP's gelu'(u), kept as fp16 = the "
                                    "10-bit mantissa of the TF32 rounding its product gets anyway (DESIGN.md section 3)"
                                    if _half_stream() else "")},
        "clocks": sampler.summary(),
        "e2e": {"value": e2e_val, "unit": "clips/s", "h2d_bytes_per_step": h2d_bytes, "d2h_bytes_per_step": 4},
        "e2e_augmented": None if ms_aug is None else {
            "value": B * world / (ms_aug / args.steps / 1e3), "unit": "clips/s",
            "what": "e2e with the recipe's augmentations on the device (BatchedATSTTrainTransform: random window, "
                    "mel, Mixup memory bank, RandomResizeCrop) instead of the plain mel"},
        "gpu_launches": launches,
        "cuda_graph": bool(args.graph),
        "roofline": {"bound": "tensor", "kernel": "gemm2_tf32_kernel (CTA pair, tcgen05 kind::tf32)", "achieved": achieved,
                     "peak": pk["tflops"], "unit": "TFLOP/s", "frac": achieved / pk["tflops"], "traffic": traffic,
                     "algorithmic_bytes_per_launch": gemm_bytes_launch,
                     "peak_source": pk["src"] + " bf16 cuBLAS sustained; TF32 tensor rate is half of bf16",
                     "launches_per_step": n_gemm, "gemm_ms_per_step": gemm_ms,
                     "gemm_share_of_step": gemm_ms / ms_step,
                     "algorithmic_gemm_tflop_per_step": gemm_flops_step / 1e12,
                     "model_flops_utilisation_of_step": step_gflop * 1e9 * B / (ms_step / 1e3) / 1e12 / pk["tflops"]},
        "roofline_attention": {
            "bound": "hbm", "kernel": "attn_fwd_tc_kernel / attn_bwd_tc_kernel<0,1> + attn_delta_kernel (tcgen05)",
            "unit": "GB/s", "peak": pk["hbm_gbs"],
            "note": "algorithmic bytes (qkv, o, dO read once; o / dqkv written once) over the in-place duration; "
                    "the TMEM read port (64 B/clk/SM) is the co-limiter, DESIGN.md 5.2",
            **{k_: {"launches_per_step": v_[0], "ms_per_step": v_[1],
                    "achieved": v_[2] / (v_[1] / 1e3) / 1e9 if v_[1] > 0 else 0.0,
                    "frac": (v_[2] / (v_[1] / 1e3) / 1e9 / pk["hbm_gbs"]) if v_[1] > 0 else 0.0}
               for k_, v_ in attn.items()}},
        "roofline_mel": {"bound": "hbm", "kernel": "mel_kernel (fused STFT/mel/dB/top_db clamp/MinMax, one launch)",
                         "achieved": mel_bytes / (mel_ms / 1e3) / 1e9 if mel_ms > 0 else 0.0, "peak": pk["hbm_gbs"],
                         "unit": "GB/s", "frac": (mel_bytes / (mel_ms / 1e3) / 1e9 / pk["hbm_gbs"]) if mel_ms > 0 else 0.0,
                         "ms_per_step": mel_ms, "algorithmic_bytes_per_step": mel_bytes, "traffic": mel_traffic,
                         "note": "per launch = one view of the batch; bound by instruction issue (ncu: issue slots 48 % busy at 37 % "
                                 "occupancy, DRAM 1.6 %), not by HBM: 40 flop/B"},
    }
    if world == 1 and not args.no_cpu_baseline and args.config == "c2":
        cores = pick_cpu_threads()
        cstep, kind = cpu_step_fn(4, cores)
        cstep()
        t0 = time.perf_counter()
        k = 0
        while k < 2 or (time.perf_counter() - t0 < 10.0 and k < 8):
            cstep()
            k += 1
        dt = (time.perf_counter() - t0) / k
        line["cpu_baseline"] = {"value": 4 / dt, "unit": "clips/s", "cores": torch.get_num_threads(), "kind": kind,
                                "sample": "%d steps of 4 clips of the same workload (%s, torch CPU fp32)"
                                          % (k, "reference modules" if kind == "reference" else "oracle port"),
                                "one_thread": one_thread_number(10.0)}
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference", "torch-gpu"])
    ap.add_argument("--no-tf32", action="store_true", help="--impl torch-gpu: strict fp32 matmuls")
    ap.add_argument("--no-augment", action="store_true", help="skip the e2e_augmented measurement")
    ap.add_argument("--config", default="c2", choices=sorted(CONFIGS), help="BASELINE.json config (c2 = benchmark)")
    ap.add_argument("--batch", type=int, default=0, help="clips per GPU (0 = the config's)")
    ap.add_argument("--arch", default="", help="override the config's architecture")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--profile", action="store_true", help="warm-up + one step only (for ncu captures)")
    ap.add_argument("--graph", action="store_true", help="replay the step (after the mel) as one CUDA graph")
    ap.add_argument("--gaps", action="store_true", help="in-situ kernel timeline of one step: busy / idle GPU time")
    ap.add_argument("--gaps-out", default="", help="per-rank output file pattern for --gaps, e.g. out/timeline_rank%%d.txt")
    ap.add_argument("--breakdown", default="", help="write a per-GEMM-shape timing table to this file")
    a = ap.parse_args()
    if a.warmup < 3 and a.impl == "ours" and not a.profile and not a.gaps:
        a.warmup = 3
    if a.impl == "reference":
        run_reference(a)
    elif a.impl == "torch-gpu":
        run_torch_gpu(a)
    else:
        run_ours(a)
