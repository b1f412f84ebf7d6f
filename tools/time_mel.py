"""mel_kernel timing: 64 / 256 clips x 10 s, CUDA events, inputs larger than L2 rotated between launches."""
import sys
import torch
sys.path.insert(0, __file__.rsplit("/", 2)[0])
from audiossl_b200 import ops

for B in (64, 256):
    wavs = [torch.randn(B, 1, 160000, device="cuda") * 0.1 for _ in range(4)]  # 4 x 41 / 164 MB: rotates through L2
    out = torch.empty(B, 64, 1001, device="cuda")
    for w in wavs:
        ops.mel_forward(w, out=out)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n = 20
    e0.record()
    for i in range(n):
        ops.mel_forward(wavs[i % 4], out=out)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / n
    nbytes = B * 896256
    print("mel %3d clips x 10 s: %.4f ms per launch, %.1f GB/s algorithmic" % (B, ms, nbytes / ms / 1e6))
