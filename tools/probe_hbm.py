"""HBM direction probe: write-only (fill), read-only (sum) and copy bandwidth of large fp32 buffers, CUDA events,
best of 5.  Context for the two-output-stream GEMM epilogues: is a write-heavy kernel bounded by less
than the copy figure of MEASURED_PEAKS.json?"""
import torch

n = 2 * 1024 ** 3  # 8 GiB of fp32
a = torch.empty(n, device="cuda")
b = torch.empty(n, device="cuda")


def best(f, nbytes):
    ts = []
    for _ in range(6):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        f()
        e1.record()
        e1.synchronize()
        ts.append(e0.elapsed_time(e1))
    t = min(ts[1:])
    return nbytes / t * 1e-6


print("write-only  fill_   %7.0f GB/s" % best(lambda: a.fill_(1.0), 4 * n))
print("read-only   sum     %7.0f GB/s" % best(lambda: a.sum(), 4 * n))
print("copy        copy_   %7.0f GB/s (read + write bytes)" % best(lambda: b.copy_(a), 8 * n))
print("two writes  fill x2 %7.0f GB/s" % best(lambda: (a.fill_(2.0), b.fill_(3.0)), 8 * n))
