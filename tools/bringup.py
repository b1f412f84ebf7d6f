"""GPU bring-up diagnostics: each kernel family against a plain torch fp32 computation.

    python tools/bringup.py <section> [...]     sections: mel gemm_nt gemm_mn ln attn bn loss optim tokens
    python tools/bringup.py all                 runs every section in its own subprocess with a timeout

Not a test-suite (tests/ has the parity tests); this prints error tables used while bringing the
kernels up on hardware, e.g. the shared-memory descriptor sweep for the TF32 MN-major GEMM.
"""
import itertools
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def rel_err(a, b):
    return ((a - b).norm() / b.norm().clamp_min(1e-30)).item()


def sec_mel():
    import numpy as np
    import torch
    from audiossl_b200 import ops
    from oracle import atst_oracle as O
    from tests.golden import detfill
    for kind, n in [("noise", 16000), ("sine_silence", 16000), ("chirp", 16000), ("zeros", 16000), ("noise", 1600),
                    ("noise", 160000)]:
        for win in (1024, 640):
            wav = detfill.signal(kind, n)[None]
            ref = O.mel_feature(wav, win_length=win)
            out = ops.mel_forward(torch.from_numpy(wav).cuda(), win_length=win).cpu().numpy()
            d = np.abs(out - ref)
            print("mel %-13s n=%6d win=%4d max|d|=%.3e mean|d|=%.3e" % (kind, n, win, d.max(), d.mean()))
    wav = torch.randn(64, 160000, device="cuda") * 0.1
    torch.cuda.synchronize()
    for _ in range(3):
        ops.mel_forward(wav)
    torch.cuda.synchronize()
    t = time.time()
    for _ in range(10):
        ops.mel_forward(wav)
    torch.cuda.synchronize()
    dt = (time.time() - t) / 10
    print("mel 64x10s: %.3f ms  -> %.1f GB/s algorithmic" % (dt * 1e3, 64 * 896256 / dt / 1e9))


def sec_gemm_nt():
    import torch
    from audiossl_b200 import ops
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.manual_seed(0)
    for (M, N, K) in [(128, 128, 32), (128, 256, 64), (256, 256, 128), (156, 384, 128), (300, 128, 256),
                      (1004, 2304, 768), (4096, 768, 3072), (70, 4096, 128), (512, 256, 4096)]:
        A = ops.round_tf32(torch.randn(M, K, device="cuda"))
        B = ops.round_tf32(torch.randn(N, K, device="cuda") * 0.05)
        ref = A @ B.t()
        out = ops.gemm_nt(A, B)
        torch.cuda.synchronize()
        print("gemm_nt %5dx%5dx%5d store rel=%.3e max=%.3e" % (M, N, K, rel_err(out, ref), (out - ref).abs().max().item()))
    M, N, K = 520, 768, 256
    A = ops.round_tf32(torch.randn(M, K, device="cuda"))
    B = ops.round_tf32(torch.randn(N, K, device="cuda") * 0.05)
    bias = torch.randn(N, device="cuda")
    resid = torch.randn(M, N, device="cuda")
    scale = torch.rand(M // 26, device="cuda") + 0.5
    base = A @ B.t() + bias
    out = ops.gemm_nt(A, B, bias=bias)
    print("epi bias      rel=%.3e" % rel_err(out, base))
    aux = torch.empty(M, N, device="cuda")
    out = ops.gemm_nt(A, B, bias=bias, epi=ops.EPI_GELU, aux=aux)
    print("epi gelu      rel=%.3e aux rel=%.3e" % (rel_err(out, torch.nn.functional.gelu(base)), rel_err(aux, base)))
    u = torch.randn(M, N, device="cuda")
    ug = u.clone().requires_grad_(True)
    torch.nn.functional.gelu(ug).sum().backward()
    out = ops.gemm_nt(A, B, epi=ops.EPI_DGELU, aux=u)
    print("epi dgelu     rel=%.3e" % rel_err(out, (A @ B.t()) * ug.grad))
    out = ops.gemm_nt(A, B, bias=bias, epi=ops.EPI_RESID, resid=resid, rowscale=scale, rows_per_seq=26)
    print("epi resid     rel=%.3e" % rel_err(out, resid + scale.repeat_interleave(26)[:, None] * base))
    out = ops.gemm_nt(A, B, bias=bias, epi=ops.EPI_RELU)
    print("epi relu      rel=%.3e" % rel_err(out, base.clamp_min(0)))
    # throughput
    for (M, N, K) in [(128512, 2304, 768), (128512, 768, 3072), (128512, 3072, 768)]:
        A = torch.randn(M, K, device="cuda")
        B = torch.randn(N, K, device="cuda")
        C = torch.empty(M, N, device="cuda")
        for _ in range(2):
            ops.gemm_nt(A, B, out=C)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5):
            ops.gemm_nt(A, B, out=C)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 5
        print("gemm_nt %dx%dx%d: %.3f ms  %.1f TFLOP/s" % (M, N, K, ms, 2.0 * M * N * K / ms / 1e9))
        del A, B, C


def sec_gemm_mn():
    """descriptor sweep for token-major (MN-major) TF32 operands."""
    import torch
    from audiossl_b200 import _lib
    from audiossl_b200._lib import ptr
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.manual_seed(0)
    from audiossl_b200 import ops
    T, M, N = 96, 128, 128
    A = ops.round_tf32(torch.randn(T, M, device="cuda"))
    B = ops.round_tf32(torch.randn(T, N, device="cuda"))
    ref_tn = A.t() @ B
    X = ops.round_tf32(torch.randn(M, T, device="cuda"))  # for nn: X[M,K=T] @ B[K=T,N]
    ref_nn = X @ B
    L = _lib.lib()
    good = []
    combos = list(itertools.product((1, 2), (4, 3, 5), ((4096, 512), (512, 4096), (4096, 1024), (1024, 4096),
                                                        (4096, 256), (256, 4096)), (1024, 512)))
    for layout, swz, (lbo, sbo), kstep in combos:
        for nn in (0, 1):
            C = torch.zeros(M, N, device="cuda")
            if nn:
                rc = L.atst_gemm_mn_debug(1, ptr(X), T, ptr(B), N, ptr(C), N, M, N, T, lbo, sbo, kstep, layout, swz, 1,
                                          _lib.stream())
            else:
                rc = L.atst_gemm_mn_debug(0, ptr(A), M, ptr(B), N, ptr(C), N, M, N, T, lbo, sbo, kstep, layout, swz, 1,
                                          _lib.stream())
            try:
                torch.cuda.synchronize()
            except Exception as e:  # sticky error: report and stop (rerun remaining combos separately)
                print("CUDA error at layout=%d swz=%d lbo=%d sbo=%d kstep=%d nn=%d: %s" % (layout, swz, lbo, sbo, kstep, nn, e))
                return
            err = rel_err(C, ref_nn if nn else ref_tn) if rc == 0 else float("nan")
            tag = "OK " if err < 2e-3 else "   "
            print("%s layout=%d swz=%d lbo=%4d sbo=%4d kstep=%4d %s rc=%d rel=%.3e" %
                  (tag, layout, swz, lbo, sbo, kstep, "nn" if nn else "tn", rc, err))
            if err < 2e-3:
                good.append((layout, swz, lbo, sbo, kstep, nn))
    print("GOOD:", good)
    # defaults at real sizes
    for (T, M, N) in [(1000, 256, 256), (4000, 768, 2304), (128512, 768, 768), (128512, 3072, 768)]:
        A = ops.round_tf32(torch.randn(T, M, device="cuda") * 0.1)
        B = ops.round_tf32(torch.randn(T, N, device="cuda") * 0.1)
        C = torch.zeros(M, N, device="cuda")
        ops.gemm_tn_acc(A, B, C)
        torch.cuda.synchronize()
        print("gemm_tn default T=%d %dx%d rel=%.3e" % (T, M, N, rel_err(C, A.t() @ B)))
        W = ops.round_tf32(torch.randn(M, N, device="cuda") * 0.1)  # weight [out=M, in=N]
        dY = A  # [T, M]
        dX = ops.gemm_nn(dY, W)
        torch.cuda.synchronize()
        print("gemm_nn default T=%d K=%d N=%d rel=%.3e" % (T, M, N, rel_err(dX, dY @ W)))


def sec_ln():
    import torch
    from audiossl_b200 import ops
    torch.manual_seed(0)
    for D in (128, 384, 768, 1024):
        rows = 333
        x = torch.randn(rows, D, device="cuda") * 2 + 0.5
        g = torch.randn(D, device="cuda")
        b = torch.randn(D, device="cuda")
        xr = x.clone().requires_grad_(True)
        gr, br = g.clone().requires_grad_(True), b.clone().requires_grad_(True)
        y_ref = torch.nn.functional.layer_norm(xr, (D,), gr, br, 1e-6)
        dy = torch.randn(rows, D, device="cuda")
        y_ref.backward(dy)
        y, mean, rstd = ops.layernorm_fwd(x, g, b, rows, D, round_out=False)
        dg, db = torch.zeros(D, device="cuda"), torch.zeros(D, device="cuda")
        dres = torch.randn(rows, D, device="cuda")
        dys = torch.empty(rows, D, device="cuda")
        cs = torch.zeros(D, device="cuda")
        sc = torch.rand((rows + 36) // 37, device="cuda") + 0.5
        dx = ops.layernorm_bwd(dy, x, mean, rstd, g, dg, db, rows, D, dres=dres, dys=dys, rowscale=sc, rows_per_seq=37,
                               colsum_out=cs)
        ref_dys = (xr.grad + dres) * sc.repeat_interleave(37)[:rows, None]
        print("ln D=%4d fwd=%.2e dx=%.2e dg=%.2e db=%.2e dys=%.2e colsum=%.2e" %
              (D, rel_err(y, y_ref), rel_err(dx, xr.grad + dres), rel_err(dg, gr.grad), rel_err(db, br.grad),
               rel_err(dys, ref_dys), rel_err(cs, ref_dys.sum(0))))


def sec_umma_probe():
    """which tcgen05 operand forms work: K-major reads of 32B-atom-swizzled tiles, A operand in tensor memory."""
    import torch
    from audiossl_b200 import _lib, ops
    from audiossl_b200._lib import ptr
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.manual_seed(0)
    L = _lib.lib()
    A = ops.round_tf32(torch.randn(128, 64, device="cuda"))
    B = ops.round_tf32(torch.randn(128, 64, device="cuda"))
    B2 = ops.round_tf32(torch.randn(64, 128, device="cuda"))
    ref = A @ B.t()
    for mode, name, r, b in ((1, "A in TMEM, B K-major", ref, B), (2, "A in TMEM, B token-major", A @ B2, B2)):
        D = torch.zeros(128, 128, device="cuda")
        rc = L.atst_umma_probe(mode, ptr(A), ptr(b), ptr(D), 0, 0, 0, 0, _lib.stream())
        torch.cuda.synchronize()
        print("mode %d (%s): rc=%d rel=%.3e" % (mode, name, rc, rel_err(D, r)))
    combos = list(itertools.product((1, 2), (16, 4096, 16384), (256, 512, 1024, 2048), (32, 64)))
    if os.environ.get("PROBE_COMBO"):  # one combination per process: a bad descriptor leaves a sticky error
        combos = [tuple(int(x) for x in os.environ["PROBE_COMBO"].split(","))]
    for layout, lbo, sbo, kstep in combos:
        D = torch.zeros(128, 128, device="cuda")
        rc = L.atst_umma_probe(0, ptr(A), ptr(B), ptr(D), layout, lbo, sbo, kstep, _lib.stream())
        try:
            torch.cuda.synchronize()
        except Exception as e:
            print("CUDA error at layout=%d lbo=%d sbo=%d kstep=%d: %s" % (layout, lbo, sbo, kstep, str(e).splitlines()[0]))
            return
        err = rel_err(D, ref)
        print("%s mode 0 layout=%d lbo=%5d sbo=%4d kstep=%3d rel=%.3e" % ("OK " if err < 2e-3 else "   ", layout, lbo, sbo, kstep, err))


def sec_gemm_trace():
    """per-tile timeline of the CTA-pair GEMM (fc1 shape) for the plain, GELU and GELU+aux epilogues."""
    import torch
    from audiossl_b200 import _lib, ops
    from audiossl_b200._lib import ptr
    L = _lib.lib()
    M, D = 128512, 768
    A = ops.round_tf32(torch.randn(M, D, device="cuda"))
    W = ops.round_tf32(torch.randn(4 * D, D, device="cuda") * 0.05)
    b = torch.randn(4 * D, device="cuda")
    u = torch.empty(M, 4 * D, device="cuda")
    g = torch.empty(M, 4 * D, device="cuda")
    x = torch.randn(M, D, device="cuda")
    W2 = ops.round_tf32(torch.randn(D, D, device="cuda") * 0.05)
    b2 = torch.randn(D, device="cuda")
    o2 = torch.empty(M, D, device="cuda")
    cases = [("fc1 plain", lambda: ops.gemm_nt(A, W, bias=b, out=u)),
             ("fc1 gelu (no aux)", lambda: ops.gemm_nt(A, W, bias=b, epi=ops.EPI_GELU, aux=None, round_out=True, out=g)),
             ("fc1 gelu + aux", lambda: ops.gemm_nt(A, W, bias=b, epi=ops.EPI_GELU, aux=u, round_out=True, out=g)),
             ("proj plain", lambda: ops.gemm_nt(A, W2, bias=b2, out=o2)),
             ("proj + resid", lambda: ops.gemm_nt(A, W2, bias=b2, epi=ops.EPI_RESID, resid=x, out=o2))]
    for name, fn in cases:
        fn()
        buf = torch.zeros(32, dtype=torch.int64, device="cuda")
        L.atst_gemm_trace(ptr(buf))
        fn()
        torch.cuda.synchronize()
        L.atst_gemm_trace(None)
        t = buf.cpu().tolist()
        t0 = t[0]
        print(name)
        for k in range(4):
            r = [v - t0 for v in t[8 * k:8 * k + 7]]
            print("  tile %d  epi: arrive %6d bias %6d acc-ready %6d stored %6d | mma: arrive %6d stage-free %6d issued %6d" %
                  (8 + k, r[0], r[1], r[2], r[3], r[4], r[5], r[6]))


def sec_copy_pattern():
    """bandwidth of the GEMM epilogue's access pattern (lane = row, 32 B) against 4 lanes per row (128 B) and a
    coalesced copy: DRAM-sized [128512, 3072] and L2-resident [8192, 1024] (repeated), fp32."""
    import torch
    from audiossl_b200 import _lib
    from audiossl_b200._lib import ptr
    L = _lib.lib()
    for (M, N, reps) in ((128512, 3072, 1), (8192, 1024, 20)):
        src = torch.randn(M, N, device="cuda")
        dst = torch.empty_like(src)
        for mode, name in ((0, "lane = row, 32 B per lane (epilogue today)"), (2, "4 lanes per row, 128 B per row"),
                           (1, "coalesced float4")):
            ts = []
            for _ in range(5):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                for _ in range(reps):
                    L.atst_copy_pattern(ptr(src), ptr(dst), M, N, mode, _lib.stream())
                e1.record()
                torch.cuda.synchronize()
                ts.append(e0.elapsed_time(e1) / reps)
            t = sorted(ts)[2]
            print("[%6d x %4d] %-46s %.4f ms  %7.1f GB/s (read + write)" % (M, N, name, t, 8.0 * M * N / t / 1e6))
            assert torch.equal(src, dst)
            dst.zero_()


def sec_ew_perf():
    """HBM-bound passes at the config-2 shape: time and achieved algorithmic GB/s."""
    import torch
    from audiossl_b200 import ops
    M, D = 128512, 768
    x = torch.randn(M, D, device="cuda")
    g = torch.randn(D, device="cuda")
    b = torch.randn(D, device="cuda")
    y, mean, rstd = ops.layernorm_fwd(x, g, b, M, D)
    dy = torch.randn(M, D, device="cuda")
    dres = torch.randn(M, D, device="cuda")
    dx = torch.empty(M, D, device="cuda")
    dys = torch.empty(M, D, device="cuda")
    dg, db, cs = (torch.zeros(D, device="cuda") for _ in range(3))
    sc = torch.rand(512, device="cuda") + 0.5
    u = torch.randn(M, 4 * D, device="cuda")
    gg = torch.empty(M, 4 * D, device="cuda")
    cs4 = torch.zeros(4 * D, device="cuda")
    flush = torch.empty(64 * 1024 * 1024, device="cuda")  # 256 MB > L2

    def tm(fn, nbytes, name):
        ts = []
        for _ in range(5):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            fn()
            e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        t = sorted(ts)[len(ts) // 2]
        print("%-44s %.3f ms  %7.1f GB/s" % (name, t, nbytes / t / 1e6))

    tm(lambda: ops.layernorm_fwd(x, g, b, M, D, out=y, mean=mean, rstd=rstd), 8.0 * M * D, "ln_fwd [M,768]")
    tm(lambda: ops.layernorm_bwd(dy, x, mean, rstd, g, dg, db, M, D, dres=dres, dx=dx, dys=dys, rowscale=sc,
                                 rows_per_seq=251, colsum_out=cs), 20.0 * M * D, "ln_bwd (+dres, dys, colsum) [M,768]")
    tm(lambda: ops.gelu_fwd(u, gg), 32.0 * M * D, "gelu_fwd [M,3072]")
    tm(lambda: ops.gelu_bwd_(gg, u, colsum_out=cs4), 48.0 * M * D, "gelu_bwd + colsum [M,3072]")
    tm(lambda: ops.colsum_acc(gg, cs4), 16.0 * M * D, "colsum [M,3072]")


def sec_attn_trace():
    """clock64 timeline of one CTA of the tcgen05 attention backward kernels at the config-2 shape."""
    import torch
    from audiossl_b200 import _lib, ops
    from audiossl_b200._lib import ptr
    L = _lib.lib()
    L.atst_set_option(b"attn_tcgen05", 3)
    S, N, H = 512, 251, 12
    qkv = torch.randn(S * N, 3 * H * 64, device="cuda")
    o, lse = ops.attention_fwd(qkv, S, N, H)
    d_o = torch.randn_like(o)
    dqkv = torch.empty_like(qkv)
    ops.attention_bwd(qkv, o, d_o, lse, S, N, H, dqkv=dqkv)
    for mode in (0, 1):
        buf = torch.zeros(80, dtype=torch.int64, device="cuda")
        L.atst_attention_trace(ptr(buf), 20, mode)  # the CTA's 21st item (steady state)
        ops.attention_bwd(qkv, o, d_o, lse, S, N, H, dqkv=dqkv)
        torch.cuda.synchronize()
        L.atst_attention_trace(None, 0, 0)
        tr = buf.cpu().tolist()
        t0 = tr[0]
        rel = [x - t0 if x else -1 for x in tr]
        print("mode %d (%s) cycles since CTA start" % (mode, "dQ" if mode == 0 else "dK dV"))
        for i in range(8):
            print("  q%d  mma: P ready %6d issued %6d | C ready %6d issued %6d || compute: ready %6d done %6d" %
                  (i, rel[1 + 4 * i], rel[2 + 4 * i], rel[3 + 4 * i], rel[4 + 4 * i], rel[40 + 2 * i], rel[41 + 2 * i]))
        for t in range(2):
            print("  tile %d acc ready %6d stored %6d | producer: row buffers free, load issued %6d" %
                  (t, rel[60 + 2 * t], rel[61 + 2 * t], rel[70 + t]))


def sec_attn_tc():
    from audiossl_b200 import _lib
    _lib.lib().atst_set_option(b"attn_tcgen05", int(os.environ.get("ATTN_TC", "3")))
    _lib.lib().atst_set_option(b"attn_l2_prefetch", int(os.environ.get("ATTN_PF", "1")))
    sec_attn()


def sec_attn():
    import torch
    from audiossl_b200 import ops
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.manual_seed(0)
    for (S, N, H, ragged) in [(2, 26, 2, False), (3, 26, 2, True), (2, 151, 6, True), (4, 251, 12, True), (2, 64, 2, False),
                              (2, 65, 1, True)]:
        D = H * 64
        qkv = ops.round_tf32(torch.randn(S * N, 3 * D, device="cuda"))
        lengths = torch.full((S,), N, dtype=torch.int32, device="cuda")
        if ragged:
            lengths = torch.randint(1, N + 1, (S,), dtype=torch.int32, device="cuda")
            lengths[0] = N
        q = qkv.clone().requires_grad_(True)
        t = q.reshape(S, N, 3, H, 64).permute(2, 0, 3, 1, 4)
        att = (t[0] @ t[1].transpose(-2, -1)) * 0.125
        mask = (torch.arange(N, device="cuda")[None, :] >= lengths[:, None]) * -10000.0
        att = (att + mask[:, None, None, :]).softmax(-1)
        o_ref = (att @ t[2]).transpose(1, 2).reshape(S * N, D)
        d_o = ops.round_tf32(torch.randn(S * N, D, device="cuda"))
        o_ref.backward(d_o)
        o, lse = ops.attention_fwd(qkv, S, N, H, lengths)
        dqkv = ops.attention_bwd(qkv, o, d_o, lse, S, N, H, lengths)
        torch.cuda.synchronize()
        g = q.grad
        print("attn S=%d N=%3d H=%2d ragged=%d fwd=%.2e dq=%.2e dk=%.2e dv=%.2e" %
              (S, N, H, ragged, rel_err(o, o_ref), rel_err(dqkv[:, :D], g[:, :D]), rel_err(dqkv[:, D:2 * D], g[:, D:2 * D]),
               rel_err(dqkv[:, 2 * D:], g[:, 2 * D:])))
    S, N, H = 512, 251, 12
    qkv = torch.randn(S * N, 3 * H * 64, device="cuda")
    o, lse = ops.attention_fwd(qkv, S, N, H)
    d_o = torch.randn_like(o)
    dqkv = torch.empty_like(qkv)
    torch.cuda.synchronize()
    e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
    e[0].record()
    ops.attention_fwd(qkv, S, N, H, out=o, lse=lse)
    e[1].record()
    ops.attention_bwd(qkv, o, d_o, lse, S, N, H, dqkv=dqkv)
    e[2].record()
    torch.cuda.synchronize()
    fl = 4.0 * S * H * N * N * 64
    print("attn c2 fwd %.3f ms (%.1f TF/s)  bwd %.3f ms (%.1f TF/s)" %
          (e[0].elapsed_time(e[1]), fl / e[0].elapsed_time(e[1]) / 1e9, e[1].elapsed_time(e[2]),
           2.5 * fl / e[1].elapsed_time(e[2]) / 1e9))


def sec_bn():
    import torch
    from audiossl_b200 import ops
    torch.manual_seed(0)
    R, C = 300, 4096
    x = torch.randn(R, C, device="cuda") * 3 + 1
    g = torch.rand(C, device="cuda") + 0.5
    b = torch.randn(C, device="cuda") * 0.3
    bn = torch.nn.BatchNorm1d(C).cuda()
    with torch.no_grad():
        bn.weight.copy_(g)
        bn.bias.copy_(b)
    xr = x.clone().requires_grad_(True)
    y_ref = torch.relu(bn(xr))
    dy = torch.randn(R, C, device="cuda")
    y_ref.backward(dy)
    rm, rv = torch.zeros(C, device="cuda"), torch.ones(C, device="cuda")
    mean, m2 = ops.bn_stats(x)
    rstd = ops.bn_finalize(mean, m2, R, rm, rv)
    y = ops.bn_relu_fwd(x, mean, rstd, g, b)
    s1, s2 = ops.bn_relu_bwd_stats(dy, x, mean, rstd, g, b)
    dx = ops.bn_relu_bwd_apply(dy, x, mean, rstd, g, b, s1, s2, R)
    print("bn fwd=%.2e dx=%.2e dgamma=%.2e dbeta=%.2e rm=%.2e rv=%.2e" %
          (rel_err(y, y_ref), rel_err(dx, xr.grad), rel_err(s2, bn.weight.grad), rel_err(s1, bn.bias.grad),
           rel_err(rm, bn.running_mean), rel_err(rv, bn.running_var)))
    X = torch.randn(5000, 768, device="cuda")
    out = torch.zeros(768, device="cuda")
    ops.colsum_acc(X, out)
    print("colsum rel=%.2e" % rel_err(out, X.sum(0)))


def sec_loss():
    import torch
    from audiossl_b200 import ops
    from oracle import atst_oracle as O
    torch.manual_seed(0)
    for ncrops, B in [(2, 8), (4, 5), (8, 16), (2, 256)]:
        s = torch.randn(ncrops * B, 256, device="cuda")
        t = torch.randn(2 * B, 256, device="cuda")
        sr = s.cpu().clone().requires_grad_(True)
        loss, std_s, std_t = O.byol_loss(sr, t.cpu(), ncrops)
        loss.backward()
        ds, acc = ops.byol_loss(s, t, ncrops, B)
        out = ops.byol_finalize(acc, ncrops * B, 2 * B, ncrops, B).cpu()
        print("loss ncrops=%d B=%3d loss %.6f/%.6f std_s %.6f/%.6f std_t %.6f/%.6f dgrad rel=%.2e" %
              (ncrops, B, out[0], loss.item(), out[1], std_s.item(), out[2], std_t.item(), rel_err(ds.cpu(), sr.grad)))


def sec_optim():
    import torch
    from audiossl_b200 import ops
    from oracle import atst_oracle as O
    torch.manual_seed(0)
    n = 100000
    p, g = torch.randn(n, device="cuda"), torch.randn(n, device="cuda")
    m, v = torch.zeros(n, device="cuda"), torch.zeros(n, device="cuda")
    pc, mc, vc = p.cpu().clone(), m.cpu().clone(), v.cpu().clone()
    for step in (1, 2, 3):
        ops.adamw_step(p, g, m, v, step, 1e-3, 0.04)
        O.hf_adamw_step(pc, g.cpu(), mc, vc, step, 1e-3, 0.04)
    print("adamw rel=%.2e" % rel_err(p.cpu(), pc))
    k, q = torch.randn(n, device="cuda"), torch.randn(n, device="cuda")
    kr = k * 0.99 + (1 - 0.99) * q
    ops.ema_update(k, q, 0.99)
    print("ema rel=%.2e" % rel_err(k, kr))


def sec_tokens():
    import torch
    from audiossl_b200 import ops
    from oracle import atst_oracle as O
    torch.manual_seed(0)
    S, T, D = 3, 101, 128
    P = T // 4
    mel = torch.randn(S, 1, 64, T, device="cuda")
    ast = O.OracleAST(D, 1, 2)
    ref = ast.patchify(mel.cpu())
    out = ops.patchify(mel).reshape(S, P, 256)
    print("patchify rel=%.2e" % rel_err(out.cpu(), ref))
    pe = torch.randn(S * P, D, device="cuda")
    cls, pos = torch.randn(D, device="cuda"), torch.randn(251, D, device="cuda")
    x = ops.tokens_fwd(pe, cls, pos, S, P, D).reshape(S, P + 1, D)
    xr = torch.cat([cls.expand(S, 1, D), pe.reshape(S, P, D)], 1) + pos[None, :P + 1]
    print("tokens fwd rel=%.2e" % rel_err(x, xr))
    dx = torch.randn(S * (P + 1), D, device="cuda")
    dpe, dpos, dcls = torch.empty(S * P, D, device="cuda"), torch.zeros(251, D, device="cuda"), torch.zeros(D, device="cuda")
    ops.tokens_bwd(dx, dpe, dpos, dcls, S, P, D)
    d3 = dx.reshape(S, P + 1, D)
    print("tokens bwd dpe=%.2e dpos=%.2e dcls=%.2e" % (rel_err(dpe.reshape(S, P, D), d3[:, 1:]), rel_err(dpos[:P + 1], d3.sum(0)),
                                                     rel_err(dcls, d3[:, 0].sum(0))))


def sec_gemm_perf():
    """isolated throughput of the three GEMM flavours at BASELINE config-2 sizes"""
    import torch
    from audiossl_b200 import ops
    M = 128512

    def tm(fn, flops, name):
        for _ in range(2):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(4):
            fn()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 4
        print("%-34s %8.3f ms  %7.1f TFLOP/s" % (name, ms, flops / ms / 1e9), flush=True)

    from audiossl_b200 import _lib
    A = torch.randn(M, 768, device="cuda")
    W = torch.randn(2304, 768, device="cuda")
    C = torch.empty(M, 2304, device="cuda")
    for epi, nm in ((0, "store"), (8, "dbg no-store"), (9, "dbg no-tmem-load")):
        tm(lambda: ops.gemm_nt(A, W, epi=epi, out=C), 2.0 * M * 2304 * 768, "nt M x 2304 x 768 epi=%s" % nm)
    u = torch.randn(M, 3072, device="cuda")
    gg = torch.empty_like(u)
    tm(lambda: ops.gelu_fwd(u, gg), 1.0, "gelu_fwd elementwise [M,3072]")
    cs = torch.zeros(3072, device="cuda")
    tm(lambda: ops.gelu_bwd_(gg, u, colsum_out=cs), 1.0, "gelu_bwd+colsum elementwise [M,3072]")
    del A, W, C, u, gg
    for pf in (0,):
        _lib.lib().atst_set_option(b"gemm_l2_prefetch", pf)
        A = torch.randn(M, 768, device="cuda")
        W = torch.randn(2304, 768, device="cuda")
        C = torch.empty(M, 2304, device="cuda")
        tm(lambda: ops.gemm_nt(A, W, out=C), 2.0 * M * 2304 * 768, "l2_prefetch=%d nt M x 2304 x 768" % pf)
        dW = torch.zeros(2304, 768, device="cuda")
        tm(lambda: ops.gemm_tn_acc(C, A, dW), 2.0 * M * 2304 * 768, "l2_prefetch=%d tn 2304 x 768" % pf)
        del A, W, C, dW
    for (N, K) in [(2304, 768), (768, 768), (3072, 768), (768, 3072)]:
        A = torch.randn(M, K, device="cuda")
        W = torch.randn(N, K, device="cuda")
        C = torch.empty(M, N, device="cuda")
        tm(lambda: ops.gemm_nt(A, W, out=C), 2.0 * M * N * K, "nt  M x %d x %d" % (N, K))
        bias = torch.randn(N, device="cuda")
        aux = torch.empty(M, N, device="cuda")
        if N == 3072:
            tm(lambda: ops.gemm_nt(A, W, bias=bias, epi=ops.EPI_GELU, aux=aux, round_out=True, out=C), 2.0 * M * N * K,
               "nt+gelu M x %d x %d" % (N, K))
        if N == 768:
            tm(lambda: ops.gemm_nt(A, W, bias=bias, epi=ops.EPI_RESID, resid=aux, out=C), 2.0 * M * N * K,
               "nt+resid M x %d x %d" % (N, K))
        # dgrad: dX[M,K] = dY[M,N] @ W[N,K]
        dX = torch.empty(M, K, device="cuda")
        tm(lambda: ops.gemm_nn(C, W, out=dX), 2.0 * M * N * K, "nn  M x %d (K=%d)" % (K, N))
        if K == 3072:
            tm(lambda: ops.gemm_nn(C, W, epi=ops.EPI_DGELU, aux=A, round_out=True, out=dX), 2.0 * M * N * K,
               "nn+dgelu M x %d (K=%d)" % (K, N))
        # wgrad: dW[N,K] += dY[M,N]^T @ X[M,K]
        dW = torch.zeros(N, K, device="cuda")
        tm(lambda: ops.gemm_tn_acc(C, A, dW), 2.0 * M * N * K, "tn  %d x %d (T=M)" % (N, K))
        del A, W, C, aux, dX, dW


def sec_heads():
    """the exact head-shaped GEMMs (few rows, 4096 / 256 features)"""
    import torch
    from audiossl_b200 import ops
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.manual_seed(0)
    for R in (4, 6, 32, 64, 128, 512):
        dz1 = ops.round_tf32(torch.randn(R, 4096, device="cuda") * 0.1)
        x = ops.round_tf32(torch.randn(R, 256, device="cuda"))
        C = torch.zeros(4096, 256, device="cuda")
        ops.gemm_tn_acc(dz1, x, C)
        W0 = ops.round_tf32(torch.randn(4096, 256, device="cuda") * 0.05)
        dx = ops.gemm_nn(dz1, W0)
        dz2 = ops.round_tf32(torch.randn(R, 256, device="cuda") * 0.1)
        a1 = ops.round_tf32(torch.randn(R, 4096, device="cuda"))
        C2 = torch.zeros(256, 4096, device="cuda")
        ops.gemm_tn_acc(dz2, a1, C2)
        W3 = ops.round_tf32(torch.randn(256, 4096, device="cuda") * 0.05)
        da1 = ops.gemm_nn(dz2, W3)
        torch.cuda.synchronize()
        print("R=%3d tn[4096x256] %.2e nn[->256] %.2e tn[256x4096] %.2e nn[->4096] %.2e" %
              (R, rel_err(C, dz1.t() @ x), rel_err(dx, dz1 @ W0), rel_err(C2, dz2.t() @ a1), rel_err(da1, dz2 @ W3)))


def sec_pair():
    """CTA-pair (cta_group::2) GEMM kernel: correctness of the three flavours + epilogues, then throughput"""
    import torch
    from audiossl_b200 import _lib, ops
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.manual_seed(0)
    L = _lib.lib()
    L.atst_set_option(b"gemm_cta_pair", 1)
    try:
        for (M, N, K) in [(256, 256, 32), (256, 256, 128), (300, 512, 256), (1004, 2304, 768), (4096, 768, 3072),
                          (70, 4096, 128), (128512, 768, 768)]:
            A = ops.round_tf32(torch.randn(M, K, device="cuda"))
            B = ops.round_tf32(torch.randn(N, K, device="cuda") * 0.05)
            out = ops.gemm_nt(A, B)
            torch.cuda.synchronize()
            print("pair nt %6dx%5dx%5d rel=%.3e" % (M, N, K, rel_err(out, A @ B.t())), flush=True)
        M, N, K = 520, 768, 256
        A = ops.round_tf32(torch.randn(M, K, device="cuda"))
        B = ops.round_tf32(torch.randn(N, K, device="cuda") * 0.05)
        bias, resid = torch.randn(N, device="cuda"), torch.randn(M, N, device="cuda")
        scale = torch.rand(M // 26, device="cuda") + 0.5
        base = A @ B.t() + bias
        out = ops.gemm_nt(A, B, bias=bias, epi=ops.EPI_RESID, resid=resid, rowscale=scale, rows_per_seq=26)
        print("pair epi resid rel=%.3e" % rel_err(out, resid + scale.repeat_interleave(26)[:, None] * base))
        aux = torch.empty(M, N, device="cuda")
        out = ops.gemm_nt(A, B, bias=bias, epi=ops.EPI_GELU, aux=aux)
        print("pair epi gelu  rel=%.3e aux=%.3e" % (rel_err(out, torch.nn.functional.gelu(base)), rel_err(aux, base)))
        for (T, Mf, Nf) in [(96, 256, 256), (1000, 256, 512), (4000, 768, 2304), (130, 4096, 256), (128512, 768, 768)]:
            A = ops.round_tf32(torch.randn(T, Mf, device="cuda") * 0.1)
            B = ops.round_tf32(torch.randn(T, Nf, device="cuda") * 0.1)
            C = torch.zeros(Mf, Nf, device="cuda")
            ops.gemm_tn_acc(A, B, C)
            W = ops.round_tf32(torch.randn(Mf, Nf, device="cuda") * 0.1)
            dX = ops.gemm_nn(A, W)
            torch.cuda.synchronize()
            print("pair tn T=%6d %4dx%4d rel=%.3e | nn rel=%.3e" % (T, Mf, Nf, rel_err(C, A.t() @ B), rel_err(dX, A @ W)),
                  flush=True)
        M = 128512

        def tm(fn, flops, name):
            for _ in range(2):
                fn()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(4):
                fn()
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / 4
            print("%-34s %8.3f ms  %7.1f TFLOP/s" % (name, ms, flops / ms / 1e9), flush=True)
        for pair in (0, 1):
            L.atst_set_option(b"gemm_cta_pair", pair)
            for (N, K) in [(2304, 768), (768, 768), (3072, 768), (768, 3072)]:
                A = torch.randn(M, K, device="cuda")
                W = torch.randn(N, K, device="cuda")
                C = torch.empty(M, N, device="cuda")
                tm(lambda: ops.gemm_nt(A, W, out=C), 2.0 * M * N * K, "pair=%d nt M x %d x %d" % (pair, N, K))
                dX = torch.empty(M, K, device="cuda")
                tm(lambda: ops.gemm_nn(C, W, out=dX), 2.0 * M * N * K, "pair=%d nn M x %d (K=%d)" % (pair, K, N))
                dW = torch.zeros(N, K, device="cuda")
                tm(lambda: ops.gemm_tn_acc(C, A, dW), 2.0 * M * N * K, "pair=%d tn %d x %d" % (pair, N, K))
                del A, W, C, dX, dW
    finally:
        L.atst_set_option(b"gemm_cta_pair", 0)


SECTIONS = {"gemm_trace": sec_gemm_trace, "copy_pattern": sec_copy_pattern, "ew_perf": sec_ew_perf, "attn_trace": sec_attn_trace, "umma_probe": sec_umma_probe, "attn_tc": sec_attn_tc, "pair": sec_pair, "heads": sec_heads, "gemm_perf": sec_gemm_perf, "mel": sec_mel, "gemm_nt": sec_gemm_nt, "gemm_mn": sec_gemm_mn, "ln": sec_ln, "attn": sec_attn,
            "bn": sec_bn, "loss": sec_loss, "optim": sec_optim, "tokens": sec_tokens}

if __name__ == "__main__":
    args = sys.argv[1:] or ["all"]
    if args == ["all"]:
        for name in SECTIONS:
            print("=" * 20, name, flush=True)
            try:
                r = subprocess.run([sys.executable, os.path.abspath(__file__), name], timeout=240)
                print("exit", r.returncode, flush=True)
            except subprocess.TimeoutExpired:
                print("TIMEOUT in section", name, flush=True)
    else:
        for a in args:
            SECTIONS[a]()
