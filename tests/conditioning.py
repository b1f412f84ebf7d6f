"""How ill-conditioned is the gradient being compared?  (test infrastructure; uses the oracle)

The BYOL loss behind train-mode BatchNorm heads is a difference of nearly equal, batch-centred quantities, and the
per-sequence loss gradients sum to zero over the batch, so weight gradients are sums over tokens that largely
cancel: an incoherent relative perturbation eps of the forward arithmetic (rounding, summation order - anything that
differs element by element) moves the gradients by kappa * eps with kappa ~ 1e2 on the fixtures and on small batches
(rounding the GEMM operands to TF32, 5e-4, moves them by 4 - 11 %; tools/tf32_sensitivity.py).  No two
implementations that differ at all in the forward pass - TF32 vs fp32, 3xTF32's 1e-5 vs fp32's 1e-7, one rounding
flip vs another - can agree on such a gradient better than kappa times their forward distance, so a fixed gradient
tolerance is either violated by correct code on an ill-conditioned case or blind on a well-conditioned one.

The parity tests therefore MEASURE kappa with the oracle: the same computation is run twice, the second time with
element-wise relative Gaussian noise on the output (and on the incoming gradient) of every Linear - what an
implementation's own arithmetic error looks like - and the relative change of each gradient is divided by the
relative change of the outputs.  A gradient may then be off by  max(fixed tolerance, 3 * kappa * observed output
error).  On well-conditioned tensors (kappa ~ 1: random data, large batches) the fixed tolerance rules; where the
problem itself amplifies, the bound says exactly by how much."""
import contextlib
import copy

import torch

from oracle import atst_oracle as O


def _rel(a, b, floor=0.0):
    a, b = a.detach().double(), b.detach().double()
    return ((a - b).norm() / max(b.norm().item(), floor, 1e-30)).item()


@contextlib.contextmanager
def gemm_noise(noise, seed=0):
    """every oracle Linear (forward result and the gradient flowing back into it) times (1 + noise * N(0, 1))."""
    gen = torch.Generator().manual_seed(seed)
    orig = O.linear

    def jitter(t):
        return t * (1.0 + noise * torch.randn(t.shape, generator=gen, dtype=t.dtype))

    def noisy_linear(x, w, b=None):
        y = jitter(orig(x, w, b))
        if y.requires_grad:
            y.register_hook(jitter)
        return y
    O.linear = noisy_linear
    try:
        yield
    finally:
        O.linear = orig


def gradient_kappa(module, run, noise=2e-5, seed=0):
    """module: an oracle nn.Module; run(m) -> (outputs tensor, {name: grad tensor}) on a private copy m.
    Returns ({name: kappa}, output change): kappa = (relative gradient change, measured like the parity tests measure
    a gradient error: against max(|g|, 1e-3 * largest gradient norm)) / (relative output change)."""
    with O.tf32_emulation(False):  # a sensitivity is a derivative: measured in plain fp32, without rounding steps
        out0, g0 = run(copy.deepcopy(module))
        with gemm_noise(noise, seed):
            out1, g1 = run(copy.deepcopy(module))
    d_out = max(_rel(out1, out0), 1e-12)
    big = max(v.norm().item() for v in g0.values())
    return {n: _rel(g1[n], g0[n], 1e-3 * big) / d_out for n in g0}, d_out


def allowed(fixed_tol, kappa, observed_forward_error, safety=3.0):
    return max(fixed_tol, safety * kappa * observed_forward_error)
