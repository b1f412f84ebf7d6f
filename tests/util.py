"""helpers shared by the parity tests (oracle side only: never imported by the product)."""
import os

import numpy as np
import torch

from tests.golden import detfill

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def gold(name):
    return np.load(os.path.join(GOLD, name), allow_pickle=False)


def load_det(module):
    sd = module.state_dict()
    new = detfill.fill_state_dict(sd)
    module.load_state_dict({k: (torch.from_numpy(new[k]) if k in new else v) for k, v in sd.items()})


def make_inputs(tag, B, widths, lens):
    crops, lengths = [], []
    for i, (w, l) in enumerate(zip(widths, lens)):
        crops.append(torch.from_numpy(detfill.det_array("%s/crop%d" % (tag, i), (B, 1, 64, w), 1.0, "uniform")))
        lengths.append(torch.tensor(l, dtype=torch.int64))
    return crops, lengths


def summary_rms(g, prefix, size):
    return float(np.sqrt(g[prefix + "/sq"] / max(size, 1)))


def check_summary(arr, g, prefix, rtol, atol, scale=None):
    """compare an array with a detfill.summarize() record stored under prefix/* in npz g.
    atol is relative to `scale` (default: the rms of the golden array itself)."""
    a = np.asarray(arr, np.float32).ravel()
    idx = g[prefix + "/idx"]
    if scale is None:
        scale = max(summary_rms(g, prefix, a.size), 1e-12)
    np.testing.assert_allclose(a[: g[prefix + "/head"].size], g[prefix + "/head"], rtol=rtol, atol=atol * scale,
                               err_msg=prefix + " head")
    np.testing.assert_allclose(a[idx], g[prefix + "/samp"], rtol=rtol, atol=atol * scale, err_msg=prefix + " samp")
    np.testing.assert_allclose(np.abs(a).astype(np.float64).sum(), g[prefix + "/abs"], rtol=max(rtol, 1e-4),
                               atol=atol * scale * a.size, err_msg=prefix + " abs")


def sample_rel_err(arr, g, prefix):
    """error on the stored sample (head + strided sample) of a summarised array, relative to the larger of the sample's
    own rms and the rms of the WHOLE golden array (a sample may consist of small entries, or - pos_embed - the
    array may be mostly structural zeros); also returns the golden array's l2 norm."""
    a = np.asarray(arr, np.float64).ravel()
    mine = np.concatenate([a[: g[prefix + "/head"].size], a[g[prefix + "/idx"]]])
    ref = np.concatenate([g[prefix + "/head"], g[prefix + "/samp"]]).astype(np.float64)
    rms = max(np.sqrt(float(g[prefix + "/sq"]) / max(a.size, 1)), np.sqrt(np.mean(ref ** 2)))
    err = np.sqrt(np.mean((mine - ref) ** 2)) / max(rms, 1e-30)  # every stored sample counts (no trimming)
    return float(err), float(np.sqrt(float(g[prefix + "/sq"])))


CASES = {
    "tiny2": dict(dim=128, depth=2, heads=2, ncrops=2, B=3, widths=[101, 101],
                  lens=[[101, 77, 50], [101, 101, 9]]),
    "tiny4": dict(dim=128, depth=2, heads=2, ncrops=4, B=2, widths=[101, 101, 41, 41],
                  lens=[[101, 90], [101, 101], [41, 33], [41, 41]]),
    "tiny2dp": dict(dim=128, depth=2, heads=2, ncrops=2, B=4, widths=[101, 101],
                    lens=[[101, 101, 60, 101], [101, 80, 101, 101]], drop_path=0.5),
    "tiny2b32": dict(dim=128, depth=2, heads=2, ncrops=2, B=32, widths=[101, 101],
                     lens=[[101 - (i * 7) % 60 for i in range(32)], [101 - (i * 11) % 45 for i in range(32)]]),
    "small2": dict(dim=384, depth=12, heads=6, ncrops=2, B=2, widths=[101, 101], lens=[[101, 64], [101, 101]]),
}


def dp_scales_from_rand(rand, depth, keep_per_block, n_groups_teacher=1, n_groups_student=1):
    """Rebuild per-block (attn, mlp) DropPath scales from the recorded torch.rand stream.

    Reference order of torch.rand calls (modules/transformer.py:48-57, 136-150): teacher encoder
    call(s) first, then student; inside an encoder call block 0 attn, block 0 mlp, block 1 attn, ...
    Blocks whose drop prob is 0 use nn.Identity and draw nothing.
    scale = floor(keep + u) / keep.
    """
    rows = list(rand)
    out = []
    for _ in range(n_groups_teacher + n_groups_student):
        blocks = []
        for i in range(depth):
            keep = keep_per_block[i]
            if keep >= 1.0:
                blocks.append(None)
                continue
            a = torch.from_numpy(np.floor(keep + rows.pop(0)) / keep).float()
            m = torch.from_numpy(np.floor(keep + rows.pop(0)) / keep).float()
            blocks.append((a, m))
        out.append(blocks)
    assert not rows
    return out[:n_groups_teacher], out[n_groups_teacher:]
