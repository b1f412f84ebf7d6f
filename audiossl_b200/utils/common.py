"""Per-step schedule tables and the weight-decay grouping of the ATST recipes.

API mirror of the three helpers the hot path uses from audiossl/utils/common.py (``cosine_scheduler_step`` :29-39,
``get_params_groups`` :41-68, ``bool_flag`` :69-80): same names, arguments and results - the tables are compared
bit for bit with the reference's in tests/golden/sched.npz.  The grouping must agree with
``audiossl_b200.params.FlatParams`` (which lays the same two groups out as contiguous segments for the fused
optimizer)."""
import argparse

import numpy as np


def cosine_scheduler_step(base_value, final_value, max_steps, warmup_steps=0, start_warmup_value=0):
    """table[step]: linear warm-up to ``base_value`` over ``warmup_steps`` entries, then half a cosine period down
    (or up) to ``final_value`` over the remaining ``max_steps - warmup_steps`` entries."""
    n_cos = max_steps - warmup_steps
    phase = np.pi * np.arange(n_cos) / n_cos
    cosine = final_value + 0.5 * (base_value - final_value) * (1 + np.cos(phase))
    warmup = np.linspace(start_warmup_value, base_value, warmup_steps) if warmup_steps > 0 else np.array([])
    table = np.concatenate((warmup, cosine))
    assert len(table) == max_steps
    return table


def _decayed(name, param):
    return not name.endswith(".bias") and param.ndim != 1


def get_params_groups(model, no_weight_decay_attr=(), debug=False):
    """[{'params': matrices / embeddings}, {'params': biases and 1-D tensors, 'weight_decay': 0.}] over the trainable
    parameters in module order; ``debug=True`` returns the two name lists instead."""
    if len(no_weight_decay_attr):
        raise NotImplementedError("no_weight_decay_attr is not used by the ATST recipes (flat-buffer segments are "
                                  "split by name/shape only)")
    trainable = [(n, p) for n, p in model.named_parameters() if p.requires_grad]
    decay = [(n, p) for n, p in trainable if _decayed(n, p)]
    no_decay = [(n, p) for n, p in trainable if not _decayed(n, p)]
    if debug:
        return [n for n, _ in decay], [n for n, _ in no_decay]
    return [{"params": [p for _, p in decay]}, {"params": [p for _, p in no_decay], "weight_decay": 0.}]


def bool_flag(s):
    """argparse type for on/off switches ("on", "true", "1" / "off", "false", "0", any case)."""
    word = s.lower()
    if word in ("on", "true", "1"):
        return True
    if word in ("off", "false", "0"):
        return False
    raise argparse.ArgumentTypeError("invalid value for a boolean flag")
