"""ctypes binding of libatst_b200.so (the C ABI in include/atst_b200.h).

The product path has no CPU fallback: importing works anywhere (so the symbol table can be
checked on a CPU box) but every compute call requires the sm_100a library and a CUDA device and
raises RuntimeError otherwise.
"""
import ctypes
import os
from ctypes import c_char_p, c_float, c_int, c_longlong, c_uint, c_void_p

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libatst_b200.so")

P, I, L, F, U = c_void_p, c_int, c_longlong, c_float, c_uint

# name -> argtypes (restype is int unless noted)
SIGNATURES = {
    "atst_version": [],
    "atst_init": [],
    "atst_set_option": [c_char_p, I],
    "atst_mel_forward": [P, I, I, L, I, P, L, P, I, P],
    "atst_gemm_nt": [P, I, P, I, P, I, I, I, I, P, I, P, I, P, I, P, I, I, P],
    "atst_gemm_nn": [P, I, P, I, P, I, I, I, I, I, P, I, P, I, I, P, P],
    "atst_gemm_tn": [P, I, P, I, P, I, I, I, I, P],
    "atst_gemm_mn_debug": [I, P, I, P, I, P, I, I, I, I, U, U, U, U, I, I, P],
    "atst_umma_probe": [I, P, P, P, U, U, U, U, P],
    "atst_layernorm_forward": [P, L, P, P, P, L, P, P, I, I, F, I, P],
    "atst_layernorm_backward": [P, L, P, L, P, P, P, P, L, P, L, P, P, I, I, P, L, P, I, P, P],
    "atst_attention_forward": [P, P, P, P, I, I, I, P],
    "atst_attention_backward": [P, P, P, P, P, P, P, I, I, I, P],
    "atst_gemm_trace": [P],
    "atst_copy_pattern": [P, P, I, I, I, P],
    "atst_attention_trace": [P, I, I],
    "atst_patchify": [P, L, I, I, P, P],
    "atst_tokens_forward": [P, P, P, P, P, P, I, I, I, I, P],
    "atst_tokens_backward": [P, P, P, P, P, P, I, I, I, I, P],
    "atst_colsum_accumulate": [P, L, I, I, P, P],
    "atst_bn_stats": [P, I, I, P, P, P],
    "atst_bn_finalize": [P, P, F, F, F, P, P, P, I, P],
    "atst_bn_relu_forward": [P, P, P, P, P, P, I, I, P],
    "atst_bn_relu_backward_stats": [P, P, P, P, P, P, I, I, P, P, P],
    "atst_bn_relu_backward_apply": [P, P, P, P, P, P, P, P, F, P, I, I, P],
    "atst_byol_loss": [P, P, I, I, P, P, P],
    "atst_byol_finalize": [P, F, F, I, I, P, P],
    "atst_ema_update": [P, P, F, L, P],
    "atst_adamw_step": [P, P, P, P, L, I, F, F, F, F, F, F, P],
    "atst_mixup_forward": [P, P, P, P, P, L, I, P],
    "atst_resize_crop_forward": [P, P, P, I, I, I, I, I, P],
    "atst_gather_rows": [P, P, P, I, I, P],
    "atst_scatter_rows": [P, P, P, I, I, P],
    "atst_gelu_forward": [P, P, L, P],
    "atst_gelu_backward": [P, P, I, I, P, P],
    "atst_round_tf32": [P, P, L, P],
    "atst_axpy": [P, P, F, L, P],
}

_lib = None
_inited = False


def load():
    """dlopen the library and declare signatures (no GPU needed)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError("%s not found: run `python -m audiossl_b200.build` (nvcc, sm_100a). "
                           "There is no CPU fallback." % LIB_PATH)
    lib = ctypes.CDLL(LIB_PATH)
    for name, args in SIGNATURES.items():
        fn = getattr(lib, name)
        fn.argtypes = args
        fn.restype = c_int
    lib.atst_last_error.argtypes = []
    lib.atst_last_error.restype = c_char_p
    _lib = lib
    return lib


def lib():
    """library handle for compute calls: checks the device once."""
    global _inited
    l = load()
    if not _inited:
        rc = l.atst_init()
        if rc != 0:
            raise RuntimeError("atst_init failed (%d): %s" % (rc, l.atst_last_error().decode()))
        _inited = True
    return l


def check(rc, what=""):
    if rc != 0:
        raise RuntimeError("%s failed (%d): %s" % (what or "atst call", rc, load().atst_last_error().decode()))


def ptr(t):
    """device pointer of a torch tensor (None -> NULL)."""
    return None if t is None else t.data_ptr()


def stream():
    import torch
    return torch.cuda.current_stream().cuda_stream
