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
# variants of the same sources (audiossl_b200/build.py): "" = the product (TF32 tcgen05 path), "precise" = 3xTF32
# validation build (fp32-equivalent products, audiossl_b200.set_precision), "debug" = bring-up entry points
VARIANT_FILES = {"": "libatst_b200.so", "precise": "libatst_b200_precise.so", "debug": "libatst_b200_debug.so"}

P, I, L, F, U = c_void_p, c_int, c_longlong, c_float, c_uint

# name -> argtypes (restype is int unless noted)
SIGNATURES = {
    "atst_version": [],
    "atst_init": [],
    "atst_set_option": [c_char_p, I],
    "atst_mel_forward": [P, I, I, L, P, I, P, L, P, I, P],
    "atst_gemm_nt": [P, I, P, I, P, I, I, I, I, P, I, P, I, P, I, P, I, I, P],
    "atst_gemm_nn": [P, I, P, I, P, I, I, I, I, I, P, I, P, I, I, P, P],
    "atst_gemm_tn": [P, I, P, I, P, I, I, I, I, P],
    "atst_layernorm_forward": [P, L, P, P, P, L, P, P, I, I, F, I, P],
    "atst_layernorm_backward": [P, L, P, L, P, P, P, P, L, P, L, P, P, I, I, P, L, P, I, P, P],
    "atst_attention_forward": [P, P, P, P, I, I, I, P],
    "atst_attention_backward": [P, P, P, P, P, P, P, I, I, I, P],
    "atst_patchify": [P, L, I, I, P, P],
    "atst_tokens_forward": [P, P, P, P, P, P, I, I, I, I, P],
    "atst_tokens_backward": [P, P, P, P, P, P, I, I, I, I, P],
    "atst_colsum_accumulate": [P, L, I, I, P, P],
    "atst_bn_stats": [P, I, I, P, P, P],
    "atst_bn_finalize": [P, P, F, F, F, P, P, P, I, P],
    "atst_bn_relu_forward": [P, P, P, P, P, P, I, I, I, P],
    "atst_bn_relu_backward_stats": [P, P, P, P, P, P, I, I, P, P, P],
    "atst_bn_relu_backward_apply": [P, P, P, P, P, P, P, P, F, P, I, I, I, P],
    "atst_byol_loss": [P, P, I, I, P, P, P],
    "atst_byol_finalize": [P, F, F, I, I, P, P],
    "atst_ema_update": [P, P, F, P, L, P],
    "atst_adamw_step": [P, P, P, P, L, I, F, F, F, F, F, F, P, P],
    "atst_mixup_forward": [P, I, P, I, P, P, P, P, P, I, I, P],
    "atst_resize_crop_forward": [P, P, P, I, I, I, I, I, P],
    "atst_gather_rows": [P, P, P, I, I, P],
    "atst_scatter_rows": [P, P, P, I, I, P],
    "atst_gelu_forward": [P, P, L, P],
    "atst_gelu_backward": [P, P, I, I, P, P],
    "atst_round_tf32": [P, P, L, P],
    "atst_is_precise": [],
    "atst_split_tf32": [P, L, I, I, P, I, I, P],
    "atst_axpy": [P, P, F, L, P],
}

# bring-up entry points (include/atst_b200_debug.h), present in the "debug" variant only
DEBUG_SIGNATURES = {
    "atst_gemm_mn_debug": [I, P, I, P, I, P, I, I, I, I, U, U, U, U, I, I, P],
    "atst_umma_probe": [I, P, P, P, U, U, U, U, P],
    "atst_gemm_trace": [P],
    "atst_copy_pattern": [P, P, I, I, I, P],
    "atst_attention_trace": [P, I, I],
}

_libs = {}
_inited = set()
_variant = os.environ.get("ATST_LIB_VARIANT", "")


def set_variant(name):
    """select the library the compute calls go to ("" product, "precise", "debug"); returns the previous one"""
    global _variant
    if name not in VARIANT_FILES:
        raise ValueError("unknown library variant %r" % (name,))
    prev, _variant = _variant, name
    return prev


def variant():
    return _variant


def load(name=None):
    """dlopen a library variant and declare signatures (no GPU needed)."""
    name = _variant if name is None else name
    if name in _libs:
        return _libs[name]
    path = os.path.join(_HERE, VARIANT_FILES[name])
    if not os.path.exists(path):
        raise RuntimeError("%s not found: run `python -m audiossl_b200.build` (nvcc, sm_100a). "
                           "There is no CPU fallback." % path)
    lib = ctypes.CDLL(path)
    sigs = dict(SIGNATURES)
    if name == "debug":
        sigs.update(DEBUG_SIGNATURES)
    for fn_name, args in sigs.items():
        fn = getattr(lib, fn_name)
        fn.argtypes = args
        fn.restype = c_int
    lib.atst_last_error.argtypes = []
    lib.atst_last_error.restype = c_char_p
    _libs[name] = lib
    return lib


def lib():
    """library handle for compute calls: checks the device once per variant."""
    l = _libs.get(_variant)
    if l is not None and _variant in _inited:
        return l
    l = load()
    if _variant not in _inited:
        rc = l.atst_init()
        if rc != 0:
            raise RuntimeError("atst_init failed (%d): %s" % (rc, l.atst_last_error().decode()))
        _inited.add(_variant)
    return l


def is_precise():
    return _variant == "precise"


def check(rc, what=""):
    if rc != 0:
        raise RuntimeError("%s failed (%d): %s" % (what or "atst call", rc, load().atst_last_error().decode()))


def ptr(t):
    """device pointer of a torch tensor (None -> NULL)."""
    return None if t is None else t.data_ptr()


_raw_stream = None


def stream():
    """the caller's current CUDA stream as a raw cudaStream_t (what every entry point takes last).  Goes through
    torch's C accessors: torch.cuda.current_stream() builds a Stream object per call (~5 us), which at ~500 launches
    per step was half of the host time of a small-batch step."""
    global _raw_stream
    if _raw_stream is None:
        import torch
        get_raw, get_dev = torch._C._cuda_getCurrentRawStream, torch._C._cuda_getDevice
        _raw_stream = lambda: get_raw(get_dev())  # noqa: E731
    return _raw_stream()
