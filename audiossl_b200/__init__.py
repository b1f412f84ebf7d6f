"""audiossl_b200 - B200-native (sm_100a) implementation of the ATST pre-training hot path of
Audio-WestlakeU/audiossl, behind the reference's own Python surface.

    audiossl_b200.transforms                 <-> audiossl.transforms
    audiossl_b200.models.atst.ATST           <-> audiossl.models.atst.ATST
    audiossl_b200.methods.atst.model         <-> audiossl.methods.atst.model (ATSTLightningModule)
    audiossl_b200.methods.atst.transform     <-> audiossl.methods.atst.transform (ATSTTrainTransform)

Compute goes through libatst_b200.so (include/atst_b200.h); there is no CPU or eager fallback.
"""
__version__ = "0.1.0"


def set_precision(mode):
    """"tf32" (default): tcgen05 kind::tf32 products, producers round operands with cvt.rna - the benchmarked path.
    "3xtf32": the validation build of the same kernels and engine (libatst_b200_precise.so): producers keep fp32 and
    every product is error-compensated (hi/lo operand split, three TF32 products, fp32 accumulation), i.e. fp32
    accuracy at about three times the GEMM time.  Returns the previous mode."""
    from . import _lib
    names = {"tf32": "", "3xtf32": "precise"}
    if mode not in names:
        raise ValueError('precision must be "tf32" or "3xtf32"')
    prev = _lib.set_variant(names[mode])
    return {"": "tf32", "precise": "3xtf32"}.get(prev, prev)


class precision:
    """context manager form of set_precision"""

    def __init__(self, mode):
        self.mode = mode

    def __enter__(self):
        self.prev = set_precision(self.mode)

    def __exit__(self, *exc):
        set_precision(self.prev)
