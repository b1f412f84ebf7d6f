"""audiossl_b200 - B200-native (sm_100a) implementation of the ATST pre-training hot path of
Audio-WestlakeU/audiossl, behind the reference's own Python surface.

    audiossl_b200.transforms                 <-> audiossl.transforms
    audiossl_b200.models.atst.ATST           <-> audiossl.models.atst.ATST
    audiossl_b200.methods.atst.model         <-> audiossl.methods.atst.model (ATSTLightningModule)
    audiossl_b200.methods.atst.transform     <-> audiossl.methods.atst.transform (ATSTTrainTransform)

Compute goes through libatst_b200.so (include/atst_b200.h); there is no CPU or eager fallback.
"""
__version__ = "0.1.0"
