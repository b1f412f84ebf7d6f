from .lmdb import LMDBDataset, write_dataset  # noqa: F401
from .prefetch import DevicePrefetcher, collate_waveforms  # noqa: F401
