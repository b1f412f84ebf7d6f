"""LMDBDataset - audiossl/datasets/lmdb.py:12-99 with the same constructor, item layout, ``cycle()`` and attributes,
reading the reference's on-disk format without liblmdb / legacy pyarrow (datasets/lmdb_format.py,
datasets/arrow_legacy.py): ``<db_path>/{train,valid,eval}.lmdb`` holding ``key -> serialize((waveform f32 [1,n],
label [1,C]))`` plus ``__keys__`` / ``__len__``."""
import os
import random
from copy import deepcopy

import numpy as np
import torch
import torch.utils.data as data

from . import arrow_legacy
from .lmdb_format import LMDBReader

random.seed(1234)


class LMDBDataset(data.Dataset):
    def __init__(self, db_path, split, subset=None, transform=None, target_transform=None, return_key=False):
        self.db_path = db_path
        self.return_key = return_key
        name = {"train": "train.lmdb", "valid": "valid.lmdb"}.get(split, "eval.lmdb")
        self.lmdb_path = os.path.join(self.db_path, name)
        self.subset = subset
        self.env = LMDBReader(self.lmdb_path)
        self.length = arrow_legacy.loads(self.env.get(b'__len__'))
        self.keys = arrow_legacy.loads(self.env.get(b'__keys__'))
        self.org_keys = deepcopy(self.keys)
        self.start = 0
        if subset is not None and subset < self.length:
            self.length = subset
            random.shuffle(self.keys)
            self.org_keys = deepcopy(self.keys)
            self.keys = self.keys[:subset]
            self.start = subset
        self.transform = transform
        self.target_transform = target_transform
        unpacked = arrow_legacy.loads(self.env.get(self.keys[0]))
        self.num_classes = unpacked[1].shape[1]
        self.sr = 16000

    def read(self, index):
        """(waveform [n] float32 tensor, label [C] tensor, key) of record ``index`` (no transform)."""
        key = self.keys[index]
        unpacked = arrow_legacy.loads(self.env.get(key))
        # the arrays are read-only views into the memory map: copy them out (one 640 KB memcpy for a 10 s clip)
        wav, label = (torch.from_numpy(np.array(a)) for a in unpacked[:2])
        return wav.squeeze(0), label.squeeze(0), key

    def __getitem__(self, index):
        waveform, label, key = self.read(index)
        if self.transform is not None:
            transformed = self.transform(waveform)
            if self.target_transform is not None:
                transformed = list(transformed)
                transformed[0], label = self.target_transform(transformed[0], label)
                transformed = tuple(transformed)
            return (transformed, label, key) if self.return_key else (transformed, label)
        return (waveform, label, key) if self.return_key else (waveform, label)

    def cycle(self):
        if self.start + self.subset > len(self.org_keys):
            self.keys = self.org_keys[self.start:] + self.org_keys[:self.start + self.subset - len(self.org_keys)]
            random.shuffle(self.org_keys)
            self.start = 0
        else:
            self.keys = self.org_keys[self.start:self.start + self.subset]
            self.start = self.start + self.subset

    def __len__(self):
        return len(self.keys)

    def __repr__(self):
        return self.__class__.__name__ + ' (' + self.db_path + ')'

    def __getstate__(self):  # DataLoader workers re-open the map instead of pickling it
        d = dict(self.__dict__)
        d["env"] = None
        return d

    def __setstate__(self, d):
        self.__dict__.update(d)
        self.env = LMDBReader(self.lmdb_path)


def write_dataset(lmdb_path, records):
    """dataset2lmdb.py:108-129 for an iterable of (name, waveform [1,n] float32 ndarray, label [1,C] ndarray):
    ``name -> serialize((waveform, label))`` plus ``__keys__`` and ``__len__``."""
    from .lmdb_format import write_lmdb
    items, keys = {}, []
    for name, wav, label in records:
        key = u'{}'.format(name).encode('ascii')
        keys.append(key)
        items[key] = arrow_legacy.dumps((wav, label))
    items[b'__keys__'] = arrow_legacy.dumps(keys)
    items[b'__len__'] = arrow_legacy.dumps(len(keys))
    return write_lmdb(lmdb_path, items)
