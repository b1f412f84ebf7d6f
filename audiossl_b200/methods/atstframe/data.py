"""FrameATSTDataModule - audiossl/methods/atstframe/data.py:20-107 with the same constructor and argparse group,
re-plumbed like the clip data module (methods/atst/data.py): DataLoader workers read raw waveforms from the LMDB,
batches go through pinned memory to the GPU, ``BatchedFrameATSTTrainTransform`` produces
``((melspecs, lengths, masks), labels)`` there."""
import torch
from torch.utils import data

from ...datasets import DevicePrefetcher, LMDBDataset, collate_waveforms
from ...utils.common import bool_flag
from ..atst.data import SyntheticWaveforms, _DeviceBatches
from .transform import BatchedFrameATSTTrainTransform


class FrameATSTDataModule:
    def __init__(self, data_path=None, batch_size_per_gpu=256, num_workers=10, subset=200000, win_length=1024,
                 aug_tea=True, aug_stu=True, freq_wrap=True, mix_up=True, mask_ratio=0.75, mask_type="block",
                 anchor_len=6., mask_len=5, min_mask_len=2, n_mels=64, clip_seconds=10.0, synthetic_clips=0, device=None,
                 **kwargs):
        if data_path is None:
            if synthetic_clips <= 0:
                raise ValueError("FrameATSTDataModule needs --data_path (LMDB directory) or synthetic_clips > 0")
            self.dataset = SyntheticWaveforms(synthetic_clips, clip_seconds)
        else:
            self.dataset = LMDBDataset(data_path, split="train", subset=subset, transform=None)
        self.batch_size, self.num_workers = batch_size_per_gpu, num_workers
        self.clip_samples = int(max(clip_seconds, anchor_len) * 16000)
        self.device = device
        self.transform = BatchedFrameATSTTrainTransform(
            win_length=win_length, aug_tea=aug_tea, aug_stu=aug_stu, freq_wrap=freq_wrap, mix_up=mix_up,
            mask_ratio=mask_ratio, anchor_len=anchor_len, mask_type=mask_type, mask_len=mask_len,
            min_mask_len=min_mask_len, n_mels=n_mels)

    def host_loader(self, rank=0, world=1, seed=0):
        sampler = None
        if world > 1:
            sampler = data.distributed.DistributedSampler(self.dataset, num_replicas=world, rank=rank, shuffle=True,
                                                          seed=seed, drop_last=True)
        n = self.clip_samples
        return data.DataLoader(self.dataset, batch_size=self.batch_size, num_workers=self.num_workers, sampler=sampler,
                               shuffle=sampler is None, drop_last=True, pin_memory=True,
                               collate_fn=lambda samples: collate_waveforms(samples, n),
                               persistent_workers=self.num_workers > 0)

    def train_dataloader(self, rank=0, world=1, seed=0):
        device = self.device or torch.device("cuda", torch.cuda.current_device())
        return _DeviceBatches(self.host_loader(rank, world, seed), device, self.transform)

    @staticmethod
    def add_data_specific_args(parent_parser):
        parser = parent_parser.add_argument_group("FrameATSTData")
        parser.add_argument("--data_path", type=str, default=None, help="data path")
        parser.add_argument('--batch_size_per_gpu', default=256, type=int,
                            help='Per-GPU batch-size : number of distinct samples loaded on one GPU.')
        parser.add_argument('--num_workers', default=10, type=int, help='Number of data loading workers per GPU.')
        parser.add_argument('--subset', default=200000, type=int, help='subset of training data')
        parser.add_argument('--win_length', default=1024, type=int, help='windown length')
        parser.add_argument('--aug_tea', default=True, type=bool_flag, help='augment the view fed into the teacher')
        parser.add_argument('--aug_stu', default=True, type=bool_flag, help='augment the view fed into the student')
        parser.add_argument('--freq_wrap', default=True, type=bool_flag, help='freq wraping or not')
        parser.add_argument('--mix_up', default=True, type=bool_flag, help='mixup or not')
        parser.add_argument('--anchor_len', default=6., type=float, help="length of training samples")
        parser.add_argument('--mask_ratio', default=0.75, type=float, help="masking ratio")
        parser.add_argument('--mask_len', default=5, type=int, help="masking block length")
        parser.add_argument('--min_mask_len', default=2, type=int, help="minimum masking block length")
        parser.add_argument('--n_mels', default=64, type=int, help="number of mel channels")
        parser.add_argument('--mask_type', default="block", type=str, help="masking type: random or block")
        parser.add_argument('--synthetic_clips', default=0, type=int,
                            help='train on this many seeded synthetic clips instead of an LMDB (no --data_path)')
        return parent_parser
