"""ATSTDataModule - audiossl/methods/atst/data.py:6-42 with the same constructor and argparse group, re-plumbed for a
GPU front-end: the reference runs ``ATSTTrainTransform`` (mel + augmentations) per sample on 10 DataLoader workers and
ships mels; here the workers only read raw waveforms from LMDB, batches are staged through pinned memory onto the GPU
(datasets/prefetch.py) and ``BatchedATSTTrainTransform`` produces the training batch there.
``train_dataloader()`` yields ``((melspecs, lengths), labels)`` - what ``ATSTLightningModule.training_step`` takes."""
import torch
from torch.utils import data

from ...datasets import DevicePrefetcher, LMDBDataset, collate_waveforms
from .transform import BatchedATSTTrainTransform


class SyntheticWaveforms(data.Dataset):
    """seeded Gaussian 16 kHz clips (SURVEY.md section 8d synthetic inputs) with a one-hot dummy label."""

    def __init__(self, n_clips, seconds=10.0, num_classes=527, seed=1234):
        self.n_clips, self.n, self.num_classes, self.seed = n_clips, int(seconds * 16000), num_classes, seed

    def __len__(self):
        return self.n_clips

    def __getitem__(self, i):
        g = torch.Generator().manual_seed(self.seed + i)
        label = torch.zeros(self.num_classes)
        label[i % self.num_classes] = 1.0
        return torch.randn(self.n, generator=g) * 0.1, label


class _DeviceBatches:
    def __init__(self, loader, device, transform):
        self.loader, self.device, self.transform = loader, device, transform

    def __len__(self):
        return len(self.loader)

    def __iter__(self):
        for wav, labels in DevicePrefetcher(self.loader, self.device):
            yield self.transform(wav), labels


class ATSTDataModule:
    def __init__(self, data_path=None, batch_size_per_gpu=256, num_workers=10, subset=200000, train_len=6.0,
                 clip_seconds=10.0, synthetic_clips=0, device=None, augment=True, **kwargs):
        if data_path is None:
            if synthetic_clips <= 0:
                raise ValueError("ATSTDataModule needs --data_path (LMDB directory) or synthetic_clips > 0")
            self.dataset = SyntheticWaveforms(synthetic_clips, clip_seconds)
        else:
            self.dataset = LMDBDataset(data_path, split="train", subset=subset, transform=None)
        self.batch_size, self.num_workers = batch_size_per_gpu, num_workers
        self.clip_samples = int(clip_seconds * 16000)
        self.device = device
        self.transform = BatchedATSTTrainTransform(anchor_len=(train_len, train_len), augment=augment)

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
        parser = parent_parser.add_argument_group("ATSTData")
        parser.add_argument("--data_path", type=str, default=None, help="data path")
        parser.add_argument('--batch_size_per_gpu', default=256, type=int,
                            help='Per-GPU batch-size : number of distinct samples loaded on one GPU.')
        parser.add_argument('--num_workers', default=10, type=int, help='Number of data loading workers per GPU.')
        parser.add_argument('--subset', default=200000, type=int, help='subset of training data')
        parser.add_argument('--train_len', default=6.0, type=float, help='length of training segment')
        parser.add_argument('--synthetic_clips', default=0, type=int,
                            help='train on this many seeded synthetic clips instead of an LMDB (no --data_path)')
        return parent_parser
