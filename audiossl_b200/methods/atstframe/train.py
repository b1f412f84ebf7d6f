"""ATST-Frame pre-training launcher - the recipe of audiossl/methods/atstframe/train.py:12-60 on the shared loop of
methods/atst/train.py (one process per GPU under torchrun, no DDP wrapper, no Lightning ``Trainer``):

    torchrun --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 -m audiossl_b200.methods.atstframe.train \\
        --save_path out --nproc 4 --data_path /data/audioset --arch small --anchor_len 10 --mask_ratio 0.65 ...
"""
from argparse import ArgumentParser

from ..atst.train import run
from .data import FrameATSTDataModule
from .model import FrameATSTLightningModule


def main(args):
    args.spec_h = args.n_mels
    return run(args, FrameATSTLightningModule, FrameATSTDataModule, ("std_frm_stu", "std_frm_tea"))


def build_parser():
    parser = ArgumentParser("FrameATST")
    parser.add_argument("--save_path", type=str, required=True)
    parser.add_argument('--nproc', type=int, default=1)
    parser.add_argument('--patch_h', type=int, default=64)
    parser.add_argument('--patch_w', type=int, default=4)
    parser.add_argument('--log_every', type=int, default=50)
    parser.add_argument('--save_every', type=int, default=1000)
    parser = FrameATSTLightningModule.add_model_specific_args(parser)
    parser = FrameATSTDataModule.add_data_specific_args(parser)
    return parser


if __name__ == "__main__":
    main(build_parser().parse_args())
