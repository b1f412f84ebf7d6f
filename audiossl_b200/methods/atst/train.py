"""Pre-training launcher - the recipe of audiossl/methods/atst/train.py:11-48 without a Lightning ``Trainer``:

    python -m audiossl_b200.methods.atst.train --save_path out --nproc 1 --data_path /data/audioset [...]
    torchrun --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 -m audiossl_b200.methods.atst.train --nproc 8 ...

One process per GPU (torchrun sets RANK / LOCAL_RANK / WORLD_SIZE).  What the reference's
``Trainer(strategy="ddp_find_unused_parameters_true", sync_batchnorm=True)`` provides is done by the module itself:
the step's explicit backward writes one flat gradient buffer that is all-reduced over NCCL, BatchNorm statistics are
exchanged across ranks, and the never-used ``mask_embed`` is simply outside the exchanged range - so the model is NOT
wrapped in DistributedDataParallel (its reducer would never fire: gradients do not come from autograd hooks).
Same hyper-parameter handling: ``learning_rate *= nproc * batch_size_per_gpu / 256``; ``last.ckpt`` in the Lightning
layout (``state_dict`` / ``hyper_parameters`` / ``optimizer_states`` / ``global_step``) is written every
``--save_every`` steps and resumed from when present."""
import os
import time
from argparse import ArgumentParser

import torch
import torch.distributed as dist

from .data import ATSTDataModule
from .model import ATSTLightningModule


def save_checkpoint(path, lm, opt, step, hparams):
    tmp = path + ".tmp"
    torch.save({"state_dict": {k: v.detach().cpu() for k, v in lm.state_dict().items()},
                "hyper_parameters": hparams, "optimizer_states": [opt.state_dict()], "global_step": step}, tmp)
    os.replace(tmp, path)


def main(args):
    return run(args, ATSTLightningModule, ATSTDataModule, ("std_cls_s", "std_cls_t"))


def run(args, module_cls, data_cls, std_names):
    """the training loop shared by the ATST-clip and ATST-Frame launchers"""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.nproc:
        raise SystemExit("--nproc %d but WORLD_SIZE=%d: launch one process per GPU with torchrun" % (args.nproc, world))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1 and not dist.is_initialized():
        dist.init_process_group("nccl", device_id=dev)
    args.learning_rate = args.learning_rate * args.nproc * args.batch_size_per_gpu / 256
    dict_args = vars(args)
    torch.manual_seed(0)  # identical initial weights on every rank (DDP broadcasts rank 0's; same effect)
    model = module_cls(**dict_args).to(dev).train()
    data = data_cls(device=dev, **dict_args)
    opt = model.configure_optimizers()[0]
    model.trainer.optimizers = [opt]
    step = 0
    os.makedirs(args.save_path, exist_ok=True)
    last_ckpt = os.path.join(args.save_path, "last.ckpt")
    if os.path.exists(last_ckpt):
        ck = torch.load(last_ckpt, map_location="cpu", weights_only=False)
        model.load_state_dict(ck["state_dict"])
        step = int(ck.get("global_step", 0))
        model.global_step = step
        if ck.get("optimizer_states"):
            model.model._runtime(dev)  # the flat buffers the optimizer state maps onto
            opt.load_state_dict(ck["optimizer_states"][0])
    epoch, t0, seen = 0, time.time(), 0
    while step < args.max_steps:
        for batch in data.train_dataloader(rank, world, seed=epoch):
            model.global_step = step
            loss = model.training_step(batch, step)
            opt.zero_grad()
            loss.backward()
            opt.step()
            model.on_train_batch_end(None, batch, step)
            step += 1
            seen += args.batch_size_per_gpu * world
            if rank == 0 and step % args.log_every == 0:
                dt = time.time() - t0
                print("step %d  loss %.4f  std_s %.3f  std_t %.3f  lr %.2e  %.0f clips/s" % (
                    step, float(loss.detach()), float(model.logged[std_names[0]]), float(model.logged[std_names[1]]),
                    model.logged["lr"], seen / dt), flush=True)
                t0, seen = time.time(), 0
            if rank == 0 and (step % args.save_every == 0 or step == args.max_steps):
                save_checkpoint(last_ckpt, model, opt, step, dict(dict_args))
            if step >= args.max_steps:
                break
        epoch += 1
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return model


def build_parser():
    parser = ArgumentParser("ATST")
    parser.add_argument("--save_path", type=str, required=True)
    parser.add_argument('--nproc', type=int, default=1)
    parser.add_argument('--log_every', type=int, default=50)
    parser.add_argument('--save_every', type=int, default=1000)
    parser = ATSTLightningModule.add_model_specific_args(parser)
    parser = ATSTDataModule.add_data_specific_args(parser)
    return parser


if __name__ == "__main__":
    main(build_parser().parse_args())
