"""ATSTLightningModule - audiossl/methods/atst/model.py:6-65 with the same constructor kwargs, hooks,
logged quantities and argparse group.  If pytorch_lightning is importable the class derives from
``LightningModule``; otherwise from a minimal stand-in with the attributes the hooks use
(``global_step``, ``trainer.optimizers``, ``log``, ``save_hyperparameters``) so that the same
training_step / configure_optimizers / on_train_batch_end code drives bench.py and the tests.
"""
import torch
from torch import nn

from ...models.atst import ATST
from ...optim import FusedHFAdamW
from ...utils.common import cosine_scheduler_step, get_params_groups

try:  # pragma: no cover - pytorch_lightning is absent from the build image
    from pytorch_lightning import LightningModule
except Exception:  # noqa: BLE001
    class _Trainer:
        def __init__(self):
            self.optimizers = []

    class LightningModule(nn.Module):
        """the slice of the LightningModule protocol the ATST recipes and loaders touch: ``global_step``,
        ``trainer.optimizers``, ``log``, ``save_hyperparameters`` (records the constructor arguments in ``hparams``)
        and ``load_from_checkpoint`` for the checkpoint layout Lightning writes (``state_dict`` +
        ``hyper_parameters``)."""

        def __init__(self):
            super().__init__()
            self.global_step = 0
            self.trainer = _Trainer()
            self.logged = {}
            self.hparams = {}

        def log(self, name, value, **kwargs):
            self.logged[name] = value

        def save_hyperparameters(self, *args, **kwargs):
            import inspect
            frame = inspect.currentframe().f_back
            info = inspect.getargvalues(frame)
            hp = {k: info.locals[k] for k in info.args if k != "self"}
            if info.keywords:
                hp.update(info.locals[info.keywords])
            self.hparams = hp

        @classmethod
        def load_from_checkpoint(cls, checkpoint_path, map_location=None, strict=True, **kwargs):
            ck = torch.load(checkpoint_path, map_location=map_location or "cpu", weights_only=False)
            hp = dict(ck.get("hyper_parameters", {}))
            hp.update(kwargs)
            module = cls(**hp)
            module.load_state_dict(ck["state_dict"], strict=strict)
            return module


class ATSTLightningModule(LightningModule):
    def __init__(self, arch="small", learning_rate: float = 5e-4, warmup_steps=1300, max_steps=39000, ema=0.99,
                 ncrops=2, **kwargs):
        super().__init__()
        # ncrops / drop_path_rate are pass-through extensions; defaults keep the reference behaviour
        model_kwargs = {k: kwargs[k] for k in ("drop_path_rate",) if k in kwargs}
        self.model = ATST(arch=arch, ncrops=ncrops, **model_kwargs)
        self.learning_rate = learning_rate
        self.warmup_steps = warmup_steps
        self.max_steps = max_steps
        self.ema_scheduler = cosine_scheduler_step(ema, 1, max_steps, 0)
        self.wd_scheduler = cosine_scheduler_step(0.04, 0.4, max_steps, 0)
        self.mylr_scheduler = cosine_scheduler_step(learning_rate, 1e-6, max_steps, warmup_steps)
        self.save_hyperparameters()

    def training_step(self, batch, batch_idx):
        self.schedule()
        (melspecs, lengths), _ = batch
        loss, std_cls_s, std_cls_t = self.model(melspecs, lengths)
        self.log("loss", loss, prog_bar=True, logger=True)
        self.log("std_cls_t", std_cls_t, prog_bar=True, logger=True)
        self.log("std_cls_s", std_cls_s, prog_bar=True, logger=True)
        self.log("ema", self.ema_scheduler[self.global_step], prog_bar=True, logger=True)
        self.log("step", self.global_step, prog_bar=True, logger=True)
        return loss

    def schedule(self):
        for i, param_group in enumerate(self.trainer.optimizers[0].param_groups):
            param_group["lr"] = self.mylr_scheduler[self.global_step]
            if i == 0:  # only the first group is regularized
                param_group["weight_decay"] = self.wd_scheduler[self.global_step]
        self.log("wd", self.wd_scheduler[self.global_step], prog_bar=True, logger=True)
        self.log("lr", param_group["lr"], prog_bar=True, logger=True)

    def configure_optimizers(self):
        def flat():
            p = next(self.model.student.parameters())
            return self.model._runtime(p.device).fs
        optimizer = FusedHFAdamW(get_params_groups(self.model.student), flat=flat, lr=self.learning_rate,
                                 weight_decay=0.)
        return [optimizer]

    def on_train_batch_end(self, outputs, batch, batch_idx: int, unused: int = 0) -> None:
        m = self.ema_scheduler[self.global_step]
        self.model.update_teacher(m)

    @staticmethod
    def add_model_specific_args(parent_parser):
        parser = parent_parser.add_argument_group("ATSTModel")
        parser.add_argument("--arch", type=str, default="small")
        parser.add_argument("--learning_rate", default=0.0005, type=float, help="""Learning rate at the end of
            linear warmup (highest LR used during training). The learning rate is linearly scaled
            with the batch size, and specified here for a reference batch size of 256.""")
        parser.add_argument('--ema', default=0.99, type=float, help="""Base EMA
            parameter for teacher update. The value is increased to 1 during training with cosine schedule.
            """)
        parser.add_argument('--warmup_steps', default=1300, type=int)
        parser.add_argument('--max_steps', default=39010, type=int)
        return parent_parser
