"""The whole ATST-clip training step as ONE CUDA graph.

A small-batch step is launch-bound: ~500 kernel launches of a few microseconds of GPU work each, every one paid for
with a Python call, a ctypes transition and a driver launch (config 1: 5.7 ms per 8-clip step of which the GPU is busy
for a fraction).  ``GraphedTrainStep`` captures ``training_step -> backward -> optimizer.step -> EMA`` once - mel is
outside: the batch is what ``training_step`` takes - and replays it with one launch per step:

  * inputs are copied into static tensors before the replay; the loss / logged statistics are static outputs;
  * the per-step scalars (learning rate and weight decay of both parameter groups with Adam's bias corrections, EMA
    momentum) live in a 5-float device tensor that the AdamW and EMA kernels read (``dyn`` / ``m_dev`` arguments of
    the C ABI), refreshed from the module's schedules before each replay;
  * the backward pass is the engine's own (``_Runtime.backward``), called directly rather than through
    ``loss.backward()`` - autograd's device thread cannot take part in a capture;
  * DropPath draws come from torch's CUDA generator, which is graph-safe (the Philox offset advances per replay);
  * the flat gradient buffer, the workspace and the optimizer moments are the eager path's own - switching between
    eager and graphed steps is allowed, checkpoints are unaffected.

Constraints: one GPU per process without a process group (the data-parallel exchange is not captured), fixed shapes
(batch size, crop widths), ATST-clip only (the frame model reads its masked-row count on the host).  Warm-up and
capture run real kernels on the model's state; it is snapshotted before and restored after, so constructing the
object does not advance training.
"""
import torch

from .distributed import world


class GraphedTrainStep:
    def __init__(self, module, optimizer, example_batch, warmup=3):
        """module: ATSTLightningModule on cuda, optimizer: its FusedHFAdamW, example_batch: ((melspecs, lengths), _)."""
        if world() > 1:
            raise RuntimeError("GraphedTrainStep captures the single-GPU step; run eager steps under torch.distributed")
        self.lm, self.opt = module, optimizer
        (mels, lens), _ = example_batch
        self.static_mels = [m.detach().clone() for m in mels]
        self.static_lens = [l.detach().clone() for l in lens]
        dev = self.static_mels[0].device
        self.scalars = torch.zeros(5, device=dev)
        self._one = torch.ones((), device=dev)  # d loss / d loss
        self._host = torch.zeros(5).pin_memory()
        model = module.model
        rt = model._runtime(dev)
        module.trainer.optimizers = [optimizer]
        optimizer.device_scalars, model.ema_device_scalar = self.scalars[:4], self.scalars[4:5]
        step0 = int(module.global_step)
        self._fill(step0, optimizer._step + 1)
        # ---- snapshot everything the warm-up / capture runs will touch
        moments = optimizer._moments(rt.fs)
        snap = [t.clone() for t in (rt.fs.data, rt.ft.data, *moments)]
        bufs = [(b, b.clone()) for b in model.buffers()]
        opt_step = optimizer._step
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):
            for _ in range(warmup):
                self._eager(step0)
        torch.cuda.current_stream(dev).wait_stream(side)
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            self.loss = self._eager(step0)
        self.logged = dict(module.logged) if hasattr(module, "logged") else {}
        # ---- restore: constructing the graph must not advance training
        for dst, src in zip((rt.fs.data, rt.ft.data, *moments), snap):
            dst.copy_(src)
        for b, saved in bufs:
            b.copy_(saved)
        optimizer._step = opt_step
        rt.fs.grads_pending = False
        torch.cuda.synchronize(dev)

    def _eager(self, step):
        lm, opt = self.lm, self.opt
        lm.global_step = step
        loss = lm.training_step(((self.static_mels, self.static_lens), None), step)
        opt.zero_grad()
        # the engine's explicit backward pass, called directly: loss.backward() would hand the same call to autograd's
        # device thread, whose stream bookkeeping creates a dependency on work outside the capture
        # (cudaErrorStreamCaptureIsolation); nothing in this step is differentiated by autograd anyway
        lm.model._rt.backward(self._one)
        opt.step()
        lm.on_train_batch_end(None, None, step)
        return loss.detach()

    def _fill(self, schedule_step, opt_step):
        lm = self.lm
        vals = self.opt.step_scalars(opt_step, float(lm.mylr_scheduler[schedule_step]), float(lm.wd_scheduler[schedule_step]))
        vals.append(float(lm.ema_scheduler[schedule_step]))
        self._host.copy_(torch.tensor(vals, dtype=torch.float32))
        self.scalars.copy_(self._host, non_blocking=True)

    def __call__(self, batch, step):
        """one training step on ``batch`` at schedule position ``step``; returns the (static) loss tensor."""
        (mels, lens), _ = batch
        for dst, src in zip(self.static_mels, mels):
            dst.copy_(src, non_blocking=True)
        for dst, src in zip(self.static_lens, lens):
            dst.copy_(src, non_blocking=True)
        self._fill(step, self.opt._step + 1)
        self.graph.replay()
        self.opt._step += 1
        self.lm.global_step = step
        return self.loss

    def release(self):
        """detach the device scalars: the module and optimizer go back to taking their scalars from the host."""
        self.opt.device_scalars = None
        self.lm.model.ema_device_scalar = None
