"""Public step API: what a user (or the Lightning-compatible shim in compat/) calls once per optimizer step.

`PretrainStep.step(host_batches)` performs, for each micro-batch: the host->device copy of the token tensors from pinned
memory, forward + backward through the fsb200 model, the engine's bucketed reduce-scatter / fp32 accumulation; then the
sharded AdamW update + parameter all-gather, the LR schedule, and returns the mean loss as a Python float (one
device->host read per step, the `self.log('train/loss', ...)` of the reference's training_step,
examples/ziya_llama/finetune_ziya_llama.py:133-148).
"""
import torch

from .engine import ZeroEngine


class PretrainStep:
    def __init__(self, model, lr_fn, lr=1e-4, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.1, grad_clip=0.0, ga_steps=1,
                 process_group=None, stage=2, comm_sms=0):
        self.model = model
        self.engine = ZeroEngine(model, lr=lr, betas=betas, eps=eps, weight_decay=weight_decay, grad_clip=grad_clip,
                                 ga_steps=ga_steps, process_group=process_group, stage=stage, comm_sms=comm_sms)
        self.lr_fn = lr_fn
        self.global_step = 0
        self.device = model.flat.params.device
        self._loss_acc = torch.zeros((), dtype=torch.float32, device=self.device)

    def step_device(self, device_batches):
        """One optimizer step from batches already resident on the device. Returns the loss as a 0-d device tensor."""
        eng = self.engine
        self._loss_acc.zero_()
        for b in device_batches:
            out = self.model(**b)
            out.loss.backward()
            eng.backward_done()
            self._loss_acc += out.loss.detach()
        eng.step(lr=self.lr_fn(self.global_step))
        self.global_step += 1
        return self._loss_acc / len(device_batches)

    def step(self, host_batches):
        """One optimizer step from pinned host batches (dicts of CPU tensors). Returns the loss as a float."""
        dev = [{k: v.to(self.device, non_blocking=True) for k, v in b.items()} for b in host_batches]
        return float(self.step_device(dev).item())
