"""Callbacks used by the hot-path scripts: LearningRateMonitor (finetune_ziya_llama.py:219) and ModelCheckpoint, the
base of fengshen.utils.universal_checkpoint.UniversalCheckpoint (universal_checkpoint.py:5-41)."""
import os

import torch


class Callback:
    def on_fit_start(self, trainer, pl_module):
        pass

    def on_train_batch_end(self, trainer, pl_module, outputs, batch, batch_idx):
        pass

    def on_train_epoch_end(self, trainer, pl_module):
        pass

    def on_fit_end(self, trainer, pl_module):
        pass


class LearningRateMonitor(Callback):
    def __init__(self, logging_interval=None, log_momentum=False):
        self.logging_interval = logging_interval

    def on_train_batch_end(self, trainer, pl_module, outputs, batch, batch_idx):
        if trainer.optimizers:
            trainer.callback_metrics["lr-" + type(trainer.optimizers[0]).__name__] = \
                trainer.optimizers[0].param_groups[0]["lr"]


class ModelCheckpoint(Callback):
    """Saves `{'state_dict', 'global_step', 'epoch', 'global_samples', optimizer shard}` under
    dirpath/<filename>.ckpt/ — a DIRECTORY per checkpoint like DeepSpeed's: mp_rank_00_model_states.pt
    (key 'module', read back by examples/pretrain_t5/convert_ckpt_to_bin.py:21) + zero_pp_rank_R_mp_rank_00_optim_states.pt."""

    def __init__(self, dirpath=None, filename=None, monitor=None, save_top_k=1, mode="min", save_last=False,
                 every_n_train_steps=None, save_weights_only=False, every_n_epochs=None, save_on_train_epoch_end=None,
                 **_):
        self.dirpath, self.filename, self.monitor, self.mode = dirpath, filename or "model-{step}", monitor, mode
        self.save_top_k, self.save_last = save_top_k, save_last
        self.every_n_train_steps = None if every_n_train_steps is None else int(every_n_train_steps)
        self.save_weights_only, self.every_n_epochs = save_weights_only, every_n_epochs
        self.saved = []

    def _path(self, trainer, last=False):
        if last:
            name = "last"
        else:
            import re
            vals = {k: (float(v) if hasattr(v, "__float__") else v) for k, v in trainer.callback_metrics.items()}
            vals.update(epoch=trainer.current_epoch, step=trainer.global_step)

            def fill(m):   # {epoch:02d}, {step:d}, {train_loss:.4f}, ... ; metrics not logged yet format as 0
                v = vals.get(m.group(1), 0)
                spec = m.group(2) or ""
                return format(int(v) if spec.endswith("d") else v, spec)
            name = re.sub(r"\{(\w+)(?::([^}]*))?\}", fill, self.filename)
        return os.path.join(self.dirpath or ".", name + ".ckpt")

    def on_train_batch_end(self, trainer, pl_module, outputs, batch, batch_idx):
        n = self.every_n_train_steps
        if n and trainer.global_step > 0 and trainer.global_step % n == 0 and trainer.just_stepped:
            trainer.save_checkpoint(self._path(trainer), weights_only=self.save_weights_only)
            self.saved.append(self._path(trainer))
            if self.save_last:
                trainer.save_checkpoint(self._path(trainer, last=True), weights_only=self.save_weights_only)
            k = int(self.save_top_k) if self.save_top_k is not None else -1
            while k >= 0 and len(self.saved) > max(k, 1):
                old = self.saved.pop(0)
                if trainer.global_rank == 0 and os.path.isdir(old):
                    import shutil
                    shutil.rmtree(old, ignore_errors=True)

    def on_fit_end(self, trainer, pl_module):
        if self.save_last:
            trainer.save_checkpoint(self._path(trainer, last=True), weights_only=self.save_weights_only)
