"""Loggers named by the hot-path scripts (WandbLogger finetune_ziya_llama.py:218; TensorBoardLogger pretrain_t5.py:154).
Neither wandb nor a network exists here: both write one JSON line per logged step to <save_dir>/metrics.jsonl."""
import json
import os


class _JsonlLogger:
    def __init__(self, save_dir=".", name="default", project=None, **_):
        self.save_dir = os.path.join(save_dir or ".", project or "", name or "")
        self._f = None

    def log_metrics(self, metrics, step):
        if self._f is None:
            os.makedirs(self.save_dir, exist_ok=True)
            self._f = open(os.path.join(self.save_dir, "metrics.jsonl"), "a")
        self._f.write(json.dumps({"step": step, **{k: float(v) for k, v in metrics.items()}}) + "\n")
        self._f.flush()


class WandbLogger(_JsonlLogger):
    pass


class TensorBoardLogger(_JsonlLogger):
    pass
