"""Minimal `pytorch_lightning` 1.x surface for the Fengshen hot-path example scripts.

pytorch_lightning is a third-party dependency of the reference (setup.py:19, `>=1.5.10`) that is not installable here;
this package provides exactly the names those scripts touch — `LightningModule`, `LightningDataModule`,
`Trainer.{add_argparse_args, from_argparse_args, fit, predict}`, callbacks, loggers, `strategies.DeepSpeedStrategy` and
the strategy-registry strings `deepspeed_stage_{1,2,3}[_offload]` / `ddp` — with the step executed by the fsb200 engine
instead of Lightning's loops + DeepSpeed. Hook order follows PL 1.8: setup_environment -> module.setup('fit') ->
configure_optimizers -> strategy setup -> fit loop (examples/ziya_llama/finetune_ziya_llama.py:185-227 is the model).
Put `fengshen-lm_b200/compat` on PYTHONPATH to activate it (it never shadows a real installation silently: importing it
when the real package is importable is the user's explicit choice of path order).
"""
from .core import LightningDataModule, LightningModule, seed_everything  # noqa: F401
from .trainer import Trainer  # noqa: F401
from . import callbacks, loggers, strategies  # noqa: F401

__version__ = "1.9.0+fsb200"
