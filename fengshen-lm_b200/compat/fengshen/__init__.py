"""`fengshen` import paths of the data-parallel pretraining hot path, backed by fsb200.

Only the modules the hot-path example scripts import exist here (examples/ziya_llama/finetune_ziya_llama.py:11-20):
    fengshen.models.model_utils, fengshen.models.llama.{modeling_llama,configuration_llama}, fengshen.models.megatron.mpu,
    fengshen.data.universal_datamodule, fengshen.utils.universal_checkpoint, fengshen.strategies.megatron_deepspeed.
Unlike the reference's fengshen/__init__.py:16-19 nothing is imported eagerly.
"""
