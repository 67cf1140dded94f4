"""Minimal `deepspeed` surface for the Fengshen hot path: the optimizer classes that
fengshen/models/model_utils.py:3 imports, the activation-checkpointing / RNG symbols that
fengshen/models/megatron/mpu/random.py:18-37 re-exports, and the topology class used at
fengshen/strategies/megatron_deepspeed.py:347. The ZeRO engine itself is fsb200.engine.ZeroEngine; DeepSpeed (3P,
`deepspeed>=0.5.10`, setup.py:20) cannot be installed offline."""
from . import checkpointing, ops, runtime  # noqa: F401

__version__ = "0.9.0+fsb200"


def init_distributed(dist_backend="nccl", **_):
    import torch.distributed as dist
    if not dist.is_initialized():
        dist.init_process_group(dist_backend)


def initialize(*args, **kwargs):
    raise NotImplementedError("deepspeed.initialize is replaced by fsb200.engine.ZeroEngine (driven by the "
                              "pytorch_lightning-compatible Trainer in fengshen-lm_b200/compat)")
