import torch


def print_rank_0(message):
    """fengshen/data/megatron_dataloader/utils.py:18-24."""
    if not torch.distributed.is_initialized() or torch.distributed.get_rank() == 0:
        print(message, flush=True)
