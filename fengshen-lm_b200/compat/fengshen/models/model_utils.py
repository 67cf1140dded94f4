"""Optimizer / LR-schedule / step-count factory shared by the example scripts — same functions, flags, defaults and
name-based weight-decay grouping as the reference's fengshen/models/model_utils.py (cited per function)."""
from deepspeed.ops.adam import DeepSpeedCPUAdam, FusedAdam
from pytorch_lightning.strategies import DeepSpeedStrategy
import torch
from torch.optim.lr_scheduler import LambdaLR

from fsb200.schedules import linear_lr, polynomial_lr


def add_module_args(parent_args):
    """model_utils.py:13-28 — identical flag names and defaults."""
    parser = parent_args.add_argument_group('Basic Module')
    parser.add_argument('--learning_rate', default=5e-5, type=float)
    parser.add_argument('--min_learning_rate', default=1e-7, type=float)
    parser.add_argument('--lr_decay_steps', default=0, type=int)
    parser.add_argument('--lr_decay_ratio', default=1.0, type=float)
    parser.add_argument('--warmup_steps', default=0, type=int)
    parser.add_argument('--warmup_ratio', default=0.1, type=float)
    parser.add_argument('--weight_decay', default=1e-1, type=float)
    parser.add_argument('--adam_beta1', default=0.9, type=float)
    parser.add_argument('--adam_beta2', default=0.999, type=float)
    parser.add_argument('--adam_epsilon', default=1e-8, type=float)
    parser.add_argument('--model_path', default=None, type=str)
    parser.add_argument('--scheduler_type', default='polynomial', type=str)
    return parent_args


def get_default_update_params(pl_model):
    """model_utils.py:39-47 — grouping BY NAME substrings."""
    no_decay = ['bias', 'LayerNorm.bias', 'LayerNorm.weight', 'layer_norm.', 'layernorm.']
    named = [(n, p) for n, p in pl_model.named_parameters() if p.requires_grad]
    return [
        {'params': [p for n, p in named if not any(nd in n for nd in no_decay)],
         'weight_decay': pl_model.hparams.weight_decay},
        {'params': [p for n, p in named if any(nd in n for nd in no_decay)], 'weight_decay': 0.0},
    ]


def get_scheduler(name, optimizer, num_warmup_steps=None, num_training_steps=None, lr_end=None):
    """model_utils.py:212-254 -> transformers' schedule functions, restated as LambdaLR over fsb200.schedules."""
    base = [g['lr'] for g in optimizer.param_groups]
    if name == 'polynomial':
        if num_warmup_steps is None or num_training_steps is None:
            raise ValueError(f"{name} requires `num_warmup_steps` and `num_training_steps`, please provide them.")
        if not base[0] > lr_end:
            raise ValueError(f"lr_end ({lr_end}) must be smaller than initial lr ({base[0]})")
        return LambdaLR(optimizer, [lambda s, b=b: polynomial_lr(s, b, num_warmup_steps, num_training_steps, lr_end) / b
                                    for b in base])
    if name == 'linear':
        return LambdaLR(optimizer, lambda s: linear_lr(s, 1.0, num_warmup_steps, num_training_steps))
    if name == 'constant':
        return LambdaLR(optimizer, lambda s: 1.0)
    if name == 'constant_with_warmup':
        return LambdaLR(optimizer, lambda s: min(1.0, float(s) / float(max(1.0, num_warmup_steps))))
    raise ValueError(f"scheduler type {name!r} is not implemented in the fsb200 compat layer "
                     "(polynomial, linear, constant, constant_with_warmup are)")


def configure_optimizers(pl_model, model_params=None):
    """model_utils.py:50-98: DeepSpeed strategy -> FusedAdam(adam_w_mode=True) (offload -> DeepSpeedCPUAdam), otherwise
    AdamW; then the step-interval scheduler. Returns ([optimizer], [{"scheduler", "interval": "step", "frequency": 1}])."""
    groups = get_default_update_params(pl_model) if model_params is None else model_params
    hp = pl_model.hparams
    if isinstance(pl_model.trainer.strategy, DeepSpeedStrategy):
        if 'offload_optimizer' in pl_model.trainer.strategy.config['zero_optimization']:
            optimizer = DeepSpeedCPUAdam(groups, adamw_mode=True, lr=hp.learning_rate,
                                         betas=(hp.adam_beta1, hp.adam_beta2), eps=hp.adam_epsilon)
        else:
            optimizer = FusedAdam(groups, adam_w_mode=True, lr=hp.learning_rate,
                                  betas=(hp.adam_beta1, hp.adam_beta2), eps=hp.adam_epsilon)
    else:
        optimizer = torch.optim.AdamW(groups, lr=hp.learning_rate, betas=(hp.adam_beta1, hp.adam_beta2),
                                      eps=hp.adam_epsilon)
    total_steps = hp.lr_decay_ratio * pl_model.total_steps if hp.lr_decay_steps == 0 else hp.lr_decay_steps
    warmup_steps = hp.warmup_ratio * pl_model.total_steps if hp.warmup_steps == 0 else hp.warmup_steps
    scheduler = get_scheduler(name=hp.scheduler_type, optimizer=optimizer, num_warmup_steps=warmup_steps,
                              num_training_steps=total_steps, lr_end=hp.min_learning_rate)
    return [optimizer], [{"scheduler": scheduler, "interval": "step", "frequency": 1}]


def get_total_steps(trainer, hparams):
    """model_utils.py:194-209 (including its integer arithmetic)."""
    train_loader = trainer._data_connector._train_dataloader_source.dataloader()
    if trainer.max_epochs > 0:
        if hasattr(hparams, 'use_mpu') and hparams.use_mpu:
            from fengshen.models.megatron import mpu
            world_size = mpu.get_data_parallel_world_size()
        else:
            world_size = trainer.world_size
        tb_size = hparams.train_batchsize * max(1, world_size)
        ab_size = trainer.accumulate_grad_batches
        total_steps = (len(train_loader.dataset) * trainer.max_epochs // tb_size) // ab_size
    else:
        total_steps = trainer.max_steps
    return total_steps
