"""UniversalCheckpoint — the argparse flags and the constructor pass-through of fengshen/utils/universal_checkpoint.py:5-41,
restated: these flags ARE the schema the example scripts and their launch files use."""
import os

from pytorch_lightning.callbacks import ModelCheckpoint

# (flag, add_argument keywords) — universal_checkpoint.py:9-22
_FLAGS = (
    ('--monitor', dict(default='step', type=str)),
    ('--mode', dict(default='max', type=str)),
    ('--save_ckpt_path', dict(default='./ckpt/', type=str)),
    ('--load_ckpt_path', dict(default='./ckpt/', type=str)),
    ('--filename', dict(default='model-ep{epoch:02d}-st{step:d}', type=str)),
    ('--save_last', dict(action='store_true', default=False)),
    ('--save_top_k', dict(default=10, type=float)),
    ('--every_n_train_steps', dict(default=None, type=float)),
    ('--save_weights_only', dict(action='store_true', default=False)),
    ('--every_n_epochs', dict(default=None, type=int)),
    ('--save_on_train_epoch_end', dict(action='store_true', default=None)),
)
_PASS_THROUGH = ('monitor', 'save_top_k', 'mode', 'every_n_train_steps', 'save_weights_only', 'filename', 'save_last',
                 'every_n_epochs', 'save_on_train_epoch_end')


class UniversalCheckpoint(ModelCheckpoint):
    @staticmethod
    def add_argparse_args(parent_args):
        group = parent_args.add_argument_group('universal checkpoint callback')
        for flag, kw in _FLAGS:
            group.add_argument(flag, **kw)
        return parent_args

    def __init__(self, args):
        super().__init__(dirpath=args.save_ckpt_path, **{k: getattr(args, k) for k in _PASS_THROUGH})
        if args.load_ckpt_path is not None and not os.path.exists(args.load_ckpt_path):   # universal_checkpoint.py:37-41
            print('--------warning no checkpoint found--------, remove args')
            args.load_ckpt_path = None
