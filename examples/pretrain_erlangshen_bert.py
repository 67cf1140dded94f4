"""Erlangshen-MegatronBERT MLM + sentence-order pretraining, written against the REFERENCE's import surface only
(transformers class names, pytorch_lightning, fengshen.*) — the structure of
fengshen/examples/pretrain_erlangshen_bert/pretrain_erlangshen.py:126-240 with the collator taken from the library
(fengshen.data.data_utils.collators) instead of being defined in the script, and the corpus read from a JSON-lines file
({"text": ...} per line) handed to UniversalDataModule through its `datasets=` argument. Run it through the launcher, which puts
the compat packages on sys.path and binds the transformers class names to the fsb200-backed classes:

  PYTHONPATH=fengshen-lm_b200 python -m fsb200.launch examples/pretrain_erlangshen_bert.py \
      --model_path <dir with config.json + vocab.txt> --train_file corpus.json --train_batchsize 32 --max_seq_length 512 \
      --max_epochs 1 --strategy deepspeed_stage_1 --replace_sampler_ddp False --learning_rate 1e-4

The recipe's DeepSpeed JSON (ZeRO stage, gradient_clipping) is read from $PL_DEEPSPEED_CONFIG_PATH as in the reference
(pretrain_erlangshen_base.sh:24-39). Dropout must be 0 in config.json: fsb200 rejects dropout > 0 loudly.
"""
import argparse
import json
import os

import torch
from transformers import (
    MegatronBertConfig,
    MegatronBertForPreTraining,
    AutoTokenizer,
)
from pytorch_lightning import (
    LightningModule,
    Trainer,
)
from pytorch_lightning.callbacks import LearningRateMonitor
from pytorch_lightning.loggers import TensorBoardLogger
from fengshen.data.universal_datamodule import UniversalDataModule
from fengshen.data.data_utils.collators import ErLangShenCollator
from fengshen.models.model_utils import (
    add_module_args,
    configure_optimizers,
    get_total_steps,
)
from fengshen.utils.universal_checkpoint import UniversalCheckpoint


class JsonLines(torch.utils.data.Dataset):
    def __init__(self, path):
        with open(path, encoding="utf8") as f:
            self.rows = [json.loads(line) for line in f if line.strip()]

    def __len__(self):
        return len(self.rows)

    def __getitem__(self, i):
        return self.rows[i]


class ErLangShenBert(LightningModule):
    @staticmethod
    def add_module_specific_args(parent_parser):
        parser = parent_parser.add_argument_group('Erlangshen Bert')
        parser.add_argument('--masked_lm_prob', type=float, default=0.15)
        parser.add_argument('--max_seq_length', type=int, default=512)
        parser.add_argument('--sample_content_key', type=str, default='text')
        return parent_parser

    def __init__(self, args, tokenizer, **kwargs) -> None:
        super().__init__()
        self.save_hyperparameters(args)
        config = MegatronBertConfig.from_pretrained(args.model_path)
        self.config = config
        self.tokenizer = tokenizer
        self.model = MegatronBertForPreTraining(config)

    def setup(self, stage) -> None:
        if stage == 'fit':
            self.total_steps = get_total_steps(self.trainer, self.hparams)
            print('Total steps: {}'.format(self.total_steps))

    def configure_optimizers(self):
        return configure_optimizers(self)

    def forward(self, **batch):
        return self.model(**batch)

    def comput_metrix(self, logits, labels):
        y_pred = torch.argmax(logits, dim=-1).view(size=(-1,))
        y_true = labels.view(size=(-1,)).float()
        return torch.sum(torch.eq(y_pred, y_true).float()) / labels.shape[0]

    def training_step(self, batch, batch_idx):
        output = self(**batch)
        self.log('train_loss', output.loss, sync_dist=True)
        label_idx = batch['labels'] != -100
        acc = self.comput_metrix(
            output.prediction_logits[label_idx].view(-1, output.prediction_logits.size(-1)), batch['labels'][label_idx])
        self.log('train_acc', acc, sync_dist=True)
        return output.loss

    def on_load_checkpoint(self, checkpoint) -> None:
        global_step_offset = checkpoint["global_step"]
        if 'global_samples' in checkpoint:
            self.consumed_samples = checkpoint['global_samples']
        self.trainer.fit_loop.epoch_loop._batches_that_stepped = global_step_offset


def main(argv=None):
    args_parser = argparse.ArgumentParser()
    args_parser = add_module_args(args_parser)
    args_parser = UniversalDataModule.add_data_specific_args(args_parser)
    args_parser = Trainer.add_argparse_args(args_parser)
    args_parser = ErLangShenBert.add_module_specific_args(args_parser)
    args_parser = UniversalCheckpoint.add_argparse_args(args_parser)
    args = args_parser.parse_args(argv)

    tokenizer = AutoTokenizer.from_pretrained(args.model_path)
    collate_fn = ErLangShenCollator(
        tokenizer=tokenizer,
        max_seq_length=args.max_seq_length,
        masked_lm_prob=args.masked_lm_prob,
        content_key=args.sample_content_key,
    )
    collate_fn.setup()
    data_module = UniversalDataModule(tokenizer=tokenizer, args=args, collate_fn=collate_fn,
                                      datasets={args.train_datasets_field: JsonLines(args.train_file)})
    print('data load complete')

    model = ErLangShenBert(args, tokenizer=tokenizer)
    print('model load complete')

    lr_monitor = LearningRateMonitor(logging_interval='step')
    checkpoint_callback = UniversalCheckpoint(args)
    if args.load_ckpt_path is not None and not os.path.exists(args.load_ckpt_path):
        print('--------warning no checkpoint found--------, remove args')
        args.load_ckpt_path = None
    logger = TensorBoardLogger(save_dir=os.path.join(args.default_root_dir or ".", 'logs'), name='erlangshen')
    trainer = Trainer.from_argparse_args(args, logger=logger, callbacks=[lr_monitor, checkpoint_callback])
    trainer.fit(model, data_module, ckpt_path=args.load_ckpt_path)
    return trainer, model


if __name__ == '__main__':
    main()
