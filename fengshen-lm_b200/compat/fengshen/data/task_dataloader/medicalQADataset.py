"""`GPT2QADataset` / `GPT2QADataModel` — the data side of the Wenzhong-GPT2 recipe (C2), restated from
fengshen/data/task_dataloader/medicalQADataset.py:9-124 (imported by examples/wenzhong_qa/finetune_wenzhong.py:10): one Python
dict literal per line with 'Question' and 'answer'; an item is question + answer tokenised, padded / truncated to max_seq_length,
with the padding positions of `labels` set to -100. The flag names, defaults and constructor are the reference's (they are the
schema its launch scripts pass); lines are parsed with ast.literal_eval where the reference calls eval()."""
import ast
import os

import pytorch_lightning as pl
from torch.utils.data import DataLoader, Dataset
from transformers import AutoTokenizer

# (flag, keyword arguments of add_argument) — medicalQADataset.py:84-92
_FLAGS = (
    ('--data_dir', dict(type=str, required=True)),
    ('--num_workers', dict(default=2, type=int)),
    ('--train_data', dict(default='train.txt', type=str)),
    ('--valid_data', dict(default='valid.txt', type=str)),
    ('--test_data', dict(default='test.txt', type=str)),
    ('--train_batchsize', dict(type=int, required=True)),
    ('--valid_batchsize', dict(type=int, required=True)),
    ('--max_seq_length', dict(default=1024, type=int)),
)


class GPT2QADataset(Dataset):
    def __init__(self, data_path, name, args):
        super().__init__()
        tok = AutoTokenizer.from_pretrained(args.pretrained_model_path)
        if tok.pad_token is None:   # GPT-2 vocabularies have no padding token: the end-of-text token stands in
            tok.add_special_tokens({'pad_token': '<|endoftext|>'})
        self.tokenizer, self.data_type_name, self.max_seq_length = tok, name, args.max_seq_length
        self.data = self.load_data(data_path)

    def load_data(self, data_path):
        with open(data_path, "rt", encoding='utf8') as lines:
            return [self.data_parse(line) for line in lines if line.strip()]

    def data_parse(self, line):
        return ast.literal_eval(line.strip())

    def encode(self, item):
        enc = self.tokenizer(item['Question'] + item['answer'], max_length=self.max_seq_length, padding='max_length',
                             truncation=True, return_tensors='pt')
        ids, mask = enc['input_ids'].squeeze(), enc['attention_mask'].squeeze()
        labels = ids.clone()
        labels[ids == self.tokenizer.pad_token_id] = -100
        return dict(input_ids=ids, attention_mask=mask, labels=labels, question=item['Question'], answer=item['answer'])

    def __getitem__(self, index):
        return self.encode(self.data[index])

    def __len__(self):
        return len(self.data)


class GPT2QADataModel(pl.LightningDataModule):
    @staticmethod
    def add_data_specific_args(parent_args):
        group = parent_args.add_argument_group('GPT2QADataModel')
        for flag, kw in _FLAGS:
            group.add_argument(flag, **kw)
        return parent_args

    def __init__(self, args):
        super().__init__()
        self.args = args
        self.train_batchsize, self.valid_batchsize = args.train_batchsize, args.valid_batchsize
        at = lambda name: os.path.join(args.data_dir, name)
        if not args.do_eval_only:
            self.train_data = GPT2QADataset(at(args.train_data), '训练集', args)
            self.valid_data = GPT2QADataset(at(args.valid_data), '验证集', args)
        self.test_data = GPT2QADataset(at(args.test_data), '测试集', args)

    def _loader(self, ds, bs, shuffle):
        return DataLoader(ds, shuffle=shuffle, batch_size=bs, pin_memory=False, num_workers=self.args.num_workers)

    def train_dataloader(self):
        return self._loader(self.train_data, self.train_batchsize, True)

    def val_dataloader(self):
        return self._loader(self.valid_data, self.valid_batchsize, False)

    def predict_dataloader(self):
        return self._loader(self.test_data, self.valid_batchsize, False)
