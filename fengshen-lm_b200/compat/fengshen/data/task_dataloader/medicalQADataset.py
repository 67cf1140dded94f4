"""`GPT2QADataset` / `GPT2QADataModel` — the data side of the Wenzhong-GPT2 recipe (C2):
fengshen/data/task_dataloader/medicalQADataset.py:9-124, imported by examples/wenzhong_qa/finetune_wenzhong.py:10. One Python
dict literal per line with 'Question' and 'answer'; each item is question+answer tokenised, padded / truncated to
max_seq_length, and the padding positions of `labels` set to -100. Same flags and constructor; lines are parsed with
ast.literal_eval (the reference calls eval() on them)."""
import ast
import os

import pytorch_lightning as pl
from torch.utils.data import DataLoader, Dataset
from transformers import AutoTokenizer


class GPT2QADataset(Dataset):
    def __init__(self, data_path, name, args):
        super().__init__()
        self.tokenizer = AutoTokenizer.from_pretrained(args.pretrained_model_path)
        if self.tokenizer.pad_token is None:
            self.tokenizer.add_special_tokens({'pad_token': '<|endoftext|>'})
        self.data_type_name = name
        self.max_seq_length = args.max_seq_length
        self.data = self.load_data(data_path)

    def __len__(self):
        return len(self.data)

    def __getitem__(self, index):
        return self.encode(self.data[index])

    def load_data(self, data_path):
        with open(data_path, "rt", encoding='utf8') as f:   # streamed; the reference's >5 GB branch differs only in its progress bar
            return [self.data_parse(line) for line in f if line.strip()]

    def data_parse(self, line):
        return ast.literal_eval(line.strip())

    def encode(self, item):
        enc = self.tokenizer(item['Question'] + item['answer'], max_length=self.max_seq_length, padding='max_length',
                             truncation=True, return_tensors='pt')
        ids = enc['input_ids']
        labels = ids.clone().detach()
        labels[ids == self.tokenizer.pad_token_id] = -100
        return {"input_ids": ids.squeeze(), "attention_mask": enc['attention_mask'].squeeze(), "labels": labels.squeeze(),
                "question": item['Question'], "answer": item['answer']}


class GPT2QADataModel(pl.LightningDataModule):
    @staticmethod
    def add_data_specific_args(parent_args):
        parser = parent_args.add_argument_group('GPT2QADataModel')
        parser.add_argument('--data_dir', type=str, required=True)
        parser.add_argument('--num_workers', default=2, type=int)
        parser.add_argument('--train_data', default='train.txt', type=str)
        parser.add_argument('--valid_data', default='valid.txt', type=str)
        parser.add_argument('--test_data', default='test.txt', type=str)
        parser.add_argument('--train_batchsize', type=int, required=True)
        parser.add_argument('--valid_batchsize', type=int, required=True)
        parser.add_argument('--max_seq_length', default=1024, type=int)
        return parent_args

    def __init__(self, args):
        super().__init__()
        self.args = args
        self.train_batchsize = args.train_batchsize
        self.valid_batchsize = args.valid_batchsize
        if not args.do_eval_only:
            self.train_data = GPT2QADataset(os.path.join(args.data_dir, args.train_data), '训练集', args)
            self.valid_data = GPT2QADataset(os.path.join(args.data_dir, args.valid_data), '验证集', args)
        self.test_data = GPT2QADataset(os.path.join(args.data_dir, args.test_data), '测试集', args)

    def _loader(self, ds, bs, shuffle):
        return DataLoader(ds, shuffle=shuffle, batch_size=bs, pin_memory=False, num_workers=self.args.num_workers)

    def train_dataloader(self):
        return self._loader(self.train_data, self.train_batchsize, True)

    def val_dataloader(self):
        return self._loader(self.valid_data, self.valid_batchsize, False)

    def predict_dataloader(self):
        return self._loader(self.test_data, self.valid_batchsize, False)
