"""T5 span-corruption batch assembly — the collate step of `UnsuperviseT5DataModel`
(fengshen/data/t5_dataloader/t5_datasets.py:282-437), i.e. the data format on the input side of BASELINE config 5
(fengshen/examples/pretrain_t5/pretrain_t5.py feeds `input_ids` / `labels` produced here to MT5ForConditionalGeneration).

Behaviour restated (integer in, integer out — bit-exact against the reference's own functions under the same numpy RNG
stream, tests/test_t5_collator_cpu.py + tests/golden/t5_collator.npz):
  * `compute_input_and_target_lengths(L, density, mean_span)`: the raw token count T whose corrupted form has exactly L input
    tokens (non-noise tokens + one sentinel per noise span + EOS), and the target length (noise tokens + sentinels + EOS);
  * per example, a noise mask of `round(T * density)` tokens in `round(noise / mean_span)` spans, alternating non-noise / noise
    starting with non-noise; span lengths come from two `np.random.shuffle` calls on "cut here" indicator vectors (the RNG
    consumption the reference has, so a seeded run draws the same masks);
  * inputs: every noise span collapses to ONE sentinel id (vocab_size - 1, vocab_size - 2, ... in order of appearance), EOS appended;
    labels: the complement (every non-noise span collapses to its sentinel), EOS appended; decoder_input_ids = labels shifted right
    behind the decoder start id.
Implementation: span bookkeeping with np.repeat / np.diff per example instead of the reference's cumsum / roll / where algebra over
the whole batch — same outputs, and what a C++ / GPU batch assembler would do per sequence."""
import numpy as np
import pytorch_lightning as pl
import torch


def _lengths_for(tokens, density, mean_span):
    noise = int(round(tokens * density))
    spans = int(round(noise / mean_span))
    return (tokens - noise) + spans + 1, noise + spans + 1       # (input length, target length), EOS included


def compute_input_and_target_lengths(inputs_length, noise_density, mean_noise_span_length):
    """-> (tokens_length, targets_length): the largest raw length whose corrupted inputs still fit `inputs_length`."""
    tokens = inputs_length
    while _lengths_for(tokens + 1, noise_density, mean_noise_span_length)[0] <= inputs_length:
        tokens += 1
    _, targets = _lengths_for(tokens, noise_density, mean_noise_span_length)
    if noise_density == 0.5 and targets > inputs_length:        # keep targets <= inputs at the symmetric setting
        tokens -= 1
        targets -= 1
    return tokens, targets


def _random_segment_lengths(n_items, n_segments):
    """n_segments positive integers summing to n_items, every composition equally likely. ONE np.random.shuffle over the
    n_items - 1 possible cut positions (the reference's RNG consumption)."""
    cuts = np.arange(n_items - 1) < (n_segments - 1)
    np.random.shuffle(cuts)
    bounds = np.flatnonzero(cuts) + 1
    return np.diff(np.concatenate(([0], bounds, [n_items])))


def random_spans_noise_mask(length, noise_density=0.15, mean_noise_span_length=3.0):
    """bool [length]: True on the tokens to corrupt."""
    n_noise = int(np.round(length * noise_density))
    n_noise = min(max(n_noise, 1), length - 1)
    n_spans = max(int(np.round(n_noise / mean_noise_span_length)), 1)
    noise_len = _random_segment_lengths(n_noise, n_spans)
    keep_len = _random_segment_lengths(length - n_noise, n_spans)
    lengths = np.stack([keep_len, noise_len], axis=1).reshape(-1)            # keep, noise, keep, noise, ...
    return np.repeat(np.tile(np.array([False, True]), n_spans), lengths)[:length]


def _collapse(tokens, drop, vocab_size, eos_id):
    """Replace every maximal run of `drop` positions by one sentinel (vocab_size - 1, - 2, ...), keep the rest, append EOS."""
    out, k, prev = [], 0, False
    for tok, d in zip(tokens.tolist(), drop.tolist()):
        if d:
            if not prev:
                k += 1
                out.append(vocab_size - k)
        else:
            out.append(tok)
        prev = d
    out.append(eos_id)
    return out


class T5SpanCorruptionCollator:
    """collate_fn of the reference's UnsuperviseT5DataModel: examples are dicts with `input_ids` of the expanded length
    (`compute_input_and_target_lengths(max_seq_length, ...)[0]`); returns input_ids [B, max_seq_length], labels [B, targets_length],
    decoder_input_ids, plus any other keys of the examples stacked."""

    def __init__(self, vocab_size, max_seq_length, pad_token_id=0, eos_token_id=1, decoder_start_token_id=0,
                 noise_density=0.15, mean_noise_span_length=3):
        self.vocab_size, self.max_seq_length = vocab_size, max_seq_length
        self.pad_token_id, self.eos_token_id, self.decoder_start_token_id = pad_token_id, eos_token_id, decoder_start_token_id
        self.noise_density, self.mean_noise_span_length = noise_density, mean_noise_span_length
        self.expanded_inputs_length, self.targets_length = compute_input_and_target_lengths(
            max_seq_length, noise_density, mean_noise_span_length)

    def __call__(self, examples):
        batch = {k: np.array([ex[k] for ex in examples]) for k in examples[0]}
        ids = np.asarray(batch["input_ids"])
        masks = [random_spans_noise_mask(ids.shape[1], self.noise_density, self.mean_noise_span_length) for _ in range(ids.shape[0])]
        inputs = [_collapse(row, m, self.vocab_size, self.eos_token_id) for row, m in zip(ids, masks)]
        labels = [_collapse(row, ~m, self.vocab_size, self.eos_token_id) for row, m in zip(ids, masks)]
        if any(len(r) != self.max_seq_length for r in inputs):
            raise ValueError(f"`input_ids` are incorrectly preprocessed: corrupted length {len(inputs[0])}, expected "
                             f"{self.max_seq_length} (examples must be {self.expanded_inputs_length} tokens long)")
        if any(len(r) != self.targets_length for r in labels):
            raise ValueError(f"`labels` are incorrectly preprocessed: length {len(labels[0])}, expected {self.targets_length}")
        batch["input_ids"] = np.array(inputs, dtype=np.int64)
        batch["labels"] = np.array(labels, dtype=np.int64)
        dec = np.zeros_like(batch["labels"])
        dec[:, 1:] = batch["labels"][:, :-1]
        dec[:, 0] = self.decoder_start_token_id
        batch["decoder_input_ids"] = np.where(dec == -100, self.pad_token_id, dec)
        return {k: torch.tensor(v) for k, v in batch.items()}


# ---- dataset + data module around the collator (t5_datasets.py:60-280) --------------------------------------------------
def _load_local(path, num_proc=None):
    """The reference reads named IDEA corpora through its private `data.fs_datasets.load_dataset`; here a path is a directory
    written by `datasets.save_to_disk` (the alternative the reference keeps in a comment, t5_datasets.py:113) or a JSON-lines
    file. Returns a mapping with at least 'train'."""
    import os
    import datasets
    if os.path.isdir(path):
        ds = datasets.load_from_disk(path)
        return ds if isinstance(ds, (dict, datasets.DatasetDict)) else {"train": ds}
    if os.path.isfile(path):
        return datasets.load_dataset("json", data_files={"train": path})
    raise FileNotFoundError(f"fsb200 compat: {path!r} is neither a datasets.save_to_disk directory nor a JSON-lines file "
                            "(IDEA's named corpora need the private fs_datasets loader and are out of scope)")


def group_texts(examples, chunk_length):
    """t5_datasets.py:160-177: concatenate every column of a tokenised batch and cut it into chunks of `chunk_length`,
    dropping the tail (when the batch holds less than one chunk the single short piece is kept, as in the reference)."""
    out = {}
    for key, rows in examples.items():
        flat = [tok for row in rows for tok in row]
        total = len(flat)
        if total >= chunk_length:
            total = (total // chunk_length) * chunk_length
        out[key] = [flat[i:i + chunk_length] for i in range(0, total, chunk_length)]
    return out


class UnsuperviseT5Dataset(torch.utils.data.Dataset):
    """load_data_type 0: raw text at `data_path` -> tokenise (no special tokens) -> group into `expanded_inputs_length` chunks;
    1: already tokenised data at `data_path`; 2: already tokenised data handed in as `data`."""

    def __init__(self, data_path, args, load_data_type=0, data=None, tokenizer=None):
        super().__init__()
        self.text_column_name = args.text_column_name
        self.dataset_num_workers = getattr(args, "dataset_num_workers", None)
        self.max_seq_length = args.max_seq_length
        self.remove_columns = list(getattr(args, "remove_columns", []) or [])
        self.noise_density, self.mean_noise_span_length = 0.15, 3
        self.load_data_type = load_data_type
        self.tokenizer = tokenizer
        if load_data_type == 0:
            if self.tokenizer is None:
                from transformers import BertTokenizer, T5Tokenizer
                cls = T5Tokenizer if args.tokenizer_type == 't5_tokenizer' else BertTokenizer
                self.tokenizer = cls.from_pretrained(getattr(args, "new_vocab_path", None) or args.pretrained_model_path)
            self.expanded_inputs_length, self.targets_length = compute_input_and_target_lengths(
                self.max_seq_length, self.noise_density, self.mean_noise_span_length)
            self.data = self.load_data(data_path)
        elif load_data_type == 1:
            self.data = _load_local(data_path)['train']
        else:
            if data is None:
                raise ValueError("UnsuperviseT5Dataset(load_data_type=2) needs `data`")
            self.data = data

    def __len__(self):
        return len(self.data)

    def __getitem__(self, index):
        return self.data[index]

    def tokenize_function(self, examples):
        # add_special_tokens=False: no EOS in the middle of a chunk
        return self.tokenizer(examples[self.text_column_name], add_special_tokens=False, return_attention_mask=False)

    def group_texts(self, examples):
        return group_texts(examples, self.expanded_inputs_length)

    def load_data(self, data_path):
        samples = _load_local(data_path)['train']
        drop = [c for c in samples.column_names if c == self.text_column_name or c in self.remove_columns]
        tokenised = samples.map(self.tokenize_function, batched=True, remove_columns=drop)
        keep = [c for c in tokenised.column_names if c != "input_ids"]
        return tokenised.map(self.group_texts, batched=True, remove_columns=keep)


class UnsuperviseT5DataModel(pl.LightningDataModule):
    _FLAGS = (   # t5_datasets.py:184-196
        ('--dataset_num_workers', dict(default=8, type=int)),
        ('--dataloader_num_workers', dict(default=4, type=int)),
        ('--train_data_path', dict(default='wudao_180g_mt5_tokenized', type=str)),
        ('--train_batchsize', dict(default=2, type=int)),
        ('--valid_batchsize', dict(default=2, type=int)),
        ('--train_split_size', dict(default=None, type=float)),
        ('--tokenizer_type', dict(default='t5_tokenizer', choices=['t5_tokenizer', 'bert_tokenizer'])),
        ('--text_column_name', dict(default='text')),
        ('--remove_columns', dict(nargs='+', default=[])),
    )

    @staticmethod
    def add_data_specific_args(parent_args):
        group = parent_args.add_argument_group('UnsuperviseT5DataModel')
        for flag, kw in UnsuperviseT5DataModel._FLAGS:
            group.add_argument(flag, **kw)
        return parent_args

    def __init__(self, args):
        super().__init__()
        self.save_hyperparameters(args)
        from transformers import MT5Config
        if args.train_split_size is not None:
            splits = _load_local(args.train_data_path)
            self.train_dataset = UnsuperviseT5Dataset('', args, load_data_type=2, data=splits['train'])
            self.test_dataset = UnsuperviseT5Dataset('', args, load_data_type=2, data=splits['test'])
        else:   # the reference stores this one under a name its own train_dataloader never reads; both names are set here
            self.train_dataset = UnsuperviseT5Dataset(args.train_data_path, args, load_data_type=1)
            self.test_dataset = self.train_dataset
        self.train_data = self.test_data = self.train_dataset
        self.config = MT5Config.from_pretrained(args.pretrained_model_path)
        self.max_seq_length = args.max_seq_length
        self.collator = T5SpanCorruptionCollator(
            vocab_size=self.config.vocab_size, max_seq_length=args.max_seq_length, pad_token_id=self.config.pad_token_id,
            eos_token_id=self.config.eos_token_id, decoder_start_token_id=self.config.decoder_start_token_id)
        self.expanded_inputs_length, self.targets_length = self.collator.expanded_inputs_length, self.collator.targets_length

    def collate_fn(self, examples):
        return self.collator([{"input_ids": ex["input_ids"]} for ex in examples])

    def train_dataloader(self):
        """Megatron sampler so that a resumed run continues the sample stream (t5_datasets.py:233-250)."""
        from fengshen.data.universal_datamodule.universal_datamodule import get_consume_samples
        from fengshen.data.universal_datamodule.universal_sampler import PretrainingSampler
        sampler = PretrainingSampler(total_samples=len(self.train_dataset), consumed_samples=get_consume_samples(self),
                                     micro_batch_size=self.hparams.train_batchsize,
                                     data_parallel_rank=self.trainer.global_rank, data_parallel_size=self.trainer.world_size)
        return torch.utils.data.DataLoader(self.train_dataset, batch_sampler=sampler, pin_memory=True,
                                           num_workers=self.hparams.dataloader_num_workers, collate_fn=self.collate_fn)

    def _eval_loader(self, ds):
        sampler = None
        if self.trainer is not None and self.trainer.world_size > 1:
            sampler = torch.utils.data.distributed.DistributedSampler(ds, num_replicas=self.trainer.world_size,
                                                                      rank=self.trainer.global_rank, shuffle=False)
        return torch.utils.data.DataLoader(ds, sampler=sampler, shuffle=False, batch_size=self.hparams.valid_batchsize,
                                           pin_memory=True, num_workers=self.hparams.dataloader_num_workers,
                                           collate_fn=self.collate_fn)

    def val_dataloader(self):
        return self._eval_loader(self.test_dataset)

    def predict_dataloader(self):
        return self._eval_loader(self.test_dataset)
