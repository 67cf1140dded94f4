"""`ErLangShenCollator` as a library class. In the reference it lives inside the example script
(fengshen/examples/pretrain_erlangshen_bert/pretrain_erlangshen.py:35-123) and the unmodified script keeps using its own copy on
top of this package's helpers; this one serves fsb200's own entry points (examples/, bench fixtures) that need MegatronBERT
MLM + sentence-order batches without importing an example script."""
import time
from dataclasses import dataclass

import numpy as np
import torch

from .mask_utils import create_masked_lm_predictions
from .sentence_split import ChineseSentenceSplitter
from .sop_utils import get_a_and_b_segments
from .token_type_utils import create_tokens_and_tokentypes
from .truncate_utils import truncate_segments


@dataclass
class ErLangShenCollator:
    tokenizer: object = None
    max_seq_length: int = 512
    masked_lm_prob: float = 0.15
    content_key: str = 'text'

    def setup(self):
        self.sentence_split = ChineseSentenceSplitter()
        self.np_rng = np.random.RandomState(seed=int(time.time()) % 2 ** 32)
        self.vocab_id_to_token_dict = {v: k for k, v in self.tokenizer.vocab.items()}
        self.vocab_id_list = list(self.vocab_id_to_token_dict.keys())

    def encode(self, text):
        """One document -> model inputs, or None when it yields no tokens (the reference skips such samples)."""
        tk, L, rng = self.tokenizer, self.max_seq_length, self.np_rng
        sents = [tk.convert_tokens_to_ids(tk.tokenize(s)) for s in self.sentence_split.tokenize(text)]
        if not sents:
            print('find empty sentence')
            return None
        if len(sents) > 1:
            a, b, swapped = get_a_and_b_segments(sents, rng)
        else:
            a, b, swapped = sents[0], [], False
        if not a:
            return None
        truncate_segments(a, b, len(a), len(b), L - 3, rng)   # room for [CLS] [SEP] [SEP]
        tokens, types = create_tokens_and_tokentypes(a, b, tk.cls_token_id, tk.sep_token_id)
        tokens, positions, targets, _, _ = create_masked_lm_predictions(
            tokens, self.vocab_id_list, self.vocab_id_to_token_dict, self.masked_lm_prob, tk.cls_token_id, tk.sep_token_id,
            tk.mask_token_id, self.masked_lm_prob * len(tokens), rng, masking_style='bert')
        n = len(tokens)
        assert n <= L and len(types) == n and len(positions) == len(targets)
        ids = np.full(L, tk.pad_token_id, dtype=np.int64)
        ids[:n] = tokens
        tt = np.full(L, tk.pad_token_id, dtype=np.int64)   # the reference pads token types with the PAD id as well
        tt[:n] = types
        labels = np.full(L, -100, dtype=np.int64)
        labels[positions] = targets
        return {'input_ids': torch.from_numpy(ids), 'attention_mask': torch.from_numpy((np.arange(L) < n).astype(np.int64)),
                'token_type_ids': torch.from_numpy(tt), 'labels': torch.from_numpy(labels),
                'next_sentence_label': int(swapped)}

    def __call__(self, samples):
        rows = [r for r in (self.encode(s[self.content_key]) for s in samples) if r is not None]
        batch = {k: torch.stack([r[k] for r in rows]) for k in ('input_ids', 'attention_mask', 'token_type_ids', 'labels')}
        batch['next_sentence_label'] = torch.tensor([r['next_sentence_label'] for r in rows], dtype=torch.int64)
        return batch
