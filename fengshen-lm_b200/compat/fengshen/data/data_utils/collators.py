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


class FastErLangShenCollator(ErLangShenCollator):
    """Same batches as `ErLangShenCollator`, with everything after tokenisation done by one C call (`fsb_bert_collate` in
    libfsb200.so: segments, truncation, token types, whole-word n-gram masking, padding) on the collator's own numpy generator:
    the rows AND the state `np_rng` is left in are bit-identical to the Python path (tests/test_bert_collator_cpu.py), so the
    two can be swapped mid-run. Use it when the Python collator would limit the step (about 50 k tokens/s per core)."""

    def setup(self):
        super().setup()
        import ctypes
        from fsb200 import lib as L
        self._L, self._ct = L, ctypes
        L.load()
        table = np.zeros(max(self.vocab_id_to_token_dict) + 1, dtype=np.uint8)
        for i, piece in self.vocab_id_to_token_dict.items():
            table[i] = piece.startswith("##")
        self._continuation = table
        # numpy's choice(p=...) arithmetic for n-gram sizes 1..3 (mask_utils.py:137-143 and :165-167), done by numpy itself
        p = 1. / np.arange(1, 4)
        p /= p.sum(keepdims=True)
        p = p / p.sum(keepdims=True)
        cdf = p.cumsum()
        cdf /= cdf[-1]
        self._ngram_cdf = np.ascontiguousarray(cdf, dtype=np.float64)

    def _ragged(self, samples):
        tk = self.tokenizer
        flat, sent_off, doc_off = [], [0], [0]
        for s in samples:
            for sent in self.sentence_split.tokenize(s[self.content_key]):
                flat.extend(tk.convert_tokens_to_ids(tk.tokenize(sent)))
                sent_off.append(len(flat))
            doc_off.append(len(sent_off) - 1)
        return (np.asarray(flat, dtype=np.int32), np.asarray(sent_off, dtype=np.int64), np.asarray(doc_off, dtype=np.int64))

    def __call__(self, samples):
        tk, L_, ct = self.tokenizer, self.max_seq_length, self._ct
        tokens, sent_off, doc_off = self._ragged(samples)
        n_docs = len(samples)
        vocab_ids = np.asarray(self.vocab_id_list, dtype=np.int32)
        name, key, pos, has_gauss, cached = self.np_rng.get_state()
        key = np.ascontiguousarray(key, dtype=np.uint32).copy()
        pos_c = ct.c_int32(int(pos))
        out = [np.empty((n_docs, L_), dtype=np.int64) for _ in range(4)]
        nsl = np.empty(n_docs, dtype=np.int64)
        ptr = lambda a: ct.c_void_p(a.ctypes.data)
        rows = self._L.load().fsb_bert_collate(
            ptr(tokens), ptr(sent_off), ptr(doc_off), n_docs, ptr(self._continuation), self._continuation.shape[0],
            ptr(vocab_ids), vocab_ids.shape[0], tk.cls_token_id, tk.sep_token_id, tk.mask_token_id, tk.pad_token_id, L_,
            float(self.masked_lm_prob), ptr(self._ngram_cdf), 3, ptr(key), ct.byref(pos_c), *(ptr(a) for a in out), ptr(nsl))
        if rows < 0:
            raise RuntimeError(f"fsb200: fsb_bert_collate failed: {self._L.last_error()}")
        self.np_rng.set_state((name, key, int(pos_c.value), has_gauss, cached))
        ids, am, tt, lab = (torch.from_numpy(a[:rows]) for a in out)
        return {'input_ids': ids, 'attention_mask': am, 'token_type_ids': tt, 'labels': lab,
                'next_sentence_label': torch.from_numpy(nsl[:rows])}
