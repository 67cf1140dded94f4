"""Dataset assembly helpers of fengshen/data/megatron_dataloader/dataset_utils.py for the BERT-style path: split arithmetic, blend
weights, the cached sentence-span index map, and the sample helpers `bert_dataset.py` imports from here. The segment / masking
functions are the ones of fengshen.data.data_utils (same code in the reference, duplicated there); `create_masked_lm_predictions`
only adds the positional `tokenizer` argument of dataset_utils.py:182-197.

One deliberate difference: the reference's `get_samples_mapping` (dataset_utils.py:731-788) only LOADS
`<prefix>_<name>_indexmap_..._<seed>s.npy` and fails when it is missing (the branch that builds it was removed there, leaving
`helpers.build_mapping` unused); here a missing map is built with `helpers.build_mapping` (rank 0) and saved under that name
first, as upstream Megatron-LM does, then loaded the same way."""
import math
import os
import time

import numpy as np
import torch

from fengshen.data.data_utils.mask_utils import MaskedLmInstance, is_start_piece  # noqa: F401
from fengshen.data.data_utils.mask_utils import create_masked_lm_predictions as _create_masked_lm_predictions
from fengshen.data.data_utils.sop_utils import get_a_and_b_segments  # noqa: F401
from fengshen.data.data_utils.token_type_utils import create_tokens_and_tokentypes  # noqa: F401
from fengshen.data.data_utils.truncate_utils import truncate_segments  # noqa: F401
from fengshen.data.megatron_dataloader.blendable_dataset import BlendableDataset
from fengshen.data.megatron_dataloader.indexed_dataset import make_dataset as make_indexed_dataset
from fengshen.data.megatron_dataloader.utils import print_rank_0

DSET_TYPE_BERT = 'standard_bert'
DSET_TYPE_ICT = 'ict'
DSET_TYPE_T5 = 't5'
DSET_TYPE_BERT_CN_WWM = 'bert_cn_wwm'
DSET_TYPE_BART = 'bart'
DSET_TYPE_COCOLM = 'coco_lm'
DSET_TYPES = [DSET_TYPE_BERT, DSET_TYPE_ICT, DSET_TYPE_T5, DSET_TYPE_BERT_CN_WWM, DSET_TYPE_BART, DSET_TYPE_COCOLM]


def compile_helper():
    """dataset_utils.py:77-87 runs `make` on helpers.cpp; the fsb200 builders are part of the prebuilt libfsb200.so."""
    from fengshen.data.megatron_dataloader import helpers  # noqa: F401  (raises if the library is missing)


def create_masked_lm_predictions(tokens, vocab_id_list, vocab_id_to_token_dict, masked_lm_prob, cls_id, sep_id, mask_id,
                                 max_predictions_per_seq, np_rng, tokenizer=None, max_ngrams=3, do_whole_word_mask=True,
                                 favor_longer_ngram=False, do_permutation=False, geometric_dist=False, masking_style="bert",
                                 zh_tokenizer=None):
    return _create_masked_lm_predictions(tokens, vocab_id_list, vocab_id_to_token_dict, masked_lm_prob, cls_id, sep_id, mask_id,
                                         max_predictions_per_seq, np_rng, max_ngrams=max_ngrams,
                                         do_whole_word_mask=do_whole_word_mask, favor_longer_ngram=favor_longer_ngram,
                                         do_permutation=do_permutation, geometric_dist=geometric_dist,
                                         masking_style=masking_style, zh_tokenizer=zh_tokenizer)


def pad_and_convert_to_numpy(tokens, tokentypes, masked_positions, masked_labels, pad_id, max_seq_length):
    """dataset_utils.py:471-501 -> (tokens, token types, labels, padding mask, loss mask), int64 [max_seq_length] each."""
    n = len(tokens)
    if n > max_seq_length or len(tokentypes) != n or len(masked_positions) != len(masked_labels):
        raise ValueError("pad_and_convert_to_numpy: inconsistent sample")
    tokens_np = np.full(max_seq_length, pad_id, dtype=np.int64)
    tokens_np[:n] = tokens
    types_np = np.full(max_seq_length, pad_id, dtype=np.int64)
    types_np[:n] = tokentypes
    labels_np = np.full(max_seq_length, -100, dtype=np.int64)
    loss_mask_np = np.zeros(max_seq_length, dtype=np.int64)
    pos = np.asarray(masked_positions, dtype=np.int64)
    if pos.size:
        if pos.max() >= n:
            raise ValueError("pad_and_convert_to_numpy: masked position beyond the sample")
        labels_np[pos] = masked_labels
        loss_mask_np[pos] = 1
    return tokens_np, types_np, labels_np, (np.arange(max_seq_length) < n).astype(np.int64), loss_mask_np


def get_datasets_weights_and_num_samples(data_prefix, train_valid_test_num_samples):
    """dataset_utils.py:46-74: ['w1', 'prefix1', 'w2', 'prefix2', ...] -> prefixes, normalised weights, per-dataset sample
    counts padded by 0.5 % so that a blend never runs dry."""
    if len(data_prefix) % 2:
        raise ValueError("data_prefix must alternate weight, prefix")
    weights = [float(w) for w in data_prefix[0::2]]
    prefixes = [p.strip() for p in data_prefix[1::2]]
    total = 0.0
    for w in weights:
        total += w
    if not total > 0.0:
        raise ValueError("weights must sum to a positive number")
    weights = [w / total for w in weights]
    counts = [[int(math.ceil(val * w * 1.005)) for val in train_valid_test_num_samples] for w in weights]
    return prefixes, weights, counts


def get_train_valid_test_split_(splits_string, size):
    """dataset_utils.py:703-728: '949,50,1' or '0.9/0.1' -> four document boundaries summing exactly to `size`."""
    sep = ',' if ',' in splits_string else ('/' if '/' in splits_string else None)
    splits = [float(s) for s in splits_string.split(sep)] if sep else [float(splits_string)]
    splits = (splits + [0.0, 0.0])[:3]
    total = sum(splits)
    if not total > 0.0:
        raise ValueError("splits must sum to a positive number")
    bounds = [0]
    for s in splits:
        bounds.append(bounds[-1] + int(round(s / total * float(size))))
    diff = bounds[-1] - size
    bounds = [bounds[0]] + [b - diff for b in bounds[1:]]
    assert len(bounds) == 4 and bounds[-1] == size
    return bounds


def get_indexed_dataset_(data_prefix, data_impl, skip_warmup):
    t0 = time.time()
    ds = make_indexed_dataset(data_prefix, data_impl, skip_warmup)
    if ds is None:
        raise FileNotFoundError(f"indexed dataset {data_prefix}(.idx|.bin) not found")
    assert ds.sizes.shape[0] == ds.doc_idx[-1]
    print_rank_0(' > finished creating indexed dataset in {:4f} seconds'.format(time.time() - t0))
    print_rank_0('    number of documents: {}'.format(ds.doc_idx.shape[0] - 1))
    print_rank_0('    number of sentences: {}'.format(ds.sizes.shape[0]))
    return ds


def indexmap_filename(data_prefix, name, num_epochs, max_num_samples, max_seq_length, short_seq_prob, seed):
    fn = data_prefix + '_{}_indexmap'.format(name)
    if num_epochs != (np.iinfo(np.int32).max - 1):
        fn += '_{}ep'.format(num_epochs)
    if max_num_samples != (np.iinfo(np.int64).max - 1):
        fn += '_{}mns'.format(max_num_samples)
    return fn + '_{}msl'.format(max_seq_length) + '_{:0.2f}ssp'.format(short_seq_prob) + '_{}s'.format(seed) + '.npy'


def get_samples_mapping(indexed_dataset, data_prefix, num_epochs, max_num_samples, max_seq_length, short_seq_prob, seed, name,
                        binary_head):
    """-> uint32 / uint64 [samples, 3] rows (first sentence, end sentence, target length), memory-mapped from the cache file."""
    if not num_epochs:
        if not max_num_samples:
            raise ValueError("Need to specify either max_num_samples or num_epochs")
        num_epochs = np.iinfo(np.int32).max - 1
    if not max_num_samples:
        max_num_samples = np.iinfo(np.int64).max - 1
    fn = indexmap_filename(data_prefix, name, num_epochs, max_num_samples, max_seq_length, short_seq_prob, seed)
    dist = torch.distributed
    rank = dist.get_rank() if dist.is_available() and dist.is_initialized() else 0
    if rank == 0 and not os.path.isfile(fn):
        from fengshen.data.megatron_dataloader import helpers
        assert indexed_dataset.doc_idx.dtype == np.int64 and indexed_dataset.sizes.dtype == np.int32
        t0 = time.time()
        mapping = helpers.build_mapping(indexed_dataset.doc_idx, indexed_dataset.sizes, num_epochs, max_num_samples,
                                        max_seq_length, short_seq_prob, seed, False, 2 if binary_head else 1)
        np.save(fn, mapping, allow_pickle=True)
        print_rank_0(' > built and saved the index map {} in {:.3f} seconds'.format(fn, time.time() - t0))
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.barrier()
    print_rank_0(' > loading indexed mapping from {}'.format(fn))
    mapping = np.load(fn, allow_pickle=True, mmap_mode='r')
    print_rank_0('    total number of samples: {}'.format(mapping.shape[0]))
    return mapping


def build_train_valid_test_datasets(data_prefix, data_impl, splits_string, train_valid_test_num_samples, max_seq_length,
                                    masked_lm_prob, short_seq_prob, seed, tokenizer, skip_warmup, binary_head=False,
                                    max_seq_length_dec=None, dataset_type='standard_bert', zh_tokenizer=None, span=None):
    """dataset_utils.py:504-564: one prefix -> (train, valid, test); ['w', 'prefix', ...] -> the three blends."""
    def one(prefix, counts):
        return _build_train_valid_test_datasets(prefix, data_impl, splits_string, counts, max_seq_length, masked_lm_prob,
                                                short_seq_prob, seed, skip_warmup, binary_head, max_seq_length_dec, tokenizer,
                                                dataset_type=dataset_type, zh_tokenizer=zh_tokenizer, span=span)
    if len(data_prefix) == 1:
        return one(data_prefix[0], train_valid_test_num_samples)
    prefixes, weights, counts = get_datasets_weights_and_num_samples(data_prefix, train_valid_test_num_samples)
    groups = [[], [], []]
    for prefix, cnt in zip(prefixes, counts):
        for g, ds in zip(groups, one(prefix, cnt)):
            if ds:
                g.append(ds)
    return tuple(BlendableDataset(g, weights) if g else None for g in groups)


def _build_train_valid_test_datasets(data_prefix, data_impl, splits_string, train_valid_test_num_samples, max_seq_length,
                                     masked_lm_prob, short_seq_prob, seed, skip_warmup, binary_head, max_seq_length_dec,
                                     tokenizer, dataset_type='standard_bert', zh_tokenizer=None, span=None):
    if dataset_type not in DSET_TYPES:
        raise ValueError("Invalid dataset_type: ", dataset_type)
    if dataset_type != DSET_TYPE_BERT:   # bert_cn_wwm strips '##' from Chinese pieces through the tokenizer (needs jieba-segmented
        raise NotImplementedError(f"fsb200 compat: dataset_type {dataset_type!r} is outside the BERT pretraining path "   # corpora)
                                  "(standard_bert is provided)")
    from fengshen.data.megatron_dataloader.bert_dataset import BertDataset
    indexed_dataset = get_indexed_dataset_(data_prefix, data_impl, skip_warmup)
    total_docs = indexed_dataset.doc_idx.shape[0] - 1
    splits = get_train_valid_test_split_(splits_string, total_docs)
    whole = indexed_dataset.get_doc_idx()
    out = []
    for i, name in enumerate(('train', 'valid', 'test')):
        ds = None
        if splits[i + 1] > splits[i]:
            print_rank_0('    {}: document indices in [{}, {}) total of {} documents'.format(
                name, splits[i], splits[i + 1], splits[i + 1] - splits[i]))
            indexed_dataset.set_doc_idx(whole[splits[i]:splits[i + 1] + 1])   # the split's view while its map is built
            try:
                ds = BertDataset(name=name, indexed_dataset=indexed_dataset, data_prefix=data_prefix, num_epochs=None,
                                 max_num_samples=train_valid_test_num_samples[i], masked_lm_prob=masked_lm_prob,
                                 max_seq_length=max_seq_length, short_seq_prob=short_seq_prob, seed=seed,
                                 binary_head=binary_head, tokenizer=tokenizer, masking_style='bert')
            finally:
                indexed_dataset.set_doc_idx(whole)
        out.append(ds)
    return tuple(out)
