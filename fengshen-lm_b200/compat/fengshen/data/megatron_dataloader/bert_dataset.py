"""`BertDataset` — fengshen/data/megatron_dataloader/bert_dataset.py:29-196: item i is the sentence span i of the index map,
split into segments A / B (sentence-order label when `binary_head`), wrapped in [CLS] / [SEP], whole-word n-gram masked and padded;
the numpy RandomState is seeded with (seed + i) so a sample does not depend on the order of access."""
import numpy as np
import torch

from fengshen.data.megatron_dataloader.dataset_utils import (
    create_masked_lm_predictions,
    create_tokens_and_tokentypes,
    get_a_and_b_segments,
    get_samples_mapping,
    pad_and_convert_to_numpy,
)


class BertDataset(torch.utils.data.Dataset):
    def __init__(self, name, indexed_dataset, data_prefix, num_epochs, max_num_samples, masked_lm_prob, max_seq_length,
                 short_seq_prob, seed, binary_head, tokenizer, masking_style):
        self.name, self.seed = name, seed
        self.masked_lm_prob, self.max_seq_length, self.short_seq_prob = masked_lm_prob, max_seq_length, short_seq_prob
        self.binary_head, self.masking_style = binary_head, masking_style
        self.indexed_dataset = indexed_dataset
        # three positions are reserved for [CLS] [SEP] [SEP]
        self.samples_mapping = get_samples_mapping(indexed_dataset, data_prefix, num_epochs, max_num_samples,
                                                   max_seq_length - 3, short_seq_prob, seed, name, binary_head)
        self.vocab_id_to_token_dict = {v: k for k, v in tokenizer.vocab.items()}
        self.vocab_id_list = list(self.vocab_id_to_token_dict.keys())
        self.cls_id, self.sep_id = tokenizer.cls_token_id, tokenizer.sep_token_id
        self.mask_id, self.pad_id = tokenizer.mask_token_id, tokenizer.pad_token_id
        self.tokenizer = tokenizer

    def __len__(self):
        return self.samples_mapping.shape[0]

    def __getitem__(self, idx):
        first, end, target_len = (int(v) for v in self.samples_mapping[idx])
        sample = [self.indexed_dataset[i] for i in range(first, end)]
        np_rng = np.random.RandomState(seed=((self.seed + idx) % 2 ** 32))
        return build_training_sample(sample, target_len, self.max_seq_length, self.vocab_id_list, self.vocab_id_to_token_dict,
                                     self.cls_id, self.sep_id, self.mask_id, self.pad_id, self.masked_lm_prob, np_rng,
                                     self.binary_head, tokenizer=self.tokenizer, masking_style=self.masking_style)


def build_training_sample(sample, target_seq_length, max_seq_length, vocab_id_list, vocab_id_to_token_dict, cls_id, sep_id, mask_id,
                          pad_id, masked_lm_prob, np_rng, binary_head, tokenizer, masking_style='bert'):
    """bert_dataset.py:81-160. Segment A is clipped to max_seq_length - 3 tokens; as in the reference there is NO pair truncation
    to the target length (its truncate_segments call is commented out), so the index map's spans must fit by construction."""
    if binary_head and len(sample) <= 1:
        raise ValueError("build_training_sample: the sentence-order head needs at least two sentences per sample")
    if target_seq_length > max_seq_length:
        raise ValueError("build_training_sample: target length beyond max_seq_length")
    if binary_head:
        tokens_a, tokens_b, swapped = get_a_and_b_segments(sample, np_rng)
    else:
        tokens_a, tokens_b, swapped = [t for sent in sample for t in sent], [], False
    tokens_a = tokens_a[:max_seq_length - 3] if len(tokens_a) >= max_seq_length - 3 else tokens_a
    tokens, tokentypes = create_tokens_and_tokentypes(tokens_a, tokens_b, cls_id, sep_id)
    tokens, positions, labels, _, _ = create_masked_lm_predictions(
        tokens, vocab_id_list, vocab_id_to_token_dict, masked_lm_prob, cls_id, sep_id, mask_id,
        masked_lm_prob * target_seq_length, np_rng, tokenizer=tokenizer, masking_style=masking_style)
    tokens_np, types_np, labels_np, padding_mask_np, _ = pad_and_convert_to_numpy(tokens, tokentypes, positions, labels, pad_id,
                                                                                   max_seq_length)
    return {'input_ids': tokens_np, 'token_type_ids': types_np, 'labels': labels_np,
            'next_sentence_label': int(swapped), 'attention_mask': padding_mask_np}
