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
