"""CPU: the T5 span-corruption collator (SURVEY.md §8f rank 3 — the data format on the input side of config 5) is bit-exact with
the UNMODIFIED reference's `UnsuperviseT5DataModel.collate_fn` / `compute_input_and_target_lengths`
(fengshen/data/t5_dataloader/t5_datasets.py:14-58,282-437) under the same numpy RNG stream: golden vectors in
tests/golden/t5_collator.npz, written by oracle/make_golden_t5_collator.py from the reference itself."""
import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
COMPAT = os.path.join(ROOT, "fengshen-lm_b200", "compat")
if COMPAT not in sys.path:
    sys.path.insert(0, COMPAT)

from fengshen.data.t5_dataloader import T5SpanCorruptionCollator, compute_input_and_target_lengths  # noqa: E402
from fengshen.data.t5_dataloader.t5_datasets import random_spans_noise_mask  # noqa: E402

G = np.load(os.path.join(ROOT, "tests", "golden", "t5_collator.npz"))


def test_lengths_match_reference():
    for L in (64, 128, 512):
        assert tuple(G[f"lengths_{L}"]) == compute_input_and_target_lengths(L, 0.15, 3)
    assert tuple(G["lengths_512_half"]) == compute_input_and_target_lengths(512, 0.5, 3)
    assert compute_input_and_target_lengths(512, 0.15, 3) == (568, 114)      # enc 512 / dec 114 by the collator


@pytest.mark.parametrize("name", ["a", "b", "c"])
def test_collator_bit_exact_with_reference_under_the_same_rng(name):
    V, L, B, seed = (int(x) for x in G[f"{name}_cfg"])
    coll = T5SpanCorruptionCollator(vocab_size=V, max_seq_length=L, pad_token_id=0, eos_token_id=1, decoder_start_token_id=0)
    raw = G[f"{name}_raw"]
    assert raw.shape[1] == coll.expanded_inputs_length
    np.random.seed(seed)
    out = coll([{"input_ids": raw[i]} for i in range(B)])
    for k in ("input_ids", "labels", "decoder_input_ids"):
        assert out[k].dtype == torch.int64
        assert np.array_equal(out[k].numpy(), G[f"{name}_{k}"]), k
    assert out["input_ids"].shape == (B, L) and out["labels"].shape == (B, coll.targets_length)


def test_noise_mask_properties_and_ragged_input_is_rejected():
    np.random.seed(0)
    for length in (17, 100, 568):
        m = random_spans_noise_mask(length)
        n_noise = min(max(int(np.round(length * 0.15)), 1), length - 1)
        assert m.shape == (length,) and m.sum() == n_noise and not m[0]          # starts with a non-noise span
        spans = int(np.sum(m[1:] & ~m[:-1]))
        assert spans == max(int(np.round(n_noise / 3.0)), 1)
    coll = T5SpanCorruptionCollator(vocab_size=1000, max_seq_length=64)
    with pytest.raises(ValueError, match="incorrectly preprocessed"):
        coll([{"input_ids": np.arange(2, 2 + 40)}])                               # wrong raw length: loud, like the reference


def test_group_texts_matches_reference():
    """The chunking step in front of the collator (t5_datasets.py:160-177), golden from the unmodified reference function."""
    from fengshen.data.t5_dataloader.t5_datasets import group_texts
    g = np.load(os.path.join(ROOT, "tests", "golden", "t5_collator.npz"))
    for name in ("g1", "g2", "g3"):
        lens, chunk = g[f"{name}_rows"][:-1], int(g[f"{name}_rows"][-1])
        flat = g[f"{name}_flat"].tolist()
        rows, k = [], 0
        for n in lens:
            rows.append(flat[k:k + n]); k += n
        out = group_texts({"input_ids": rows}, chunk)["input_ids"]
        assert [len(c) for c in out] == g[f"{name}_lens"].tolist()
        assert [t for c in out for t in c] == g[f"{name}_out"].tolist()
    assert g["g2_lens"].tolist() == [int(g["g2_rows"][:-1].sum())] and g["g2_lens"][0] < 568    # the short-batch branch
    assert g["g3_lens"].tolist() == [50, 50] and g["g3_rows"][:-1].sum() == 113                 # tail of 13 tokens dropped
