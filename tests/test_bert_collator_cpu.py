"""BERT-family sample assembly (`ErLangShenCollator` and the data_utils it calls) — the compat restatement
against golden vectors produced by the UNMODIFIED reference (oracle/make_golden_bert_collator.py; same seeded cases, see
oracle/bert_collator_cases.py). Everything here is integer / string work: the bar is exact equality, including the state the
shared numpy RandomState is left in (every case draws from the generator the previous one advanced)."""
import importlib
import json
import os
import sys
from dataclasses import dataclass

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
for p in ("fengshen-lm_b200", os.path.join("fengshen-lm_b200", "compat")):
    sys.path.insert(0, os.path.join(ROOT, p))

NAMES = ("sentence_split", "sop_utils", "truncate_utils", "token_type_utils", "mask_utils")


@pytest.fixture(scope="module")
def golden():
    with open(os.path.join(ROOT, "tests", "golden", "bert_collator.json"), encoding="utf8") as fh:
        return json.load(fh)


@pytest.fixture(scope="module")
def ours(tmp_path_factory):
    import bert_collator_cases as C
    from transformers import BertTokenizer
    mods = {n: importlib.import_module(f"fengshen.data.data_utils.{n}") for n in NAMES}
    for m in mods.values():
        assert os.path.join("fengshen-lm_b200", "compat") in m.__file__, m.__file__
    d = tmp_path_factory.mktemp("vocab")
    with open(d / "vocab.txt", "w", encoding="utf8") as fh:
        fh.write("\n".join(C.build_vocab()) + "\n")
    tok = BertTokenizer(str(d / "vocab.txt"), do_lower_case=True)
    from fengshen.data.data_utils.collators import ErLangShenCollator
    return json.loads(json.dumps(C.run_cases(mods, ErLangShenCollator, tok), ensure_ascii=False))


@pytest.mark.parametrize("key", ["sentences", "segments", "truncate", "tokentypes"])
def test_helpers_match_reference(golden, ours, key):
    assert ours[key] == golden[key]


def test_masking_matches_reference_draw_for_draw(golden, ours):
    assert len(ours["masks"]) == len(golden["masks"])
    for i, (a, b) in enumerate(zip(ours["masks"], golden["masks"])):
        assert a == b, f"masking case {i} differs"
    # the cases are not vacuous: most of them mask something, and the [MASK] / keep / random split is all exercised
    assert sum(1 for m in golden["masks"] if len(m) == 5 and m[1]) > 80


def test_collator_batches_match_reference(golden, ours):
    assert len(ours["batches"]) == len(golden["batches"])
    for a, b in zip(ours["batches"], golden["batches"]):
        assert a.keys() == b.keys() == {"input_ids", "attention_mask", "token_type_ids", "labels", "next_sentence_label"}
        for k in a:
            assert a[k] == b[k], k
    lab = np.array(golden["batches"][1]["labels"])
    assert (lab != -100).any() and (lab == -100).any()


def test_masking_properties():
    """Size-independent invariants at the full C3 length (512): positions sorted and unique, labels are the original tokens,
    untouched positions unchanged, budget respected, [CLS]/[SEP] never masked."""
    from fengshen.data.data_utils.mask_utils import create_masked_lm_predictions
    rs = np.random.RandomState(0)
    V = 21128
    inv = {i: (f"##p{i}" if i % 3 == 0 else f"w{i}") for i in range(V)}
    ids = list(inv.keys())
    for _ in range(20):
        toks = [101] + rs.randint(1000, V, size=510).tolist() + [102]
        toks[200] = 102
        out, pos, lab, boundary, spans = create_masked_lm_predictions(toks, ids, inv, 0.15, 101, 102, 103, 0.15 * 512, rs)
        assert pos == sorted(set(pos)) and 0 < len(pos) <= int(round(512 * 0.15))
        assert lab == [toks[p] for p in pos]
        assert all(toks[p] not in (101, 102) for p in pos)
        keep = set(range(512)) - set(pos)
        assert all(out[p] == toks[p] for p in keep)
        assert sorted(p for s in spans for p in s.index) == pos
        assert len(boundary) == 512


def test_native_collator_is_bit_identical_to_the_python_one_and_keeps_the_generator_in_step(tmp_path):
    """fsb_bert_collate (C++ inside libfsb200.so, numpy's MT19937 stream and derived draws re-implemented) against the Python
    collator that the goldens above pin to the unmodified reference: same rows, and the SAME generator state afterwards, over many
    seeds, sequence lengths (truncation on / off), probabilities, one-sentence and empty documents, '##' pieces."""
    import torch
    import bert_collator_cases as C
    from transformers import BertTokenizer
    from fengshen.data.data_utils.collators import ErLangShenCollator, FastErLangShenCollator
    with open(tmp_path / "vocab.txt", "w", encoding="utf8") as fh:
        fh.write("\n".join(C.build_vocab()) + "\n")
    tok = BertTokenizer(str(tmp_path / "vocab.txt"), do_lower_case=True)
    docs = [{"text": t} for t in C.TEXTS] + [{"text": C.TEXTS[0] + C.TEXTS[2] + C.TEXTS[5]}, {"text": "   "},
                                             {"text": "unbelievable results。playing games。" * 7}]
    checked = 0
    for seed in range(25):
        for L, prob in ((16, 0.15), (32, 0.15), (64, 0.4), (128, 0.15), (512, 0.15), (24, 0.05)):
            slow = ErLangShenCollator(tokenizer=tok, max_seq_length=L, masked_lm_prob=prob)
            fast = FastErLangShenCollator(tokenizer=tok, max_seq_length=L, masked_lm_prob=prob)
            for c in (slow, fast):
                c.setup()
                c.np_rng = np.random.RandomState(seed)
                c.vocab_id_list = sorted(c.vocab_id_list)
            for _ in range(2):                           # the second batch continues the generator of the first
                a, b = slow(docs), fast(docs)
                assert a.keys() == b.keys()
                for k in a:
                    assert a[k].dtype == b[k].dtype and torch.equal(a[k], b[k]), (seed, L, prob, k)
                sa, sb = slow.np_rng.get_state(), fast.np_rng.get_state()
                assert sa[2] == sb[2] and np.array_equal(sa[1], sb[1]), (seed, L, prob)
                checked += a["input_ids"].shape[0]
    assert checked > 2000
    # the generator really is shared: mixing the two implementations batch by batch stays on the same stream
    slow.np_rng = np.random.RandomState(99); fast.np_rng = np.random.RandomState(99)
    x = slow(docs); fast.np_rng.set_state(slow.np_rng.get_state()); y1 = fast(docs); y2 = slow(docs)
    assert all(torch.equal(y1[k], y2[k]) for k in y1)
