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
