"""TEST INFRASTRUCTURE — golden vectors for the BERT-family sample assembly from the UNMODIFIED reference:
fengshen/data/data_utils/{sentence_split,sop_utils,truncate_utils,token_type_utils,mask_utils}.py and the `ErLangShenCollator`
class of fengshen/examples/pretrain_erlangshen_bert/pretrain_erlangshen.py:35-123 (imported from the script file itself).
The reference's `fengshen` package cannot be imported whole, so its five data_utils files are loaded by path and
registered under their real module names BEFORE the script is imported — the collator then runs on reference code only; the
remaining imports of the script (pytorch_lightning, fengshen.data.universal_datamodule, ...) resolve to the compat surface and are
not exercised. Run in the authoring container:  python oracle/make_golden_bert_collator.py"""
import importlib.util
import json
import os
import sys
import tempfile
import types

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
REF = "/root/reference"
NAMES = ("sentence_split", "sop_utils", "truncate_utils", "token_type_utils", "mask_utils")


def _load(name, path):
    spec = importlib.util.spec_from_file_location(name, path)
    mod = importlib.util.module_from_spec(spec)
    sys.modules[name] = mod
    spec.loader.exec_module(mod)
    return mod


def make_tokenizer(vocab_list):
    from transformers import BertTokenizer
    d = tempfile.mkdtemp()
    with open(os.path.join(d, "vocab.txt"), "w", encoding="utf8") as fh:
        fh.write("\n".join(vocab_list) + "\n")
    return BertTokenizer(os.path.join(d, "vocab.txt"), do_lower_case=True)


def main():
    sys.path.insert(0, HERE)
    sys.path.insert(0, os.path.join(ROOT, "fengshen-lm_b200", "compat"))
    sys.path.insert(0, os.path.join(ROOT, "fengshen-lm_b200"))
    import fengshen.data   # compat package object; the reference files are hung below it
    pkg = types.ModuleType("fengshen.data.data_utils")
    pkg.__path__ = []
    sys.modules["fengshen.data.data_utils"] = pkg
    mods = {n: _load(f"fengshen.data.data_utils.{n}", os.path.join(REF, "fengshen/data/data_utils", n + ".py")) for n in NAMES}
    for n, m in mods.items():
        assert m.__file__.startswith(REF), m.__file__
    script = _load("ref_pretrain_erlangshen",
                   os.path.join(REF, "fengshen/examples/pretrain_erlangshen_bert/pretrain_erlangshen.py"))
    import bert_collator_cases as C
    tok = make_tokenizer(C.build_vocab())
    out = C.run_cases(mods, script.ErLangShenCollator, tok)
    path = os.path.join(ROOT, "tests", "golden", "bert_collator.json")
    with open(path, "w", encoding="utf8") as fh:
        json.dump(out, fh, ensure_ascii=False, separators=(",", ":"))
    print({k: len(v) for k, v in out.items()}, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
