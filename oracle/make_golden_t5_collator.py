"""TEST INFRASTRUCTURE — golden vectors for the T5 span-corruption collator from the UNMODIFIED reference
(/root/reference/fengshen/data/t5_dataloader/t5_datasets.py: compute_input_and_target_lengths, UnsuperviseT5DataModel.collate_fn /
random_spans_noise_mask). The data module's __init__ needs corpora and a checkpoint directory, so the object is created without
it and given exactly the attributes collate_fn reads. Run in the authoring container:  python oracle/make_golden_t5_collator.py"""
import os
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
REF = "/root/reference"
sys.path.insert(0, os.path.join(ROOT, "fengshen-lm_b200", "compat"))     # pytorch_lightning shim for the module's imports
sys.path.insert(0, os.path.join(ROOT, "fengshen-lm_b200"))


def load_reference_module():
    import importlib.util
    import transformers
    if not hasattr(transformers, "MT5Tokenizer"):   # removed in transformers 5.x; the collate path never touches it
        transformers.MT5Tokenizer = type("MT5Tokenizer", (), {})
    spec = importlib.util.spec_from_file_location("ref_t5_datasets", os.path.join(REF, "fengshen/data/t5_dataloader/t5_datasets.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def main():
    mod = load_reference_module()
    rec = {}
    for L in (64, 128, 512):
        rec[f"lengths_{L}"] = np.array(mod.compute_input_and_target_lengths(L, 0.15, 3), dtype=np.int64)
    rec["lengths_512_half"] = np.array(mod.compute_input_and_target_lengths(512, 0.5, 3), dtype=np.int64)
    for name, (V, L, B, seed) in {"a": (32598, 128, 4, 11), "b": (1000, 64, 3, 5), "c": (32598, 512, 2, 99)}.items():
        exp_len, tgt_len = mod.compute_input_and_target_lengths(L, 0.15, 3)
        dm = object.__new__(mod.UnsuperviseT5DataModel)
        dm.noise_density, dm.mean_noise_span_length = 0.15, 3
        dm.pad_token_id, dm.decoder_start_token_id, dm.eos_token_id, dm.vocab_size = 0, 0, 1, V
        dm.max_seq_length, dm.targets_length = L, tgt_len
        rs = np.random.RandomState(seed)
        raw = rs.randint(2, V - 200, size=(B, exp_len)).astype(np.int64)
        np.random.seed(seed)
        out = dm.collate_fn([{"input_ids": raw[i]} for i in range(B)])
        rec[f"{name}_cfg"] = np.array([V, L, B, seed], dtype=np.int64)
        rec[f"{name}_raw"] = raw
        for k in ("input_ids", "labels", "decoder_input_ids"):
            rec[f"{name}_{k}"] = out[k].numpy().astype(np.int64)
    # group_texts (t5_datasets.py:160-177) called unbound on a stand-in `self` that carries the one attribute it reads
    from types import SimpleNamespace
    rs = np.random.RandomState(3)
    for name, (chunk, n_rows) in {"g1": (16, 40), "g2": (568, 9), "g3": (50, 2)}.items():
        rows = [rs.randint(0, 1000, size=int(rs.randint(0, 60))).tolist() for _ in range(n_rows)]
        got = mod.UnsuperviseT5Dataset.group_texts(SimpleNamespace(expanded_inputs_length=chunk), {"input_ids": rows})
        rec[f"{name}_rows"] = np.array([len(r) for r in rows] + [chunk], dtype=np.int64)
        rec[f"{name}_flat"] = np.array([t for r in rows for t in r], dtype=np.int64)
        rec[f"{name}_lens"] = np.array([len(c) for c in got["input_ids"]], dtype=np.int64)
        rec[f"{name}_out"] = np.array([t for c in got["input_ids"] for t in c], dtype=np.int64)
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "t5_collator.npz"), **rec)
    print({k: v.shape for k, v in rec.items()})


if __name__ == "__main__":
    main()
