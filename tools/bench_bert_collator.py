"""Host-side throughput of the MegatronBERT sample assembly: Python collator vs the native one (fsb_bert_collate), one core.
Tokenisation (transformers' BertTokenizer) is timed separately — it is common to both."""
import os
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, "oracle"), os.path.join(ROOT, "fengshen-lm_b200", "compat"), os.path.join(ROOT, "fengshen-lm_b200")]
import bert_collator_cases as C  # noqa: E402
from fengshen.data.data_utils.collators import ErLangShenCollator, FastErLangShenCollator  # noqa: E402


def main():
    from transformers import BertTokenizer
    d = tempfile.mkdtemp()
    with open(os.path.join(d, "vocab.txt"), "w", encoding="utf8") as fh:
        fh.write("\n".join(C.build_vocab()) + "\n")
    tok = BertTokenizer(os.path.join(d, "vocab.txt"))
    doc = "".join(t for t in C.TEXTS if "。" in t) * 6            # > 512 tokens: every row is a full 512-token sample
    rows = [{"text": doc}] * 32
    res = {}
    for name, cls in (("python", ErLangShenCollator), ("native", FastErLangShenCollator)):
        c = cls(tokenizer=tok, max_seq_length=512)
        c.setup()
        c.np_rng = np.random.RandomState(0)
        c(rows)
        t = time.perf_counter(); n = 0
        for _ in range(5):
            n += int(c(rows)["attention_mask"].sum())
        res[name] = (n, time.perf_counter() - t)
    t = time.perf_counter()
    for _ in range(5):
        c._ragged(rows)
    t_tok = time.perf_counter() - t
    n = res["python"][0]
    print(f"tokens per batch of 32: {n // 5}")
    print(f"tokenisation only:            {n / t_tok:10.0f} tokens/s")
    for name, (n, dt) in res.items():
        print(f"{name:7s} collator end to end: {n / dt:10.0f} tokens/s   assembly + masking alone: {n / max(dt - t_tok, 1e-9):12.0f} tokens/s")


if __name__ == "__main__":
    main()
