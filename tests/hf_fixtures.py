"""Fixtures for the HF-backed recipe tests (C2 Wenzhong-GPT2, C3 Erlangshen-MegatronBERT): tiny model directories in the layout
the scripts' `--model_path` / `--pretrained_model_path` expect, built offline (no hub): config.json + vocab / tokenizer files,
and small corpora in the formats the reference's data modules read."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle"))


def bert_dir(path, hidden=256, layers=2, heads=4, vocab_size=512, ff=512, max_pos=128):
    """MegatronBERT directory: config.json (dropout 0 — fsb200 rejects dropout) + vocab.txt (characters of the collator cases)."""
    import bert_collator_cases as C
    os.makedirs(path, exist_ok=True)
    vocab = C.build_vocab()
    with open(os.path.join(path, "vocab.txt"), "w", encoding="utf8") as f:
        f.write("\n".join(vocab) + "\n")
    assert vocab_size >= len(vocab) and vocab_size % 8 == 0
    cfg = {"model_type": "megatron-bert", "vocab_size": vocab_size, "hidden_size": hidden,
           "num_hidden_layers": layers, "num_attention_heads": heads, "intermediate_size": ff,
           "hidden_act": "gelu_new", "hidden_dropout_prob": 0.0, "attention_probs_dropout_prob": 0.0,
           "max_position_embeddings": max_pos, "type_vocab_size": 2, "layer_norm_eps": 1e-12}
    with open(os.path.join(path, "config.json"), "w") as f:
        json.dump(cfg, f)
    return cfg


def bert_corpus(path, n=64):
    """JSON lines {"text": ...}: the collator-case paragraphs, rotated so that documents differ."""
    import bert_collator_cases as C
    docs = [t for t in C.TEXTS if "。" in t]
    with open(path, "w", encoding="utf8") as f:
        for i in range(n):
            a, b = docs[i % len(docs)], docs[(i * 3 + 1) % len(docs)]
            f.write(json.dumps({"text": a + b}, ensure_ascii=False) + "\n")
    return path


def gpt2_tokenizer_dir(path):
    """A byte-level BPE tokenizer without merges (257 entries: 256 bytes + <|endoftext|>), saved as tokenizer.json so that
    AutoTokenizer.from_pretrained(path) loads it offline."""
    from tokenizers import Tokenizer, decoders, models, pre_tokenizers
    from transformers import PreTrainedTokenizerFast
    vocab = {c: i for i, c in enumerate(sorted(pre_tokenizers.ByteLevel.alphabet()))}
    vocab["<|endoftext|>"] = len(vocab)
    tk = Tokenizer(models.BPE(vocab=vocab, merges=[]))
    tk.pre_tokenizer = pre_tokenizers.ByteLevel(add_prefix_space=False)
    tk.decoder = decoders.ByteLevel()
    fast = PreTrainedTokenizerFast(tokenizer_object=tk, eos_token="<|endoftext|>", bos_token="<|endoftext|>",
                                   unk_token="<|endoftext|>")
    os.makedirs(path, exist_ok=True)
    fast.save_pretrained(path)
    return len(vocab)


GPT2_CFG = {"model_type": "gpt2", "vocab_size": 512, "n_embd": 256, "n_layer": 2, "n_head": 4, "n_positions": 128,
            "resid_pdrop": 0.0, "embd_pdrop": 0.0, "attn_pdrop": 0.0, "activation_function": "gelu_new",
            "layer_norm_epsilon": 1e-5}


def qa_files(data_dir, n=48):
    """train / valid / test files of the medical-QA format: one Python dict literal per line (medicalQADataset.py:54-59)."""
    os.makedirs(data_dir, exist_ok=True)
    rows = [{"Question": f"q{i % 7}: what helps a headache number {i % 5}?", "answer": f" rest, water and {i % 3} hours of sleep."}
            for i in range(n)]
    for name in ("train.txt", "valid.txt", "test.txt"):
        with open(os.path.join(data_dir, name), "w", encoding="utf8") as f:
            for r in rows:
                f.write(repr(r) + "\n")
    return rows


MT5_CFG = {"model_type": "mt5", "vocab_size": 512, "d_model": 256, "d_kv": 64, "d_ff": 512, "num_layers": 2,
           "num_decoder_layers": 2, "num_heads": 4, "relative_attention_num_buckets": 32, "dropout_rate": 0.0,
           "feed_forward_proj": "gated-gelu", "pad_token_id": 0, "eos_token_id": 1, "decoder_start_token_id": 0,
           "tie_word_embeddings": True}


def t5_dir(path):
    """mT5 directory for the `bert_tokenizer` branch of pretrain_t5.py: config.json + vocab.txt."""
    import bert_collator_cases as C
    os.makedirs(path, exist_ok=True)
    with open(os.path.join(path, "vocab.txt"), "w", encoding="utf8") as f:
        f.write("\n".join(C.build_vocab()) + "\n")
    with open(os.path.join(path, "config.json"), "w") as f:
        json.dump(MT5_CFG, f)


def t5_tokenised_dir(path, n_train=48, n_test=8, length=70, vocab=400, seed=0):
    """A `datasets.save_to_disk` directory with 'train' / 'test' splits of already tokenised chunks: `input_ids` of the expanded
    length for max_seq_length 64 (compute_input_and_target_lengths(64, .15, 3) = (70, 14))."""
    import datasets
    import numpy as np
    rs = np.random.RandomState(seed)
    # a learnable stream: every chunk repeats a short motif, so that the span targets are predictable from the inputs
    def rows(n):
        out = []
        for _ in range(n):
            motif = rs.randint(5, vocab, size=4)
            out.append(np.tile(motif, length // 4 + 1)[:length].tolist())
        return out
    datasets.DatasetDict({"train": datasets.Dataset.from_dict({"input_ids": rows(n_train)}),
                          "test": datasets.Dataset.from_dict({"input_ids": rows(n_test)})}).save_to_disk(str(path))
