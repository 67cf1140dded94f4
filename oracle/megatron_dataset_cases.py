"""TEST INFRASTRUCTURE — seeded cases for the Megatron BERT data path (mmap indexed dataset files, split / blend arithmetic,
BertDataset samples). `run_cases(M, helpers, workdir)` runs them against ANY implementation of the
`fengshen.data.megatron_dataloader` modules: oracle/make_golden_megatron_dataset.py feeds it the unmodified reference's Python
modules + its compiled C++ helpers (-> tests/golden/megatron_dataset.npz), tests/test_megatron_dataset_cpu.py the compat ones."""
import os
from types import SimpleNamespace

import numpy as np
import torch

import bert_collator_cases as B


def ordered_tokenizer():
    """What BertDataset reads from a tokenizer, with `vocab` in id order (a tokenizers-backed BertTokenizer's has none)."""
    vocab = {t: i for i, t in enumerate(B.build_vocab())}
    return SimpleNamespace(vocab=vocab, cls_token_id=vocab["[CLS]"], sep_token_id=vocab["[SEP]"],
                           mask_token_id=vocab["[MASK]"], pad_token_id=vocab["[PAD]"])


def corpus_sentences(seed, n_docs):
    """Documents of 0..7 sentences of 1..12 token ids (specials excluded), a few empty documents."""
    rs = np.random.RandomState(seed)
    V = len(B.build_vocab())
    docs = []
    for _ in range(n_docs):
        k = int(rs.randint(0, 8)) if rs.rand() > 0.05 else 0
        docs.append([rs.randint(5, V, size=int(rs.randint(1, 13))).astype(np.int64) for _ in range(k)])
    return docs


def write_dataset(M, prefix, docs, dtype=None, vocab_size=None):
    I = M["indexed_dataset"]
    if dtype is None:
        builder = I.make_builder(I.data_file_path(prefix), "mmap", vocab_size=vocab_size)
    else:
        builder = I.MMapIndexedDatasetBuilder(I.data_file_path(prefix), dtype=dtype)
    for d in docs:
        for s in d:
            builder.add_item(torch.from_numpy(s))
        builder.end_document()
    builder.finalize(I.index_file_path(prefix))


def file_bytes(path):
    with open(path, "rb") as f:
        return np.frombuffer(f.read(), dtype=np.uint8).copy()


def run_cases(M, helpers, workdir):
    out = {}
    I, U = M["indexed_dataset"], M["dataset_utils"]
    # ---- files: byte-for-byte, three dtypes, the vocab-size rule of make_builder
    docs = corpus_sentences(1, 60)
    for tag, kw in (("u16", dict(vocab_size=30000)), ("i32", dict(vocab_size=100000)), ("i64", dict(dtype=np.int64))):
        prefix = os.path.join(workdir, f"corpus_{tag}")
        write_dataset(M, prefix, docs, **kw)
        out[f"idx_{tag}"], out[f"bin_{tag}"] = file_bytes(prefix + ".idx"), file_bytes(prefix + ".bin")
    prefix = os.path.join(workdir, "corpus_u16")
    ds = I.make_dataset(prefix, "infer", skip_warmup=True)
    out["sizes"], out["doc_idx"] = np.array(ds.sizes), np.array(ds.doc_idx)
    out["item_7"], out["get_7"] = np.array(ds[7]), np.array(ds.get(7, offset=1, length=None))
    out["slice_cat"] = np.concatenate(ds[3:9])
    out["n"] = np.array([len(ds), int(I.infer_dataset_impl(prefix) == "mmap"), int(I.dataset_exists(prefix, "mmap"))])
    # ---- merge_file_: the second file's sentences are appended (its document boundaries are not, in the reference)
    other = os.path.join(workdir, "corpus_other")
    write_dataset(M, other, corpus_sentences(2, 10), vocab_size=30000)
    merged = os.path.join(workdir, "corpus_merged")
    b = I.make_builder(I.data_file_path(merged), "mmap", vocab_size=30000)
    for s in docs[1]:
        b.add_item(torch.from_numpy(s))
    b.end_document()
    b.merge_file_(other)
    b.finalize(I.index_file_path(merged))
    out["idx_merged"], out["bin_merged"] = file_bytes(merged + ".idx"), file_bytes(merged + ".bin")
    # ---- split / blend arithmetic
    out["splits"] = np.array([U.get_train_valid_test_split_(s, n) for s, n in
                              (("949,50,1", 1000), ("0.9/0.1", 77), ("1", 5), ("3,3,3", 10), ("98,1,1", 3))])
    pre, w, cnt = U.get_datasets_weights_and_num_samples(["0.3", " a ", "0.7", "b"], [1000, 100, 10])
    out["blend_w"], out["blend_cnt"] = np.array(w), np.array(cnt)
    # ---- BertDataset: the index map is built by `helpers.build_mapping` and cached under the name get_samples_mapping looks for
    # (the reference's build_training_sample does not truncate the A/B pair — its truncate_segments call is commented out — and
    # asserts when a span overshoots; the sentence-order case therefore uses documents shorter than the sequence length, the
    # plain-MLM case relies on the clipping of segment A)
    tok = ordered_tokenizer()
    for tag, (binary_head, max_len, p_short, seed, cap) in {"sop": (True, 128, 0.1, 1234, 150), "mlm": (False, 48, 0.0, 7, 90)}.items():
        epochs, max_seq = np.iinfo(np.int32).max - 1, max_len - 3
        mapping = helpers.build_mapping(np.array(ds.doc_idx), np.array(ds.sizes), epochs, cap, max_seq, p_short, seed, False,
                                        2 if binary_head else 1)
        fn = prefix + "_{}_indexmap_{}mns_{}msl_{:0.2f}ssp_{}s.npy".format(tag, cap, max_seq, p_short, seed)
        np.save(fn, mapping, allow_pickle=True)
        bd = M["bert_dataset"].BertDataset(name=tag, indexed_dataset=ds, data_prefix=prefix, num_epochs=None, max_num_samples=cap,
                                           masked_lm_prob=0.15, max_seq_length=max_len, short_seq_prob=p_short, seed=seed,
                                           binary_head=binary_head, tokenizer=tok, masking_style="bert")
        out[f"{tag}_len"] = np.array([len(bd)])
        picks = [0, 1, 2, len(bd) // 2, len(bd) - 1]
        for k in ("input_ids", "token_type_ids", "labels", "attention_mask"):
            out[f"{tag}_{k}"] = np.stack([bd[i][k] for i in picks])
        out[f"{tag}_nsl"] = np.array([bd[i]["next_sentence_label"] for i in picks])
        os.remove(fn)
    return out
