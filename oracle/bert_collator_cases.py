"""TEST INFRASTRUCTURE — the seeded cases that pin the BERT-family sample assembly (sentence split, A/B segments, truncation,
token types, whole-word n-gram masking, and the full ErLangShenCollator batch). `run_cases(mods, collator_cls)` executes them
against ANY implementation of the five `fengshen.data.data_utils` modules: oracle/make_golden_bert_collator.py feeds it the
unmodified reference (-> tests/golden/bert_collator.json), tests/test_bert_collator_cpu.py feeds it fengshen-lm_b200/compat."""
import numpy as np

TEXTS = [
    "今天天气很好。我们去公园散步吧！你觉得怎么样？他说：“我不想去。”然后就走了。",
    "“你好！”她笑着说。“很高兴认识你。”我们握了握手……之后再也没见过。",
    "深度学习模型的训练需要大量的数据。数据的质量决定了模型的上限!好的数据胜过好的模型。优化器只是工具。",
    "单句没有标点",
    "第一句。第二句。",
    "他问：“这是什么？是新的吗？”我答：“是的。”",
    "unbelievable results were reported。playing games is fun。",
]
PIECES = ["##ing", "##ed", "##s", "##able", "un", "believ", "play", "report", "result", "game", "were", "is", "fun"]


def build_vocab():
    chars = sorted({c for t in TEXTS for c in t if not c.isspace() and not c.isascii()})
    punct = sorted({c for t in TEXTS for c in t if c.isascii() and not c.isalnum() and not c.isspace()})
    return ["[PAD]", "[UNK]", "[CLS]", "[SEP]", "[MASK]"] + chars + punct + PIECES


def fake_segmenter(text, HMM=True):
    """Stands in for jieba.lcut: greedy two-character words, deterministic."""
    return [text[i:i + 2] for i in range(0, len(text), 2)]


def _py(x):
    if isinstance(x, (list, tuple)):
        return [_py(v) for v in x]
    if isinstance(x, (np.integer,)):
        return int(x)
    if isinstance(x, (np.bool_,)):
        return bool(x)
    if hasattr(x, "_asdict"):
        return _py(list(x))
    return x


def run_cases(mods, collator_cls, tokenizer):
    """mods: dict name -> module for sentence_split, sop_utils, truncate_utils, token_type_utils, mask_utils."""
    out = {}
    split = mods["sentence_split"].ChineseSentenceSplitter()
    out["sentences"] = [split.tokenize(t) for t in TEXTS]

    rs = np.random.RandomState(7)
    seg = []
    for n in (2, 3, 5, 9):
        for _ in range(6):
            sample = [list(range(100 * i, 100 * i + 1 + (i * 7) % 5)) for i in range(n)]
            seg.append(_py(mods["sop_utils"].get_a_and_b_segments(sample, rs)))
    out["segments"] = seg

    rs = np.random.RandomState(8)
    tr = []
    for la, lb, mx in ((10, 3, 20), (30, 5, 12), (7, 40, 16), (25, 25, 9), (3, 0, 2)):
        a, b = list(range(la)), list(range(1000, 1000 + lb))
        r = mods["truncate_utils"].truncate_segments(a, b, la, lb, mx, rs)
        tr.append([bool(r), a, b])
    out["truncate"] = tr
    out["tokentypes"] = [_py(mods["token_type_utils"].create_tokens_and_tokentypes(a, b, 2, 3))
                         for a, b in (([5, 6, 7], [8, 9]), ([5], []), ([], [4]))]

    vocab = tokenizer.vocab
    inv = {v: k for k, v in sorted(vocab.items(), key=lambda kv: kv[1])}
    ids = list(inv.keys())   # sorted: `tokenizer.vocab` of a tokenizers-backed BertTokenizer has no stable order across processes
    f = mods["mask_utils"].create_masked_lm_predictions
    cls_id, sep_id, mask_id = vocab["[CLS]"], vocab["[SEP]"], vocab["[MASK]"]
    masks = []
    rs = np.random.RandomState(9)
    for ti, text in enumerate(TEXTS):
        toks = [cls_id] + tokenizer.convert_tokens_to_ids(tokenizer.tokenize(text)) + [sep_id]
        variants = [dict(), dict(favor_longer_ngram=True), dict(geometric_dist=True), dict(masking_style="t5", max_ngrams=5),
                    dict(do_whole_word_mask=False), dict(zh_tokenizer=fake_segmenter)]
        for kw in variants:
            for prob, cap in ((0.15, 0.15 * len(toks)), (0.4, 6), (0, 3)):
                masks.append(_py(f(toks, ids, inv, prob, cls_id, sep_id, mask_id, cap, rs, **kw)))
    np.random.seed(3)   # the permutation pass draws its n-gram sizes from the global generator
    toks = [cls_id] + tokenizer.convert_tokens_to_ids(tokenizer.tokenize(TEXTS[2])) + [sep_id]
    for _ in range(4):
        masks.append(_py(f(toks, ids, inv, 0.2, cls_id, sep_id, mask_id, 20, rs, do_permutation=True)))
    out["masks"] = masks

    batches = []
    for seed, L in ((11, 32), (12, 64), (13, 16)):
        coll = collator_cls(tokenizer=tokenizer, max_seq_length=L, masked_lm_prob=0.15, content_key="text")
        coll.setup()
        coll.np_rng = np.random.RandomState(seed)
        coll.vocab_id_list = sorted(coll.vocab_id_list)   # see above: pin the order the random-replacement draw indexes
        for _ in range(2):
            b = coll([{"text": t} for t in TEXTS])
            batches.append({k: v.tolist() for k, v in b.items()})
    out["batches"] = batches
    return out
