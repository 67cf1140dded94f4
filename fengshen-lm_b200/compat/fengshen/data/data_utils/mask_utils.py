"""Masked-LM target selection (whole-word, n-gram) — fengshen/data/data_utils/mask_utils.py:19-285, itself Megatron-LM's
`create_masked_lm_predictions`. Restated around explicit word spans; the order and number of draws from `np_rng` is the
reference's (shuffle of the candidate list, one `choice` / `geometric` per visited candidate, one or two uniforms + an optional
`randint` per masked token, a second shuffle before the optional permutation pass), so a shared RandomState stays in lock-step
with the reference across samples."""
import collections

import numpy as np

MaskedLmInstance = collections.namedtuple("MaskedLmInstance", ["index", "label"])


def is_start_piece(piece):
    """WordPiece continuation pieces carry a leading '##' (mask_utils.py:9-15)."""
    return not piece.startswith("##")


def _wordpiece_spans(tokens, id2tok, cls_id, sep_id, whole_word):
    """Candidate words from '##' continuation marks (mask_utils.py:58-77). [CLS]/[SEP] are boundaries but never candidates."""
    spans, boundary = [], [0] * len(tokens)
    for i, t in enumerate(tokens):
        if t == cls_id or t == sep_id:
            boundary[i] = 1
            continue
        starts = is_start_piece(id2tok[t])
        if whole_word and spans and not starts:
            spans[-1].append(i)
        else:
            spans.append([i])
            if starts:
                boundary[i] = 1
    return spans, boundary


def _segmenter_spans(tokens, id2tok, cls_id, sep_id, zh_tokenizer):
    """Candidate words from a Chinese word segmenter such as jieba.lcut (mask_utils.py:78-123): at each position take the
    longest segmenter word that the following single-character tokens spell out; '##' runs (old-style vocabularies) win."""
    text = ''.join(id2tok[t] for t in tokens if t != cls_id and t != sep_id)
    words = set(zh_tokenizer(text, HMM=True))
    longest = {}
    for w in words:
        if w and longest.get(w[0], 0) < len(w):
            longest[w[0]] = len(w)
    spans, boundary = [], [0] * len(tokens)
    n, i = len(tokens), 0
    while i < n:
        piece = id2tok[tokens[i]]
        if len(piece) == 0 or tokens[i] == cls_id or tokens[i] == sep_id:
            boundary[i] = 1
            i += 1
            continue
        end = i + 1
        while end < n and id2tok[tokens[end]].startswith('##'):
            end += 1
        if end == i + 1:   # no '##' run: grow a word character by character, remember the longest hit
            acc = ''
            for j in range(min(longest.get(piece[0], 1), n - i)):
                acc += id2tok[tokens[i + j]]
                if acc in words:
                    end = i + j + 1
        spans.append(list(range(i, end)))
        boundary[i] = 1
        i = end
    return spans, boundary


def create_masked_lm_predictions(tokens, vocab_id_list, vocab_id_to_token_dict, masked_lm_prob, cls_id, sep_id, mask_id,
                                 max_predictions_per_seq, np_rng, max_ngrams=3, do_whole_word_mask=True,
                                 favor_longer_ngram=False, do_permutation=False, geometric_dist=False,
                                 masking_style="bert", zh_tokenizer=None):
    """Returns (output_tokens, masked_positions, masked_labels, token_boundary, masked_spans); the 4-tuple without spans when
    masked_lm_prob == 0 (mask_utils.py:129-131). Tokens are vocabulary ids."""
    if zh_tokenizer is None:
        spans, boundary = _wordpiece_spans(tokens, vocab_id_to_token_dict, cls_id, sep_id, do_whole_word_mask)
    else:
        spans, boundary = _segmenter_spans(tokens, vocab_id_to_token_dict, cls_id, sep_id, zh_tokenizer)
    out = list(tokens)
    if masked_lm_prob == 0:
        return out, [], [], boundary
    if masking_style not in ("bert", "t5"):
        raise ValueError("invalid value of masking style")

    budget = min(max_predictions_per_seq, max(1, int(round(len(tokens) * masked_lm_prob))))
    sizes = np.arange(1, max_ngrams + 1, dtype=np.int64)
    pvals = None
    if not geometric_dist:   # shorter n-grams are favoured: p(n) ~ 1/n (mask_utils.py:137-143)
        pvals = 1. / np.arange(1, max_ngrams + 1)
        pvals /= pvals.sum(keepdims=True)
        if favor_longer_ngram:
            pvals = pvals[::-1]

    def window(start, n):   # token positions of the n-gram of words starting at word `start` (clipped at the end)
        return [p for word in spans[start:start + n] for p in word]

    def shrink_to(start, n, used):
        """The reference's retry loop (mask_utils.py:178-188): try n, n-1, ..., 1 words until the n-gram fits the budget."""
        picked = window(start, n)
        n -= 1
        while used + len(picked) > budget and n > 0:
            picked = window(start, n)
            n -= 1
        return picked

    order = list(range(len(spans)))
    np_rng.shuffle(order)
    picked_lm, picked_spans, covered = [], [], set()
    for start in order:
        if len(picked_lm) >= budget:
            break
        if geometric_dist:   # SpanBERT: p = 0.2, clipped
            n = min(np_rng.geometric(0.2), max_ngrams)
        else:
            n = np_rng.choice(sizes, p=pvals / pvals.sum(keepdims=True))
        cand = shrink_to(start, int(n), len(picked_lm))
        if len(picked_lm) + len(cand) > budget or any(p in covered for p in cand):
            continue
        for p in cand:
            covered.add(p)
            if masking_style == "t5":
                new = mask_id
            elif np_rng.random() < 0.8:            # 80 %: [MASK]
                new = mask_id
            elif np_rng.random() < 0.5:            # 10 %: keep
                new = tokens[p]
            else:                                   # 10 %: a random vocabulary entry
                new = vocab_id_list[np_rng.randint(0, len(vocab_id_list))]
            out[p] = new
            picked_lm.append(MaskedLmInstance(index=p, label=tokens[p]))
        picked_spans.append(MaskedLmInstance(index=cand, label=[tokens[p] for p in cand]))
    assert len(picked_lm) <= budget
    np_rng.shuffle(order)   # drawn whether or not the permutation pass runs (mask_utils.py:230)

    if do_permutation:
        chosen = set()
        for start in order:
            if len(chosen) >= budget:
                break
            # the reference draws this one from numpy's GLOBAL generator (mask_utils.py:245), not from np_rng
            n = np.random.choice(sizes, p=pvals / pvals.sum(keepdims=True))
            cand = shrink_to(start, int(n), len(chosen))
            if len(chosen) + len(cand) > budget or any(p in covered or p in chosen for p in cand):
                continue
            chosen.update(cand)
        assert len(chosen) <= budget
        src = sorted(chosen)
        dst = list(src)
        np_rng.shuffle(dst)
        before = list(out)
        for s, d in zip(src, dst):
            out[s] = before[d]
            picked_lm.append(MaskedLmInstance(index=s, label=before[s]))

    picked_lm.sort(key=lambda m: m.index)
    picked_spans.sort(key=lambda m: m.index[0])
    return out, [m.index for m in picked_lm], [m.label for m in picked_lm], boundary, picked_spans
