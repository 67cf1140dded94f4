def get_a_and_b_segments(sample, np_rng):
    """fengshen/data/data_utils/sop_utils.py:2-32. `sample` is a list of tokenised sentences (>= 2). The first `cut` sentences
    form segment A, the rest segment B; with three or more sentences `cut` is drawn from [1, n) (ONE randint), then ONE uniform
    decides whether the two segments swap (the sentence-order-prediction label)."""
    n = len(sample)
    assert n > 1, 'make sure each sample has at least two sentences.'
    cut = np_rng.randint(1, n) if n >= 3 else 1
    first = [tok for sent in sample[:cut] for tok in sent]
    second = [tok for sent in sample[cut:] for tok in sent]
    if np_rng.random() < 0.5:
        return second, first, True
    return first, second, False
