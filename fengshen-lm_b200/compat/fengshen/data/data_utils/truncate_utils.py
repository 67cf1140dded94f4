def truncate_segments(tokens_a, tokens_b, len_a, len_b, max_num_tokens, np_rng):
    """fengshen/data/data_utils/truncate_utils.py:2-19. Shortens the longer segment (ties: B) one token at a time until the
    pair fits, IN PLACE; every removed token costs ONE uniform (front if < 0.5, else back). Returns whether anything was cut."""
    assert len_a > 0
    excess = len_a + len_b - max_num_tokens
    if excess <= 0:
        return False
    for _ in range(excess):
        if len_a > len_b:
            victim, len_a = tokens_a, len_a - 1
        else:
            victim, len_b = tokens_b, len_b - 1
        if np_rng.random() < 0.5:
            del victim[0]
        else:
            victim.pop()
    return True
