def create_tokens_and_tokentypes(tokens_a, tokens_b, cls_id, sep_id):
    """fengshen/data/data_utils/token_type_utils.py:1-25: [CLS] A [SEP] (type 0), then B [SEP] (type 1) when B is non-empty."""
    tokens = [cls_id, *tokens_a, sep_id]
    types = [0] * len(tokens)
    if tokens_b:
        tokens += [*tokens_b, sep_id]
        types += [1] * (len(tokens_b) + 1)
    return tokens, types
