def padding_to_maxlength(ids, max_length, pad_id):
    """fengshen/data/data_utils/common_utils.py:1-4: right-pad `ids`, return (padded ids, 1/0 attention mask)."""
    n, fill = len(ids), max_length - len(ids)
    return ids + [pad_id] * fill, [1] * n + [0] * fill
