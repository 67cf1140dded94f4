"""Relative-position bias of the T5 / mT5 attention stacks (BASELINE config 5) — host-side part.

The reference's `pretrain_t5.py:57-59` builds transformers' MT5ForConditionalGeneration, whose first self-attention layer of
each stack owns a `[num_buckets, heads]` embedding; every layer of the stack adds `bias[h, q, k] = table[bucket(k - q), h]` to the
UNSCALED scores (transformers models/t5/modeling_t5.py: T5Attention._relative_position_bucket / compute_bias). Because the
bias depends on (q, k) only through the offset r = k - q, the attention kernels take it as one vector per head over
r in [-(S_q - 1), S_kv - 1] (`rel_bias_vector`), and its gradient comes back in the same shape (`scatter_rel_grad`).
Pure integer / indexing logic: tested on CPU against transformers (tests/test_t5_bias_cpu.py)."""
import math

import torch


def relative_position_bucket(rel, bidirectional, num_buckets=32, max_distance=128):
    """rel = key_position - query_position (int64 tensor) -> bucket index, same arithmetic (fp32 log, truncation) as T5."""
    rel = rel.to(torch.int64)
    out = torch.zeros_like(rel)
    nb = num_buckets
    if bidirectional:
        nb //= 2
        out = out + (rel > 0).to(torch.int64) * nb
        rel = rel.abs()
    else:
        rel = -torch.clamp(rel, max=0)
    max_exact = nb // 2
    is_small = rel < max_exact
    large = max_exact + (torch.log(rel.clamp(min=1).float() / max_exact) / math.log(max_distance / max_exact)
                         * (nb - max_exact)).to(torch.int64)
    large = torch.minimum(large, torch.full_like(large, nb - 1))
    return out + torch.where(is_small, rel, large)


def rel_offsets(seq_q, seq_kv, device=None):
    """All offsets k - q that occur: [-(S_q - 1), S_kv - 1]; index i of the bias vector is offset i - (S_q - 1)."""
    return torch.arange(-(seq_q - 1), seq_kv, dtype=torch.int64, device=device)


_bucket_cache = {}


def offset_buckets(seq_q, seq_kv, bidirectional, num_buckets=32, max_distance=128, device=None):
    """Bucket index of every offset k - q in [-(S_q - 1), S_kv - 1] (cached: pure function of its integer arguments)."""
    key = (seq_q, seq_kv, bool(bidirectional), num_buckets, max_distance, str(device))
    b = _bucket_cache.get(key)
    if b is None:
        b = relative_position_bucket(rel_offsets(seq_q, seq_kv, device), bidirectional, num_buckets, max_distance)
        _bucket_cache[key] = b
    return b


def rel_bias_vector(table, seq_q, seq_kv, bidirectional, num_buckets=32, max_distance=128):
    """table [num_buckets, heads] -> fp32 [heads, S_q + S_kv - 1]: bias[h, q, k] == vec[h, k - q + S_q - 1]."""
    b = offset_buckets(seq_q, seq_kv, bidirectional, num_buckets, max_distance, table.device)
    return table.float()[b].t().contiguous()


def scatter_rel_grad(dvec, seq_q, seq_kv, bidirectional, num_buckets=32, max_distance=128):
    """Gradient w.r.t. the bias vector [heads, S_q + S_kv - 1] -> gradient of the table [num_buckets, heads] (fp32,
    deterministic: index_add over a sorted, fixed index)."""
    b = offset_buckets(seq_q, seq_kv, bidirectional, num_buckets, max_distance, dvec.device)
    out = torch.zeros(num_buckets, dvec.shape[0], dtype=torch.float32, device=dvec.device)
    out.index_add_(0, b, dvec.float().t().contiguous())
    return out


def dense_bias(vec, seq_q, seq_kv):
    """[heads, S_q + S_kv - 1] -> [heads, S_q, S_kv] (checker / small-shape helper; the kernels never materialise this)."""
    q = torch.arange(seq_q, device=vec.device)[:, None]
    k = torch.arange(seq_kv, device=vec.device)[None, :]
    return vec[:, (k - q + seq_q - 1)]
