"""C5 groundwork (CPU): the relative-position bias vector the attention kernels will consume equals transformers'
T5Attention.compute_bias, and its gradient scatters back onto the bucket table like autograd does."""
import os
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "fengshen-lm_b200"))

from fsb200.models import t5_bias as TB  # noqa: E402


def _hf_attention(is_decoder, heads=4):
    from transformers import MT5Config
    from transformers.models.mt5.modeling_mt5 import MT5Attention
    cfg = MT5Config(vocab_size=64, d_model=32, d_kv=8, d_ff=64, num_layers=1, num_heads=heads, dropout_rate=0.0,
                    relative_attention_num_buckets=32, relative_attention_max_distance=128, is_decoder=is_decoder)
    cfg.is_decoder = is_decoder
    torch.manual_seed(0)
    return MT5Attention(cfg, has_relative_attention_bias=True)


@pytest.mark.parametrize("is_decoder,sq,skv", [(False, 96, 96), (True, 48, 48), (False, 7, 300)])
def test_bias_vector_matches_transformers_compute_bias(is_decoder, sq, skv):
    att = _hf_attention(is_decoder)
    ref = att.compute_bias(sq, skv)[0]                                   # [heads, sq, skv]
    table = att.relative_attention_bias.weight.detach()                  # [buckets, heads]
    vec = TB.rel_bias_vector(table, sq, skv, bidirectional=not is_decoder)
    assert vec.shape == (table.shape[1], sq + skv - 1)
    assert torch.equal(TB.dense_bias(vec, sq, skv), ref.detach())


def test_bucket_function_matches_golden_tables():
    import numpy as np
    g = np.load(os.path.join(ROOT, "tests", "golden", "mt5_small.npz"))
    rel = torch.arange(-300, 301)
    assert np.array_equal(TB.relative_position_bucket(rel, True).numpy(), g["bucket_bidirectional"])
    assert np.array_equal(TB.relative_position_bucket(rel, False).numpy(), g["bucket_causal"])


def test_gradient_scatter_matches_autograd():
    att = _hf_attention(False)
    sq, skv = 40, 56
    table = att.relative_attention_bias.weight
    ref = att.compute_bias(sq, skv)[0]
    g = torch.randn_like(ref)
    ref.backward(g)
    want = table.grad.clone()                                            # [buckets, heads]
    # gradient w.r.t. the vector: sum of g over the (q, k) pairs of each offset
    q = torch.arange(sq)[:, None]; k = torch.arange(skv)[None, :]
    idx = (k - q + sq - 1).reshape(-1)
    dvec = torch.zeros(g.shape[0], sq + skv - 1).index_add_(1, idx, g.reshape(g.shape[0], -1))
    got = TB.scatter_rel_grad(dvec, sq, skv, bidirectional=True)
    assert torch.allclose(got, want, atol=1e-4, rtol=1e-5)
