"""GPU parity of the fused attention kernels against the reference's un-fused formula
(baddbmm(alpha=1/sqrt(hn)) -> masked softmax -> bmm, fengshen/models/megatron/layers/transformer.py:307-408) in fp32."""
import math

import pytest
import torch

pytestmark = pytest.mark.gpu

from fsb200 import ops  # noqa: E402

DEV = "cuda"


def _ref_attention(q, k, v, scale, causal, kv_mask=None):
    # q,k,v [B,S,H,D] fp32
    s = torch.einsum("bqhd,bkhd->bhqk", q, k) * scale
    Sq, Sk = s.shape[-2:]
    if causal:
        s = s.masked_fill(torch.triu(torch.ones(Sq, Sk, dtype=torch.bool, device=s.device), 1), float("-inf"))
    if kv_mask is not None:
        s = s.masked_fill(~kv_mask.bool()[:, None, None, :], float("-inf"))
    p = torch.softmax(s, -1)
    return torch.einsum("bhqk,bkhd->bqhd", p, v), torch.logsumexp(s, -1)


@pytest.mark.parametrize("B,S,H,D,causal,layout", [
    (2, 256, 2, 128, True, "llama"),
    (1, 512, 3, 128, True, "llama"),
    (2, 200, 2, 64, True, "gpt2"),
    (2, 384, 4, 64, False, "gpt2"),
    (1, 1024, 2, 64, True, "gpt2"),
    (1, 2048, 1, 128, True, "llama"),
    (3, 77, 2, 64, False, "gpt2"),
])
def test_sdpa_fwd(B, S, H, D, causal, layout):
    g = torch.Generator().manual_seed(0)
    h = H * D
    if layout == "llama":   # [B,S,H,3,D] per-head interleaved (transformer.py:488-497)
        qkv = torch.randn(B, S, H, 3, D, generator=g).to(torch.bfloat16).to(DEV)
        q, k, v = qkv[:, :, :, 0], qkv[:, :, :, 1], qkv[:, :, :, 2]
    else:                   # [B,S,3,H,D] contiguous thirds (GPT-2 c_attn)
        qkv = torch.randn(B, S, 3, H, D, generator=g).to(torch.bfloat16).to(DEV)
        q, k, v = qkv[:, :, 0], qkv[:, :, 1], qkv[:, :, 2]
    scale = 1.0 / math.sqrt(D)
    out, lse = ops.sdpa_fwd(q, k, v, scale, causal)
    torch.cuda.synchronize()
    ref, ref_lse = _ref_attention(q.float(), k.float(), v.float(), scale, causal)
    err = (out.float() - ref).abs().max().item()
    # P is rounded to bf16 before the PV product and O is stored in bf16: 2^-8 relative on values of O(1)
    assert err < 2e-2, f"max |O - ref| = {err}"
    lse_err = (lse * math.log(2.0) - ref_lse).abs().max().item()
    assert lse_err < 2e-3, f"max |lse - ref| = {lse_err}"
    assert h == out.shape[2] * out.shape[3]


def test_sdpa_fwd_padding_mask():
    B, S, H, D = 2, 320, 2, 64
    g = torch.Generator().manual_seed(1)
    qkv = torch.randn(B, S, 3, H, D, generator=g).to(torch.bfloat16).to(DEV)
    q, k, v = qkv[:, :, 0], qkv[:, :, 1], qkv[:, :, 2]
    mask = torch.ones(B, S, dtype=torch.uint8, device=DEV)
    mask[0, 250:] = 0
    mask[1, 17:40] = 0
    out, _ = ops.sdpa_fwd(q, k, v, 0.125, False, kv_mask=mask)
    ref, _ = _ref_attention(q.float(), k.float(), v.float(), 0.125, False, mask)
    assert (out.float() - ref).abs().max().item() < 2e-2
