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


@pytest.mark.parametrize("B,S,H,D,causal,layout", [
    (2, 256, 2, 128, True, "llama"),
    (1, 512, 2, 128, True, "llama"),
    (2, 200, 2, 64, True, "gpt2"),
    (2, 384, 2, 64, False, "gpt2"),
    (1, 1024, 1, 64, True, "gpt2"),
    (1, 2048, 1, 128, True, "llama"),
    (3, 77, 2, 64, False, "gpt2"),
])
def test_sdpa_bwd(B, S, H, D, causal, layout):
    g = torch.Generator().manual_seed(0)
    if layout == "llama":
        qkv = torch.randn(B, S, H, 3, D, generator=g).to(torch.bfloat16).to(DEV)
        sel = lambda t, i: t[:, :, :, i]
    else:
        qkv = torch.randn(B, S, 3, H, D, generator=g).to(torch.bfloat16).to(DEV)
        sel = lambda t, i: t[:, :, i]
    q, k, v = sel(qkv, 0), sel(qkv, 1), sel(qkv, 2)
    scale = 1.0 / math.sqrt(D)
    out, lse = ops.sdpa_fwd(q, k, v, scale, causal)
    dout = torch.randn(B, S, H, D, generator=g).to(torch.bfloat16).to(DEV)
    dqkv = torch.full_like(qkv, float("nan"))
    ops.sdpa_bwd(q, k, v, out, dout, lse, scale, causal, sel(dqkv, 0), sel(dqkv, 1), sel(dqkv, 2))
    torch.cuda.synchronize()
    qf, kf, vf = (t.float().detach().requires_grad_(True) for t in (q, k, v))
    ref, _ = _ref_attention(qf, kf, vf, scale, causal)
    ref.backward(dout.float())
    assert not torch.isnan(dqkv.float()).any(), "unwritten gradient elements"
    for name, got, want in (("dq", sel(dqkv, 0), qf.grad), ("dk", sel(dqkv, 1), kf.grad), ("dv", sel(dqkv, 2), vf.grad)):
        err = (got.float() - want).abs().max().item()
        tol = 3e-2 * max(1.0, want.abs().max().item())  # P, dS rounded to bf16 before the second GEMMs
        assert err < tol, f"{name}: max err {err} (tol {tol})"


def test_sdpa_bwd_padding_mask():
    B, S, H, D = 2, 320, 2, 64
    g = torch.Generator().manual_seed(1)
    qkv = torch.randn(B, S, 3, H, D, generator=g).to(torch.bfloat16).to(DEV)
    q, k, v = qkv[:, :, 0], qkv[:, :, 1], qkv[:, :, 2]
    mask = torch.ones(B, S, dtype=torch.uint8, device=DEV)
    mask[0, 250:] = 0
    mask[1, 17:40] = 0
    out, lse = ops.sdpa_fwd(q, k, v, 0.125, False, kv_mask=mask)
    dout = torch.randn(B, S, H, D, generator=g).to(torch.bfloat16).to(DEV)
    dqkv = torch.zeros_like(qkv)
    ops.sdpa_bwd(q, k, v, out, dout, lse, 0.125, False, dqkv[:, :, 0], dqkv[:, :, 1], dqkv[:, :, 2], kv_mask=mask)
    qf, kf, vf = (t.float().detach().requires_grad_(True) for t in (q, k, v))
    ref, _ = _ref_attention(qf, kf, vf, 0.125, False, mask)
    ref.backward(dout.float())
    for got, want in ((dqkv[:, :, 0], qf.grad), (dqkv[:, :, 1], kf.grad), (dqkv[:, :, 2], vf.grad)):
        assert (got.float() - want).abs().max().item() < 3e-2 * max(1.0, want.abs().max().item())


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


# ------------------------------------------------------------------------------------------------ T5 relative-position bias
def _ref_attention_bias(q, k, v, scale, causal, rel, kv_mask=None):
    """transformers mt5/modeling_mt5.py:300-323: scores (unscaled when scale == 1) + position_bias (+ mask) -> fp32 softmax.
    rel [H, Sq + Skv - 1]: bias[h, q, k] = rel[h, k - q + Sq - 1]."""
    Sq, Sk = q.shape[1], k.shape[1]
    qi = torch.arange(Sq, device=q.device)[:, None]
    ki = torch.arange(Sk, device=q.device)[None, :]
    bias = rel[:, ki - qi + Sq - 1]                                   # [H, Sq, Sk]
    s = torch.einsum("bqhd,bkhd->bhqk", q, k) * scale + bias[None]
    if causal:
        s = s.masked_fill(torch.triu(torch.ones(Sq, Sk, dtype=torch.bool, device=s.device), 1), float("-inf"))
    if kv_mask is not None:
        s = s.masked_fill(~kv_mask.bool()[:, None, None, :], float("-inf"))
    p = torch.softmax(s, -1)
    return torch.einsum("bhqk,bkhd->bqhd", p, v), torch.logsumexp(s, -1)


@pytest.mark.parametrize("B,Sq,Sk,H,D,causal,masked", [
    (2, 256, 256, 2, 64, False, False),     # encoder self-attention
    (2, 200, 200, 3, 64, True, False),      # decoder self-attention, ragged tile
    (1, 512, 512, 4, 64, False, True),      # encoder with a padded tail (the collator pads, t5_datasets.py:230-234)
    (2, 114, 114, 2, 64, True, False),      # the reference collator's decoder length
    (1, 384, 384, 2, 128, False, False),    # head dim 128 (t5-11b style)
])
def test_sdpa_rel_bias_fwd_bwd(B, Sq, Sk, H, D, causal, masked):
    g = torch.Generator().manual_seed(3)
    q = (torch.randn(B, Sq, H, D, generator=g) * 0.5).to(torch.bfloat16).to(DEV)
    k = (torch.randn(B, Sk, H, D, generator=g) * 0.5).to(torch.bfloat16).to(DEV)
    v = torch.randn(B, Sk, H, D, generator=g).to(torch.bfloat16).to(DEV)
    rel = (torch.randn(H, Sq + Sk - 1, generator=g) * 1.5).to(DEV)
    mask = None
    if masked:
        mask = torch.ones(B, Sk, dtype=torch.uint8, device=DEV)
        mask[0, Sk - 93:] = 0
    scale = 1.0   # T5 attention is unscaled
    out, lse = ops.sdpa_fwd(q, k, v, scale, causal, kv_mask=mask, rel_bias=rel)
    qf, kf, vf = (t.float().detach().requires_grad_(True) for t in (q, k, v))
    relf = rel.clone().requires_grad_(True)
    ref, ref_lse = _ref_attention_bias(qf, kf, vf, scale, causal, relf, mask)
    assert (out.float() - ref).abs().max().item() < 2e-2
    assert (lse * math.log(2.0) - ref_lse).abs().max().item() < 2e-3
    dout = torch.randn(B, Sq, H, D, generator=g).to(torch.bfloat16).to(DEV)
    dq, dk, dv = (torch.full_like(t, float("nan")) for t in (q, k, v))
    drel = torch.full_like(rel, 0.25)          # accumulated into: the initial content must survive
    ops.sdpa_bwd(q, k, v, out, dout, lse, scale, causal, dq, dk, dv, kv_mask=mask, rel_bias=rel, drel_bias=drel)
    torch.cuda.synchronize()
    ref.backward(dout.float())
    for name, got, want in (("dq", dq, qf.grad), ("dk", dk, kf.grad), ("dv", dv, vf.grad)):
        assert not torch.isnan(got.float()).any(), name
        err = (got.float() - want).abs().max().item()
        assert err < 3e-2 * max(1.0, want.abs().max().item()), f"{name}: {err}"
    want = relf.grad + 0.25
    err = (drel - want).abs().max().item()
    # fp32 diagonal sums of fp32 dS (before its bf16 rounding): only P's ex2.approx and the bf16 inputs separate the two
    assert err < 2e-2 * max(1.0, want.abs().max().item()), f"drel: {err} vs max {want.abs().max().item()}"
    # deterministic: a second run gives the same bits
    drel2 = torch.full_like(rel, 0.25)
    ops.sdpa_bwd(q, k, v, out, dout, lse, scale, causal, dq, dk, dv, kv_mask=mask, rel_bias=rel, drel_bias=drel2)
    assert torch.equal(drel, drel2)


def test_sdpa_cross_attention_no_bias():
    """Decoder -> encoder attention of mT5: seq_q != seq_kv, not causal, zero position bias, encoder padding mask."""
    B, Sq, Sk, H, D = 2, 114, 512, 2, 64
    g = torch.Generator().manual_seed(5)
    q = (torch.randn(B, Sq, H, D, generator=g) * 0.5).to(torch.bfloat16).to(DEV)
    kv = (torch.randn(B, Sk, 2, H, D, generator=g) * 0.5).to(torch.bfloat16).to(DEV)   # fused K|V projection output
    k, v = kv[:, :, 0], kv[:, :, 1]
    mask = torch.ones(B, Sk, dtype=torch.uint8, device=DEV)
    mask[1, 400:] = 0
    out, lse = ops.sdpa_fwd(q, k, v, 1.0, False, kv_mask=mask)
    qf, kf, vf = (t.float().detach().requires_grad_(True) for t in (q, k, v))
    ref, _ = _ref_attention(qf, kf, vf, 1.0, False, mask)
    assert (out.float() - ref).abs().max().item() < 2e-2
    dout = torch.randn(B, Sq, H, D, generator=g).to(torch.bfloat16).to(DEV)
    dq = torch.empty_like(q)
    dkv = torch.empty_like(kv)
    ops.sdpa_bwd(q, k, v, out, dout, lse, 1.0, False, dq, dkv[:, :, 0], dkv[:, :, 1], kv_mask=mask)
    ref.backward(dout.float())
    for got, want in ((dq, qf.grad), (dkv[:, :, 0], kf.grad), (dkv[:, :, 1], vf.grad)):
        assert (got.float() - want).abs().max().item() < 3e-2 * max(1.0, want.abs().max().item())
