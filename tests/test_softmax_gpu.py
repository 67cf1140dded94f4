"""GPU: the reference's own two CUDA ops (scaled_masked_softmax_cuda, scaled_upper_triang_masked_softmax_cuda) served by
libfsb200.so through the same Python module names and signatures. The reference's only stated tolerance for these kernels
is mean |diff| <= 1e-3 vs torch softmax (fused_kernels/tests/test_fused_kernels.py:97-104,197-204); we assert that AND a
tighter elementwise bound (one bf16 rounding of a probability <= 2^-9)."""
import os
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "fengshen-lm_b200", "compat"))

import scaled_masked_softmax_cuda as sms  # noqa: E402
import scaled_upper_triang_masked_softmax_cuda as suts  # noqa: E402


@pytest.mark.parametrize("b,np_,sq,sk,mb", [(2, 4, 64, 64, 2), (1, 3, 128, 256, 1), (2, 2, 40, 2048, 1), (1, 2, 16, 1000, 1)])
def test_scaled_masked_softmax_fwd_bwd(b, np_, sq, sk, mb):
    if sk % 8:
        with pytest.raises(RuntimeError):
            sms.forward(torch.randn(b, np_, sq, sk, device="cuda", dtype=torch.bfloat16),
                        torch.zeros(mb, 1, sq, sk, dtype=torch.uint8, device="cuda"), 1.0)
        return
    g = torch.Generator(device="cuda").manual_seed(0)
    x = (torch.randn(b, np_, sq, sk, device="cuda", generator=g) * 3).to(torch.bfloat16)
    mask = torch.rand(mb, 1, sq, sk, device="cuda", generator=g) < 0.2
    scale = 0.125
    y = sms.forward(x, mask, scale)
    xf = x.float().requires_grad_(True)
    ref = torch.softmax((xf * scale).masked_fill(mask, -10000.0), dim=-1)     # fused_softmax.py mask_func semantics
    assert (y.float() - ref).abs().mean().item() <= 1e-3
    assert (y.float() - ref).abs().max().item() <= 2.0 ** -8
    dy = torch.randn(b, np_, sq, sk, device="cuda", generator=g).to(torch.bfloat16)
    ref.backward(dy.float())
    dy_buf = dy.clone()
    dx = sms.backward(dy_buf, y, scale)
    assert dx.data_ptr() == dy_buf.data_ptr()                                 # in place, like the reference
    # the reference kernel differentiates through its own bf16 y: compare against the same formula in fp32
    yy = y.float()
    want = scale * (dy.float() * yy - yy * (dy.float() * yy).sum(-1, keepdim=True))
    assert (dx.float() - want).abs().max().item() <= 2e-2 * max(1.0, want.abs().max().item())
    assert (dx.float() - xf.grad).abs().mean().item() <= 1e-3


def test_get_batch_per_block_matches_reference_formula():
    # scaled_masked_softmax.h:337-349
    assert sms.get_batch_per_block(128, 2048, 2, 40) == 4
    assert sms.get_batch_per_block(128, 128, 2, 40) == 8
    assert sms.get_batch_per_block(128, 64, 2, 40) == 8
    assert sms.get_batch_per_block(128, 16, 2, 40) == 16


@pytest.mark.parametrize("ab,s", [(8, 64), (4, 512), (2, 2048)])
def test_scaled_upper_triang_masked_softmax(ab, s):
    g = torch.Generator(device="cuda").manual_seed(1)
    x = (torch.randn(ab, s, s, device="cuda", generator=g) * 2).to(torch.bfloat16)
    y = suts.forward(x, 0.25)
    causal = torch.triu(torch.ones(s, s, dtype=torch.bool, device="cuda"), 1)
    xf = x.float().requires_grad_(True)
    ref = torch.softmax((xf * 0.25).masked_fill(causal, float("-inf")), dim=-1)
    assert (y.float() - ref).abs().max().item() <= 2.0 ** -8
    assert y.float().masked_select(causal).abs().max().item() == 0.0        # exact zeros above the diagonal
    dy = torch.randn(ab, s, s, device="cuda", generator=g).to(torch.bfloat16)
    ref.backward(dy.float())
    dx = suts.backward(dy.clone(), y, 0.25)
    assert (dx.float() - xf.grad).abs().mean().item() <= 1e-3
    assert dx.float().masked_select(causal).abs().max().item() == 0.0
