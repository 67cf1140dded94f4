"""GPU parity of the fsb200 GPT-2 step against the implementation the reference actually calls:
`transformers.GPT2LMHeadModel` (fengshen/examples/wenzhong_qa/finetune_wenzhong.py:56), run here on CPU in fp32 with
attn_implementation="eager" and dropout 0. transformers is a third-party dependency of the
reference (setup.py:17, `transformers>=4.17.0`, unpinned); parity is pinned to the installed 5.5.0 and to the golden
loss values in tests/golden/gpt2_small.npz produced by oracle/make_golden_hf.py.

Tolerances as in test_llama_gpu.py (bf16 activations vs fp32 reference)."""
import os
import sys

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import hf_oracle as H  # noqa: E402  (checker only)

from fsb200.engine import ZeroEngine  # noqa: E402
from fsb200.models.gpt2 import GPT2LMHeadModel  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden", "gpt2_small.npz")


def _mine(ref):
    m = GPT2LMHeadModel(ref.config, device="cuda")
    m.load_reference_state_dict(ref.state_dict())
    return m


def test_gpt2_forward_backward_vs_transformers():
    g = np.load(GOLD)
    ref = H.build_gpt2(H.GPT2_SMALL)
    batch = H.make_lm_batch(H.GPT2_SMALL["vocab_size"], 2, 96, seed=1234)
    out_ref = ref(input_ids=batch["input_ids"], labels=batch["labels"])
    assert abs(out_ref.loss.item() - float(g["loss"])) < 1e-5  # live HF == committed golden
    out_ref.loss.backward()
    mine = _mine(ref)
    out = mine(input_ids=batch["input_ids"].cuda(), labels=batch["labels"].cuda(), return_logits=True)
    assert abs(out.loss.item() - out_ref.loss.item()) <= 3e-3
    tol = 4 * 2.0 ** -8 * out_ref.logits.abs().max().item()
    assert (out.logits.float().cpu() - out_ref.logits).abs().max().item() <= tol
    out.loss.backward()
    torch.cuda.synchronize()
    ref_params = dict(ref.named_parameters())
    for name, prm in mine.named_parameters():
        got = prm.main_grad.float().cpu().flatten()
        want = ref_params[name].grad.flatten()
        cos = torch.dot(got, want) / (got.norm() * want.norm() + 1e-30)
        assert cos.item() >= 0.998, (name, cos.item())
        assert abs(got.norm().item() / (want.norm().item() + 1e-30) - 1.0) <= 0.03, name


def test_gpt2_padding_mask_matches_transformers():
    ref = H.build_gpt2(H.GPT2_SMALL)
    batch = H.make_lm_batch(H.GPT2_SMALL["vocab_size"], 2, 64, seed=7)
    am = torch.ones(2, 64, dtype=torch.int64)
    am[1, 50:] = 0
    labels = batch["labels"].clone()
    labels[am == 0] = -100
    out_ref = ref(input_ids=batch["input_ids"], attention_mask=am, labels=labels)
    mine = _mine(ref)
    out = mine(input_ids=batch["input_ids"].cuda(), attention_mask=am.cuda(), labels=labels.cuda())
    assert abs(out.loss.item() - out_ref.loss.item()) <= 3e-3


def test_gpt2_loss_curve_vs_transformers_golden():
    g = np.load(GOLD)
    ref = H.build_gpt2(H.GPT2_SMALL)
    mine = _mine(ref)
    lr, b1, b2, eps, wd = (float(x) for x in g["train_hparams"])
    steps = len(g["loss_curve"])
    eng = ZeroEngine(mine, lr=lr, betas=(b1, b2), eps=eps, weight_decay=wd)
    batches = [H.make_lm_batch(H.GPT2_SMALL["vocab_size"], 2, 96, seed=1234 + i) for i in range(4)]
    curve = []
    for it in range(steps):
        b = batches[it % 4]
        out = mine(input_ids=b["input_ids"].cuda(), labels=b["labels"].cuda())
        out.loss.backward()
        eng.backward_done()
        eng.step(lr=H.linear_lr(it, lr, 2, steps))
        curve.append(out.loss.item())
    err = np.abs(np.array(curve) - g["loss_curve"]).max()
    assert err <= 2e-2, (err, curve[:3], g["loss_curve"][:3])
