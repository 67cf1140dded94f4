"""GPU parity of the fsb200 BERT / MegatronBERT steps against the classes the reference's MLM pretraining scripts call
(transformers.BertForMaskedLM, examples/pretrain_bert/pretrain_bert.py:137; transformers.MegatronBertForPreTraining,
examples/pretrain_erlangshen_bert/pretrain_erlangshen.py:141), run on CPU in fp32 (eager attention, dropout 0) on
bf16-exact weights. Golden losses: tests/golden/bert_small.npz (oracle/make_golden_hf.py). Tolerances as in
test_llama_gpu.py."""
import os
import sys

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import hf_oracle as H  # noqa: E402  (checker only)

from fsb200.models.bert import BertForMaskedLM, MegatronBertForPreTraining  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden", "bert_small.npz")


def _compare(ref, mine, batch, loss_key):
    g = np.load(GOLD)
    out_ref = ref(**batch)
    assert abs(out_ref.loss.item() - float(g[loss_key])) < 1e-5          # live HF == committed golden
    out_ref.loss.backward()
    out = mine(**{k: v.cuda() for k, v in batch.items()}, return_logits=True)
    assert abs(out.loss.item() - out_ref.loss.item()) <= 4e-3, (out.loss.item(), out_ref.loss.item())
    ref_logits = out_ref.logits if hasattr(out_ref, "logits") and out_ref.logits is not None else out_ref.prediction_logits
    tol = 4 * 2.0 ** -8 * ref_logits.abs().max().item()
    assert (out.logits.float().cpu() - ref_logits).abs().max().item() <= tol
    out.loss.backward()
    torch.cuda.synchronize()
    refp = dict(ref.named_parameters())
    for name, prm in mine.named_parameters():
        want = refp[name].grad
        got = prm.main_grad.float().cpu()
        if want is None or want.norm().item() < 1e-7:
            assert got.norm().item() < 1e-4, name
            continue
        cos = torch.dot(got.flatten(), want.flatten()) / (got.norm() * want.norm() + 1e-30)
        assert cos.item() >= 0.998, (name, cos.item())
        assert abs(got.norm().item() / want.norm().item() - 1.0) <= 0.03, (name, got.norm().item(), want.norm().item())


def test_bert_mlm_vs_transformers():
    ref = H.build_bert(H.BERT_SMALL)
    mine = BertForMaskedLM(ref.config, device="cuda")
    mine.load_reference_state_dict(ref.state_dict())
    _compare(ref, mine, H.make_mlm_batch(H.BERT_SMALL["vocab_size"], 3, 96, seed=5, pad_tail=20), "bert_loss")


def test_megatron_bert_pretraining_vs_transformers():
    ref = H.build_megatron_bert(H.BERT_SMALL)
    mine = MegatronBertForPreTraining(ref.config, device="cuda")
    mine.load_reference_state_dict(ref.state_dict())
    _compare(ref, mine, H.make_mlm_batch(H.BERT_SMALL["vocab_size"], 3, 96, seed=6, nsp=True, pad_tail=11),
             "megatron_loss")


def test_megatron_bert_gelu_new_variant():
    # workspace/erlangshen-bert-base/pretrain/config.json uses hidden_act = gelu_new
    ref = H.build_megatron_bert(H.BERT_SMALL, hidden_act="gelu_new")
    mine = MegatronBertForPreTraining(ref.config, device="cuda")
    mine.load_reference_state_dict(ref.state_dict())
    b = H.make_mlm_batch(H.BERT_SMALL["vocab_size"], 2, 64, seed=7, nsp=True)
    out_ref = ref(**b)
    out = mine(**{k: v.cuda() for k, v in b.items()})
    assert abs(out.loss.item() - out_ref.loss.item()) <= 4e-3
