"""GPU parity at the BASELINE shapes (round 1 stopped at h=512 / s=320).

  C4  Ziya-LLaMA-13B: ONE full-width layer (h=5120, 40 heads x 128, ff=13824, s=2048, b=1; SURVEY.md §6: a full-width
      layer is what the host can run) between a reduced-vocabulary embedding and head — against the pinned CPU oracle
      (oracle/llama_oracle.py, fp32) on bf16-exact weights: loss, logits, every parameter gradient.
  C2  Wenzhong-GPT2 at its real width / sequence / vocabulary (h=768, 12 heads, s=1024, V=50264) with 2 layers — against
      transformers.GPT2LMHeadModel on CPU (the class the reference calls, finetune_wenzhong.py:56).
  C1  Erlangshen-BERT-base at its real width / sequence / vocabulary (h=768, s=128, V=21128, batch 8) with 2 layers —
      against transformers.BertForMaskedLM (pretrain_bert.py:137).
The CPU sides take 10-40 s each on the box's host cores. Tolerances as in the small-shape tests (bf16 activations vs an fp32
reference): loss 3e-3, logits 4 * 2^-8 * max|logit|, per-parameter gradient cosine >= 0.998 / norm within 3 %.
"""
import os
import sys
from types import SimpleNamespace

import pytest
import torch

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import hf_oracle as H  # noqa: E402  (checker only)
import llama_oracle as O  # noqa: E402  (checker only)


def _check_grads(mine_named, ref_grads, cos_min=0.998, ratio_tol=0.03):
    worst = (1.0, "")
    for name, prm in mine_named:
        got = prm.main_grad.float().cpu().flatten()
        want = ref_grads[name].flatten()
        if want.norm().item() < 1e-9:
            continue
        cos = (torch.dot(got, want) / (got.norm() * want.norm() + 1e-30)).item()
        worst = min(worst, (cos, name))
        assert cos >= cos_min, (name, cos)
        assert abs(got.norm().item() / want.norm().item() - 1.0) <= ratio_tol, (name, got.norm().item(), want.norm().item())
    return worst


def test_ziya_llama_13b_full_width_layer_vs_oracle():
    from fsb200.models.llama import LlamaForCausalLM
    torch.set_num_threads(max(1, min(32, os.cpu_count() or 1)))
    V, h, L, nh, B, S = 4096, 5120, 1, 40, 1, 2048
    sd = O.make_weights(V, h, L, seed=0)
    batch = O.make_batch(V, B, S, seed=1234)
    cfg = SimpleNamespace(vocab_size=V, hidden_size=h, num_hidden_layers=L, num_attention_heads=nh, rms_norm_epsilon=1e-6,
                          max_position_embeddings=2048, rotary_emb_base=10000, llama_mlp_multiple_of=256)
    model = LlamaForCausalLM(cfg, device="cuda")
    assert model.ff == 13824 and model.hn == 128                      # the Ziya-13B layer geometry (transformer.py:589-590)
    model.load_reference_state_dict(sd)
    out = model(input_ids=batch["input_ids"].cuda(), position_ids=batch["position_ids"].cuda(), labels=batch["labels"].cuda(),
                return_logits=True)
    out.loss.backward()
    torch.cuda.synchronize()
    osd = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
    oloss, ologits = O.forward(osd, batch, nh)
    oloss.backward()
    assert abs(out.loss.item() - oloss.item()) <= 3e-3, (out.loss.item(), oloss.item())
    tol = 4 * 2.0 ** -8 * ologits.detach().abs().max().item()
    assert (out.logits.float().cpu() - ologits.detach().view_as(out.logits.cpu())).abs().max().item() <= tol
    _check_grads(model.named_parameters(), {k: v.grad for k, v in osd.items()}, cos_min=0.999, ratio_tol=0.02)


def test_wenzhong_gpt2_real_shape_vs_transformers():
    from fsb200.models.gpt2 import GPT2LMHeadModel
    torch.set_num_threads(max(1, min(32, os.cpu_count() or 1)))
    cfg = dict(vocab_size=50264, n_positions=1024, n_embd=768, n_layer=2, n_head=12)
    ref = H.build_gpt2(cfg)
    batch = H.make_lm_batch(cfg["vocab_size"], 2, 1024, seed=77)
    out_ref = ref(input_ids=batch["input_ids"], labels=batch["labels"])
    out_ref.loss.backward()
    mine = GPT2LMHeadModel(ref.config, device="cuda")
    mine.load_reference_state_dict(ref.state_dict())
    out = mine(input_ids=batch["input_ids"].cuda(), labels=batch["labels"].cuda(), return_logits=True)
    assert abs(out.loss.item() - out_ref.loss.item()) <= 3e-3, (out.loss.item(), out_ref.loss.item())
    tol = 4 * 2.0 ** -8 * out_ref.logits.abs().max().item()
    assert (out.logits.float().cpu() - out_ref.logits).abs().max().item() <= tol
    out.loss.backward()
    torch.cuda.synchronize()
    _check_grads(mine.named_parameters(), {n: p.grad for n, p in ref.named_parameters()})


def test_erlangshen_bert_base_real_shape_vs_transformers():
    from fsb200.models.bert import BertForMaskedLM
    torch.set_num_threads(max(1, min(32, os.cpu_count() or 1)))
    cfg = dict(vocab_size=21128, hidden_size=768, num_hidden_layers=2, num_attention_heads=12, intermediate_size=3072,
               max_position_embeddings=512, type_vocab_size=2)
    ref = H.build_bert(cfg)
    batch = H.make_mlm_batch(cfg["vocab_size"], 8, 128, seed=9, pad_tail=17)    # BASELINE configs[0]: seq 128, batch 8
    out_ref = ref(**batch)
    out_ref.loss.backward()
    mine = BertForMaskedLM(ref.config, device="cuda")
    mine.load_reference_state_dict(ref.state_dict())
    out = mine(**{k: v.cuda() for k, v in batch.items()}, return_logits=True)
    assert abs(out.loss.item() - out_ref.loss.item()) <= 4e-3, (out.loss.item(), out_ref.loss.item())
    tol = 4 * 2.0 ** -8 * out_ref.logits.abs().max().item()
    assert (out.logits.float().cpu() - out_ref.logits).abs().max().item() <= tol
    out.loss.backward()
    torch.cuda.synchronize()
    refg = {n: p.grad for n, p in ref.named_parameters()}
    _check_grads([(n, p) for n, p in mine.named_parameters() if refg.get(n) is not None], refg)
