"""GPU parity of the fsb200 mT5 step (BASELINE config 5) against the implementation the reference calls:
`transformers.MT5ForConditionalGeneration` (fengshen/examples/pretrain_t5/pretrain_t5.py:57-59,81-87), run on CPU in fp32 with
eager attention and dropout 0, and against the golden values in tests/golden/mt5_small.npz produced by
oracle/make_golden_hf.py from the same library (transformers 5.5.0; the reference pins only >=4.17.0 — parity unpinned by the
reference, pinned by us). Tolerances as in test_gpt2_gpu.py (bf16 activations vs an fp32 reference)."""
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
from fsb200.models.t5 import MT5ForConditionalGeneration  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden", "mt5_small.npz")


def _mine(ref, **kw):
    m = MT5ForConditionalGeneration(ref.config, device="cuda", **kw)
    m.load_reference_state_dict(ref.state_dict())
    return m


def _cuda(b):
    return {k: v.cuda() for k, v in b.items()}


def _loss_close(got, want):
    """HF's MT5 init draws the (tied) embedding / LM head from N(0, 1) and applies no d^-0.5 rescale, so a random-init model has
    logits of O(100) and a loss of O(150): the bf16 rounding of such logits (2^-8 relative) is what bounds the agreement.
    Tolerance: 3e-3 absolute (the LLaMA / GPT-2 bar at loss ~6) + 5e-4 relative."""
    return abs(got - want) <= 3e-3 + 5e-4 * abs(want)


def test_mt5_forward_backward_vs_transformers_and_golden():
    g = np.load(GOLD)
    ref = H.build_mt5(H.MT5_SMALL)
    batch = H.make_t5_batch(H.MT5_SMALL["vocab_size"], 2, 96, 48, seed=7, pad_tail=13)
    out_ref = ref(**batch)
    assert abs(out_ref.loss.item() - float(g["loss"])) < 1e-5          # live HF == committed golden
    assert np.abs(out_ref.logits.detach()[:, :, :64].numpy() - g["logits_slice"]).max() < 1e-4
    out_ref.loss.backward()
    mine = _mine(ref)
    out = mine(**_cuda(batch), return_logits=True)
    assert _loss_close(out.loss.item(), out_ref.loss.item()), (out.loss.item(), out_ref.loss.item())
    tol = 4 * 2.0 ** -8 * out_ref.logits.abs().max().item()
    assert (out.logits.float().cpu() - out_ref.logits).abs().max().item() <= tol
    out.loss.backward()
    torch.cuda.synchronize()
    ref_params = dict(ref.named_parameters())
    checked = 0
    for name, prm in mine.named_parameters():
        got = prm.main_grad.float().cpu().flatten()
        want = ref_params[name].grad.flatten()
        cos = torch.dot(got, want) / (got.norm() * want.norm() + 1e-30)
        assert cos.item() >= 0.998, (name, cos.item())
        assert abs(got.norm().item() / (want.norm().item() + 1e-30) - 1.0) <= 0.03, name
        assert abs(want.norm().item() - float(g["gradnorm/" + name])) <= 1e-4 * max(1.0, float(g["gradnorm/" + name])), name
        checked += 1
    # every parameter of the HF model has a counterpart (embed_tokens / the tied lm_head are aliases of shared.weight)
    assert checked == len(ref_params) == 51


def test_mt5_untied_head_variant():
    """Older transformers releases (and the official mT5 checkpoints) keep lm_head separate from the shared embedding; the
    installed 5.5 ties it (configuration_mt5.py). The untied path is checked against the same HF model with the head weight
    cloned: logits and loss only depend on which matrix the head multiplies by."""
    import copy
    ref = H.build_mt5(H.MT5_SMALL)
    batch = H.make_t5_batch(H.MT5_SMALL["vocab_size"], 2, 64, 32, seed=21)
    cfg = copy.copy(ref.config)
    cfg.tie_word_embeddings = False
    mine = MT5ForConditionalGeneration(cfg, device="cuda")
    sd = dict(ref.state_dict())
    sd["lm_head.weight"] = ref.shared.weight.detach().clone()
    mine.load_reference_state_dict(sd)
    assert "lm_head.weight" in dict(mine.named_parameters())
    out = mine(**_cuda(batch))
    assert _loss_close(out.loss.item(), ref(**batch).loss.item())
    out.loss.backward()
    g_head = mine.P("lm_head.weight").main_grad.float()
    g_emb = mine.P("shared.weight").main_grad.float()
    ref.zero_grad(); ref(**batch).loss.backward()
    tot = (g_head + g_emb).cpu().flatten(); want = ref.shared.weight.grad.flatten()
    assert (torch.dot(tot, want) / (tot.norm() * want.norm())).item() >= 0.998


@pytest.mark.parametrize("se,sd", [(128, 128), (200, 57)])
def test_mt5_shapes_and_shift_right(se, sd):
    ref = H.build_mt5(H.MT5_SMALL, seed=3)
    batch = H.make_t5_batch(H.MT5_SMALL["vocab_size"], 3, se, sd, seed=11, pad_tail=0)
    out_ref = ref(**batch)
    mine = _mine(ref)
    out = mine(**_cuda(batch))
    assert _loss_close(out.loss.item(), out_ref.loss.item()), (out.loss.item(), out_ref.loss.item())
    # integer side: decoder inputs are bit-exact with HF's _shift_right
    want = ref._shift_right(batch["labels"])
    assert torch.equal(mine._shift_right(batch["labels"].cuda()).cpu(), want)


def test_mt5_training_curve_tracks_transformers():
    """12 AdamW steps (fp32 master weights, weight decay by name as fengshen/models/model_utils.py:39-47) against the same
    loop on the HF model in fp32 on CPU."""
    from fsb200.flat import is_no_decay
    ref = H.build_mt5(H.MT5_SMALL)
    mine = _mine(ref)
    V = H.MT5_SMALL["vocab_size"]
    batches = [H.make_t5_batch(V, 2, 64, 32, seed=40 + i) for i in range(4)]
    named = list(ref.named_parameters())
    opt = torch.optim.AdamW([{"params": [p for n, p in named if not is_no_decay(n)], "weight_decay": 0.1},
                             {"params": [p for n, p in named if is_no_decay(n)], "weight_decay": 0.0}], lr=1e-3)
    eng = ZeroEngine(mine, lr=1e-3, weight_decay=0.1)
    want, got = [], []
    for it in range(12):
        b = batches[it % 4]
        loss = ref(**b).loss
        opt.zero_grad(set_to_none=True)
        loss.backward()
        opt.step()
        want.append(loss.item())
        out = mine(**_cuda(b))
        out.loss.backward()
        eng.backward_done()
        eng.step()
        got.append(out.loss.item())
    err = (np.abs(np.array(got) - np.array(want)) / np.abs(np.array(want))).max()
    assert got[-1] < 0.5 * got[0] and err <= 1e-2, (err, got, want)   # relative: the curve runs from ~160 down to ~25


def test_mt5_grad_accumulation_rotating_slots_match_full_batch():
    """ZeRO-2 with GA: the enc*/dec* layer buckets share rotating gradient slots; two micro-batches of 2 must give the same
    update as what the fp32 accumulation of their gradients implies (checked against a one-shot batch of 4)."""
    ref = H.build_mt5(dict(H.MT5_SMALL, num_layers=3, num_decoder_layers=3))   # >= 3 layers per stack: the slots rotate
    V = H.MT5_SMALL["vocab_size"]
    b4 = H.make_t5_batch(V, 4, 64, 32, seed=90)
    one = _mine(ref)
    e1 = ZeroEngine(one, lr=1e-3, weight_decay=0.1, grad_clip=1.0)
    out = one(**_cuda(b4))
    out.loss.backward(); e1.backward_done(); e1.step()
    two = _mine(ref)
    e2 = ZeroEngine(two, lr=1e-3, weight_decay=0.1, grad_clip=1.0, ga_steps=2)
    assert two.flat.grads.numel() < two.flat.total          # rotating slots in use
    for half in (slice(0, 2), slice(2, 4)):
        mb = {k: v[half] for k, v in b4.items()}
        # the mean over valid labels differs per micro-batch; equalise by using the same number of valid labels (make_t5_batch
        # masks the last 3 positions of every row)
        o = two(**_cuda(mb))
        o.loss.backward(); e2.backward_done()
    e2.step()
    d = (one.flat.params.float() - two.flat.params.float()).abs().max().item()
    assert d <= 2e-2, d   # bf16 parameters after one lr=1e-3 step: one ulp at |w| ~ 2 is 1.6e-2


def test_mt5_save_pretrained_is_readable_by_transformers(tmp_path):
    """pretrain_t5.py:105-112 ends a run with `self.model.save_pretrained(...)`: the exported directory must load into the HF class
    the reference uses, and give the same loss there (CPU, fp32 arithmetic on the exported bf16 weights)."""
    from transformers import MT5ForConditionalGeneration as HFMT5
    from fsb200.models.export import from_pretrained
    ref = H.build_mt5(H.MT5_SMALL, seed=5)
    mine = _mine(ref)
    batch = H.make_t5_batch(H.MT5_SMALL["vocab_size"], 2, 64, 32, seed=77)
    mine.save_pretrained(str(tmp_path / "exp"))
    hf = HFMT5.from_pretrained(str(tmp_path / "exp"), torch_dtype=torch.float32, attn_implementation="eager")
    hf.eval()
    with torch.no_grad():
        want = ref(**batch).loss.item()
        got_hf = hf(**batch).loss.item()
    assert abs(got_hf - want) <= 1e-4 * max(1.0, abs(want)), (got_hf, want)
    back = from_pretrained(MT5ForConditionalGeneration, str(tmp_path / "exp"), config_cls=type(ref.config), device="cuda")
    assert torch.equal(back.flat.params, mine.flat.params)
