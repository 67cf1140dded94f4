"""TEST INFRASTRUCTURE — the HF-backed configs (C1 BERT, C2 GPT-2, C3 MegatronBERT, C5 mT5) have no in-tree
arithmetic: the reference's example scripts call `transformers` model classes directly
(examples/wenzhong_qa/finetune_wenzhong.py:56, examples/pretrain_bert/pretrain_bert.py:137,
examples/pretrain_erlangshen_bert/pretrain_erlangshen.py:141, examples/pretrain_t5/pretrain_t5.py:57-59). The oracle for
them is therefore the installed third-party library itself (transformers 5.5.0; the reference pins only `>=4.17.0`,
setup.py:17), run on CPU in fp32 with eager attention and dropout 0, plus the optimizer /
schedule the scripts build around it, restated below with file:line. Only tests/, smoke() and bench.py's CPU-baseline
legs import this module.
"""
import numpy as np
import torch

GPT2_SMALL = dict(vocab_size=512, n_positions=128, n_embd=256, n_layer=2, n_head=4)
GPT2_110M = dict(vocab_size=50264, n_positions=1024, n_embd=768, n_layer=12, n_head=12)  # vocab 50257 padded to /8


def _bf16_exact_(model):
    with torch.no_grad():
        for p in model.parameters():
            p.copy_(p.to(torch.bfloat16).to(torch.float32))
    return model


def build_gpt2(cfg, seed=0, bf16_exact=True):
    """GPT2LMHeadModel(config) with HF's own init under a fixed seed; dropout 0; eager attention."""
    from transformers import GPT2Config, GPT2LMHeadModel
    torch.manual_seed(seed)
    config = GPT2Config(resid_pdrop=0.0, embd_pdrop=0.0, attn_pdrop=0.0, activation_function="gelu_new",
                        attn_implementation="eager", **cfg)
    model = GPT2LMHeadModel(config)
    model.train()
    return _bf16_exact_(model) if bf16_exact else model


def make_lm_batch(V, B, S, seed=1234):
    """Synthetic causal-LM batch: labels = input_ids, attention_mask = 1."""
    rs = np.random.RandomState(seed)
    ids = torch.from_numpy(rs.randint(0, V, size=(B, S)).astype(np.int64))
    return {"input_ids": ids, "labels": ids.clone(), "attention_mask": torch.ones_like(ids)}


def wenzhong_param_groups(named_params, weight_decay):
    """GPT2FinetuneMedicalQA.configure_optimizers, examples/wenzhong_qa/finetune_wenzhong.py:89-100."""
    no_decay = ['bias', 'LayerNorm.bias', 'LayerNorm.weight']
    named = [(n, p) for n, p in named_params if p.requires_grad]
    return [{"params": [p for n, p in named if not any(nd in n for nd in no_decay)], "weight_decay": weight_decay},
            {"params": [p for n, p in named if any(nd in n for nd in no_decay)], "weight_decay": 0.0}]


def linear_lr(step, base_lr, warmup, total):
    """transformers.get_linear_schedule_with_warmup (finetune_wenzhong.py:102-104)."""
    if step < warmup:
        return base_lr * float(step) / float(max(1, warmup))
    return base_lr * max(0.0, float(total - step) / float(max(1, total - warmup)))


def train_gpt2(model, batches, steps, lr=1e-3, weight_decay=0.1, warmup=2):
    """torch.optim.AdamW + linear schedule exactly as finetune_wenzhong.py:89-113 wires them; returns the loss curve."""
    opt = torch.optim.AdamW(wenzhong_param_groups(model.named_parameters(), weight_decay), lr=lr)
    losses = []
    for it in range(steps):
        cur = linear_lr(it, lr, warmup, steps)
        for g in opt.param_groups:
            g["lr"] = cur
        b = batches[it % len(batches)]
        loss = model(input_ids=b["input_ids"], attention_mask=b["attention_mask"], labels=b["labels"]).loss
        opt.zero_grad(set_to_none=True)
        loss.backward()
        opt.step()
        losses.append(float(loss.detach()))
    return losses
