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


BERT_SMALL = dict(vocab_size=512, hidden_size=256, num_hidden_layers=2, num_attention_heads=4, intermediate_size=512,
                  max_position_embeddings=128, type_vocab_size=2)


def build_bert(cfg, seed=0, bf16_exact=True):
    """BertForMaskedLM as examples/pretrain_bert/pretrain_bert.py:135-137 builds it (config -> random init)."""
    from transformers import BertConfig, BertForMaskedLM
    torch.manual_seed(seed)
    config = BertConfig(hidden_dropout_prob=0.0, attention_probs_dropout_prob=0.0, hidden_act="gelu",
                        attn_implementation="eager", **cfg)
    model = BertForMaskedLM(config)
    model.train()
    return _bf16_exact_(model) if bf16_exact else model


def build_megatron_bert(cfg, seed=0, bf16_exact=True, hidden_act="gelu"):
    """MegatronBertForPreTraining as examples/pretrain_erlangshen_bert/pretrain_erlangshen.py:138-141 builds it."""
    from transformers import MegatronBertConfig, MegatronBertForPreTraining
    torch.manual_seed(seed)
    config = MegatronBertConfig(hidden_dropout_prob=0.0, attention_probs_dropout_prob=0.0, hidden_act=hidden_act,
                                attn_implementation="eager", **cfg)
    model = MegatronBertForPreTraining(config)
    model.train()
    return _bf16_exact_(model) if bf16_exact else model


def make_mlm_batch(V, B, S, seed=1234, nsp=False, pad_tail=0):
    """Synthetic MLM batch: ids uniform in [1, V), labels = -100 except a Bernoulli(0.15) subset where
    labels = ids; token types 0/1 split at the middle; optional NSP labels and a padded tail on the last sample."""
    rs = np.random.RandomState(seed)
    ids = torch.from_numpy(rs.randint(1, V, size=(B, S)).astype(np.int64))
    sel = torch.from_numpy(rs.rand(B, S) < 0.15)
    labels = torch.where(sel, ids, torch.full_like(ids, -100))
    tt = torch.zeros_like(ids)
    tt[:, S // 2:] = 1
    am = torch.ones_like(ids)
    if pad_tail:
        am[-1, S - pad_tail:] = 0
        labels[-1, S - pad_tail:] = -100
    b = {"input_ids": ids, "attention_mask": am, "token_type_ids": tt, "labels": labels}
    if nsp:
        b["next_sentence_label"] = torch.from_numpy(rs.randint(0, 2, size=(B,)).astype(np.int64))
    return b


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


# ------------------------------------------------------------------------------------------------ C5: mT5 (round-2 row)
# examples/pretrain_t5/pretrain_t5.py:57-59 builds transformers.MT5ForConditionalGeneration(MT5Config); training_step
# (:81-87) is `self.model(input_ids=..., labels=...)` — HF shifts the labels right itself, T5LayerNorm is an RMSNorm without
# mean subtraction, attention is UNSCALED with an additive relative-position bias shared by all layers of a stack, the FFN is
# gated-GELU (wi_0, wi_1, wo), nothing has a bias. NOTE: transformers 5.x forces tie_word_embeddings=True for MT5 configs
# (configuration_mt5.py __post_init__), so the model built here — exactly what the reference script would get with this library —
# has lm_head tied to `shared` (51 parameters) and no d^-0.5 rescale of the decoder output; the keyword below is ignored by 5.x.
MT5_SMALL = dict(vocab_size=512, d_model=256, d_kv=64, d_ff=512, num_layers=2, num_decoder_layers=2, num_heads=4,
                 relative_attention_num_buckets=32, relative_attention_max_distance=128)


def build_mt5(cfg, seed=0, bf16_exact=True):
    from transformers import MT5Config, MT5ForConditionalGeneration
    torch.manual_seed(seed)
    config = MT5Config(dropout_rate=0.0, feed_forward_proj="gated-gelu", tie_word_embeddings=False,
                       attn_implementation="eager", decoder_start_token_id=0, pad_token_id=0, **cfg)
    model = MT5ForConditionalGeneration(config)
    model.train()
    return _bf16_exact_(model) if bf16_exact else model


def make_t5_batch(V, B, S_enc, S_dec, seed=1234, pad_tail=0):
    """Synthetic span-corruption-shaped batch: encoder ids + attention mask, decoder labels with -100 on the padded tail."""
    rng = np.random.RandomState(seed)
    ids = rng.randint(2, V, size=(B, S_enc)).astype(np.int64)
    am = np.ones((B, S_enc), dtype=np.int64)
    if pad_tail:
        ids[-1, -pad_tail:] = 0
        am[-1, -pad_tail:] = 0
    labels = rng.randint(2, V, size=(B, S_dec)).astype(np.int64)
    labels[:, -3:] = -100
    return {"input_ids": torch.from_numpy(ids), "attention_mask": torch.from_numpy(am), "labels": torch.from_numpy(labels)}


def t5_relative_position_bucket(rel, bidirectional, num_buckets=32, max_distance=128):
    """Restatement (numpy, integer in / integer out) of transformers' T5Attention._relative_position_bucket
    (models/t5/modeling_t5.py, static method): rel = key_pos - query_pos. This is the function the CUDA attention kernels
    will evaluate per (q, k) pair to index the [num_buckets, heads] bias table; tests pin it to HF bit-exactly."""
    rel = np.asarray(rel, dtype=np.int64)
    out = np.zeros_like(rel)
    nb = num_buckets
    if bidirectional:
        nb //= 2
        out += (rel > 0).astype(np.int64) * nb
        rel = np.abs(rel)
    else:
        rel = -np.minimum(rel, 0)
    max_exact = nb // 2
    is_small = rel < max_exact
    # float32 arithmetic exactly as torch does it (log in fp32, truncation towards zero)
    relf = np.maximum(rel, 1).astype(np.float32)
    large = max_exact + (np.log(relf / np.float32(max_exact)) / np.float32(np.log(max_distance / max_exact))
                         * np.float32(nb - max_exact)).astype(np.int64)
    large = np.minimum(large, nb - 1)
    return out + np.where(is_small, rel, large)
