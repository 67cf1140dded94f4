"""TEST INFRASTRUCTURE — CPU restatement (plain PyTorch, fp32 unless told otherwise) of the reference's Ziya-LLaMA
training step. Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import
this; the product path (fengshen-lm_b200/) never does.

Pinned against the UNMODIFIED reference: tests/golden/llama_*.npz are produced by oracle/make_golden.py from
/root/reference (fengshen.models.llama.modeling_llama.LlamaForCausalLM, attention_config="global") and
tests/test_oracle.py checks this file against them (max |delta| <= 1e-5 on logits, 1e-6 on loss in fp32).

Each function cites the reference lines it follows.
"""
import math

import numpy as np
import torch
import torch.nn.functional as F


def ff_dim(hidden_size, multiple_of=256):
    """LLaMAParallelMLP.__init__, fengshen/models/megatron/layers/transformer.py:589-590."""
    ff = int(2 * hidden_size * 4 / 3)
    return multiple_of * ((ff + multiple_of - 1) // multiple_of)


def param_shapes(V, h, L):
    """State-dict keys and shapes of the reference model (utils/llama_convert/hf_to_fs.py:136-147)."""
    ff = ff_dim(h)
    out = [("llama.embed_in.word_embeddings.weight", (V, h))]
    for i in range(L):
        p = f"llama.layers.{i}."
        out += [(p + "input_layernorm.scale", (h,)), (p + "attention.query_key_value.weight", (3 * h, h)),
                (p + "attention.dense.weight", (h, h)), (p + "post_attention_layernorm.scale", (h,)),
                (p + "mlp.w1.weight", (ff, h)), (p + "mlp.w3.weight", (ff, h)), (p + "mlp.w2.weight", (h, ff))]
    out += [("llama.final_layer_norm.scale", (h,)), ("embed_out.final_linear.weight", (V, h))]
    return out


def make_weights(V, h, L, seed=0, bf16_exact=True):
    """Deterministic, platform-independent synthetic weights (numpy RandomState) with the reference's init scales
    (layers/init_functions.py:121-142: small_init std sqrt(2/(5h)), wang_init std 2/(L sqrt(h)); norm scales near 1).
    With bf16_exact the values are rounded to bf16 so that a bf16 GPU model and the fp32 oracle hold identical weights."""
    rs = np.random.RandomState(seed)
    small, wang = math.sqrt(2.0 / (5.0 * h)), 2.0 / (L * math.sqrt(h))
    sd = {}
    for name, shape in param_shapes(V, h, L):
        if name.endswith(".scale"):
            w = 1.0 + 0.05 * rs.standard_normal(shape)
        elif name.endswith("dense.weight") or name.endswith("w2.weight"):
            w = wang * rs.standard_normal(shape)
        else:
            w = small * rs.standard_normal(shape)
        t = torch.from_numpy(w.astype(np.float32))
        sd[name] = t.to(torch.bfloat16).to(torch.float32) if bf16_exact else t
    return sd


def make_batch(V, B, S, seed=1234):
    """Synthetic causal-LM batch as LlamaSFTCollator emits it (examples/ziya_llama/finetune_ziya_llama.py:79-84):
    labels = input_ids, attention_mask = 1, position_ids = arange."""
    rs = np.random.RandomState(seed)
    ids = torch.from_numpy(rs.randint(0, V, size=(B, S)).astype(np.int64))
    return {"input_ids": ids, "labels": ids.clone(), "attention_mask": torch.ones_like(ids),
            "position_ids": torch.arange(S, dtype=torch.int64)[None].expand(B, S).contiguous()}


# ------------------------------------------------------------------------------------------------------- model
def rmsnorm(x, scale, eps):
    """RMSNorm.forward, layers/norms.py:44-52: fp32 variance, cast to the scale's 16-bit dtype (if any), then scale."""
    var = x.to(torch.float32).pow(2).mean(-1, keepdim=True)
    y = x * torch.rsqrt(var + eps)
    if scale.dtype in (torch.float16, torch.bfloat16):
        y = y.to(scale.dtype)
    return scale * y


def rope_tables(hn, max_pos, base=10000.0):
    """RotaryEmbedding.__init__, layers/positional_embeddings.py:38-52."""
    inv_freq = 1.0 / (base ** (torch.arange(0, hn, 2).float() / hn))
    t = torch.arange(max_pos, dtype=inv_freq.dtype)
    freqs = torch.einsum("i,j->ij", t, inv_freq)
    emb = torch.cat((freqs, freqs), dim=-1)
    return emb.cos(), emb.sin()


def rotate_half(x):
    """layers/positional_embeddings.py:71-75."""
    x1, x2 = x[..., : x.shape[-1] // 2], x[..., x.shape[-1] // 2:]
    return torch.cat((-x2, x1), dim=-1)


def attention(x, wqkv, wd, pos, nh, cos, sin, hn=None):
    """ParallelSelfAttention.forward, layers/transformer.py:476-568 with the `global` core (:307-408):
    per-head interleaved QKV split (:488-497), rotary on q,k gathered by position_ids
    (positional_embeddings.py:78-87), scores = q k^T / sqrt(hn) (:338-344), causal masked softmax, context, dense."""
    B, S, h = x.shape
    hn = h // nh if hn is None else hn      # hn given: a tensor-parallel shard (nh = LOCAL heads, wd = [h, nh * hn] row-parallel slice)
    mixed = F.linear(x, wqkv).view(B, S, nh, 3 * hn)
    q, k, v = mixed[..., :hn], mixed[..., hn:2 * hn], mixed[..., 2 * hn:]
    c = cos[pos].to(x.dtype)[:, :, None, :]  # [B,S,1,hn]
    s = sin[pos].to(x.dtype)[:, :, None, :]
    q = q * c + rotate_half(q) * s
    k = k * c + rotate_half(k) * s
    scores = torch.einsum("bqhd,bkhd->bhqk", q, k) * (1.0 / math.sqrt(hn))
    causal = torch.triu(torch.ones(S, S, dtype=torch.bool), diagonal=1)
    scores = scores.masked_fill(causal, torch.finfo(scores.dtype).min)  # modeling_llama.py:71-74
    probs = torch.softmax(scores, dim=-1)
    ctx = torch.einsum("bhqk,bkhd->bqhd", probs, v).reshape(B, S, nh * hn)
    return F.linear(ctx, wd)


def mlp(x, w1, w3, w2):
    """LLaMAParallelMLP.forward, layers/transformer.py:620-623."""
    return F.linear(F.silu(F.linear(x, w1)) * F.linear(x, w3), w2)


def forward(sd, batch, nh, eps=1e-6, max_pos=2048):
    """LlamaForCausalLM.forward (modeling_llama.py:272-351) = LlamaModel.forward (:135-236) + LM head (:332) +
    shifted mean cross-entropy (:334-339). Pre-norm residual block per ParallelTransformerLayer.forward (:753-797)."""
    ids, pos, labels = batch["input_ids"], batch["position_ids"], batch.get("labels")
    L = 1 + max(int(k.split(".")[2]) for k in sd if k.startswith("llama.layers."))
    x = F.embedding(ids, sd["llama.embed_in.word_embeddings.weight"])
    h = x.shape[-1]
    cos, sin = rope_tables(h // nh, max_pos)
    for i in range(L):
        p = f"llama.layers.{i}."
        a = attention(rmsnorm(x, sd[p + "input_layernorm.scale"], eps), sd[p + "attention.query_key_value.weight"],
                      sd[p + "attention.dense.weight"], pos, nh, cos, sin)
        x = a + x                                                     # transformer.py:775-778 (dropout p = 0)
        m = mlp(rmsnorm(x, sd[p + "post_attention_layernorm.scale"], eps), sd[p + "mlp.w1.weight"],
                sd[p + "mlp.w3.weight"], sd[p + "mlp.w2.weight"])
        x = m + x                                                     # transformer.py:785-788
    x = rmsnorm(x, sd["llama.final_layer_norm.scale"], eps)
    logits = F.linear(x, sd["embed_out.final_linear.weight"])
    loss = None
    if labels is not None:
        V = logits.shape[-1]
        loss = F.cross_entropy(logits[:, :-1, :].reshape(-1, V), labels[:, 1:].reshape(-1))
    return loss, logits


# ------------------------------------------------------------------------------------------------------- optimiser
NO_DECAY = ['bias', 'LayerNorm.bias', 'LayerNorm.weight', 'layer_norm.', 'layernorm.']  # model_utils.py:40


def param_groups(named_params, weight_decay):
    """get_default_update_params, fengshen/models/model_utils.py:39-47 (grouping by NAME substrings)."""
    named = list(named_params)
    return [{"params": [p for n, p in named if not any(nd in n for nd in NO_DECAY)], "weight_decay": weight_decay},
            {"params": [p for n, p in named if any(nd in n for nd in NO_DECAY)], "weight_decay": 0.0}]


def polynomial_lr(step, base_lr, warmup, total, lr_end=1e-7, power=1.0):
    """transformers.optimization.get_polynomial_decay_schedule_with_warmup (selected at model_utils.py:250-252)."""
    if step < warmup:
        return base_lr * float(step) / float(max(1, warmup))
    if step > total:
        return lr_end
    span, remaining = total - warmup, 1 - (step - warmup) / (total - warmup)
    return (base_lr - lr_end) * remaining ** power + lr_end if span > 0 else lr_end


def train(sd, batches, nh, steps, lr=1e-3, betas=(0.9, 0.95), eps_adam=1e-8, weight_decay=0.1, warmup_ratio=0.1,
          lr_end=1e-7, eps=1e-6):
    """configure_optimizers (model_utils.py:50-98) restated with torch.optim.AdamW (== FusedAdam adam_w_mode=True) and
    the per-step polynomial schedule; one optimizer step per batch. Returns the list of losses (before each update)."""
    params = {k: torch.nn.Parameter(v.clone()) for k, v in sd.items()}
    opt = torch.optim.AdamW(param_groups(params.items(), weight_decay), lr=lr, betas=betas, eps=eps_adam)
    warmup = warmup_ratio * steps
    losses = []
    for it in range(steps):
        cur = polynomial_lr(it, lr, warmup, steps, lr_end)
        for g in opt.param_groups:
            g["lr"] = cur
        loss, _ = forward(params, batches[it % len(batches)], nh, eps)
        opt.zero_grad(set_to_none=True)
        loss.backward()
        opt.step()
        losses.append(float(loss.detach()))
    return losses, {k: v.detach() for k, v in params.items()}
