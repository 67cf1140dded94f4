"""HF <-> Fengshen LLaMA weight-layout converters (behaviour of fengshen/utils/llama_convert/hf_to_fs.py:84-150 and
fs_to_hf.py:40-108), restated as pure state-dict functions so that they run without instantiating either model:

  hf_to_fs_state_dict : transformers LlamaForCausalLM keys -> the reference's keys. q/k/v projections are stacked PER HEAD into
                        the interleaved `attention.query_key_value.weight` ([heads, {q,k,v}, head_dim] rows,
                        layers/transformer.py:488-497; hf_to_fs.py:126-130); gate/up/down -> mlp.w1/w3/w2; RMSNorm weight ->
                        `.scale`. Both sides use the half-rotation RoPE convention, so nothing is permuted (the reference
                        defines `permute_rotary` and never calls it, hf_to_fs.py:110-114).
  fs_to_hf_state_dict : the inverse (fs_to_hf.py:73-98).
`save_pretrained_fs` writes config.json + pytorch_model.bin the way `LlamaForCausalLM.from_pretrained` (compat) reads them —
the HF-style export the training scripts call at the end of a run."""
import json
import os

import torch


def hf_to_fs_state_dict(hf_sd, num_heads):
    out = {"llama.embed_in.word_embeddings.weight": hf_sd["model.embed_tokens.weight"],
           "embed_out.final_linear.weight": hf_sd["lm_head.weight"],
           "llama.final_layer_norm.scale": hf_sd["model.norm.weight"]}
    n_layers = 1 + max(int(k.split(".")[2]) for k in hf_sd if k.startswith("model.layers."))
    for i in range(n_layers):
        h, f = f"model.layers.{i}.", f"llama.layers.{i}."
        q, k, v = (hf_sd[h + f"self_attn.{n}_proj.weight"] for n in "qkv")
        hidden = q.shape[1]
        hn = q.shape[0] // num_heads
        qkv = torch.stack([w.view(num_heads, hn, hidden) for w in (q, k, v)], dim=1)      # [heads, 3, hn, hidden]
        out[f + "attention.query_key_value.weight"] = qkv.reshape(num_heads * 3 * hn, hidden).clone()
        out[f + "attention.dense.weight"] = hf_sd[h + "self_attn.o_proj.weight"].clone()
        out[f + "mlp.w1.weight"] = hf_sd[h + "mlp.gate_proj.weight"].clone()
        out[f + "mlp.w3.weight"] = hf_sd[h + "mlp.up_proj.weight"].clone()
        out[f + "mlp.w2.weight"] = hf_sd[h + "mlp.down_proj.weight"].clone()
        out[f + "input_layernorm.scale"] = hf_sd[h + "input_layernorm.weight"].clone()
        out[f + "post_attention_layernorm.scale"] = hf_sd[h + "post_attention_layernorm.weight"].clone()
    return out


def fs_to_hf_state_dict(fs_sd, num_heads):
    out = {"model.embed_tokens.weight": fs_sd["llama.embed_in.word_embeddings.weight"],
           "lm_head.weight": fs_sd["embed_out.final_linear.weight"],
           "model.norm.weight": fs_sd["llama.final_layer_norm.scale"]}
    n_layers = 1 + max(int(k.split(".")[2]) for k in fs_sd if k.startswith("llama.layers."))
    for i in range(n_layers):
        h, f = f"model.layers.{i}.", f"llama.layers.{i}."
        qkv = fs_sd[f + "attention.query_key_value.weight"]
        hidden = qkv.shape[1]
        hn = qkv.shape[0] // (3 * num_heads)
        q, k, v = qkv.view(num_heads, 3, hn, hidden).unbind(1)
        out[h + "self_attn.q_proj.weight"] = q.reshape(num_heads * hn, hidden).clone()
        out[h + "self_attn.k_proj.weight"] = k.reshape(num_heads * hn, hidden).clone()
        out[h + "self_attn.v_proj.weight"] = v.reshape(num_heads * hn, hidden).clone()
        out[h + "self_attn.o_proj.weight"] = fs_sd[f + "attention.dense.weight"].clone()
        out[h + "mlp.gate_proj.weight"] = fs_sd[f + "mlp.w1.weight"].clone()
        out[h + "mlp.up_proj.weight"] = fs_sd[f + "mlp.w3.weight"].clone()
        out[h + "mlp.down_proj.weight"] = fs_sd[f + "mlp.w2.weight"].clone()
        out[h + "input_layernorm.weight"] = fs_sd[f + "input_layernorm.scale"].clone()
        out[h + "post_attention_layernorm.weight"] = fs_sd[f + "post_attention_layernorm.scale"].clone()
    return out


def save_pretrained_fs(state_dict, config, path):
    """config.json + pytorch_model.bin in the reference's key layout (what `from_pretrained` reads back)."""
    os.makedirs(path, exist_ok=True)
    cfg = config if isinstance(config, dict) else {k: v for k, v in vars(config).items() if not k.startswith("_")}
    cfg = {k: v for k, v in cfg.items() if isinstance(v, (int, float, str, bool, list, type(None)))}
    with open(os.path.join(path, "config.json"), "w") as f:
        json.dump(cfg, f, indent=1)
    torch.save({k: v.detach().cpu() for k, v in state_dict.items()}, os.path.join(path, "pytorch_model.bin"))


def split_state_dict_tp(fs_sd, tp, num_heads):
    """The shard layout of utils/llama_convert/convert_fs_llama_tp.py:143-181 (`part_{rank}` directories): embedding and LM head
    split along the vocabulary (dim 0); query_key_value viewed as [tp, heads/tp * 3 * head_dim, hidden] (whole heads per rank,
    interleave kept); dense and w2 split along their INPUT dim (dim 1, row-parallel); w1 / w3 along dim 0 (column-parallel);
    norms and inv_freq duplicated. Returns a list of `tp` state dicts."""
    out = [dict() for _ in range(tp)]
    for k, v in fs_sd.items():
        if k in ("llama.embed_in.word_embeddings.weight", "embed_out.final_linear.weight"):
            parts = v.chunk(tp, dim=0)
        elif "query_key_value" in k:
            parts = v.view(tp, v.shape[0] // tp, *v.shape[1:]).unbind(0)
        elif k.endswith("attention.dense.weight") or k.endswith("mlp.w2.weight"):
            parts = v.chunk(tp, dim=1)
        elif k.endswith("mlp.w1.weight") or k.endswith("mlp.w3.weight"):
            parts = v.chunk(tp, dim=0)
        else:   # layernorm scales, rotary inv_freq: duplicated
            parts = [v] * tp
        for r in range(tp):
            out[r][k] = parts[r].clone()
    return out


def merge_state_dict_tp(shards, num_heads):
    """Inverse of split_state_dict_tp (what fs_merge_weight.py does for released shards)."""
    tp = len(shards)
    out = {}
    for k in shards[0]:
        vs = [s[k] for s in shards]
        if k in ("llama.embed_in.word_embeddings.weight", "embed_out.final_linear.weight") or "query_key_value" in k \
                or k.endswith("mlp.w1.weight") or k.endswith("mlp.w3.weight"):
            out[k] = torch.cat(vs, dim=0)
        elif k.endswith("attention.dense.weight") or k.endswith("mlp.w2.weight"):
            out[k] = torch.cat(vs, dim=1)
        else:
            out[k] = vs[0].clone()
    return out
