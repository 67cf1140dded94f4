"""Ziya-LLaMA on the fsb200 kernels — drop-in for `fengshen.models.llama.modeling_llama.LlamaForCausalLM`.

Same constructor (`LlamaForCausalLM(config)`), same `forward(input_ids, attention_mask, position_ids, labels)` ->
`CausalLMOutputWithPast`-like result (reference: fengshen/models/llama/modeling_llama.py:272-351), same state-dict key
names (`llama.embed_in.word_embeddings.weight`, `llama.layers.N.attention.query_key_value.weight`, ...,
`embed_out.final_linear.weight`; utils/llama_convert/hf_to_fs.py:136-147), same per-head interleaved QKV weight layout
(layers/transformer.py:488-497). What differs is everything underneath:
  * activations stay [b, s, h] token-major — no [s, b, h] transposes (modeling_llama.py:201,222), no q/k/v repacking;
  * all parameters are views into one flat bf16 buffer, gradients go straight into the parallel flat grad buffer;
  * the whole network is ONE autograd node (`_LlamaStep`): forward runs the hand-scheduled kernel sequence and keeps
    the activations it needs; `loss.backward()` runs the matching hand-written backward (dgrad / wgrad GEMMs, fused
    attention backward, norm backward with the residual-gradient add fused in).
The causal mask is implicit (flash path semantics, transformer.py:441-448: `attention_mask` is not applied; identical to
the `global` path when the mask is all ones — SURVEY.md Appendix B).
"""
import math
from types import SimpleNamespace

import torch
from torch import nn

from .. import lib as L
from .. import ops
from ..flat import FlatBuffers, FlatSpec


def llama_ff_dim(hidden_size, multiple_of=256):
    ff = int(2 * hidden_size * 4 / 3)  # layers/transformer.py:589-590
    return multiple_of * ((ff + multiple_of - 1) // multiple_of)


class _Holder(nn.Module):
    """Bare container so that named_parameters()/state_dict() reproduce the reference's key names."""


def _init_normal_(t, std, gen):
    t.copy_(torch.empty(t.shape, dtype=torch.float32).normal_(0.0, std, generator=gen).to(t.dtype))


class LlamaForCausalLM(nn.Module):
    def __init__(self, config, device=None, world_size=None, seed=0, tp_group=None):
        """tp_group: the tensor-model-parallel process group (mpu.get_model_parallel_group()) or None. With t = its size > 1
        this rank holds the shard the reference's `part_{rank}` checkpoints hold (utils/llama_convert/convert_fs_llama_tp.py
        :143-181): heads / ff columns / vocabulary rows split t ways (ColumnParallelLinear mpu/layers.py:261-360 for QKV,
        w1, w3 and the LM head; RowParallelLinear :363-470 for dense and w2; VocabParallelEmbedding :62-130), norms replicated.
        `world_size` is then the DATA-parallel size (ranks that share a tensor-parallel rank)."""
        super().__init__()
        self.config = config
        import torch.distributed as dist
        self.tp_group = tp_group
        self.tp = dist.get_world_size(tp_group) if tp_group is not None else 1
        self.tp_rank = dist.get_rank(tp_group) if tp_group is not None else 0
        if world_size is None:  # laid out for the job's data-parallel world (the scripts build the model in setup())
            world_size = dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1
            world_size //= self.tp
        dev = torch.device(device if device is not None else f"cuda:{torch.cuda.current_device()}"
                           if torch.cuda.is_available() else "cuda")
        if dev.type != "cuda":
            raise RuntimeError("fsb200 LlamaForCausalLM runs on CUDA only (no CPU fallback on the product path)")
        h, V, nl, nh = config.hidden_size, config.vocab_size, config.num_hidden_layers, config.num_attention_heads
        self.h, self.V, self.nl, self.nh = h, V, nl, nh
        self.hn = h // nh
        self.ff = llama_ff_dim(h, getattr(config, "llama_mlp_multiple_of", 256))
        self.eps = getattr(config, "rms_norm_epsilon", 1e-6)
        if self.hn not in (64, 128):
            raise RuntimeError(f"fsb200: head dim {self.hn} unsupported (64 or 128)")
        if V % 8 or h % 8:
            raise RuntimeError("fsb200: vocab_size and hidden_size must be multiples of 8")
        t = self.tp
        if nh % t or self.ff % (8 * t) or V % (8 * t):
            raise RuntimeError(f"fsb200: heads {nh}, ff {self.ff} and vocab {V} must split evenly over tensor-parallel size {t}")
        # local (per tensor-parallel rank) extents
        self.nh_l, self.ff_l, self.V_l = nh // t, self.ff // t, V // t
        self.h_l = self.nh_l * self.hn
        self.tp_replicated_buckets = ("no_decay",)   # norms: identical on every tensor-parallel rank (counted once in the norm)

        spec = FlatSpec()
        spec.add("llama.embed_in.word_embeddings.weight", (self.V_l, h), "embed_in")
        for i in range(nl):
            p, bk = f"llama.layers.{i}.", f"layer{i}"
            spec.add(p + "input_layernorm.scale", (h,), bk)          # *.scale names match 'layernorm.' -> no-decay bucket
            spec.add(p + "attention.query_key_value.weight", (3 * self.h_l, h), bk)
            spec.add(p + "attention.dense.weight", (h, self.h_l), bk)
            spec.add(p + "post_attention_layernorm.scale", (h,), bk)
            spec.add(p + "mlp.w1.weight", (self.ff_l, h), bk)   # w1 | w3 adjacent: one [2ff, h] GEMM operand
            spec.add(p + "mlp.w3.weight", (self.ff_l, h), bk)
            spec.add(p + "mlp.w2.weight", (h, self.ff_l), bk)
        spec.add("llama.final_layer_norm.scale", (h,), "head")
        spec.add("embed_out.final_linear.weight", (self.V_l, h), "head")
        self.flat = FlatBuffers(spec, dev, world_size=world_size)

        # module tree mirroring the reference (modeling_llama.py:97-127, :239-252)
        def P(name):
            prm = nn.Parameter(self.flat.view(name), requires_grad=True)
            prm.main_grad = self.flat.view(name, grad=True)
            prm.fsb_name = name
            return prm

        self.llama = _Holder()
        self.llama.embed_in = _Holder()
        self.llama.embed_in.word_embeddings = _Holder()
        self.llama.embed_in.word_embeddings.weight = P("llama.embed_in.word_embeddings.weight")
        self.llama.layers = nn.ModuleList()
        inv_freq = 1.0 / (getattr(config, "rotary_emb_base", 10000) ** (torch.arange(0, self.hn, 2).float() / self.hn))
        for i in range(nl):
            p = f"llama.layers.{i}."
            lyr = _Holder()
            lyr.input_layernorm = _Holder(); lyr.input_layernorm.scale = P(p + "input_layernorm.scale")
            lyr.attention = _Holder()
            lyr.attention.query_key_value = _Holder()
            lyr.attention.query_key_value.weight = P(p + "attention.query_key_value.weight")
            lyr.attention.rotary_emb = _Holder()
            lyr.attention.rotary_emb.register_buffer("inv_freq", inv_freq.clone().to(dev))
            lyr.attention.dense = _Holder(); lyr.attention.dense.weight = P(p + "attention.dense.weight")
            lyr.post_attention_layernorm = _Holder()
            lyr.post_attention_layernorm.scale = P(p + "post_attention_layernorm.scale")
            lyr.mlp = _Holder()
            lyr.mlp.w1 = _Holder(); lyr.mlp.w1.weight = P(p + "mlp.w1.weight")
            lyr.mlp.w3 = _Holder(); lyr.mlp.w3.weight = P(p + "mlp.w3.weight")
            lyr.mlp.w2 = _Holder(); lyr.mlp.w2.weight = P(p + "mlp.w2.weight")
            lyr.w13 = None  # filled below (non-parameter view)
            self.llama.layers.append(lyr)
        self.llama.final_layer_norm = _Holder()
        self.llama.final_layer_norm.scale = P("llama.final_layer_norm.scale")
        self.embed_out = _Holder()
        self.embed_out.final_linear = _Holder()
        self.embed_out.final_linear.weight = P("embed_out.final_linear.weight")

        self._w13 = [self.flat.span(f"llama.layers.{i}.mlp.w1.weight", 2 * self.ff_l, h)
                     for i in range(nl)]
        self._dw13 = [self.flat.span(f"llama.layers.{i}.mlp.w1.weight", 2 * self.ff_l, h, grad=True) for i in range(nl)]

        # RoPE tables exactly as RotaryEmbedding builds them (layers/positional_embeddings.py:38-52), fp32
        self._inv_freq = inv_freq
        self._rope_rows = 0
        self._ensure_rope(getattr(config, "max_position_embeddings", 2048), dev)

        self.reset_parameters(seed)
        self.accumulate_grads = False   # set by the engine for micro-batches after the first
        self.loss_scale = 1.0           # 1 / (gradient_accumulation_steps * world_size), folded into dlogits
        self.grad_hook = None           # engine callback: grad_hook(bucket_name) when a bucket's gradients are final

    def _ensure_rope(self, n, dev=None):
        """cos/sin tables with at least n rows. RotaryEmbedding regrows its cache when a longer sequence arrives
        (positional_embeddings.py:54-68); fsb_rope_inplace never reads past the table (rows beyond it become NaN)."""
        if n <= self._rope_rows:
            return
        dev = dev if dev is not None else self.flat.params.device
        t = torch.arange(n, dtype=self._inv_freq.dtype)
        freqs = torch.einsum("i,j->ij", t, self._inv_freq)
        self._cos = freqs.cos().contiguous().to(dev)
        self._sin = freqs.sin().contiguous().to(dev)
        self._rope_rows = n

    # ---- init (layers/init_functions.py:121-142: small_init std sqrt(2/(5h)); wang_init std 2/(L*sqrt(h))) -------------
    @torch.no_grad()
    def reset_parameters(self, seed=0):
        h, nl = self.h, self.nl
        small = math.sqrt(2.0 / (5.0 * h))
        wang = 2.0 / (nl * math.sqrt(h))
        big = self.flat.total > 300_000_000  # billions of CPU randn take minutes: draw on the device instead
        gen = torch.Generator(device=self.flat.params.device if big else "cpu").manual_seed(seed + 7919 * self.tp_rank)
        for name, prm in self.named_parameters():
            if name.endswith(".scale"):
                prm.fill_(1.0)
                continue
            std = wang if (name.endswith("dense.weight") or name.endswith("w2.weight")) else small
            if big:
                prm.normal_(0.0, std, generator=gen)
            else:
                _init_normal_(prm.data, std, gen)

    # The reference scripts call `.from_pretrained(..., torch_dtype=torch.half).cuda()`; parameters here are views into the
    # flat bf16 CUDA buffer and must never be re-allocated by nn.Module._apply.
    def cuda(self, device=None):
        return self

    def half(self):
        return self

    def bfloat16(self):
        return self

    def to(self, *args, **kwargs):
        return self

    def train(self, mode=True):   # reference defect (modeling_llama.py:257-260: `mode` is required there); accept both
        return super().train(mode)

    # HF-style loading of a reference state dict (fp32/fp16/bf16 tensors on any device)
    @torch.no_grad()
    def load_reference_state_dict(self, sd):
        own = dict(self.named_parameters())
        missing = [k for k in own if k not in sd]
        if missing:
            raise KeyError(f"missing keys in state dict: {missing[:4]}...")
        for k, prm in own.items():
            if tuple(sd[k].shape) != tuple(prm.shape):
                raise ValueError(f"shape mismatch for {k}: {tuple(sd[k].shape)} vs {tuple(prm.shape)}")
            prm.copy_(sd[k].to(device=prm.device, dtype=prm.dtype))

    # ---- forward ----------------------------------------------------------------------------------------------------
    def forward(self, input_ids=None, attention_mask=None, position_ids=None, labels=None, return_logits=False, **_):
        B, S = input_ids.shape
        dev = self.flat.params.device
        ids = input_ids.to(device=dev, dtype=torch.int64).contiguous().view(-1)
        if position_ids is None:
            self._ensure_rope(S)
            pos = torch.arange(S, device=dev, dtype=torch.int64).repeat(B)
        else:
            if not position_ids.is_cuda:   # host tensor (the collators emit CPU batches): exact bound, no device sync
                lo, hi = int(position_ids.min()), int(position_ids.max())
                if lo < 0:
                    raise ValueError(f"position_ids must be non-negative, got {lo}")
                self._ensure_rope(hi + 1)
            pos = position_ids.to(device=dev, dtype=torch.int64).expand(B, S).contiguous().view(-1)
            if position_ids.is_cuda:       # device tensor: asynchronous device-side check against the table size
                self._ensure_rope(S)
                torch._assert_async(((pos >= 0) & (pos < self._rope_rows)).all(),
                                    "fsb200 LlamaForCausalLM: position_ids outside [0, rope table rows); pass them on the "
                                    "host or raise config.max_position_embeddings")
        lab = None if labels is None else labels.to(device=dev, dtype=torch.int64).contiguous().view(-1)
        if lab is not None and torch.is_grad_enabled():
            # a leaf that requires grad makes the node differentiable; real gradients go to the flat grad buffer
            loss, logits = _LlamaStep.apply(self, ids, pos, lab, B, S, return_logits,
                                            self.llama.final_layer_norm.scale)
        else:
            loss, logits, _ = self._forward_impl(ids, pos, lab, B, S, save=False, want_logits=True)
        return SimpleNamespace(loss=loss, logits=None if logits is None else logits.view(B, S, self.V),
                               past_key_values=None, hidden_states=None, attentions=None)

    def _forward_impl(self, ids, pos, lab, B, S, save, want_logits):
        h, nh, hn, ff = self.h, self.nh_l, self.hn, self.ff_l          # LOCAL heads / ff columns under tensor parallelism
        hl = self.h_l
        T = B * S
        acts = []
        self._need("no_decay"); self._need("embed_in")
        ids_l, emb_keep = self._local_ids(ids)
        x = ops.embedding_fwd(ids_l, self.llama.embed_in.word_embeddings.weight.data)
        if emb_keep is not None:      # VocabParallelEmbedding.forward (mpu/layers.py:104-130): foreign rows are zero, then all-reduce
            x.mul_(emb_keep)
            self._tp_all_reduce(x)
        prev_m = None
        for i, lyr in enumerate(self.llama.layers):
            self._need(f"layer{i}")
            h1, rstd1, x = ops.rmsnorm_fwd(x if prev_m is None else prev_m, lyr.input_layernorm.scale.data, self.eps,
                                           residual=None if prev_m is None else x)
            qkv = ops.gemm(L.GEMM_NT, h1, lyr.attention.query_key_value.weight.data)
            ops.rope_inplace(qkv, self._cos, self._sin, pos, nh, hn, 3 * hl, 3 * hn, offset=0)
            ops.rope_inplace(qkv, self._cos, self._sin, pos, nh, hn, 3 * hl, 3 * hn, offset=hn)
            q5 = qkv.view(B, S, nh, 3, hn)
            o, lse = ops.sdpa_fwd(q5[:, :, :, 0], q5[:, :, :, 1], q5[:, :, :, 2], 1.0 / math.sqrt(hn), True)
            o2 = o.view(T, hl)
            a = ops.gemm(L.GEMM_NT, o2, lyr.attention.dense.weight.data)
            self._tp_all_reduce(a)        # RowParallelLinear: partial sums over the head shards (mpu/layers.py:451-470)
            h2, rstd2, x1 = ops.rmsnorm_fwd(a, lyr.post_attention_layernorm.scale.data, self.eps, residual=x)
            gu = ops.gemm(L.GEMM_NT, h2, self._w13[i])
            act = ops.glu_fwd(L.ACT_SILU, gu[:, :ff], gu[:, ff:])
            m = ops.gemm(L.GEMM_NT, act, lyr.mlp.w2.weight.data)
            self._tp_all_reduce(m)        # RowParallelLinear (w2)
            if save:
                acts.append((x, rstd1, h1, qkv, o, lse, x1, rstd2, h2, gu, act))
            x, prev_m = x1, m
        self._need("head")
        hf, rstdf, xf = ops.rmsnorm_fwd(prev_m, self.llama.final_layer_norm.scale.data, self.eps, residual=x)
        logits = ops.gemm(L.GEMM_NT, hf, self.embed_out.final_linear.weight.data)
        logits = self._tp_gather_columns(logits)   # ParallelLinear(parallel_output=False): full-vocabulary logits on every rank
        loss = None
        ctx = None
        if lab is not None:
            keep = logits.clone() if (want_logits and save) else None
            loss, dlogits, _ = ops.softmax_xent(logits, lab, S, shift=1, grad_scale=self.loss_scale,
                                                dlogits="inplace" if save else None)
            if save:
                ctx = (acts, hf, rstdf, xf, dlogits, ids, pos, B, S)
                logits = keep
        return loss, (logits if want_logits else None), ctx

    # ---- KV-cache inference -----------------------------------------------------------------------
    # The reference decodes through HF's GenerationMixin: `prepare_inputs_for_generation` (modeling_llama.py:353-377) feeds the
    # last token with position_ids = cumsum(attention_mask) - 1, every layer concatenates the new key / value onto `layer_past`
    # (layers/transformer.py:529-537) and `llama_generate.generate` (examples/ziya_llama/llama_generate.py:16-39) left-pads the
    # prompts. Here the cache is a pre-allocated [batch, max_length, heads, head_dim] pair per layer; a decode step runs the
    # same kernels as training with one query row per sequence, the unused tail of the cache (and the left padding) hidden by
    # the key mask. Padded keys ARE masked (the reference's `global` attention path; its flash path ignores the mask).
    @property
    def device(self):
        return self.flat.params.device

    def _layer_infer(self, i, x, prev_m, pos, B, S, kc, vc, at, kv_mask, causal):
        """One layer forward without saving activations; writes this call's keys / values into the cache at [at, at + S)."""
        if self.tp > 1:
            raise NotImplementedError("fsb200: KV-cache decoding under tensor parallelism is not implemented")
        h, nh, hn, ff = self.h, self.nh, self.hn, self.ff
        lyr = self.llama.layers[i]
        h1, _, x = ops.rmsnorm_fwd(x if prev_m is None else prev_m, lyr.input_layernorm.scale.data, self.eps,
                                   residual=None if prev_m is None else x)
        qkv = ops.gemm(L.GEMM_NT, h1, lyr.attention.query_key_value.weight.data)
        ops.rope_inplace(qkv, self._cos, self._sin, pos, nh, hn, 3 * h, 3 * hn, offset=0)
        ops.rope_inplace(qkv, self._cos, self._sin, pos, nh, hn, 3 * h, 3 * hn, offset=hn)
        q5 = qkv.view(B, S, nh, 3, hn)
        kc[:, at:at + S].copy_(q5[:, :, :, 1])
        vc[:, at:at + S].copy_(q5[:, :, :, 2])
        if causal:   # prefill: attend inside the prompt (causal + left-padding mask)
            o, _ = ops.sdpa_fwd(q5[:, :, :, 0], q5[:, :, :, 1], q5[:, :, :, 2], 1.0 / math.sqrt(hn), True,
                                kv_mask=None if kv_mask is None else kv_mask[:, :S].contiguous())
        else:        # decode: one query row per sequence against the whole cache; unwritten slots are masked
            o, _ = ops.sdpa_fwd(q5[:, :, :, 0], kc, vc, 1.0 / math.sqrt(hn), False, kv_mask=kv_mask)
        a = ops.gemm(L.GEMM_NT, o.view(B * S, h), lyr.attention.dense.weight.data)
        h2, _, x1 = ops.rmsnorm_fwd(a, lyr.post_attention_layernorm.scale.data, self.eps, residual=x)
        gu = ops.gemm(L.GEMM_NT, h2, self._w13[i])
        act = ops.glu_fwd(L.ACT_SILU, gu[:, :ff], gu[:, ff:])
        m = ops.gemm(L.GEMM_NT, act, lyr.mlp.w2.weight.data)
        return x1, m

    def _infer(self, ids, pos, B, S, cache, at, kv_mask, causal):
        """ids / pos flattened [B*S]; returns fp32 logits of the LAST position of every sequence, [B, V]."""
        for b in ("no_decay", "embed_in"):
            self._need(b)
        x, prev_m = ops.embedding_fwd(ids, self.llama.embed_in.word_embeddings.weight.data), None
        for i in range(self.nl):
            self._need(f"layer{i}")
            x, prev_m = self._layer_infer(i, x, prev_m, pos, B, S, cache[i][0], cache[i][1], at, kv_mask, causal)
        self._need("head")
        hf, _, _ = ops.rmsnorm_fwd(prev_m, self.llama.final_layer_norm.scale.data, self.eps, residual=x)
        last = hf.view(B, S, self.h)[:, -1].contiguous()                       # [B, h]
        rows = max(8, B)                                                        # the GEMM wants >= 8 aligned rows
        if rows != B:
            pad = torch.zeros((rows, self.h), dtype=last.dtype, device=last.device)
            pad[:B] = last
            last = pad
        return ops.gemm(L.GEMM_NT, last, self.embed_out.final_linear.weight.data)[:B].float()

    @torch.no_grad()
    def generate(self, input_ids, attention_mask=None, max_length=None, max_new_tokens=None, do_sample=False,
                 temperature=1.0, top_k=0, top_p=1.0, repetition_penalty=1.0, pad_token_id=None, eos_token_id=None,
                 generator=None, **_):
        """Greedy / sampling decode with a KV cache; the keyword surface `llama_generate.generate` passes to HF's
        `model.generate` (do_sample, top_p, top_k, max_length, repetition_penalty, temperature, pad_token_id, eos_token_id).
        Prompts are LEFT-padded (attention_mask 0 on the pads). Returns [batch, <= max_length] token ids, prompt included,
        finished rows filled with pad_token_id — HF's GenerationMixin conventions."""
        dev = self.device
        ids = input_ids.to(device=dev, dtype=torch.int64)
        B, S0 = ids.shape
        mask = torch.ones((B, S0), dtype=torch.uint8, device=dev) if attention_mask is None else \
            attention_mask.to(device=dev).to(torch.uint8)
        if max_length is None:
            max_length = S0 + (max_new_tokens if max_new_tokens is not None else 20)
        if max_length <= S0:
            return ids
        pad_id = pad_token_id if pad_token_id is not None else (eos_token_id if eos_token_id is not None else 0)
        Lmax = (max_length + 63) // 64 * 64
        self._ensure_rope(max_length)
        cache = [(torch.zeros((B, Lmax, self.nh, self.hn), dtype=torch.bfloat16, device=dev),
                  torch.zeros((B, Lmax, self.nh, self.hn), dtype=torch.bfloat16, device=dev)) for _ in range(self.nl)]
        kv_mask = torch.zeros((B, Lmax), dtype=torch.uint8, device=dev)
        kv_mask[:, :S0] = mask
        # position_ids = cumsum(mask) - 1, pads -> 1 (prepare_inputs_for_generation, modeling_llama.py:360-366)
        pos = mask.long().cumsum(-1) - 1
        pos = pos.masked_fill(mask == 0, 1)
        logits = self._infer(ids.reshape(-1), pos.reshape(-1).contiguous(), B, S0, cache, 0, kv_mask if not bool(mask.all()) else None,
                             True)
        count = mask.long().sum(-1)                                             # real tokens so far = next position id
        seqs = ids
        done = torch.zeros(B, dtype=torch.bool, device=dev)
        for cur in range(S0, max_length):
            nxt = self._pick(logits, seqs, do_sample, temperature, top_k, top_p, repetition_penalty, generator)
            if eos_token_id is not None:
                nxt = torch.where(done, torch.full_like(nxt, pad_id), nxt)
                done = done | (nxt == eos_token_id)
            seqs = torch.cat([seqs, nxt[:, None]], dim=1)
            if cur + 1 >= max_length or bool(done.all()):
                break
            kv_mask[:, cur] = 1
            logits = self._infer(nxt, count.clone(), B, 1, cache, cur, kv_mask, False)
            count = count + 1
        return seqs

    @staticmethod
    def _pick(logits, seqs, do_sample, temperature, top_k, top_p, repetition_penalty, generator):
        """HF logits-processor order: repetition penalty -> temperature -> top-k -> top-p -> sample (or argmax)."""
        if repetition_penalty != 1.0:
            seen = torch.gather(logits, 1, seqs)
            seen = torch.where(seen < 0, seen * repetition_penalty, seen / repetition_penalty)
            logits = logits.scatter(1, seqs, seen)
        if not do_sample:
            return logits.argmax(-1)
        if temperature != 1.0:
            logits = logits / temperature
        if top_k and top_k > 0:
            kth = torch.topk(logits, min(top_k, logits.shape[-1]), dim=-1).values[:, -1:]
            logits = logits.masked_fill(logits < kth, float("-inf"))
        if top_p < 1.0:
            srt, idx = torch.sort(logits, descending=False, dim=-1)
            cum = torch.softmax(srt, -1).cumsum(-1)
            remove = cum <= (1.0 - top_p)
            remove[:, -1] = False
            logits = logits.masked_fill(remove.scatter(1, idx, remove), float("-inf"))
        return torch.multinomial(torch.softmax(logits, -1), 1, generator=generator).squeeze(1)

    # ---- backward ---------------------------------------------------------------------------------------------------
    def _backward_impl(self, ctx, gloss):
        acts, hf, rstdf, xf, dlogits, ids, pos, B, S = ctx
        h, nh, hn, ff = self.h, self.nh_l, self.hn, self.ff_l
        hl = self.h_l
        T = B * S
        acc = self.accumulate_grads
        self._begin_backward()
        if gloss is not None:
            ops.scale_inplace(dlogits, gloss)  # upstream scalar; the kernel exits immediately when it is 1.0
        if self.tp > 1:   # this rank's vocabulary columns of dlogits (a strided view: the GEMMs take the row stride)
            dlogits = dlogits[:, self.tp_rank * self.V_l:(self.tp_rank + 1) * self.V_l]
        W_out = self.embed_out.final_linear.weight
        dhf = ops.gemm(L.GEMM_NN, dlogits, W_out.data)
        self._tp_all_reduce(dhf)          # backward of copy_to_model_parallel_region (mpu/mappings.py): sum the partial dgrads
        ops.gemm(L.GEMM_TN, dlogits, hf, out=W_out.main_grad, accumulate=acc)
        del dlogits
        self._done("head")
        fscale = self.llama.final_layer_norm.scale
        dx = ops.rmsnorm_bwd(dhf, xf, fscale.data, rstdf, fscale.main_grad, accumulate=acc)
        for i in reversed(range(self.nl)):
            lyr = self.llama.layers[i]
            x, rstd1, h1, qkv, o, lse, x1, rstd2, h2, gu, act = acts[i]
            acts[i] = None
            # x_next = x1 + m  ->  dm = dx, residual gradient into x1 = dx
            w2 = lyr.mlp.w2.weight
            dact = ops.gemm(L.GEMM_NN, dx, w2.data)
            ops.gemm(L.GEMM_TN, dx, act, out=w2.main_grad, accumulate=acc)
            dgu = torch.empty_like(gu)
            ops.glu_bwd(L.ACT_SILU, dact, gu[:, :ff], gu[:, ff:], dgu[:, :ff], dgu[:, ff:])
            dh2 = ops.gemm(L.GEMM_NN, dgu, self._w13[i])
            self._tp_all_reduce(dh2)      # column-parallel w1|w3: dgrad partial sums
            ops.gemm(L.GEMM_TN, dgu, h2, out=self._dw13[i], accumulate=acc)
            s2 = lyr.post_attention_layernorm.scale
            dx1 = ops.rmsnorm_bwd(dh2, x1, s2.data, rstd2, s2.main_grad, accumulate=acc, dres=dx)
            # x1 = x + a  ->  da = dx1
            wd = lyr.attention.dense.weight
            do = ops.gemm(L.GEMM_NN, dx1, wd.data)
            ops.gemm(L.GEMM_TN, dx1, o.view(T, hl), out=wd.main_grad, accumulate=acc)
            dqkv = torch.empty_like(qkv)
            q5, d5 = qkv.view(B, S, nh, 3, hn), dqkv.view(B, S, nh, 3, hn)
            ops.sdpa_bwd(q5[:, :, :, 0], q5[:, :, :, 1], q5[:, :, :, 2], o, do.view(B, S, nh, hn), lse,
                         1.0 / math.sqrt(hn), True, d5[:, :, :, 0], d5[:, :, :, 1], d5[:, :, :, 2])
            ops.rope_inplace(dqkv, self._cos, self._sin, pos, nh, hn, 3 * hl, 3 * hn, backward=True, offset=0)
            ops.rope_inplace(dqkv, self._cos, self._sin, pos, nh, hn, 3 * hl, 3 * hn, backward=True, offset=hn)
            wq = lyr.attention.query_key_value.weight
            dh1 = ops.gemm(L.GEMM_NN, dqkv, wq.data)
            self._tp_all_reduce(dh1)      # column-parallel QKV: dgrad partial sums
            ops.gemm(L.GEMM_TN, dqkv, h1, out=wq.main_grad, accumulate=acc)
            s1 = lyr.input_layernorm.scale
            dx = ops.rmsnorm_bwd(dh1, x, s1.data, rstd1, s1.main_grad, accumulate=acc, dres=dx1)
            self._done(f"layer{i}")
        W_in = self.llama.embed_in.word_embeddings.weight
        if not acc:
            W_in.main_grad.zero_()
        ids_l, emb_keep = self._local_ids(ids)
        if emb_keep is not None:
            dx = dx * emb_keep            # rows of foreign vocabulary shards contribute nothing here
        ops.embedding_bwd(ids_l, dx, W_in.main_grad)
        self._done("embed_in")
        self._done("no_decay")

    # ---- tensor-parallel exchanges (fengshen/models/megatron/mpu/mappings.py:29-192), NCCL on the compute stream -------------
    def _tp_all_reduce(self, t):
        if self.tp > 1:
            import torch.distributed as dist
            dist.all_reduce(t, group=self.tp_group)

    def _tp_gather_columns(self, t):
        """[rows, n/tp] on every rank -> [rows, n] (gather_from_model_parallel_region, last-dimension concatenation)."""
        if self.tp == 1:
            return t
        import torch.distributed as dist
        buf = torch.empty((self.tp,) + tuple(t.shape), dtype=t.dtype, device=t.device)
        dist.all_gather_into_tensor(buf, t.contiguous(), group=self.tp_group)
        return buf.permute(1, 0, 2).reshape(t.shape[0], self.tp * t.shape[1])

    def _local_ids(self, ids):
        """VocabParallelEmbedding index arithmetic (mpu/layers.py:104-121): ids of this rank's vocabulary range re-based to 0,
        foreign ids clamped (their rows are zeroed by the returned [rows, 1] bf16 mask). (ids, None) without tensor parallelism."""
        if self.tp == 1:
            return ids, None
        lo = self.tp_rank * self.V_l
        mine = (ids >= lo) & (ids < lo + self.V_l)
        return torch.where(mine, ids - lo, torch.zeros_like(ids)), mine.to(torch.bfloat16)[:, None]

    def _done(self, bucket):
        if self.grad_hook is not None:
            self.grad_hook(bucket)

    def _need(self, bucket):
        """Forward is about to read this bucket's parameters (the engine may still be all-gathering them)."""
        hook = getattr(self, "param_hook", None)
        if hook is not None:
            hook(bucket)

    def _begin_backward(self):
        hook = getattr(self, "backward_begin_hook", None)
        if hook is not None:
            hook()


class _LlamaStep(torch.autograd.Function):
    """The whole network as one autograd node; `flat_params` is the differentiable handle (its .grad is never
    materialised: gradients are written into model.flat.grads by the kernels)."""

    @staticmethod
    def forward(ctx, model, ids, pos, lab, B, S, want_logits, flat_params):
        loss, logits, saved = model._forward_impl(ids, pos, lab, B, S, save=True, want_logits=want_logits)
        ctx.model = model
        ctx.saved = saved
        ctx.mark_non_differentiable(*([logits] if logits is not None else []))
        return loss, logits

    @staticmethod
    def backward(ctx, gloss, _glogits):
        model, saved = ctx.model, ctx.saved
        ctx.saved = None
        model._backward_impl(saved, None if gloss is None else gloss)
        return (None,) * 8
