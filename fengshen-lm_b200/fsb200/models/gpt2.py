"""Wenzhong-GPT2 on the fsb200 kernels — drop-in for `transformers.GPT2LMHeadModel` as the reference uses it
(fengshen/examples/wenzhong_qa/finetune_wenzhong.py:56: `GPT2LMHeadModel.from_pretrained(...)`, training_step :89-113).

The arithmetic restated here lives in 3P `transformers` (gpt2/modeling_gpt2.py):
pre-LN blocks `x + attn(ln_1(x))`, `x + mlp(ln_2(x))`, final `ln_f`; fused `c_attn` Conv1D with weight stored [in, out]
(q | k | v contiguous thirds); `gelu_new`; learned positions `wte[ids] + wpe[pos]`; LM head tied to `wte`; shifted mean
cross-entropy with ignore_index -100. State-dict keys follow HF (`transformer.h.N.attn.c_attn.weight`, ...).
Conv1D's [in, out] layout maps onto the GEMM layouts without any transpose: forward = NN, dgrad = NT, wgrad = TN.
Dropout probabilities must be 0 (parity / benchmark setting); a non-zero value is rejected loudly.
"""
import math
from types import SimpleNamespace

import torch
from torch import nn

from .. import lib as L
from .. import ops
from ..flat import FlatBuffers, FlatSpec


class _Holder(nn.Module):
    pass


class GPT2LMHeadModel(nn.Module):
    def __init__(self, config, device=None, world_size=None, seed=0):
        super().__init__()
        self.config = config
        if world_size is None:
            import torch.distributed as dist
            world_size = dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1
        dev = torch.device(device if device is not None else f"cuda:{torch.cuda.current_device()}"
                           if torch.cuda.is_available() else "cuda")
        if dev.type != "cuda":
            raise RuntimeError("fsb200 GPT2LMHeadModel runs on CUDA only (no CPU fallback on the product path)")
        g = lambda k, d=None: getattr(config, k, d)
        self.h, self.nl, self.nh = g("n_embd", g("hidden_size")), g("n_layer", g("num_hidden_layers")), \
            g("n_head", g("num_attention_heads"))
        self.V, self.npos = g("vocab_size"), g("n_positions", g("max_position_embeddings", 1024))
        self.eps = g("layer_norm_epsilon", 1e-5)
        self.inner = g("n_inner") or 4 * self.h
        for k in ("resid_pdrop", "embd_pdrop", "attn_pdrop"):
            if g(k, 0.0) not in (0, 0.0):
                raise RuntimeError(f"fsb200 GPT2: {k}={g(k)} — dropout is not implemented; set it to 0")
        if g("activation_function", "gelu_new") != "gelu_new":
            raise RuntimeError("fsb200 GPT2: only activation_function='gelu_new' is implemented")
        h, V = self.h, self.V
        self.hn = h // self.nh
        if self.hn not in (64, 128) or V % 8 or h % 8:
            raise RuntimeError("fsb200 GPT2: head dim must be 64/128 and vocab/hidden multiples of 8 (pad the vocab)")

        spec = FlatSpec()
        spec.add("transformer.wte.weight", (V, h), "wte")
        spec.add("transformer.wpe.weight", (self.npos, h), "wte")
        for i in range(self.nl):
            p, bk = f"transformer.h.{i}.", f"layer{i}"
            for n, s in (("ln_1.weight", (h,)), ("ln_1.bias", (h,)), ("attn.c_attn.weight", (h, 3 * h)),
                         ("attn.c_attn.bias", (3 * h,)), ("attn.c_proj.weight", (h, h)), ("attn.c_proj.bias", (h,)),
                         ("ln_2.weight", (h,)), ("ln_2.bias", (h,)), ("mlp.c_fc.weight", (h, self.inner)),
                         ("mlp.c_fc.bias", (self.inner,)), ("mlp.c_proj.weight", (self.inner, h)),
                         ("mlp.c_proj.bias", (h,))):
                spec.add(p + n, s, bk)
        spec.add("transformer.ln_f.weight", (h,), "wte")
        spec.add("transformer.ln_f.bias", (h,), "wte")
        self.flat = FlatBuffers(spec, dev, world_size=world_size)

        def P(name):
            prm = nn.Parameter(self.flat.view(name), requires_grad=True)
            prm.main_grad = self.flat.view(name, grad=True)
            return prm

        tr = self.transformer = _Holder()
        tr.wte = _Holder(); tr.wte.weight = P("transformer.wte.weight")
        tr.wpe = _Holder(); tr.wpe.weight = P("transformer.wpe.weight")
        tr.h = nn.ModuleList()
        for i in range(self.nl):
            p = f"transformer.h.{i}."
            blk = _Holder()
            blk.ln_1 = _Holder(); blk.ln_1.weight = P(p + "ln_1.weight"); blk.ln_1.bias = P(p + "ln_1.bias")
            blk.attn = _Holder()
            blk.attn.c_attn = _Holder()
            blk.attn.c_attn.weight = P(p + "attn.c_attn.weight"); blk.attn.c_attn.bias = P(p + "attn.c_attn.bias")
            blk.attn.c_proj = _Holder()
            blk.attn.c_proj.weight = P(p + "attn.c_proj.weight"); blk.attn.c_proj.bias = P(p + "attn.c_proj.bias")
            blk.ln_2 = _Holder(); blk.ln_2.weight = P(p + "ln_2.weight"); blk.ln_2.bias = P(p + "ln_2.bias")
            blk.mlp = _Holder()
            blk.mlp.c_fc = _Holder()
            blk.mlp.c_fc.weight = P(p + "mlp.c_fc.weight"); blk.mlp.c_fc.bias = P(p + "mlp.c_fc.bias")
            blk.mlp.c_proj = _Holder()
            blk.mlp.c_proj.weight = P(p + "mlp.c_proj.weight"); blk.mlp.c_proj.bias = P(p + "mlp.c_proj.bias")
            tr.h.append(blk)
        tr.ln_f = _Holder(); tr.ln_f.weight = P("transformer.ln_f.weight"); tr.ln_f.bias = P("transformer.ln_f.bias")
        self.lm_head = _Holder()
        self.lm_head.weight = tr.wte.weight  # tied (modeling_gpt2.py:646)

        self.reset_parameters(seed)
        self.accumulate_grads = False
        self.loss_scale = 1.0
        self.grad_hook = None

    @torch.no_grad()
    def reset_parameters(self, seed=0):
        """HF GPT2 _init_weights: N(0, initializer_range) for Linear/Embedding, c_proj scaled by 1/sqrt(2*n_layer),
        LayerNorm weight 1 / bias 0, biases 0."""
        std = getattr(self.config, "initializer_range", 0.02)
        gen = torch.Generator(device=self.flat.params.device).manual_seed(seed)
        for name, prm in self.named_parameters():
            if name.endswith("bias"):
                prm.zero_()
            elif ".ln_" in name:
                prm.fill_(1.0)
            else:
                s = std / math.sqrt(2 * self.nl) if name.endswith("c_proj.weight") else std
                prm.normal_(0.0, s, generator=gen)

    @torch.no_grad()
    def load_reference_state_dict(self, sd):
        for k, prm in self.named_parameters():
            if tuple(sd[k].shape) != tuple(prm.shape):
                raise ValueError(f"shape mismatch for {k}: {tuple(sd[k].shape)} vs {tuple(prm.shape)}")
            prm.copy_(sd[k].to(device=prm.device, dtype=prm.dtype))

    # ---- forward ----------------------------------------------------------------------------------------------------
    def forward(self, input_ids=None, attention_mask=None, labels=None, position_ids=None, return_logits=False, **_):
        B, S = input_ids.shape
        dev = self.flat.params.device
        ids = input_ids.to(device=dev, dtype=torch.int64).contiguous().view(-1)
        pos = None if position_ids is None else \
            position_ids.to(device=dev, dtype=torch.int64).expand(B, S).contiguous().view(-1)
        mask = None
        if attention_mask is not None and not bool(attention_mask.all()):
            mask = attention_mask.to(device=dev, dtype=torch.uint8).contiguous()
        lab = None if labels is None else labels.to(device=dev, dtype=torch.int64).contiguous().view(-1)
        if lab is not None and torch.is_grad_enabled():
            loss, logits = _GPT2Step.apply(self, ids, pos, mask, lab, B, S, return_logits, self.transformer.ln_f.weight)
        else:
            loss, logits, _ = self._forward_impl(ids, pos, mask, lab, B, S, save=False, want_logits=True)
        return SimpleNamespace(loss=loss, logits=None if logits is None else logits.view(B, S, self.V),
                               past_key_values=None, hidden_states=None, attentions=None)

    def _forward_impl(self, ids, pos, mask, lab, B, S, save, want_logits):
        h, nh, hn = self.h, self.nh, self.hn
        T = B * S
        tr = self.transformer
        self._need("no_decay"); self._need("wte")
        x = ops.embedding_fwd(ids, tr.wte.weight.data, pos=pos, P=tr.wpe.weight.data, seq_len=S)
        acts, prev_m = [], None
        scale = 1.0 / math.sqrt(hn)
        for i, blk in enumerate(tr.h):
            self._need(f"layer{i}")
            h1, st1, x = ops.layernorm_fwd(x if prev_m is None else prev_m, blk.ln_1.weight.data, blk.ln_1.bias.data,
                                           self.eps, residual=None if prev_m is None else x)
            qkv = ops.gemm(L.GEMM_NN, h1, blk.attn.c_attn.weight.data, bias=blk.attn.c_attn.bias.data)
            q5 = qkv.view(B, S, 3, nh, hn)
            o, lse = ops.sdpa_fwd(q5[:, :, 0], q5[:, :, 1], q5[:, :, 2], scale, True, kv_mask=mask)
            a = ops.gemm(L.GEMM_NN, o.view(T, h), blk.attn.c_proj.weight.data, bias=blk.attn.c_proj.bias.data)
            h2, st2, x1 = ops.layernorm_fwd(a, blk.ln_2.weight.data, blk.ln_2.bias.data, self.eps, residual=x)
            pre = torch.empty((T, self.inner), dtype=torch.bfloat16, device=x.device) if save else None
            f = ops.gemm(L.GEMM_NN, h2, blk.mlp.c_fc.weight.data, bias=blk.mlp.c_fc.bias.data,
                         epilogue=L.EPI_GELU_TANH, aux=pre)
            m = ops.gemm(L.GEMM_NN, f, blk.mlp.c_proj.weight.data, bias=blk.mlp.c_proj.bias.data)
            if save:
                acts.append((x, st1, h1, qkv, o, lse, x1, st2, h2, pre, f))
            x, prev_m = x1, m
        hf, stf, xf = ops.layernorm_fwd(prev_m, tr.ln_f.weight.data, tr.ln_f.bias.data, self.eps, residual=x)
        logits = ops.gemm(L.GEMM_NT, hf, tr.wte.weight.data)
        loss, ctx = None, None
        if lab is not None:
            keep = logits.clone() if (want_logits and save) else None
            loss, dlogits, _ = ops.softmax_xent(logits, lab, S, shift=1, grad_scale=self.loss_scale,
                                                dlogits="inplace" if save else None)
            if save:
                ctx = (acts, hf, stf, xf, dlogits, ids, pos, mask, B, S)
                logits = keep
        return loss, (logits if want_logits else None), ctx

    # ---- backward ---------------------------------------------------------------------------------------------------
    def _backward_impl(self, ctx, gloss):
        acts, hf, stf, xf, dlogits, ids, pos, mask, B, S = ctx
        h, nh, hn = self.h, self.nh, self.hn
        T = B * S
        acc = self.accumulate_grads
        self._begin_backward()
        tr = self.transformer
        scale = 1.0 / math.sqrt(hn)
        if gloss is not None:
            ops.scale_inplace(dlogits, gloss)  # upstream scalar; the kernel exits immediately when it is 1.0
        wte = tr.wte.weight
        dhf = ops.gemm(L.GEMM_NN, dlogits, wte.data)
        ops.gemm(L.GEMM_TN, dlogits, hf, out=wte.main_grad, accumulate=acc)   # tied head: written first, embedding adds later
        del dlogits
        dx = ops.layernorm_bwd(dhf, xf, tr.ln_f.weight.data, stf, tr.ln_f.weight.main_grad, tr.ln_f.bias.main_grad,
                               accumulate=acc)
        for i in reversed(range(self.nl)):
            blk = tr.h[i]
            x, st1, h1, qkv, o, lse, x1, st2, h2, pre, f = acts[i]
            acts[i] = None
            w = blk.mlp.c_proj
            df = ops.gemm(L.GEMM_NT, dx, w.weight.data)
            ops.gemm(L.GEMM_TN, f, dx, out=w.weight.main_grad, accumulate=acc)
            ops.colsum(dx, w.bias.main_grad, accumulate=acc)
            w = blk.mlp.c_fc
            dpre = ops.act_bwd_bias(L.ACT_GELU_TANH, df, pre, w.bias.main_grad, accumulate=acc)   # dGELU + c_fc bias grad
            dh2 = ops.gemm(L.GEMM_NT, dpre, w.weight.data)
            ops.gemm(L.GEMM_TN, h2, dpre, out=w.weight.main_grad, accumulate=acc)
            dx1 = ops.layernorm_bwd(dh2, x1, blk.ln_2.weight.data, st2, blk.ln_2.weight.main_grad,
                                    blk.ln_2.bias.main_grad, accumulate=acc, dres=dx)
            w = blk.attn.c_proj
            do = ops.gemm(L.GEMM_NT, dx1, w.weight.data)
            ops.gemm(L.GEMM_TN, o.view(T, h), dx1, out=w.weight.main_grad, accumulate=acc)
            ops.colsum(dx1, w.bias.main_grad, accumulate=acc)
            dqkv = torch.empty_like(qkv)
            q5, d5 = qkv.view(B, S, 3, nh, hn), dqkv.view(B, S, 3, nh, hn)
            ops.sdpa_bwd(q5[:, :, 0], q5[:, :, 1], q5[:, :, 2], o, do.view(B, S, nh, hn), lse, scale, True,
                         d5[:, :, 0], d5[:, :, 1], d5[:, :, 2], kv_mask=mask)
            w = blk.attn.c_attn
            dh1 = ops.gemm(L.GEMM_NT, dqkv, w.weight.data)
            ops.gemm(L.GEMM_TN, h1, dqkv, out=w.weight.main_grad, accumulate=acc)
            ops.colsum(dqkv, w.bias.main_grad, accumulate=acc)
            dx = ops.layernorm_bwd(dh1, x, blk.ln_1.weight.data, st1, blk.ln_1.weight.main_grad,
                                   blk.ln_1.bias.main_grad, accumulate=acc, dres=dx1)
            self._done(f"layer{i}")
        ops.embedding_bwd(ids, dx, wte.main_grad)  # accumulates onto the LM-head wgrad (tied weights)
        wpe = tr.wpe.weight
        if pos is None:
            # dP[s] = sum_b dx[b, s]: column sum of dx viewed as [B, S*h]
            ops.colsum(dx.view(B, S * h), wpe.main_grad[:S].reshape(-1), accumulate=acc)
            if not acc and S < self.npos:
                wpe.main_grad[S:].zero_()
        else:
            if not acc:
                wpe.main_grad.zero_()
            ops.embedding_bwd(pos, dx, wpe.main_grad)
        self._done("wte")
        self._done("no_decay")

    def save_pretrained(self, path, **_):
        """HF-style export (config.json + pytorch_model.bin with this class's HF key names): fsb200/models/export.py."""
        from .export import save_pretrained
        save_pretrained(self, path)

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


class _GPT2Step(torch.autograd.Function):
    @staticmethod
    def forward(ctx, model, ids, pos, mask, lab, B, S, want_logits, _anchor):
        loss, logits, saved = model._forward_impl(ids, pos, mask, lab, B, S, save=True, want_logits=want_logits)
        ctx.model, ctx.saved = model, saved
        ctx.mark_non_differentiable(*([logits] if logits is not None else []))
        return loss, logits

    @staticmethod
    def backward(ctx, gloss, _glogits):
        model, saved = ctx.model, ctx.saved
        ctx.saved = None
        model._backward_impl(saved, gloss)
        return (None,) * 9
