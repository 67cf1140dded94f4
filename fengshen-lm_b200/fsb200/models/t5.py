"""Randeng-T5 / mT5 on the fsb200 kernels — drop-in for `transformers.MT5ForConditionalGeneration` as the reference uses it
(fengshen/examples/pretrain_t5/pretrain_t5.py:57-59 builds it from an MT5Config; training_step :81-87 is
`self.model(input_ids=..., labels=...)`). BASELINE config 5.

The arithmetic restated here lives in 3P `transformers` (mt5/modeling_mt5.py):
  * pre-RMSNorm sub-layers `x + f(norm(x))` (MT5LayerNorm :47-70 == the in-tree RMSNorm formula), final RMSNorm per stack;
  * attention WITHOUT the 1/sqrt(d) scale (:300) plus an additive relative-position bias `table[bucket(k - q), h]` (:181-235)
    owned by the FIRST self-attention layer of each stack and reused by every layer of that stack; fp32 softmax (:323);
    decoder cross-attention over the encoder's final hidden states with zero bias and the encoder padding mask;
  * gated-GeLU FFN `wo(gelu_new(wi_0 x) * wi_1 x)` (:96-123); nothing has a bias vector;
  * one shared token embedding for both stacks, `decoder_input_ids = shift_right(labels)` (start id = pad id, -100 -> pad;
    :592, :1123-1125), untied `lm_head` without the d^-0.5 rescale (:1143), CrossEntropyLoss(ignore_index=-100) over every
    decoder position (:1147-1150).
State-dict keys follow HF (`shared.weight`, `encoder.block.N.layer.0.SelfAttention.q.weight`, ...).

Kernel mapping: q|k|v (self), k|v (cross) and wi_0|wi_1 are adjacent in the flat buffer, so each is ONE GEMM; the bias reaches
the attention kernels as one fp32 vector per head over the offset k - q (`t5_bias.rel_bias_vector`), and its gradient comes
back the same way (deterministic diagonal sums inside the dQ kernel, csrc/attention_bwd.cu) and is scattered onto the
[buckets, heads] table on the host side of the step (`t5_bias.scatter_rel_grad`). The encoder-output gradient is the sum of
every decoder layer's K|V-projection dgrad: accumulated in fp32 by the GEMM epilogue, rounded to bf16 once.
Dropout must be 0 (parity / benchmark setting); a non-zero value is rejected loudly.
"""
import math
from types import SimpleNamespace

import torch
from torch import nn

from .. import lib as L
from .. import ops
from ..flat import FlatBuffers, FlatSpec
from . import t5_bias as TB


class _Holder(nn.Module):
    pass


class MT5ForConditionalGeneration(nn.Module):
    def __init__(self, config, device=None, world_size=None, seed=0):
        super().__init__()
        self.config = config
        if world_size is None:
            import torch.distributed as dist
            world_size = dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1
        dev = torch.device(device if device is not None else f"cuda:{torch.cuda.current_device()}"
                           if torch.cuda.is_available() else "cuda")
        if dev.type != "cuda":
            raise RuntimeError("fsb200 MT5ForConditionalGeneration runs on CUDA only (no CPU fallback on the product path)")
        g = lambda k, d=None: getattr(config, k, d)
        self.d, self.dk, self.nh, self.ff = g("d_model"), g("d_kv"), g("num_heads"), g("d_ff")
        self.ne = g("num_layers")
        self.nd = g("num_decoder_layers") or self.ne
        self.V = g("vocab_size")
        self.eps = g("layer_norm_epsilon", 1e-6)
        self.nbuckets = g("relative_attention_num_buckets", 32)
        self.maxdist = g("relative_attention_max_distance", 128)
        self.pad_id = g("pad_token_id", 0)
        self.start_id = g("decoder_start_token_id", 0)
        if self.start_id is None:
            self.start_id = self.pad_id
        if g("dropout_rate", 0.0) not in (0, 0.0):
            raise RuntimeError(f"fsb200 MT5: dropout_rate={g('dropout_rate')} — dropout is not implemented; set it to 0")
        if g("feed_forward_proj", "gated-gelu") != "gated-gelu":
            raise RuntimeError("fsb200 MT5: only feed_forward_proj='gated-gelu' (T5 v1.1 / mT5) is implemented")
        # transformers 5.x FORCES tie_word_embeddings=True for MT5 configs (configuration_mt5.py __post_init__: "we have to tie
        # always") and applies NO d^-0.5 rescale to the decoder output (modeling_mt5.py:1141-1143) — that is what the reference
        # script gets from `MT5ForConditionalGeneration(config)` with the installed library, and what the goldens pin. Older
        # releases honoured mT5's untied head; both are implemented, neither rescales.
        self.tied = bool(g("tie_word_embeddings", False))
        if self.dk not in (64, 128) or self.V % 8 or self.d % 8 or self.ff % 8:
            raise RuntimeError("fsb200 MT5: d_kv must be 64/128 and vocab/d_model/d_ff multiples of 8 (pad the vocab)")
        d, inner, ff, V = self.d, self.nh * self.dk, self.ff, self.V
        self.inner = inner

        spec = FlatSpec()
        spec.add("shared.weight", (V, d), "shared")
        for i in range(self.ne):
            p, bk = f"encoder.block.{i}.layer.", f"enc{i}"
            for n in ("q", "k", "v"):                               # adjacent: one [3*inner, d] GEMM operand
                spec.add(p + f"0.SelfAttention.{n}.weight", (inner, d), bk)
            spec.add(p + "0.SelfAttention.o.weight", (d, inner), bk)
            if i == 0:
                spec.add(p + "0.SelfAttention.relative_attention_bias.weight", (self.nbuckets, self.nh), bk)
            spec.add(p + "0.layer_norm.weight", (d,), bk)
            spec.add(p + "1.DenseReluDense.wi_0.weight", (ff, d), bk)   # wi_0 | wi_1 adjacent: one [2*ff, d] operand
            spec.add(p + "1.DenseReluDense.wi_1.weight", (ff, d), bk)
            spec.add(p + "1.DenseReluDense.wo.weight", (d, ff), bk)
            spec.add(p + "1.layer_norm.weight", (d,), bk)
        spec.add("encoder.final_layer_norm.weight", (d,), "head")
        for i in range(self.nd):
            p, bk = f"decoder.block.{i}.layer.", f"dec{i}"
            for n in ("q", "k", "v"):
                spec.add(p + f"0.SelfAttention.{n}.weight", (inner, d), bk)
            spec.add(p + "0.SelfAttention.o.weight", (d, inner), bk)
            if i == 0:
                spec.add(p + "0.SelfAttention.relative_attention_bias.weight", (self.nbuckets, self.nh), bk)
            spec.add(p + "0.layer_norm.weight", (d,), bk)
            spec.add(p + "1.EncDecAttention.q.weight", (inner, d), bk)
            spec.add(p + "1.EncDecAttention.k.weight", (inner, d), bk)  # k | v adjacent: one [2*inner, d] operand
            spec.add(p + "1.EncDecAttention.v.weight", (inner, d), bk)
            spec.add(p + "1.EncDecAttention.o.weight", (d, inner), bk)
            spec.add(p + "1.layer_norm.weight", (d,), bk)
            spec.add(p + "2.DenseReluDense.wi_0.weight", (ff, d), bk)
            spec.add(p + "2.DenseReluDense.wi_1.weight", (ff, d), bk)
            spec.add(p + "2.DenseReluDense.wo.weight", (d, ff), bk)
            spec.add(p + "2.layer_norm.weight", (d,), bk)
        spec.add("decoder.final_layer_norm.weight", (d,), "head")
        if not self.tied:
            spec.add("lm_head.weight", (V, d), "head")
        self.flat = FlatBuffers(spec, dev, world_size=world_size)

        # module tree mirroring HF's parameter names (state_dict / named_parameters / weight-decay grouping BY NAME)
        self._p = {}
        for name in self.flat.offsets:
            prm = nn.Parameter(self.flat.view(name), requires_grad=True)
            prm.main_grad = self.flat.view(name, grad=True)
            self._p[name] = prm
            mod = self
            parts = name.split(".")
            for part in parts[:-1]:
                if not hasattr(mod, part):
                    setattr(mod, part, _Holder())
                mod = getattr(mod, part)
            setattr(mod, parts[-1], prm)

        if self.tied:
            self.lm_head = _Holder()
            self.lm_head.weight = self._p["shared.weight"]      # same Parameter object: named_parameters() lists it once
        self._head = "shared.weight" if self.tied else "lm_head.weight"
        fl = self.flat
        E, D = "encoder.block.{}.layer.", "decoder.block.{}.layer."
        self._e_qkv = [fl.span(E.format(i) + "0.SelfAttention.q.weight", 3 * inner, d) for i in range(self.ne)]
        self._e_dqkv = [fl.span(E.format(i) + "0.SelfAttention.q.weight", 3 * inner, d, grad=True) for i in range(self.ne)]
        self._e_wi = [fl.span(E.format(i) + "1.DenseReluDense.wi_0.weight", 2 * ff, d) for i in range(self.ne)]
        self._e_dwi = [fl.span(E.format(i) + "1.DenseReluDense.wi_0.weight", 2 * ff, d, grad=True) for i in range(self.ne)]
        self._d_qkv = [fl.span(D.format(i) + "0.SelfAttention.q.weight", 3 * inner, d) for i in range(self.nd)]
        self._d_dqkv = [fl.span(D.format(i) + "0.SelfAttention.q.weight", 3 * inner, d, grad=True) for i in range(self.nd)]
        self._d_kv = [fl.span(D.format(i) + "1.EncDecAttention.k.weight", 2 * inner, d) for i in range(self.nd)]
        self._d_dkv = [fl.span(D.format(i) + "1.EncDecAttention.k.weight", 2 * inner, d, grad=True) for i in range(self.nd)]
        self._d_wi = [fl.span(D.format(i) + "2.DenseReluDense.wi_0.weight", 2 * ff, d) for i in range(self.nd)]
        self._d_dwi = [fl.span(D.format(i) + "2.DenseReluDense.wi_0.weight", 2 * ff, d, grad=True) for i in range(self.nd)]

        self.reset_parameters(seed)
        self.accumulate_grads, self.loss_scale, self.grad_hook = False, 1.0, None

    def P(self, name):
        return self._p[name]

    @torch.no_grad()
    def reset_parameters(self, seed=0):
        """HF MT5PreTrainedModel._init_weights with initializer_factor 1 (mt5/modeling_mt5.py)."""
        d, dk, nh, ff = self.d, self.dk, self.nh, self.ff
        gen = torch.Generator(device=self.flat.params.device).manual_seed(seed)
        for name, prm in self._p.items():
            if name.endswith("layer_norm.weight"):
                prm.fill_(1.0)
                continue
            if name in ("shared.weight", "lm_head.weight"):
                std = 1.0
            elif name.endswith(".q.weight"):
                std = (d * dk) ** -0.5
            elif name.endswith(".k.weight") or name.endswith(".v.weight") or "relative_attention_bias" in name:
                std = d ** -0.5
            elif name.endswith("Attention.o.weight"):
                std = (nh * dk) ** -0.5
            elif "wi_" in name:
                std = d ** -0.5
            else:  # DenseReluDense.wo
                std = ff ** -0.5
            prm.normal_(0.0, std, generator=gen)

    def cuda(self, device=None):
        return self

    def half(self):
        return self

    def bfloat16(self):
        return self

    def to(self, *args, **kwargs):
        return self

    @torch.no_grad()
    def load_reference_state_dict(self, sd):
        """HF state dict (fp32 / bf16, any device). `encoder.embed_tokens.weight` / `decoder.embed_tokens.weight` are aliases
        of `shared.weight` in HF and are ignored."""
        for k, prm in self._p.items():
            if k not in sd:
                raise KeyError(f"missing key in state dict: {k}")
            if tuple(sd[k].shape) != tuple(prm.shape):
                raise ValueError(f"shape mismatch for {k}: {tuple(sd[k].shape)} vs {tuple(prm.shape)}")
            prm.copy_(sd[k].to(device=prm.device, dtype=prm.dtype))

    def save_pretrained(self, path, **_):
        """HF-style export (config.json + pytorch_model.bin with this class's HF key names): fsb200/models/export.py."""
        from .export import save_pretrained
        save_pretrained(self, path)

    # ---- engine hooks -----------------------------------------------------------------------------------------------
    def _done(self, bucket):
        if self.grad_hook is not None and bucket in self.flat.bucket_index:   # "head" does not exist with a tied LM head
            self.grad_hook(bucket)

    def _need(self, bucket):
        hook = getattr(self, "param_hook", None)
        if hook is not None:
            hook(bucket)

    def _begin_backward(self):
        hook = getattr(self, "backward_begin_hook", None)
        if hook is not None:
            hook()

    # ---- forward ----------------------------------------------------------------------------------------------------
    def _shift_right(self, labels):
        """MT5 _shift_right (:592): decoder_input_ids = [start] + labels[:-1], with -100 replaced by the pad id."""
        dec = labels.new_zeros(labels.shape)
        dec[:, 1:] = labels[:, :-1]
        dec[:, 0] = self.start_id
        return dec.masked_fill(dec == -100, self.pad_id)

    def forward(self, input_ids=None, attention_mask=None, labels=None, decoder_input_ids=None, return_logits=False, **_):
        dev = self.flat.params.device
        B, Se = input_ids.shape
        ids = input_ids.to(device=dev, dtype=torch.int64).contiguous()
        mask = None
        if attention_mask is not None and not bool(attention_mask.all()):
            mask = attention_mask.to(device=dev, dtype=torch.uint8).contiguous()
        lab = None if labels is None else labels.to(device=dev, dtype=torch.int64).contiguous()
        if decoder_input_ids is None:
            if lab is None:
                raise ValueError("fsb200 MT5: pass labels or decoder_input_ids")
            dec_ids = self._shift_right(lab)
        else:
            dec_ids = decoder_input_ids.to(device=dev, dtype=torch.int64).contiguous()
        Sd = dec_ids.shape[1]
        if lab is not None and torch.is_grad_enabled():
            loss, logits = _T5Step.apply(self, ids.view(-1), dec_ids.view(-1), mask, lab.view(-1), B, Se, Sd, return_logits,
                                         self.P("decoder.final_layer_norm.weight"))
        else:
            loss, logits, _ = self._forward_impl(ids.view(-1), dec_ids.view(-1), mask, None if lab is None else lab.view(-1),
                                                 B, Se, Sd, save=False, want_logits=True)
        return SimpleNamespace(loss=loss, logits=None if logits is None else logits.view(B, Sd, self.V),
                               past_key_values=None, encoder_last_hidden_state=None)

    def _norm(self, prev, x, name):
        """pre-norm with the pending residual add fused in: returns (normed, rstd, residual stream)."""
        if prev is None:
            return ops.rmsnorm_fwd(x, self.P(name).data, self.eps)
        return ops.rmsnorm_fwd(prev, self.P(name).data, self.eps, residual=x)

    def _forward_impl(self, ids, dec_ids, mask, lab, B, Se, Sd, save, want_logits):
        d, nh, dk, inner, ff = self.d, self.nh, self.dk, self.inner, self.ff
        P = self.P
        Te, Td = B * Se, B * Sd
        self._need("no_decay"); self._need("shared")
        W = P("shared.weight").data
        # relative-position bias vectors (fp32 [heads, 2S - 1]) from the two [buckets, heads] tables
        rel_e = TB.rel_bias_vector(P("encoder.block.0.layer.0.SelfAttention.relative_attention_bias.weight").data, Se, Se, True,
                                   self.nbuckets, self.maxdist)
        rel_d = TB.rel_bias_vector(P("decoder.block.0.layer.0.SelfAttention.relative_attention_bias.weight").data, Sd, Sd, False,
                                   self.nbuckets, self.maxdist)
        # ---- encoder
        x, prev = ops.embedding_fwd(ids, W), None
        eacts = []
        for i in range(self.ne):
            p = f"encoder.block.{i}.layer."
            self._need(f"enc{i}")
            h1, r1, x = self._norm(prev, x, p + "0.layer_norm.weight")
            qkv = ops.gemm(L.GEMM_NT, h1, self._e_qkv[i])
            q5 = qkv.view(B, Se, 3, nh, dk)
            o, lse = ops.sdpa_fwd(q5[:, :, 0], q5[:, :, 1], q5[:, :, 2], 1.0, False, kv_mask=mask, rel_bias=rel_e)
            a = ops.gemm(L.GEMM_NT, o.view(Te, inner), P(p + "0.SelfAttention.o.weight").data)
            h2, r2, x1 = self._norm(a, x, p + "1.layer_norm.weight")
            gu = ops.gemm(L.GEMM_NT, h2, self._e_wi[i])
            act = ops.glu_fwd(L.ACT_GELU_TANH, gu[:, :ff], gu[:, ff:])
            m = ops.gemm(L.GEMM_NT, act, P(p + "1.DenseReluDense.wo.weight").data)
            if save:
                eacts.append((x, r1, h1, qkv, o, lse, x1, r2, h2, gu, act))
            x, prev = x1, m
        self._need("head")
        enc_h, rfe, xfe = self._norm(prev, x, "encoder.final_layer_norm.weight")
        # ---- decoder
        y, prev = ops.embedding_fwd(dec_ids, W), None
        dacts = []
        for i in range(self.nd):
            p = f"decoder.block.{i}.layer."
            self._need(f"dec{i}")
            h1, r1, y = self._norm(prev, y, p + "0.layer_norm.weight")
            qkv = ops.gemm(L.GEMM_NT, h1, self._d_qkv[i])
            q5 = qkv.view(B, Sd, 3, nh, dk)
            o, lse = ops.sdpa_fwd(q5[:, :, 0], q5[:, :, 1], q5[:, :, 2], 1.0, True, rel_bias=rel_d)
            a = ops.gemm(L.GEMM_NT, o.view(Td, inner), P(p + "0.SelfAttention.o.weight").data)
            h2, r2, y1 = self._norm(a, y, p + "1.layer_norm.weight")
            qc = ops.gemm(L.GEMM_NT, h2, P(p + "1.EncDecAttention.q.weight").data)
            kvc = ops.gemm(L.GEMM_NT, enc_h, self._d_kv[i])
            kv5 = kvc.view(B, Se, 2, nh, dk)
            oc, lsec = ops.sdpa_fwd(qc.view(B, Sd, nh, dk), kv5[:, :, 0], kv5[:, :, 1], 1.0, False, kv_mask=mask)
            ac = ops.gemm(L.GEMM_NT, oc.view(Td, inner), P(p + "1.EncDecAttention.o.weight").data)
            h3, r3, y2 = self._norm(ac, y1, p + "2.layer_norm.weight")
            gu = ops.gemm(L.GEMM_NT, h3, self._d_wi[i])
            act = ops.glu_fwd(L.ACT_GELU_TANH, gu[:, :ff], gu[:, ff:])
            m = ops.gemm(L.GEMM_NT, act, P(p + "2.DenseReluDense.wo.weight").data)
            if save:
                dacts.append((y, r1, h1, qkv, o, lse, y1, r2, h2, qc, kvc, oc, lsec, y2, r3, h3, gu, act))
            y, prev = y2, m
        hf, rfd, xfd = self._norm(prev, y, "decoder.final_layer_norm.weight")
        logits = ops.gemm(L.GEMM_NT, hf, P(self._head).data)
        loss, ctx = None, None
        if lab is not None:
            keep = logits.clone() if (want_logits and save) else None
            loss, dlogits, _ = ops.softmax_xent(logits, lab, Sd, shift=0, grad_scale=self.loss_scale,
                                                dlogits="inplace" if save else None)
            if save:
                ctx = (eacts, dacts, enc_h, rfe, xfe, hf, rfd, xfd, dlogits, ids, dec_ids, mask, rel_e, rel_d, B, Se, Sd)
                logits = keep
        return loss, (logits if want_logits else None), ctx

    # ---- backward ---------------------------------------------------------------------------------------------------
    def _backward_impl(self, ctx, gloss):
        eacts, dacts, enc_h, rfe, xfe, hf, rfd, xfd, dlogits, ids, dec_ids, mask, rel_e, rel_d, B, Se, Sd = ctx
        d, nh, dk, inner, ff = self.d, self.nh, self.dk, self.inner, self.ff
        P = self.P
        Te, Td = B * Se, B * Sd
        acc = self.accumulate_grads
        dev = self.flat.params.device
        self._begin_backward()
        if gloss is not None:
            ops.scale_inplace(dlogits, gloss)
        Wlm = P(self._head)
        dhf = ops.gemm(L.GEMM_NN, dlogits, Wlm.data)
        ops.gemm(L.GEMM_TN, dlogits, hf, out=Wlm.main_grad, accumulate=acc)   # tied head: written first, the embeddings add later
        del dlogits
        self._done("head")                      # lm_head is the bucket's only decayed parameter (the norms are no-decay)
        fs = P("decoder.final_layer_norm.weight")
        dy = ops.rmsnorm_bwd(dhf, xfd, fs.data, rfd, fs.main_grad, accumulate=acc)
        drel_e = torch.zeros_like(rel_e)
        drel_d = torch.zeros_like(rel_d)
        denc32 = torch.empty((Te, d), dtype=torch.float32, device=dev)   # sum over decoder layers of the K|V dgrads
        for i in reversed(range(self.nd)):
            p = f"decoder.block.{i}.layer."
            y, r1, h1, qkv, o, lse, y1, r2, h2, qc, kvc, oc, lsec, y2, r3, h3, gu, act = dacts[i]
            dacts[i] = None
            wo = P(p + "2.DenseReluDense.wo.weight")
            dact = ops.gemm(L.GEMM_NN, dy, wo.data)
            ops.gemm(L.GEMM_TN, dy, act, out=wo.main_grad, accumulate=acc)
            dgu = torch.empty_like(gu)
            ops.glu_bwd(L.ACT_GELU_TANH, dact, gu[:, :ff], gu[:, ff:], dgu[:, :ff], dgu[:, ff:])
            dh3 = ops.gemm(L.GEMM_NN, dgu, self._d_wi[i])
            ops.gemm(L.GEMM_TN, dgu, h3, out=self._d_dwi[i], accumulate=acc)
            s3 = P(p + "2.layer_norm.weight")
            dy2 = ops.rmsnorm_bwd(dh3, y2, s3.data, r3, s3.main_grad, accumulate=acc, dres=dy)
            # cross-attention
            woc = P(p + "1.EncDecAttention.o.weight")
            doc = ops.gemm(L.GEMM_NN, dy2, woc.data)
            ops.gemm(L.GEMM_TN, dy2, oc.view(Td, inner), out=woc.main_grad, accumulate=acc)
            dqc = torch.empty_like(qc)
            dkvc = torch.empty_like(kvc)
            kv5, dkv5 = kvc.view(B, Se, 2, nh, dk), dkvc.view(B, Se, 2, nh, dk)
            ops.sdpa_bwd(qc.view(B, Sd, nh, dk), kv5[:, :, 0], kv5[:, :, 1], oc, doc.view(B, Sd, nh, dk), lsec, 1.0, False,
                         dqc.view(B, Sd, nh, dk), dkv5[:, :, 0], dkv5[:, :, 1], kv_mask=mask)
            wqc = P(p + "1.EncDecAttention.q.weight")
            dh2 = ops.gemm(L.GEMM_NN, dqc, wqc.data)
            ops.gemm(L.GEMM_TN, dqc, h2, out=wqc.main_grad, accumulate=acc)
            ops.gemm(L.GEMM_NN, dkvc, self._d_kv[i], out=denc32, accumulate=(i != self.nd - 1))
            ops.gemm(L.GEMM_TN, dkvc, enc_h, out=self._d_dkv[i], accumulate=acc)
            s2 = P(p + "1.layer_norm.weight")
            dy1 = ops.rmsnorm_bwd(dh2, y1, s2.data, r2, s2.main_grad, accumulate=acc, dres=dy2)
            # causal self-attention with the decoder's relative-position bias
            wos = P(p + "0.SelfAttention.o.weight")
            do = ops.gemm(L.GEMM_NN, dy1, wos.data)
            ops.gemm(L.GEMM_TN, dy1, o.view(Td, inner), out=wos.main_grad, accumulate=acc)
            dqkv = torch.empty_like(qkv)
            q5, d5 = qkv.view(B, Sd, 3, nh, dk), dqkv.view(B, Sd, 3, nh, dk)
            ops.sdpa_bwd(q5[:, :, 0], q5[:, :, 1], q5[:, :, 2], o, do.view(B, Sd, nh, dk), lse, 1.0, True,
                         d5[:, :, 0], d5[:, :, 1], d5[:, :, 2], rel_bias=rel_d, drel_bias=drel_d)
            dh1 = ops.gemm(L.GEMM_NN, dqkv, self._d_qkv[i])
            ops.gemm(L.GEMM_TN, dqkv, h1, out=self._d_dqkv[i], accumulate=acc)
            s1 = P(p + "0.layer_norm.weight")
            dy = ops.rmsnorm_bwd(dh1, y, s1.data, r1, s1.main_grad, accumulate=acc, dres=dy1)
            if i == 0:
                self._table_grad("decoder.block.0.layer.0.SelfAttention.relative_attention_bias.weight", drel_d, Sd, False, acc)
            self._done(f"dec{i}")
        ddec_emb = dy                                                     # gradient w.r.t. the decoder's input embeddings
        # ---- encoder
        if self.nd > 0:
            denc = ops.cast_f32_to_bf16(denc32)
        else:
            denc = torch.zeros((Te, d), dtype=torch.bfloat16, device=dev)
        del denc32
        es = P("encoder.final_layer_norm.weight")
        dx = ops.rmsnorm_bwd(denc, xfe, es.data, rfe, es.main_grad, accumulate=acc)
        for i in reversed(range(self.ne)):
            p = f"encoder.block.{i}.layer."
            x, r1, h1, qkv, o, lse, x1, r2, h2, gu, act = eacts[i]
            eacts[i] = None
            wo = P(p + "1.DenseReluDense.wo.weight")
            dact = ops.gemm(L.GEMM_NN, dx, wo.data)
            ops.gemm(L.GEMM_TN, dx, act, out=wo.main_grad, accumulate=acc)
            dgu = torch.empty_like(gu)
            ops.glu_bwd(L.ACT_GELU_TANH, dact, gu[:, :ff], gu[:, ff:], dgu[:, :ff], dgu[:, ff:])
            dh2 = ops.gemm(L.GEMM_NN, dgu, self._e_wi[i])
            ops.gemm(L.GEMM_TN, dgu, h2, out=self._e_dwi[i], accumulate=acc)
            s2 = P(p + "1.layer_norm.weight")
            dx1 = ops.rmsnorm_bwd(dh2, x1, s2.data, r2, s2.main_grad, accumulate=acc, dres=dx)
            wos = P(p + "0.SelfAttention.o.weight")
            do = ops.gemm(L.GEMM_NN, dx1, wos.data)
            ops.gemm(L.GEMM_TN, dx1, o.view(Te, inner), out=wos.main_grad, accumulate=acc)
            dqkv = torch.empty_like(qkv)
            q5, d5 = qkv.view(B, Se, 3, nh, dk), dqkv.view(B, Se, 3, nh, dk)
            ops.sdpa_bwd(q5[:, :, 0], q5[:, :, 1], q5[:, :, 2], o, do.view(B, Se, nh, dk), lse, 1.0, False,
                         d5[:, :, 0], d5[:, :, 1], d5[:, :, 2], kv_mask=mask, rel_bias=rel_e, drel_bias=drel_e)
            dh1 = ops.gemm(L.GEMM_NN, dqkv, self._e_qkv[i])
            ops.gemm(L.GEMM_TN, dqkv, h1, out=self._e_dqkv[i], accumulate=acc)
            s1 = P(p + "0.layer_norm.weight")
            dx = ops.rmsnorm_bwd(dh1, x, s1.data, r1, s1.main_grad, accumulate=acc, dres=dx1)
            if i == 0:
                self._table_grad("encoder.block.0.layer.0.SelfAttention.relative_attention_bias.weight", drel_e, Se, True, acc)
            self._done(f"enc{i}")
        Wg = P("shared.weight").main_grad
        if not acc and not self.tied:
            Wg.zero_()
        ops.embedding_bwd(ids, dx, Wg)          # encoder inputs
        ops.embedding_bwd(dec_ids, ddec_emb, Wg)  # decoder inputs share the table
        self._done("shared")
        self._done("no_decay")

    def _table_grad(self, name, drel, S, bidirectional, acc):
        """[heads, 2S - 1] gradient of the bias vector -> [buckets, heads] gradient of the embedding table (fixed-index
        index_add: deterministic), written to the flat gradient buffer."""
        g = TB.scatter_rel_grad(drel, S, S, bidirectional, self.nbuckets, self.maxdist)
        mg = self.P(name).main_grad
        if acc:
            mg.copy_((mg.float() + g).to(mg.dtype))
        else:
            mg.copy_(g.to(mg.dtype))


class _T5Step(torch.autograd.Function):
    """The whole encoder-decoder as one autograd node (cf. _LlamaStep)."""

    @staticmethod
    def forward(ctx, model, ids, dec_ids, mask, lab, B, Se, Sd, want_logits, anchor):
        loss, logits, saved = model._forward_impl(ids, dec_ids, mask, lab, B, Se, Sd, save=True, want_logits=want_logits)
        ctx.model, ctx.saved = model, saved
        ctx.mark_non_differentiable(*([logits] if logits is not None else []))
        return loss, logits

    @staticmethod
    def backward(ctx, gloss, _glogits):
        model, saved = ctx.model, ctx.saved
        ctx.saved = None
        model._backward_impl(saved, None if gloss is None else gloss)
        return (None,) * 10


def t5_flops_per_step(cfg, B, Se, Sd):
    """Algorithmic FLOPs of one forward + backward (3x forward) over B samples — SURVEY.md §8(d) C5 formula: matmul parameters
    touched per token x 6, plus attention (encoder full, decoder causal-counted, cross full)."""
    d, inner, ff, V = cfg["d_model"], cfg["num_heads"] * cfg["d_kv"], cfg["d_ff"], cfg["vocab_size"]
    Le, Ld = cfg["num_layers"], cfg.get("num_decoder_layers") or cfg["num_layers"]
    enc_mm = Le * (4 * d * inner + 3 * d * ff)
    dec_mm = Ld * (4 * d * inner + 2 * d * inner + 3 * d * ff) + V * d      # self + cross q/o + FFN + head, per decoder token
    cross_kv = Ld * 2 * d * inner                                            # per ENCODER token
    mm = 6.0 * B * (Se * (enc_mm + cross_kv) + Sd * dec_mm)
    attn = 3.0 * 4.0 * inner * B * (Le * Se * Se + Ld * Sd * Sd / 2 + Ld * Se * Sd)
    return mm + attn
