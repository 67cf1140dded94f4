"""Erlangshen-BERT / Erlangshen-MegatronBERT on the fsb200 kernels — drop-ins for the two HF classes the reference's
MLM pretraining scripts instantiate:
    transformers.BertForMaskedLM            examples/pretrain_bert/pretrain_bert.py:135-137           (config 1)
    transformers.MegatronBertForPreTraining examples/pretrain_erlangshen_bert/pretrain_erlangshen.py:138-141 (config 3)
The arithmetic restated here lives in 3P `transformers`:
  BERT          post-LN blocks  x = LN(x + W_o attn(x)); x = LN(x + W_2 act(W_1 x)); embeddings word+pos+type -> LN.
  MegatronBERT  pre-LN blocks   x + W_o attn(ln(x)); x + W_2 act(W_1 ln(x)); final encoder ln; embeddings WITHOUT LN;
                pooler tanh(W x[:,0]) + NSP head; loss = CE(mlm) + CE(nsp).
Both: separate q/k/v Linear(h,h)+bias (laid out adjacently -> one [3h,h] GEMM), softmax(QK^T/sqrt(hn) + padding mask),
LayerNorm eps 1e-12, MLM head = dense + act + LN + decoder tied to the word embeddings + bias, CE ignore_index -100.
State-dict keys follow HF. Dropout must be 0 (rejected loudly otherwise).
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


def _set(root, dotted, value):
    parts = dotted.split(".")
    mod = root
    for p in parts[:-1]:
        if not hasattr(mod, p):
            setattr(mod, p, nn.ModuleList() if False else _Holder())
        mod = getattr(mod, p)
    setattr(mod, parts[-1], value)


class _BertFamily(nn.Module):
    PRE_LN = False

    def __init__(self, config, device=None, world_size=None, seed=0):
        super().__init__()
        self.config = config
        if world_size is None:
            import torch.distributed as dist
            world_size = dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1
        dev = torch.device(device if device is not None else f"cuda:{torch.cuda.current_device()}"
                           if torch.cuda.is_available() else "cuda")
        if dev.type != "cuda":
            raise RuntimeError("fsb200 BERT models run on CUDA only (no CPU fallback on the product path)")
        g = lambda k, d=None: getattr(config, k, d)
        self.h, self.nl, self.nh, self.V = g("hidden_size"), g("num_hidden_layers"), g("num_attention_heads"), g("vocab_size")
        self.ff, self.npos, self.ntype = g("intermediate_size"), g("max_position_embeddings", 512), g("type_vocab_size", 2)
        self.eps = g("layer_norm_eps", 1e-12)
        for k in ("hidden_dropout_prob", "attention_probs_dropout_prob"):
            if g(k, 0.0) not in (0, 0.0):
                raise RuntimeError(f"fsb200 BERT: {k}={g(k)} — dropout is not implemented; set it to 0")
        act = g("hidden_act", "gelu")
        if act not in ("gelu", "gelu_new"):
            raise RuntimeError(f"fsb200 BERT: hidden_act={act!r} not implemented (gelu, gelu_new)")
        self.epi = L.EPI_GELU_ERF if act == "gelu" else L.EPI_GELU_TANH
        self.act = L.ACT_GELU_ERF if act == "gelu" else L.ACT_GELU_TANH
        h = self.h
        self.hn = h // self.nh
        if self.hn not in (64, 128) or self.V % 8 or h % 128:
            raise RuntimeError("fsb200 BERT: head dim must be 64/128, vocab a multiple of 8, hidden a multiple of 128")
        pre = self.PRE_LN
        spec = FlatSpec()
        E = "bert.embeddings."
        spec.add(E + "word_embeddings.weight", (self.V, h), "emb")
        spec.add(E + "position_embeddings.weight", (self.npos, h), "emb")
        spec.add(E + "token_type_embeddings.weight", (self.ntype, h), "emb")
        if not pre:
            spec.add(E + "LayerNorm.weight", (h,), "emb"); spec.add(E + "LayerNorm.bias", (h,), "emb")
        for i in range(self.nl):
            p, bk = f"bert.encoder.layer.{i}.", f"layer{i}"
            if pre:
                spec.add(p + "attention.ln.weight", (h,), bk); spec.add(p + "attention.ln.bias", (h,), bk)
            for n in ("query", "key", "value"):            # adjacent: one [3h, h] operand
                spec.add(p + f"attention.self.{n}.weight", (h, h), bk)
            for n in ("query", "key", "value"):            # adjacent in the no-decay bucket: one [3h] bias
                spec.add(p + f"attention.self.{n}.bias", (h,), bk)
            spec.add(p + "attention.output.dense.weight", (h, h), bk); spec.add(p + "attention.output.dense.bias", (h,), bk)
            if pre:
                spec.add(p + "ln.weight", (h,), bk); spec.add(p + "ln.bias", (h,), bk)
            else:
                spec.add(p + "attention.output.LayerNorm.weight", (h,), bk)
                spec.add(p + "attention.output.LayerNorm.bias", (h,), bk)
            spec.add(p + "intermediate.dense.weight", (self.ff, h), bk); spec.add(p + "intermediate.dense.bias", (self.ff,), bk)
            spec.add(p + "output.dense.weight", (h, self.ff), bk); spec.add(p + "output.dense.bias", (h,), bk)
            if not pre:
                spec.add(p + "output.LayerNorm.weight", (h,), bk); spec.add(p + "output.LayerNorm.bias", (h,), bk)
        if pre:
            spec.add("bert.encoder.ln.weight", (h,), "head"); spec.add("bert.encoder.ln.bias", (h,), "head")
            spec.add("bert.pooler.dense.weight", (h, h), "head"); spec.add("bert.pooler.dense.bias", (h,), "head")
        spec.add("cls.predictions.bias", (self.V,), "head")
        spec.add("cls.predictions.transform.dense.weight", (h, h), "head")
        spec.add("cls.predictions.transform.dense.bias", (h,), "head")
        spec.add("cls.predictions.transform.LayerNorm.weight", (h,), "head")
        spec.add("cls.predictions.transform.LayerNorm.bias", (h,), "head")
        if pre:
            spec.add("cls.seq_relationship.weight", (2, h), "head"); spec.add("cls.seq_relationship.bias", (2,), "head")
        self.flat = FlatBuffers(spec, dev, world_size=world_size)
        self._p = {}
        for name in self.flat.offsets:
            prm = nn.Parameter(self.flat.view(name), requires_grad=True)
            prm.main_grad = self.flat.view(name, grad=True)
            self._p[name] = prm
            parts = name.split(".")
            mod = self
            for j, part in enumerate(parts[:-1]):
                if not hasattr(mod, part):
                    setattr(mod, part, _Holder())
                mod = getattr(mod, part)
            setattr(mod, parts[-1], prm)
        # fused q|k|v operands
        self._wqkv = [self.flat.span(f"bert.encoder.layer.{i}.attention.self.query.weight", 3 * h, h) for i in range(self.nl)]
        self._dwqkv = [self.flat.span(f"bert.encoder.layer.{i}.attention.self.query.weight", 3 * h, h, grad=True)
                       for i in range(self.nl)]
        self._bqkv = [self.flat.span(f"bert.encoder.layer.{i}.attention.self.query.bias", 1, 3 * h).view(-1)
                      for i in range(self.nl)]
        self._dbqkv = [self.flat.span(f"bert.encoder.layer.{i}.attention.self.query.bias", 1, 3 * h, grad=True).view(-1)
                       for i in range(self.nl)]
        if pre:  # NSP classifier padded to 8 outputs (pad logits = -30000 -> zero probability); parameters stay [2, h]
            self._nsp_w = torch.zeros(8, h, dtype=torch.bfloat16, device=dev)
            self._nsp_b = torch.full((8,), -30000.0, dtype=torch.bfloat16, device=dev)
        self.reset_parameters(seed)
        self.accumulate_grads, self.loss_scale, self.grad_hook = False, 1.0, None

    def P(self, name):
        return self._p[name]

    @torch.no_grad()
    def reset_parameters(self, seed=0):
        std = getattr(self.config, "initializer_range", 0.02)
        gen = torch.Generator(device=self.flat.params.device).manual_seed(seed)
        for name, prm in self._p.items():
            if name.endswith("bias"):
                prm.zero_()
            elif "LayerNorm.weight" in name or name.endswith("ln.weight"):
                prm.fill_(1.0)
            else:
                prm.normal_(0.0, std, generator=gen)

    @torch.no_grad()
    def load_reference_state_dict(self, sd):
        for k, prm in self._p.items():
            if tuple(sd[k].shape) != tuple(prm.shape):
                raise ValueError(f"shape mismatch for {k}: {tuple(sd[k].shape)} vs {tuple(prm.shape)}")
            prm.copy_(sd[k].to(device=prm.device, dtype=prm.dtype))

    def cuda(self, device=None):
        return self

    def to(self, *a, **k):
        return self

    # ---- forward ----------------------------------------------------------------------------------------------------
    def forward(self, input_ids=None, attention_mask=None, token_type_ids=None, position_ids=None, labels=None,
                next_sentence_label=None, return_logits=False, **_):
        B, S = input_ids.shape
        dev = self.flat.params.device
        c = lambda t: None if t is None else t.to(device=dev, dtype=torch.int64).contiguous().view(-1)
        ids, tt, pos, lab = c(input_ids), c(token_type_ids), c(position_ids), c(labels)
        nsl = c(next_sentence_label)
        mask = None
        if attention_mask is not None and not bool(attention_mask.all()):
            mask = attention_mask.to(device=dev, dtype=torch.uint8).contiguous()
        if lab is not None and torch.is_grad_enabled():
            anchor = self.P("cls.predictions.transform.LayerNorm.weight")
            loss, logits, nsp = _BertStep.apply(self, ids, tt, pos, mask, lab, nsl, B, S, return_logits, anchor)
        else:
            loss, logits, nsp, _ = self._forward_impl(ids, tt, pos, mask, lab, nsl, B, S, save=False, want_logits=True)
        out = SimpleNamespace(loss=loss, logits=None if logits is None else logits.view(B, S, self.V),
                              hidden_states=None, attentions=None)
        out.prediction_logits = out.logits
        out.seq_relationship_logits = None if nsp is None else nsp[:, :2]
        return out

    def _forward_impl(self, ids, tt, pos, mask, lab, nsl, B, S, save, want_logits):
        h, nh, hn, pre = self.h, self.nh, self.hn, self.PRE_LN
        T = B * S
        P = self.P
        E = "bert.embeddings."
        scale = 1.0 / math.sqrt(hn)
        self._need("no_decay"); self._need("emb")
        emb = ops.embedding_fwd(ids, P(E + "word_embeddings.weight").data, pos=pos,
                                P=P(E + "position_embeddings.weight").data, token_type=tt,
                                T=P(E + "token_type_embeddings.weight").data, seq_len=S)
        acts = []
        if pre:
            x, prev_m, emb_ctx = emb, None, None
        else:
            x, st_e, _ = ops.layernorm_fwd(emb, P(E + "LayerNorm.weight").data, P(E + "LayerNorm.bias").data, self.eps)
            emb_ctx = (emb, st_e)
        for i in range(self.nl):
            p = f"bert.encoder.layer.{i}."
            self._need(f"layer{i}")
            if pre:
                h1, st1, x = ops.layernorm_fwd(x if prev_m is None else prev_m, P(p + "attention.ln.weight").data,
                                               P(p + "attention.ln.bias").data, self.eps,
                                               residual=None if prev_m is None else x)
                attn_in = h1
            else:
                attn_in = x
            qkv = ops.gemm(L.GEMM_NT, attn_in, self._wqkv[i], bias=self._bqkv[i])
            q5 = qkv.view(B, S, 3, nh, hn)
            o, lse = ops.sdpa_fwd(q5[:, :, 0], q5[:, :, 1], q5[:, :, 2], scale, False, kv_mask=mask)
            a = ops.gemm(L.GEMM_NT, o.view(T, h), P(p + "attention.output.dense.weight").data,
                         bias=P(p + "attention.output.dense.bias").data)
            if pre:
                h2, st2, x1 = ops.layernorm_fwd(a, P(p + "ln.weight").data, P(p + "ln.bias").data, self.eps, residual=x)
            else:
                h2, st2, x1 = ops.layernorm_fwd(a, P(p + "attention.output.LayerNorm.weight").data,
                                                P(p + "attention.output.LayerNorm.bias").data, self.eps, residual=x)
            prea = torch.empty((T, self.ff), dtype=torch.bfloat16, device=x.device) if save else None
            f = ops.gemm(L.GEMM_NT, h2, P(p + "intermediate.dense.weight").data, bias=P(p + "intermediate.dense.bias").data,
                         epilogue=self.epi, aux=prea)
            m = ops.gemm(L.GEMM_NT, f, P(p + "output.dense.weight").data, bias=P(p + "output.dense.bias").data)
            if pre:
                if save:
                    acts.append((x, st1, h1, qkv, o, lse, x1, st2, h2, prea, f))
                x, prev_m = x1, m
            else:
                xo, st3, s2 = ops.layernorm_fwd(m, P(p + "output.LayerNorm.weight").data,
                                                P(p + "output.LayerNorm.bias").data, self.eps, residual=h2)
                if save:
                    acts.append((x, qkv, o, lse, x1, st2, h2, prea, f, s2, st3))
                x = xo
        self._need("head")
        if pre:
            hf, stf, xf = ops.layernorm_fwd(prev_m, P("bert.encoder.ln.weight").data, P("bert.encoder.ln.bias").data,
                                            self.eps, residual=x)
        else:
            hf, stf, xf = x, None, None
        # MLM head: dense + act + LN + tied decoder + bias (on every position)
        tpre = torch.empty((T, h), dtype=torch.bfloat16, device=hf.device) if save else None
        tf = ops.gemm(L.GEMM_NT, hf, P("cls.predictions.transform.dense.weight").data,
                      bias=P("cls.predictions.transform.dense.bias").data, epilogue=self.epi, aux=tpre)
        tn, stt, _ = ops.layernorm_fwd(tf, P("cls.predictions.transform.LayerNorm.weight").data,
                                       P("cls.predictions.transform.LayerNorm.bias").data, self.eps)
        logits = ops.gemm(L.GEMM_NT, tn, P(E + "word_embeddings.weight").data, bias=P("cls.predictions.bias").data)
        loss, ctx, nsp_logits = None, None, None
        nsp_ctx = None
        if pre:
            first = hf.view(B, S, h)[:, 0, :]                     # strided [B, h] view, row stride S*h
            ppre = torch.empty((B, h), dtype=torch.bfloat16, device=hf.device)
            ops.gemm(L.GEMM_NT, first, P("bert.pooler.dense.weight").data, bias=P("bert.pooler.dense.bias").data, out=ppre)
            pooled = ops.act_fwd(L.ACT_TANH, ppre)
            self._nsp_w[:2].copy_(P("cls.seq_relationship.weight").data)
            self._nsp_b[:2].copy_(P("cls.seq_relationship.bias").data)
            nsp_logits = ops.gemm(L.GEMM_NT, pooled, self._nsp_w, bias=self._nsp_b)   # [B, 8]
            nsp_ctx = (first, ppre, pooled)
        if lab is not None:
            keep = logits.clone() if (want_logits and save) else None
            loss, dlogits, _ = ops.softmax_xent(logits, lab, S, shift=0, grad_scale=self.loss_scale,
                                                dlogits="inplace" if save else None)
            dnsp = None
            if pre and nsl is not None:
                keep_nsp = nsp_logits.clone()
                nloss, dnsp, _ = ops.softmax_xent(nsp_logits, nsl, 1, shift=0, grad_scale=self.loss_scale,
                                                  dlogits="inplace" if save else None)
                loss = loss + nloss                               # modeling_megatron_bert.py:776-779
                nsp_logits = keep_nsp
            if save:
                ctx = (acts, emb_ctx, hf, stf, xf, tpre, tf, stt, tn, dlogits, nsp_ctx, dnsp, ids, tt, pos, mask, B, S)
                logits = keep
        return loss, (logits if want_logits else None), nsp_logits, ctx

    # ---- backward ---------------------------------------------------------------------------------------------------
    def _backward_impl(self, ctx, gloss):
        acts, emb_ctx, hf, stf, xf, tpre, tf, stt, tn, dlogits, nsp_ctx, dnsp, ids, tt, pos, mask, B, S = ctx
        h, nh, hn, pre = self.h, self.nh, self.hn, self.PRE_LN
        T = B * S
        P = self.P
        acc = self.accumulate_grads
        self._begin_backward()
        E = "bert.embeddings."
        scale = 1.0 / math.sqrt(hn)
        if gloss is not None:
            ops.scale_inplace(dlogits, gloss)
            if dnsp is not None:
                ops.scale_inplace(dnsp, gloss)
        wte = P(E + "word_embeddings.weight")
        dtn = ops.gemm(L.GEMM_NN, dlogits, wte.data)
        ops.gemm(L.GEMM_TN, dlogits, tn, out=wte.main_grad, accumulate=acc)    # tied decoder: written first
        ops.colsum(dlogits, P("cls.predictions.bias").main_grad, accumulate=acc)
        del dlogits
        lnw, lnb = P("cls.predictions.transform.LayerNorm.weight"), P("cls.predictions.transform.LayerNorm.bias")
        dtf = ops.layernorm_bwd(dtn, tf, lnw.data, stt, lnw.main_grad, lnb.main_grad, accumulate=acc)
        dtpre = ops.act_bwd(self.act, dtf, tpre)
        tw, tb = P("cls.predictions.transform.dense.weight"), P("cls.predictions.transform.dense.bias")
        dhf = ops.gemm(L.GEMM_NN, dtpre, tw.data)
        ops.gemm(L.GEMM_TN, dtpre, hf, out=tw.main_grad, accumulate=acc)
        ops.colsum(dtpre, tb.main_grad, accumulate=acc)
        if pre and dnsp is not None:
            first, ppre, pooled = nsp_ctx
            sw, sb = P("cls.seq_relationship.weight"), P("cls.seq_relationship.bias")
            dpooled = ops.gemm(L.GEMM_NN, dnsp, self._nsp_w)                     # [B, h]
            dw8 = ops.gemm(L.GEMM_TN, dnsp, pooled, out_dtype=torch.float32)     # [8, h]
            db8 = torch.zeros(8, dtype=torch.float32, device=dnsp.device)
            ops.colsum(dnsp, db8)
            if acc:
                sw.main_grad.add_(dw8[:2].to(sw.main_grad.dtype)); sb.main_grad.add_(db8[:2].to(sb.main_grad.dtype))
            else:
                sw.main_grad.copy_(dw8[:2]); sb.main_grad.copy_(db8[:2])
            dppre = ops.act_bwd(L.ACT_TANH, dpooled, ppre)
            pw, pb = P("bert.pooler.dense.weight"), P("bert.pooler.dense.bias")
            ops.gemm(L.GEMM_TN, dppre, first, out=pw.main_grad, accumulate=acc)
            ops.colsum(dppre, pb.main_grad, accumulate=acc)
            ops.gemm(L.GEMM_NN, dppre, pw.data, out=dhf.view(B, S, h)[:, 0, :], accumulate=True)   # += into token 0 rows
        elif pre:
            for n in ("cls.seq_relationship.weight", "cls.seq_relationship.bias", "bert.pooler.dense.weight",
                      "bert.pooler.dense.bias"):
                if not acc:
                    P(n).main_grad.zero_()
        self._done("head")
        if pre:
            ew, eb = P("bert.encoder.ln.weight"), P("bert.encoder.ln.bias")
            dx = ops.layernorm_bwd(dhf, xf, ew.data, stf, ew.main_grad, eb.main_grad, accumulate=acc)
        else:
            dx = dhf
        for i in reversed(range(self.nl)):
            p = f"bert.encoder.layer.{i}."
            w1, b1 = P(p + "intermediate.dense.weight"), P(p + "intermediate.dense.bias")
            w2, b2 = P(p + "output.dense.weight"), P(p + "output.dense.bias")
            wo, bo = P(p + "attention.output.dense.weight"), P(p + "attention.output.dense.bias")
            if pre:
                x, st1, h1, qkv, o, lse, x1, st2, h2, prea, f = acts[i]
                dm, dres_in = dx, dx                       # x_next = x1 + m
            else:
                x, qkv, o, lse, x1, st2, h2, prea, f, s2, st3 = acts[i]
                lw, lb = P(p + "output.LayerNorm.weight"), P(p + "output.LayerNorm.bias")
                dm = ops.layernorm_bwd(dx, s2, lw.data, st3, lw.main_grad, lb.main_grad, accumulate=acc)  # d(h2 + m)
                dres_in = None
            acts[i] = None
            df = ops.gemm(L.GEMM_NN, dm, w2.data)
            ops.gemm(L.GEMM_TN, dm, f, out=w2.main_grad, accumulate=acc)
            ops.colsum(dm, b2.main_grad, accumulate=acc)
            dprea = ops.act_bwd_bias(self.act, df, prea, b1.main_grad, accumulate=acc)   # dGELU + intermediate.dense bias grad
            ops.gemm(L.GEMM_TN, dprea, h2, out=w1.main_grad, accumulate=acc)
            if pre:
                dh2 = ops.gemm(L.GEMM_NN, dprea, w1.data)
                lw, lb = P(p + "ln.weight"), P(p + "ln.bias")
                dx1 = ops.layernorm_bwd(dh2, x1, lw.data, st2, lw.main_grad, lb.main_grad, accumulate=acc, dres=dres_in)
                da = dx1
            else:
                ops.gemm(L.GEMM_NN, dprea, w1.data, out=dm, accumulate=True)      # dh2 = d(h2+m) + dgrad(fc1)
                lw, lb = P(p + "attention.output.LayerNorm.weight"), P(p + "attention.output.LayerNorm.bias")
                da = ops.layernorm_bwd(dm, x1, lw.data, st2, lw.main_grad, lb.main_grad, accumulate=acc)  # d(x + a)
            do = ops.gemm(L.GEMM_NN, da, wo.data)
            ops.gemm(L.GEMM_TN, da, o.view(T, h), out=wo.main_grad, accumulate=acc)
            ops.colsum(da, bo.main_grad, accumulate=acc)
            dqkv = torch.empty_like(qkv)
            q5, d5 = qkv.view(B, S, 3, nh, hn), dqkv.view(B, S, 3, nh, hn)
            ops.sdpa_bwd(q5[:, :, 0], q5[:, :, 1], q5[:, :, 2], o, do.view(B, S, nh, hn), lse, scale, False,
                         d5[:, :, 0], d5[:, :, 1], d5[:, :, 2], kv_mask=mask)
            attn_in = h1 if pre else x
            ops.gemm(L.GEMM_TN, dqkv, attn_in, out=self._dwqkv[i], accumulate=acc)
            ops.colsum(dqkv, self._dbqkv[i], accumulate=acc)
            if pre:
                dh1 = ops.gemm(L.GEMM_NN, dqkv, self._wqkv[i])
                lw, lb = P(p + "attention.ln.weight"), P(p + "attention.ln.bias")
                dx = ops.layernorm_bwd(dh1, x, lw.data, st1, lw.main_grad, lb.main_grad, accumulate=acc, dres=dx1)
            else:
                ops.gemm(L.GEMM_NN, dqkv, self._wqkv[i], out=da, accumulate=True)  # dx_in = d(x+a) + dgrad(qkv)
                dx = da
            self._done(f"layer{i}")
        if not pre:
            emb, st_e = emb_ctx
            lw, lb = P(E + "LayerNorm.weight"), P(E + "LayerNorm.bias")
            dx = ops.layernorm_bwd(dx, emb, lw.data, st_e, lw.main_grad, lb.main_grad, accumulate=acc)
        ops.embedding_bwd(ids, dx, wte.main_grad)           # adds onto the tied decoder's weight gradient
        wpe, wtt = P(E + "position_embeddings.weight"), P(E + "token_type_embeddings.weight")
        if pos is None:
            ops.colsum(dx.view(B, S * h), wpe.main_grad[:S].reshape(-1), accumulate=acc)
            if not acc and S < self.npos:
                wpe.main_grad[S:].zero_()
        else:
            if not acc:
                wpe.main_grad.zero_()
            ops.embedding_bwd(pos, dx, wpe.main_grad)
        # token types: dT = onehot(tt)^T dx as a (tiny-M) GEMM; all-zero types reduce to a column sum
        if tt is None:
            ops.colsum(dx, wtt.main_grad[0], accumulate=acc)
            if not acc:
                wtt.main_grad[1:].zero_()
        else:
            onehot = torch.zeros(T, 8, dtype=torch.bfloat16, device=dx.device)
            onehot.scatter_(1, tt.view(-1, 1), 1.0)
            d8 = ops.gemm(L.GEMM_TN, onehot, dx, out_dtype=torch.float32)
            if acc:
                wtt.main_grad.add_(d8[:self.ntype].to(wtt.main_grad.dtype))
            else:
                wtt.main_grad.copy_(d8[:self.ntype])
        self._done("emb")
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


class BertForMaskedLM(_BertFamily):
    PRE_LN = False


class MegatronBertForPreTraining(_BertFamily):
    PRE_LN = True


class _BertStep(torch.autograd.Function):
    @staticmethod
    def forward(ctx, model, ids, tt, pos, mask, lab, nsl, B, S, want_logits, _anchor):
        loss, logits, nsp, saved = model._forward_impl(ids, tt, pos, mask, lab, nsl, B, S, save=True,
                                                       want_logits=want_logits)
        ctx.model, ctx.saved = model, saved
        nd = [t for t in (logits, nsp) if t is not None]
        ctx.mark_non_differentiable(*nd)
        return loss, logits, nsp

    @staticmethod
    def backward(ctx, gloss, _gl, _gn):
        model, saved = ctx.model, ctx.saved
        ctx.saved = None
        model._backward_impl(saved, gloss)
        return (None,) * 11
