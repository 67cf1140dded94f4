"""Test double for fsb200.engine's `kernels` argument: the same four entry points implemented with torch CPU ops
(AdamW in torch.optim.AdamW's operation order). Lives under tests/ — it is a CHECKER for the host-side sharding /
bucketing logic over gloo, never a product fallback."""
import math

import torch


def accumulate(acc32, x16, scale=1.0, overwrite=False):
    if overwrite:
        acc32.copy_(x16.float() * scale)
    else:
        acc32.add_(x16.float() * scale)


def sumsq(x, out, accumulate=False):
    s = x.float().pow(2).sum()
    out.copy_(out + s if accumulate else s)


def clip_coef(sumsq_t, max_norm, coef_out, norm_out=None):
    nrm = sumsq_t.sqrt()
    if norm_out is not None:
        norm_out.copy_(nrm)
    c = max_norm / (nrm + 1e-6)
    coef_out.copy_(torch.clamp(c, max=1.0) if max_norm > 0 else torch.ones_like(c))


def adamw_flat(master, m, v, grad, param16, lr, beta1, beta2, eps, weight_decay, step, grad_scale=None, hyper=None):
    g = grad.float() * (grad_scale if grad_scale is not None else 1.0)
    master.mul_(1 - lr * weight_decay)
    m.mul_(beta1).add_(g, alpha=1 - beta1)
    v.mul_(beta2).addcmul_(g, g, value=1 - beta2)
    bc1 = 1 - beta1 ** step
    bc2s = math.sqrt(1 - beta2 ** step)
    master.addcdiv_(m, v.sqrt() / bc2s + eps, value=-lr / bc1)
    if param16 is not None:
        param16.copy_(master.to(param16.dtype))
