"""Tensor-level wrappers over the C ABI: torch is used only for device memory and the current stream.

Each function validates what the C side cannot know (dtype, device, contiguity) and forwards raw pointers.
No function here has a PyTorch fallback path.
"""
import torch

from . import lib as L

_bf16 = torch.bfloat16
_ws_cache = {}


def _stream():
    return torch.cuda.current_stream().cuda_stream


def _p(t):
    return None if t is None else t.data_ptr()


def _chk(t, dtype=None, name="tensor"):
    if not t.is_cuda:
        raise RuntimeError(f"fsb200: {name} must be a CUDA tensor (the hot path has no CPU fallback)")
    if dtype is not None and t.dtype != dtype:
        raise RuntimeError(f"fsb200: {name} must be {dtype}, got {t.dtype}")


def _rows2d(t, name):
    """View as [rows, cols] with unit inner stride; returns (rows, cols, ld)."""
    if t.dim() != 2 or t.stride(1) != 1:
        raise RuntimeError(f"fsb200: {name} must be 2-D with unit inner stride, got shape {tuple(t.shape)} "
                           f"strides {t.stride()}")
    return t.shape[0], t.shape[1], t.stride(0)


def workspace(nbytes, device, tag="default"):
    """Grow-only scratch buffer per (device, tag); caller-owned from the library's point of view."""
    key = (device, tag)
    buf = _ws_cache.get(key)
    if buf is None or buf.numel() < nbytes:
        buf = torch.empty(max(int(nbytes), 1 << 20), dtype=torch.uint8, device=device)
        _ws_cache[key] = buf
    return buf


# ------------------------------------------------------------------------------------------------------ GEMM
def gemm(layout, a, b, out=None, out_dtype=_bf16, bias=None, epilogue=L.EPI_NONE, accumulate=False, aux=None):
    """layout NT: a[M,K] b[N,K]; NN: a[M,K] b[K,N]; TN: a[K,M] b[K,N]. Returns out[M,N]."""
    _chk(a, _bf16, "a"); _chk(b, _bf16, "b")
    ar, ac, lda = _rows2d(a, "a")
    br, bc, ldb = _rows2d(b, "b")
    if layout == L.GEMM_NT:
        M, K, N = ar, ac, br
        if bc != K: raise RuntimeError(f"fsb200 gemm NT: K mismatch {ac} vs {bc}")
    elif layout == L.GEMM_NN:
        M, K, N = ar, ac, bc
        if br != K: raise RuntimeError(f"fsb200 gemm NN: K mismatch {ac} vs {br}")
    else:
        K, M, N = ar, ac, bc
        if br != K: raise RuntimeError(f"fsb200 gemm TN: K mismatch {ar} vs {br}")
    if out is None:
        if accumulate: raise RuntimeError("fsb200 gemm: accumulate needs an existing `out`")
        out = torch.empty((M, N), dtype=out_dtype, device=a.device)
    _chk(out, None, "out")
    if out.dtype not in (_bf16, torch.float32): raise RuntimeError("fsb200 gemm: out must be bf16 or fp32")
    orr, occ, ldd = _rows2d(out, "out")
    if (orr, occ) != (M, N): raise RuntimeError(f"fsb200 gemm: out shape {tuple(out.shape)} != ({M},{N})")
    bias_dt = L.BF16
    if bias is not None:
        _chk(bias, None, "bias")
        if bias.numel() != N or not bias.is_contiguous(): raise RuntimeError("fsb200 gemm: bias must be contiguous [N]")
        bias_dt = L.F32 if bias.dtype == torch.float32 else L.BF16
    ldaux = 0
    if aux is not None:
        _chk(aux, _bf16, "aux")
        _, _, ldaux = _rows2d(aux, "aux")
    ws, ws_bytes = None, 0
    if layout == L.GEMM_TN and bias is None and aux is None and epilogue == L.EPI_NONE:
        key = (M, N, K)
        ws_bytes = _splitk_bytes.get(key)
        if ws_bytes is None:
            ws_bytes = _splitk_bytes[key] = int(L.load().fsb_gemm_workspace_bytes(layout, M, N, K))
        if ws_bytes:
            ws = workspace(ws_bytes, a.device, "gemm_splitk")   # one stream issues the step's GEMMs: a shared scratch is safe
    prof = _profiler
    if prof is not None:
        ev0 = torch.cuda.Event(enable_timing=True); ev0.record()
    L.call("fsb_gemm_bf16", layout, M, N, K, _p(a), lda, _p(b), ldb, _p(out), ldd,
           L.F32 if out.dtype == torch.float32 else L.BF16, _p(bias), bias_dt, epilogue, int(bool(accumulate)),
           _p(aux), ldaux, 1, 0, 0, 0, 0, _p(ws), ws_bytes, _stream(),
           tag=(f"{('NT', 'NN', 'TN')[layout]} {M}x{N}x{K} epi{epilogue} acc{int(bool(accumulate))} "
                f"{'f32' if out.dtype == torch.float32 else 'bf16'}{' bias' if bias is not None else ''}"
                f"{' aux' if aux is not None else ''}") if L.call_profiler is not None else None)
    if prof is not None:
        ev1 = torch.cuda.Event(enable_timing=True); ev1.record()
        prof.add("gemm_bf16_kernel", ev0, ev1, 2.0 * M * N * K)
    return out


_splitk_bytes = {}


def set_reserved_sms(n):
    """Leave n SMs (2n for CTA-pair kernels) of every persistent GEMM grid to overlapping communication kernels."""
    L.call("fsb_set_reserved_sms", int(n))


class KernelProfiler:
    """CUDA-event timing of individual launches on the launching stream (bench.py's roofline block)."""

    def __init__(self):
        self.rec = {}

    def add(self, name, ev0, ev1, work):
        self.rec.setdefault(name, []).append((ev0, ev1, work))

    def summary(self):
        out = {}
        for name, items in self.rec.items():
            ms = sum(a.elapsed_time(b) for a, b, _ in items)
            out[name] = {"launches": len(items), "ms": ms, "work": sum(w for _, _, w in items)}
        return out


_profiler = None


def set_profiler(p):
    global _profiler
    _profiler = p


# ------------------------------------------------------------------------------------------------------ norms
def rmsnorm_fwd(x, scale, eps, residual=None):
    """x [rows, cols] bf16. Returns (y, rstd, x_sum) where x_sum = x + residual (or x itself when residual is None)."""
    _chk(x, _bf16, "x"); _chk(scale, _bf16, "scale")
    rows, cols = x.shape
    y = torch.empty_like(x)
    rstd = torch.empty(rows, dtype=torch.float32, device=x.device)
    xs = torch.empty_like(x) if residual is not None else None
    L.call("fsb_rmsnorm_fwd", _p(x), _p(residual), _p(scale), _p(y), _p(xs), _p(rstd), rows, cols, float(eps), _stream())
    return y, rstd, (xs if residual is not None else x)


def rmsnorm_bwd(dy, x, scale, rstd, dscale_out, accumulate=False, dres=None):
    rows, cols = x.shape
    dx = torch.empty_like(x)
    nbytes = L.load().fsb_norm_bwd_workspace_bytes(rows, cols, 0)
    ws = workspace(nbytes, x.device, "norm")
    L.call("fsb_rmsnorm_bwd", _p(dy), _p(x), _p(scale), _p(rstd), _p(dres), _p(dx), _p(dscale_out),
           L.F32 if dscale_out.dtype == torch.float32 else L.BF16, int(bool(accumulate)), _p(ws), ws.numel(), rows, cols,
           _stream())
    return dx


def layernorm_fwd(x, gamma, beta, eps, residual=None):
    _chk(x, _bf16, "x"); _chk(gamma, _bf16, "gamma"); _chk(beta, _bf16, "beta")
    rows, cols = x.shape
    y = torch.empty_like(x)
    stats = torch.empty(rows, 2, dtype=torch.float32, device=x.device)
    xs = torch.empty_like(x) if residual is not None else None
    L.call("fsb_layernorm_fwd", _p(x), _p(residual), _p(gamma), _p(beta), _p(y), _p(xs), _p(stats), rows, cols,
           float(eps), _stream())
    return y, stats, (xs if residual is not None else x)


def layernorm_bwd(dy, x, gamma, stats, dgamma_out, dbeta_out, accumulate=False, dres=None):
    rows, cols = x.shape
    dx = torch.empty_like(x)
    nbytes = L.load().fsb_norm_bwd_workspace_bytes(rows, cols, 1)
    ws = workspace(nbytes, x.device, "norm")
    L.call("fsb_layernorm_bwd", _p(dy), _p(x), _p(gamma), _p(stats), _p(dres), _p(dx), _p(dgamma_out), _p(dbeta_out),
           L.F32 if dgamma_out.dtype == torch.float32 else L.BF16, int(bool(accumulate)), _p(ws), ws.numel(), rows, cols,
           _stream())
    return dx


# ------------------------------------------------------------------------------------------------------ pointwise
def rope_inplace(x, cos, sin, positions, nheads, head_dim, row_stride, head_stride, backward=False, offset=0):
    """Rotate `nheads` heads per row in place. x is the flat packed buffer; `offset` (elements) selects q or k."""
    _chk(x, _bf16, "x")
    rows = positions.numel()
    L.call("fsb_rope_inplace", x.data_ptr() + 2 * offset, _p(cos), _p(sin), _p(positions), rows, nheads, head_dim,
           row_stride, head_stride, cos.shape[0], int(bool(backward)), _stream())


def glu_fwd(act, gate, up):
    rows, cols, ldg = _rows2d(gate, "gate")
    _, _, ldu = _rows2d(up, "up")
    out = torch.empty((rows, cols), dtype=_bf16, device=gate.device)
    L.call("fsb_glu_fwd", act, _p(gate), _p(up), _p(out), rows, cols, ldg, ldu, cols, _stream())
    return out


def glu_bwd(act, dout, gate, up, dgate, dup):
    rows, cols, ldg = _rows2d(gate, "gate")
    _, _, ldu = _rows2d(up, "up")
    _, _, ldo = _rows2d(dout, "dout")
    _, _, ldg2 = _rows2d(dgate, "dgate")
    _, _, ldu2 = _rows2d(dup, "dup")
    L.call("fsb_glu_bwd", act, _p(dout), _p(gate), _p(up), _p(dgate), _p(dup), rows, cols, ldo, ldg, ldu, ldg2, ldu2,
           _stream())


def act_fwd(act, x):
    y = torch.empty_like(x)
    L.call("fsb_act_fwd", act, _p(x), _p(y), x.numel(), _stream())
    return y


def act_bwd(act, dy, x):
    dx = torch.empty_like(x)
    L.call("fsb_act_bwd", act, _p(dy), _p(x), _p(dx), x.numel(), _stream())
    return dx


def act_bwd_bias(act, dy, x, dbias, accumulate=False):
    """dx = dy * act'(x) and dbias[c] (+)= sum_r dx[r, c] in one pass (x, dy contiguous [rows, cols])."""
    _chk(x, _bf16, "x"); _chk(dy, _bf16, "dy")
    if not (x.is_contiguous() and dy.is_contiguous()) or x.dim() != 2 or dy.shape != x.shape:
        raise RuntimeError("fsb200 act_bwd_bias: x and dy must be contiguous [rows, cols] of the same shape")
    rows, cols = x.shape
    dx = torch.empty_like(x)
    nbytes = L.load().fsb_act_bwd_bias_workspace_bytes(rows, cols)
    ws = workspace(nbytes, x.device, "act_bwd_bias")
    L.call("fsb_act_bwd_bias", act, _p(dy), _p(x), _p(dx), rows, cols, _p(dbias),
           L.F32 if dbias.dtype == torch.float32 else L.BF16, int(bool(accumulate)), _p(ws), ws.numel(), _stream())
    return dx


def add(a, b, out=None):
    if out is None: out = torch.empty_like(a)
    L.call("fsb_add", _p(a), _p(b), _p(out), a.numel(), _stream())
    return out


def accumulate(acc32, x16, scale=1.0, overwrite=False):
    """acc32 (fp32) = (0 if overwrite else acc32) + scale * x16 (bf16)."""
    L.call("fsb_accumulate", _p(acc32), _p(x16), acc32.numel(), float(scale), int(bool(overwrite)), _stream())


def scale_inplace(x16, scale_dev):
    """x16 (bf16, contiguous) *= scale_dev (0-d fp32 CUDA tensor); free when the scalar is 1."""
    if scale_dev.dtype != torch.float32 or not scale_dev.is_cuda:
        scale_dev = scale_dev.to(device=x16.device, dtype=torch.float32)
    L.call("fsb_scale_inplace", _p(x16), x16.numel(), _p(scale_dev), _stream())


def colsum(x, out, accumulate=False):
    """out[c] (+)= sum_r x[r, c]; x bf16 [rows, cols] (unit inner stride); out bf16 or fp32 [cols]."""
    rows, cols, ld = _rows2d(x, "x")
    nbytes = L.load().fsb_colsum_workspace_bytes(rows, cols)
    ws = workspace(nbytes, x.device, "colsum")
    L.call("fsb_colsum", _p(x), rows, cols, ld, _p(out), L.F32 if out.dtype == torch.float32 else L.BF16,
           int(bool(accumulate)), _p(ws), ws.numel(), _stream())


def embedding_fwd(ids, W, pos=None, P=None, token_type=None, T=None, seq_len=1):
    rows = ids.numel()
    cols = W.shape[1]
    out = torch.empty((rows, cols), dtype=_bf16, device=W.device)
    L.call("fsb_embedding_fwd", _p(ids), _p(pos), _p(token_type), _p(W), _p(P), _p(T), _p(out), rows, cols, seq_len,
           _stream())
    return out


def embedding_bwd(ids, dout, dW, idx_mod=0):
    """dW[ids[t]] += dout[t]. With ids: deterministic — the ids are sorted (torch.sort: integer index plumbing) and every
    distinct row is summed in fp32 in a fixed order, one bf16 rounding. ids=None: row t % idx_mod (bf16 atomics)."""
    rows, cols = dout.shape
    if ids is None:
        L.call("fsb_embedding_bwd", None, _p(dout), _p(dW), rows, cols, idx_mod, _stream())
        return
    ids_sorted, order = torch.sort(ids.view(-1), stable=True)
    L.call("fsb_embedding_bwd_sorted", _p(ids_sorted), _p(order), _p(dout), _p(dW), rows, cols, _stream())


def cast_f32_to_bf16(x32, out=None):
    if out is None:
        out = torch.empty(x32.shape, dtype=_bf16, device=x32.device)
    L.call("fsb_cast_f32_to_bf16", _p(x32), _p(out), x32.numel(), _stream())
    return out


# ------------------------------------------------------------------------------------------------------ loss / optim
def softmax_xent(logits, labels, seq_len, shift=1, ignore_index=-100, grad_scale=1.0, dlogits="inplace"):
    """logits [rows, V] bf16 (rows = b*seq_len), labels int64 [rows]. Returns (loss scalar tensor, dlogits, n_valid)."""
    _chk(logits, _bf16, "logits")
    rows, V, ld = _rows2d(logits, "logits")
    dev = logits.device
    row_loss = torch.empty(rows, dtype=torch.float32, device=dev)
    loss = torch.empty((), dtype=torch.float32, device=dev)
    n_valid = torch.empty((), dtype=torch.int32, device=dev)
    if isinstance(dlogits, str):
        dl = logits if dlogits == "inplace" else None
    else:
        dl = dlogits
    L.call("fsb_softmax_xent_fwd_bwd", _p(logits), _p(labels), _p(dl), _p(row_loss), _p(loss), _p(n_valid), rows, V, ld,
           seq_len, shift, ignore_index, float(grad_scale), _stream())
    return loss, dl, n_valid


def adamw_flat(master, m, v, grad, param16, lr, beta1, beta2, eps, weight_decay, step, grad_scale=None, hyper=None):
    """hyper: optional fp32 CUDA tensor [lr, 1 - beta1^t, sqrt(1 - beta2^t)] read by the kernel instead of lr / step."""
    L.call("fsb_adamw_flat", _p(master), _p(m), _p(v), _p(grad), L.F32 if grad.dtype == torch.float32 else L.BF16,
           _p(param16), master.numel(), float(lr), float(beta1), float(beta2), float(eps), float(weight_decay), int(step),
           _p(grad_scale), _p(hyper), _stream())


def sumsq(x, out, accumulate=False):
    nbytes = L.load().fsb_sumsq_workspace_bytes()
    ws = workspace(nbytes, x.device, "sumsq")
    L.call("fsb_sumsq", _p(x), L.F32 if x.dtype == torch.float32 else L.BF16, x.numel(), _p(out), int(bool(accumulate)),
           _p(ws), ws.numel(), _stream())


def clip_coef(sumsq_t, max_norm, coef_out, norm_out=None):
    L.call("fsb_clip_coef", _p(sumsq_t), float(max_norm), _p(coef_out), _p(norm_out), _stream())


# ------------------------------------------------------------------------------------------------------ attention
def _bshd(t, name):
    """[batch, seq, heads, dim] view with unit inner stride and batch stride == seq * row stride."""
    if t.dim() != 4 or t.stride(3) != 1:
        raise RuntimeError(f"fsb200: {name} must be [batch, seq, heads, dim] with unit inner stride")
    B, S, H, D = t.shape
    if B > 1 and t.stride(0) != S * t.stride(1):
        raise RuntimeError(f"fsb200: {name} batch stride {t.stride(0)} != seq*row_stride {S * t.stride(1)}")
    return B, S, H, D, t.stride(1), t.stride(2)


def _chk_rel(rel, H, Sq, Skv, name):
    _chk(rel, torch.float32, name)
    if tuple(rel.shape) != (H, Sq + Skv - 1) or not rel.is_contiguous():
        raise RuntimeError(f"fsb200: {name} must be contiguous fp32 [heads, seq_q + seq_kv - 1] = [{H}, {Sq + Skv - 1}], "
                           f"got {tuple(rel.shape)}")


def sdpa_fwd(q, k, v, scale, causal, kv_mask=None, out=None, rel_bias=None):
    """q,k,v: strided [B,S,H,D] bf16 views (e.g. slices of the packed QKV projection). Returns (out [B,Sq,H,D], lse).
    rel_bias: optional fp32 [H, Sq + Skv - 1] additive bias over the offset k - q (T5 relative-position bias)."""
    _chk(q, _bf16, "q"); _chk(k, _bf16, "k"); _chk(v, _bf16, "v")
    B, Sq, H, D, q_rs, q_hs = _bshd(q, "q")
    _, Skv, _, _, k_rs, k_hs = _bshd(k, "k")
    _, _, _, _, v_rs, v_hs = _bshd(v, "v")
    if out is None:
        out = torch.empty((B, Sq, H, D), dtype=_bf16, device=q.device)
    _, _, _, _, o_rs, o_hs = _bshd(out, "out")
    lse = torch.empty((B, H, Sq), dtype=torch.float32, device=q.device)
    if kv_mask is not None:
        _chk(kv_mask, torch.uint8, "kv_mask")
        if tuple(kv_mask.shape) != (B, Skv) or not kv_mask.is_contiguous():
            raise RuntimeError("fsb200: kv_mask must be contiguous uint8 [batch, seq_kv]")
    if rel_bias is not None:
        _chk_rel(rel_bias, H, Sq, Skv, "rel_bias")
    L.call("fsb_sdpa_fwd", _p(q), _p(k), _p(v), _p(out), _p(lse), B, Sq, Skv, H, D, q_rs, k_rs, v_rs, o_rs, q_hs, k_hs,
           v_hs, o_hs, float(scale), int(bool(causal)), _p(kv_mask), _p(rel_bias), _stream())
    return out, lse


def sdpa_bwd(q, k, v, out, dout, lse, scale, causal, dq, dk, dv, kv_mask=None, rel_bias=None, drel_bias=None):
    """All tensors strided [B,S,H,D] bf16 views; dq/dk/dv are written (e.g. slices of a packed dQKV buffer).
    rel_bias as in sdpa_fwd; drel_bias (fp32 [H, Sq + Skv - 1]) is accumulated into (+=), deterministically."""
    B, Sq, H, D, q_rs, q_hs = _bshd(q, "q")
    _, Skv, _, _, k_rs, k_hs = _bshd(k, "k")
    _, _, _, _, v_rs, v_hs = _bshd(v, "v")
    _, _, _, _, o_rs, o_hs = _bshd(out, "out")
    _, _, _, _, do_rs, do_hs = _bshd(dout, "dout")
    _, _, _, _, dq_rs, dq_hs = _bshd(dq, "dq")
    _, _, _, _, dk_rs, dk_hs = _bshd(dk, "dk")
    _, _, _, _, dv_rs, dv_hs = _bshd(dv, "dv")
    for t, n in ((dout, "dout"), (dq, "dq"), (dk, "dk"), (dv, "dv")):
        _chk(t, _bf16, n)
    delta = torch.empty((B, H, Sq), dtype=torch.float32, device=q.device)
    ws, ws_bytes = None, 0
    if rel_bias is not None:
        _chk_rel(rel_bias, H, Sq, Skv, "rel_bias")
    if drel_bias is not None:
        _chk_rel(drel_bias, H, Sq, Skv, "drel_bias")
        ws_bytes = int(L.load().fsb_sdpa_bwd_workspace_bytes(B, Sq, Skv, H))
        ws = workspace(ws_bytes, q.device, "sdpa_dbias")
    L.call("fsb_sdpa_bwd", _p(q), _p(k), _p(v), _p(out), _p(dout), _p(lse), _p(delta), _p(dq), _p(dk), _p(dv), B, Sq, Skv,
           H, D, q_rs, k_rs, v_rs, o_rs, do_rs, dq_rs, dk_rs, dv_rs, q_hs, k_hs, v_hs, o_hs, do_hs, dq_hs, dk_hs, dv_hs,
           float(scale), int(bool(causal)), _p(kv_mask), _p(rel_bias), _p(drel_bias), _p(ws), ws_bytes, _stream())
