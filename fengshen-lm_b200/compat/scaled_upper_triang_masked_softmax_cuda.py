"""Drop-in for the reference's pybind module `scaled_upper_triang_masked_softmax_cuda`
(fengshen/models/megatron/fused_kernels/scaled_upper_triang_masked_softmax.cpp:62-70): forward(input [attn_batches, s, s],
scale) / backward(output_grads, softmax_results, scale) with the gradient written in place; executed by libfsb200.so."""
import torch

from fsb200 import lib as _L


def _stream():
    return torch.cuda.current_stream().cuda_stream


def _check(t, name):
    if not t.is_cuda or t.dtype != torch.bfloat16 or not t.is_contiguous():
        raise RuntimeError(f"{name} must be a contiguous CUDA bfloat16 tensor (got {t.dtype}, cuda={t.is_cuda})")
    if t.dim() != 3 or t.size(1) != t.size(2):
        raise RuntimeError("expected 3D tensor [attn_batches, seq_len, seq_len]")   # ..._cuda.cu:31-36


def forward(input, scale_factor):
    _check(input, "input")
    out = torch.empty_like(input)
    _L.call("fsb_scaled_upper_triang_masked_softmax_fwd", input.data_ptr(), out.data_ptr(), input.size(0), input.size(1),
            float(scale_factor), _stream())
    return out


def backward(output_grads, softmax_results, scale_factor):
    _check(output_grads, "output_grads"); _check(softmax_results, "softmax_results")
    _L.call("fsb_scaled_upper_triang_masked_softmax_bwd", output_grads.data_ptr(), softmax_results.data_ptr(),
            output_grads.size(0), output_grads.size(1), float(scale_factor), _stream())
    return output_grads
