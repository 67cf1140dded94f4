"""Dev tool: time fsb_gemm_bf16 against torch.matmul (cuBLAS) on the hot-path shapes. Not part of the product path."""
import os
import sys

import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "fengshen-lm_b200"))
from fsb200 import lib as L, ops  # noqa: E402


def timeit(fn, iters=20, warmup=3):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(iters):
        fn()
    e.record()
    torch.cuda.synchronize()
    return s.elapsed_time(e) / iters


def main():
    dev = "cuda"
    shapes = [  # (name, layout, M, N, K)
        ("llama qkv fwd", L.GEMM_NT, 8192, 15360, 5120),
        ("llama w13 fwd", L.GEMM_NT, 8192, 27648, 5120),
        ("llama w2 fwd", L.GEMM_NT, 8192, 5120, 13824),
        ("llama w2 dgrad", L.GEMM_NN, 8192, 13824, 5120),
        ("llama w13 wgrad", L.GEMM_TN, 27648, 5120, 8192),
        ("llama head fwd", L.GEMM_NT, 8192, 39424, 5120),
        ("gpt2 fc fwd", L.GEMM_NT, 32768, 3072, 768),
        ("gpt2 fc wgrad", L.GEMM_TN, 3072, 768, 32768),
        ("gpt2 head fwd", L.GEMM_NT, 32768, 50264, 768),
        ("square 8192", L.GEMM_NT, 8192, 8192, 8192),
    ]
    if len(sys.argv) > 1 and sys.argv[1] == "gpt2":   # every GEMM of a Wenzhong-GPT2-110M step (tokens = 32 x 1024)
        T, h, V = 32768, 768, 50264
        shapes = []
        for nm, i, o in [("c_attn", h, 3 * h), ("c_proj", h, h), ("c_fc", h, 4 * h), ("mlp_proj", 4 * h, h)]:
            shapes += [(nm + " fwd", L.GEMM_NN, T, o, i), (nm + " dgrad", L.GEMM_NT, T, i, o), (nm + " wgrad", L.GEMM_TN, i, o, T)]
        shapes += [("head fwd", L.GEMM_NT, T, V, h), ("head dgrad", L.GEMM_NN, T, h, V), ("head wgrad", L.GEMM_TN, V, h, T)]
    if len(sys.argv) > 1 and sys.argv[1] == "epi":    # epilogue variants of the GPT2 c_fc forward GEMM
        T, h = 32768, 768
        a = torch.randn(T, h, device=dev, dtype=torch.bfloat16); w = torch.randn(h, 4 * h, device=dev, dtype=torch.bfloat16) * 0.05
        bias = torch.randn(4 * h, device=dev, dtype=torch.bfloat16)
        out = torch.empty(T, 4 * h, device=dev, dtype=torch.bfloat16); aux = torch.empty_like(out)
        for nm, kw in [("plain", {}), ("bias", dict(bias=bias)), ("bias+aux", dict(bias=bias, aux=aux)),
                       ("bias+gelu_tanh", dict(bias=bias, epilogue=L.EPI_GELU_TANH)),
                       ("bias+gelu_tanh+aux", dict(bias=bias, epilogue=L.EPI_GELU_TANH, aux=aux)),
                       ("bias+gelu_erf+aux", dict(bias=bias, epilogue=L.EPI_GELU_ERF, aux=aux))]:
            t = timeit(lambda: ops.gemm(L.GEMM_NN, a, w, out=out, **kw))
            print(f"c_fc fwd {nm:22s} {t:8.3f} ms {2.0 * T * h * 4 * h / t / 1e9:8.1f} TF", flush=True)
        return
    for name, layout, M, N, K in shapes:
        if layout == L.GEMM_NT:
            a = torch.randn(M, K, device=dev, dtype=torch.bfloat16); b = torch.randn(N, K, device=dev, dtype=torch.bfloat16)
            ref = lambda: torch.matmul(a, b.t())
        elif layout == L.GEMM_NN:
            a = torch.randn(M, K, device=dev, dtype=torch.bfloat16); b = torch.randn(K, N, device=dev, dtype=torch.bfloat16)
            ref = lambda: torch.matmul(a, b)
        else:
            a = torch.randn(K, M, device=dev, dtype=torch.bfloat16); b = torch.randn(K, N, device=dev, dtype=torch.bfloat16)
            ref = lambda: torch.matmul(a.t(), b)
        out = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
        mine = lambda: ops.gemm(layout, a, b, out=out)
        t_ref = timeit(ref)
        t_mine = timeit(mine)
        err = (out.float() - ref().float()).abs().max().item()
        fl = 2.0 * M * N * K
        print(f"{name:18s} M={M:6d} N={N:6d} K={K:6d}  fsb {t_mine:8.3f} ms {fl / t_mine / 1e9:8.1f} TF | "
              f"cublas {t_ref:8.3f} ms {fl / t_ref / 1e9:8.1f} TF | maxerr {err:.3g}", flush=True)
        del a, b, out


if __name__ == "__main__":
    main()
