"""Dev tool: time the attention kernels at a bench-like shape (CUDA-graph replay of 10 launches, so host launch overhead
is excluded); also usable as the ncu target."""
import math, os, sys, torch
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "fengshen-lm_b200"))
from fsb200 import ops
B, S, H, D = (int(x) for x in (sys.argv[1:5] if len(sys.argv) > 4 else (8, 1024, 12, 64)))
qkv = torch.randn(B, S, 3, H, D, device="cuda", dtype=torch.bfloat16)
q, k, v = qkv[:, :, 0], qkv[:, :, 1], qkv[:, :, 2]
dout = torch.randn(B, S, H, D, device="cuda", dtype=torch.bfloat16)
dqkv = torch.empty_like(qkv)
sc = 1 / math.sqrt(D)
for _ in range(3):
    out, lse = ops.sdpa_fwd(q, k, v, sc, True)
    ops.sdpa_bwd(q, k, v, out, dout, lse, sc, True, dqkv[:, :, 0], dqkv[:, :, 1], dqkv[:, :, 2])
torch.cuda.synchronize()
N = 10


def timed(fn):
    g = torch.cuda.CUDAGraph()
    side = torch.cuda.Stream()
    with torch.cuda.stream(side):
        with torch.cuda.graph(g, stream=side):
            for _ in range(N):
                fn()
    g.replay()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record(); g.replay(); e.record(); torch.cuda.synchronize()
    return s.elapsed_time(e) * 1e-3 / N


fl = 4 * S * S * D * H * B / 2
t = timed(lambda: ops.sdpa_fwd(q, k, v, sc, True))
print(f"B{B} S{S} H{H} D{D} fwd {t*1e6:.1f} us  {fl/t/1e12:.1f} TF/s")
t = timed(lambda: ops.sdpa_bwd(q, k, v, out, dout, lse, sc, True, dqkv[:, :, 0], dqkv[:, :, 1], dqkv[:, :, 2]))
print(f"B{B} S{S} H{H} D{D} bwd {t*1e6:.1f} us  {2.5*fl/t/1e12:.1f} TF/s (5-GEMM count)")
