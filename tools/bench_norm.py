"""Dev tool: time the norm kernels (CUDA-graph replay of 10 launches) against their HBM roofline."""
import os, sys, torch
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "fengshen-lm_b200"))
from fsb200 import ops

def timed(fn, n=10):
    fn(); torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph(); side = torch.cuda.Stream()
    with torch.cuda.stream(side):
        with torch.cuda.graph(g, stream=side):
            for _ in range(n): fn()
    g.replay(); torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record(); g.replay(); e.record(); torch.cuda.synchronize()
    return s.elapsed_time(e) * 1e-3 / n

for kind, rows, cols in [("ln", 16384, 2048), ("rms", 8192, 5120), ("ln", 32768, 768), ("ln", 1024, 768)]:
    x = torch.randn(rows, cols, device="cuda", dtype=torch.bfloat16); r = torch.randn_like(x); dy = torch.randn_like(x)
    g = torch.ones(cols, device="cuda", dtype=torch.bfloat16); b = torch.zeros_like(g)
    dg = torch.zeros(cols, device="cuda", dtype=torch.float32); db = torch.zeros_like(dg)
    nbytes = rows * cols * 2
    if kind == "ln":
        y, st, xs = ops.layernorm_fwd(x, g, b, 1e-5, residual=r)
        tf = timed(lambda: ops.layernorm_fwd(x, g, b, 1e-5, residual=r))
        tb = timed(lambda: ops.layernorm_bwd(dy, xs, g, st, dg, db, dres=r))
    else:
        y, st, xs = ops.rmsnorm_fwd(x, g, 1e-6, residual=r)
        tf = timed(lambda: ops.rmsnorm_fwd(x, g, 1e-6, residual=r))
        tb = timed(lambda: ops.rmsnorm_bwd(dy, xs, g, st, dg, dres=r))
    print(f"{kind} {rows}x{cols}: fwd {tf*1e6:7.1f} us = {4*nbytes/tf/1e9:6.0f} GB/s | bwd {tb*1e6:7.1f} us = {4*nbytes/tb/1e9:6.0f} GB/s  (4 tensors of {nbytes/1e6:.0f} MB each way)")
