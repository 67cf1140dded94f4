"""Dev tool: run the attention kernels once at a bench-like shape (for ncu captures)."""
import math, os, sys, torch
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "fengshen-lm_b200"))
from fsb200 import ops
B, S, H, D = (int(x) for x in (sys.argv[1:5] if len(sys.argv) > 4 else (8, 1024, 12, 64)))
qkv = torch.randn(B, S, 3, H, D, device="cuda", dtype=torch.bfloat16)
q, k, v = qkv[:, :, 0], qkv[:, :, 1], qkv[:, :, 2]
dout = torch.randn(B, S, H, D, device="cuda", dtype=torch.bfloat16)
dqkv = torch.empty_like(qkv)
for _ in range(3):
    out, lse = ops.sdpa_fwd(q, k, v, 1 / math.sqrt(D), True)
    ops.sdpa_bwd(q, k, v, out, dout, lse, 1 / math.sqrt(D), True, dqkv[:, :, 0], dqkv[:, :, 1], dqkv[:, :, 2])
torch.cuda.synchronize()
s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
s.record(); out, lse = ops.sdpa_fwd(q, k, v, 1 / math.sqrt(D), True); e.record(); torch.cuda.synchronize()
fl = 4 * S * S * D * H * B / 2
print(f"fwd {s.elapsed_time(e)*1e3:.1f} us  {fl/s.elapsed_time(e)/1e9:.1f} TF/s")
s.record(); ops.sdpa_bwd(q, k, v, out, dout, lse, 1 / math.sqrt(D), True, dqkv[:, :, 0], dqkv[:, :, 1], dqkv[:, :, 2]); e.record(); torch.cuda.synchronize()
print(f"bwd {s.elapsed_time(e)*1e3:.1f} us  {2.5*fl/s.elapsed_time(e)/1e9:.1f} TF/s (5-GEMM count)")
