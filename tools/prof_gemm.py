"""Dev tool: launch fsb_gemm_bf16 a few times at one shape (target for `ncu --set full -k regex:gemm_bf16`)."""
import os
import sys

import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "fengshen-lm_b200"))
from fsb200 import lib as L, ops  # noqa: E402

layout = {"NT": L.GEMM_NT, "NN": L.GEMM_NN, "TN": L.GEMM_TN}[sys.argv[1]]
M, N, K = (int(x) for x in sys.argv[2:5])
iters = int(sys.argv[5]) if len(sys.argv) > 5 else 6
dev = "cuda"
if layout == L.GEMM_NT:
    a, b = torch.randn(M, K, device=dev, dtype=torch.bfloat16), torch.randn(N, K, device=dev, dtype=torch.bfloat16)
elif layout == L.GEMM_NN:
    a, b = torch.randn(M, K, device=dev, dtype=torch.bfloat16), torch.randn(K, N, device=dev, dtype=torch.bfloat16)
else:
    a, b = torch.randn(K, M, device=dev, dtype=torch.bfloat16), torch.randn(K, N, device=dev, dtype=torch.bfloat16)
out = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
for _ in range(iters):
    ops.gemm(layout, a, b, out=out)
torch.cuda.synchronize()
print("done", M, N, K)
