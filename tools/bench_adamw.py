"""Dev tool: achieved HBM bandwidth of fsb_adamw_flat / fsb_accumulate / fsb_sumsq on bucket-sized and shard-sized inputs."""
import os
import sys

import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "fengshen-lm_b200"))
from fsb200 import ops  # noqa: E402


def timeit(fn, iters=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters * 1e-3


for n in (7_000_064, 12_850_176, 100_000_000, 400_000_000):
    for gdt in (torch.bfloat16, torch.float32):
        p, m, v = (torch.randn(n, device="cuda") for _ in range(3))
        g = torch.randn(n, device="cuda").to(gdt)
        p16 = torch.empty(n, dtype=torch.bfloat16, device="cuda")
        coef = torch.ones((), device="cuda")
        t = timeit(lambda: ops.adamw_flat(p, m, v, g, p16, 1e-4, 0.9, 0.95, 1e-8, 0.1, 3, coef))
        nbytes = n * (12 + g.element_size() + 12 + 2)
        print(f"adamw n={n:>11,d} grad={str(gdt):15s} {t*1e6:9.1f} us  {nbytes/t/1e12:6.2f} TB/s")
        del p, m, v, g, p16
