"""Dev tool: build a traced copy of libfsb200 (-DFSB_ATTN_TRACE), run the forward kernel once, print per-step cycle deltas
of CTA (0,0,0): softmax warp 8 (slot 1) and the MMA warp."""
import ctypes
import math
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
csrc = os.path.join(ROOT, "fengshen-lm_b200", "csrc")
out = os.path.join(ROOT, "fengshen-lm_b200", "fsb200", "lib", "libfsb200_trace.so")   # build it on the CPU box: make -C csrc trace
if not os.path.exists(out):
    raise SystemExit("build the traced library first: make -C fengshen-lm_b200/csrc trace")
sys.path.insert(0, os.path.join(ROOT, "fengshen-lm_b200"))
from fsb200 import lib  # noqa: E402
lib.LIB_PATH = out
import numpy as np  # noqa: E402
import torch  # noqa: E402
from fsb200 import ops  # noqa: E402

B, S, H, D = (int(x) for x in sys.argv[1:5]) if len(sys.argv) > 4 else (8, 1024, 12, 64)
qkv = torch.randn(B, S, 3, H, D, device="cuda", dtype=torch.bfloat16)
q, k, v = qkv[:, :, 0], qkv[:, :, 1], qkv[:, :, 2]
for _ in range(2):
    o, lse = ops.sdpa_fwd(q, k, v, 1 / math.sqrt(D), True)
torch.cuda.synchronize()
buf = (ctypes.c_longlong * (2 * 64 * 8))()
fn = lib.load().fsb_debug_attn_fwd_trace
fn.argtypes = [ctypes.c_void_p]
fn.restype = ctypes.c_int
assert fn(buf) == 0
t = np.array(buf, dtype=np.int64).reshape(2, 64, 8)
n = min(S // 64, 63)
base = t[0, 0, 0]
c = t[0, 63]
print("CTA(0,0,0) warp 4: entry->setup %d | setup->first step %d | steps %d | o_done wait %d | epilogue %d | exit sync %d | total %d" % (
    c[1]-c[0], base-c[1], c[2]-base, c[3]-c[2], c[4]-c[3], c[5]-c[4], c[5]-c[0]))
print("softmax warp 4: step | decide(+o_done)  prefetch(s_full)  exp_lo  ld_wait+mask  exp_hi+max  fence+arrive | step total")
for j in range(n):
    r = t[0, j]
    nxt = t[0, j + 1, 0] if j + 1 < n else r[6]
    print(f"{j:3d} | " + " ".join(f"{r[i+1]-r[i]:6d}" for i in range(6)) + f" | {nxt - r[0]:6d}   t0={r[0]-base}")
print("issue times relative to the softmax warp's step-0 start (slot 1 softmax step start | S1 PV1 | S0 PV0 | K load  V load):")
for j in range(n):
    r = t[1, j] - base
    print(f"{j:3d} | step start {t[0, j, 0]-base:7d} | S1 {r[0]:7d} PV1 {r[1]:7d} | S0 {r[2]:7d} PV0 {r[3]:7d} | K {r[4]:7d} V {r[5]:7d}")
