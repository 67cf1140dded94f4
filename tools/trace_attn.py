"""Dev tool: build a traced copy of libfsb200 (-DFSB_ATTN_TRACE), run the dq kernel once, print per-step cycle deltas."""
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
dout = torch.randn(B, S, H, D, device="cuda", dtype=torch.bfloat16)
dqkv = torch.empty_like(qkv)
for _ in range(2):
    o, lse = ops.sdpa_fwd(q, k, v, 1 / math.sqrt(D), True)
    ops.sdpa_bwd(q, k, v, o, dout, lse, 1 / math.sqrt(D), True, dqkv[:, :, 0], dqkv[:, :, 1], dqkv[:, :, 2])
torch.cuda.synchronize()
buf = (ctypes.c_longlong * (2 * 64 * 8))()
fn = lib.load().fsb_debug_attn_trace
fn.argtypes = [ctypes.c_void_p]
fn.restype = ctypes.c_int
assert fn(buf) == 0
t = np.array(buf, dtype=np.int64).reshape(2, 64, 8)
n = S // 64
base = t[0, 0, 0]
print("math warp0: step | wait_s  tmem_ld  compute+st  fence  arrive | step total")
for j in range(n):
    r = t[0, j]
    nxt = t[0, j + 1, 0] if j + 1 < n else r[5]
    print(f"{j:3d} | {r[1]-r[0]:6d} {r[2]-r[1]:6d} {r[3]-r[2]:6d} {r[4]-r[3]:6d} {r[5]-r[4]:6d} | {nxt - r[0]:6d}   t0={r[0]-base}")
print("mma thread: step | wait_kv  issue_S  wait_t_ready  issue_dQ | total")
for j in range(n):
    r = t[1, j]
    print(f"{j:3d} | {r[1]-r[0]:6d} {r[2]-r[1]:6d} {r[3]-r[2]:6d} {r[4]-r[3]:6d} | {r[4]-r[0]:6d}   t0={r[0]-base}")
