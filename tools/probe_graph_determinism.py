"""Dev tool: is the training step bit-reproducible run to run (eager vs eager), and does the CUDA-graph step match it?"""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for d in ("tests", "fengshen-lm_b200", "oracle", ""):
    sys.path.insert(0, os.path.join(ROOT, d))
import importlib.util
spec = importlib.util.spec_from_file_location("t", os.path.join(ROOT, "tests", "test_llama_gpu.py"))
t = importlib.util.module_from_spec(spec); spec.loader.exec_module(t)
from fsb200.schedules import polynomial_lr
from fsb200.trainer import PretrainStep
g = np.load(t.GOLDEN[0])
ga = int(sys.argv[1]) if len(sys.argv) > 1 else 2
clip = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0


def run(graph):
    model, sd, (V, h, L, nh, B, S) = t._build(g)
    st = PretrainStep(model, lambda s_: polynomial_lr(s_, 1e-3, 2, 20, 1e-7), lr=1e-3, betas=(0.9, 0.95), weight_decay=0.1,
                      grad_clip=clip, ga_steps=ga, cuda_graph=graph)
    snaps = []
    for it in range(6):
        mbs = [{k: v.cuda() for k, v in t.O.make_batch(V, B, S, seed=300 + ga * it + m).items() if k in ("input_ids", "labels")}
               for m in range(ga)]
        loss = float(st.step_device(mbs))
        snaps.append((loss, model.flat.params.clone(), st.engine.master.clone(), st.engine.exp_avg.clone()))
    return snaps


a, b, c = run(False), run(False), run(True)
for name, x, y in (("eager vs eager", a, b), ("eager vs graph", a, c)):
    for i, (sx, sy) in enumerate(zip(x, y)):
        print(name, "step", i, "dloss %.3e" % abs(sx[0] - sy[0]), "params equal", bool(torch.equal(sx[1], sy[1])),
              "max|dmaster| %.3e" % (sx[2] - sy[2]).abs().max().item(), "max|dm| %.3e" % (sx[3] - sy[3]).abs().max().item())
