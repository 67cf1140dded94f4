#!/usr/bin/env python
"""bench.py — tokens/s of the data-parallel pretraining step (forward + backward + sharded AdamW) on synthetic tokens.

    python bench.py --gpus N --steps K --warmup W [--workload NAME] [--impl reference]

Workloads (BASELINE.json `configs`):
    gpt2-110m        C2  Wenzhong-GPT2-110M, seq 1024, 32 sequences per GPU, bf16            (default at any N; weak scaling)
    ziya-llama-13b   C4  Ziya-LLaMA-13B, seq 2048, 32 sequences per GPU, ZeRO-2, bf16        (needs >= 4 GPUs of HBM)
    ziya-llama-13b-L{n}  same width, n layers (NOT a BASELINE config; kernel bring-up / profiling only)
A "step" is one optimizer step over the per-GPU batch (micro-batches x gradient accumulation). `value` is whole-job
tokens/s with the token tensors already resident in HBM; `e2e` repeats the measurement through the public step API
(fsb200.trainer.PretrainStep.step) from pinned host memory, with the host->device copies and the loss read-back inside
the timed region. Timing: CUDA events, barrier + synchronize on both sides, max over ranks.
`--impl reference` times the reference's own CPU implementation of the same step on the host cores.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(ROOT, "fengshen-lm_b200"))

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

WORKLOADS = {
    # name: family, dims, seq, per-GPU sequences, micro-batch
    "gpt2-110m": dict(family="gpt2", vocab_size=50264, n_positions=1024, n_embd=768, n_layer=12, n_head=12, seq=1024,
                      per_gpu=32, micro=32, lr=1e-4, betas=(0.9, 0.999), wd=0.1, clip=0.0,
                      label="Wenzhong-GPT2-110M pretrain, seq 1024, batch 32/GPU (BASELINE configs[1])"),
    "bert-base": dict(family="bert", variant="bert", vocab_size=21128, hidden_size=768, num_hidden_layers=12,
                      num_attention_heads=12, intermediate_size=3072, hidden_act="gelu", seq=128, per_gpu=8, micro=8,
                      lr=1e-4, betas=(0.9, 0.999), wd=0.1, clip=0.0,
                      label="Erlangshen-BERT-base MLM, seq 128, batch 8 (BASELINE configs[0])"),
    "megatronbert-1.3b": dict(family="bert", variant="megatron", vocab_size=21128, hidden_size=2048,
                              num_hidden_layers=24, num_attention_heads=32, intermediate_size=8192, hidden_act="gelu",
                              seq=512, per_gpu=128, micro=32, lr=1e-4, betas=(0.9, 0.999), wd=0.1, clip=1.0,
                              label="Erlangshen-MegatronBERT-1.3B MLM+SOP pretrain, seq 512, batch 128/GPU, ZeRO-1 "
                                    "(BASELINE configs[2])"),
    "ziya-llama-13b": dict(family="llama", vocab_size=39424, hidden_size=5120, num_hidden_layers=40,
                           num_attention_heads=40, seq=2048, per_gpu=32, micro=4, lr=1e-4, betas=(0.9, 0.95), wd=0.1,
                           clip=1.0, label="Ziya-LLaMA-13B pretrain, seq 2048, global batch 32/GPU, ZeRO-2 (BASELINE configs[3])"),
}


def workload(name):
    if name in WORKLOADS:
        return dict(WORKLOADS[name])
    if name.startswith("ziya-llama-13b-L"):
        w = dict(WORKLOADS["ziya-llama-13b"])
        w["num_hidden_layers"] = int(name.split("-L")[1])
        w["label"] = f"Ziya-LLaMA-13B WIDTH with {w['num_hidden_layers']} layers (bring-up only, not a BASELINE config)"
        return w
    raise SystemExit(f"unknown workload {name}")


def flops_per_token(w):
    """F_tok = 6*N_mm + 3*F_attn_fwd (causal-counted), SURVEY.md §8(d) / BASELINE.md §3."""
    if w["family"] == "gpt2":
        h, L, V, s = w["n_embd"], w["n_layer"], w["vocab_size"], w["seq"]
        n_mm = L * 12 * h * h + V * h
        attn = 4 * s * h * L / 2
    elif w["family"] == "bert":
        h, L, V, s, ff = w["hidden_size"], w["num_hidden_layers"], w["vocab_size"], w["seq"], w["intermediate_size"]
        n_mm = L * (4 * h * h + 2 * h * ff) + h * h + V * h      # + MLM transform dense + tied decoder
        attn = 4 * s * h * L                                      # bidirectional: full s x s
    else:
        h, L, V, s = w["hidden_size"], w["num_hidden_layers"], w["vocab_size"], w["seq"]
        ff = 256 * ((int(2 * h * 4 / 3) + 255) // 256)
        n_mm = L * (4 * h * h + 3 * h * ff) + V * h
        attn = 4 * s * h * L / 2
    return 6.0 * n_mm + 3.0 * attn


# DRAM traffic of the dominant kernel's largest-share launch, from the committed `ncu --set full` capture
# (profiles/r01_ncu_gemm_v3.summary.txt; dram__bytes_read.sum + dram__bytes_write.sum, one launch), next to its algorithmic
# bytes (A + B read once, D written once). bench.py cannot run ncu itself, so these are constants tied to that capture.
NCU_TRAFFIC = {
    "gpt2": {"launch": "fsb::gemm_bf16_kernel<NN,256,pair> 32768x3072x768 (c_fc forward)", "traffic": 60.866304e6 + 146.682112e6,
             "algorithmic": (32768 * 768 + 768 * 3072 + 32768 * 3072) * 2.0, "tensor_pipe_active_pct": 78.4,
             "source": "profiles/r01_ncu_gemm_final.summary.txt"},
    "bert": {"launch": "fsb::gemm_bf16_kernel<NN,256,pair> 32768x3072x768 (same MLP shape family)", "traffic": 60.866304e6 + 146.682112e6,
             "algorithmic": (32768 * 768 + 768 * 3072 + 32768 * 3072) * 2.0, "tensor_pipe_active_pct": 78.4,
             "source": "profiles/r01_ncu_gemm_final.summary.txt"},
    "llama": {"launch": "fsb::gemm_bf16_kernel<NT,256,pair> 8192x15360x5120 (QKV forward)", "traffic": 711.888384e6 + 240.509952e6,
              "algorithmic": (8192 * 5120 + 15360 * 5120 + 8192 * 15360) * 2.0, "tensor_pipe_active_pct": 98.4,
              "source": "profiles/r01_ncu_gemm_final.summary.txt"},
}


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return {"bf16_tflops": d.get("bf16_tflops_sustained", d.get("bf16_tflops")), "hbm_gbs": d.get("hbm_gbs"),
                "src": "measured (MEASURED_PEAKS.json, sustained)"}
    return {"bf16_tflops": 1400.0, "hbm_gbs": 6650.0, "src": "fallback (B200_PROFILING.md)"}


class ClockSampler:
    """nvidia-smi clocks/throttle sampling DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.p = None
        try:
            self.p = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms",
                                       "100", "-i", str(gpu_index)], stdout=self.f, stderr=subprocess.DEVNULL)
        except Exception:
            self.p = None

    def stop(self):
        if self.p is None:
            return None
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        self.f.flush()
        rows = [r.strip().split(",") for r in open(self.f.name) if r.strip()]
        os.unlink(self.f.name)
        sm, mx, reasons = [], [], set()
        for r in rows:
            try:
                sm.append(float(r[1])); mx.append(float(r[2]))
            except Exception:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[4:8]):
                if v.strip().lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return None
        return {"sm_mhz": statistics.median(sm), "sm_max_mhz": max(mx), "reasons": sorted(reasons), "samples": len(sm)}


def make_host_batches(w, n_pool, rank):
    """Per-rank synthetic token batches in pinned host memory: g = manual_seed(1234 + dp_rank)."""
    g = torch.Generator().manual_seed(1234 + rank)
    out = []
    for _ in range(n_pool):
        ids = torch.randint(1, w["vocab_size"] - 8, (w["micro"], w["seq"]), generator=g, dtype=torch.int64)
        if w["family"] == "bert":   # MLM: labels = -100 except a Bernoulli(0.15) subset; C3 adds NSP labels
            sel = torch.rand(ids.shape, generator=g) < 0.15
            b = {"input_ids": ids.pin_memory(), "labels": torch.where(sel, ids, torch.full_like(ids, -100)).pin_memory(),
                 "token_type_ids": torch.zeros_like(ids).pin_memory()}
            if w["variant"] == "megatron":
                b["next_sentence_label"] = torch.randint(0, 2, (w["micro"],), generator=g, dtype=torch.int64).pin_memory()
        else:
            b = {"input_ids": ids.pin_memory(), "labels": ids.clone().pin_memory()}
        out.append(b)
    return out


def build_model(w, device, world):
    from types import SimpleNamespace
    if w["family"] == "gpt2":
        from fsb200.models.gpt2 import GPT2LMHeadModel
        cfg = SimpleNamespace(vocab_size=w["vocab_size"], n_positions=w["n_positions"], n_embd=w["n_embd"],
                              n_layer=w["n_layer"], n_head=w["n_head"], layer_norm_epsilon=1e-5, initializer_range=0.02,
                              resid_pdrop=0.0, embd_pdrop=0.0, attn_pdrop=0.0, activation_function="gelu_new")
        return GPT2LMHeadModel(cfg, device=device, world_size=world)
    if w["family"] == "bert":
        from fsb200.models.bert import BertForMaskedLM, MegatronBertForPreTraining
        cfg = SimpleNamespace(vocab_size=w["vocab_size"], hidden_size=w["hidden_size"],
                              num_hidden_layers=w["num_hidden_layers"], num_attention_heads=w["num_attention_heads"],
                              intermediate_size=w["intermediate_size"], hidden_act=w["hidden_act"],
                              max_position_embeddings=512, type_vocab_size=2, layer_norm_eps=1e-12,
                              hidden_dropout_prob=0.0, attention_probs_dropout_prob=0.0, initializer_range=0.02)
        cls = MegatronBertForPreTraining if w["variant"] == "megatron" else BertForMaskedLM
        return cls(cfg, device=device, world_size=world)
    from fsb200.models.llama import LlamaForCausalLM
    cfg = SimpleNamespace(vocab_size=w["vocab_size"], hidden_size=w["hidden_size"],
                          num_hidden_layers=w["num_hidden_layers"], num_attention_heads=w["num_attention_heads"],
                          rms_norm_epsilon=1e-6, max_position_embeddings=2048, rotary_emb_base=10000,
                          llama_mlp_multiple_of=256)
    return LlamaForCausalLM(cfg, device=device, world_size=world)


def timed(fn, steps, world):
    """barrier + sync, K steps between CUDA events, sync + barrier; returns max-over-ranks milliseconds."""
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(steps):
        fn(i)
    e1.record()
    torch.cuda.synchronize()
    ms = torch.tensor([e0.elapsed_time(e1)], device="cuda")
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        dist.barrier()
    return float(ms.item())


def cpu_reference_tokens_per_s(w, budget_s=20.0):
    """The reference's own CPU path for this workload on the host cores (bounded sample)."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    # all the host threads the process may use — but containers often report more CPUs than they can schedule, and an
    # oversubscribed OpenMP pool is many times slower: calibrate on a GEMM and keep the fastest thread count.
    avail = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    best_n, best_t = avail, float("inf")
    a = torch.randn(1536, 1536)
    for n in sorted({avail, max(1, avail // 2), max(1, avail // 4), min(avail, 32), min(avail, 16), min(avail, 8)}):
        torch.set_num_threads(n)
        (a @ a)
        t0 = time.time()
        for _ in range(3):
            (a @ a)
        dt = time.time() - t0
        if dt < best_t:
            best_n, best_t = n, dt
    torch.set_num_threads(best_n)
    if w["family"] == "gpt2":
        import hf_oracle as H
        cfg = dict(vocab_size=w["vocab_size"], n_positions=w["n_positions"], n_embd=w["n_embd"], n_layer=w["n_layer"],
                   n_head=w["n_head"])
        model = H.build_gpt2(cfg, bf16_exact=False)
        opt = torch.optim.AdamW(H.wenzhong_param_groups(model.named_parameters(), w["wd"]), lr=w["lr"])
        B, S = 1, w["seq"]
        batch = H.make_lm_batch(w["vocab_size"], B, S)

        def one():
            loss = model(input_ids=batch["input_ids"], labels=batch["labels"]).loss
            opt.zero_grad(set_to_none=True)
            loss.backward()
            opt.step()
        kind, sample, scale = "reference", f"transformers GPT2LMHeadModel fp32 + torch AdamW, batch {B} x seq {S}", 1.0
    elif w["family"] == "bert":
        import hf_oracle as H
        cfg = dict(vocab_size=w["vocab_size"], hidden_size=w["hidden_size"], num_hidden_layers=w["num_hidden_layers"],
                   num_attention_heads=w["num_attention_heads"], intermediate_size=w["intermediate_size"],
                   max_position_embeddings=512, type_vocab_size=2)
        meg = w["variant"] == "megatron"
        model = (H.build_megatron_bert if meg else H.build_bert)(cfg, bf16_exact=False)
        opt = torch.optim.AdamW(H.wenzhong_param_groups(model.named_parameters(), w["wd"]), lr=w["lr"])
        B, S = (1 if meg else 8), w["seq"]
        batch = H.make_mlm_batch(w["vocab_size"], B, S, nsp=meg)

        def one():
            loss = model(**batch).loss
            opt.zero_grad(set_to_none=True)
            loss.backward()
            opt.step()
        kind, scale = "reference", 1.0
        sample = f"transformers {type(model).__name__} fp32 + torch AdamW, batch {B} x seq {S}"
    else:
        import llama_oracle as O
        Lr = 1  # one full-width layer + head, extrapolated linearly in L (full size does not fit host RAM)
        V, h, nh = w["vocab_size"], w["hidden_size"], w["num_attention_heads"]
        sd = {k: torch.nn.Parameter(v) for k, v in O.make_weights(V, h, Lr, bf16_exact=False).items()}
        B, S = 1, w["seq"]
        batch = O.make_batch(V, B, S)
        opt = torch.optim.AdamW(O.param_groups(sd.items(), w["wd"]), lr=w["lr"], betas=w["betas"])

        def one():
            loss, _ = O.forward(sd, batch, nh)
            opt.zero_grad(set_to_none=True)
            loss.backward()
            opt.step()
        kind = "port"
        sample = (f"oracle/llama_oracle.py fp32, {Lr} of {w['num_hidden_layers']} layers at full width + LM head, "
                  f"batch {B} x seq {S}; per-layer time extrapolated x{w['num_hidden_layers']}")
        scale = None
    one()  # warm-up
    t0 = time.time(); n = 0
    while True:
        one(); n += 1
        if time.time() - t0 > budget_s or n >= 8:
            break
    dt = (time.time() - t0) / n
    if scale is None:
        # time(L layers) ~ head + L * layer: measure the head-only cost by difference is too slow; report the
        # conservative linear extrapolation of the whole 1-layer step (over-estimates CPU speed slightly)
        dt = dt * w["num_hidden_layers"]
    return B * S / dt, kind, sample, best_n


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--workload", default="gpt2-110m")
    ap.add_argument("--impl", default="fsb200", choices=["fsb200", "reference"])
    ap.add_argument("--micro-batch", type=int, default=0)
    ap.add_argument("--per-gpu-batch", type=int, default=0, help="override sequences per GPU per step (profiling only)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--breakdown", action="store_true",
                    help="after the timed runs, profile 2 more steps with CUDA events around every fsb_* call and print the "
                         "per-entry-point time table to stderr (diagnostic; not part of the JSON line)")
    args = ap.parse_args()
    w = workload(args.workload)
    if args.micro_batch:
        w["micro"] = args.micro_batch
    if args.per_gpu_batch:
        w["per_gpu"] = args.per_gpu_batch
        w["label"] += f" [per-GPU batch overridden to {args.per_gpu_batch}: profiling only]"
    w["micro"] = min(w["micro"], w["per_gpu"])
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}: launch one rank per GPU, e.g. python -m torch.distributed.run "
                         f"--nnodes=1 --nproc-per-node {args.gpus} --master-addr 127.0.0.1 --master-port 29500 bench.py --gpus {args.gpus}")
    tokens_per_step_gpu = w["per_gpu"] * w["seq"]
    ftok = flops_per_token(w)
    cfg_common = {"workload": w["label"], "name": args.workload, "seq_len": w["seq"], "per_gpu_batch": w["per_gpu"],
                  "global_batch": w["per_gpu"] * max(1, args.gpus), "micro_batch": w["micro"],
                  "grad_accum": w["per_gpu"] // w["micro"], "dropout": 0.0,
                  "flops_per_token": ftok, "l2": "working set (weights + activations, GBs) exceeds the 126 MB L2; no flush"}

    if args.impl == "reference":
        if rank != 0:
            return
        v, kind, sample, cores = cpu_reference_tokens_per_s(w, budget_s=max(10.0, 4.0 * args.steps))
        line = {"impl": "reference", "metric": "tokens_per_sec", "value": v, "unit": "tokens/s", "n_gpus": args.gpus,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1000.0 * tokens_per_step_gpu / v,
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": cfg_common,
                "cpu_baseline": {"value": v, "unit": "tokens/s", "cores": cores, "kind": kind, "sample": sample},
                "e2e": {"value": v, "unit": "tokens/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        print(json.dumps(line))
        return

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device — the fsb200 path has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    device = torch.device("cuda", local_rank)
    pg = None
    if world > 1:
        dist.init_process_group("nccl", device_id=device)
    from fsb200 import lib as L, ops
    from fsb200.trainer import PretrainStep
    L.load()

    model = build_model(w, device, world)
    ga = w["per_gpu"] // w["micro"]
    total_steps = 1000
    from fsb200.schedules import polynomial_lr
    stepper = PretrainStep(model, lambda s: polynomial_lr(s, w["lr"], 0.01 * total_steps, total_steps, 1e-7), lr=w["lr"],
                           betas=w["betas"], weight_decay=w["wd"], grad_clip=w["clip"], ga_steps=ga, process_group=pg)
    pool = 2
    host = [make_host_batches(w, ga, rank) for _ in range(pool)]
    dev = [[{k: v.to(device) for k, v in b.items()} for b in hb] for hb in host]
    h2d = sum(t.numel() * t.element_size() for b in host[0] for t in b.values())

    losses = []
    for i in range(max(3, args.warmup)):
        losses.append(stepper.step_device(dev[i % pool]))
    torch.cuda.synchronize()

    # ---- device-resident measurement (value) with per-launch GEMM timing for the roofline block
    prof = ops.KernelProfiler()
    ops.set_profiler(prof)
    k0 = L.kernel_launches
    sampler = ClockSampler(local_rank) if rank == 0 else None
    ms = timed(lambda i: losses.append(stepper.step_device(dev[i % pool])), args.steps, world)
    clocks = sampler.stop() if sampler else None
    launches = L.kernel_launches - k0
    ops.set_profiler(None)
    ms_per_step = ms / args.steps
    value = args.gpus * tokens_per_step_gpu / (ms_per_step / 1000.0)
    gsum = prof.summary().get("gemm_bf16_kernel", {"launches": 0, "ms": 0.0, "work": 0.0})

    # ---- end-to-end through the public API: pinned host batches, H2D + loss read-back inside the timed region
    e2e = None
    if not args.no_e2e:
        ms2 = timed(lambda i: stepper.step(host[i % pool]), args.steps, world)
        e2e = {"value": args.gpus * tokens_per_step_gpu / (ms2 / args.steps / 1000.0), "unit": "tokens/s",
               "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 4}

    if args.breakdown and rank == 0:
        bp = ops.KernelProfiler()
        L.call_profiler = bp
        torch.cuda.synchronize()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev0.record()
        for i in range(2):
            stepper.step_device(dev[i % pool])
        ev1.record()
        torch.cuda.synchronize()
        L.call_profiler = None
        tot = ev0.elapsed_time(ev1)
        rows = sorted(bp.summary().items(), key=lambda kv: -kv[1]["ms"])
        acc = sum(v["ms"] for _, v in rows)
        print(f"[breakdown] 2 steps: {tot:.2f} ms wall on the stream; {acc:.2f} ms inside fsb_* calls "
              f"({100 * acc / tot:.1f}%); the rest is torch-native kernels, NCCL and launch gaps", file=sys.stderr)
        for name, v in rows:
            print(f"[breakdown] {v['ms']:9.3f} ms {100 * v['ms'] / tot:6.2f}%  n={v['launches']:5d}  {name}", file=sys.stderr)
    elif args.breakdown:
        for i in range(2):
            stepper.step_device(dev[i % pool])

    final_loss = float(losses[-1].item())
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    pk = peaks()
    gemm_tf = gsum["work"] / (gsum["ms"] / 1000.0) / 1e12 if gsum["ms"] > 0 else 0.0
    step_tf = value * ftok / args.gpus / 1e12
    line = {"metric": "tokens_per_sec", "value": value, "unit": "tokens/s", "n_gpus": args.gpus, "steps": args.steps,
            "warmup": max(3, args.warmup), "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "bf16", "data": "synthetic", "config": cfg_common,
            "roofline": {"bound": "tensor", "kernel": "fsb::gemm_bf16_kernel (tcgen05)", "achieved": gemm_tf,
                         "peak": pk["bf16_tflops"], "unit": "TFLOP/s", "frac": gemm_tf / pk["bf16_tflops"],
                         "traffic": (NCU_TRAFFIC.get(w["family"]) or {}).get("traffic"),
                         "traffic_detail": NCU_TRAFFIC.get(w["family"]),
                         "peak_source": pk["src"], "launches_per_step": gsum["launches"] / args.steps,
                         "kernel_share_of_step": gsum["ms"] / ms if ms > 0 else None,
                         "step_achieved_tflops_per_gpu": step_tf, "step_frac": step_tf / pk["bf16_tflops"]},
            "e2e": e2e, "gpu_launches": launches, "clocks": clocks, "final_loss": final_loss}
    if args.gpus == 1 and not args.no_cpu_baseline:
        v, kind, sample, cores = cpu_reference_tokens_per_s(w, budget_s=15.0)
        line["cpu_baseline"] = {"value": v, "unit": "tokens/s", "cores": cores, "kind": kind, "sample": sample}
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
