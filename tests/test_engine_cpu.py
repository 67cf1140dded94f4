"""CPU: host-side logic of the ZeRO engine — bucket / shard planning, and the world_size-2 data path over gloo
(reduce-scatter per bucket, fp32 shard accumulation, clip, sharded AdamW, in-place all-gather) against a single-process
AdamW over the mean gradient (the exact-arithmetic oracle of SURVEY.md Appendix D). The compute kernels are replaced by
the torch-CPU test double in tests/cpu_kernels.py; the product default (CUDA library) is exercised by the -m gpu tests."""
import os
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (os.path.join(ROOT, "fengshen-lm_b200"), os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)

from fsb200.flat import ALIGN, FlatBuffers, FlatSpec, is_no_decay  # noqa: E402


def _spec():
    s = FlatSpec()
    s.add("emb.weight", (50, 24), "emb")
    for i in range(3):
        s.add(f"layers.{i}.input_layernorm.scale", (24,), f"layer{i}")
        s.add(f"layers.{i}.w1.weight", (64, 24), f"layer{i}")
        s.add(f"layers.{i}.w3.weight", (64, 24), f"layer{i}")
        s.add(f"layers.{i}.proj.bias", (24,), f"layer{i}")
    s.add("head.weight", (50, 24), "head")
    return s


def test_no_decay_grouping_follows_reference_substrings():
    # fengshen/models/model_utils.py:39-47
    assert is_no_decay("llama.layers.0.input_layernorm.scale")
    assert is_no_decay("llama.final_layer_norm.scale")
    assert is_no_decay("transformer.h.0.attn.c_attn.bias")
    assert is_no_decay("bert.embeddings.LayerNorm.weight")
    assert not is_no_decay("transformer.h.0.ln_1.weight")      # GPT-2 LN weights ARE decayed
    assert not is_no_decay("llama.layers.0.mlp.w1.weight")


@pytest.mark.parametrize("world", [1, 2, 8])
def test_flat_plan_buckets_and_shards(world):
    fb = FlatBuffers(_spec(), "cpu", world_size=world)
    names = [b[0] for b in fb.buckets]
    assert names == ["emb", "layer0", "layer1", "layer2", "head", "no_decay"]
    end = 0
    for (name, start, length, decay) in fb.buckets:
        assert start == end and length % (world * ALIGN) == 0
        assert decay == (name != "no_decay")
        end = start + length
    assert end == fb.total and fb.shard_numel * world == fb.total
    # every parameter is aligned and lies in the bucket of its weight-decay class
    for name, (off, shape) in fb.offsets.items():
        assert off % ALIGN == 0
        b = next(b for b in fb.buckets if b[1] <= off < b[1] + b[2])
        assert (b[0] == "no_decay") == is_no_decay(name)
    # w1 | w3 adjacency for the fused GEMM operand
    w13 = fb.span("layers.1.w1.weight", 128, 24)
    fb.view("layers.1.w3.weight").fill_(3.0)
    assert float(w13[64:].float().min()) == 3.0 and float(w13[:64].float().abs().max()) == 0.0
    # shard segments tile the local shard exactly
    assert fb.shard_offsets[0] == 0
    for i in range(1, len(fb.buckets)):
        assert fb.shard_offsets[i] == fb.shard_offsets[i - 1] + fb.buckets[i - 1][2] // world


class _ToyModel:
    def __init__(self, world):
        self.flat = FlatBuffers(_spec(), "cpu", world_size=world)
        g = torch.Generator().manual_seed(0)
        for name in self.flat.offsets:
            self.flat.view(name).copy_(torch.randn(self.flat.offsets[name][1], generator=g).to(torch.bfloat16))
        self.grad_hook, self.loss_scale = None, 1.0

    def fake_backward(self, rank, micro):
        """Deterministic per-(rank, micro) gradients, reported bucket by bucket in backward order."""
        g = torch.Generator().manual_seed(1000 + 10 * rank + micro)
        grads = {name: torch.randn(self.flat.offsets[name][1], generator=g) * self.loss_scale for name in self.flat.offsets}
        if getattr(self, "backward_begin_hook", None):
            self.backward_begin_hook()
        # a layer's gradients are written, THEN its bucket is reported — layer buckets may share rotating slots (ZeRO-2)
        for b in ["head", "layer2", "layer1", "layer0", "emb", "no_decay"]:
            i = self.flat.bucket_index[b]
            _, start, length, _ = self.flat.buckets[i]
            for name, (off, _) in self.flat.offsets.items():
                if start <= off < start + length:
                    gv = self.flat.view(name, grad=True)
                    if getattr(self, "accumulate_grads", False):
                        gv.copy_((gv.float() + grads[name].to(torch.bfloat16).float()).to(torch.bfloat16))
                    else:
                        gv.copy_(grads[name].to(torch.bfloat16))
            self.grad_hook(b)


def _reference(world, ga, steps, lr, wd, clip, stage=2):
    """Single process: fp32 master AdamW over the sum of every rank's (already 1/(world*ga)-scaled, bf16) gradients."""
    import cpu_kernels as K
    ref = _ToyModel(1)
    fb = ref.flat
    master = fb.params.float().clone()
    m, v = torch.zeros_like(master), torch.zeros_like(master)
    scale = 1.0 / (world * ga)
    for step in range(1, steps + 1):
        total = torch.zeros_like(master)

        def rank_grads(r, micro):
            g = torch.Generator().manual_seed(1000 + 10 * r + micro)
            tmp = torch.zeros(fb.total, dtype=torch.bfloat16)
            for name, (off, shape) in fb.offsets.items():
                n = 1
                for s in shape:
                    n *= s
                tmp[off:off + n] = (torch.randn(shape, generator=g) * scale).to(torch.bfloat16).flatten()
            return tmp
        if stage == 2:
            for micro in range(ga):
                # bf16 wire reduction of the ranks' bucket gradients, then fp32 accumulation over micro-steps
                wire = torch.zeros(fb.total, dtype=torch.bfloat16)
                for r in range(world):
                    wire = (wire.float() + rank_grads(r, micro).float()).to(torch.bfloat16)
                total += wire.float()
        else:
            # stage 1: every rank accumulates its micro-batches in bf16, ONE bf16 wire reduction at the GA boundary
            wire = torch.zeros(fb.total, dtype=torch.bfloat16)
            for r in range(world):
                acc = torch.zeros(fb.total, dtype=torch.bfloat16)
                for micro in range(ga):
                    acc = (acc.float() + rank_grads(r, micro).float()).to(torch.bfloat16)
                wire = (wire.float() + acc.float()).to(torch.bfloat16)
            total = wire.float()
        coef = None
        if clip > 0:
            nrm = total.pow(2).sum().sqrt()
            coef = torch.clamp(clip / (nrm + 1e-6), max=1.0)
        for (_, start, length, decay) in fb.buckets:
            sl = slice(start, start + length)
            K.adamw_flat(master[sl], m[sl], v[sl], total[sl], None, lr, 0.9, 0.95, 1e-8, wd if decay else 0.0, step, coef)
    return master.to(torch.bfloat16)


def _worker(rank, world, ga, steps, clip, port, q, stage=2):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import cpu_kernels as K
    from fsb200.engine import ZeroEngine
    model = _ToyModel(world)
    eng = ZeroEngine(model, lr=1e-2, betas=(0.9, 0.95), eps=1e-8, weight_decay=0.1, grad_clip=clip, ga_steps=ga,
                     kernels=K, overlap_comm=False, stage=stage)
    if stage == 2:   # per-layer gradient buckets share two rotating slots: no full-size gradient buffer
        assert model.flat.grads.numel() < model.flat.total and eng.grad_bytes_released > 0
    else:
        assert model.flat.grads.numel() == model.flat.total
    assert model.loss_scale == 1.0 / (world * ga)
    for _ in range(steps):
        for micro in range(ga):
            model.fake_backward(rank, micro)
            eng.backward_done()
        eng.step()
    q.put((rank, model.flat.params.clone()))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("ga,clip,stage", [(1, 0.0, 2), (2, 1.0, 2), (2, 1.0, 1)])
def test_zero_engine_world2_gloo_matches_single_process_adamw(ga, clip, stage):
    world, steps = 2, 3
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29600 + ga + 10 * stage
    procs = [ctx.Process(target=_worker, args=(r, world, ga, steps, clip, port, q, stage)) for r in range(world)]
    for p in procs:
        p.start()
    got = dict(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert torch.equal(got[0], got[1]), "ranks disagree on the gathered parameters"
    ref = _reference(world, ga, steps, 1e-2, 0.1, clip, stage)
    # identical arithmetic up to the order of the two-rank bf16 sum -> bit-exact here
    assert torch.equal(got[0], ref), (got[0].float() - ref.float()).abs().max()


def test_engine_rejects_wrong_world_and_micro_count():
    import cpu_kernels as K
    from fsb200.engine import ZeroEngine
    model = _ToyModel(1)
    eng = ZeroEngine(model, ga_steps=2, kernels=K)
    model.fake_backward(0, 0)
    eng.backward_done()
    with pytest.raises(RuntimeError, match="micro-batches"):
        eng.step()
    with pytest.raises(ValueError, match="world_size"):
        ZeroEngine(_ToyModel(2), kernels=K)


def test_compact_grads_folds_layer_buckets_onto_rotating_slots():
    fb = FlatBuffers(_spec(), "cpu", world_size=2)
    v0 = fb.view("layers.0.w1.weight", grad=True)
    v2 = fb.view("layers.2.w1.weight", grad=True)
    v1 = fb.span("layers.1.w1.weight", 128, 24, grad=True)
    head = fb.view("head.weight", grad=True)
    full = fb.grads.numel()
    released = fb.compact_grads(slots=2)
    layer_len = fb.buckets[fb.bucket_index["layer0"]][2]
    assert released == 2 * layer_len and fb.grads.numel() == full - layer_len
    # views handed out before the call were re-pointed in place: layers 0 and 2 alias, layer 1 does not
    v0.fill_(1.0)
    assert float(v2.float().sum()) == v2.numel() and float(v1.float().abs().sum()) == 0.0
    assert float(fb.bucket_view(fb.bucket_index["layer2"], grad=True).float().sum()) == v0.numel()
    head.fill_(2.0)
    assert float(fb.bucket_view(fb.bucket_index["head"], grad=True).float().sum()) == 2.0 * head.numel()
    assert [g for g in fb.rot_group if g] == [("layer", 0), ("layer", 1), ("layer", 0)]
    # parameters keep their full-size layout
    assert fb.params.numel() == fb.total


# ---- tensor parallelism x data parallelism: the clipping norm sums over BOTH groups, replicated buckets counted once ----------
def _tp_grads(tp_rank, dp_rank, name, shape):
    """Shard gradients differ per (tp, dp) rank; the replicated bucket's are the same on both tensor-parallel ranks."""
    seed = 500 + 10 * dp_rank + (0 if name.startswith("norm") else 1 + tp_rank)
    return torch.randn(shape, generator=torch.Generator().manual_seed(seed))


class _TpToy:
    tp_replicated_buckets = ("norms",)

    def __init__(self, world):
        s = FlatSpec()
        s.add("w.weight", (40, 24), "shard")
        s.add("norm.scale", (24,), "norms")
        self.flat = FlatBuffers(s, "cpu", world_size=world)
        self.grad_hook, self.loss_scale = None, 1.0

    def fake_backward(self, tp_rank, dp_rank):
        if getattr(self, "backward_begin_hook", None):
            self.backward_begin_hook()
        for name, bucket in (("w.weight", "shard"), ("norm.scale", "norms")):
            g = _tp_grads(tp_rank, dp_rank, name, self.flat.offsets[name][1]) * self.loss_scale
            self.flat.view(name, grad=True).copy_(g.to(torch.bfloat16))
            self.grad_hook(bucket)


def _tp_worker(rank, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=4)
    sys.path.insert(0, os.path.join(ROOT, "fengshen-lm_b200", "compat"))
    import cpu_kernels as K
    from fengshen.models.megatron import mpu
    from fsb200.engine import ZeroEngine
    mpu.initialize_model_parallel(2)
    tp_rank, dp_rank = mpu.get_model_parallel_rank(), mpu.get_data_parallel_rank()
    model = _TpToy(world=2)
    eng = ZeroEngine(model, lr=1e-2, grad_clip=0.5, kernels=K, overlap_comm=False, process_group=mpu.get_data_parallel_group(),
                     tp_group=mpu.get_model_parallel_group())
    assert eng.world == 2 and eng.tp == 2 and eng.tp_replicated == [True if b[0] == "norms" else False for b in model.flat.buckets]
    model.fake_backward(tp_rank, dp_rank)
    eng.backward_done()
    eng.step()
    q.put((rank, float(eng.grad_norm), float(eng.coef)))
    dist.barrier()
    dist.destroy_process_group()


def test_clipping_norm_under_tensor_parallelism_counts_replicated_buckets_once():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_tp_worker, args=(r, 29683, q)) for r in range(4)]
    for p in procs:
        p.start()
    try:
        got = dict((r[0], r[1:]) for r in (q.get(timeout=180) for _ in range(4)))
    finally:
        for p in procs:
            p.join(timeout=60)
            if p.is_alive():
                p.terminate()
    assert all(p.exitcode == 0 for p in procs)
    # expectation: the data-parallel mean of the bf16 gradients (wire sum of the two ranks' halves), then
    # ||g||^2 = sum over tensor-parallel ranks of the shard part + the replicated part ONCE
    def reduced(tp_rank, name, shape):
        parts = [(_tp_grads(tp_rank, d, name, shape) * 0.5).to(torch.bfloat16) for d in range(2)]
        return (parts[0] + parts[1]).float()
    total = sum(reduced(t, "w.weight", (40, 24)).pow(2).sum() for t in range(2)) + reduced(0, "norm.scale", (24,)).pow(2).sum()
    want = float(total.sqrt())
    for rank in range(4):
        norm, coef = got[rank]
        assert abs(norm - want) < 1e-4 * want, (rank, norm, want)
        assert abs(coef - min(1.0, 0.5 / (want + 1e-6))) < 1e-5
