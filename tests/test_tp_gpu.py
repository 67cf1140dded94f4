"""GPU (>= 2 devices): tensor parallelism (SURVEY.md §8f rank 1 — the reference's other Ziya recipe, finetune_with_tp.sh).
Two ranks each hold the `part_{rank}` shard (convert_fs_llama_tp.py layout: whole heads, ff columns and vocabulary rows split) of
the same model; column-parallel QKV / w1|w3 / LM head, row-parallel dense / w2 with an NCCL all-reduce, vocabulary-parallel
embedding, gathered logits (mpu/layers.py:62-470, mappings.py:29-192). The sharded step must reproduce the un-sharded one:
same loss, shard gradients == the matching slices of the full gradients, and after clipped AdamW steps the merged parameters
match the single-GPU run (the gradient norm sums over the tensor-parallel group with the replicated norms counted once)."""
import os
import sys
from types import SimpleNamespace

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (os.path.join(ROOT, "fengshen-lm_b200"), os.path.join(ROOT, "fengshen-lm_b200", "compat"), os.path.join(ROOT, "oracle")):
    if p not in sys.path:
        sys.path.insert(0, p)

V, H, NL, NH, S, STEPS = 512, 256, 2, 4, 64, 3


def _cfg():
    return SimpleNamespace(vocab_size=V, hidden_size=H, num_hidden_layers=NL, num_attention_heads=NH, rms_norm_epsilon=1e-6,
                           max_position_embeddings=2048, rotary_emb_base=10000, llama_mlp_multiple_of=256)


def _run(rank, world, port, q):
    import llama_oracle as O
    from fengshen.utils.llama_convert import merge_state_dict_tp, split_state_dict_tp
    from fsb200.engine import ZeroEngine
    from fsb200.models.llama import LlamaForCausalLM
    dev = torch.device("cuda", rank)
    torch.cuda.set_device(dev)
    tp_group = None
    if world > 1:
        os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
        tp_group = dist.new_group(list(range(world)))
    full = O.make_weights(V, H, NL, seed=0)
    model = LlamaForCausalLM(_cfg(), device=dev, world_size=1, tp_group=tp_group)
    model.load_reference_state_dict(split_state_dict_tp(full, world, NH)[rank] if world > 1 else full)
    dp_group = None
    if world > 1:   # data-parallel size 1: each rank is alone in its data-parallel group (every rank creates every group)
        dp_group = [dist.new_group([r]) for r in range(world)][rank]
    eng = ZeroEngine(model, lr=1e-3, betas=(0.9, 0.95), weight_decay=0.1, grad_clip=1.0, process_group=dp_group, tp_group=tp_group)
    losses, g0 = [], None
    for it in range(STEPS):
        ids = O.make_batch(V, 2, S, seed=50 + it)["input_ids"].to(dev)
        out = model(input_ids=ids, labels=ids)
        out.loss.backward()
        if it == 0:
            g0 = {n: p.main_grad.float().cpu().numpy().copy() for n, p in model.named_parameters()}
        eng.backward_done()
        eng.step()
        losses.append(out.loss.item())
    eng.wait_params()
    # numpy arrays (pickled by value): torch tensors travel through shared-memory handles that die with the child
    q.put((world, rank, losses, g0, {n: p.detach().float().cpu().numpy().copy() for n, p in model.named_parameters()},
           float(eng.grad_norm.item())))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_tensor_parallel_2_matches_single_gpu():
    from fengshen.utils.llama_convert import merge_state_dict_tp, split_state_dict_tp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_run, args=(r, 2, 29851, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted((q.get(timeout=300) for _ in range(2)), key=lambda r: r[1])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    p1 = ctx.Process(target=_run, args=(0, 1, 0, q))
    p1.start()
    single = q.get(timeout=300)
    p1.join(timeout=60)
    # identical loss on both tensor-parallel ranks, equal to the single-GPU loss up to bf16 all-reduce order
    assert res[0][2] == res[1][2]
    for a, b in zip(res[0][2], single[2]):
        assert abs(a - b) < 5e-3, (res[0][2], single[2])
    # first-step gradients: every shard gradient is the matching slice of the full gradient
    T = lambda d: {k: torch.from_numpy(v) for k, v in d.items()}
    full_g = T(single[3])
    merged_g = merge_state_dict_tp([T(res[0][3]), T(res[1][3])], NH)
    for name, want in full_g.items():
        got = merged_g[name]
        cos = torch.dot(got.flatten(), want.flatten()) / (got.norm() * want.norm() + 1e-30)
        assert cos.item() > 0.999, (name, cos.item())
    # clipped AdamW: same global gradient norm (replicated norms counted once), merged parameters track the single-GPU run
    assert abs(res[0][5] - single[5]) < 2e-2 * single[5], (res[0][5], single[5])
    merged_p = merge_state_dict_tp([T(res[0][4]), T(res[1][4])], NH)
    for name, want in T(single[4]).items():
        assert (merged_p[name] - want).abs().max().item() < 2e-2, name
