"""CPU: the drop-in surface (fengshen-lm_b200/compat) — import paths, argparse flags, sampler index arithmetic,
optimizer/scheduler construction and strategy config discovery behave like the reference (cited per test).
No CUDA is touched: the model itself is exercised by the -m gpu tests."""
import argparse
import json
import os
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
COMPAT = os.path.join(ROOT, "fengshen-lm_b200", "compat")
if COMPAT not in sys.path:
    sys.path.insert(0, COMPAT)

import pytorch_lightning as pl  # noqa: E402
from fengshen.data.universal_datamodule import PretrainingRandomSampler, PretrainingSampler, UniversalDataModule  # noqa: E402
from fengshen.models import model_utils  # noqa: E402
from fengshen.strategies.megatron_deepspeed import DeepSpeedStrategy  # noqa: E402
from fengshen.utils.universal_checkpoint import UniversalCheckpoint  # noqa: E402


def _parser():
    p = argparse.ArgumentParser()
    p = model_utils.add_module_args(p)
    p = pl.Trainer.add_argparse_args(p)
    p = UniversalDataModule.add_data_specific_args(p)
    p = UniversalCheckpoint.add_argparse_args(p)
    return p


def test_argparse_surface_matches_reference_flags_and_defaults():
    a = _parser().parse_args([])
    # fengshen/models/model_utils.py:13-28
    assert (a.learning_rate, a.min_learning_rate, a.warmup_ratio, a.weight_decay) == (5e-5, 1e-7, 0.1, 0.1)
    assert (a.adam_beta1, a.adam_beta2, a.adam_epsilon, a.scheduler_type) == (0.9, 0.999, 1e-8, "polynomial")
    # universal_datamodule.py:22-45, universal_checkpoint.py:7-21
    assert (a.train_batchsize, a.sampler_type, a.use_mpu) == (16, "random", False)
    assert (a.monitor, a.save_ckpt_path, a.save_top_k) == ("step", "./ckpt/", 10)
    assert a.max_steps == -1 and a.accumulate_grad_batches == 1


def test_pretraining_sampler_partitions_global_batches_by_rank():
    # universal_sampler.py:22-68
    got = [list(PretrainingSampler(20, 4, 2, r, 2)) for r in (0, 1)]
    assert got[0] == [[4, 5], [8, 9], [12, 13], [16, 17]]
    assert got[1] == [[6, 7], [10, 11], [14, 15], [18, 19]]
    assert len(PretrainingSampler(20, 0, 2, 0, 2)) == 5


def test_pretraining_random_sampler_buckets_and_resume():
    # universal_sampler.py:71-125: rank r owns bucket r; a resumed sampler continues inside the same permutation
    full = [list(PretrainingRandomSampler(40, 0, 2, r, 2, epoch=3)) for r in (0, 1)]
    flat0 = [i for b in full[0] for i in b]
    flat1 = [i for b in full[1] for i in b]
    assert sorted(flat0) == list(range(0, 20)) and sorted(flat1) == list(range(20, 40))
    resumed = list(PretrainingRandomSampler(40, 8, 2, 0, 2, epoch=3))   # 8 consumed samples = 2 global batches
    assert resumed == full[0][2:]
    other_epoch = [i for b in PretrainingRandomSampler(40, 0, 2, 0, 2, epoch=4) for i in b]
    assert other_epoch != flat0 and sorted(other_epoch) == sorted(flat0)


class _Lit(pl.LightningModule):
    def __init__(self, args):
        super().__init__()
        self.save_hyperparameters(args)
        self.layer = torch.nn.Linear(4, 4)
        self.input_layernorm = torch.nn.Module()
        self.input_layernorm.scale = torch.nn.Parameter(torch.ones(4))
        self.total_steps = 100


def test_configure_optimizers_groups_by_name_and_builds_polynomial_schedule():
    a = _parser().parse_args(["--learning_rate", "1e-3", "--warmup_steps", "10", "--adam_beta2", "0.95"])
    m = _Lit(a)
    m.trainer = pl.Trainer(strategy=DeepSpeedStrategy(pipe_model_parallel_size=1, tensor_model_parallel_size=1, mpu_seed=42))
    (opt,), (sch,) = model_utils.configure_optimizers(m)
    assert type(opt).__name__ == "FusedAdam"                       # model_utils.py:69-72
    assert sch["interval"] == "step" and sch["frequency"] == 1     # :97
    decay, no_decay = opt.param_groups
    assert decay["weight_decay"] == 0.1 and no_decay["weight_decay"] == 0.0
    assert len(decay["params"]) == 1 and len(no_decay["params"]) == 2   # weight | bias + layernorm.scale (:39-47)
    assert opt.param_groups[0]["betas"] == (0.9, 0.95)
    from fsb200.schedules import polynomial_lr
    for step in range(0, 120, 7):
        assert abs(opt.param_groups[0]["lr"] - polynomial_lr(step, 1e-3, 10, 100, 1e-7)) < 1e-12
        for _ in range(7):
            opt.step(); sch["scheduler"].step()


def test_ddp_strategy_selects_adamw():
    a = _parser().parse_args([])
    m = _Lit(a)
    m.trainer = pl.Trainer(strategy="ddp")
    (opt,), _ = model_utils.configure_optimizers(m)
    assert isinstance(opt, torch.optim.AdamW)                      # model_utils.py:80-83


def test_get_total_steps_integer_arithmetic():
    # model_utils.py:194-209
    a = _parser().parse_args(["--max_epochs", "3", "--train_batchsize", "4", "--accumulate_grad_batches", "2"])
    tr = pl.Trainer.from_argparse_args(a)
    tr.world_size = 2
    tr._train_loader = torch.utils.data.DataLoader(list(range(101)), batch_size=4)
    assert model_utils.get_total_steps(tr, a) == (101 * 3 // 8) // 2
    a2 = _parser().parse_args(["--max_epochs", "-1", "--max_steps", "77"])
    tr2 = pl.Trainer.from_argparse_args(a2)
    tr2._train_loader = tr._train_loader
    assert model_utils.get_total_steps(tr2, a2) == 77


def test_strategy_reads_deepspeed_json_from_env(tmp_path, monkeypatch):
    cfg = {"zero_optimization": {"stage": 2, "reduce_bucket_size": 2e8}, "gradient_clipping": 1,
           "bf16": {"enabled": True}, "activation_checkpointing": {"partition_activations": False}}
    path = tmp_path / "ds.json"
    path.write_text(json.dumps(cfg))
    monkeypatch.setenv("PL_DEEPSPEED_CONFIG_PATH", str(path))          # megatron_deepspeed.py:53
    s = DeepSpeedStrategy(tensor_model_parallel_size=1, pipe_model_parallel_size=1, mpu_seed=42)
    assert s.stage == 2 and s.gradient_clipping == 1.0 and s.precision == "bf16"
    assert "offload_optimizer" not in s.config["zero_optimization"]
    assert DeepSpeedStrategy(tensor_model_parallel_size=8, pipe_model_parallel_size=1, mpu_seed=42).tensor_model_parallel_size == 8
    with pytest.raises(NotImplementedError):   # pipeline parallelism stays outside the hot path
        DeepSpeedStrategy(tensor_model_parallel_size=1, pipe_model_parallel_size=2, mpu_seed=42)
    monkeypatch.setenv("PL_DEEPSPEED_CONFIG_PATH", str(tmp_path / "missing.json"))
    with pytest.raises(FileNotFoundError):
        DeepSpeedStrategy(tensor_model_parallel_size=1, pipe_model_parallel_size=1, mpu_seed=42)


def test_strategy_registry_strings():
    from pytorch_lightning.strategies import strategy_from_string
    assert strategy_from_string("deepspeed_stage_1").stage == 1
    assert strategy_from_string("deepspeed_stage_2").stage == 2
    assert strategy_from_string("ddp").stage == 1
    with pytest.raises(NotImplementedError):
        strategy_from_string("deepspeed_stage_3")
    with pytest.raises(NotImplementedError):
        strategy_from_string("deepspeed_stage_2_offload")


def test_universal_checkpoint_drops_missing_resume_path(tmp_path):
    a = _parser().parse_args(["--load_ckpt_path", str(tmp_path / "nope"), "--save_ckpt_path", str(tmp_path)])
    cb = UniversalCheckpoint(a)
    assert a.load_ckpt_path is None and cb.dirpath == str(tmp_path)   # universal_checkpoint.py:37-41


def test_example_script_parses_on_the_reference_surface():
    sys.path.insert(0, os.path.join(ROOT, "examples"))
    import importlib
    mod = importlib.import_module("pretrain_ziya_llama")
    assert hasattr(mod, "Llama") and hasattr(mod.Llama, "training_step")


def test_llama_hf_fs_converters_roundtrip_and_match_transformers():
    """fengshen/utils/llama_convert/{hf_to_fs,fs_to_hf}.py: the per-head interleaved QKV layout. A transformers LLaMA and the
    pinned oracle (the reference's own arithmetic) give the same loss on weights that went through the converter, and the
    round trip is the identity."""
    import sys as _sys
    _sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import llama_oracle as O
    from fengshen.utils.llama_convert import fs_to_hf_state_dict, hf_to_fs_state_dict
    from transformers import LlamaConfig as HFLlamaConfig, LlamaForCausalLM as HFLlama
    V, h, L, nh, S = 256, 128, 2, 4, 24
    fs = O.make_weights(V, h, L, seed=3, bf16_exact=False)
    hf_sd = fs_to_hf_state_dict(fs, nh)
    back = hf_to_fs_state_dict(hf_sd, nh)
    assert set(back) == set(fs) and all(torch.equal(back[k], fs[k]) for k in fs)
    ff = fs["llama.layers.0.mlp.w1.weight"].shape[0]
    hf = HFLlama(HFLlamaConfig(vocab_size=V, hidden_size=h, intermediate_size=ff, num_hidden_layers=L, num_attention_heads=nh,
                               num_key_value_heads=nh, rms_norm_eps=1e-6, max_position_embeddings=2048,
                               attn_implementation="eager", tie_word_embeddings=False))
    missing, unexpected = hf.load_state_dict(hf_sd, strict=False)
    assert not unexpected and all("rotary" in m or "inv_freq" in m for m in missing), (missing, unexpected)
    batch = O.make_batch(V, 2, S, seed=4)
    want, _ = O.forward(fs, batch, nh)
    got = hf(input_ids=batch["input_ids"], labels=batch["labels"]).loss
    assert abs(got.item() - want.item()) < 2e-5, (got.item(), want.item())


def test_tp_shard_layout_matches_reference_and_sharded_math_equals_full():
    """convert_fs_llama_tp.py:143-181 shard format + the column / row-parallel algebra of mpu/layers.py: running every shard's
    slice of the layer and summing the row-parallel partial outputs reproduces the un-sharded oracle layer exactly (fp32)."""
    import sys as _sys
    _sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import math
    import llama_oracle as O
    import torch.nn.functional as F
    from fengshen.utils.llama_convert import merge_state_dict_tp, split_state_dict_tp
    V, h, L, nh, S, tp = 64, 128, 1, 4, 12, 2
    hn = h // nh
    sd = O.make_weights(V, h, L, seed=1, bf16_exact=False)
    shards = split_state_dict_tp(sd, tp, nh)
    assert shards[0]["llama.layers.0.attention.query_key_value.weight"].shape == (3 * h // tp, h)
    assert shards[1]["llama.layers.0.attention.dense.weight"].shape == (h, h // tp)
    assert shards[0]["llama.layers.0.mlp.w2.weight"].shape[1] * tp == sd["llama.layers.0.mlp.w2.weight"].shape[1]
    assert shards[1]["llama.embed_in.word_embeddings.weight"].shape == (V // tp, h)
    merged = merge_state_dict_tp(shards, nh)
    assert all(torch.equal(merged[k], sd[k]) for k in sd)
    # sharded attention + MLP == full (the all-reduce is the sum over shards)
    batch = O.make_batch(V, 2, S, seed=2)
    x = F.embedding(batch["input_ids"], sd["llama.embed_in.word_embeddings.weight"])
    cos, sin = O.rope_tables(hn, 64)
    p = "llama.layers.0."
    xin = O.rmsnorm(x, sd[p + "input_layernorm.scale"], 1e-6)
    full = O.attention(xin, sd[p + "attention.query_key_value.weight"], sd[p + "attention.dense.weight"], batch["position_ids"], nh,
                       cos, sin)
    part = sum(O.attention(xin, s_[p + "attention.query_key_value.weight"], s_[p + "attention.dense.weight"], batch["position_ids"],
                           nh // tp, cos, sin, hn=hn) for s_ in shards)
    assert torch.allclose(part, full, atol=1e-5)
    fm = O.mlp(xin, sd[p + "mlp.w1.weight"], sd[p + "mlp.w3.weight"], sd[p + "mlp.w2.weight"])
    pm = sum(O.mlp(xin, s_[p + "mlp.w1.weight"], s_[p + "mlp.w3.weight"], s_[p + "mlp.w2.weight"]) for s_ in shards)
    assert torch.allclose(pm, fm, atol=1e-5)


def _mpu_worker(rank, world, port, q):
    import torch
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from fengshen.models.megatron import mpu
    mpu.initialize_model_parallel(2)
    t = torch.tensor([float(rank)])
    dist.all_reduce(t, group=mpu.get_model_parallel_group())
    d = torch.tensor([float(rank)])
    dist.all_reduce(d, group=mpu.get_data_parallel_group())
    q.put((rank, mpu.get_model_parallel_world_size(), mpu.get_model_parallel_rank(), mpu.get_data_parallel_world_size(),
           mpu.get_data_parallel_rank(), float(t), float(d)))
    dist.barrier()
    dist.destroy_process_group()


def test_mpu_groups_tensor_parallel_ranks_consecutive_data_parallel_strided():
    """mpu/initialize.py:37-118 on 4 gloo ranks with tensor parallelism 2: TP groups {0,1} {2,3}, DP groups {0,2} {1,3}."""
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_mpu_worker, args=(r, 4, 29671, q)) for r in range(4)]
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
    for rank in range(4):
        tp_world, tp_rank, dp_world, dp_rank, tp_sum, dp_sum = got[rank]
        assert (tp_world, tp_rank, dp_world, dp_rank) == (2, rank % 2, 2, rank // 2)
        assert tp_sum == {0: 1.0, 1: 1.0, 2: 5.0, 3: 5.0}[rank]              # 0+1 / 2+3: consecutive ranks share a TP group
        assert dp_sum == {0: 2.0, 1: 4.0, 2: 2.0, 3: 4.0}[rank]              # 0+2 / 1+3: same TP rank, strided
